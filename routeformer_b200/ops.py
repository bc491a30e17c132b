"""Raw (non-autograd) tensor-level wrappers of the C ABI: one function per entry point.

PyTorch supplies device memory and the stream; all arithmetic happens in the CUDA library.  Every wrapper
raises if a tensor is not a CUDA tensor -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import check

F32, F16, BF16, U8, U8_F16 = 0, 1, 2, 3, 4
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
ACT_GELU_SAVE_GRAD = 3  # forward: out = gelu(x), `preact` receives gelu'(x)
DACT_SAVED = 3          # backward: multiply by the saved derivative in `dact_aux`
ATTN_PROB, ATTN_PROB_MASKED, ATTN_FULL = 0, 1, 2
LAYOUT_BLHD, LAYOUT_BHLD = 0, 1
ACT_CODES = {None: ACT_NONE, "none": ACT_NONE, "relu": ACT_RELU, "gelu": ACT_GELU}

_DTYPES = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16, torch.uint8: U8}

# number of kernels enqueued through this module (bench.py reports it as `gpu_launches`)
launch_count = 0


def _count(n: int = 1) -> None:
    global launch_count
    launch_count += n


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("routeformer_b200 ops need CUDA tensors (no CPU fallback exists)")
    return t.data_ptr()


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t


def _row_pitch(t: torch.Tensor, name: str) -> int:
    """Leading dimension of a 2-D view with unit inner stride."""
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError(f"{name} must be 2-D with a contiguous inner dimension, got shape {tuple(t.shape)} strides {t.stride()}")
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


# ---------------------------------------------------------------------------------------------
def fov_crop(frames: torch.Tensor, centers: torch.Tensor, windows: torch.Tensor, out_size: int, mean, std,
             patch: int = 0, frame_ids: Optional[torch.Tensor] = None, n_frames: Optional[int] = None,
             out_dtype=torch.float32, out: Optional[torch.Tensor] = None, u8_as_f16: bool = False) -> torch.Tensor:
    """frames [*,3,H,W] (fp16/fp32/u8, contiguous) -> [n,3,S,S] or patch-major [n*G*G, 3*p*p].
    u8_as_f16: uint8 pixels are converted like the reference's loader, fp16(v / 255), before they are interpolated."""
    lib = _lib.load()
    assert frames.is_contiguous() and frames.shape[-3] == 3
    H, W = frames.shape[-2:]
    n = int(n_frames if n_frames is not None else (frame_ids.numel() if frame_ids is not None else frames.numel() // (3 * H * W)))
    centers = _f32(centers, "centers").contiguous()
    windows = _f32(windows, "windows").contiguous()
    assert centers.shape == (n, 2) and windows.shape == (n, 2)
    if frame_ids is not None:
        assert frame_ids.dtype == torch.int32 and frame_ids.is_contiguous()
    if out is None:
        if patch > 0:
            G = out_size // patch
            out = torch.empty(n * G * G, 3 * patch * patch, device=frames.device, dtype=out_dtype)
        else:
            out = torch.empty(n, 3, out_size, out_size, device=frames.device, dtype=out_dtype)
    p = _lib.RfFovCropParams()
    p.frames, p.src_dtype, p.frame_ids = _ptr(frames), (U8_F16 if (u8_as_f16 and frames.dtype == torch.uint8) else _DTYPES[frames.dtype]), _ptr(frame_ids)
    p.n_frames, p.H, p.W = n, H, W
    p.centers, p.windows = _ptr(centers), _ptr(windows)
    for c in range(3):
        p.mean[c] = float(mean[c])
        p.inv_std[c] = 1.0 / float(std[c])
    p.out_size, p.patch, p.out, p.out_dtype = out_size, patch, _ptr(out), _DTYPES[out.dtype]
    p.out_ld = out.stride(0) if patch > 0 else 0
    check(lib.rf_fov_crop(C.byref(p), _stream()), "rf_fov_crop")
    _count()
    return out


def area_resize_u8(frames: torch.Tensor, factor: Optional[float] = None, out_hw=None, rows=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Dataset-side down-scaling on the device: uint8 [..., H, W] -> uint8 [..., H', W'], bit-exact with the reference loader's
    `cv2.resize(frame, (int(W * factor), int(H * factor)), interpolation=cv2.INTER_AREA)` (io/dataset.py:1440-1501).
    rows = (r0, r1): scale only rows [r0, r1) of every plane (the GoPro row crop of dataset.py:1324-1338, no copy)."""
    lib = _lib.load()
    if frames.dtype != torch.uint8 or not frames.is_contiguous():
        raise TypeError("area_resize_u8: contiguous uint8 frames expected")
    H, W = frames.shape[-2:]
    r0, r1 = (0, H) if rows is None else rows
    Hc = r1 - r0
    if out_hw is None:
        if factor is None or not factor < 1:
            raise ValueError("area_resize_u8: give out_hw or a factor < 1 (the reference takes INTER_LINEAR for factor > 1)")
        out_hw = (int(Hc * factor), int(W * factor))
    dH, dW = out_hw
    n_planes = frames.numel() // (H * W)
    if out is None:
        out = torch.empty(*frames.shape[:-2], dH, dW, device=frames.device, dtype=torch.uint8)
    p = _lib.RfAreaResizeParams()
    p.src, p.src_plane_stride, p.src_row_pitch = _ptr(frames) + r0 * W, H * W, W
    p.n_planes, p.H, p.W = n_planes, Hc, W
    p.dst, p.dH, p.dW = _ptr(out), dH, dW
    check(lib.rf_area_resize_u8(C.byref(p), _stream()), "rf_area_resize_u8")
    _count()
    return out


# ---- precise mode (parity instrument) ------------------------------------------------------------
# RF_PRECISE=1 (or `with ops.precise():`) runs every fp32 GEMM as a 3xTF32 split on the SAME tcgen05 kernel:
#   A = A_hi + A_lo, B = B_hi + B_lo (hi = operand rounded to tf32, lo = the remainder),
#   A.B^T ~= A_hi.B_hi^T + A_lo.B_hi^T + A_hi.B_lo^T        (dropped term A_lo.B_lo^T ~ 2^-22 relative)
# which restores fp32-level accuracy (fp32 accumulate throughout).  ProbSparse top-u selection is a discontinuous function of
# the scores, so the raw (no-replay) comparison against the reference's golden outputs is asserted in this mode, where a
# TF32 rounding cannot flip a marginal query.  3 launches + two elementwise splits per GEMM: a test mode, not a fast path.
import os as _os

PRECISE = _os.environ.get("RF_PRECISE", "0") == "1"


class precise:
    def __init__(self, on: bool = True):
        self.on = on

    def __enter__(self):
        global PRECISE
        self.prev, PRECISE = PRECISE, self.on
        # the tensor-core attention forward rounds P and V to tf32: precise mode takes the fp32 FMA kernels instead
        self.prev_tc = _os.environ.get("RF_ATTN_TC")
        if self.on:
            _os.environ["RF_ATTN_TC"] = "0"
        return self

    def __exit__(self, *exc):
        global PRECISE
        PRECISE = self.prev
        if self.on:
            if self.prev_tc is None:
                _os.environ.pop("RF_ATTN_TC", None)
            else:
                _os.environ["RF_ATTN_TC"] = self.prev_tc


# ---- bf16 operand mode (north star: "bf16 operands with a stated looser tolerance") ---------------------------------
# RF_BF16=1 (or `with ops.bf16_operands():`): a GEMM operand is taken in bf16 wherever it can be PRODUCED in bf16 without an extra
# pass over memory (fp32 accumulate, fp32 epilogue everywhere):
#   * the patch embedding: the crop kernel writes bf16 patches, the projection weight is cast once per weight version
#     (instead of the fp16 operands the reference's autocast uses);
#   * inference (torch.no_grad): the patch embedding stores its features in bf16 and the frame encoder's token convolution
#     (K = 1024, the other tensor-bound GEMM of the model) multiplies them as bf16 against a bf16 copy of its weight.
# Everything downstream (the D = 128 layers, HBM-bound; attention scores and the top-u selection) stays fp32 / TF32: casting
# their activations would cost a pass that the GEMM cannot win back.  Tolerance: tests/test_gpu_model.py::test_bf16_operand_mode
# (measured errors in profiles/).
BF16_MODE = _os.environ.get("RF_BF16", "0") == "1"


class bf16_operands:
    def __init__(self, on: bool = True):
        self.on = on

    def __enter__(self):
        global BF16_MODE
        self.prev, BF16_MODE = BF16_MODE, self.on
        return self

    def __exit__(self, *exc):
        global BF16_MODE
        BF16_MODE = self.prev


def _tf32_split(t: torch.Tensor):
    """(hi, lo): hi = t rounded to 10 mantissa bits (round-half-away in magnitude), lo = t - hi (exact in fp32)."""
    rows, cols = t.shape
    pitch = (cols + 3) // 4 * 4  # TMA needs a 16 B row pitch: keep the padding a column-slice view had
    hi = torch.zeros(rows, pitch, device=t.device, dtype=torch.float32)[:, :cols]
    lo = torch.zeros(rows, pitch, device=t.device, dtype=torch.float32)[:, :cols]
    c = t.contiguous()
    hi.copy_(((c.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32))
    torch.sub(c, hi, out=lo) if pitch == cols else lo.copy_(c - hi)
    return hi, lo


def _gemm_precise(A, B, out, *, a_mn, b_mn, residual, accumulate, colsum_a, **kw):
    a_hi, a_lo = _tf32_split(A)
    b_hi, b_lo = _tf32_split(B)
    if colsum_a is not None:  # the fused column sums would see the tf32-rounded tile: take them from the fp32 operand instead
        colsum_accumulate(A, colsum_a)
    if accumulate:
        for x, y in ((a_lo, b_hi), (a_hi, b_lo), (a_hi, b_hi)):
            _gemm(x, y, out, a_mn=a_mn, b_mn=b_mn, accumulate=True, **kw)
        return out
    Mv = kw.get("M") or (A.shape[1] if a_mn else A.shape[0])
    N = B.shape[1] if b_mn else B.shape[0]
    corr = torch.empty(Mv, N, device=out.device, dtype=torch.float32)
    plain = dict(M=kw.get("M"), split_k=1)
    _gemm(a_lo, b_hi, corr, a_mn=a_mn, b_mn=b_mn, **plain)
    _gemm(a_hi, b_lo, corr, a_mn=a_mn, b_mn=b_mn, residual=corr, **plain)
    if residual is not None:
        corr.add_(residual)
    return _gemm(a_hi, b_hi, out, a_mn=a_mn, b_mn=b_mn, residual=corr, **kw)


def gemm(A: torch.Tensor, B: torch.Tensor, out: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False,
         bias: Optional[torch.Tensor] = None, rowadd: Optional[torch.Tensor] = None, rowadd_period: int = 0,
         residual: Optional[torch.Tensor] = None, act: int = ACT_NONE, preact: Optional[torch.Tensor] = None,
         dact_aux: Optional[torch.Tensor] = None, dact: int = ACT_NONE, accumulate: bool = False, split_k: int = 0,
         out_group=(0, 0, 0), round_f16: bool = False, M: Optional[int] = None, colsum_a: Optional[torch.Tensor] = None) -> torch.Tensor:
    """See `_gemm`; in precise mode fp32 operands take the 3xTF32 route."""
    if PRECISE and A.dtype == torch.float32:
        return _gemm_precise(A, B, out, a_mn=a_mn, b_mn=b_mn, residual=residual, accumulate=accumulate, colsum_a=colsum_a, bias=bias,
                             rowadd=rowadd, rowadd_period=rowadd_period, act=act, preact=preact, dact_aux=dact_aux, dact=dact,
                             split_k=split_k, out_group=out_group, round_f16=round_f16, M=M)
    return _gemm(A, B, out, a_mn=a_mn, b_mn=b_mn, bias=bias, rowadd=rowadd, rowadd_period=rowadd_period, residual=residual, act=act,
                 preact=preact, dact_aux=dact_aux, dact=dact, accumulate=accumulate, split_k=split_k, out_group=out_group,
                 round_f16=round_f16, M=M, colsum_a=colsum_a)


def _gemm(A: torch.Tensor, B: torch.Tensor, out: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False,
         bias: Optional[torch.Tensor] = None, rowadd: Optional[torch.Tensor] = None, rowadd_period: int = 0,
         residual: Optional[torch.Tensor] = None, act: int = ACT_NONE, preact: Optional[torch.Tensor] = None,
         dact_aux: Optional[torch.Tensor] = None, dact: int = ACT_NONE, accumulate: bool = False, split_k: int = 0,
         out_group=(0, 0, 0), round_f16: bool = False, M: Optional[int] = None, colsum_a: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] (+)= epilogue(A . B^T) on the tcgen05 kernel.  A, B are 2-D fp32 views (multiplied as TF32) or both fp16 / both
    bf16 (kind::f16, K-major only); the epilogue operands are fp32; out is fp32, or bf16 for 16-bit operands (rounded once, after
    the epilogue: the next GEMM takes it as a bf16 operand).

    a_mn=False: A is [M,K] memory; True: A is [K,M] memory (logical A^T).  Same for B with N.
    """
    lib = _lib.load()
    f16 = A.dtype in (torch.float16, torch.bfloat16)
    if f16:
        if B.dtype != A.dtype or a_mn or b_mn:
            raise TypeError("gemm: 16-bit operands need A and B in the same 16-bit type, K-major")
    else:
        _f32(A, "A"), _f32(B, "B")
    if out.dtype == torch.bfloat16:
        if not f16 or accumulate or preact is not None:
            raise TypeError("gemm: a bf16 output needs 16-bit operands and no accumulate / preact")
    else:
        _f32(out, "out")
    if a_mn:
        K, Ma = A.shape
    else:
        Ma, K = A.shape
    if b_mn:
        Kb, N = B.shape
    else:
        N, Kb = B.shape
    if K != Kb:
        raise ValueError(f"gemm: reduction mismatch {K} vs {Kb}")
    Mv = Ma if M is None else M
    p = _lib.RfGemmParams()
    p.A, p.lda, p.a_mn_major = _ptr(A), _row_pitch(A, "A"), int(a_mn)
    p.B, p.ldb, p.b_mn_major = _ptr(B), _row_pitch(B, "B"), int(b_mn)
    p.C, p.ldc = _ptr(out), _row_pitch(out, "out")
    p.M, p.N, p.K = Mv, N, K
    p.bias = _ptr(bias)
    if rowadd is not None:
        p.rowadd, p.rowadd_period, p.ld_rowadd = _ptr(rowadd), rowadd_period, _row_pitch(rowadd, "rowadd")
    if residual is not None:
        p.residual, p.ld_res = _ptr(residual), _row_pitch(residual, "residual")
    p.act = act
    if preact is not None:
        p.preact, p.ld_pre = _ptr(preact), _row_pitch(preact, "preact")
    if dact != ACT_NONE:
        p.dact_aux, p.ld_aux, p.dact = _ptr(dact_aux), _row_pitch(dact_aux, "dact_aux"), dact
    # inference (no_grad) asks for the bit-reproducible policy (forward GEMMs are never split): metrics repeat exactly run to run
    p.accumulate, p.split_k = int(accumulate), (-1 if (split_k == 0 and not accumulate and not torch.is_grad_enabled()) else split_k)
    p.out_group_in, p.out_group_out, p.out_row_offset = out_group
    p.round_f16 = int(round_f16)
    p.ab_dtype = _DTYPES[A.dtype] if f16 else F32
    p.c_dtype = BF16 if out.dtype == torch.bfloat16 else F32
    if colsum_a is not None:  # colsum_a[k] += sum_m A[m, k]: the bias gradient when A is an output gradient
        if colsum_a.numel() != K or not colsum_a.is_contiguous():
            raise ValueError("gemm: colsum_a must be a contiguous [K] vector")
        p.colsum_a = _ptr(_f32(colsum_a, "colsum_a"))
    check(lib.rf_gemm_tf32(C.byref(p), _stream()), "rf_gemm_tf32")
    _count()
    return out


def conv3_assemble_fwd(z, y, n_seq, L, D, pad, bias=None, pe=None, wtime=None):
    lib = _lib.load()
    p = _lib.RfConv3AssembleParams()
    p.z, p.ldz, p.y, p.ldy = _ptr(z), _row_pitch(z, "z"), _ptr(y), _row_pitch(y, "y")
    p.n_seq, p.L, p.D, p.pad = n_seq, L, D, pad
    p.bias, p.wtime = _ptr(bias), _ptr(wtime)
    if pe is not None:
        p.pe, p.ld_pe = _ptr(pe), _row_pitch(pe, "pe")
    check(lib.rf_conv3_assemble_fwd(C.byref(p), _stream()), "rf_conv3_assemble_fwd")
    _count()
    return y


def conv3_assemble_bwd(dy, dz, n_seq, L, D, pad, dbias=None, dwtime=None):
    lib = _lib.load()
    p = _lib.RfConv3AssembleBwdParams()
    p.dy, p.ldy, p.dz, p.ldz = _ptr(dy), _row_pitch(dy, "dy"), _ptr(dz), _row_pitch(dz, "dz")
    p.n_seq, p.L, p.D, p.pad = n_seq, L, D, pad
    p.dbias, p.dwtime = _ptr(dbias), _ptr(dwtime)
    check(lib.rf_conv3_assemble_bwd(C.byref(p), _stream()), "rf_conv3_assemble_bwd")
    _count(1 + (dbias is not None) + (dwtime is not None))
    return dz


def conv3_pack_weight(w: torch.Tensor, wcat: torch.Tensor):
    D, Cin, _ = w.shape
    check(_lib.load().rf_conv3_pack_weight(_ptr(w), _ptr(wcat), D, Cin, wcat.stride(0), _stream()), "rf_conv3_pack_weight")
    _count()
    return wcat


def conv3_unpack_grad(dwcat: torch.Tensor, dw: torch.Tensor):
    D, Cin, _ = dw.shape
    check(_lib.load().rf_conv3_unpack_grad(_ptr(dwcat), _ptr(dw), D, Cin, dwcat.stride(0), _stream()), "rf_conv3_unpack_grad")
    _count()
    return dw


def _attn_params(q, k, v, B, H, Lq, Lk, dh, mode, layout, idx, idx_group, U, u, out, top, measure, forced_top):
    """q/k/v: fp32 tensors whose element (b,l,h,e) sits at base + b*bs + l*ls + h*dh + e; given as (tensor, bs, ls)."""
    p = _lib.RfAttnParams()
    (qt, p.q_bs, p.q_ls), (kt, p.k_bs, p.k_ls), (vt, p.v_bs, p.v_ls) = q, k, v
    p.q, p.k, p.v = _ptr(qt), _ptr(kt), _ptr(vt)
    p.B, p.H, p.Lq, p.Lk, p.dh = B, H, Lq, Lk, dh
    p.mode, p.out_layout = mode, layout
    p.idx, p.idx_group, p.U, p.u = _ptr(idx), idx_group, U, u
    p.out, p.top, p.measure, p.forced_top = _ptr(out), _ptr(top), _ptr(measure), _ptr(forced_top)
    return p


def _set_attn_dropout(p, dropout):
    """dropout = (p, seed, offset[, base]): base = device scalar added to the offset when the kernel runs (CUDA-graph replays)."""
    if dropout is not None and dropout[0] > 0.0:
        p.dropout_p, p.dropout_seed, p.dropout_offset = float(dropout[0]), int(dropout[1]), int(dropout[2])
        if len(dropout) > 3 and dropout[3] is not None:
            p.dropout_offset_base = _ptr(dropout[3])


def attention_fwd(q, k, v, B, H, Lq, Lk, dh, mode, layout, idx, idx_group, U, u, out, top, measure=None, forced_top=None, dropout=None,
                  tail_only: bool = False):
    """dropout = (p, seed, offset): probability dropout, full attention only.  tail_only: only the last query's context is needed
    (the other rows of `out` may stay unwritten)."""
    p = _attn_params(q, k, v, B, H, Lq, Lk, dh, mode, layout, idx, idx_group, U, u, out, top, measure, forced_top)
    _set_attn_dropout(p, dropout)
    p.tail_only = int(tail_only)
    check(_lib.load().rf_attention_fwd(C.byref(p), _stream()), "rf_attention_fwd")
    _count()
    return out


def attention_bwd(q, k, v, B, H, Lq, Lk, dh, mode, layout, U, u, top, dout, dq, dk, dv, dropout=None, tail_only: bool = False):
    """tail_only: rows 0..Lq-2 of `dout` are zero by construction and need not be read."""
    bp = _lib.RfAttnBwdParams()
    bp.f = _attn_params(q, k, v, B, H, Lq, Lk, dh, mode, layout, None, 0, U, u, None, top, None, None)
    _set_attn_dropout(bp.f, dropout)
    bp.f.tail_only = int(tail_only)
    bp.dout, bp.dq, bp.dk, bp.dv = _ptr(dout), _ptr(dq), _ptr(dk), _ptr(dv)
    check(_lib.load().rf_attention_bwd(C.byref(bp), _stream()), "rf_attention_bwd")
    _count()


def dropout(x: torch.Tensor, out: torch.Tensor, p: float, seed: int, offset: int, base: Optional[torch.Tensor] = None,
            residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = residual + keep * x / (1 - p) over a 2-D fp32 view (out may alias x).  The mask is a pure function of (seed, offset
    [+ *base], logical element index): calling it again on a gradient is the backward pass.  `base` is an optional int64 device
    scalar added to `offset` on the device -- the per-step part of the Philox counter of a CUDA-graph-captured step."""
    M, N = x.shape
    _f32(x, "x"), _f32(out, "out")
    check(_lib.load().rf_dropout(_ptr(x), _row_pitch(x, "x"), _ptr(residual), _row_pitch(residual, "residual") if residual is not None else 0,
                                 _ptr(out), _row_pitch(out, "out"), M, N, float(p), int(seed), int(offset), _ptr(base), _stream()), "rf_dropout")
    _count()
    return out


class DropoutStream:
    """Hands out (seed, offset, base) triples for dropout call sites.

    seed   = the CUDA generator's seed (torch.manual_seed sets it without touching the CPU stream that feeds the ProbSparse draws);
    offset = index of the call site within the current step (host side, baked into a captured launch);
    base   = int64 device scalar, advanced by STEP at the start of every training forward (`begin_step`, a device-side add that
             is captured with the step) -- so every step, and every replay of a captured step, draws fresh masks.
    The Philox counter a kernel uses is offset + *base; `log` (tests) records that effective value."""
    STEP = 1 << 20
    site = 0
    _base = {}       # device index -> int64 device scalar
    _base_host = {}  # device index -> host mirror of the value the NEXT kernels will read
    log = None       # tests: list that receives (where, seed, effective offset, rows, cols)

    @classmethod
    def _dev(cls, device) -> int:
        return device.index if device.index is not None else torch.cuda.current_device()

    @classmethod
    def base(cls, device) -> torch.Tensor:
        d = cls._dev(device)
        if d not in cls._base:
            cls._base[d] = torch.zeros(1, dtype=torch.int64, device=torch.device("cuda", d))
            cls._base_host[d] = 0
        return cls._base[d]

    @classmethod
    def begin_step(cls, device) -> None:
        """Start of a training forward: new per-step base (device-side add: capturable), call-site counter back to 0."""
        cls.base(device).add_(cls.STEP)
        if not torch.cuda.is_current_stream_capturing():  # a captured add runs at replay time (note_replay), not now
            cls._base_host[cls._dev(device)] += cls.STEP
        cls.site = 0

    @classmethod
    def snapshot(cls, device):
        return cls.base(device).clone(), cls._base_host[cls._dev(device)]

    @classmethod
    def restore(cls, device, snap) -> None:
        """Undo the advances of forwards that do not count as steps (the warm-up passes of a CUDA-graph capture)."""
        cls.base(device).copy_(snap[0])
        cls._base_host[cls._dev(device)] = snap[1]

    @classmethod
    def note_replay(cls, device) -> None:
        """A captured step (which contains the device-side add of `begin_step`) has been replayed: advance the host mirror."""
        cls._base_host[cls._dev(device)] += cls.STEP

    @classmethod
    def next(cls, device, where: str = "", rows: int = 0, cols: int = 0):
        seed = torch.cuda.default_generators[cls._dev(device)].initial_seed()
        base = cls.base(device)
        cls.site += 1
        if cls.log is not None:
            cls.log.append((where, seed, cls._base_host[cls._dev(device)] + cls.site, rows, cols))
        return seed, cls.site, base


def layernorm_fwd(x, gamma, beta, y, mean, rstd):
    M, D = x.shape
    check(_lib.load().rf_layernorm_fwd(_ptr(x), _row_pitch(x, "x"), _ptr(gamma), _ptr(beta), _ptr(y), _row_pitch(y, "y"),
                                       _ptr(mean), _ptr(rstd), M, D, _stream()), "rf_layernorm_fwd")
    _count()
    return y


def layernorm_bwd(dy, x, gamma, mean, rstd, dx, dgamma, dbeta):
    M, D = x.shape
    check(_lib.load().rf_layernorm_bwd(_ptr(dy), _row_pitch(dy, "dy"), _ptr(x), _row_pitch(x, "x"), _ptr(gamma), _ptr(mean),
                                       _ptr(rstd), _ptr(dx), _row_pitch(dx, "dx"), _ptr(dgamma), _ptr(dbeta), M, D, _stream()),
          "rf_layernorm_bwd")
    _count()
    return dx


def distil_fwd(z, B, Lz, D, gamma, beta, running_mean, running_var, training, mean, rstd, out, argmax, momentum=0.1, eps=1e-5):
    p = _lib.RfDistilParams()
    p.z, p.B, p.Lz, p.D = _ptr(z), B, Lz, D
    p.gamma, p.beta, p.running_mean, p.running_var = _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var)
    p.training, p.momentum, p.eps = int(training), momentum, eps
    p.mean, p.rstd, p.out, p.argmax = _ptr(mean), _ptr(rstd), _ptr(out), _ptr(argmax)
    check(_lib.load().rf_distil_fwd(C.byref(p), _stream()), "rf_distil_fwd")
    _count(3 if training else 2)
    return out


def distil_bwd(z, B, Lz, D, gamma, beta, mean, rstd, training, argmax, dout, dz, dgamma, dbeta, scratch):
    p = _lib.RfDistilBwdParams()
    p.z, p.B, p.Lz, p.D = _ptr(z), B, Lz, D
    p.gamma, p.beta, p.mean, p.rstd, p.training = _ptr(gamma), _ptr(beta), _ptr(mean), _ptr(rstd), int(training)
    p.argmax, p.dout, p.dz, p.dgamma, p.dbeta, p.scratch = _ptr(argmax), _ptr(dout), _ptr(dz), _ptr(dgamma), _ptr(dbeta), _ptr(scratch)
    check(_lib.load().rf_distil_bwd(C.byref(p), _stream()), "rf_distil_bwd")
    _count(2)
    return dz


def motion_features(gps, visual, x, origin, E, rotate, normalize, mean, std, input_is_motion=False):
    B, T, _ = gps.shape
    ld_vis = visual.stride(1) if visual is not None else 0
    check(_lib.load().rf_motion_features(_ptr(gps), _ptr(visual), ld_vis, _ptr(x), x.stride(1), _ptr(origin), B, T, E, int(rotate),
                                         int(normalize), float(mean), float(std), int(input_is_motion), _stream()), "rf_motion_features")
    _count()
    return x


def decoder_input_fwd(x, xdec, P, smart):
    B, T, ld = x.shape
    check(_lib.load().rf_decoder_input_fwd(_ptr(x), _ptr(xdec), B, T, P, ld, int(smart), _stream()), "rf_decoder_input_fwd")
    _count()
    return xdec


def decoder_input_bwd(dxdec, dx, P, smart):
    B, T, ld = dx.shape
    check(_lib.load().rf_decoder_input_bwd(_ptr(dxdec), _ptr(dx), B, T, P, ld, int(smart), _stream()), "rf_decoder_input_bwd")
    _count()
    return dx


def stream_tokens_fwd(src, F, first, step, dense, emb, tokens, B, T, E, tokens_per_clip, t_off):
    check(_lib.load().rf_stream_tokens_fwd(_ptr(src), F, first, step, int(dense), _ptr(emb), _ptr(tokens), B, T, E, tokens_per_clip,
                                           t_off, _stream()), "rf_stream_tokens_fwd")
    _count()
    return tokens


def stream_tokens_bwd(dtokens, dsrc, F, first, step, dense, demb, B, T, E, tokens_per_clip, t_off):
    check(_lib.load().rf_stream_tokens_bwd(_ptr(dtokens), _ptr(dsrc), F, first, step, int(dense), _ptr(demb), B, T, E,
                                           tokens_per_clip, t_off, _stream()), "rf_stream_tokens_bwd")
    _count((dsrc is not None) + (demb is not None))


def decode_waypoints_fwd(out, origin, last_gps, waypoints, motion, rotate, normalize, mean, std):
    B, P, _ = out.shape
    check(_lib.load().rf_decode_waypoints_fwd(_ptr(out), out.stride(1), _ptr(origin), _ptr(last_gps), _ptr(waypoints), _ptr(motion), B, P,
                                              int(rotate), int(normalize), float(mean), float(std), _stream()), "rf_decode_waypoints_fwd")
    _count()


def decode_waypoints_bwd(dwaypoints, origin, dout, rotate, normalize, std):
    B, P, _ = dout.shape
    check(_lib.load().rf_decode_waypoints_bwd(_ptr(dwaypoints), _ptr(origin), _ptr(dout), dout.stride(1), B, P, int(rotate),
                                              int(normalize), float(std), _stream()), "rf_decode_waypoints_bwd")
    _count()


def median_downsample(x: torch.Tensor, target: int) -> torch.Tensor:
    B, S, Cc = x.shape
    x = _f32(x, "x").contiguous()
    y = torch.empty(B, target, Cc, device=x.device, dtype=torch.float32)
    check(_lib.load().rf_median_downsample(_ptr(x), _ptr(y), B, S, Cc, target, _stream()), "rf_median_downsample")
    _count()
    return y


def ade_fde(pred: torch.Tensor, truth: torch.Tensor, per_sample: bool = False):
    B, T, _ = pred.shape
    pred, truth = _f32(pred, "pred").contiguous(), _f32(truth, "truth").contiguous()
    result = torch.empty(2, device=pred.device, dtype=torch.float32)
    ps = torch.empty(B, 2, device=pred.device, dtype=torch.float32) if per_sample else None
    check(_lib.load().rf_ade_fde(_ptr(pred), _ptr(truth), B, T, _ptr(result), _ptr(ps), _stream()), "rf_ade_fde")
    _count()
    return result, ps


LOSS_KINDS = {"smooth_l1": 0, "mse": 1, "mae": 2}


def eval_samples(preds: torch.Tensor, truth: torch.Tensor, gamma: float, epsilon: float, kind: str = "smooth_l1"):
    """preds [S,B,T,2] stacked stochastic forwards, truth [B,T,2] -> (mean prediction [B,T,2], per clip [B,3] = loss, ade, fde)."""
    S, B, T, _ = preds.shape
    preds, truth = _f32(preds, "preds").contiguous(), _f32(truth, "truth").contiguous()
    mean = torch.empty(B, T, 2, device=preds.device, dtype=torch.float32)
    per_clip = torch.empty(B, 3, device=preds.device, dtype=torch.float32)
    check(_lib.load().rf_eval_samples(_ptr(preds), _ptr(truth), S, B, T, float(gamma), float(epsilon), LOSS_KINDS[kind], _ptr(mean),
                                      _ptr(per_clip), _stream()), "rf_eval_samples")
    _count()
    return mean, per_clip


def discounted_loss_fwd(pred, truth, gamma, epsilon, kind):
    B, T = pred.shape[:2]
    Cc = pred[0, 0].numel()
    p2, t2 = pred.reshape(B * T, Cc), truth.reshape(B * T, Cc)
    loss = torch.empty(1, device=pred.device, dtype=torch.float32)
    check(_lib.load().rf_discounted_loss_fwd(_ptr(p2), _row_pitch(p2, "pred"), _ptr(t2), _row_pitch(t2, "truth"), B, T, Cc, float(gamma),
                                             float(epsilon), LOSS_KINDS[kind], _ptr(loss), _stream()), "rf_discounted_loss_fwd")
    _count()
    return loss


def discounted_loss_bwd(pred, truth, gamma, epsilon, kind, dloss, scale, dpred, accumulate=False):
    B, T = pred.shape[:2]
    Cc = pred[0, 0].numel()
    p2, t2, d2 = pred.reshape(B * T, Cc), truth.reshape(B * T, Cc), dpred.reshape(B * T, Cc)
    check(_lib.load().rf_discounted_loss_bwd(_ptr(p2), _row_pitch(p2, "pred"), _ptr(t2), _row_pitch(t2, "truth"), B, T, Cc, float(gamma),
                                             float(epsilon), LOSS_KINDS[kind], _ptr(dloss), float(scale), _ptr(d2),
                                             _row_pitch(d2, "dpred"), int(accumulate), _stream()), "rf_discounted_loss_bwd")
    _count()
    return dpred


def colsum_accumulate(src: torch.Tensor, dst: torch.Tensor):
    M, N = src.shape
    check(_lib.load().rf_colsum_accumulate(_ptr(src), _row_pitch(src, "src"), M, N, _ptr(dst), _stream()), "rf_colsum_accumulate")
    _count()
    return dst


def sumsq_accumulate(x: torch.Tensor, out: torch.Tensor):
    check(_lib.load().rf_sumsq_accumulate(_ptr(x), x.numel(), _ptr(out), _stream()), "rf_sumsq_accumulate")
    _count()
    return out


def adamw_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, gnorm_sq=None,
               max_norm=0.0):
    check(_lib.load().rf_adamw_step(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(), float(lr), float(beta1),
                                    float(beta2), float(eps), float(weight_decay), int(step), float(grad_scale), _ptr(gnorm_sq),
                                    float(max_norm), _stream()), "rf_adamw_step")
    _count()


def stage_frames_h2d(host_video: torch.Tensor, times, device, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Pinned host video [B,T,3,H,W] -> device [B,len(times),3,H,W] holding only the frames at `times` (async on the current stream)."""
    if host_video.is_cuda or not host_video.is_pinned() or not host_video.is_contiguous():
        raise ValueError("stage_frames_h2d needs a contiguous pinned host tensor")
    B, T = host_video.shape[:2]
    times = [int(t) for t in times]
    shape = (B, len(times)) + tuple(host_video.shape[2:])
    dst = out if out is not None else torch.empty(shape, device=device, dtype=host_video.dtype)
    if tuple(dst.shape) != shape or dst.dtype != host_video.dtype or not dst.is_contiguous():
        raise ValueError("stage_frames_h2d: `out` does not match the staged shape")
    frame_bytes = host_video[0, 0].numel() * host_video.element_size()
    arr = (C.c_int * len(times))(*times)
    with torch.cuda.device(device):
        check(_lib.load().rf_stage_frames_h2d(dst.data_ptr(), host_video.data_ptr(), B, T, arr, len(times), frame_bytes, _stream()),
              "rf_stage_frames_h2d")
    return dst
