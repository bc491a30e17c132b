"""GPU parity tests, one per C-ABI entry point: CUDA kernel (through ctypes) vs the CPU oracle / plain PyTorch.

Tolerances: integer/index results bit-exact; fp32 elementwise kernels 1e-5; TF32 tensor-core GEMM 2e-3 relative
(10-bit mantissa operands, fp32 accumulate -- SURVEY 7, hard part 7).
"""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import routeformer_oracle as O
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from routeformer_b200 import ops as _ops

    return _ops


def g(seed):
    return torch.Generator().manual_seed(seed)


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (300, 200, 96), (1000, 384, 1024), (70, 66, 832), (4480, 832, 200), (513, 64, 64)])
def test_gemm_layouts(ops, a_mn, b_mn, M, N, K):
    if (a_mn and M % 4) or (b_mn and N % 4):
        pytest.skip("MN-major operand needs a pitch that is a multiple of 4")
    gen = g(M + N + K)
    A = torch.randn(M, K, generator=gen)
    B = torch.randn(N, K, generator=gen)
    ref = (A.double() @ B.double().t()).float()
    Ad = (A.t().contiguous() if a_mn else A).to(DEV)
    Bd = (B.t().contiguous() if b_mn else B).to(DEV)
    out = torch.full((M, N), float("nan"), device=DEV)
    ops.gemm(Ad, Bd, out, a_mn=a_mn, b_mn=b_mn)
    torch.cuda.synchronize()
    assert rel_err(out.cpu(), ref) < 2e-3


def test_gemm_random_shapes(ops):
    """Random problem sizes (ragged edge tiles in M and N, one- and few-row problems, K from one MMA to the split-K range) in every
    operand layout the pitches allow, with the epilogues of the model's layers; the output is pre-filled with NaN so that a
    tile that is never written shows."""
    gen = g(77)
    ri = lambda lo, hi: int(torch.randint(lo, hi, (1,), generator=gen))
    failures = []
    for i in range(40):
        M, N, K = ri(1, 2500), ri(1, 700), 4 * ri(1, 375)
        if i % 8 == 0:
            M, N = ri(1, 9), ri(1, 9)
        a_mn, b_mn = bool(ri(0, 2)) and M % 4 == 0, bool(ri(0, 2)) and N % 4 == 0
        A, B = torch.randn(M, K, generator=gen), torch.randn(N, K, generator=gen) / math.sqrt(K)
        bias, res, c0 = torch.randn(N, generator=gen), torch.randn(M, N, generator=gen), torch.randn(M, N, generator=gen)
        Ad = (A.t().contiguous() if a_mn else A).to(DEV)
        Bd = (B.t().contiguous() if b_mn else B).to(DEV)
        prod = (A.double() @ B.double().t()).float()
        epi = i % 4
        if epi == 0:
            out = torch.full((M, N), float("nan"), device=DEV)
            ops.gemm(Ad, Bd, out, a_mn=a_mn, b_mn=b_mn)
            ref = prod
        elif epi == 1:
            out = torch.full((M, N), float("nan"), device=DEV)
            ops.gemm(Ad, Bd, out, a_mn=a_mn, b_mn=b_mn, bias=bias.to(DEV), residual=res.to(DEV))
            ref = prod + bias + res
        elif epi == 2:
            out = c0.clone().to(DEV)
            ops.gemm(Ad, Bd, out, a_mn=a_mn, b_mn=b_mn, accumulate=True)
            ref = prod + c0
        else:
            out = torch.full((M, N), float("nan"), device=DEV)
            ops.gemm(Ad, Bd, out, a_mn=a_mn, b_mn=b_mn, bias=bias.to(DEV), act=ops.ACT_GELU)
            ref = F.gelu(prod + bias)
        got = out.cpu()
        if not (torch.isfinite(got).all() and (got - ref).norm() < 2e-3 * max(float(ref.norm()), 1e-3 * math.sqrt(M * N))):
            failures.append((i, M, N, K, a_mn, b_mn, epi, rel_err(got, ref)))
    assert not failures, failures


def test_gemm_epilogue_forward(ops):
    gen = g(1)
    M, N, K, period = 260, 136, 72, 65
    A, B = torch.randn(M, K, generator=gen), torch.randn(N, K, generator=gen) / math.sqrt(K)
    bias, pe, res = torch.randn(N, generator=gen), torch.randn(period, N, generator=gen), torch.randn(M, N, generator=gen)
    pre_ref = (A.double() @ B.double().t()).float() + bias + pe.repeat(M // period, 1) + res
    for act, fn in [(ops.ACT_GELU, F.gelu), (ops.ACT_RELU, F.relu), (ops.ACT_NONE, lambda x: x)]:
        out = torch.empty(M, N, device=DEV)
        pre = torch.empty(M, N, device=DEV)
        ops.gemm(A.to(DEV), B.to(DEV), out, bias=bias.to(DEV), rowadd=pe.to(DEV), rowadd_period=period, residual=res.to(DEV),
                 act=act, preact=pre)
        assert rel_err(pre.cpu(), pre_ref) < 2e-3
        assert rel_err(out.cpu(), fn(pre_ref)) < 2e-3


def test_gemm_fp16_operands(ops):
    """kind::f16 path (fp16 A and B, fp32 accumulate): the patch-embedding GEMM under the reference's fp16 autocast."""
    gen = g(4)
    for (M, N, K) in [(300, 136, 200), (640, 1024, 3072), (130, 48, 192)]:
        A = torch.randn(M, K, generator=gen).half()
        B = (torch.randn(N, K, generator=gen) / math.sqrt(K)).half()
        bias = torch.randn(N, generator=gen)
        ref = (A.double() @ B.double().t()).float() + bias
        out = torch.empty(M, N, device=DEV)
        ops.gemm(A.to(DEV), B.to(DEV), out, bias=bias.to(DEV))
        assert rel_err(out.cpu(), ref) < 1e-5  # exact products of fp16 values, fp32 accumulation
    # row remap + fp16 rounding of the result, as encode_views uses it
    n, K, N = 6, 192, 40
    A = torch.randn(n * 8, K, generator=gen).half()
    W = (torch.randn(N, K, generator=gen) / math.sqrt(K)).half()
    buf = torch.full((n * 9, N), -1.0, device=DEV)
    ops.gemm(A.to(DEV), W.to(DEV), buf, out_group=(8, 9, 0), round_f16=True)
    ref = (A.float() @ W.float().t()).half().float().view(n, 8, N)
    got = buf.cpu().view(n, 9, N)
    assert rel_err(got[:, :8], ref) < 1e-3
    assert torch.equal(got[:, 8], torch.full((n, N), -1.0))
    with pytest.raises(TypeError):
        ops.gemm(A.to(DEV), W.float().to(DEV), buf)
    with pytest.raises(TypeError):
        ops.gemm(A.to(DEV), W.to(DEV), buf, b_mn=True)


def test_gemm_operand_column_sums(ops):
    """colsum_a: bias gradients taken from the A tiles of a dgrad GEMM (persistent kernel: reducer warp; long K: separate pass)."""
    gen = g(6)
    for (M, N, K) in [(5000, 160, 96), (777, 384, 132), (99, 64, 40), (300, 128, 1000)]:
        A = torch.randn(M, K, generator=gen)
        B = torch.randn(K, N, generator=gen) / math.sqrt(K)   # dgrad form: B is [K][N] memory (b_mn)
        res = torch.randn(M, N, generator=gen)
        out = torch.empty(M, N, device=DEV)
        cs = torch.full((K,), 3.0, device=DEV)
        ops.gemm(A.to(DEV), B.to(DEV), out, b_mn=True, residual=res.to(DEV), colsum_a=cs)
        assert rel_err(out.cpu(), (A.double() @ B.double()).float() + res) < 2e-3
        # short reductions sum the TF32-rounded operand tiles (unbiased RN rounding, 2^-11 per element), long ones the fp32 data
        assert rel_err(cs.cpu() - 3.0, A.sum(0)) < 1e-3, (M, N, K)
    with pytest.raises(RuntimeError):
        ops.gemm(A.t().contiguous().to(DEV), B.to(DEV), out, a_mn=True, b_mn=True, colsum_a=cs)


def test_gemm_split_linear_epilogue(ops):
    """Few tiles + long reduction + linear epilogue (the Informer layers): the library splits K on its own; bias, positional
    rows and the residual must be added exactly once, and a residual aliasing the output must keep working."""
    gen = g(8)
    for (M, N, K, period) in [(192, 832, 3328, 0), (704, 832, 832, 64), (130, 136, 2000, 0)]:
        A, B = torch.randn(M, K, generator=gen), torch.randn(N, K, generator=gen) / math.sqrt(K)
        bias, res = torch.randn(N, generator=gen), torch.randn(M, N, generator=gen)
        pe = torch.randn(max(period, 1), N, generator=gen)
        ref = (A.double() @ B.double().t()).float() + bias + res + (pe.repeat(M // period, 1) if period else 0)
        out = torch.full((M, N), float("nan"), device=DEV)
        ops.gemm(A.to(DEV), B.to(DEV), out, bias=bias.to(DEV), residual=res.to(DEV),
                 rowadd=pe.to(DEV) if period else None, rowadd_period=period)
        assert rel_err(out.cpu(), ref) < 2e-3, (M, N, K)
        with torch.no_grad():  # inference policy: no split -> bit-reproducible
            o1, o2 = torch.empty(M, N, device=DEV), torch.empty(M, N, device=DEV)
            for o in (o1, o2):
                ops.gemm(A.to(DEV), B.to(DEV), o, bias=bias.to(DEV), residual=res.to(DEV), rowadd=pe.to(DEV) if period else None,
                         rowadd_period=period)
            assert torch.equal(o1, o2) and rel_err(o1.cpu(), ref) < 2e-3
        acc = res.to(DEV).clone()  # out aliases the residual (dx += ... chains): must not be zero-filled
        ops.gemm(A.to(DEV), B.to(DEV), acc, bias=bias.to(DEV), residual=acc)
        assert rel_err(acc.cpu(), (A.double() @ B.double().t()).float() + bias + res) < 2e-3
        wide = torch.full((M, N + 8), 7.0, device=DEV)  # strided output: the columns beyond N stay untouched
        ops.gemm(A.to(DEV), B.to(DEV), wide[:, :N], bias=bias.to(DEV))
        assert rel_err(wide[:, :N].cpu(), (A.double() @ B.double().t()).float() + bias) < 2e-3
        assert torch.equal(wide[:, N:].cpu(), torch.full((M, 8), 7.0))


def test_gemm_strided_rows_and_row_remap(ops):
    gen = g(2)
    n, K, N = 6, 96, 40
    # A = last token of each 9-token frame: rows strided by 9*K
    tok = torch.randn(n * 9, K, generator=gen)
    W = torch.randn(N, K, generator=gen)
    Ad = tok.to(DEV)
    view = Ad.view(n, 9, K)[:, 8, :]
    out = torch.empty(n, N, device=DEV)
    ops.gemm(view, W.to(DEV), out)
    assert rel_err(out.cpu(), tok.view(n, 9, K)[:, 8] @ W.t()) < 2e-3
    # row remap: 8 rows per group in, 9 per group out (the appended "-1 token" row stays untouched)
    A = torch.randn(n * 8, K, generator=gen)
    buf = torch.full((n * 9, N), -1.0, device=DEV)
    ops.gemm(A.to(DEV), W.to(DEV), buf, out_group=(8, 9, 0), round_f16=True)
    ref = (A @ W.t()).half().float().view(n, 8, N)
    got = buf.cpu().view(n, 9, N)
    assert rel_err(got[:, :8], ref) < 2e-3
    assert torch.equal(got[:, 8], torch.full((n, N), -1.0))
    assert torch.equal(got[:, :8].half().float(), got[:, :8])  # representable in fp16


def test_gemm_bf16_operands_and_bf16_output(ops):
    """bf16 operand mode (north star): kind::f16 MMAs on bf16 operands with fp32 accumulation; the result optionally stored as bf16
    (rounded once, after the fp32 epilogue) with the patch embedding's 64 -> 65 row remap.  Against fp64 arithmetic on the SAME
    bf16-rounded operands the fp32 result is exact to accumulation order; the bf16 result is within half a bf16 ulp (2^-9)."""
    gen = g(17)
    M, N, K = 3 * 64, 200, 328
    A = (torch.randn(M, K, generator=gen) / math.sqrt(K)).to(torch.bfloat16)
    W = torch.randn(N, K, generator=gen).to(torch.bfloat16)
    bias = torch.randn(N, generator=gen)
    ref = A.double() @ W.double().t() + bias.double()
    out = torch.empty(M, N, device=DEV)
    ops.gemm(A.to(DEV), W.to(DEV), out, bias=bias.to(DEV))
    assert rel_err(out.cpu(), ref.float()) < 2e-6
    # bf16 C through the row remap (groups of 64 rows stored with pitch 65, the 65th row untouched)
    buf = torch.full((3 * 65, 208), -1.0, device=DEV, dtype=torch.bfloat16)
    ops.gemm(A.to(DEV), W.to(DEV), buf[:, :N], bias=bias.to(DEV), out_group=(64, 65, 0))
    got = buf.cpu().float().view(3, 65, 208)
    want = ref.float().view(3, 64, N)
    # (fp32 accumulation order differs from the fp64 reference: a value that sits on a bf16 rounding boundary may land on either
    #  side, so the bound is one bf16 ulp, 2^-8 relative, and nearly all values are the correctly rounded ones)
    assert ((got[:, :64, :N] - want).abs() <= 2.0 ** -8 * want.abs() + 1e-6).all()
    assert (got[:, :64, :N] == want.to(torch.bfloat16).float()).float().mean() > 0.99
    assert torch.equal(got[:, 64], torch.full((3, 208), -1.0)) and torch.equal(got[:, :, N:], torch.full((3, 65, 208 - N), -1.0))
    with pytest.raises(TypeError):
        ops.gemm(A.to(DEV), W.to(DEV).half(), out)
    with pytest.raises(TypeError):
        ops.gemm(A.to(DEV).float(), W.to(DEV).float(), buf[:, :N])


def test_gemm_dact_and_accumulate(ops):
    gen = g(3)
    M, N, K = 5000, 96, 160
    dY, X = torch.randn(M, N, generator=gen), torch.randn(M, K, generator=gen)
    # wgrad: dW[N,K] += dY^T X, split over the long reduction, atomically accumulated
    dW = torch.ones(N, K, device=DEV)
    ops.gemm(dY.to(DEV), X.to(DEV), dW, a_mn=True, b_mn=True, accumulate=True)
    assert rel_err(dW.cpu(), 1.0 + (dY.double().t() @ X.double()).float()) < 2e-3
    # dgrad with activation derivative in the epilogue
    W = torch.randn(N, K, generator=gen)
    pre = torch.randn(M, K, generator=gen)
    pg = pre.clone().requires_grad_()
    F.gelu(pg).sum().backward()
    for code, grad in [(ops.ACT_RELU, (pre > 0).float()), (ops.ACT_GELU, pg.grad)]:
        out = torch.empty(M, K, device=DEV)
        ops.gemm(dY.to(DEV), W.to(DEV), out, b_mn=True, dact=code, dact_aux=pre.to(DEV))
        assert rel_err(out.cpu(), (dY.double() @ W.double()).float() * grad) < 2e-3
    # the FFN pairing: forward stores gelu'(pre) beside gelu(pre), backward multiplies by the saved derivative
    Xs, W1 = torch.randn(M, N, generator=gen), torch.randn(K, N, generator=gen) / math.sqrt(N)
    b1 = torch.randn(K, generator=gen)
    pre_ref = ((Xs.double() @ W1.double().t()) + b1).float().requires_grad_()
    F.gelu(pre_ref).sum().backward()
    h, saved = torch.empty(M, K, device=DEV), torch.empty(M, K, device=DEV)
    ops.gemm(Xs.to(DEV), W1.to(DEV), h, bias=b1.to(DEV), act=ops.ACT_GELU_SAVE_GRAD, preact=saved)
    assert rel_err(h.cpu(), F.gelu(pre_ref.detach())) < 2e-3
    assert rel_err(saved.cpu(), pre_ref.grad) < 2e-3
    out = torch.empty(M, K, device=DEV)
    ops.gemm(dY.to(DEV), W.to(DEV), out, b_mn=True, dact=ops.DACT_SAVED, dact_aux=saved)
    assert rel_err(out.cpu(), (dY.double() @ W.double()).float() * pre_ref.grad) < 2e-3
    with pytest.raises(RuntimeError):
        ops.gemm(Xs.to(DEV), W1.to(DEV), h, act=ops.ACT_GELU_SAVE_GRAD)


# ------------------------------------------------------------------------------------------------ FoV crop
# the crop kernels: separable column walker over a cp.async-staged source window with four (default) or two output columns per
# thread, the same walk reading global memory (unaligned frames), and the round-1 direct 4-tap gather (tall frames)
CROP_KERNELS = ["staged", "pairs", "walk", "direct"]


def select_crop_kernel(monkeypatch, kernel):
    monkeypatch.setenv("RF_CROP_DIRECT", "1" if kernel == "direct" else "0")
    monkeypatch.setenv("RF_CROP_NOSTAGE", "1" if kernel == "walk" else "0")
    monkeypatch.setenv("RF_CROP_PAIRS", "1" if kernel == "pairs" else "0")
    monkeypatch.setenv("RF_CROP_WALK", "1")  # (the walker also for tall frames, which the dispatcher would hand to the direct kernel)


@pytest.mark.parametrize("H,W,S,patch", [(324, 326, 224, 28), (86, 384, 256, 32), (36, 34, 32, 8), (240, 320, 64, 0), (37, 35, 32, 8),
                                         (216, 768, 256, 32), (64, 1088, 224, 28)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.uint8])
@pytest.mark.parametrize("kernel", CROP_KERNELS)
def test_fov_crop(ops, monkeypatch, H, W, S, patch, dtype, kernel):
    select_crop_kernel(monkeypatch, kernel)
    gen = g(H + W)
    n = 5
    if dtype == torch.uint8:
        frames = torch.randint(0, 256, (n + 2, 3, H, W), generator=gen, dtype=torch.uint8)
        ref_frames = frames.float() / 255.0
    else:
        frames = torch.rand(n + 2, 3, H, W, generator=gen).to(dtype)
        ref_frames = frames.float()
    ids = torch.tensor([6, 0, 3, 3, 5], dtype=torch.int32)
    centers = (0.5 + 0.25 * torch.randn(n, 2, generator=gen)).clamp(0, 1)
    cx, cy, fw, fh = O.frame_window(H, W)  # the reference-like "pad to square and resize" window
    windows = torch.tensor([[0.5, 0.5], [1.0, 1.0], [fw, fh], [0.3, 0.7], [1.4, 0.2]])
    centers[2, 0], centers[2, 1] = cx, cy
    spec = O.BackboneSpec()
    ref = O.fov_crop(ref_frames[ids.long()], centers, windows, S, spec.mean, spec.std)
    out = ops.fov_crop(frames.to(DEV), centers.to(DEV), windows.to(DEV), S, spec.mean, spec.std, patch=patch, frame_ids=ids.to(DEV))
    got = out.cpu()
    if patch:
        G = S // patch
        ref = ref.view(n, 3, G, patch, G, patch).permute(0, 2, 4, 1, 3, 5).reshape(n * G * G, 3 * patch * patch)
    # white-noise frames are the worst case: |dI/dx| ~ 1/std per pixel, so a 1e-4 px difference in the fp32
    # sample position (ATen builds the grid with linspace + matmul) moves the value by ~5e-4
    assert (got - ref).abs().max() < 1e-3 and (got - ref).abs().mean() < 2.5e-5
    bf = ops.fov_crop(frames.to(DEV), centers.to(DEV), windows.to(DEV), S, spec.mean, spec.std, patch=patch, frame_ids=ids.to(DEV),
                      out_dtype=torch.bfloat16)
    assert (bf.float().cpu() - ref).abs().max() < 3e-2
    hf = ops.fov_crop(frames.to(DEV), centers.to(DEV), windows.to(DEV), S, spec.mean, spec.std, patch=patch, frame_ids=ids.to(DEV),
                      out_dtype=torch.float16)
    assert torch.equal(hf.cpu(), got.half())  # same values, rounded to fp16 (the A operand of the fp16 patch GEMM)


@pytest.mark.parametrize("kernel", CROP_KERNELS)
def test_fov_crop_uint8_frames_like_the_reference_loader(ops, monkeypatch, kernel):
    """raw uint8 frames staged on the device (half the host->device bytes) and converted inside the crop kernel
    exactly as the reference's loader converts them on the host, `astype(float16) / 255.0` (io/dataset.py:1505-1522): the crop of
    the uint8 frames must equal the crop of the host-converted fp16 frames BIT FOR BIT."""
    import numpy as np

    select_crop_kernel(monkeypatch, kernel)
    gen = g(5)
    n, H, W, S, patch = 4, 60, 62, 32, 8
    u8 = torch.randint(0, 256, (n, 3, H, W), generator=gen, dtype=torch.uint8)
    f16 = torch.from_numpy(u8.numpy().astype(np.float16) / 255.0)
    assert f16.dtype == torch.float16
    centers = (0.5 + 0.2 * torch.randn(n, 2, generator=gen)).clamp(0, 1).to(DEV)
    windows = torch.full((n, 2), 0.6).to(DEV)
    spec = O.BackboneSpec()
    for p_, dt in ((patch, torch.float16), (0, torch.float32)):
        a = ops.fov_crop(u8.to(DEV), centers, windows, S, spec.mean, spec.std, patch=p_, out_dtype=dt, u8_as_f16=True)
        b = ops.fov_crop(f16.to(DEV), centers, windows, S, spec.mean, spec.std, patch=p_, out_dtype=dt)
        assert torch.equal(a, b)
    # every one of the 256 pixel values converts like numpy does
    ramp = torch.arange(256, dtype=torch.uint8).view(1, 1, 16, 16).repeat(1, 3, 1, 1)
    ident = ops.fov_crop(ramp.to(DEV), torch.tensor([[0.5, 0.5]], device=DEV), torch.tensor([[1.0, 1.0]], device=DEV), 16,
                         (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), out_dtype=torch.float32, u8_as_f16=True)
    want = torch.from_numpy(np.arange(256, dtype=np.uint8).astype(np.float16) / 255.0).float().view(16, 16)
    assert torch.equal(ident[0, 0].cpu(), want)


@pytest.mark.parametrize("kernel", CROP_KERNELS)
def test_fov_crop_mirrored_and_offscreen(ops, monkeypatch, kernel):
    """Windows off the walker kernel's common path: mirrored (fw < 0: sample positions decrease with the index), entirely
    outside the frame (no loads at all), touching the left / right / top / bottom border (per-tap predicates)."""
    select_crop_kernel(monkeypatch, kernel)
    gen = g(77)
    H, W, S = 48, 50, 32
    frames = torch.rand(6, 3, H, W, generator=gen).half()
    centers = torch.tensor([[0.5, 0.5], [3.0, 0.5], [0.0, 0.0], [1.0, 1.0], [0.5, -2.0], [0.02, 0.98]])
    windows = torch.tensor([[-0.6, 0.8], [0.5, 0.5], [0.5, 0.5], [0.4, 0.4], [0.5, 0.5], [0.1, 0.1]])
    spec = O.BackboneSpec()
    ref = O.fov_crop(frames.float(), centers, windows, S, spec.mean, spec.std)
    for patch in (0, 8):
        out = ops.fov_crop(frames.to(DEV), centers.to(DEV), windows.to(DEV), S, spec.mean, spec.std, patch=patch).cpu()
        want = ref
        if patch:
            G = S // patch
            want = ref.view(6, 3, G, patch, G, patch).permute(0, 2, 4, 1, 3, 5).reshape(6 * G * G, 3 * patch * patch)
        assert (out - want).abs().max() < 1e-3, patch


def test_fov_crop_random_windows(ops):
    """The dispatcher's own kernel choice on random problems: frame sizes with odd widths / heights (rows that are not 4-byte
    aligned take the unstaged walk, tall frames the direct gather), fp16 / fp32 / uint8 sources, crop sizes 20..256 planar and
    patch-major, centres inside and outside the frame, windows from a twentieth of the frame to 1.6 frames, some mirrored."""
    gen = g(2024)
    spec = O.BackboneSpec()
    ri = lambda lo, hi: int(torch.randint(lo, hi, (1,), generator=gen))
    crops = [(32, 8), (32, 0), (64, 16), (128, 32), (224, 28), (256, 32), (96, 0), (20, 4), (44, 0)]
    dtypes = [torch.float16, torch.float32, torch.uint8]
    for i in range(36):
        H, W = ri(8, 200), ri(8, 400)
        S, patch = crops[i % len(crops)]
        dtype = dtypes[ri(0, 3)]
        n = 3
        if dtype == torch.uint8:
            frames = torch.randint(0, 256, (n, 3, H, W), generator=gen, dtype=torch.uint8)
            ref_frames = frames.float() / 255.0
        else:
            frames = torch.rand(n, 3, H, W, generator=gen).to(dtype)
            ref_frames = frames.float()
        centers = 0.5 + 0.4 * torch.randn(n, 2, generator=gen)
        windows = torch.exp(torch.empty(n, 2).uniform_(math.log(0.05), math.log(1.6), generator=gen))
        windows = windows * torch.where(torch.rand(n, 2, generator=gen) < 0.1, -1.0, 1.0)
        ref = O.fov_crop(ref_frames, centers, windows, S, spec.mean, spec.std)
        got = ops.fov_crop(frames.to(DEV), centers.to(DEV), windows.to(DEV), S, spec.mean, spec.std, patch=patch).cpu()
        if patch:
            G = S // patch
            ref = ref.view(n, 3, G, patch, G, patch).permute(0, 2, 4, 1, 3, 5).reshape(n * G * G, 3 * patch * patch)
        err = (got - ref).abs()
        assert err.max() < 1e-3 and err.mean() < 5e-5, (i, H, W, S, patch, dtype, float(err.max()), float(err.mean()), centers.tolist(), windows.tolist())


# ------------------------------------------------------------------------------------------------ circular conv
@pytest.mark.parametrize("n,L,C,D,pad", [(3, 65, 48, 128, 1), (4, 40, 8, 64, 1), (2, 21, 64, 64, 2), (2, 5, 32, 32, 2), (2, 4, 32, 32, 2)])
def test_conv3_assemble(ops, n, L, C, D, pad):
    gen = g(L + D)
    x = torch.randn(n, L, C, generator=gen)
    w = torch.randn(D, C, 3, generator=gen) / math.sqrt(3 * C)
    bias, wt = torch.randn(D, generator=gen), torch.randn(D, generator=gen)
    L_out = L + 2 * pad - 2
    pe = O.pe_table(L_out + 3, D)
    xr, wr, br, wtr = x.clone().requires_grad_(), w.clone().requires_grad_(), bias.clone().requires_grad_(), wt.clone().requires_grad_()
    t = torch.arange(L_out, dtype=torch.float32).view(1, L_out, 1)
    ref = O.circular_conv3(xr, wr, br, pad) + pe[:L_out] + t * wtr
    dy = torch.randn(n, L_out, D, generator=gen)
    ref.backward(dy)
    # CUDA: pack weight -> GEMM -> assemble
    Cp = (C + 3) // 4 * 4
    wcat = torch.empty(3 * D, Cp, device=DEV)
    ops.conv3_pack_weight(w.to(DEV), wcat)
    assert torch.equal(wcat.cpu()[:, :C].view(3, D, C), w.permute(2, 0, 1))
    xd = torch.zeros(n * L, Cp, device=DEV)
    xd[:, :C] = x.view(n * L, C).to(DEV)
    z = torch.empty(n * L, 3 * D, device=DEV)
    ops.gemm(xd, wcat, z)
    y = torch.empty(n * L_out, D, device=DEV)
    ops.conv3_assemble_fwd(z, y, n, L, D, pad, bias=bias.to(DEV), pe=pe.to(DEV), wtime=wt.to(DEV))
    assert rel_err(y.cpu().view(n, L_out, D), ref.detach()) < 2e-3
    # backward
    dz = torch.empty(n * L, 3 * D, device=DEV)
    dbias, dwt = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    ops.conv3_assemble_bwd(dy.view(n * L_out, D).to(DEV), dz, n, L, D, pad, dbias=dbias, dwtime=dwt)
    assert rel_err(dbias.cpu(), br.grad) < 1e-5 and rel_err(dwt.cpu(), wtr.grad) < 1e-5
    dx = torch.empty(n * L, Cp, device=DEV)
    ops.gemm(dz, wcat, dx, b_mn=True)
    assert rel_err(dx.cpu()[:, :C].reshape(n, L, C), xr.grad) < 2e-3
    dwcat = torch.zeros(3 * D, Cp, device=DEV)
    ops.gemm(dz, xd, dwcat, a_mn=True, b_mn=True, accumulate=True)
    dw = torch.zeros(D, C, 3, device=DEV)
    ops.conv3_unpack_grad(dwcat, dw)
    assert rel_err(dw.cpu(), wr.grad) < 2e-3


# ------------------------------------------------------------------------------------------------ attention
ATTN_CASES = [
    # B, H, Lq, Lk, dh, factor, mode, layout
    (6, 8, 65, 65, 16, 5, "prob", "blhd"),
    (3, 8, 160, 160, 16, 5, "prob", "blhd"),
    (4, 8, 40, 40, 8, 5, "prob_masked", "blhd"),
    (3, 8, 40, 40, 104, 4, "prob", "bhld"),
    (3, 8, 70, 70, 104, 4, "prob_masked", "bhld"),
    (3, 8, 70, 4, 104, 4, "prob", "bhld"),
    (2, 4, 7, 7, 16, 4, "prob", "bhld"),
    (3, 8, 40, 30, 8, 5, "full", "blhd"),
    # register-blocked small-problem path (dh 8/16, Lk <= 96): 1 / 2 / 3 key slots, Lq != Lk, both layouts
    (4, 8, 40, 40, 8, 5, "prob", "blhd"),
    (3, 4, 40, 70, 16, 5, "prob", "bhld"),
    (3, 8, 90, 33, 8, 4, "prob", "blhd"),
    (6, 2, 20, 96, 16, 5, "prob", "blhd"),
    # shapes the tensor-core forward takes (unmasked, dh 16, Lq == Lk <= 79): gaze encoder, odd length, LP = 80 limit, 2 heads
    (4, 8, 40, 40, 16, 5, "prob", "blhd"),
    (5, 8, 33, 33, 16, 4, "prob", "blhd"),
    (3, 2, 79, 79, 16, 5, "prob", "bhld"),
    (7, 4, 48, 48, 16, 5, "prob", "blhd"),
    # head sizes outside the model's own (8 / 16 / 104): the generic kernels with a run-time head dimension, every mode
    (2, 4, 40, 40, 32, 5, "prob", "blhd"),
    (3, 2, 50, 50, 64, 4, "prob_masked", "bhld"),
    (3, 3, 24, 30, 12, 5, "full", "blhd"),
    (2, 4, 130, 130, 20, 5, "prob", "bhld"),
    (2, 2, 33, 70, 48, 4, "prob", "blhd"),
]


def _tc_eligible(Lq, Lk, dh, mode, H):
    return mode == "prob" and dh == 16 and Lq == Lk and 8 <= Lq <= 79 and H % 2 == 0


@pytest.mark.parametrize("tc", ["0", "1"])  # fp32 FMA kernels / tcgen05 forward (3xTF32 scores, tf32 P.V) where it applies
@pytest.mark.parametrize("B,H,Lq,Lk,dh,factor,mode,layout", ATTN_CASES)
def test_attention_forward_backward(ops, monkeypatch, B, H, Lq, Lk, dh, factor, mode, layout, tc):
    on_tc = tc == "1" and _tc_eligible(Lq, Lk, dh, mode, H)
    if tc == "1" and not on_tc:
        pytest.skip("shape not taken by the tensor-core forward")
    monkeypatch.setenv("RF_ATTN_TC", tc)
    out_tol = 1e-3 if on_tc else 1e-5  # P and V enter the P.V product rounded to tf32 (2^-11), like every GEMM of the model
    gen = g(Lq * 7 + Lk + dh)
    D = H * dh
    # q/k/v live inside fused [rows, 3D] buffers exactly as the modules produce them
    qkv_q = torch.randn(B * Lq, 3 * D, generator=gen)
    qkv_k = qkv_q if Lq == Lk and mode != "full" else torch.randn(B * Lk, 3 * D, generator=gen)
    q = qkv_q[:, :D].reshape(B, Lq, H, dh).clone().requires_grad_()
    k = qkv_k[:, D:2 * D].reshape(B, Lk, H, dh).clone().requires_grad_()
    v = qkv_k[:, 2 * D:].reshape(B, Lk, H, dh).clone().requires_grad_()
    U, u = O.sparse_budget(Lk, factor), O.sparse_budget(Lq, factor)
    groups = 3 if B % 3 == 0 else 1
    idx = torch.randint(Lk, (groups, Lq, U), generator=gen)
    if mode == "full":
        ref = O.full_attention(q, k, v)  # [B,Lq,H,dh]
        tops = None
    else:
        outs, tops, meas = [], [], []
        per = B // groups
        for gi in range(groups):
            sl = slice(gi * per, (gi + 1) * per)
            c, t, m = O.prob_attention(q[sl], k[sl], v[sl], idx[gi], factor, mode == "prob_masked")
            outs.append(c), tops.append(t), meas.append(m)
        ctx, tops, meas = torch.cat(outs), torch.cat(tops), torch.cat(meas)
        ref = ctx.contiguous() if layout == "bhld" else ctx.transpose(1, 2).contiguous()
    dout = torch.randn(ref.shape, generator=gen)
    ref.backward(dout)

    qd, kd = qkv_q.to(DEV), qkv_k.to(DEV)
    qa = (qd, Lq * 3 * D, 3 * D)
    ka = (kd[:, D:], Lk * 3 * D, 3 * D)
    va = (kd[:, 2 * D:], Lk * 3 * D, 3 * D)
    out = torch.full(ref.shape, float("nan"), device=DEV)
    top = torch.zeros(B, H, max(u, 1), dtype=torch.int32, device=DEV)
    measure = torch.empty(B, H, Lq, device=DEV)
    code = {"prob": ops.ATTN_PROB, "prob_masked": ops.ATTN_PROB_MASKED, "full": ops.ATTN_FULL}[mode]
    lay = ops.LAYOUT_BHLD if layout == "bhld" else ops.LAYOUT_BLHD
    idx_d = idx.to(torch.int32).to(DEV)
    ops.attention_fwd(qa, ka, va, B, H, Lq, Lk, dh, code, lay, idx_d, B // groups if groups > 1 else 0, U, u, out, top, measure)
    torch.cuda.synchronize()
    if mode != "full":
        assert rel_err(measure.cpu(), meas) < 1e-5
        assert torch.equal(top.cpu().long().sort(-1).values, tops.sort(-1).values)  # same selected set (index work: exact)
    assert rel_err(out.cpu(), ref.detach()) < out_tol, rel_err(out.cpu(), ref.detach())
    if on_tc:
        print(f"tcgen05 attention forward B={B} H={H} L={Lq}: context rel err {rel_err(out.cpu(), ref.detach()):.2e}, measure rel err "
              f"{rel_err(measure.cpu(), meas):.2e}")
    dqkv_q = torch.zeros_like(qd)
    dqkv_k = dqkv_q if qkv_k is qkv_q else torch.zeros_like(kd)
    ops.attention_bwd(qa, ka, va, B, H, Lq, Lk, dh, code, lay, U, u, top, dout.to(DEV), dqkv_q, dqkv_k[:, D:], dqkv_k[:, 2 * D:])
    torch.cuda.synchronize()
    assert rel_err(dqkv_q.cpu()[:, :D].reshape(B, Lq, H, dh), q.grad) < 2e-5
    assert rel_err(dqkv_k.cpu()[:, D:2 * D].reshape(B, Lk, H, dh), k.grad) < 2e-5
    assert rel_err(dqkv_k.cpu()[:, 2 * D:].reshape(B, Lk, H, dh), v.grad) < 2e-5
    if mode != "full":
        # forced selection in ARBITRARY order (torch.topk(sorted=False) gives no order; Routeformer.forced_tops feeds the reference's
        # picks straight in): forward and backward must not depend on the order of the u indices
        forced = tops.gather(-1, torch.argsort(torch.rand(tops.shape, generator=gen), -1)).to(torch.int32).to(DEV)
        out2 = torch.full(ref.shape, float("nan"), device=DEV)
        top2 = torch.zeros_like(top)
        ops.attention_fwd(qa, ka, va, B, H, Lq, Lk, dh, code, lay, idx_d, B // groups if groups > 1 else 0, U, u, out2, top2, forced_top=forced)
        assert torch.equal(top2, forced)
        assert rel_err(out2.cpu(), ref.detach()) < out_tol
        g_q = torch.zeros_like(qd)
        g_k = g_q if qkv_k is qkv_q else torch.zeros_like(kd)
        ops.attention_bwd(qa, ka, va, B, H, Lq, Lk, dh, code, lay, U, u, top2, dout.to(DEV), g_q, g_k[:, D:], g_k[:, 2 * D:])
        torch.cuda.synchronize()
        assert rel_err(g_q.cpu()[:, :D].reshape(B, Lq, H, dh), q.grad) < 2e-5
        assert rel_err(g_k.cpu()[:, D:2 * D].reshape(B, Lk, H, dh), k.grad) < 2e-5
        assert rel_err(g_k.cpu()[:, 2 * D:].reshape(B, Lk, H, dh), v.grad) < 2e-5


@pytest.mark.parametrize("B,H,L,dh", [(6, 8, 65, 16), (4, 4, 40, 8), (3, 2, 20, 16)])
def test_attention_last_query_only(ops, B, H, L, dh):
    """tail_only: the caller keeps the last query's context only (PerceiveEncoder(out_len=1)); forward writes that row, the
    backward treats every other row of dout as zero.  Both the selected and the mean-filled case occur over the (b, h) problems."""
    gen = g(L + dh)
    factor, D = 5, H * dh
    q, k, v = (torch.randn(B, L, H, dh, generator=gen).requires_grad_() for _ in range(3))
    U = u = O.sparse_budget(L, factor)
    idx = torch.randint(L, (L, U), generator=gen)
    ctx, tops, _ = O.prob_attention(q, k, v, idx, factor, False)
    ref = ctx.transpose(1, 2)  # [B, L, H, dh]
    selected_last = (tops == L - 1).any(-1)
    assert selected_last.any() and (~selected_last).any()  # both code paths are exercised
    dlast = torch.randn(B, H, dh, generator=gen)
    dout = torch.zeros(B, L, H, dh)
    dout[:, L - 1] = dlast
    ref.backward(dout)
    mk = lambda t: (t.detach().reshape(B * L, D).to(DEV), L * D, D)
    out = torch.full((B, L, H, dh), float("nan"), device=DEV)
    top = torch.zeros(B, H, u, dtype=torch.int32, device=DEV)
    ops.attention_fwd(mk(q), mk(k), mk(v), B, H, L, L, dh, ops.ATTN_PROB, ops.LAYOUT_BLHD, idx.int().to(DEV), 0, U, u, out, top,
                      tail_only=True)
    assert torch.equal(top.cpu().long().sort(-1).values, tops.sort(-1).values)
    assert rel_err(out[:, L - 1].cpu(), ref[:, L - 1].detach()) < 1e-5
    dq, dk, dv = (torch.full((B * L, D), float("nan"), device=DEV) for _ in range(3))
    garbage = torch.full((B, L, H, dh), float("nan"))  # rows other than the last must not be read
    garbage[:, L - 1] = dlast
    ops.attention_bwd(mk(q), mk(k), mk(v), B, H, L, L, dh, ops.ATTN_PROB, ops.LAYOUT_BLHD, U, u, top, garbage.to(DEV), dq, dk, dv,
                      tail_only=True)
    assert rel_err(dq.cpu().view(B, L, H, dh), q.grad) < 2e-5
    assert rel_err(dk.cpu().view(B, L, H, dh), k.grad) < 2e-5
    assert rel_err(dv.cpu().view(B, L, H, dh), v.grad) < 2e-5


def test_attention_forced_selection(ops):
    B, H, L, dh, factor = 2, 4, 21, 16, 4
    gen = g(5)
    q, k, v = (torch.randn(B, L, H, dh, generator=gen) for _ in range(3))
    U, u = O.sparse_budget(L, factor), O.sparse_budget(L, factor)
    idx = torch.randint(L, (L, U), generator=gen)
    forced = torch.stack([torch.randperm(L, generator=gen)[:u] for _ in range(B * H)]).view(B, H, u)
    ctx, _, _ = O.prob_attention(q, k, v, idx, factor, False, forced_top=forced)
    out = torch.empty(B, H, L, dh, device=DEV)
    top = torch.empty(B, H, u, dtype=torch.int32, device=DEV)
    D = H * dh
    mk = lambda t: (t.reshape(B * L, D).to(DEV), L * D, D)
    ops.attention_fwd(mk(q), mk(k), mk(v), B, H, L, L, dh, ops.ATTN_PROB, ops.LAYOUT_BHLD, idx.int().to(DEV), 0, U, u, out, top,
                      forced_top=forced.int().to(DEV))
    assert torch.equal(top.cpu().long(), forced)
    assert rel_err(out.cpu(), ctx) < 1e-5


def test_reductions_eval_samples(ops):
    """rf_eval_samples: mean of S stochastic forwards + per-clip loss / ADE / FDE (full_comparison.py:654-679)."""
    gen = g(77)
    S, B, T = 5, 37, 30
    preds = torch.randn(S, B, T, 2, generator=gen) * 3 + 50
    truth = torch.randn(B, T, 2, generator=gen) * 3 + 50
    mean_ref = torch.stack(list(preds)).mean(dim=0)
    for kind, eps in (("smooth_l1", 1.0), ("mse", 0.3), ("mae", 0.3)):
        mean, per_clip = ops.eval_samples(preds.to(DEV), truth.to(DEV), 0.97, eps, kind)
        assert rel_err(mean.cpu(), mean_ref) < 1e-6
        for i in range(B):
            p, t = mean_ref[i:i + 1], truth[i:i + 1]
            ref = torch.stack([O.future_discounted_loss(p, t, 0.97, kind, eps), O.ade(p, t), O.fde(p, t)])
            assert torch.allclose(per_clip[i].cpu(), ref, rtol=2e-5, atol=1e-6), (kind, i)


def test_reductions_dropout(ops):
    """rf_dropout: Philox mask as a pure function of (seed, offset, logical index); out = residual + keep * x / (1 - p)."""
    gen = g(91)
    M, N, p = 517, 130, 0.3
    x = torch.randn(M, N, generator=gen).to(DEV)
    res = torch.randn(M, N, generator=gen).to(DEV)
    y = ops.dropout(x, torch.empty_like(x), p, 1234, 7)
    keep = (y != 0)
    assert abs(keep.float().mean().item() - (1 - p)) < 0.01
    assert torch.allclose(y[keep], x[keep] / (1 - p), rtol=1e-6)
    assert torch.equal(ops.dropout(x, torch.empty_like(x), p, 1234, 7), y)                     # deterministic
    assert not torch.equal(ops.dropout(x, torch.empty_like(x), p, 1234, 8) != 0, keep)         # another call site
    assert not torch.equal(ops.dropout(x, torch.empty_like(x), p, 1235, 7) != 0, keep)         # another seed
    wide = torch.zeros(M, N + 6, device=DEV)
    wide[:, :N] = x
    assert torch.equal(ops.dropout(wide[:, :N], torch.empty_like(x), p, 1234, 7), y)           # independent of the pitch
    assert torch.allclose(ops.dropout(x, torch.empty_like(x), p, 1234, 7, residual=res), y + res)
    z = x.clone()
    ops.dropout(z, z, p, 1234, 7)                                                              # in place
    assert torch.equal(z, y)
    assert torch.equal(ops.dropout(x, torch.empty_like(x), 0.0, 1, 1), x)
    # the backward pass is the same call on the gradient
    dy = torch.randn(M, N, generator=gen).to(DEV)
    assert torch.allclose(ops.dropout(dy, torch.empty_like(dy), p, 1234, 7), dy * keep / (1 - p), rtol=1e-6)


def test_attention_full_probability_dropout(ops):
    """Full attention with dropout on the softmax probabilities (cross_modal_transformer.py:63), forward and backward, against
    autograd with the same mask (= rf_dropout's mask of a [B*H*Lq, Lk] tensor)."""
    gen = g(92)
    B, H, Lq, Lk, dh, p, seed, off = 3, 8, 40, 30, 8, 0.25, 99, 5
    D = H * dh
    q = torch.randn(B, Lq, H, dh, generator=gen, requires_grad=True)
    k = torch.randn(B, Lk, H, dh, generator=gen, requires_grad=True)
    v = torch.randn(B, Lk, H, dh, generator=gen, requires_grad=True)
    ones = torch.ones(B * H * Lq, Lk, device=DEV)
    mask = ops.dropout(ones, torch.empty_like(ones), p, seed, off).cpu().view(B, H, Lq, Lk)
    attn = torch.softmax(torch.einsum("blhe,bshe->bhls", q, k) / math.sqrt(dh), dim=-1) * mask
    ref = torch.einsum("bhls,bshd->blhd", attn, v)
    dout = torch.randn(ref.shape, generator=gen)
    ref.backward(dout)
    mk = lambda t, L: (t.detach().reshape(B * L, D).to(DEV), L * D, D)
    out = torch.empty(B, Lq, H, dh, device=DEV)
    ops.attention_fwd(mk(q, Lq), mk(k, Lk), mk(v, Lk), B, H, Lq, Lk, dh, ops.ATTN_FULL, ops.LAYOUT_BLHD, None, 0, 0, 0, out, None,
                      dropout=(p, seed, off))
    assert rel_err(out.cpu(), ref.detach()) < 1e-5
    dq, dk, dv = (torch.empty(B * L, D, device=DEV) for L in (Lq, Lk, Lk))
    ops.attention_bwd(mk(q, Lq), mk(k, Lk), mk(v, Lk), B, H, Lq, Lk, dh, ops.ATTN_FULL, ops.LAYOUT_BLHD, 0, 0, None, dout.to(DEV),
                      dq, dk, dv, dropout=(p, seed, off))
    assert rel_err(dq.cpu().view(B, Lq, H, dh), q.grad) < 2e-5
    assert rel_err(dk.cpu().view(B, Lk, H, dh), k.grad) < 2e-5
    assert rel_err(dv.cpu().view(B, Lk, H, dh), v.grad) < 2e-5
    with pytest.raises(RuntimeError):  # ProbSparse attention has no probability dropout in the reference
        ops.attention_fwd(mk(q, Lq), mk(q, Lq), mk(q, Lq), B, H, Lq, Lq, dh, ops.ATTN_PROB, ops.LAYOUT_BLHD,
                          torch.zeros(Lq, 20, dtype=torch.int32, device=DEV), 0, 20, 20, out, torch.zeros(B, H, 20, dtype=torch.int32, device=DEV),
                          dropout=(p, seed, off))


# ------------------------------------------------------------------------------------------------ norms
@pytest.mark.parametrize("M,D", [(1000, 128), (77, 64), (300, 832), (33, 30)])
def test_layernorm(ops, M, D):
    gen = g(M + D)
    x = (torch.randn(M, D, generator=gen) * 3 + 1).requires_grad_()
    gamma, beta = (1 + 0.1 * torch.randn(D, generator=gen)).requires_grad_(), torch.randn(D, generator=gen).requires_grad_()
    ref = F.layer_norm(x, (D,), gamma, beta, 1e-5)
    dy = torch.randn(M, D, generator=gen)
    ref.backward(dy)
    y, mean, rstd = torch.empty(M, D, device=DEV), torch.empty(M, device=DEV), torch.empty(M, device=DEV)
    xd, gd, bd = x.detach().to(DEV), gamma.detach().to(DEV), beta.detach().to(DEV)
    ops.layernorm_fwd(xd, gd, bd, y, mean, rstd)
    assert (y.cpu() - ref.detach()).abs().max() < 2e-5
    dx, dg, db = torch.empty(M, D, device=DEV), torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    ops.layernorm_bwd(dy.to(DEV), xd, gd, mean, rstd, dx, dg, db)
    assert rel_err(dx.cpu(), x.grad) < 1e-5 and rel_err(dg.cpu(), gamma.grad) < 1e-5 and rel_err(db.cpu(), beta.grad) < 1e-5


def test_layernorm_random_shapes(ops):
    """Row counts that leave partial warp passes (the rows kernel takes four rows per warp for D <= 128), widths from 1 to
    the kernels' limit of 1 024 including odd ones, single rows."""
    gen = g(31)
    ri = lambda lo, hi: int(torch.randint(lo, hi, (1,), generator=gen))
    failures = []
    for i in range(24):
        M = ri(1, 8) if i % 6 == 0 else ri(1, 700)
        D = [ri(1, 130), ri(1, 130), ri(130, 1025), 4 * ri(1, 33)][i % 4]  # (the kernels take D <= 1024)
        x = (torch.randn(M, D, generator=gen) * 3 + 1).requires_grad_()
        gamma, beta = (1 + 0.1 * torch.randn(D, generator=gen)).requires_grad_(), torch.randn(D, generator=gen).requires_grad_()
        ref = F.layer_norm(x, (D,), gamma, beta, 1e-5)
        dy = torch.randn(M, D, generator=gen)
        ref.backward(dy)
        y, mean, rstd = torch.empty(M, D, device=DEV), torch.empty(M, device=DEV), torch.empty(M, device=DEV)
        xd, gd, bd = x.detach().to(DEV), gamma.detach().to(DEV), beta.detach().to(DEV)
        ops.layernorm_fwd(xd, gd, bd, y, mean, rstd)
        dx, dg, db = torch.empty(M, D, device=DEV), torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
        ops.layernorm_bwd(dy.to(DEV), xd, gd, mean, rstd, dx, dg, db)
        if D == 1:  # one channel: y = beta, dx = 0 exactly; nothing relative to compare
            ok = torch.allclose(y.cpu(), ref.detach(), atol=2e-5) and float(dx.abs().max()) < 1e-4
        else:
            ok = ((y.cpu() - ref.detach()).abs().max() < 5e-5 and rel_err(dx.cpu(), x.grad) < 2e-5 and
                  rel_err(dg.cpu(), gamma.grad) < 2e-5 and rel_err(db.cpu(), beta.grad) < 2e-5)
        if not ok:
            failures.append((i, M, D, float((y.cpu() - ref.detach()).abs().max()), rel_err(dx.cpu(), x.grad), rel_err(dg.cpu(), gamma.grad),
                             rel_err(db.cpu(), beta.grad)))
    assert not failures, failures


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("B,Lz,D", [(4, 42, 64), (3, 23, 832), (5, 7, 32), (2, 6, 32)])
def test_distil_tail(ops, B, Lz, D, training):
    gen = g(B + Lz + D)
    z = torch.randn(B, Lz, D, generator=gen).requires_grad_()
    gamma, beta = (1 + 0.1 * torch.randn(D, generator=gen)).requires_grad_(), (0.1 * torch.randn(D, generator=gen)).requires_grad_()
    rm, rv = 0.1 * torch.randn(D, generator=gen), 0.5 + torch.rand(D, generator=gen)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    zt = F.batch_norm(z.transpose(1, 2), rm_ref, rv_ref, gamma, beta, training, 0.1, 1e-5)
    ref = F.max_pool1d(F.elu(zt), 3, 2, 1).transpose(1, 2)
    Lp = ref.shape[1]
    dout = torch.randn(B, Lp, D, generator=gen)
    ref.backward(dout)
    zd, gd, bd, rmd, rvd = z.detach().to(DEV), gamma.detach().to(DEV), beta.detach().to(DEV), rm.to(DEV), rv.to(DEV)
    mean, rstd = torch.empty(D, device=DEV), torch.empty(D, device=DEV)
    out, arg = torch.empty(B * Lp, D, device=DEV), torch.empty(B * Lp, D, dtype=torch.int8, device=DEV)
    ops.distil_fwd(zd.view(B * Lz, D), B, Lz, D, gd, bd, rmd, rvd, training, mean, rstd, out, arg)
    assert (out.cpu().view(B, Lp, D) - ref.detach()).abs().max() < 2e-5
    assert torch.allclose(rmd.cpu(), rm_ref, atol=1e-5) and torch.allclose(rvd.cpu(), rv_ref, atol=1e-5)
    dz, dg, db = torch.empty(B * Lz, D, device=DEV), torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    ops.distil_bwd(zd.view(B * Lz, D), B, Lz, D, gd, bd, mean, rstd, training, arg, dout.view(B * Lp, D).to(DEV), dz, dg, db,
                   torch.empty(2 * D, device=DEV))
    assert rel_err(dz.cpu().view(B, Lz, D), z.grad) < 2e-4
    assert rel_err(dg.cpu(), gamma.grad) < 1e-4 and rel_err(db.cpu(), beta.grad) < 1e-4


# ------------------------------------------------------------------------------------------------ glue
@pytest.mark.parametrize("rotate,normalize", [(False, False), (True, False), (False, True), (True, True)])
def test_motion_features_and_decode(ops, rotate, normalize):
    gen = g(11)
    B, T, E, P = 7, 40, 64, 30
    gps = torch.cumsum(1.83 + 0.91 * torch.randn(B, T, 2, generator=gen), 1)
    visual = torch.randn(B, T, E, generator=gen)
    mean, std = 1.83, 0.91
    motion = gps[:, 1:] - gps[:, :-1]
    if normalize:
        motion = (motion - mean) / std
    motion = F.pad(motion, (0, 0, 1, 0))
    feats, origin = O.motion_features(motion, rotate)
    ld = 72
    x = torch.full((B, T, ld), float("nan"), device=DEV)
    org = torch.empty(B, device=DEV)
    ops.motion_features(gps.to(DEV), visual.to(DEV), x, org, E, rotate, normalize, mean, std)
    got = x.cpu()
    assert (got[:, :, :5] - feats).abs().max() < 2e-5
    assert torch.equal(got[:, :, 5:5 + E], visual) and torch.equal(got[:, :, 5 + E:], torch.zeros(B, T, ld - 5 - E))
    assert (org.cpu() - origin.view(B)).abs().max() < 1e-6
    # decode
    out = torch.randn(B, P, 68, generator=gen).requires_grad_()
    o = out[:, :, :66]
    if rotate:
        o = torch.cat([O.rotate2d(o[:, :, :2], origin), o[:, :, 2:]], -1)
    mv = o[:, :, :2]
    if normalize:
        mv = mv * std + mean
    wp = gps[:, -1:, :] + torch.cumsum(mv, 1)
    dwp = torch.randn(B, P, 2, generator=gen)
    wp.backward(dwp)
    wpd, mvd = torch.empty(B, P, 2, device=DEV), torch.empty(B, P, 2, device=DEV)
    ops.decode_waypoints_fwd(out.detach().to(DEV), org, gps[:, -1, :].contiguous().to(DEV), wpd, mvd, rotate, normalize, mean, std)
    assert rel_err(wpd.cpu(), wp.detach()) < 1e-6 and rel_err(mvd.cpu(), mv.detach()) < 1e-5
    dout = torch.zeros(B, P, 68, device=DEV)
    ops.decode_waypoints_bwd(dwp.to(DEV), org, dout, rotate, normalize, std)
    assert rel_err(dout.cpu(), out.grad) < 1e-5


@pytest.mark.parametrize("smart", [True, False])
def test_decoder_input(ops, smart):
    gen = g(12)
    B, T, P, ld = 5, 40, 30, 72
    x = torch.randn(B, T, ld, generator=gen).requires_grad_()
    ref = torch.cat([x, x[:, -1:].repeat(1, P, 1) if smart else torch.zeros(B, P, ld)], 1)
    d = torch.randn(B, T + P, ld, generator=gen)
    ref.backward(d)
    xdec = torch.empty(B, T + P, ld, device=DEV)
    ops.decoder_input_fwd(x.detach().to(DEV), xdec, P, smart)
    assert torch.equal(xdec.cpu(), ref.detach())
    dx = torch.ones(B, T, ld, device=DEV)
    ops.decoder_input_bwd(d.to(DEV), dx, P, smart)
    assert rel_err(dx.cpu(), 1 + x.grad) < 1e-6


def test_stream_tokens(ops):
    gen = g(13)
    B, T, E, F_ = 4, 40, 32, 8
    idx = O.frame_indices(T, 5)
    left = torch.randn(B, F_, E, generator=gen).requires_grad_()
    gaze = torch.randn(B, T, E, generator=gen).requires_grad_()
    el, eg, eo = (torch.randn(1, 1, E, generator=gen).requires_grad_() for _ in range(3))
    full = torch.zeros(B, T, E).index_copy(1, idx, left)
    tokens = torch.cat([full + el, gaze + eg, torch.zeros(B, T, E) + eo], 1)
    d = torch.randn(B, 3 * T, E, generator=gen)
    tokens.backward(d)
    tok = torch.full((B, 3 * T, E), float("nan"), device=DEV)
    first, step = int(idx[0]), int(idx[1] - idx[0])
    ops.stream_tokens_fwd(left.detach().to(DEV), F_, first, step, False, el.detach().view(E).to(DEV), tok, B, T, E, 3 * T, 0)
    ops.stream_tokens_fwd(gaze.detach().to(DEV), 0, 0, 1, True, eg.detach().view(E).to(DEV), tok, B, T, E, 3 * T, T)
    ops.stream_tokens_fwd(None, 0, 0, 1, False, eo.detach().view(E).to(DEV), tok, B, T, E, 3 * T, 2 * T)
    assert torch.equal(tok.cpu(), tokens.detach())
    dd = d.to(DEV)
    dleft, dgaze = torch.empty(B, F_, E, device=DEV), torch.empty(B, T, E, device=DEV)
    del_, deg, deo = (torch.zeros(E, device=DEV) for _ in range(3))
    ops.stream_tokens_bwd(dd, dleft, F_, first, step, False, del_, B, T, E, 3 * T, 0)
    ops.stream_tokens_bwd(dd, dgaze, 0, 0, 1, True, deg, B, T, E, 3 * T, T)
    ops.stream_tokens_bwd(dd, None, 0, 0, 1, False, deo, B, T, E, 3 * T, 2 * T)
    assert torch.equal(dleft.cpu(), left.grad) and torch.equal(dgaze.cpu(), gaze.grad)
    for got, ref in ((del_, el), (deg, eg), (deo, eo)):
        assert rel_err(got.cpu(), ref.grad.view(E)) < 1e-5


def test_median_metrics_loss(ops):
    gen = g(14)
    x = torch.randn(6, 1600, 2, generator=gen)
    assert torch.equal(ops.median_downsample(x.to(DEV), 40).cpu(), O.median_downsample(x, 40))
    x2 = torch.randn(3, 83, 2, generator=gen)
    assert torch.equal(ops.median_downsample(x2.to(DEV), 20).cpu(), O.median_downsample(x2, 20))
    assert torch.equal(ops.median_downsample(x2[:, :80].contiguous().to(DEV), 40).cpu(), O.median_downsample(x2[:, :80], 40))
    with pytest.raises(Exception):
        ops.median_downsample(x2.to(DEV), 83)
    p, t = torch.randn(9, 30, 2, generator=gen) * 3, torch.randn(9, 30, 2, generator=gen)
    res, ps = ops.ade_fde(p.to(DEV), t.to(DEV), per_sample=True)
    assert abs(res[0].item() - O.ade(p, t).item()) < 1e-5 and abs(res[1].item() - O.fde(p, t).item()) < 1e-4
    for b in range(9):
        assert abs(ps[b, 0].item() - O.ade(p[b:b + 1], t[b:b + 1]).item()) < 1e-5
        assert abs(ps[b, 1].item() - O.fde(p[b:b + 1], t[b:b + 1]).item()) < 1e-4
    pred, truth = (torch.randn(5, 30, 66, generator=gen) * 2).requires_grad_(), torch.randn(5, 30, 66, generator=gen)
    for kind, eps in (("smooth_l1", 1.0), ("mse", 0.3), ("mae", 0.3)):
        pred.grad = None
        ref = O.future_discounted_loss(pred, truth, 0.97, kind, eps)
        (ref * 1.7).backward()
        loss = ops.discounted_loss_fwd(pred.detach().to(DEV), truth.to(DEV), 0.97, eps, kind)
        assert abs(loss.item() - ref.item()) < 1e-5 * max(1, abs(ref.item()))
        dp = torch.empty(5, 30, 66, device=DEV)
        ops.discounted_loss_bwd(pred.detach().to(DEV), truth.to(DEV), 0.97, eps, kind, torch.tensor([1.7], device=DEV), 1.0, dp)
        assert rel_err(dp.cpu(), pred.grad) < 1e-5


def test_reductions_and_adamw(ops):
    gen = g(15)
    src = torch.randn(3000, 200, generator=gen)
    dst = torch.ones(200, device=DEV)
    ops.colsum_accumulate(src.to(DEV), dst)
    assert rel_err(dst.cpu(), 1 + src.double().sum(0).float()) < 1e-5
    acc = torch.zeros(1, device=DEV)
    ops.sumsq_accumulate(src.to(DEV), acc)
    assert abs(acc.item() / (src.double() ** 2).sum().item() - 1) < 1e-5
    n = 10007
    p0, g0 = torch.randn(n, generator=gen), torch.randn(n, generator=gen) * 5
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref_p], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    pd, m, v = p0.to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 4):
        grad = g0 * step
        ref_p.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([ref_p], 2.5)
        opt.step()
        gn = torch.zeros(1, device=DEV)
        gd = grad.to(DEV)
        ops.sumsq_accumulate(gd, gn)
        ops.adamw_step(pd, gd, m, v, 1e-3, 0.9, 0.999, 1e-8, 1e-2, step, 1.0, gn, 2.5)
        assert rel_err(pd.cpu(), ref_p.detach()) < 1e-5


# ------------------------------------------------------------------------------------------------ N4: dataset-side scaling
def test_area_resize_bit_exact(ops):
    """rf_area_resize_u8 against the golden vectors produced by cv2.resize(INTER_AREA) (tests/golden/area_resize.npz) and against
    the oracle on the reference's frame sizes: BIT-EXACT uint8 (integral scales: box sums; otherwise OpenCV's float32 cell sums
    in OpenCV's order, no fused multiply-add)."""
    import os
    import numpy as np
    from oracle import area_resize as A

    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "area_resize.npz"))
    names = sorted({k.split("/")[0] for k in z.files if "/" in k})
    for name in names:
        x, y, f = z[name + "/x"], z[name + "/y"], float(z[name + "/factor"])
        got = ops.area_resize_u8(torch.from_numpy(x).to(DEV), f).cpu().numpy()
        assert got.shape == y.shape and np.array_equal(got, y), (name, int((got != y).sum()))
    # the reference's own frame sizes (the GoPro views with their row crop expressed as a row range, no copy)
    rng = np.random.default_rng(3)
    for (H, W), f, crop in (((2160, 3840), 0.1, True), ((1080, 1088), 0.3, False), ((1080, 1920), 0.4, True), ((720, 960), 1 / 3.0, False)):
        x = rng.integers(0, 256, size=(2, 3, H, W), dtype=np.uint8)
        rows = (int(0.3 * H), int(0.7 * H)) if crop else None
        want = A.scale_video(np.ascontiguousarray(A.crop_gopro_rows(x)) if crop else x, f)
        got = ops.area_resize_u8(torch.from_numpy(x).to(DEV), f, rows=rows).cpu().numpy()
        assert got.shape == want.shape and np.array_equal(got, want), (H, W, f, int((got != want).sum()))
    with pytest.raises(RuntimeError):
        ops.area_resize_u8(torch.zeros(1, 3, 8, 8, dtype=torch.uint8, device=DEV), out_hw=(16, 16))
    with pytest.raises(TypeError):
        ops.area_resize_u8(torch.zeros(1, 3, 8, 8, device=DEV), 0.5)


def test_area_resize_random_shapes(ops):
    """rf_area_resize_u8 against the oracle (itself equal to cv2 on thousands of random shapes) away from the reference's frame
    sizes: every cell-width class of the word-streamed kernel (4 / 5 / 9 / 10 / 13 bytes) and the byte kernel behind it (rows that
    are not a multiple of 4 bytes wide, cells wider than 13 bytes), integral scale on one axis only, identity on one axis,
    one-pixel outputs, a row range; BIT-EXACT uint8."""
    import numpy as np
    from oracle import area_resize as A

    rng = np.random.default_rng(11)
    cases = []
    for lo, hi in ((1.01, 2.0), (2.0, 3.0), (3.0, 7.0), (7.0, 8.0), (8.0, 11.0), (11.0, 20.0)):   # horizontal scale ranges
        for _ in range(6):
            dW = int(rng.integers(3, 40))
            W = 4 * max(1, int(round(dW * rng.uniform(lo, hi) / 4)))
            dW = min(dW, W)
            H = int(rng.integers(5, 70))
            cases.append((H, W, int(rng.integers(1, H + 1)), dW))
    for _ in range(30):                                                                      # anything goes (byte kernel when W % 4 != 0)
        H, W = int(rng.integers(2, 90)), int(rng.integers(2, 130))
        cases.append((H, W, int(rng.integers(1, H + 1)), int(rng.integers(1, W + 1))))
    for _ in range(12):                                                                      # integral scales: both axes / x only / y only
        iy, ix = int(rng.integers(1, 6)), int(rng.integers(1, 14))
        dH, dW = int(rng.integers(1, 20)), 4 * int(rng.integers(1, 12))
        cases += [(dH * iy, dW * ix, dH, dW), (dH * iy + 3, dW * ix, dH, dW), (dH * iy, dW * ix + 4, dH, dW)]
    cases += [(16, 64, 16, 64), (16, 64, 16, 20), (16, 64, 5, 64), (9, 12, 1, 1), (40, 128, 1, 128), (40, 128, 40, 1)]
    for H, W, dH, dW in cases:
        x = rng.integers(0, 256, size=(2, 2, H, W), dtype=np.uint8)
        want = A.area_resize_u8(x, (dH, dW))
        got = ops.area_resize_u8(torch.from_numpy(x).to(DEV), out_hw=(dH, dW)).cpu().numpy()
        assert got.shape == want.shape and np.array_equal(got, want), (H, W, dH, dW, int((got != want).sum()))
    # a row range (source pointer moved by r0 rows, plane stride unchanged) at a non-integral and an integral scale
    for (H, W, r0, r1, dH, dW) in ((60, 96, 7, 51, 13, 29), (60, 96, 12, 52, 10, 24)):
        x = rng.integers(0, 256, size=(3, H, W), dtype=np.uint8)
        want = A.area_resize_u8(np.ascontiguousarray(x[:, r0:r1]), (dH, dW))
        got = ops.area_resize_u8(torch.from_numpy(x).to(DEV), out_hw=(dH, dW), rows=(r0, r1)).cpu().numpy()
        assert np.array_equal(got, want), (H, W, r0, r1, dH, dW, int((got != want).sum()))


def test_area_resize_then_crop_equals_the_host_pipeline(ops):
    """SURVEY 8(f) N4 end to end: raw uint8 camera frames -> device area-resize -> FoV crop with the loader's fp16(v / 255)
    conversion, against the host pipeline of the reference (cv2.resize INTER_AREA -> astype(float16) / 255 -> crop): bit-equal."""
    import numpy as np
    from oracle import area_resize as A

    rng = np.random.default_rng(5)
    raw = rng.integers(0, 256, size=(3, 3, 360, 362), dtype=np.uint8)            # front camera at a third of its size
    host_small = A.scale_video(raw, 0.3)                                         # what the loader keeps ...
    host_f16 = torch.from_numpy(host_small.astype(np.float16) / 255.0)           # ... and hands to the model
    dev_small = ops.area_resize_u8(torch.from_numpy(raw).to(DEV), 0.3)
    assert np.array_equal(dev_small.cpu().numpy(), host_small)
    centers = torch.tensor([[0.5, 0.5], [0.3, 0.6], [0.7, 0.4]], device=DEV)
    windows = torch.full((3, 2), 0.5, device=DEV)
    spec = O.BackboneSpec()
    a = ops.fov_crop(dev_small, centers, windows, 32, spec.mean, spec.std, patch=8, out_dtype=torch.float16, u8_as_f16=True)
    b = ops.fov_crop(host_f16.to(DEV), centers, windows, 32, spec.mean, spec.std, patch=8, out_dtype=torch.float16)
    assert torch.equal(a, b)
