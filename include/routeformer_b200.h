/*
 * routeformer_b200 -- C ABI of the B200-native (sm_100a) Routeformer forward/backward hot path.
 *
 * The reference (meakbiyik/routeformer) is pure Python/PyTorch and has no FFI; its only "plugin API"
 * is the Python module boundary (`Routeformer`, `PerceiveEncoder`, `Informer`, `VideoBackboneModule`).
 * This header is the boundary UNDER that Python API: every entry point replaces the ATen call
 * sequence of one reference function, cited as  // replaces: <file:line>  (paths relative to the
 * reference repository root).  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes only.  All data pointers are DEVICE pointers unless a
 *    field says "host".  fp32 everywhere unless a dtype field says otherwise.
 *  - the caller owns all memory; the library never allocates device memory, never retains a
 *    pointer after return, never synchronises: all work is enqueued on `stream` (a cudaStream_t
 *    passed as void*), so whole forward/backward passes are CUDA-graph capturable.
 *  - every entry returns RF_OK (0) or a negative RF_ERR_* code; rf_last_error() (thread-local)
 *    describes the failure.  Nothing throws, nothing calls exit().
 *  - "ld*" = leading dimension (row pitch) in ELEMENTS.
 */
#ifndef ROUTEFORMER_B200_H_
#define ROUTEFORMER_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define RF_ABI_VERSION 1

#define RF_OK 0
#define RF_ERR_INVALID_ARGUMENT (-1)
#define RF_ERR_CUDA (-2)
#define RF_ERR_UNSUPPORTED (-3)

/* dtypes */
#define RF_F32 0
#define RF_F16 1
#define RF_BF16 2
#define RF_U8 3
#define RF_U8_F16 4 /* crop source only: uint8 pixel converted like the reference loader, fp16(v / 255) (io/dataset.py:1505-1522) */

/* activations */
#define RF_ACT_NONE 0
#define RF_ACT_RELU 1
#define RF_ACT_GELU 2 /* exact erf GELU */
#define RF_ACT_GELU_SAVE_GRAD 3 /* forward only: out = gelu(x) and the `preact` buffer receives gelu'(x) instead of x (the
                                  derivative shares erf / exp with the activation, so the backward pass need not recompute it) */
#define RF_DACT_SAVED 3         /* backward only (`dact`): result *= dact_aux (a derivative saved by RF_ACT_GELU_SAVE_GRAD) */

int rf_abi_version(void);
const char* rf_last_error(void);
/* sizeof() of the parameter structs below, in declaration order (0 = RfFovCropParams ... 7 = RfDistilBwdParams, 8 = RfAreaResizeParams);
 * lets a foreign-language binding verify its struct mirror.  Returns -1 for an unknown index. */
int rf_struct_size(int which);

/* ------------------------------------------------------------------------------------------------
 * (1) Field-of-view crop / resample  (HBM-bound)
 * replaces: routeformer/models/video_backbone/TimmBackbone.py:164-177 (pad-to-square + resize +
 *           normalise) and, in "gaze" mode, adds the gaze-centred window of the north-star.
 * Semantics = F.grid_sample(frames, affine_grid(theta), bilinear, zeros, align_corners=False)
 * with theta = [[fw,0,2cx-1],[0,fh,2cy-1]], then (x - mean[c]) * inv_std[c].
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const void* frames;       /* planar frames, [*, 3, H, W]; frame i starts at frames + frame_ids[i]*3*H*W */
  int src_dtype;            /* RF_F16 | RF_F32 | RF_U8 (scaled by 1/255 in fp32) | RF_U8_F16 (scaled by 1/255 and rounded to fp16:
                               raw uint8 frames staged on the device give bit-identical taps to the reference's host-side conversion) */
  const int* frame_ids;     /* [n_frames] source frame numbers, or NULL for 0..n_frames-1 */
  int n_frames, H, W;
  const float* centers;     /* [n_frames,2] (cx,cy) as fractions of the frame */
  const float* windows;     /* [n_frames,2] (fw,fh) as fractions of the frame */
  float mean[3], inv_std[3];
  int out_size;             /* S: output is S x S */
  int patch;                /* 0: planar [n,3,S,S];  p>0: patch-major rows [(n*G+py)*G+px][(c*p+iy)*p+ix], G=S/p */
  void* out;
  int out_dtype;            /* RF_F32 | RF_F16 | RF_BF16 */
  long long out_ld;         /* row pitch in elements when patch>0 (>= 3*p*p) */
} RfFovCropParams;
int rf_fov_crop(const RfFovCropParams* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (2) Tensor-core GEMM (tcgen05.mma kind::tf32, or kind::f16 for fp16 / bf16 operands; TMA-fed, TMEM accumulators, fp32 accumulate)
 * replaces: every aten::addmm / mm / 1x1 aten::convolution on the path:
 *           cross_modal_transformer.py:177-198,297-299,356-368; SelfAttentionFamily.py:176-194;
 *           TransformerEncoderDecoder.py:12-18,48-50; Embedding.py:32-45; and their autograd.
 *   C[M,N] (+)= epilogue( sum_k A[m,k] * B[n,k] )
 * A is logical [M,K]: a_mn_major=0 -> memory [M][K] (pitch lda);  1 -> memory [K][M] (pitch lda).
 * B is logical [N,K]: b_mn_major=0 -> memory [N][K] (pitch ldb);  1 -> memory [K][N] (pitch ldb).
 * Forward  y = x W^T : A=x (0), B=W (0).   dgrad dx = dy W : A=dy (0), B=W (1).
 * wgrad  dW = dy^T x : A=dy (1), B=x (1), accumulate=1, split over the (long) reduction.
 * All pitches must be multiples of 4 elements and all bases 16-byte aligned (TMA).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* A; long long lda; int a_mn_major;
  const float* B; long long ldb; int b_mn_major;
  float* C; long long ldc;
  int M, N, K;
  const float* bias;                         /* [N] or NULL */
  const float* rowadd; int rowadd_period; long long ld_rowadd; /* + rowadd[m % period][n] (positional table) or NULL */
  const float* residual; long long ld_res;   /* + residual[m][n] or NULL */
  int act;                                   /* RF_ACT_*; applied after bias/rowadd/residual */
  float* preact; long long ld_pre;           /* if non-NULL the pre-activation value is also stored here */
  const float* dact_aux; long long ld_aux; int dact; /* if dact!=0: result *= act'(dact_aux[m][n]) (RELU: aux>0) */
  int accumulate;                            /* 1: C += result (TMA reduce-add in L2, or fp32 atomics for unaligned C); C must be initialised */
  int split_k;                               /* 0 = auto: >1 for accumulate=1 (wgrad), and for few-tile / long-K GEMMs whose epilogue is
                                                linear (bias, rowadd, residual): output zero-filled, split 0 adds the linear terms,
                                                splits meet in fp32 reduce-adds (order not fixed).  -1 = auto, but bit-reproducible:
                                                non-accumulating calls are never split.  >1 = explicit (accumulate=1 only) */
  int out_group_in, out_group_out, out_row_offset; /* if out_group_in>0: row m is stored at (m/gi)*go + m%gi + offset */
  int round_f16;                             /* 1: round the result through fp16 (backbone plugin returns input dtype) */
  int ab_dtype;                              /* RF_F32: A, B are fp32, multiplied as TF32 (kind::tf32).  RF_F16 / RF_BF16: A, B point to
                                                fp16 / bf16 (kind::f16, fp32 accumulate; the reference backbone runs under fp16
                                                autocast, TimmBackbone.py:106-145; bf16 = the north star's "bf16 operands" mode);
                                                K-major operands only, pitches multiples of 8 */
  float* colsum_a;                           /* optional [K]: colsum_a[k] += sum_m A[m][k] (fp32, K-major A, no split-K).  In a dgrad
                                                call A is the output gradient, so this is the bias gradient of the producing
                                                layer, taken from the operand tiles already in shared memory instead of a
                                                second pass over A (short reductions: sums of the TF32-rounded tiles; otherwise a separate
                                                fp32 kernel runs) */
  int c_dtype;                               /* RF_F32 (0, default): C is fp32.  RF_BF16: C points to bf16 and ldc counts bf16 elements
                                                (the result is rounded once, after the epilogue; 16-bit operands only, no accumulate /
                                                preact): lets a producer hand its output to the next GEMM as a bf16 operand */
} RfGemmParams;
int rf_gemm_tf32(const RfGemmParams* p, void* stream);
/* Profiling hook: installs (or clears with NULL) a device buffer of >= 8 u64; CTA (0,0,0) of every later GEMM launch records
 * %globaltimer (ns) at: start, after setup, first operands landed, last MMA issued, accumulator ready, epilogue done, exit. */
int rf_debug_gemm_stamps(unsigned long long* device_buffer);
/* Profiling hook: bottleneck probe of the persistent GEMM kernel (bit mask, see gemm_tf32.cu g_probe).  TIMING ONLY: while a bit
 * is set the kernel skips work and its results are wrong; 0 restores the normal kernel.  tools/gemm_probe.py --stages. */
int rf_debug_gemm_probe(int mode);

/* ------------------------------------------------------------------------------------------------
 * (3) Circular Conv1d(k=3) assembly.  The conv is computed as ONE GEMM Z = X [W_0;W_1;W_2]^T
 * (N = 3*D) followed by this shift-add:
 *   y[s,t,d] = sum_j Z[s,(t-pad+j) mod L, j*D+d] + bias[d] + pe[t,d] + t*wtime[d],  t in [0, L+2*pad-2)
 * replaces: cross_modal_transformer.py:356-368,425 (pad=1,+bias,+pe); Embedding.py:32-45,124
 *           (pad=1,+pe,+t*w_time); TransformerEncoderDecoder.py:12-18,24 (pad=2,+bias).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* z; long long ldz;   /* [n_seq*L, 3*D] */
  float* y; long long ldy;         /* [n_seq*L_out, D] */
  int n_seq, L, D, pad;
  const float* bias;               /* [D] or NULL */
  const float* pe; long long ld_pe;/* [>=L_out, D] or NULL */
  const float* wtime;              /* [D] or NULL */
} RfConv3AssembleParams;
int rf_conv3_assemble_fwd(const RfConv3AssembleParams* p, void* stream);
typedef struct {
  const float* dy; long long ldy;  /* [n_seq*L_out, D] */
  float* dz; long long ldz;        /* [n_seq*L, 3*D] (overwritten) */
  int n_seq, L, D, pad;
  float* dbias;                    /* [D] += , or NULL */
  float* dwtime;                   /* [D] += , or NULL */
} RfConv3AssembleBwdParams;
int rf_conv3_assemble_bwd(const RfConv3AssembleBwdParams* p, void* stream);
/* W [D,C,3] (reference Conv1d layout) <-> Wcat [3*D, ldw] (row j*D+d, col c; cols >= C are zero) */
int rf_conv3_pack_weight(const float* w, float* wcat, int D, int C, long long ldw, void* stream);
int rf_conv3_unpack_grad(const float* dwcat, float* dw, int D, int C, long long ldw, void* stream); /* dw += */

/* ------------------------------------------------------------------------------------------------
 * (4) Fused ProbSparse / full attention
 * replaces: cross_modal_transformer.py:36-69 (FullAttention), :72-166 (ProbAttention, ProbMask :22-33);
 *           SelfAttentionFamily.py:71-165 (Informer ProbAttention, [B,H,L,dh] output layout).
 * One launch does: sampled scores Q.K[idx], sparsity measure M = max - sum/Lk, top-u selection,
 * scaled scores of the selected queries against all keys, (causal) softmax, PV, and the
 * mean(V) / cumsum(V) fill of the unselected queries.
 * Element (b,l,h,e) of q is at q + b*q_bs + l*q_ls + h*dh + e (same for k, v, and the grads).
 * ---------------------------------------------------------------------------------------------- */
#define RF_ATTN_PROB 0
#define RF_ATTN_PROB_MASKED 1
#define RF_ATTN_FULL 2
#define RF_LAYOUT_BLHD 0 /* Perceive*: context [B,L,H,dh] */
#define RF_LAYOUT_BHLD 1 /* Informer : context [B,H,L,dh], later viewed as [B,L,H*dh] */
typedef struct {
  const float* q; long long q_bs; long long q_ls;
  const float* k; long long k_bs; long long k_ls;
  const float* v; long long v_bs; long long v_ls;
  int B, H, Lq, Lk, dh;
  int mode, out_layout;
  const int* idx;          /* [n_groups, Lq, U] sampled key index per (query, j); shared by all heads */
  int idx_group;           /* batch entries [g*idx_group, (g+1)*idx_group) use table g; 0 = one table */
  int U, u;
  float* out;              /* context, dense, layout per out_layout */
  int* top;                /* out [B,H,u] selected queries (unused for FULL) */
  float* measure;          /* optional out [B,H,Lq] (NULL to skip) */
  const int* forced_top;   /* optional in [B,H,u]: use this selection instead (test hook) */
  float dropout_p;         /* RF_ATTN_FULL only: dropout on the softmax probabilities (cross_modal_transformer.py:63), 0 = off */
  unsigned long long dropout_seed, dropout_offset; /* mask = rf_dropout's for a [B*H*Lq, Lk] tensor with the same seed / offset */
  const unsigned long long* dropout_offset_base;   /* optional DEVICE scalar added to dropout_offset (see rf_dropout) */
  int tail_only;           /* hint: the caller consumes only the context of the LAST query of every sequence (PerceiveEncoder with
                              out_len = 1, cross_modal_transformer.py:433).  Forward: rows 0..Lq-2 of `out` may be left unwritten
                              (top / measure are complete).  Backward: rows 0..Lq-2 of `dout` are taken as zero and not read. */
} RfAttnParams;
int rf_attention_fwd(const RfAttnParams* p, void* stream);
/* Profiling hook: installs (or clears with NULL) a device buffer of >= 16 i64; the middle CTA of every later tensor-core
 * attention forward records clock64 at its phase boundaries (setup, tiles landed, operands prepared, scores ready, per head:
 * measure / soft-max / P written / P.V issued, contexts written, exit). */
int rf_debug_attn_stamps(long long* device_buffer);
typedef struct {
  RfAttnParams f;          /* same problem description; f.top is now an INPUT; f.out unused */
  const float* dout;       /* gradient of the context, same layout as f.out */
  float* dq; float* dk; float* dv; /* same strides as q/k/v; fully overwritten */
} RfAttnBwdParams;
int rf_attention_bwd(const RfAttnBwdParams* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (5) LayerNorm (eps 1e-5) forward / backward, rows of D <= 4096
 * replaces: aten::native_layer_norm at cross_modal_transformer.py:283-301,217-233,421,494;
 *           TransformerEncoderDecoder.py:39-52,99-115; Informer.py:68,101
 * ---------------------------------------------------------------------------------------------- */
int rf_layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float* y,
                     long long ldy, float* mean, float* rstd, int M, int D, void* stream);
int rf_layernorm_bwd(const float* dy, long long lddy, const float* x, long long ldx, const float* gamma,
                     const float* mean, const float* rstd, float* dx, long long lddx, float* dgamma,
                     float* dbeta, int M, int D, void* stream); /* dgamma/dbeta += */

/* ------------------------------------------------------------------------------------------------
 * (6) Informer distilling block tail: BatchNorm1d -> ELU -> MaxPool1d(3,2,1) over [B, Lz, D]
 * replaces: TransformerEncoderDecoder.py:19-21,25-28
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* z;            /* [B*Lz, D] dense */
  int B, Lz, D;
  const float* gamma; const float* beta;
  float* running_mean; float* running_var; /* updated in place when training */
  int training; float momentum, eps;
  float* mean; float* rstd;  /* out [D]: statistics used (batch stats if training) */
  float* out;                /* [B*Lp, D], Lp = (Lz - 1)/2 + 1 */
  signed char* argmax;       /* out [B*Lp, D]: offset (-1,0,1) of the pooled element */
} RfDistilParams;
int rf_distil_fwd(const RfDistilParams* p, void* stream);
typedef struct {
  const float* z; int B, Lz, D;
  const float* gamma; const float* beta; const float* mean; const float* rstd; int training;
  const signed char* argmax;
  const float* dout;         /* [B*Lp, D] */
  float* dz;                 /* out [B*Lz, D] */
  float* dgamma; float* dbeta; /* += */
  float* scratch;            /* [2*D] workspace (zeroed by the call) */
} RfDistilBwdParams;
int rf_distil_bwd(const RfDistilBwdParams* p, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (7) Routeformer glue
 * ---------------------------------------------------------------------------------------------- */
/* replaces: routeformer.py:279-292 (first difference, normalise, zero pad) + :209-235 (angle, norm,
 * acceleration, optional rotation, concat with the visual features).  gps [B,T,2] -> x [B,T,ldx]:
 * cols 0..4 motion features, cols 5..5+E-1 visual (copied from `visual` [B,T,E], or zeros if NULL/only_motion),
 * remaining cols (padding up to ldx) zero.  origin [B] receives the per-clip origin angle.
 * input_is_motion=1: `gps` already holds the motion dynamics [B,T,2] (differenced, normalised, zero row first) --
 * the autoregressive re-entry of routeformer.py:176-190. */
int rf_motion_features(const float* gps, const float* visual, long long ld_vis, float* x, long long ldx,
                       float* origin, int B, int T, int E, int rotate, int normalize, float mean, float std,
                       int input_is_motion, void* stream);
/* replaces: gps_backbone/Informer.py:125-149. x [B,T,ld] -> xdec [B,T+P,ld]; smart: repeat last row, else zeros */
int rf_decoder_input_fwd(const float* x, float* xdec, int B, int T, int P, long long ld, int smart, void* stream);
int rf_decoder_input_bwd(const float* dxdec, float* dx, int B, int T, int P, long long ld, int smart, void* stream); /* dx += */
/* replaces: routeformer.py:443-459,524-531,334-345.  Writes one stream of the fused token buffer:
 * tokens[b, t_off + t, :] = (frame_t ? src[b, f(t), :] : dense ? src[b,t,:] : 0) + emb[:],
 * where frame_t: t == first + f*step for f in [0,F).  src/emb may be NULL. */
int rf_stream_tokens_fwd(const float* src, int F, int first, int step, int dense, const float* emb,
                         float* tokens, int B, int T, int E, int tokens_per_clip, int t_off, void* stream);
int rf_stream_tokens_bwd(const float* dtokens, float* dsrc, int F, int first, int step, int dense, float* demb,
                         int B, int T, int E, int tokens_per_clip, int t_off, void* stream); /* dsrc overwritten, demb += */
/* replaces: routeformer.py:246-252 (rotate back) + :350-395 (de-normalise, cumsum).  out [B,P,ld]:
 * first two channels -> waypoints [B,P,2] = last_gps + cumsum(mv); mv also written back to motion [B,P,2]. */
int rf_decode_waypoints_fwd(const float* out, long long ld, const float* origin, const float* last_gps,
                            float* waypoints, float* motion, int B, int P, int rotate, int normalize,
                            float mean, float std, void* stream);
int rf_decode_waypoints_bwd(const float* dwaypoints, const float* origin, float* dout, long long ld, int B,
                            int P, int rotate, int normalize, float std, void* stream); /* writes dout[:,:,0:2] */
/* replaces: utils/filter.py:5-43 (lower median of consecutive windows). x [B,S,C] -> y [B,target,C] */
int rf_median_downsample(const float* x, float* y, int B, int S, int C, int target, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (8) Metrics, loss, reductions, optimiser
 * ---------------------------------------------------------------------------------------------- */
/* replaces: score/error.py:10-51.  pred/truth [B,T,2].  result[0] = ADE (mean L2 over b,t),
 * result[1] = reference "FDE" = Frobenius norm of the LAST batch element's [T,2] error;
 * per_sample (optional, [B,2]) = per-clip (ADE, FDE) as computed by full_comparison.py:667-674. */
int rf_ade_fde(const float* pred, const float* truth, int B, int T, float* result, float* per_sample, void* stream);
/* replaces: nn.Dropout in training mode (cross_modal_transformer.py:224-231,295-299): out = residual + keep * x / (1 - p), keep
 * drawn per element from Philox4x32-10(seed; element index / 4, offset).  x, out: [M,N] with row pitches ldx / ldo (out may alias
 * x); residual optional.  The same call with x = dy and residual = NULL is the backward pass (the mask is a function of
 * seed / offset / logical element index only). */
int rf_dropout(const float* x, long long ldx, const float* residual, long long ldr, float* out, long long ldo, int M, int N,
               float p, unsigned long long seed, unsigned long long offset, const unsigned long long* offset_base, void* stream);
/* offset_base (optional): DEVICE scalar added to `offset` when the kernel runs.  A training step captured in a CUDA graph bakes
 * `offset` (the call site) into the launch; the per-step part of the Philox counter lives in device memory and is advanced by
 * the graph itself, so every replay draws fresh masks. */
/* replaces: experiments/full_comparison.py:654-679 (_eval_step): mean of S stochastic forwards, then per clip the
 * FutureDiscountedLoss, ADE and "FDE" of the [1,T,2] slices.  preds [S,B,T,2] (sample-major), truth [B,T,2];
 * mean_pred (optional, [B,T,2]) = stack(preds).mean(0); per_clip [B,3] = (loss, ade, fde).  kind / gamma / epsilon as below. */
int rf_eval_samples(const float* preds, const float* truth, int S, int B, int T, float gamma, float epsilon, int kind,
                    float* mean_pred, float* per_clip, void* stream);
/* replaces: losses/future_discounted_mse.py:56-95 (smooth_l1 | mse | mae with gamma^t weights).
 * kind: 0 smooth_l1, 1 mse, 2 mae.  loss[0] = mean over B*T*C.  bwd writes dpred = dloss * d loss/d pred. */
int rf_discounted_loss_fwd(const float* pred, long long ldp, const float* truth, long long ldt, int B, int T,
                           int C, float gamma, float epsilon, int kind, float* loss, void* stream);
int rf_discounted_loss_bwd(const float* pred, long long ldp, const float* truth, long long ldt, int B, int T,
                           int C, float gamma, float epsilon, int kind, const float* dloss, float scale,
                           float* dpred, long long lddp, int accumulate, void* stream);
/* dst[n] += sum_m src[m][n]  (bias gradients) */
int rf_colsum_accumulate(const float* src, long long ld, int M, int N, float* dst, void* stream);
/* out[0] += sum x^2 over n elements (global grad-norm for clipping) */
int rf_sumsq_accumulate(const float* x, long long n, float* out, void* stream);
/* Fused AdamW over a flat arena; grads are first scaled by min(1, max_norm / sqrt(*gnorm_sq)) if gnorm_sq != NULL,
 * and by grad_scale (1/world_size after an NCCL sum).  replaces torch.optim.AdamW (full_comparison.py:694-702)
 * + clip_grad_norm (:829-830). */
int rf_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                  const float* gnorm_sq, float max_norm, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (9) Host staging: copies only the frames the model consumes (routeformer.py:415-421: every 5th frame, never frame 0)
 * from a PINNED host video [B, T, frame_bytes] to a compact device buffer [B, n_sel, frame_bytes], one strided async
 * copy per selected frame time.  `src_host` is a HOST pointer, `times` a HOST int array [n_sel].
 * ---------------------------------------------------------------------------------------------- */
int rf_stage_frames_h2d(void* dst_dev, const void* src_host, int B, int T, const int* times, int n_sel,
                        long long frame_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (10) Dataset-side video down-scaling on the device (raw uint8 frames are uploaded once at camera resolution)
 * replaces: routeformer/io/dataset.py:1440-1501 (`_apply_scaling`: cv2.resize(frame, (int(W*f), int(H*f)), INTER_AREA) per
 *           frame) -- bit-exact for uint8 (OpenCV's resizeAreaFast_ / resizeArea_); the row crop of the GoPro views
 *           (:1324-1338, rows [int(0.3 H), int(0.7 H))) is expressed by the caller through `src` / `H` / `src_plane_stride`;
 *           the conversion `astype(float16) / 255` (:1503-1523) is RF_U8_F16 of rf_fov_crop.
 * Planes are independent (cv2 treats the channels of an HWC image independently): src is [n_planes] planes of H rows of W
 * bytes, dst dense [n_planes, dH, dW].
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const unsigned char* src; long long src_plane_stride; long long src_row_pitch; /* bytes */
  int n_planes, H, W;
  unsigned char* dst; int dH, dW;
} RfAreaResizeParams;
int rf_area_resize_u8(const RfAreaResizeParams* p, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ROUTEFORMER_B200_H_ */
