// LayerNorm forward/backward (warp per row, shuffle reductions) and the Informer distilling-block tail
// BatchNorm1d -> ELU -> MaxPool1d(3,2,1) forward/backward.  See include/routeformer_b200.h (5), (6).
#include "common.cuh"

namespace rf {
namespace norm {

constexpr int WARPS = 8;

// ---------------------------------------------------------------------------------------------
// LayerNorm: one warp per row; the row is read from L1/L2 three times (mean, centred variance, normalise),
// which matches ATen's two-pass statistics closely and keeps registers independent of D.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WARPS * 32)
layernorm_fwd_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float* __restrict__ y, long long ldy, float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int D) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const bool vec = ((D & 3) == 0) && ((ldx & 3) == 0) && ((ldy & 3) == 0);
  for (long long row = static_cast<long long>(blockIdx.x) * WARPS + warp; row < M; row += static_cast<long long>(gridDim.x) * WARPS) {
    const float* xr = x + row * ldx;
    float* yr = y + row * ldy;
    float s = 0.f;
    if (vec) {
      for (int d = lane * 4; d < D; d += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + d);
        s += (v.x + v.y) + (v.z + v.w);
      }
    } else {
      for (int d = lane; d < D; d += 32) s += xr[d];
    }
    const float mean = warp_sum(s) / D;
    float q = 0.f;
    if (vec) {
      for (int d = lane * 4; d < D; d += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + d);
        const float a = v.x - mean, b = v.y - mean, c = v.z - mean, e = v.w - mean;
        q += (a * a + b * b) + (c * c + e * e);
      }
    } else {
      for (int d = lane; d < D; d += 32) { const float a = xr[d] - mean; q += a * a; }
    }
    const float rstd = rsqrtf(warp_sum(q) / D + 1e-5f);
    if (vec) {
      for (int d = lane * 4; d < D; d += 128) {
        const float4 v = *reinterpret_cast<const float4*>(xr + d);
        const float4 g = *reinterpret_cast<const float4*>(gamma + d);
        const float4 b = *reinterpret_cast<const float4*>(beta + d);
        float4 o;
        o.x = (v.x - mean) * rstd * g.x + b.x;
        o.y = (v.y - mean) * rstd * g.y + b.y;
        o.z = (v.z - mean) * rstd * g.z + b.z;
        o.w = (v.w - mean) * rstd * g.w + b.w;
        *reinterpret_cast<float4*>(yr + d) = o;
      }
    } else {
      for (int d = lane; d < D; d += 32) yr[d] = (xr[d] - mean) * rstd * gamma[d] + beta[d];
    }
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
  }
}

// Rows of D <= 128 (every Perceive layer: D = 128): a warp takes R rows per pass, one float4 per lane and row.  All R loads are
// issued before the first use (R x 512 B in flight per warp instead of one row behind a chain of two shuffle reductions), the
// row stays in registers for the three passes, and the R butterfly reductions interleave.  Per-row arithmetic -- lane partial
// (x + y) + (z + w), butterfly, centred second pass -- is exactly that of the generic kernel above: results are bit-identical.
template <int R>
__global__ void __launch_bounds__(WARPS * 32, 3)
layernorm_fwd_rows_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                          float* __restrict__ y, long long ldy, float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int D) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int d = lane * 4;
  const bool valid = d < D;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 g = valid ? *reinterpret_cast<const float4*>(gamma + d) : zero4;
  const float4 b = valid ? *reinterpret_cast<const float4*>(beta + d) : zero4;
  for (long long row0 = (static_cast<long long>(blockIdx.x) * WARPS + warp) * R; row0 < M; row0 += static_cast<long long>(gridDim.x) * WARPS * R) {
    float4 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = (valid && row0 + r < M) ? *reinterpret_cast<const float4*>(x + (row0 + r) * ldx + d) : zero4;
    float s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) s[r] = (v[r].x + v[r].y) + (v[r].z + v[r].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < R; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
    float mean[R], q[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      mean[r] = s[r] / D;
      const float a0 = v[r].x - mean[r], a1 = v[r].y - mean[r], a2 = v[r].z - mean[r], a3 = v[r].w - mean[r];
      q[r] = valid ? (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3) : 0.f;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < R; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float rstd = rsqrtf(q[r] / D + 1e-5f);
      if (valid && row0 + r < M) {
        float4 o4;
        o4.x = (v[r].x - mean[r]) * rstd * g.x + b.x;
        o4.y = (v[r].y - mean[r]) * rstd * g.y + b.y;
        o4.z = (v[r].z - mean[r]) * rstd * g.z + b.z;
        o4.w = (v[r].w - mean[r]) * rstd * g.w + b.w;
        *reinterpret_cast<float4*>(y + (row0 + r) * ldy + d) = o4;
      }
      if (lane == 0 && row0 + r < M) {
        if (mean_out) mean_out[row0 + r] = mean[r];
        if (rstd_out) rstd_out[row0 + r] = rstd;
      }
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma;  dgamma += sum dy*xhat;  dbeta += sum dy.
// A lane owns the same columns (lane, lane+32, ...) for every row its warp processes, so the dgamma/dbeta partials live in
// registers across rows (CPL = columns per lane, template) and leave the CTA as one shared-memory reduction + one global
// atomic per column.
template <int CPL>
__global__ void __launch_bounds__(WARPS * 32)
layernorm_bwd_kernel(const float* __restrict__ dy, long long lddy, const float* __restrict__ x, long long ldx,
                     const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                     float* __restrict__ dx, long long lddx, float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int D) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sh[];  // [2*D]: dgamma partial, dbeta partial
  float* sg = sh;
  float* sb = sh + D;
  for (int d = threadIdx.x; d < 2 * D; d += blockDim.x) sh[d] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  float gm[CPL], ag[CPL], ab[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int d = lane + 32 * c;
    gm[c] = d < D ? gamma[d] : 0.f;
    ag[c] = 0.f;
    ab[c] = 0.f;
  }
  for (long long row = static_cast<long long>(blockIdx.x) * WARPS + warp; row < M; row += static_cast<long long>(gridDim.x) * WARPS) {
    const float* xr = x + row * ldx;
    const float* gr = dy + row * lddy;
    float* dr = dx + row * lddx;
    const float mu = mean[row], rs = rstd[row];
    float go[CPL], xh[CPL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int d = lane + 32 * c;
      go[c] = d < D ? gr[d] : 0.f;
      xh[c] = d < D ? (xr[d] - mu) * rs : 0.f;
      const float g = go[c] * gm[c];
      s1 += g;
      s2 += g * xh[c];
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      const int d = lane + 32 * c;
      if (d < D) dr[d] = rs * (go[c] * gm[c] - s1 - xh[c] * s2);
      ag[c] += go[c] * xh[c];
      ab[c] += go[c];
    }
  }
#pragma unroll
  for (int c = 0; c < CPL; ++c) {
    const int d = lane + 32 * c;
    if (d < D) {
      atomicAdd(sg + d, ag[c]);
      atomicAdd(sb + d, ab[c]);
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + d, sg[d]);
    if (dbeta) atomicAdd(dbeta + d, sb[d]);
  }
}

// ---------------------------------------------------------------------------------------------
// Distil tail
// ---------------------------------------------------------------------------------------------
// per-channel sum and sum of squares over R rows, deterministic: CTA row `by` writes its partials to part[by][d] (sums) and
// part[gy + by][d] (squares); bn_finalize adds them in a fixed order (no atomics: batch statistics feed the forward pass).
__global__ void column_moments_kernel(const float* __restrict__ z, int R, int D, float* __restrict__ part) {
  __shared__ float p1[8][33], p2[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f, b = 0.f;
  if (col < D)
    for (int r = blockIdx.y * 8 + threadIdx.y; r < R; r += gridDim.y * 8) {
      const float v = z[static_cast<long long>(r) * D + col];
      a += v;
      b += v * v;
    }
  p1[threadIdx.y][threadIdx.x] = a;
  p2[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && col < D) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) { s1 += p1[y][threadIdx.x]; s2 += p2[y][threadIdx.x]; }
    part[static_cast<long long>(blockIdx.y) * D + col] = s1;
    part[static_cast<long long>(gridDim.y + blockIdx.y) * D + col] = s2;
  }
}

// (mean, rstd) from the partial moments + running-statistics update (momentum, unbiased variance)
__global__ void bn_finalize_kernel(const float* __restrict__ part, int gy, float* __restrict__ mean, float* __restrict__ rstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, int R, int D, float momentum, float eps) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  float s1 = 0.f, s2 = 0.f;
  for (int y = 0; y < gy; ++y) {
    s1 += part[static_cast<long long>(y) * D + d];
    s2 += part[static_cast<long long>(gy + y) * D + d];
  }
  const float m = s1 / R;
  float var = s2 / R - m * m;
  var = fmaxf(var, 0.f);
  mean[d] = m;
  rstd[d] = rsqrtf(var + eps);
  running_mean[d] = (1.f - momentum) * running_mean[d] + momentum * m;
  const float unbiased = R > 1 ? var * (static_cast<float>(R) / (R - 1)) : var;
  running_var[d] = (1.f - momentum) * running_var[d] + momentum * unbiased;
}
__global__ void bn_eval_stats_kernel(float* __restrict__ mean, float* __restrict__ rstd, const float* __restrict__ running_mean,
                                     const float* __restrict__ running_var, int D, float eps) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  mean[d] = running_mean[d];
  rstd[d] = rsqrtf(running_var[d] + eps);
}

__device__ __forceinline__ float elu1(float y) { return y > 0.f ? y : expm1f(y); }

__global__ void distil_apply_kernel(const float* __restrict__ z, int Lz, int Lp, int D, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ rstd,
                                    float* __restrict__ out, signed char* __restrict__ argmax, long long total) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D);
    const long long r = i / D;
    const int tp = static_cast<int>(r % Lp);
    const long long b = r / Lp;
    const float sc = rstd[d] * gamma[d], sh = beta[d] - mean[d] * sc;
    float best = -INFINITY;
    int off_best = 0;
#pragma unroll
    for (int off = -1; off <= 1; ++off) {
      const int t = 2 * tp + off;
      if (t < 0 || t >= Lz) continue;
      const float e = elu1(z[(b * Lz + t) * D + d] * sc + sh);
      if (e > best) { best = e; off_best = off; }
    }
    out[i] = best;
    argmax[i] = static_cast<signed char>(off_best);
  }
}

// pass 1: gy = pooled-gradient routed to its argmax * elu'(y); stored in dz; column sums of gy and gy*xhat -> scratch
__global__ void distil_bwd_pass1_kernel(const float* __restrict__ z, int Lz, int Lp, int D, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, const float* __restrict__ mean,
                                        const float* __restrict__ rstd, const signed char* __restrict__ argmax,
                                        const float* __restrict__ dout, float* __restrict__ dz, float* __restrict__ scratch, int R) {
  __shared__ float p1[8][33], p2[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f, bsum = 0.f;
  if (col < D) {
    const float mu = mean[col], rs = rstd[col], gm = gamma[col];
    for (int r = blockIdx.y * 8 + threadIdx.y; r < R; r += gridDim.y * 8) {
      const int t = r % Lz;
      const long long b = r / Lz;
      float g = 0.f;
#pragma unroll
      for (int off = -1; off <= 1; ++off) {
        const int tt = t - off;
        if (tt < 0 || (tt & 1)) continue;
        const int tp = tt >> 1;
        if (tp >= Lp) continue;
        const long long o = (b * Lp + tp) * D + col;
        if (argmax[o] == off) g += dout[o];
      }
      const float xh = (z[static_cast<long long>(r) * D + col] - mu) * rs;
      const float y = xh * gm + beta[col];
      const float gy = g * (y > 0.f ? 1.f : expf(y));
      dz[static_cast<long long>(r) * D + col] = gy;
      a += gy;
      bsum += gy * xh;
    }
  }
  p1[threadIdx.y][threadIdx.x] = a;
  p2[threadIdx.y][threadIdx.x] = bsum;
  __syncthreads();
  if (threadIdx.y == 0 && col < D) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) { s1 += p1[y][threadIdx.x]; s2 += p2[y][threadIdx.x]; }
    atomicAdd(scratch + col, s1);
    atomicAdd(scratch + D + col, s2);
  }
}
// pass 2: dz = gamma*rstd*(gy - (S1 + xhat*S2)/R) (training) or gamma*rstd*gy (eval); block 0 also adds dgamma/dbeta
__global__ void distil_bwd_pass2_kernel(const float* __restrict__ z, int D, const float* __restrict__ gamma, const float* __restrict__ mean,
                                        const float* __restrict__ rstd, int training, float* __restrict__ dz,
                                        const float* __restrict__ scratch, float* __restrict__ dgamma, float* __restrict__ dbeta, int R,
                                        long long total) {
  const float invR = 1.f / R;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D);
    const float rs = rstd[d];
    const float gy = dz[i];
    float v = gy;
    if (training) {
      const float xh = (z[i] - mean[d]) * rs;
      v = gy - (scratch[d] + xh * scratch[D + d]) * invR;
    }
    dz[i] = gamma[d] * rs * v;
  }
  if (blockIdx.x == 0)
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      if (dbeta) dbeta[d] += scratch[d];
      if (dgamma) dgamma[d] += scratch[D + d];
    }
}

}  // namespace norm
}  // namespace rf

using namespace rf;
using namespace rf::norm;

extern "C" int rf_layernorm_fwd(const float* x, long long ldx, const float* gamma, const float* beta, float* y, long long ldy,
                                float* mean, float* rstd, int M, int D, void* stream) {
  RF_CHECK_ARG(x && gamma && beta && y && M > 0 && D > 0 && ldx >= D && ldy >= D, "rf_layernorm_fwd: bad arguments");
  const bool al16 = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(gamma) |
                      reinterpret_cast<uintptr_t>(beta)) & 15) == 0;
  static const bool rows_kernel = [] { const char* e = getenv("RF_LN_ROWS"); return !(e && e[0] == '0'); }();
  if (rows_kernel && D <= 128 && (D & 3) == 0 && (ldx & 3) == 0 && (ldy & 3) == 0 && al16 && M >= 4 * WARPS) {
    constexpr int R = 4;
    int grid = ceil_div(M, WARPS * R);
    grid = grid > 148 * 8 ? 148 * 8 : grid;
    RF_CUDA_OK(launch_pdl(layernorm_fwd_rows_kernel<R>, dim3(grid), dim3(WARPS * 32), 0, static_cast<cudaStream_t>(stream), x, ldx, gamma, beta, y,
                          ldy, mean, rstd, M, D));
    return RF_OK;
  }
  int grid = ceil_div(M, WARPS);
  grid = grid > 148 * 8 ? 148 * 8 : grid;
  RF_CUDA_OK(launch_pdl(layernorm_fwd_kernel, dim3(grid), dim3(WARPS * 32), 0, static_cast<cudaStream_t>(stream), x, ldx, gamma, beta, y, ldy, mean, rstd,
                        M, D));
  return RF_OK;
}

extern "C" int rf_layernorm_bwd(const float* dy, long long lddy, const float* x, long long ldx, const float* gamma, const float* mean,
                                const float* rstd, float* dx, long long lddx, float* dgamma, float* dbeta, int M, int D, void* stream) {
  RF_CHECK_ARG(dy && x && gamma && mean && rstd && dx && M > 0 && D > 0 && D <= 1024, "rf_layernorm_bwd: bad arguments (D <= 1024)");
  int grid = ceil_div(M, WARPS * 4);  // >= 4 rows per warp so the per-CTA column atomics amortise
  grid = grid < 1 ? 1 : (grid > 148 * 4 ? 148 * 4 : grid);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t sm = 2 * D * sizeof(float);
  // (a float4 / four-rows-per-pass variant like the forward's was measured: 43.0 vs 44.1 us on [99840, 128], not kept)
  if (D <= 64) RF_CUDA_OK(launch_pdl(layernorm_bwd_kernel<2>, dim3(grid), dim3(WARPS * 32), sm, st, dy, lddy, x, ldx, gamma, mean, rstd, dx, lddx, dgamma, dbeta, M, D));
  else if (D <= 128) RF_CUDA_OK(launch_pdl(layernorm_bwd_kernel<4>, dim3(grid), dim3(WARPS * 32), sm, st, dy, lddy, x, ldx, gamma, mean, rstd, dx, lddx, dgamma, dbeta, M, D));
  else if (D <= 256) RF_CUDA_OK(launch_pdl(layernorm_bwd_kernel<8>, dim3(grid), dim3(WARPS * 32), sm, st, dy, lddy, x, ldx, gamma, mean, rstd, dx, lddx, dgamma, dbeta, M, D));
  else if (D <= 1024) RF_CUDA_OK(launch_pdl(layernorm_bwd_kernel<32>, dim3(grid), dim3(WARPS * 32), sm, st, dy, lddy, x, ldx, gamma, mean, rstd, dx, lddx, dgamma, dbeta, M, D));
  else { set_error("rf_layernorm_bwd: D=%d > 1024 not supported", D); return RF_ERR_UNSUPPORTED; }
  RF_LAUNCH_OK();
  return RF_OK;
}

extern "C" int rf_distil_fwd(const RfDistilParams* p, void* stream) {
  RF_CHECK_ARG(p && p->z && p->gamma && p->beta && p->running_mean && p->running_var && p->mean && p->rstd && p->out && p->argmax,
               "rf_distil_fwd: null pointer");
  RF_CHECK_ARG(p->B > 0 && p->Lz > 0 && p->D > 0, "rf_distil_fwd: bad shape");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int R = p->B * p->Lz, D = p->D;
  const int Lp = (p->Lz - 1) / 2 + 1;
  if (p->training) {
    // the pooled-output buffer doubles as scratch for the partial moments (it is overwritten by the apply kernel afterwards)
    int gy = ceil_div(R, 8 * 16);
    const int cap = (p->B * Lp) / 2;
    gy = gy > 64 ? 64 : gy;
    gy = gy > cap ? cap : gy;
    gy = gy < 1 ? 1 : gy;
    RF_CHECK_ARG(static_cast<long long>(p->B) * Lp >= 2, "rf_distil_fwd: needs at least 2 pooled rows");
    column_moments_kernel<<<dim3(ceil_div(D, 32), gy), dim3(32, 8), 0, s>>>(p->z, R, D, p->out);
    RF_LAUNCH_OK();
    bn_finalize_kernel<<<ceil_div(D, 128), 128, 0, s>>>(p->out, gy, p->mean, p->rstd, p->running_mean, p->running_var, R, D, p->momentum,
                                                       p->eps);
    RF_LAUNCH_OK();
  } else {
    bn_eval_stats_kernel<<<ceil_div(D, 128), 128, 0, s>>>(p->mean, p->rstd, p->running_mean, p->running_var, D, p->eps);
    RF_LAUNCH_OK();
  }
  const long long total = static_cast<long long>(p->B) * Lp * D;
  long long blocks = (total + 255) / 256;
  blocks = blocks > 148 * 16 ? 148 * 16 : blocks;
  distil_apply_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(p->z, p->Lz, Lp, D, p->gamma, p->beta, p->mean, p->rstd, p->out, p->argmax,
                                                               total);
  RF_LAUNCH_OK();
  return RF_OK;
}

extern "C" int rf_distil_bwd(const RfDistilBwdParams* p, void* stream) {
  RF_CHECK_ARG(p && p->z && p->gamma && p->beta && p->mean && p->rstd && p->argmax && p->dout && p->dz && p->scratch, "rf_distil_bwd: null pointer");
  RF_CHECK_ARG(p->B > 0 && p->Lz > 0 && p->D > 0, "rf_distil_bwd: bad shape");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int R = p->B * p->Lz, D = p->D;
  const int Lp = (p->Lz - 1) / 2 + 1;
  RF_CUDA_OK(cudaMemsetAsync(p->scratch, 0, 2 * D * sizeof(float), s));
  int gy = ceil_div(R, 8 * 16);
  gy = gy < 1 ? 1 : (gy > 64 ? 64 : gy);
  distil_bwd_pass1_kernel<<<dim3(ceil_div(D, 32), gy), dim3(32, 8), 0, s>>>(p->z, p->Lz, Lp, D, p->gamma, p->beta, p->mean, p->rstd, p->argmax,
                                                                           p->dout, p->dz, p->scratch, R);
  RF_LAUNCH_OK();
  const long long total = static_cast<long long>(R) * D;
  long long blocks = (total + 255) / 256;
  blocks = blocks > 148 * 16 ? 148 * 16 : blocks;
  distil_bwd_pass2_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(p->z, D, p->gamma, p->mean, p->rstd, p->training, p->dz, p->scratch,
                                                                   p->dgamma, p->dbeta, R, total);
  RF_LAUNCH_OK();
  return RF_OK;
}
