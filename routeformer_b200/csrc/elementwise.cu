// HBM-bound glue kernels of the Routeformer path: circular-conv assembly, weight re-layout, motion features,
// token-stream assembly, waypoint decode, median filter, metrics / loss, column reductions and fused AdamW.
// See include/routeformer_b200.h (3), (7), (8) for semantics and the reference lines each entry replaces.
#include "common.cuh"

namespace rf {
namespace ew {

constexpr int TPB = 256;
inline int blocks_for(long long n, int per_block = TPB, int cap = 148 * 16) {
  long long b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  return static_cast<int>(b < cap ? b : cap);
}

// ---------------------------------------------------------------------------------------------
// (3) circular conv assembly
// ---------------------------------------------------------------------------------------------
__global__ void conv3_assemble_fwd_kernel(const RfConv3AssembleParams a, int L_out, long long total4) {
  pdl_wait();
  pdl_trigger();
  const int D4 = a.D >> 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D4) << 2;
    const long long row = i / D4;
    const int t = static_cast<int>(row % L_out);
    const long long s = row / L_out;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      int src = (t - a.pad + j) % a.L;
      if (src < 0) src += a.L;
      const float4 z = *reinterpret_cast<const float4*>(a.z + (s * a.L + src) * a.ldz + j * a.D + d);
      acc.x += z.x; acc.y += z.y; acc.z += z.z; acc.w += z.w;
    }
    if (a.bias) {
      const float4 b = *reinterpret_cast<const float4*>(a.bias + d);
      acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
    }
    if (a.wtime) {
      const float4 w = *reinterpret_cast<const float4*>(a.wtime + d);
      const float tf = static_cast<float>(t);
      acc.x += tf * w.x; acc.y += tf * w.y; acc.z += tf * w.z; acc.w += tf * w.w;
    }
    if (a.pe) {
      const float4 p = *reinterpret_cast<const float4*>(a.pe + static_cast<long long>(t) * a.ld_pe + d);
      acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    *reinterpret_cast<float4*>(a.y + row * a.ldy + d) = acc;
  }
}

__global__ void conv3_assemble_bwd_kernel(const RfConv3AssembleBwdParams a, int L_out, long long total4) {
  pdl_wait();
  pdl_trigger();
  const int D4 = a.D >> 2;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % D4) << 2;
    long long r = i / D4;
    const int j = static_cast<int>(r % 3);
    r /= 3;
    const int src = static_cast<int>(r % a.L);
    const long long s = r / a.L;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int t = (src + a.pad - j) % a.L;
    if (t < 0) t += a.L;
    for (; t < L_out; t += a.L) {
      const float4 g = *reinterpret_cast<const float4*>(a.dy + (s * L_out + t) * a.ldy + d);
      acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
    }
    *reinterpret_cast<float4*>(a.dz + (s * a.L + src) * a.ldz + j * a.D + d) = acc;
  }
}

// dst[n] += sum_m w(m) * src[m][n];  w(m) = 1 or (m % period) (time-feature weight gradient).
// 32 columns x 8 row-lanes per CTA; every thread keeps 4 independent loads in flight; one atomic per column per CTA.
__global__ void colsum_kernel(const float* __restrict__ src, long long ld, int M, int N, float* __restrict__ dst,
                              int weight_period) {
  pdl_wait();
  pdl_trigger();
  __shared__ float part[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (col < N) {
    const int stride = gridDim.y * 8;
    int m = blockIdx.y * 8 + threadIdx.y;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (weight_period > 0) {
      for (; m < M; m += stride) a0 += src[static_cast<long long>(m) * ld + col] * static_cast<float>(m % weight_period);
    } else {
      for (; m + 3 * stride < M; m += 4 * stride) {
        const float v0 = src[static_cast<long long>(m) * ld + col];
        const float v1 = src[static_cast<long long>(m + stride) * ld + col];
        const float v2 = src[static_cast<long long>(m + 2 * stride) * ld + col];
        const float v3 = src[static_cast<long long>(m + 3 * stride) * ld + col];
        a0 += v0; a1 += v1; a2 += v2; a3 += v3;
      }
      for (; m < M; m += stride) a0 += src[static_cast<long long>(m) * ld + col];
    }
    acc = (a0 + a1) + (a2 + a3);
  }
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) s += part[y][threadIdx.x];
    atomicAdd(dst + col, s);
  }
}
static void launch_colsum(const float* src, long long ld, int M, int N, float* dst, int weight_period, cudaStream_t s) {
  dim3 block(32, 8);
  int gy = ceil_div(M, 8 * 8);  // ~8 rows per thread
  gy = gy < 1 ? 1 : (gy > 1024 ? 1024 : gy);
  dim3 grid(ceil_div(N, 32), gy);
  launch_pdl(colsum_kernel, grid, block, 0, s, src, ld, M, N, dst, weight_period);
}

__global__ void pack_weight_kernel(const float* __restrict__ w, float* __restrict__ wcat, int D, int C, long long ldw) {
  const long long total = 3ll * D * ldw;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % ldw);
    const long long row = i / ldw;
    const int d = static_cast<int>(row % D), j = static_cast<int>(row / D);
    wcat[i] = c < C ? w[(static_cast<long long>(d) * C + c) * 3 + j] : 0.f;
  }
}
__global__ void unpack_grad_kernel(const float* __restrict__ dwcat, float* __restrict__ dw, int D, int C, long long ldw) {
  const long long total = 3ll * D * C;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % 3);
    const long long dc = i / 3;
    const int c = static_cast<int>(dc % C), d = static_cast<int>(dc / C);
    dw[i] += dwcat[(static_cast<long long>(j) * D + d) * ldw + c];
  }
}

// ---------------------------------------------------------------------------------------------
// (7) glue
// ---------------------------------------------------------------------------------------------
__global__ void motion_features_kernel(const float* __restrict__ gps, const float* __restrict__ visual, long long ld_vis,
                                       float* __restrict__ x, long long ldx, float* __restrict__ origin, int T, int E,
                                       int rotate, int normalize, float mean, float inv_std, int input_is_motion) {
  const int b = blockIdx.x;
  const float* g = gps + static_cast<long long>(b) * T * 2;
  float* xb = x + static_cast<long long>(b) * T * ldx;
  __shared__ float s_origin;
  auto motion = [&](int t, float& mx, float& my) {
    if (input_is_motion) {  // already differenced / normalised / zero-padded (autoregressive re-entry, routeformer.py:176-190)
      if (t < 0) { mx = 0.f; my = 0.f; return; }
      mx = g[2 * t]; my = g[2 * t + 1];
      return;
    }
    if (t <= 0) { mx = 0.f; my = 0.f; return; }
    mx = g[2 * t] - g[2 * t - 2];
    my = g[2 * t + 1] - g[2 * t - 1];
    if (normalize) { mx = (mx - mean) * inv_std; my = (my - mean) * inv_std; }
  };
  if (threadIdx.x == 0) {
    float mx, my;
    motion(rotate ? T - 1 : 0, mx, my);
    s_origin = atan2f(my, mx);
    origin[b] = s_origin;
  }
  __syncthreads();
  const float o = s_origin;
  float so, co;
  sincosf(-o, &so, &co);
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    float mx, my, px, py;
    motion(t, mx, my);
    motion(t - 1, px, py);
    const float nrm = sqrtf(mx * mx + my * my);
    const float pnrm = sqrtf(px * px + py * py);
    const float ang = atan2f(my, mx);
    float* row = xb + static_cast<long long>(t) * ldx;
    if (rotate) {
      row[0] = co * mx - so * my;
      row[1] = so * mx + co * my;
    } else {
      row[0] = mx;
      row[1] = my;
    }
    row[2] = (ang - o) / 3.14159265358979323846f;
    row[3] = nrm;
    row[4] = t > 0 ? nrm - pnrm : 0.f;
  }
  const int rest = static_cast<int>(ldx) - 5;
  for (int i = threadIdx.x; i < T * rest; i += blockDim.x) {
    const int t = i / rest, e = i % rest;
    float v = 0.f;
    if (visual && e < E) v = visual[(static_cast<long long>(b) * T + t) * ld_vis + e];
    xb[static_cast<long long>(t) * ldx + 5 + e] = v;
  }
}

__global__ void decoder_input_fwd_kernel(const float* __restrict__ x, float* __restrict__ xdec, int T, int P, long long ld,
                                         int smart, long long total) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % ld);
    const long long r = i / ld;
    const int t = static_cast<int>(r % (T + P));
    const long long b = r / (T + P);
    float v;
    if (t < T) v = x[(b * T + t) * ld + c];
    else v = smart ? x[(b * T + T - 1) * ld + c] : 0.f;
    xdec[i] = v;
  }
}
__global__ void decoder_input_bwd_kernel(const float* __restrict__ dxdec, float* __restrict__ dx, int T, int P, long long ld,
                                         int smart, long long total) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % ld);
    const long long r = i / ld;
    const int t = static_cast<int>(r % T);
    const long long b = r / T;
    float v = dxdec[(b * (T + P) + t) * ld + c];
    if (smart && t == T - 1)
      for (int p = 0; p < P; ++p) v += dxdec[(b * (T + P) + T + p) * ld + c];
    dx[i] += v;
  }
}

__global__ void stream_tokens_fwd_kernel(const float* __restrict__ src, int F, int first, int step, int dense,
                                         const float* __restrict__ emb, float* __restrict__ tokens, int T, int E,
                                         int tokens_per_clip, int t_off, long long total) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(i % E);
    const long long r = i / E;
    const int t = static_cast<int>(r % T);
    const long long b = r / T;
    float v = emb ? emb[e] : 0.f;
    if (src) {
      if (dense) v += src[(b * T + t) * E + e];
      else {
        const int rel = t - first;
        if (rel >= 0 && rel % step == 0 && rel / step < F) v += src[(b * F + rel / step) * E + e];
      }
    }
    tokens[(b * tokens_per_clip + t_off + t) * E + e] = v;
  }
}
__global__ void stream_tokens_bwd_kernel(const float* __restrict__ dtokens, float* __restrict__ dsrc, int F, int first, int step,
                                         int dense, int T, int E, int tokens_per_clip, int t_off, long long total) {
  const int rows = dense ? T : F;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(i % E);
    const long long r = i / E;
    const int f = static_cast<int>(r % rows);
    const long long b = r / rows;
    const int t = dense ? f : first + f * step;
    dsrc[i] = dtokens[(b * tokens_per_clip + t_off + t) * E + e];
  }
}

// demb[e] += sum over clips b and the T rows of one stream of dtokens[b, t_off + t, e]
__global__ void stream_emb_grad_kernel(const float* __restrict__ dtokens, float* __restrict__ demb, int B, int T, int E,
                                       int tokens_per_clip, int t_off) {
  __shared__ float part[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (col < E) {
    const long long rows = static_cast<long long>(B) * T;
    for (long long r = blockIdx.y * 8 + threadIdx.y; r < rows; r += gridDim.y * 8) {
      const long long b = r / T;
      const int t = static_cast<int>(r % T);
      acc += dtokens[(b * tokens_per_clip + t_off + t) * E + col];
    }
  }
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && col < E) {
    float s = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) s += part[y][threadIdx.x];
    atomicAdd(demb + col, s);
  }
}

__global__ void decode_waypoints_fwd_kernel(const float* __restrict__ out, long long ld, const float* __restrict__ origin,
                                            const float* __restrict__ last_gps, float* __restrict__ wp, float* __restrict__ motion,
                                            int B, int P, int rotate, int normalize, float mean, float std) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float so = 0.f, co = 1.f;
  if (rotate) sincosf(origin[b], &so, &co);
  float ax = last_gps[2 * b], ay = last_gps[2 * b + 1];
  float cx = 0.f, cy = 0.f;  // running cumsum, added to last_gps afterwards (as torch: last + cumsum)
  for (int t = 0; t < P; ++t) {
    const float* o = out + (static_cast<long long>(b) * P + t) * ld;
    float mx = o[0], my = o[1];
    if (rotate) {
      const float rx = co * mx - so * my, ry = so * mx + co * my;
      mx = rx; my = ry;
    }
    if (normalize) { mx = mx * std + mean; my = my * std + mean; }
    cx += mx; cy += my;
    const long long w = (static_cast<long long>(b) * P + t) * 2;
    motion[w] = mx; motion[w + 1] = my;
    wp[w] = ax + cx; wp[w + 1] = ay + cy;
  }
}
__global__ void decode_waypoints_bwd_kernel(const float* __restrict__ dwp, const float* __restrict__ origin, float* __restrict__ dout,
                                            long long ld, int B, int P, int rotate, int normalize, float std) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float so = 0.f, co = 1.f;
  if (rotate) sincosf(origin[b], &so, &co);
  float gx = 0.f, gy = 0.f;
  for (int t = P - 1; t >= 0; --t) {
    const long long w = (static_cast<long long>(b) * P + t) * 2;
    gx += dwp[w]; gy += dwp[w + 1];
    float mx = gx, my = gy;
    if (normalize) { mx *= std; my *= std; }
    if (rotate) {  // transpose of the forward rotation
      const float rx = co * mx + so * my, ry = -so * mx + co * my;
      mx = rx; my = ry;
    }
    float* o = dout + (static_cast<long long>(b) * P + t) * ld;
    o[0] = mx; o[1] = my;
  }
}

__global__ void median_kernel(const float* __restrict__ x, float* __restrict__ y, int S, int C, int target, int stride, long long total) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long r = i / C;
    const int w = static_cast<int>(r % target);
    const long long b = r / target;
    const float* base = x + (b * S + static_cast<long long>(w) * stride) * C + c;
    const int k = (stride - 1) >> 1;  // lower median
    float res = base[0];
    for (int p = 0; p < stride; ++p) {
      const float v = base[static_cast<long long>(p) * C];
      int rank = 0;
      for (int q = 0; q < stride; ++q) {
        const float u = base[static_cast<long long>(q) * C];
        rank += (u < v) || (u == v && q < p);
      }
      if (rank == k) res = v;
    }
    y[i] = res;
  }
}

// ---------------------------------------------------------------------------------------------
// (8) metrics, loss, reductions, optimiser
// ---------------------------------------------------------------------------------------------
__global__ void ade_fde_kernel(const float* __restrict__ pred, const float* __restrict__ truth, int B, int T, float* __restrict__ result,
                               float* __restrict__ per_sample) {
  __shared__ float s_ade;
  if (threadIdx.x == 0) s_ade = 0.f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int b = warp; b < B; b += nwarps) {
    float sn = 0.f, sq = 0.f;
    for (int t = lane; t < T; t += 32) {
      const long long i = (static_cast<long long>(b) * T + t) * 2;
      const float ex = pred[i] - truth[i], ey = pred[i + 1] - truth[i + 1];
      const float q = ex * ex + ey * ey;
      sn += sqrtf(q);
      sq += q;
    }
    sn = warp_sum(sn);
    sq = warp_sum(sq);
    if (lane == 0) {
      atomicAdd(&s_ade, sn);
      if (per_sample) { per_sample[2 * b] = sn / T; per_sample[2 * b + 1] = sqrtf(sq); }
      if (b == B - 1) result[1] = sqrtf(sq);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) result[0] = s_ade / (static_cast<float>(B) * T);
}

__device__ __forceinline__ float loss_term(float e, int kind, float eps) {
  if (kind == 0) { const float a = fabsf(e); return a < 1.f ? 0.5f * e * e : a - 0.5f; }
  if (fabsf(e) < eps) e = 0.f;
  return kind == 1 ? e * e : fabsf(e);
}
__device__ __forceinline__ float loss_grad(float e, int kind, float eps) {
  if (kind == 0) return fminf(fmaxf(e, -1.f), 1.f);
  if (fabsf(e) < eps) return 0.f;
  return kind == 1 ? 2.f * e : (e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f));
}
// one warp per clip: mean over the S stochastic forwards, then loss / ADE / FDE of that clip (full_comparison.py:654-679)
__global__ void eval_samples_kernel(const float* __restrict__ preds, const float* __restrict__ truth, int S, int B, int T, float gamma,
                                    float eps, int kind, float* __restrict__ mean_pred, float* __restrict__ per_clip) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const long long stride = static_cast<long long>(B) * T * 2;
  for (int b = blockIdx.x * nwarps + warp; b < B; b += gridDim.x * nwarps) {
    float sl = 0.f, sn = 0.f, sq = 0.f;
    for (int t = lane; t < T; t += 32) {
      const long long i = (static_cast<long long>(b) * T + t) * 2;
      float px = 0.f, py = 0.f;
      for (int s = 0; s < S; ++s) { px += preds[s * stride + i]; py += preds[s * stride + i + 1]; }  // stack(...).mean(0)
      px /= S; py /= S;
      if (mean_pred) { mean_pred[i] = px; mean_pred[i + 1] = py; }
      const float ex = px - truth[i], ey = py - truth[i + 1];
      sl += (loss_term(ex, kind, eps) + loss_term(ey, kind, eps)) * powf(gamma, static_cast<float>(t));
      const float q = ex * ex + ey * ey;
      sn += sqrtf(q);
      sq += q;
    }
    sl = warp_sum(sl); sn = warp_sum(sn); sq = warp_sum(sq);
    if (lane == 0) {
      per_clip[3 * b] = sl / (2.f * T);
      per_clip[3 * b + 1] = sn / T;
      per_clip[3 * b + 2] = sqrtf(sq);
    }
  }
}

__global__ void discounted_loss_fwd_kernel(const float* __restrict__ pred, long long ldp, const float* __restrict__ truth, long long ldt,
                                           int T, int C, float gamma, float eps, int kind, float inv_count, float* __restrict__ loss,
                                           long long total) {
  float acc = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long r = i / C;
    const int t = static_cast<int>(r % T);
    acc += loss_term(pred[r * ldp + c] - truth[r * ldt + c], kind, eps) * powf(gamma, static_cast<float>(t));
  }
  acc = warp_sum(acc);
  __shared__ float part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(loss, v * inv_count);
  }
}
__global__ void discounted_loss_bwd_kernel(const float* __restrict__ pred, long long ldp, const float* __restrict__ truth, long long ldt,
                                           int T, int C, float gamma, float eps, int kind, const float* __restrict__ dloss, float scale,
                                           float* __restrict__ dpred, long long lddp, int accumulate, long long total) {
  const float g0 = (dloss ? dloss[0] : 1.f) * scale;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long r = i / C;
    const int t = static_cast<int>(r % T);
    const float g = g0 * loss_grad(pred[r * ldp + c] - truth[r * ldt + c], kind, eps) * powf(gamma, static_cast<float>(t));
    float* d = dpred + r * lddp + c;
    *d = accumulate ? *d + g : g;
  }
}

// (both kernels below stream flat arenas: 16-byte loads, several of them in flight per thread -- the scalar grid-stride versions
//  ran at 1.8 TB/s (sum of squares) and 2.9 TB/s (AdamW) of the 6.5 TB/s the copy kernel reaches)
__global__ void __launch_bounds__(TPB) sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  float acc = 0.f;
  if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const long long n4 = n >> 2;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    long long i = tid;
    for (; i + 3 * stride < n4; i += 4 * stride) {  // four independent 16 B loads per thread and pass
      const float4 u0 = x4[i], u1 = x4[i + stride], u2 = x4[i + 2 * stride], u3 = x4[i + 3 * stride];
      a0 += (u0.x * u0.x + u0.y * u0.y) + (u0.z * u0.z + u0.w * u0.w);
      a1 += (u1.x * u1.x + u1.y * u1.y) + (u1.z * u1.z + u1.w * u1.w);
      a2 += (u2.x * u2.x + u2.y * u2.y) + (u2.z * u2.z + u2.w * u2.w);
      a3 += (u3.x * u3.x + u3.y * u3.y) + (u3.z * u3.z + u3.w * u3.w);
    }
    for (; i < n4; i += stride) {
      const float4 u0 = x4[i];
      a0 += (u0.x * u0.x + u0.y * u0.y) + (u0.z * u0.z + u0.w * u0.w);
    }
    acc = (a0 + a1) + (a2 + a3);
    for (long long j = (n4 << 2) + tid; j < n; j += stride) acc += x[j] * x[j];
  } else {
    for (long long i = tid; i < n; i += stride) acc += x[i] * x[i];
  }
  acc = warp_sum(acc);
  __shared__ float part[32];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

struct AdamW {
  float lr, b1, b2, eps, wd, bc1, bc2_sqrt, gs;
  __device__ __forceinline__ void update(float& p, float g, float& m, float& v) const {
    const float gi = g * gs;
    float pi = p * (1.f - lr * wd);
    const float mi = b1 * m + (1.f - b1) * gi;
    const float vi = b2 * v + (1.f - b2) * gi * gi;
    m = mi;
    v = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p = pi;
  }
};

__global__ void __launch_bounds__(TPB) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                    long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                                                    float grad_scale, const float* __restrict__ gnorm_sq, float max_norm, int vec) {
  AdamW a;
  a.lr = lr; a.b1 = b1; a.b2 = b2; a.eps = eps; a.wd = wd; a.bc1 = bc1; a.bc2_sqrt = bc2_sqrt; a.gs = grad_scale;
  if (gnorm_sq) {
    const float norm = sqrtf(gnorm_sq[0]) * grad_scale;
    const float clip = max_norm / (norm + 1e-6f);  // torch.nn.utils.clip_grad_norm_
    if (clip < 1.f) a.gs *= clip;
  }
  const long long tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long done = 0;
  if (vec) {  // all four arenas 16-byte aligned: one float4 of each per thread and pass (same per-element arithmetic)
    const long long n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (long long i = tid; i < n4; i += stride) {
      float4 pv = p4[i], mv = m4[i], vv = v4[i];
      const float4 gv = g4[i];
      a.update(pv.x, gv.x, mv.x, vv.x);
      a.update(pv.y, gv.y, mv.y, vv.y);
      a.update(pv.z, gv.z, mv.z, vv.z);
      a.update(pv.w, gv.w, mv.w, vv.w);
      p4[i] = pv; m4[i] = mv; v4[i] = vv;
    }
    done = n4 << 2;
  }
  for (long long i = done + tid; i < n; i += stride) a.update(p[i], g[i], m[i], v[i]);
}

}  // namespace ew
}  // namespace rf

using namespace rf;
using namespace rf::ew;

extern "C" int rf_conv3_assemble_fwd(const RfConv3AssembleParams* p, void* stream) {
  RF_CHECK_ARG(p && p->z && p->y, "rf_conv3_assemble_fwd: null pointer");
  RF_CHECK_ARG(p->n_seq > 0 && p->L > 0 && p->D > 0 && (p->pad == 1 || p->pad == 2), "rf_conv3_assemble_fwd: bad shape");
  RF_CHECK_ARG(p->D % 4 == 0 && p->ldz % 4 == 0 && p->ldy % 4 == 0 && (!p->pe || p->ld_pe % 4 == 0), "rf_conv3_assemble_fwd: D/ld must be multiples of 4");
  const int L_out = p->L + 2 * p->pad - 2;
  const long long total4 = static_cast<long long>(p->n_seq) * L_out * (p->D / 4);
  RF_CUDA_OK(launch_pdl(conv3_assemble_fwd_kernel, dim3(blocks_for(total4)), dim3(TPB), 0, static_cast<cudaStream_t>(stream), *p, L_out, total4));
  return RF_OK;
}

extern "C" int rf_conv3_assemble_bwd(const RfConv3AssembleBwdParams* p, void* stream) {
  RF_CHECK_ARG(p && p->dy && p->dz, "rf_conv3_assemble_bwd: null pointer");
  RF_CHECK_ARG(p->n_seq > 0 && p->L > 0 && p->D > 0 && (p->pad == 1 || p->pad == 2), "rf_conv3_assemble_bwd: bad shape");
  RF_CHECK_ARG(p->D % 4 == 0 && p->ldz % 4 == 0 && p->ldy % 4 == 0, "rf_conv3_assemble_bwd: D/ld must be multiples of 4");
  const int L_out = p->L + 2 * p->pad - 2;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long total4 = static_cast<long long>(p->n_seq) * p->L * 3 * (p->D / 4);
  RF_CUDA_OK(launch_pdl(conv3_assemble_bwd_kernel, dim3(blocks_for(total4)), dim3(TPB), 0, s, *p, L_out, total4));
  const int rows = p->n_seq * L_out;
  if (p->dbias) { launch_colsum(p->dy, p->ldy, rows, p->D, p->dbias, 0, s); RF_LAUNCH_OK(); }
  if (p->dwtime) { launch_colsum(p->dy, p->ldy, rows, p->D, p->dwtime, L_out, s); RF_LAUNCH_OK(); }
  return RF_OK;
}

extern "C" int rf_conv3_pack_weight(const float* w, float* wcat, int D, int C, long long ldw, void* stream) {
  RF_CHECK_ARG(w && wcat && D > 0 && C > 0 && ldw >= C, "rf_conv3_pack_weight: bad arguments");
  pack_weight_kernel<<<blocks_for(3ll * D * ldw), TPB, 0, static_cast<cudaStream_t>(stream)>>>(w, wcat, D, C, ldw);
  RF_LAUNCH_OK();
  return RF_OK;
}
extern "C" int rf_conv3_unpack_grad(const float* dwcat, float* dw, int D, int C, long long ldw, void* stream) {
  RF_CHECK_ARG(dwcat && dw && D > 0 && C > 0 && ldw >= C, "rf_conv3_unpack_grad: bad arguments");
  unpack_grad_kernel<<<blocks_for(3ll * D * C), TPB, 0, static_cast<cudaStream_t>(stream)>>>(dwcat, dw, D, C, ldw);
  RF_LAUNCH_OK();
  return RF_OK;
}

extern "C" int rf_motion_features(const float* gps, const float* visual, long long ld_vis, float* x, long long ldx, float* origin,
                                  int B, int T, int E, int rotate, int normalize, float mean, float std, int input_is_motion,
                                  void* stream) {
  RF_CHECK_ARG(gps && x && origin && B > 0 && T > 0 && ldx >= 5 + (visual ? E : 0), "rf_motion_features: bad arguments");
  motion_features_kernel<<<B, 128, 0, static_cast<cudaStream_t>(stream)>>>(gps, visual, ld_vis, x, ldx, origin, T, E, rotate, normalize,
                                                                          mean, 1.0f / std, input_is_motion);
  RF_LAUNCH_OK();
  return RF_OK;
}

extern "C" int rf_decoder_input_fwd(const float* x, float* xdec, int B, int T, int P, long long ld, int smart, void* stream) {
  RF_CHECK_ARG(x && xdec && B > 0 && T > 0 && P >= 0 && ld > 0, "rf_decoder_input_fwd: bad arguments");
  const long long total = static_cast<long long>(B) * (T + P) * ld;
  decoder_input_fwd_kernel<<<blocks_for(total), TPB, 0, static_cast<cudaStream_t>(stream)>>>(x, xdec, T, P, ld, smart, total);
  RF_LAUNCH_OK();
  return RF_OK;
}
extern "C" int rf_decoder_input_bwd(const float* dxdec, float* dx, int B, int T, int P, long long ld, int smart, void* stream) {
  RF_CHECK_ARG(dxdec && dx && B > 0 && T > 0 && P >= 0 && ld > 0, "rf_decoder_input_bwd: bad arguments");
  const long long total = static_cast<long long>(B) * T * ld;
  decoder_input_bwd_kernel<<<blocks_for(total), TPB, 0, static_cast<cudaStream_t>(stream)>>>(dxdec, dx, T, P, ld, smart, total);
  RF_LAUNCH_OK();
  return RF_OK;
}

extern "C" int rf_stream_tokens_fwd(const float* src, int F, int first, int step, int dense, const float* emb, float* tokens, int B,
                                    int T, int E, int tokens_per_clip, int t_off, void* stream) {
  RF_CHECK_ARG(tokens && B > 0 && T > 0 && E > 0 && t_off >= 0 && t_off + T <= tokens_per_clip, "rf_stream_tokens_fwd: bad arguments");
  RF_CHECK_ARG(!src || dense || (F > 0 && step > 0 && first >= 0 && first + (F - 1) * step < T), "rf_stream_tokens_fwd: frame grid outside [0,T)");
  const long long total = static_cast<long long>(B) * T * E;
  stream_tokens_fwd_kernel<<<blocks_for(total), TPB, 0, static_cast<cudaStream_t>(stream)>>>(src, F, first, step, dense, emb, tokens, T, E,
                                                                                            tokens_per_clip, t_off, total);
  RF_LAUNCH_OK();
  return RF_OK;
}
extern "C" int rf_stream_tokens_bwd(const float* dtokens, float* dsrc, int F, int first, int step, int dense, float* demb, int B, int T,
                                    int E, int tokens_per_clip, int t_off, void* stream) {
  RF_CHECK_ARG(dtokens && B > 0 && T > 0 && E > 0 && t_off >= 0 && t_off + T <= tokens_per_clip, "rf_stream_tokens_bwd: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dsrc) {
    const long long total = static_cast<long long>(B) * (dense ? T : F) * E;
    stream_tokens_bwd_kernel<<<blocks_for(total), TPB, 0, s>>>(dtokens, dsrc, F, first, step, dense, T, E, tokens_per_clip, t_off, total);
    RF_LAUNCH_OK();
  }
  if (demb) {
    dim3 block(32, 8);
    int gy = ceil_div(B * T, 8 * 16);
    gy = gy < 1 ? 1 : (gy > 128 ? 128 : gy);
    stream_emb_grad_kernel<<<dim3(ceil_div(E, 32), gy), block, 0, s>>>(dtokens, demb, B, T, E, tokens_per_clip, t_off);
    RF_LAUNCH_OK();
  }
  return RF_OK;
}

extern "C" int rf_decode_waypoints_fwd(const float* out, long long ld, const float* origin, const float* last_gps, float* waypoints,
                                       float* motion, int B, int P, int rotate, int normalize, float mean, float std, void* stream) {
  RF_CHECK_ARG(out && last_gps && waypoints && motion && B > 0 && P > 0 && ld >= 2 && (!rotate || origin), "rf_decode_waypoints_fwd: bad arguments");
  decode_waypoints_fwd_kernel<<<ceil_div(B, 64), 64, 0, static_cast<cudaStream_t>(stream)>>>(out, ld, origin, last_gps, waypoints, motion, B,
                                                                                           P, rotate, normalize, mean, std);
  RF_LAUNCH_OK();
  return RF_OK;
}
extern "C" int rf_decode_waypoints_bwd(const float* dwaypoints, const float* origin, float* dout, long long ld, int B, int P, int rotate,
                                       int normalize, float std, void* stream) {
  RF_CHECK_ARG(dwaypoints && dout && B > 0 && P > 0 && ld >= 2 && (!rotate || origin), "rf_decode_waypoints_bwd: bad arguments");
  decode_waypoints_bwd_kernel<<<ceil_div(B, 64), 64, 0, static_cast<cudaStream_t>(stream)>>>(dwaypoints, origin, dout, ld, B, P, rotate,
                                                                                           normalize, std);
  RF_LAUNCH_OK();
  return RF_OK;
}

extern "C" int rf_median_downsample(const float* x, float* y, int B, int S, int C, int target, void* stream) {
  RF_CHECK_ARG(x && y && B > 0 && C > 0 && target > 0, "rf_median_downsample: bad arguments");
  RF_CHECK_ARG(target < S, "rf_median_downsample: Target length must be less than the current time steps.");
  const int stride = S / target;
  const long long total = static_cast<long long>(B) * target * C;
  median_kernel<<<blocks_for(total, 64), 64, 0, static_cast<cudaStream_t>(stream)>>>(x, y, S, C, target, stride, total);
  RF_LAUNCH_OK();
  return RF_OK;
}

extern "C" int rf_ade_fde(const float* pred, const float* truth, int B, int T, float* result, float* per_sample, void* stream) {
  RF_CHECK_ARG(pred && truth && result && B > 0 && T > 0, "rf_ade_fde: bad arguments");
  ade_fde_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(pred, truth, B, T, result, per_sample);
  RF_LAUNCH_OK();
  return RF_OK;
}

__global__ void dropout_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ residual, long long ldr,
                               float* __restrict__ out, long long ldo, int M, int N, uint32_t threshold, float scale,
                               unsigned long long seed, unsigned long long offset, const unsigned long long* __restrict__ offset_base) {
  // one thread per group of 4 consecutive LOGICAL elements (row-major over [M,N]): one Philox call, 4 keep decisions
  if (offset_base) offset += __ldg(offset_base);
  const long long total = static_cast<long long>(M) * N;
  const long long groups = (total + 3) >> 2;
  const bool vec = (N & 3) == 0 && ((ldx | ldo | ldr) & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(residual)) & 15) == 0;
  const int n4 = N >> 2;
  for (long long g = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; g < groups; g += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 r = philox4x32_10(seed, static_cast<unsigned long long>(g), offset);
    if (vec) {  // the 4 elements of a group share a row: one 16 B load / store per operand
      const long long row = g / n4;
      const int col = static_cast<int>(g - row * n4) << 2;
      const float4 v = *reinterpret_cast<const float4*>(x + row * ldx + col);
      float4 o = make_float4(r.x >= threshold ? v.x * scale : 0.0f, r.y >= threshold ? v.y * scale : 0.0f,
                             r.z >= threshold ? v.z * scale : 0.0f, r.w >= threshold ? v.w * scale : 0.0f);
      if (residual) {
        const float4 q = *reinterpret_cast<const float4*>(residual + row * ldr + col);
        o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
      }
      *reinterpret_cast<float4*>(out + row * ldo + col) = o;
      continue;
    }
    const uint32_t rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long e = 4 * g + i;
      if (e < total) {
        const long long row = e / N;
        const int col = static_cast<int>(e - row * N);
        float v = rv[i] >= threshold ? x[row * ldx + col] * scale : 0.0f;
        if (residual) v += residual[row * ldr + col];
        out[row * ldo + col] = v;
      }
    }
  }
}

extern "C" int rf_dropout(const float* x, long long ldx, const float* residual, long long ldr, float* out, long long ldo, int M, int N, float p,
                          unsigned long long seed, unsigned long long offset, const unsigned long long* offset_base, void* stream) {
  RF_CHECK_ARG(x && out && M > 0 && N > 0 && ldx >= N && ldo >= N, "rf_dropout: bad arguments");
  RF_CHECK_ARG(p >= 0.0f && p < 1.0f, "rf_dropout: p=%f must be in [0, 1)", p);
  const long long groups = (static_cast<long long>(M) * N + 3) / 4;
  dropout_kernel<<<blocks_for(groups, TPB, 148 * 16), TPB, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ldx, residual, ldr, out, ldo, M, N, dropout_threshold(p), 1.0f / (1.0f - p), seed, offset, offset_base);
  RF_LAUNCH_OK();
  return RF_OK;
}

extern "C" int rf_eval_samples(const float* preds, const float* truth, int S, int B, int T, float gamma, float epsilon, int kind,
                               float* mean_pred, float* per_clip, void* stream) {
  RF_CHECK_ARG(preds && truth && per_clip && S > 0 && B > 0 && T > 0 && kind >= 0 && kind <= 2, "rf_eval_samples: bad arguments");
  eval_samples_kernel<<<min(ceil_div(B, 8), 1184), 256, 0, static_cast<cudaStream_t>(stream)>>>(preds, truth, S, B, T, gamma, epsilon, kind,
                                                                                                   mean_pred, per_clip);
  RF_LAUNCH_OK();
  return RF_OK;
}

extern "C" int rf_discounted_loss_fwd(const float* pred, long long ldp, const float* truth, long long ldt, int B, int T, int C, float gamma,
                                      float epsilon, int kind, float* loss, void* stream) {
  RF_CHECK_ARG(pred && truth && loss && B > 0 && T > 0 && C > 0 && kind >= 0 && kind <= 2, "rf_discounted_loss_fwd: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  RF_CUDA_OK(cudaMemsetAsync(loss, 0, sizeof(float), s));
  const long long total = static_cast<long long>(B) * T * C;
  discounted_loss_fwd_kernel<<<blocks_for(total, TPB, 256), TPB, 0, s>>>(pred, ldp, truth, ldt, T, C, gamma, epsilon, kind,
                                                                         1.0f / static_cast<float>(total), loss, total);
  RF_LAUNCH_OK();
  return RF_OK;
}
extern "C" int rf_discounted_loss_bwd(const float* pred, long long ldp, const float* truth, long long ldt, int B, int T, int C, float gamma,
                                      float epsilon, int kind, const float* dloss, float scale, float* dpred, long long lddp,
                                      int accumulate, void* stream) {
  RF_CHECK_ARG(pred && truth && dpred && B > 0 && T > 0 && C > 0 && kind >= 0 && kind <= 2, "rf_discounted_loss_bwd: bad arguments");
  const long long total = static_cast<long long>(B) * T * C;
  discounted_loss_bwd_kernel<<<blocks_for(total), TPB, 0, static_cast<cudaStream_t>(stream)>>>(
      pred, ldp, truth, ldt, T, C, gamma, epsilon, kind, dloss, scale / static_cast<float>(total), dpred, lddp, accumulate, total);
  RF_LAUNCH_OK();
  return RF_OK;
}

extern "C" int rf_colsum_accumulate(const float* src, long long ld, int M, int N, float* dst, void* stream) {
  RF_CHECK_ARG(src && dst && M > 0 && N > 0 && ld >= N, "rf_colsum_accumulate: bad arguments");
  launch_colsum(src, ld, M, N, dst, 0, static_cast<cudaStream_t>(stream));
  RF_LAUNCH_OK();
  return RF_OK;
}

extern "C" int rf_sumsq_accumulate(const float* x, long long n, float* out, void* stream) {
  RF_CHECK_ARG(x && out && n > 0, "rf_sumsq_accumulate: bad arguments");
  sumsq_kernel<<<blocks_for((n + 15) / 16, TPB, 148 * 8), TPB, 0, static_cast<cudaStream_t>(stream)>>>(x, n, out);
  RF_LAUNCH_OK();
  return RF_OK;
}

extern "C" int rf_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int step, float grad_scale, const float* gnorm_sq, float max_norm,
                             void* stream) {
  RF_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1, "rf_adamw_step: bad arguments");
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2_sqrt = sqrtf(1.0f - powf(beta2, static_cast<float>(step)));
  const int vec = ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
                    reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0;
  adamw_kernel<<<blocks_for((n + 3) / 4, TPB, 148 * 8), TPB, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1,
                                                                                                   beta2, eps, weight_decay, bc1, bc2_sqrt,
                                                                                                   grad_scale, gnorm_sq, max_norm, vec);
  RF_LAUNCH_OK();
  return RF_OK;
}
