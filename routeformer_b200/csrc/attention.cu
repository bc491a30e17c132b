// Fused ProbSparse / full attention, forward and backward.  See include/routeformer_b200.h (4).
//
// One CTA owns one (clip b, head h) problem; Q, K, V of that head are staged once in shared memory and the whole chain
//   sampled scores -> sparsity measure -> top-u selection -> scaled scores of the selected queries
//   -> (causal) softmax -> P.V -> mean(V)/cumsum(V) fill of the unselected queries
// runs out of shared memory with fp32 FMA arithmetic: the selection is a discontinuous function of the scores, so it is
// computed at full fp32 precision rather than on the tensor cores (hard part 2).  The reference materialises
// K_expand[..., index_sample, :] = [B,H,Lq,U,dh] in HBM (426 MB per call at B=64); here nothing but Q/K/V/context (and the
// u selected indices) ever touches HBM.
//
// Problems are tiny (L <= 160, dh in {4..104}), so throughput = instruction efficiency x resident CTAs:
//   * head dim is a template parameter (8, 16, 104; 0 = run-time) -> every dot product is an unrolled float4 loop,
//     rows are padded to dh+4 floats (16 B aligned, conflict-free for 128-bit shared loads);
//   * no integer division by run-time values in inner loops: work is walked as (warp -> row, lane -> column);
//   * scores + softmax of one selected query stay in the registers of one warp (shuffle reductions), P is written once.
#include "common.cuh"

namespace rf {
namespace attn {

constexpr int THREADS = 128;
constexpr int NWARPS = THREADS / 32;
constexpr int MAX_KEYS_PER_LANE = 8;  // Lk <= 256

struct Smem {
  float* q; float* k; float* v;  // [L][pitch]
  float* s;                      // [u][Lk] probabilities
  float* m;                      // [Lq] sparsity measure
  float* acc;                    // [dh] column accumulator (mean fill / its gradient)
  int* top;                      // [u]
  int* sel;                      // [Lq] rank of the query in the selection or -1
  float* end;                    // first float after the forward layout (backward appends dO and dS)
};

template <int DH>
struct Dims {
  int dh_rt;
  __device__ __forceinline__ int dh() const { return DH ? DH : dh_rt; }
  __device__ __forceinline__ int dh4() const { return dh() >> 2; }
  __device__ __forceinline__ int pitch() const { return dh() + 4; }
};

__host__ __device__ __forceinline__ int round4(int x) { return (x + 3) & ~3; }

template <int DH>
__device__ __forceinline__ void load_tile(const Dims<DH> d, float* dst, const float* src, long long ls, int L) {
  const int dh4 = d.dh4(), pitch = d.pitch();
  for (int i = threadIdx.x; i < L * dh4; i += THREADS) {
    const int l = i / dh4, c = i - l * dh4;  // dh4 is a compile-time constant on the fast paths
    *reinterpret_cast<float4*>(dst + l * pitch + 4 * c) = __ldg(reinterpret_cast<const float4*>(src + static_cast<long long>(l) * ls) + c);
  }
}

template <int DH>
__device__ __forceinline__ float dot4(const Dims<DH> d, const float* a, const float* b) {
  float acc0 = 0.f, acc1 = 0.f;
  if (DH != 0) {
#pragma unroll
    for (int c = 0; c < (DH ? DH / 4 : 1); ++c) {
      const float4 x = reinterpret_cast<const float4*>(a)[c], y = reinterpret_cast<const float4*>(b)[c];
      acc0 = fmaf(x.x, y.x, acc0); acc1 = fmaf(x.y, y.y, acc1); acc0 = fmaf(x.z, y.z, acc0); acc1 = fmaf(x.w, y.w, acc1);
    }
  } else {
    const int dh4 = d.dh4();
    for (int c = 0; c < dh4; ++c) {
      const float4 x = reinterpret_cast<const float4*>(a)[c], y = reinterpret_cast<const float4*>(b)[c];
      acc0 = fmaf(x.x, y.x, acc0); acc1 = fmaf(x.y, y.y, acc1); acc0 = fmaf(x.z, y.z, acc0); acc1 = fmaf(x.w, y.w, acc1);
    }
  }
  return acc0 + acc1;
}

__device__ __forceinline__ Smem carve(float* base, int Lq, int Lk, int u, int dh, int pitch) {
  Smem sm;
  sm.q = base;
  sm.k = sm.q + Lq * pitch;
  sm.v = sm.k + Lk * pitch;
  sm.s = sm.v + Lk * pitch;
  sm.m = sm.s + round4(u * Lk);
  sm.acc = sm.m + round4(Lq);
  sm.top = reinterpret_cast<int*>(sm.acc + round4(dh) + THREADS * 4);
  sm.sel = sm.top + round4(u);
  sm.end = reinterpret_cast<float*>(sm.sel + round4(Lq));
  return sm;
}

__device__ __forceinline__ long long out_offset(const RfAttnParams& p, int b, int h, int l) {
  return p.out_layout == RF_LAYOUT_BLHD ? ((static_cast<long long>(b) * p.Lq + l) * p.H + h) * p.dh
                                        : ((static_cast<long long>(b) * p.H + h) * p.Lq + l) * p.dh;
}

// acc[0:dh] = sum over rows l in [0,L) with keep(l) of src[l][0:dh].  All threads participate: thread (g, c) sums rows
// g, g+groups, ... of 4 channels into scratch[g][4c..4c+3]; then one thread per channel adds the groups in a FIXED order, so
// the result is deterministic (no atomics: a 1-ulp difference here could flip a top-u selection in a later layer).
// scratch = acc + round4(dh), THREADS*4 floats.
template <int DH, typename Keep>
__device__ __forceinline__ void column_sum(const Dims<DH> d, const float* src, int L, float* acc, Keep keep) {
  const int dh = d.dh(), dh4 = d.dh4(), pitch = d.pitch();
  float* scratch = acc + round4(dh);
  const int groups = dh4 <= THREADS ? THREADS / dh4 : 1;
  const int g = threadIdx.x / dh4, c = threadIdx.x - g * dh4;
  if (g < groups) {
    float4 part = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int l = g; l < L; l += groups)
      if (keep(l)) {
        const float4 vv = *reinterpret_cast<const float4*>(src + l * pitch + 4 * c);
        part.x += vv.x; part.y += vv.y; part.z += vv.z; part.w += vv.w;
      }
    *reinterpret_cast<float4*>(scratch + (g * dh4 + c) * 4) = part;
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < dh; ch += THREADS) {
    float sum = 0.f;
    for (int gg = 0; gg < groups; ++gg) sum += scratch[gg * dh + ch];
    acc[ch] = sum;
  }
  __syncthreads();
}

// Selection (or its replay) followed by the probability rows of the selected queries: P[r][:] = softmax(scale * Q[top_r] K^T).
template <int DH>
__device__ void select_and_softmax(const RfAttnParams& p, const Dims<DH> d, const Smem& sm, int b, int u, bool compute_selection,
                                   const int* top_in, float* measure_out) {
  const int Lq = p.Lq, Lk = p.Lk, pitch = d.pitch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.mode == RF_ATTN_FULL) {
    for (int i = threadIdx.x; i < Lq; i += THREADS) { sm.top[i] = i; sm.sel[i] = i; }
  } else if (compute_selection && !p.forced_top) {
    const int group = p.idx_group > 0 ? b / p.idx_group : 0;
    const int* idx = p.idx + static_cast<long long>(group) * Lq * p.U;
    // the [Lq, U] table is copied once, coalesced, into the (still unused) probability buffer: u*Lk >= ... is not guaranteed,
    // so it is only staged when it fits; otherwise it is read through L1
    int* sidx = reinterpret_cast<int*>(sm.s);
    const bool staged = Lq * p.U <= round4(u * Lk);
    if (staged) {
      for (int i = threadIdx.x; i < Lq * p.U; i += THREADS) sidx[i] = __ldg(idx + i);
      __syncthreads();
    }
    for (int i = threadIdx.x; i < Lq; i += THREADS) {
      const float* qi = sm.q + i * pitch;
      const int* row = (staged ? sidx : idx) + i * p.U;
      float mx = -INFINITY, sum = 0.f;
#pragma unroll 4
      for (int j = 0; j < p.U; ++j) {
        const float s = dot4(d, qi, sm.k + row[j] * pitch);
        mx = fmaxf(mx, s);
        sum += s;
      }
      const float mval = mx - sum / Lk;
      sm.m[i] = mval;
      if (measure_out) measure_out[i] = mval;
    }
    __syncthreads();
    // rank-based top-u: rank = number of queries that beat this one (ties -> lower index first)
    for (int i = threadIdx.x; i < Lq; i += THREADS) {
      const float mi = sm.m[i];
      int rank = 0;
#pragma unroll 8
      for (int j = 0; j < Lq; ++j) {
        const float mj = sm.m[j];
        rank += (mj > mi) || (mj == mi && j < i);
      }
      if (rank < u) { sm.top[rank] = i; sm.sel[i] = rank; }
      else sm.sel[i] = -1;
    }
  } else {
    for (int i = threadIdx.x; i < Lq; i += THREADS) sm.sel[i] = -1;
    __syncthreads();
    for (int r = threadIdx.x; r < u; r += THREADS) {
      const int i = top_in[r];
      sm.top[r] = i;
      sm.sel[i] = r;
    }
  }
  __syncthreads();
  // one warp per selected query: scores in registers (lane <-> keys lane, lane+32, ...), shuffle max / sum, P written once
  const float scale = rsqrtf(static_cast<float>(d.dh()));
  for (int r = warp; r < u; r += NWARPS) {
    const int qi = sm.top[r];
    const float* qrow = sm.q + qi * pitch;
    float sc[MAX_KEYS_PER_LANE];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < MAX_KEYS_PER_LANE; ++t) {
      if (32 * t >= Lk) break;  // warp-uniform: no work (and no issue slots) for key slots beyond Lk
      const int j = lane + 32 * t;
      float s = -INFINITY;
      if (j < Lk && !(p.mode == RF_ATTN_PROB_MASKED && j > qi)) s = dot4(d, qrow, sm.k + j * pitch) * scale;
      sc[t] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < MAX_KEYS_PER_LANE; ++t) {
      if (32 * t >= Lk) break;  // warp-uniform: no work (and no issue slots) for key slots beyond Lk
      const float e = (lane + 32 * t < Lk) ? expf(sc[t] - mx) : 0.f;
      sc[t] = e;
      sum += e;
    }
    const float inv = 1.f / warp_sum(sum);
#pragma unroll
    for (int t = 0; t < MAX_KEYS_PER_LANE; ++t) {
      if (32 * t >= Lk) break;  // warp-uniform: no work (and no issue slots) for key slots beyond Lk
      const int j = lane + 32 * t;
      if (j < Lk) sm.s[r * Lk + j] = sc[t] * inv;
    }
  }
  __syncthreads();
}

template <int DH>
__global__ void __launch_bounds__(THREADS) attention_fwd_kernel(const RfAttnParams p) {
  extern __shared__ __align__(16) float smem_f[];
  const Dims<DH> d{p.dh};
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int Lq = p.Lq, Lk = p.Lk, dh = d.dh(), dh4 = d.dh4(), pitch = d.pitch();
  const int u = p.mode == RF_ATTN_FULL ? Lq : p.u;
  Smem sm = carve(smem_f, Lq, Lk, u, dh, pitch);
  load_tile(d, sm.q, p.q + b * p.q_bs + h * dh, p.q_ls, Lq);
  load_tile(d, sm.k, p.k + b * p.k_bs + h * dh, p.k_ls, Lk);
  load_tile(d, sm.v, p.v + b * p.v_bs + h * dh, p.v_ls, Lk);
  __syncthreads();
  const long long bh = static_cast<long long>(b) * p.H + h;
  select_and_softmax(p, d, sm, b, u, true, p.forced_top ? p.forced_top + bh * u : nullptr, p.measure ? p.measure + bh * Lq : nullptr);
  if (p.mode != RF_ATTN_FULL && p.top)
    for (int r = threadIdx.x; r < u; r += THREADS) p.top[bh * u + r] = sm.top[r];

  // selected queries: context[top_r] = P[r] . V, one thread per (r, 4 channels)
  for (int i = threadIdx.x; i < u * dh4; i += THREADS) {
    const int r = i / dh4, c = i - r * dh4;
    const float* prow = sm.s + r * Lk;
    const float* vcol = sm.v + 4 * c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int j = 0; j < Lk; ++j) {
      const float w = prow[j];
      const float4 vv = *reinterpret_cast<const float4*>(vcol + j * pitch);
      acc.x = fmaf(w, vv.x, acc.x); acc.y = fmaf(w, vv.y, acc.y); acc.z = fmaf(w, vv.z, acc.z); acc.w = fmaf(w, vv.w, acc.w);
    }
    *reinterpret_cast<float4*>(p.out + out_offset(p, b, h, sm.top[r]) + 4 * c) = acc;
  }
  // unselected queries: mean(V) (unmasked) or cumsum(V) (masked)
  if (p.mode == RF_ATTN_PROB) {
    column_sum(d, sm.v, Lk, sm.acc, [](int) { return true; });
    const float inv = 1.f / Lk;
    for (int i = threadIdx.x; i < Lq * dh4; i += THREADS) {
      const int l = i / dh4, c = i - l * dh4;
      if (sm.sel[l] < 0) {
        const float4 a4 = *reinterpret_cast<const float4*>(sm.acc + 4 * c);
        *reinterpret_cast<float4*>(p.out + out_offset(p, b, h, l) + 4 * c) = make_float4(a4.x * inv, a4.y * inv, a4.z * inv, a4.w * inv);
      }
    }
  } else if (p.mode == RF_ATTN_PROB_MASKED) {
    for (int c = threadIdx.x; c < dh; c += THREADS) {
      float acc = 0.f;
      for (int l = 0; l < Lq; ++l) {
        acc += sm.v[l * pitch + c];
        if (sm.sel[l] < 0) p.out[out_offset(p, b, h, l) + c] = acc;
      }
    }
  }
}

// Backward: recomputes P from Q, K and the saved selection, then
//   dV[j]  = sum_r P[r][j] dO[top_r]  +  fill-path gradient (mean: sum of unselected dO / Lk; cumsum: reverse cumsum of unselected dO)
//   dS     = P o (dP - rowsum(P o dP)) * scale,  dP[r][j] = dO[top_r] . V[j]
//   dQ[top_r] = sum_j dS[r][j] K[j]   (other rows 0),   dK[j] = sum_r dS[r][j] Q[top_r]
template <int DH>
__global__ void __launch_bounds__(THREADS) attention_bwd_kernel(const RfAttnBwdParams bp) {
  extern __shared__ __align__(16) float smem_f[];
  const RfAttnParams& p = bp.f;
  const Dims<DH> d{p.dh};
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int Lq = p.Lq, Lk = p.Lk, dh = d.dh(), dh4 = d.dh4(), pitch = d.pitch();
  const int u = p.mode == RF_ATTN_FULL ? Lq : p.u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Smem sm = carve(smem_f, Lq, Lk, u, dh, pitch);
  float* s_do = sm.end;               // [Lq][pitch] context gradient of this head
  float* s_ds = s_do + Lq * pitch;    // [u][Lk]
  load_tile(d, sm.q, p.q + b * p.q_bs + h * dh, p.q_ls, Lq);
  load_tile(d, sm.k, p.k + b * p.k_bs + h * dh, p.k_ls, Lk);
  load_tile(d, sm.v, p.v + b * p.v_bs + h * dh, p.v_ls, Lk);
  for (int i = threadIdx.x; i < Lq * dh4; i += THREADS) {
    const int l = i / dh4, c = i - l * dh4;
    *reinterpret_cast<float4*>(s_do + l * pitch + 4 * c) = __ldg(reinterpret_cast<const float4*>(bp.dout + out_offset(p, b, h, l)) + c);
  }
  __syncthreads();
  const long long bh = static_cast<long long>(b) * p.H + h;
  select_and_softmax(p, d, sm, b, u, false, p.mode == RF_ATTN_FULL ? nullptr : p.top + bh * u, nullptr);

  // dP -> dS, one warp per selected query (row sums by shuffle)
  const float scale = rsqrtf(static_cast<float>(dh));
  for (int r = warp; r < u; r += NWARPS) {
    const float* dorow = s_do + sm.top[r] * pitch;
    float dp[MAX_KEYS_PER_LANE];
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < MAX_KEYS_PER_LANE; ++t) {
      if (32 * t >= Lk) break;  // warp-uniform: no work (and no issue slots) for key slots beyond Lk
      const int j = lane + 32 * t;
      float v = 0.f;
      if (j < Lk) {
        v = dot4(d, dorow, sm.v + j * pitch);
        acc = fmaf(sm.s[r * Lk + j], v, acc);
      }
      dp[t] = v;
    }
    acc = warp_sum(acc);
#pragma unroll
    for (int t = 0; t < MAX_KEYS_PER_LANE; ++t) {
      if (32 * t >= Lk) break;  // warp-uniform: no work (and no issue slots) for key slots beyond Lk
      const int j = lane + 32 * t;
      if (j < Lk) s_ds[r * Lk + j] = sm.s[r * Lk + j] * (dp[t] - acc) * scale;
    }
  }
  // fill-path gradient of the unmasked mean: sum of the unselected context gradients (scaled by 1/Lk below)
  if (p.mode == RF_ATTN_PROB) column_sum(d, s_do, Lq, sm.acc, [&](int l) { return sm.sel[l] < 0; });
  else __syncthreads();

  // dQ: selected rows get sum_j dS K, the others zero
  float* dq = bp.dq + b * p.q_bs + h * dh;
  for (int i = threadIdx.x; i < Lq * dh4; i += THREADS) {
    const int l = i / dh4, c = i - l * dh4;
    const int r = sm.sel[l];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= 0) {
      const float* row = s_ds + r * Lk;
#pragma unroll 4
      for (int j = 0; j < Lk; ++j) {
        const float w = row[j];
        const float4 kk = *reinterpret_cast<const float4*>(sm.k + j * pitch + 4 * c);
        acc.x = fmaf(w, kk.x, acc.x); acc.y = fmaf(w, kk.y, acc.y); acc.z = fmaf(w, kk.z, acc.z); acc.w = fmaf(w, kk.w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(dq + static_cast<long long>(l) * p.q_ls + 4 * c) = acc;
  }
  // dK and dV: one thread per (key j, 4 channels), loop over the selected rows
  float* dk = bp.dk + b * p.k_bs + h * dh;
  float* dv = bp.dv + b * p.v_bs + h * dh;
  const float inv_lk = 1.f / Lk;
  for (int i = threadIdx.x; i < Lk * dh4; i += THREADS) {
    const int j = i / dh4, c = i - j * dh4;
    float4 ak = make_float4(0.f, 0.f, 0.f, 0.f), av = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int r = 0; r < u; ++r) {
      const int qi = sm.top[r];
      const float ws = s_ds[r * Lk + j], wp = sm.s[r * Lk + j];
      const float4 qq = *reinterpret_cast<const float4*>(sm.q + qi * pitch + 4 * c);
      const float4 oo = *reinterpret_cast<const float4*>(s_do + qi * pitch + 4 * c);
      ak.x = fmaf(ws, qq.x, ak.x); ak.y = fmaf(ws, qq.y, ak.y); ak.z = fmaf(ws, qq.z, ak.z); ak.w = fmaf(ws, qq.w, ak.w);
      av.x = fmaf(wp, oo.x, av.x); av.y = fmaf(wp, oo.y, av.y); av.z = fmaf(wp, oo.z, av.z); av.w = fmaf(wp, oo.w, av.w);
    }
    if (p.mode == RF_ATTN_PROB) {
      const float4 f4 = *reinterpret_cast<const float4*>(sm.acc + 4 * c);
      av.x += f4.x * inv_lk; av.y += f4.y * inv_lk; av.z += f4.z * inv_lk; av.w += f4.w * inv_lk;
    }
    *reinterpret_cast<float4*>(dk + static_cast<long long>(j) * p.k_ls + 4 * c) = ak;
    *reinterpret_cast<float4*>(dv + static_cast<long long>(j) * p.v_ls + 4 * c) = av;
  }
  if (p.mode == RF_ATTN_PROB_MASKED) {
    // cumsum fill: dV[j] += sum over unselected l >= j of dO[l]  (reverse running sum, one thread per channel)
    __syncthreads();
    for (int c = threadIdx.x; c < dh; c += THREADS) {
      float rev = 0.f;
      for (int j = Lk - 1; j >= 0; --j) {
        if (sm.sel[j] < 0) rev += s_do[j * pitch + c];
        dv[static_cast<long long>(j) * p.v_ls + c] += rev;
      }
    }
  }
}

static size_t fwd_smem(const RfAttnParams* p) {
  const int pitch = p->dh + 4;
  const int u = p->mode == RF_ATTN_FULL ? p->Lq : p->u;
  return sizeof(float) * (static_cast<size_t>(p->Lq) * pitch + 2 * static_cast<size_t>(p->Lk) * pitch + round4(u * p->Lk) + round4(p->Lq) +
                          round4(p->dh) + THREADS * 4 + round4(u) + round4(p->Lq));
}
static size_t bwd_smem(const RfAttnParams* p) {
  const int pitch = p->dh + 4;
  const int u = p->mode == RF_ATTN_FULL ? p->Lq : p->u;
  return fwd_smem(p) + sizeof(float) * (static_cast<size_t>(p->Lq) * pitch + round4(u * p->Lk));
}

static int validate(const RfAttnParams* p, const char* who, bool forward) {
  RF_CHECK_ARG(p->q && p->k && p->v, "%s: null q/k/v", who);
  RF_CHECK_ARG(p->B > 0 && p->H > 0 && p->Lq > 0 && p->Lk > 0 && p->dh > 0, "%s: bad shape", who);
  RF_CHECK_ARG(p->dh % 4 == 0 && p->q_ls % 4 == 0 && p->k_ls % 4 == 0 && p->v_ls % 4 == 0 && p->q_bs % 4 == 0 && p->k_bs % 4 == 0 &&
                   p->v_bs % 4 == 0,
               "%s: head dim and strides must be multiples of 4 (16 B vector access), dh=%d", who, p->dh);
  RF_CHECK_ARG(((reinterpret_cast<uintptr_t>(p->q) | reinterpret_cast<uintptr_t>(p->k) | reinterpret_cast<uintptr_t>(p->v)) & 15) == 0,
               "%s: q/k/v must be 16-byte aligned", who);
  RF_CHECK_ARG(p->dh <= 4 * THREADS, "%s: head dim %d > %d", who, p->dh, 4 * THREADS);
  RF_CHECK_ARG(p->Lk <= 32 * MAX_KEYS_PER_LANE, "%s: at most %d keys per problem", who, 32 * MAX_KEYS_PER_LANE);
  RF_CHECK_ARG(p->mode >= 0 && p->mode <= 2, "%s: bad mode %d", who, p->mode);
  if (p->mode != RF_ATTN_FULL) {
    RF_CHECK_ARG(p->u > 0 && p->u <= p->Lq && p->U > 0 && p->U <= p->Lk, "%s: bad budgets u=%d U=%d", who, p->u, p->U);
    RF_CHECK_ARG(!forward || p->idx || p->forced_top, "%s: idx table missing", who);
    RF_CHECK_ARG(p->top, "%s: top buffer missing", who);
  }
  RF_CHECK_ARG(p->mode != RF_ATTN_PROB_MASKED || p->Lq == p->Lk, "%s: masked ProbSparse attention requires Lq == Lk", who);
  RF_CHECK_ARG(static_cast<long long>(p->B) * p->H <= 2147483647LL, "%s: too many problems", who);
  return RF_OK;
}

// Opt-in to > 48 KiB of dynamic shared memory once per kernel (not per launch: keeps launches capturable in CUDA graphs).
template <typename K>
static int configure(K kernel, size_t smem) {
  static const void* configured[16] = {nullptr};
  static int n_configured = 0;
  if (smem <= 48 * 1024) return RF_OK;
  const void* key = reinterpret_cast<const void*>(kernel);
  for (int i = 0; i < n_configured; ++i)
    if (configured[i] == key) return RF_OK;
  RF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  if (n_configured < 16) configured[n_configured++] = key;
  return RF_OK;
}

}  // namespace attn
}  // namespace rf

#define RF_ATTN_LAUNCH(KERNEL, DHT, ARG, GRID, SMEM, STREAM)                       \
  {                                                                                \
    rc = attn::configure(attn::KERNEL<DHT>, SMEM);                                 \
    if (rc != RF_OK) return rc;                                                    \
    attn::KERNEL<DHT><<<GRID, attn::THREADS, SMEM, STREAM>>>(ARG);                 \
  }
#define RF_ATTN_DISPATCH(KERNEL, ARG, DHVAL, GRID, SMEM, STREAM)                   \
  switch (DHVAL) {                                                                 \
    case 8: RF_ATTN_LAUNCH(KERNEL, 8, ARG, GRID, SMEM, STREAM) break;              \
    case 16: RF_ATTN_LAUNCH(KERNEL, 16, ARG, GRID, SMEM, STREAM) break;            \
    case 104: RF_ATTN_LAUNCH(KERNEL, 104, ARG, GRID, SMEM, STREAM) break;          \
    default: RF_ATTN_LAUNCH(KERNEL, 0, ARG, GRID, SMEM, STREAM) break;             \
  }

extern "C" int rf_attention_fwd(const RfAttnParams* p, void* stream) {
  using namespace rf;
  RF_CHECK_ARG(p && p->out, "rf_attention_fwd: null pointer");
  int rc = attn::validate(p, "rf_attention_fwd", true);
  if (rc != RF_OK) return rc;
  const size_t smem = attn::fwd_smem(p);
  RF_CHECK_ARG(smem <= 220 * 1024, "rf_attention_fwd: problem needs %zu B of shared memory (> 220 KiB)", smem);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  RF_ATTN_DISPATCH(attention_fwd_kernel, *p, p->dh, p->B * p->H, smem, s)
  RF_LAUNCH_OK();
  return RF_OK;
}

extern "C" int rf_attention_bwd(const RfAttnBwdParams* p, void* stream) {
  using namespace rf;
  RF_CHECK_ARG(p && p->dout && p->dq && p->dk && p->dv, "rf_attention_bwd: null pointer");
  int rc = attn::validate(&p->f, "rf_attention_bwd", false);
  if (rc != RF_OK) return rc;
  const size_t smem = attn::bwd_smem(&p->f);
  RF_CHECK_ARG(smem <= 220 * 1024, "rf_attention_bwd: problem needs %zu B of shared memory (> 220 KiB)", smem);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  RF_ATTN_DISPATCH(attention_bwd_kernel, *p, p->f.dh, p->f.B * p->f.H, smem, s)
  RF_LAUNCH_OK();
  return RF_OK;
}
