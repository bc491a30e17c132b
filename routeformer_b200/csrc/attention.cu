// Fused ProbSparse / full attention, forward and backward.  See include/routeformer_b200.h (4).
//
// One CTA owns one (clip b, head h) problem; Q, K, V of that head are staged once in shared memory
// (rows padded to dh+1 floats: conflict-free column walks), and the whole chain
//   sampled scores -> sparsity measure -> top-u selection -> scaled scores of the selected queries
//   -> (causal) softmax -> P.V -> mean(V)/cumsum(V) fill of the unselected queries
// runs out of shared memory with fp32 FMA arithmetic: the selection is a discontinuous function of the
// scores, so it is computed at full fp32 precision rather than on the tensor cores (hard part 2).
// The reference materialises K_expand[..., index_sample, :] = [B,H,Lq,U,dh] in HBM (426 MB per call at
// B=64); here nothing but Q/K/V/context (and the u selected indices) ever touches HBM.
// Problems are tiny (L <= 160, dh <= 128), so the grid (B*H CTAs) is what fills the machine.
#include "common.cuh"

namespace rf {
namespace attn {

constexpr int THREADS = 128;

struct Smem {
  float* q; float* k; float* v;  // [L][dh+1]
  float* s;                      // [u][Lk] scores / probabilities
  float* m;                      // [Lq] sparsity measure
  int* top;                      // [u]
  int* sel;                      // [Lq] rank of the query in the selection or -1
};

__device__ __forceinline__ void load_tile(float* dst, const float* src, long long ls, int L, int dh, int pitch) {
  // rows of dh contiguous floats, row stride ls
  for (int i = threadIdx.x; i < L * dh; i += blockDim.x) {
    const int l = i / dh, e = i % dh;
    dst[l * pitch + e] = src[static_cast<long long>(l) * ls + e];
  }
}

__device__ __forceinline__ float dot_rows(const float* a, const float* b, int dh) {
  float acc = 0.f;
#pragma unroll 4
  for (int e = 0; e < dh; ++e) acc = fmaf(a[e], b[e], acc);
  return acc;
}

// Steps shared by forward and backward: selection (or replay of it) and the probability matrix of the selected rows.
__device__ void select_and_softmax(const RfAttnParams& p, const Smem& sm, int b, int h, int pitch, int u, bool compute_selection,
                                   const int* top_in, float* measure_out) {
  const int Lq = p.Lq, Lk = p.Lk, dh = p.dh;
  if (p.mode == RF_ATTN_FULL) {
    for (int i = threadIdx.x; i < Lq; i += blockDim.x) { sm.top[i] = i; sm.sel[i] = i; }
  } else if (compute_selection && !p.forced_top) {
    const int group = p.idx_group > 0 ? b / p.idx_group : 0;
    const int* idx = p.idx + static_cast<long long>(group) * Lq * p.U;
    for (int i = threadIdx.x; i < Lq; i += blockDim.x) {
      float mx = -INFINITY, sum = 0.f;
      for (int j = 0; j < p.U; ++j) {
        const float s = dot_rows(sm.q + i * pitch, sm.k + idx[i * p.U + j] * pitch, dh);
        mx = fmaxf(mx, s);
        sum += s;
      }
      const float mval = mx - sum / Lk;
      sm.m[i] = mval;
      if (measure_out) measure_out[i] = mval;
    }
    __syncthreads();
    // rank-based top-u: rank = number of queries that beat this one (ties -> lower index first)
    for (int i = threadIdx.x; i < Lq; i += blockDim.x) {
      const float mi = sm.m[i];
      int rank = 0;
      for (int j = 0; j < Lq; ++j) {
        const float mj = sm.m[j];
        rank += (mj > mi) || (mj == mi && j < i);
      }
      if (rank < u) { sm.top[rank] = i; sm.sel[i] = rank; }
      else sm.sel[i] = -1;
    }
  } else {
    for (int i = threadIdx.x; i < Lq; i += blockDim.x) sm.sel[i] = -1;
    __syncthreads();
    for (int r = threadIdx.x; r < u; r += blockDim.x) {
      const int i = top_in[r];
      sm.top[r] = i;
      sm.sel[i] = r;
    }
  }
  __syncthreads();
  // scaled scores of the selected queries against every key (causal: keys after the query are masked)
  const float scale = rsqrtf(static_cast<float>(dh));
  for (int i = threadIdx.x; i < u * Lk; i += blockDim.x) {
    const int r = i / Lk, j = i % Lk;
    const int qi = sm.top[r];
    float s = dot_rows(sm.q + qi * pitch, sm.k + j * pitch, dh) * scale;
    if (p.mode == RF_ATTN_PROB_MASKED && j > qi) s = -INFINITY;
    sm.s[i] = s;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = warp; r < u; r += nwarps) {
    float* row = sm.s + r * Lk;
    float mx = -INFINITY;
    for (int j = lane; j < Lk; j += 32) mx = fmaxf(mx, row[j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Lk; j += 32) {
      const float e = expf(row[j] - mx);
      row[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < Lk; j += 32) row[j] *= inv;
  }
  __syncthreads();
}

__device__ __forceinline__ Smem carve(float* base, int Lq, int Lk, int dh, int u, int pitch) {
  Smem sm;
  sm.q = base;
  sm.k = sm.q + Lq * pitch;
  sm.v = sm.k + Lk * pitch;
  sm.s = sm.v + Lk * pitch;
  sm.m = sm.s + u * Lk;
  sm.top = reinterpret_cast<int*>(sm.m + Lq);
  sm.sel = sm.top + u;
  return sm;
}

__device__ __forceinline__ long long out_offset(const RfAttnParams& p, int b, int h, int l) {
  return p.out_layout == RF_LAYOUT_BLHD ? ((static_cast<long long>(b) * p.Lq + l) * p.H + h) * p.dh
                                        : ((static_cast<long long>(b) * p.H + h) * p.Lq + l) * p.dh;
}

__global__ void __launch_bounds__(THREADS) attention_fwd_kernel(const RfAttnParams p) {
  extern __shared__ float smem_f[];
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int Lq = p.Lq, Lk = p.Lk, dh = p.dh, pitch = dh + 1;
  const int u = p.mode == RF_ATTN_FULL ? Lq : p.u;
  Smem sm = carve(smem_f, Lq, Lk, dh, u, pitch);
  load_tile(sm.q, p.q + b * p.q_bs + h * dh, p.q_ls, Lq, dh, pitch);
  load_tile(sm.k, p.k + b * p.k_bs + h * dh, p.k_ls, Lk, dh, pitch);
  load_tile(sm.v, p.v + b * p.v_bs + h * dh, p.v_ls, Lk, dh, pitch);
  __syncthreads();
  const long long bh = static_cast<long long>(b) * p.H + h;
  select_and_softmax(p, sm, b, h, pitch, u, true, p.forced_top ? p.forced_top + bh * u : nullptr,
                     p.measure ? p.measure + bh * Lq : nullptr);
  if (p.mode != RF_ATTN_FULL && p.top)
    for (int r = threadIdx.x; r < u; r += blockDim.x) p.top[bh * u + r] = sm.top[r];

  // selected queries: context = P . V
  for (int i = threadIdx.x; i < u * dh; i += blockDim.x) {
    const int r = i / dh, d = i % dh;
    const float* prow = sm.s + r * Lk;
    float acc = 0.f;
    for (int j = 0; j < Lk; ++j) acc = fmaf(prow[j], sm.v[j * pitch + d], acc);
    p.out[out_offset(p, b, h, sm.top[r]) + d] = acc;
  }
  // unselected queries: mean(V) (unmasked) or cumsum(V) (masked)
  if (p.mode == RF_ATTN_PROB) {
    for (int d = threadIdx.x; d < dh; d += blockDim.x) {
      float acc = 0.f;
      for (int j = 0; j < Lk; ++j) acc += sm.v[j * pitch + d];
      const float mean = acc / Lk;
      for (int l = 0; l < Lq; ++l)
        if (sm.sel[l] < 0) p.out[out_offset(p, b, h, l) + d] = mean;
    }
  } else if (p.mode == RF_ATTN_PROB_MASKED) {
    for (int d = threadIdx.x; d < dh; d += blockDim.x) {
      float acc = 0.f;
      for (int l = 0; l < Lq; ++l) {
        acc += sm.v[l * pitch + d];
        if (sm.sel[l] < 0) p.out[out_offset(p, b, h, l) + d] = acc;
      }
    }
  }
}

// Backward: recomputes P from Q, K and the saved selection, then
//   dV[j]  = sum_r P[r][j] dO[top_r]  +  fill-path gradient (mean: sum of unselected dO / Lk; cumsum: reverse cumsum of unselected dO)
//   dS     = P o (dP - rowsum(P o dP)),  dP[r][j] = dO[top_r] . V[j]
//   dQ[top_r] = scale * sum_j dS[r][j] K[j]   (other rows 0),   dK[j] = scale * sum_r dS[r][j] Q[top_r]
__global__ void __launch_bounds__(THREADS) attention_bwd_kernel(const RfAttnBwdParams bp) {
  extern __shared__ float smem_f[];
  const RfAttnParams& p = bp.f;
  const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
  const int Lq = p.Lq, Lk = p.Lk, dh = p.dh, pitch = dh + 1;
  const int u = p.mode == RF_ATTN_FULL ? Lq : p.u;
  Smem sm = carve(smem_f, Lq, Lk, dh, u, pitch);
  float* s_do = reinterpret_cast<float*>(sm.sel + Lq);  // [Lq][pitch] context gradient of this head
  float* s_ds = s_do + Lq * pitch;                      // [u][Lk]
  load_tile(sm.q, p.q + b * p.q_bs + h * dh, p.q_ls, Lq, dh, pitch);
  load_tile(sm.k, p.k + b * p.k_bs + h * dh, p.k_ls, Lk, dh, pitch);
  load_tile(sm.v, p.v + b * p.v_bs + h * dh, p.v_ls, Lk, dh, pitch);
  for (int i = threadIdx.x; i < Lq * dh; i += blockDim.x) {
    const int l = i / dh, d = i % dh;
    s_do[l * pitch + d] = bp.dout[out_offset(p, b, h, l) + d];
  }
  __syncthreads();
  const long long bh = static_cast<long long>(b) * p.H + h;
  select_and_softmax(p, sm, b, h, pitch, u, false, p.mode == RF_ATTN_FULL ? nullptr : p.top + bh * u, nullptr);

  // dP, then dS in place
  for (int i = threadIdx.x; i < u * Lk; i += blockDim.x) {
    const int r = i / Lk, j = i % Lk;
    s_ds[i] = dot_rows(s_do + sm.top[r] * pitch, sm.v + j * pitch, dh);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float scale = rsqrtf(static_cast<float>(dh));
  for (int r = warp; r < u; r += nwarps) {
    float acc = 0.f;
    for (int j = lane; j < Lk; j += 32) acc += sm.s[r * Lk + j] * s_ds[r * Lk + j];
    acc = warp_sum(acc);
    for (int j = lane; j < Lk; j += 32) s_ds[r * Lk + j] = sm.s[r * Lk + j] * (s_ds[r * Lk + j] - acc) * scale;
  }
  __syncthreads();

  // dQ
  float* dq = bp.dq + b * p.q_bs + h * dh;
  for (int i = threadIdx.x; i < Lq * dh; i += blockDim.x) {
    const int l = i / dh, d = i % dh;
    const int r = sm.sel[l];
    float acc = 0.f;
    if (r >= 0) {
      const float* row = s_ds + r * Lk;
      for (int j = 0; j < Lk; ++j) acc = fmaf(row[j], sm.k[j * pitch + d], acc);
    }
    dq[static_cast<long long>(l) * p.q_ls + d] = acc;
  }
  // dK
  float* dk = bp.dk + b * p.k_bs + h * dh;
  for (int i = threadIdx.x; i < Lk * dh; i += blockDim.x) {
    const int j = i / dh, d = i % dh;
    float acc = 0.f;
    for (int r = 0; r < u; ++r) acc = fmaf(s_ds[r * Lk + j], sm.q[sm.top[r] * pitch + d], acc);
    dk[static_cast<long long>(j) * p.k_ls + d] = acc;
  }
  // dV
  float* dv = bp.dv + b * p.v_bs + h * dh;
  for (int d = threadIdx.x; d < dh; d += blockDim.x) {
    float fill = 0.f;
    if (p.mode == RF_ATTN_PROB) {
      for (int l = 0; l < Lq; ++l)
        if (sm.sel[l] < 0) fill += s_do[l * pitch + d];
      fill /= Lk;
    }
    float rev = 0.f;  // reverse cumsum of the unselected rows (masked mode, Lq == Lk)
    for (int j = Lk - 1; j >= 0; --j) {
      float acc = 0.f;
      for (int r = 0; r < u; ++r) acc = fmaf(sm.s[r * Lk + j], s_do[sm.top[r] * pitch + d], acc);
      if (p.mode == RF_ATTN_PROB) acc += fill;
      else if (p.mode == RF_ATTN_PROB_MASKED) {
        if (sm.sel[j] < 0) rev += s_do[j * pitch + d];
        acc += rev;
      }
      dv[static_cast<long long>(j) * p.v_ls + d] = acc;
    }
  }
}

static size_t fwd_smem(const RfAttnParams* p) {
  const int pitch = p->dh + 1;
  const int u = p->mode == RF_ATTN_FULL ? p->Lq : p->u;
  return sizeof(float) * (static_cast<size_t>(p->Lq) * pitch + 2 * static_cast<size_t>(p->Lk) * pitch + static_cast<size_t>(u) * p->Lk + p->Lq) +
         sizeof(int) * (static_cast<size_t>(u) + p->Lq);
}
static size_t bwd_smem(const RfAttnParams* p) {
  const int pitch = p->dh + 1;
  const int u = p->mode == RF_ATTN_FULL ? p->Lq : p->u;
  return fwd_smem(p) + sizeof(float) * (static_cast<size_t>(p->Lq) * pitch + static_cast<size_t>(u) * p->Lk);
}

static int validate(const RfAttnParams* p, const char* who, bool forward) {
  RF_CHECK_ARG(p->q && p->k && p->v, "%s: null q/k/v", who);
  RF_CHECK_ARG(p->B > 0 && p->H > 0 && p->Lq > 0 && p->Lk > 0 && p->dh > 0, "%s: bad shape", who);
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

}  // namespace attn
}  // namespace rf

extern "C" int rf_attention_fwd(const RfAttnParams* p, void* stream) {
  using namespace rf;
  RF_CHECK_ARG(p && p->out, "rf_attention_fwd: null pointer");
  int rc = attn::validate(p, "rf_attention_fwd", true);
  if (rc != RF_OK) return rc;
  const size_t smem = attn::fwd_smem(p);
  RF_CHECK_ARG(smem <= 220 * 1024, "rf_attention_fwd: problem needs %zu B of shared memory (> 220 KiB)", smem);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    RF_CUDA_OK(cudaFuncSetAttribute(attn::attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = 220 * 1024;
  }
  attn::attention_fwd_kernel<<<p->B * p->H, attn::THREADS, smem, static_cast<cudaStream_t>(stream)>>>(*p);
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
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    RF_CUDA_OK(cudaFuncSetAttribute(attn::attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = 220 * 1024;
  }
  attn::attention_bwd_kernel<<<p->f.B * p->f.H, attn::THREADS, smem, static_cast<cudaStream_t>(stream)>>>(*p);
  RF_LAUNCH_OK();
  return RF_OK;
}
