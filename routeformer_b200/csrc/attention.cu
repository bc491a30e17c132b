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
//
// Two families: the generic kernels (any mode, L <= 256, any dh) described above, and -- further down -- register-blocked
// kernels for the small ProbSparse problems that dominate the step (frame encoder: L = 65, dh = 16), with a last-query-only
// variant for the encoder layer whose caller keeps one token.  Full attention optionally drops probabilities (Philox mask).
#include <cstdlib>

#include "common.cuh"

namespace rf {
// tensor-core forward for the small unmasked ProbSparse problems (attention_tc.cu)
bool attention_tc_fwd_eligible(const RfAttnParams* p);
int attention_tc_fwd(const RfAttnParams* p, cudaStream_t stream);

namespace attn {

constexpr int THREADS = 128;
constexpr int NWARPS = THREADS / 32;
constexpr int MAX_KEYS_PER_LANE = 8;  // Lk <= 256

template <int DH>
struct Dims {
  int dh_rt;
  __device__ __forceinline__ int dh() const { return DH ? DH : dh_rt; }
  __device__ __forceinline__ int dh4() const { return dh() >> 2; }
  __device__ __forceinline__ int pitch() const { return dh() + 4; }
};

__host__ __device__ __forceinline__ int round4(int x) { return (x + 3) & ~3; }

// Packed fp32 arithmetic (Blackwell FFMA2: two IEEE fp32 FMAs per issue slot).  Each half is an ordinary fma.rn, so results are
// bit-identical to the scalar formulation with the same accumulation order; these kernels are issue-bound, halving the FMA
// instruction count is a direct win.
#ifndef RF_ATTN_FFMA2
#define RF_ATTN_FFMA2 1
#endif
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi);
__device__ __forceinline__ float2 unpack2(f32x2 v);
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
#if RF_ATTN_FFMA2
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
#else
  const float2 x = unpack2(a), y = unpack2(b), z = unpack2(c);
  return pack2(fmaf(x.x, y.x, z.x), fmaf(x.y, y.y, z.y));
#endif
}
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float2 unpack2(f32x2 v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
// acc(4 channels as two pairs) += w * x(4 channels); the scalar w is a broadcast operand of FFMA2
struct Acc4 {
  f32x2 lo, hi;
  __device__ __forceinline__ Acc4() : lo(0ull), hi(0ull) {}
  __device__ __forceinline__ void fma(float w, const float* x4) {
    const ulonglong2 x = *reinterpret_cast<const ulonglong2*>(x4);
    const f32x2 ww = pack2(w, w);
    lo = ffma2(ww, x.x, lo);
    hi = ffma2(ww, x.y, hi);
  }
  __device__ __forceinline__ float4 get() const {
    const float2 a = unpack2(lo), b = unpack2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
  }
};

// a . b over dh channels: lanes (0,1) of one packed accumulator take channels (4c, 4c+1) then (4c+2, 4c+3)
template <int DH>
__device__ __forceinline__ float dot4(const Dims<DH> d, const float* a, const float* b) {
  f32x2 acc = 0ull;
  if (DH != 0) {
#pragma unroll
    for (int c = 0; c < (DH ? DH / 4 : 1); ++c) {
      const ulonglong2 x = reinterpret_cast<const ulonglong2*>(a)[c], y = reinterpret_cast<const ulonglong2*>(b)[c];
      acc = ffma2(x.x, y.x, acc);
      acc = ffma2(x.y, y.y, acc);
    }
  } else {
    const int dh4 = d.dh4();
    for (int c = 0; c < dh4; ++c) {
      const ulonglong2 x = reinterpret_cast<const ulonglong2*>(a)[c], y = reinterpret_cast<const ulonglong2*>(b)[c];
      acc = ffma2(x.x, y.x, acc);
      acc = ffma2(x.y, y.y, acc);
    }
  }
  const float2 r = unpack2(acc);
  return r.x + r.y;
}

__device__ __forceinline__ long long out_offset(const RfAttnParams& p, int b, int h, int l) {
  return p.out_layout == RF_LAYOUT_BLHD ? ((static_cast<long long>(b) * p.Lq + l) * p.H + h) * p.dh
                                        : ((static_cast<long long>(b) * p.H + h) * p.Lq + l) * p.dh;
}

// (nested namespaces: argument-dependent lookup on Dims<> must not see both instances)
namespace narrow {
#include "attention_generic.inc"
}  // namespace narrow
using namespace narrow;

namespace wide {  // the same kernels with 256 threads per CTA
constexpr int THREADS = 256;
constexpr int NWARPS = THREADS / 32;
#include "attention_generic.inc"
}  // namespace wide

// ---------------------------------------------------------------------------------------------
// Small-problem ProbSparse path: mode PROB, head dim 8 / 16, Lk <= 96, Lq*Lk scores fit in shared memory.
//
// The generic kernels above are bound by shared-memory bandwidth, not arithmetic (ncu: LSU shared wavefronts ~77 % busy, FMA
// pipe < 20 %): every dot product re-reads a key row (4 x LDS.128 per lane).  For the frame encoder (1536 sequences x 8 heads,
// L = 65, dh = 16: 2/3 of all attention time) the operands are re-organised around registers instead:
//   * lane j of a "key-slot" warp keeps key row j (and, in the backward pass, value row j) in REGISTERS; the query rows arrive
//     as warp-wide broadcast loads (1 wavefront each), so one raw score costs 0.16 wavefronts instead of ~0.6;
//   * the full raw score matrix S = Q K^T [Lq][Lk] is produced ONCE: the sparsity measure gathers its sampled entries from it
//     and the selected rows are soft-maxed in place, so the selected queries need no second round of dot products;
//   * dK / dV accumulate in registers of the lane that owns key j and leave through one 64 B row store each.
// Arithmetic order of every reduction is fixed (deterministic) and identical between forward and the backward recompute.
// ---------------------------------------------------------------------------------------------
// A row of DH floats held in registers as DH/2 packed pairs.
template <int DH>
struct RegRow {
  f32x2 v[DH / 2];
  __device__ __forceinline__ void load(const float* src, bool ok) {  // global or shared, 16 B aligned
#pragma unroll
    for (int c = 0; c < DH / 4; ++c) {
      const ulonglong2 t = ok ? *reinterpret_cast<const ulonglong2*>(src + 4 * c) : make_ulonglong2(0ull, 0ull);
      v[2 * c] = t.x;
      v[2 * c + 1] = t.y;
    }
  }
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int c = 0; c < DH / 2; ++c) v[c] = 0ull;
  }
  // this . x: pairs (4c,4c+1) then (4c+2,4c+3) into one packed accumulator, the order of dot4
  __device__ __forceinline__ float dot(const RegRow& x) const {
    f32x2 acc = 0ull;
#pragma unroll
    for (int c = 0; c < DH / 2; ++c) acc = ffma2(x.v[c], v[c], acc);
    const float2 r = unpack2(acc);
    return r.x + r.y;
  }
  __device__ __forceinline__ void axpy(float w, const RegRow& x) {  // this += w * x
    const f32x2 ww = pack2(w, w);
#pragma unroll
    for (int c = 0; c < DH / 2; ++c) v[c] = ffma2(ww, x.v[c], v[c]);
  }
  __device__ __forceinline__ void store(float* dst) const {
#pragma unroll
    for (int c = 0; c < DH / 4; ++c) *reinterpret_cast<ulonglong2*>(dst + 4 * c) = make_ulonglong2(v[2 * c], v[2 * c + 1]);
  }
};
template <int DH>
__device__ __forceinline__ void stage_rows(float* dst, const float* src, long long ls, int L) {  // [L][DH] dense in shared memory
  constexpr int DH4 = DH / 4;
  for (int i = threadIdx.x; i < L * DH4; i += THREADS) {
    const int l = i / DH4, c = i - l * DH4;
    reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src + static_cast<long long>(l) * ls) + c);
  }
}
// One warp: acc[0:DH] = sum over rows l < L with keep(l) of src[l][0:DH] (dense [L][DH]).  Lane (g, c) sums rows g, g+G, ...
// of channel group c, then the G partial sums are combined by a butterfly (a + b == b + a: all lanes agree, fixed order).
template <int DH, typename Keep>
__device__ __forceinline__ void warp_column_sum(const float* src, int L, float* acc, Keep keep) {
  constexpr int DH4 = DH / 4, G = 32 / DH4;
  const int lane = threadIdx.x & 31;
  const int g = lane / DH4, c = lane - g * DH4;
  float4 part = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int l = g; l < L; l += G)
    if (keep(l)) {
      const float4 v = reinterpret_cast<const float4*>(src)[l * DH4 + c];
      part.x += v.x; part.y += v.y; part.z += v.z; part.w += v.w;
    }
#pragma unroll
  for (int off = DH4; off < 32; off <<= 1) {
    part.x += __shfl_xor_sync(0xffffffffu, part.x, off); part.y += __shfl_xor_sync(0xffffffffu, part.y, off);
    part.z += __shfl_xor_sync(0xffffffffu, part.z, off); part.w += __shfl_xor_sync(0xffffffffu, part.w, off);
  }
  if (lane < DH4) reinterpret_cast<float4*>(acc)[lane] = part;
}
// One warp: acc[0:DH] = sum_l w[l] * src[l][0:DH] (dense [L][DH] rows, weights in shared memory); fixed reduction order.
template <int DH>
__device__ __forceinline__ float4 warp_weighted_column_sum(const float* src, const float* w, int L) {
  constexpr int DH4 = DH / 4, G = 32 / DH4;
  const int lane = threadIdx.x & 31;
  const int g = lane / DH4, c = lane - g * DH4;
  float4 part = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int l = g; l < L; l += G) {
    const float wl = w[l];
    const float4 v = reinterpret_cast<const float4*>(src)[l * DH4 + c];
    part.x = fmaf(wl, v.x, part.x); part.y = fmaf(wl, v.y, part.y); part.z = fmaf(wl, v.z, part.z); part.w = fmaf(wl, v.w, part.w);
  }
#pragma unroll
  for (int off = DH4; off < 32; off <<= 1) {
    part.x += __shfl_xor_sync(0xffffffffu, part.x, off); part.y += __shfl_xor_sync(0xffffffffu, part.y, off);
    part.z += __shfl_xor_sync(0xffffffffu, part.z, off); part.w += __shfl_xor_sync(0xffffffffu, part.w, off);
  }
  return part;  // valid in every lane; lane c < DH4 holds channels 4c..4c+3
}
// in-place softmax of one probability row by one warp (Lk <= 96); returns nothing, row[j] = softmax(scale * row)[j]
__device__ __forceinline__ void warp_softmax_row(float* row, int Lk, float scale) {
  const int lane = threadIdx.x & 31;
  float sc[3];
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    const int j = lane + 32 * t;
    sc[t] = j < Lk ? row[j] * scale : -INFINITY;
    mx = fmaxf(mx, sc[t]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    sc[t] = (lane + 32 * t < Lk) ? __expf(sc[t] - mx) : 0.f;
    sum += sc[t];
  }
  const float inv = 1.f / warp_sum(sum);
#pragma unroll
  for (int t = 0; t < 3; ++t)
    if (lane + 32 * t < Lk) row[lane + 32 * t] = sc[t] * inv;
}

// out[rows of one 4-row tile][4c..4c+3] = sum_j W[row][j] * X[j][4c..4c+3]: the X row is loaded once per j for four output rows
// (a shared-memory load costs one wavefront per 128 B quarter-warp whether or not lanes coincide, so operand re-use in
// registers is the only way to cut shared-memory traffic).  wrow[k] = shared row of weights of output row k (clamped rows repeat).
template <int DH>
__device__ __forceinline__ void tile4_matvec(const float* const (&wrow)[4], const float* x, int L, float4 (&out)[4]) {
  Acc4 acc[4];
#pragma unroll 4
  for (int j = 0; j < L; ++j) {
    const ulonglong2 xv = *reinterpret_cast<const ulonglong2*>(x + j * DH);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float w = wrow[k][j];
      const f32x2 ww = pack2(w, w);
      acc[k].lo = ffma2(ww, xv.x, acc[k].lo);
      acc[k].hi = ffma2(ww, xv.y, acc[k].hi);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) out[k] = acc[k].get();
}

// Key ownership: lane l of every warp owns keys l, l+32, ... (NT register rows); the rows of the left operand are split over the
// warps.  A short remainder of <= 8 keys (the 65th token of the frame encoder) would cost a whole register slot for almost no
// work, so those "tail" keys are handled transposed (lane per row).  NT is chosen on the host: see small_nt().
__host__ __device__ __forceinline__ int small_nt(int Lk) { return (Lk >> 5) + ((Lk & 31) > 8 ? 1 : 0); }

template <int DH, int NT>
__global__ void __launch_bounds__(THREADS, 7) attention_small_fwd_kernel(const RfAttnParams p) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) float smem_f[];
  constexpr int DH4 = DH / 4;
  const int b = blockIdx.x / p.H, h = blockIdx.x - b * p.H;
  const int Lq = p.Lq, Lk = p.Lk, u = p.u, U = p.U;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_q = smem_f;                     // [Lq][DH]
  float* s_v = s_q + Lq * DH;              // [Lk][DH]
  float* s_s = s_v + Lk * DH;              // [Lq][Lk] raw scores; rows of the selected queries become probabilities in place
  float* s_m = s_s + round4(Lq * Lk);      // [Lq]
  float* s_acc = s_m + round4(Lq);         // [DH]
  int* s_top = reinterpret_cast<int*>(s_acc + DH);
  int* s_sel = s_top + round4(u);
  unsigned char* s_idx = reinterpret_cast<unsigned char*>(s_sel + round4(Lq));  // [Lq][U] sampled key ids (Lk <= 96 < 256)
  const float* gq = p.q + b * p.q_bs + h * DH;
  const float* gk = p.k + b * p.k_bs + h * DH;
  const float* gv = p.v + b * p.v_bs + h * DH;
  const long long bh = static_cast<long long>(b) * p.H + h;
  const bool select = p.forced_top == nullptr;
  const int tail_start = min(Lk, NT * 32), tail = Lk - tail_start;

  stage_rows<DH>(s_q, gq, p.q_ls, Lq);
  stage_rows<DH>(s_v, gv, p.v_ls, Lk);
  if (select) {
    const int group = p.idx_group > 0 ? b / p.idx_group : 0;
    const int* idx = p.idx + static_cast<long long>(group) * Lq * U;
    for (int i = threadIdx.x; i < Lq * U; i += THREADS) s_idx[i] = static_cast<unsigned char>(__ldg(idx + i));
  }
  RegRow<DH> kr[NT > 0 ? NT : 1];
#pragma unroll
  for (int t = 0; t < NT; ++t) kr[t].load(gk + static_cast<long long>(lane + 32 * t) * p.k_ls, lane + 32 * t < Lk);
  __syncthreads();

  // raw scores S[i][j] = q_i . k_j: every query row is read once per warp and meets NT register-resident keys per lane
  if (NT > 0) {
    const int chunk_rows = (Lq + NWARPS - 1) / NWARPS;
    const int i0 = warp * chunk_rows, i1 = min(Lq, i0 + chunk_rows);
#pragma unroll 2
    for (int i = i0; i < i1; ++i) {
      RegRow<DH> qr;
      qr.load(s_q + i * DH, true);
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const float sc = kr[t].dot(qr);
        if (lane + 32 * t < Lk) s_s[i * Lk + lane + 32 * t] = sc;
      }
    }
  }
  {  // tail keys, transposed: lane per query row
    const int qblocks = (Lq + 31) >> 5;
    for (int it = NWARPS - 1 - warp; it < tail * qblocks; it += NWARPS) {
      const int jt = tail_start + it / qblocks, i = (it % qblocks) * 32 + lane;
      RegRow<DH> kt, qr;
      kt.load(gk + static_cast<long long>(jt) * p.k_ls, true);
      if (i < Lq) {
        qr.load(s_q + i * DH, true);
        s_s[i * Lk + jt] = kt.dot(qr);
      }
    }
  }
  if (warp == NWARPS - 1) warp_column_sum<DH>(s_v, Lk, s_acc, [](int) { return true; });  // mean(V) numerator
  __syncthreads();

  if (select) {
    // sparsity measure from the sampled entries of S (cross_modal_transformer.py:95-103)
    for (int i = threadIdx.x; i < Lq; i += THREADS) {
      const float* row = s_s + i * Lk;
      const unsigned char* irow = s_idx + i * U;
      float mx = -INFINITY, sum = 0.f;
#pragma unroll 5
      for (int jj = 0; jj < U; ++jj) {
        const float sc = row[irow[jj]];
        mx = fmaxf(mx, sc);
        sum += sc;
      }
      const float mval = mx - sum / Lk;
      s_m[i] = mval;
      if (p.measure) p.measure[bh * Lq + i] = mval;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Lq; i += THREADS) {  // rank-based top-u, ties -> lower index first
      const float mi = s_m[i];
      int rank = 0;
#pragma unroll 8
      for (int j = 0; j < i; ++j) rank += s_m[j] >= mi;
#pragma unroll 8
      for (int j = i + 1; j < Lq; ++j) rank += s_m[j] > mi;
      if (rank < u) { s_top[rank] = i; s_sel[i] = rank; }
      else s_sel[i] = -1;
    }
  } else {
    for (int i = threadIdx.x; i < Lq; i += THREADS) s_sel[i] = -1;
    __syncthreads();
    const int* top_in = p.forced_top + bh * u;
    for (int r = threadIdx.x; r < u; r += THREADS) {
      const int i = top_in[r];
      s_top[r] = i;
      s_sel[i] = r;
    }
  }
  __syncthreads();
  if (p.top)
    for (int r = threadIdx.x; r < u; r += THREADS) p.top[bh * u + r] = s_top[r];

  const float scale = rsqrtf(static_cast<float>(DH));
  if (p.tail_only) {
    // only the last query's context is consumed: one probability row (if that query was selected) or mean(V)
    const int last = Lq - 1;
    if (warp == 0) {
      float4 o;
      if (s_sel[last] >= 0) {  // warp-uniform
        float* row = s_s + last * Lk;
        warp_softmax_row(row, Lk, scale);
        __syncwarp();
        o = warp_weighted_column_sum<DH>(s_v, row, Lk);
      } else {
        const float inv = 1.f / Lk;
        const float4 a4 = reinterpret_cast<const float4*>(s_acc)[lane < DH4 ? lane : 0];
        o = make_float4(a4.x * inv, a4.y * inv, a4.z * inv, a4.w * inv);
      }
      if (lane < DH4) *reinterpret_cast<float4*>(p.out + out_offset(p, b, h, last) + 4 * lane) = o;
    }
    return;
  }

  // selected rows: probabilities in place (one warp per row, registers + shuffles)
  for (int r = warp; r < u; r += NWARPS) {
    float* row = s_s + s_top[r] * Lk;
    float sc[3];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const int j = lane + 32 * t;
      sc[t] = j < Lk ? row[j] * scale : -INFINITY;
      mx = fmaxf(mx, sc[t]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      sc[t] = (lane + 32 * t < Lk) ? __expf(sc[t] - mx) : 0.f;
      sum += sc[t];
    }
    const float inv = 1.f / warp_sum(sum);
#pragma unroll
    for (int t = 0; t < 3; ++t)
      if (lane + 32 * t < Lk) row[lane + 32 * t] = sc[t] * inv;
  }
  __syncthreads();

  // context of the selected queries: P V in 4-row x 4-channel register tiles (28 threads at u = 25: the first warp);
  // the other warps meanwhile write mean(V) to every other query
  const int ntiles = ((u + 3) >> 2) * DH4;
  for (int it = threadIdx.x; it < ntiles; it += THREADS) {
    const int r0 = (it / DH4) * 4, c = it % DH4;
    const float* wrow[4];
    int qi[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      qi[k] = s_top[min(r0 + k, u - 1)];
      wrow[k] = s_s + qi[k] * Lk;
    }
    float4 o[4];
    tile4_matvec<DH>(wrow, s_v + 4 * c, Lk, o);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (r0 + k < u) *reinterpret_cast<float4*>(p.out + out_offset(p, b, h, qi[k]) + 4 * c) = o[k];
  }
  const float inv_lk = 1.f / Lk;
  for (int i = THREADS - 1 - threadIdx.x; i < Lq * DH4; i += THREADS) {  // reversed: the tile threads come last
    const int l = i / DH4, c = i - l * DH4;
    if (s_sel[l] < 0) {
      const float4 a4 = reinterpret_cast<const float4*>(s_acc)[c];
      *reinterpret_cast<float4*>(p.out + out_offset(p, b, h, l) + 4 * c) = make_float4(a4.x * inv_lk, a4.y * inv_lk, a4.z * inv_lk, a4.w * inv_lk);
    }
  }
}

template <int DH, int NT>
__global__ void __launch_bounds__(THREADS, 7) attention_small_bwd_kernel(const RfAttnBwdParams bp) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) float smem_f[];
  constexpr int DH4 = DH / 4;
  const RfAttnParams& p = bp.f;
  const int b = blockIdx.x / p.H, h = blockIdx.x - b * p.H;
  const int Lq = p.Lq, Lk = p.Lk, u = p.u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* s_q = smem_f;                      // [Lq][DH]
  float* s_k = s_q + Lq * DH;               // [Lk][DH]
  float* s_do = s_k + Lk * DH;              // [Lq][DH] context gradient of this head
  float* s_p = s_do + Lq * DH;              // [u][Lk] raw scores -> probabilities
  float* s_ds = s_p + round4(u * Lk);       // [u][Lk] dP -> dS
  float* s_acc = s_ds + round4(u * Lk);     // [DH] sum of the unselected context gradients
  int* s_top = reinterpret_cast<int*>(s_acc + DH);
  int* s_sel = s_top + round4(u);
  const float* gq = p.q + b * p.q_bs + h * DH;
  const float* gk = p.k + b * p.k_bs + h * DH;
  const float* gv = p.v + b * p.v_bs + h * DH;
  const long long bh = static_cast<long long>(b) * p.H + h;
  const int tail_start = min(Lk, NT * 32), tail = Lk - tail_start;

  if (p.tail_only) {
    // Only the last query's context carried a gradient.  If that query was not selected its context was mean(V):
    // dV[j] = dO / Lk, dQ = dK = 0.  If it was, a single probability row is involved: everything is O(Lk * dh).
    const int last = Lq - 1;
    float* dq = bp.dq + b * p.q_bs + h * DH;
    float* dk = bp.dk + b * p.k_bs + h * DH;
    float* dv = bp.dv + b * p.v_bs + h * DH;
    const float inv_lk = 1.f / Lk;
    if (threadIdx.x == 0) s_sel[0] = -1;
    if (threadIdx.x < DH4) reinterpret_cast<float4*>(s_acc)[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(bp.dout + out_offset(p, b, h, last)) + threadIdx.x);
    __syncthreads();
    for (int r = threadIdx.x; r < u; r += THREADS)
      if (p.top[bh * u + r] == last) s_sel[0] = r;
    __syncthreads();
    const bool selected = s_sel[0] >= 0;  // CTA-uniform
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = threadIdx.x; i < (Lq - (selected ? 1 : 0)) * DH4; i += THREADS) {  // dQ rows without a gradient
      const int l = i / DH4, c = i - l * DH4;
      *reinterpret_cast<float4*>(dq + static_cast<long long>(l) * p.q_ls + 4 * c) = zero4;
    }
    if (!selected) {
      for (int i = threadIdx.x; i < Lk * DH4; i += THREADS) {
        const int j = i / DH4, c = i - j * DH4;
        const float4 g4 = reinterpret_cast<const float4*>(s_acc)[c];
        *reinterpret_cast<float4*>(dk + static_cast<long long>(j) * p.k_ls + 4 * c) = zero4;
        *reinterpret_cast<float4*>(dv + static_cast<long long>(j) * p.v_ls + 4 * c) = make_float4(g4.x * inv_lk, g4.y * inv_lk, g4.z * inv_lk, g4.w * inv_lk);
      }
      return;
    }
    stage_rows<DH>(s_k, gk, p.k_ls, Lk);
    if (threadIdx.x < DH4) reinterpret_cast<float4*>(s_q)[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(gq + static_cast<long long>(last) * p.q_ls) + threadIdx.x);
    __syncthreads();
    {
      RegRow<DH> qr, dor;
      qr.load(s_q, true);
      dor.load(s_acc, true);
      for (int j = threadIdx.x; j < Lk; j += THREADS) {  // raw scores and dP of the one row
        RegRow<DH> kt, vt;
        kt.load(s_k + j * DH, true);
        vt.load(gv + static_cast<long long>(j) * p.v_ls, true);
        s_p[j] = kt.dot(qr);
        s_ds[j] = vt.dot(dor);
      }
    }
    __syncthreads();
    const float scale = rsqrtf(static_cast<float>(DH));
    if (warp == 0) {
      warp_softmax_row(s_p, Lk, scale);
      __syncwarp();
      float acc = 0.f;
      for (int j = lane; j < Lk; j += 32) acc = fmaf(s_p[j], s_ds[j], acc);
      acc = warp_sum(acc);
      for (int j = lane; j < Lk; j += 32) s_ds[j] = s_p[j] * (s_ds[j] - acc) * scale;
    }
    __syncthreads();
    if (warp == 0) {  // dQ[last] = dS K
      const float4 o = warp_weighted_column_sum<DH>(s_k, s_ds, Lk);
      if (lane < DH4) *reinterpret_cast<float4*>(dq + static_cast<long long>(last) * p.q_ls + 4 * lane) = o;
    }
    for (int i = threadIdx.x; i < Lk * DH4; i += THREADS) {  // dK[j] = dS[j] q_last, dV[j] = P[j] dO_last
      const int j = i / DH4, c = i - j * DH4;
      const float4 q4 = reinterpret_cast<const float4*>(s_q)[c], g4 = reinterpret_cast<const float4*>(s_acc)[c];
      const float ws = s_ds[j], wp = s_p[j];
      *reinterpret_cast<float4*>(dk + static_cast<long long>(j) * p.k_ls + 4 * c) = make_float4(ws * q4.x, ws * q4.y, ws * q4.z, ws * q4.w);
      *reinterpret_cast<float4*>(dv + static_cast<long long>(j) * p.v_ls + 4 * c) = make_float4(wp * g4.x, wp * g4.y, wp * g4.z, wp * g4.w);
    }
    return;
  }

  stage_rows<DH>(s_q, gq, p.q_ls, Lq);
  stage_rows<DH>(s_k, gk, p.k_ls, Lk);
  for (int i = threadIdx.x; i < Lq * DH4; i += THREADS) {
    const int l = i / DH4, c = i - l * DH4;
    reinterpret_cast<float4*>(s_do)[i] = __ldg(reinterpret_cast<const float4*>(bp.dout + out_offset(p, b, h, l)) + c);
  }
  for (int i = threadIdx.x; i < Lq; i += THREADS) s_sel[i] = -1;
  for (int r = threadIdx.x; r < u; r += THREADS) s_top[r] = p.top[bh * u + r];
  __syncthreads();
  for (int r = threadIdx.x; r < u; r += THREADS) s_sel[s_top[r]] = r;

  // raw scores S[r][j] = q_top_r . k_j, then dP[r][j] = dO_top_r . v_j: key / value rows in registers, selected rows split
  // over the warps, each q / dO row read once per warp
  if (NT > 0) {
    const int chunk_rows = (u + NWARPS - 1) / NWARPS;
    const int r0 = warp * chunk_rows, r1 = min(u, r0 + chunk_rows);
    RegRow<DH> kv[NT > 0 ? NT : 1];
#pragma unroll
    for (int t = 0; t < NT; ++t) kv[t].load(s_k + (lane + 32 * t) * DH, lane + 32 * t < Lk);
#pragma unroll 2
    for (int r = r0; r < r1; ++r) {
      RegRow<DH> x;
      x.load(s_q + s_top[r] * DH, true);
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const float sc = kv[t].dot(x);
        if (lane + 32 * t < Lk) s_p[r * Lk + lane + 32 * t] = sc;
      }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t) kv[t].load(gv + static_cast<long long>(lane + 32 * t) * p.v_ls, lane + 32 * t < Lk);
#pragma unroll 2
    for (int r = r0; r < r1; ++r) {
      RegRow<DH> x;
      x.load(s_do + s_top[r] * DH, true);
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const float dp = kv[t].dot(x);
        if (lane + 32 * t < Lk) s_ds[r * Lk + lane + 32 * t] = dp;
      }
    }
  }
  {  // tail keys, transposed: lane per selected row
    const int rblocks = (u + 31) >> 5;
    for (int it = NWARPS - 1 - warp; it < tail * rblocks; it += NWARPS) {
      const int jt = tail_start + it / rblocks, r = (it % rblocks) * 32 + lane;
      RegRow<DH> kt, vt, x;
      kt.load(s_k + jt * DH, true);
      vt.load(gv + static_cast<long long>(jt) * p.v_ls, true);
      if (r < u) {
        const int qi = s_top[r];
        x.load(s_q + qi * DH, true);
        s_p[r * Lk + jt] = kt.dot(x);
        x.load(s_do + qi * DH, true);
        s_ds[r * Lk + jt] = vt.dot(x);
      }
    }
  }
  __syncthreads();  // s_sel complete as well
  if (warp == NWARPS - 1) warp_column_sum<DH>(s_do, Lq, s_acc, [&](int l) { return s_sel[l] < 0; });

  // P = softmax(scale S), dS = P o (dP - rowsum(P o dP)) * scale, in place, one warp per row
  const float scale = rsqrtf(static_cast<float>(DH));
  for (int r = warp; r < u; r += NWARPS) {
    float* prow = s_p + r * Lk;
    float* drow = s_ds + r * Lk;
    float sc[3], dp[3];
    float mx = -INFINITY;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const int j = lane + 32 * t;
      sc[t] = j < Lk ? prow[j] * scale : -INFINITY;
      dp[t] = j < Lk ? drow[j] : 0.f;
      mx = fmaxf(mx, sc[t]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      sc[t] = (lane + 32 * t < Lk) ? __expf(sc[t] - mx) : 0.f;
      sum += sc[t];
    }
    const float inv = 1.f / warp_sum(sum);
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      sc[t] *= inv;
      acc = fmaf(sc[t], dp[t], acc);
    }
    acc = warp_sum(acc);
#pragma unroll
    for (int t = 0; t < 3; ++t)
      if (lane + 32 * t < Lk) {
        prow[lane + 32 * t] = sc[t];
        drow[lane + 32 * t] = sc[t] * (dp[t] - acc) * scale;
      }
  }
  __syncthreads();

  float* dq = bp.dq + b * p.q_bs + h * DH;
  float* dk = bp.dk + b * p.k_bs + h * DH;
  float* dv = bp.dv + b * p.v_bs + h * DH;
  const float inv_lk = 1.f / Lk;
  if (NT > 0 && warp < 2) {
    // warp 0: dK[j] = sum_r dS[r][j] q_top_r; warp 1: dV[j] = sum_r P[r][j] dO_top_r + fill gradient.  Accumulators live in
    // the registers of the lane that owns key j; each q / dO row is read once.
    const float* w = warp == 0 ? s_ds : s_p;
    const float* xs = warp == 0 ? s_q : s_do;
    RegRow<DH> acc[NT > 0 ? NT : 1];
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[t].zero();
#pragma unroll 2
    for (int r = 0; r < u; ++r) {
      RegRow<DH> x;
      x.load(xs + s_top[r] * DH, true);
#pragma unroll
      for (int t = 0; t < NT; ++t) acc[t].axpy(lane + 32 * t < Lk ? w[r * Lk + lane + 32 * t] : 0.f, x);
    }
    if (warp == 1) {
      RegRow<DH> fill;
      fill.load(s_acc, true);
#pragma unroll
      for (int t = 0; t < NT; ++t) acc[t].axpy(inv_lk, fill);
    }
    float* dst = warp == 0 ? dk : dv;
    const long long ls = warp == 0 ? p.k_ls : p.v_ls;
#pragma unroll
    for (int t = 0; t < NT; ++t)
      if (lane + 32 * t < Lk) acc[t].store(dst + static_cast<long long>(lane + 32 * t) * ls);
  } else {
    // the other warps: dQ in 4-row x 4-channel register tiles, zero rows, and the tail keys' dK / dV
    const int first = NT > 0 ? 64 : 0;
    const int tid2 = threadIdx.x - first, n2 = THREADS - first;
    const int ntiles = ((u + 3) >> 2) * DH4;
    for (int it = tid2; it < ntiles; it += n2) {
      const int r0 = (it / DH4) * 4, c = it % DH4;
      const float* wrow[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) wrow[k] = s_ds + min(r0 + k, u - 1) * Lk;
      float4 o[4];
      tile4_matvec<DH>(wrow, s_k + 4 * c, Lk, o);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (r0 + k < u) *reinterpret_cast<float4*>(dq + static_cast<long long>(s_top[r0 + k]) * p.q_ls + 4 * c) = o[k];
    }
    for (int i = n2 - 1 - tid2; i < Lq * DH4; i += n2) {  // every other row: zero (reversed: the tile threads come last)
      const int l = i / DH4, c = i - l * DH4;
      if (s_sel[l] < 0) *reinterpret_cast<float4*>(dq + static_cast<long long>(l) * p.q_ls + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = n2 - 1 - tid2; i < tail * DH4; i += n2) {
      const int jt = tail_start + i / DH4, c = i % DH4;
      Acc4 ak, av;
#pragma unroll 5
      for (int r = 0; r < u; ++r) {
        const int qi = s_top[r];
        ak.fma(s_ds[r * Lk + jt], s_q + qi * DH + 4 * c);
        av.fma(s_p[r * Lk + jt], s_do + qi * DH + 4 * c);
      }
      const float4 f4 = reinterpret_cast<const float4*>(s_acc)[c];
      float4 vv = av.get();
      vv.x += f4.x * inv_lk; vv.y += f4.y * inv_lk; vv.z += f4.z * inv_lk; vv.w += f4.w * inv_lk;
      *reinterpret_cast<float4*>(dk + static_cast<long long>(jt) * p.k_ls + 4 * c) = ak.get();
      *reinterpret_cast<float4*>(dv + static_cast<long long>(jt) * p.v_ls + 4 * c) = vv;
    }
  }
}

static bool small_path(const RfAttnParams* p) {
  static const bool enabled = [] { const char* e = getenv("RF_ATTN_SMALL"); return !(e && e[0] == '0'); }();
  return enabled && p->mode == RF_ATTN_PROB && (p->dh == 8 || p->dh == 16) && p->Lk <= 96 && p->Lq <= 255 &&
         static_cast<long long>(p->Lq) * p->Lk * 4 <= 40 * 1024;
}
static size_t small_fwd_smem(const RfAttnParams* p) {
  return sizeof(float) * (static_cast<size_t>(p->Lq + p->Lk) * p->dh + round4(p->Lq * p->Lk) + round4(p->Lq) + p->dh + round4(p->u) + round4(p->Lq)) +
         static_cast<size_t>(round4(p->Lq * p->U));
}
static size_t small_bwd_smem(const RfAttnParams* p) {
  return sizeof(float) * (static_cast<size_t>(2 * p->Lq + p->Lk) * p->dh + 2 * round4(p->u * p->Lk) + p->dh + round4(p->u) + round4(p->Lq));
}

static size_t fwd_smem(const RfAttnParams* p, int threads = THREADS) {
  const int pitch = p->dh + 4;
  const int u = p->mode == RF_ATTN_FULL ? p->Lq : p->u;
  return sizeof(float) * (static_cast<size_t>(p->Lq) * pitch + 2 * static_cast<size_t>(p->Lk) * pitch + round4(u * p->Lk) + round4(p->Lq) +
                          round4(p->dh) + threads * 4 + round4(u) + round4(p->Lq));
}
static size_t bwd_smem(const RfAttnParams* p, int threads = THREADS) {
  const int pitch = p->dh + 4;
  const int u = p->mode == RF_ATTN_FULL ? p->Lq : p->u;
  return fwd_smem(p, threads) + sizeof(float) * (static_cast<size_t>(p->Lq) * pitch + round4(u * p->Lk));
}
// 256-thread variant of the generic kernels: launches with few, large problems leave most warp slots of the 148 SMs empty with
// 128-thread CTAs.  Measured at 64 clips (tools/attn_bench.py --generic, profiles/r2_attention_generic_threads.txt), 128 -> 256:
//   video encoder  L = 160, dh = 16 : fwd 74.7 -> 54.8 us, bwd 101.3 -> 80.2 us
//   Informer       L = 40,  dh = 104: fwd 49.1 -> 56.9 us (worse), bwd 74.2 -> 59.3 us
//   Informer dec.  L = 70,  dh = 104: fwd 73.3 -> 92.6 us (worse), bwd 284.9 -> 201.4 us
// hence: forward only for long sequences, backward also for wide heads.  RF_ATTN_WIDE=0 / 1 forces the choice (read per call).
static bool wide_path(const RfAttnParams* p, bool backward) {
  const char* e = getenv("RF_ATTN_WIDE");
  if (e && (e[0] == '0' || e[0] == '1')) return e[0] == '1';
  if (static_cast<long long>(p->B) * p->H > 8 * num_sms()) return false;  // enough CTAs to fill the machine as it is
  const bool long_seq = static_cast<long long>(p->Lq) * p->Lk >= 128 * 128;
  return long_seq || (backward && p->dh >= 64 && p->Lq >= 16);
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
  RF_CHECK_ARG(p->dropout_p >= 0.f && p->dropout_p < 1.f, "%s: dropout_p=%f must be in [0, 1)", who, p->dropout_p);
  RF_CHECK_ARG(p->dropout_p == 0.f || p->mode == RF_ATTN_FULL, "%s: probability dropout exists for full attention only (the reference's "
               "ProbAttention never applies its dropout)", who);
  RF_CHECK_ARG(static_cast<long long>(p->B) * p->H <= 2147483647LL, "%s: too many problems", who);
  return RF_OK;
}

template <typename K>
static int configure(K kernel, size_t smem) {
  if (smem <= 48 * 1024) return RF_OK;
  RF_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), 220 * 1024));
  return RF_OK;
}

}  // namespace attn
}  // namespace rf

#define RF_ATTN_LAUNCH(KERNEL, DHT, ARG, GRID, SMEM, STREAM)                       \
  {                                                                                \
    rc = attn::configure(attn::KERNEL<DHT>, SMEM);                                 \
    if (rc != RF_OK) return rc;                                                    \
    RF_CUDA_OK(rf::launch_pdl(attn::KERNEL<DHT>, dim3(GRID), dim3(attn::THREADS), SMEM, STREAM, ARG)); \
  }
#define RF_ATTN_DISPATCH(KERNEL, ARG, DHVAL, GRID, SMEM, STREAM)                   \
  switch (DHVAL) {                                                                 \
    case 8: RF_ATTN_LAUNCH(KERNEL, 8, ARG, GRID, SMEM, STREAM) break;              \
    case 16: RF_ATTN_LAUNCH(KERNEL, 16, ARG, GRID, SMEM, STREAM) break;            \
    case 104: RF_ATTN_LAUNCH(KERNEL, 104, ARG, GRID, SMEM, STREAM) break;          \
    default: RF_ATTN_LAUNCH(KERNEL, 0, ARG, GRID, SMEM, STREAM) break;             \
  }
#define RF_ATTN_WIDE_LAUNCH(KERNEL, DHT, ARG, GRID, SMEM, STREAM)                  \
  {                                                                                \
    rc = attn::configure(attn::wide::KERNEL<DHT>, SMEM);                           \
    if (rc != RF_OK) return rc;                                                    \
    RF_CUDA_OK(rf::launch_pdl(attn::wide::KERNEL<DHT>, dim3(GRID), dim3(attn::wide::THREADS), SMEM, STREAM, ARG)); \
  }
#define RF_ATTN_WIDE_DISPATCH(KERNEL, ARG, DHVAL, GRID, SMEM, STREAM)              \
  switch (DHVAL) {                                                                 \
    case 16: RF_ATTN_WIDE_LAUNCH(KERNEL, 16, ARG, GRID, SMEM, STREAM) break;       \
    case 104: RF_ATTN_WIDE_LAUNCH(KERNEL, 104, ARG, GRID, SMEM, STREAM) break;     \
    default: RF_ATTN_WIDE_LAUNCH(KERNEL, 0, ARG, GRID, SMEM, STREAM) break;        \
  }

#define RF_ATTN_SMALL_LAUNCH(KERNEL, DHT, NTT, ARG, GRID, SMEM, STREAM)             \
  {                                                                                \
    rc = attn::configure(attn::KERNEL<DHT, NTT>, SMEM);                            \
    if (rc != RF_OK) return rc;                                                    \
    RF_CUDA_OK(rf::launch_pdl(attn::KERNEL<DHT, NTT>, dim3(GRID), dim3(attn::THREADS), SMEM, STREAM, ARG)); \
  }
#define RF_ATTN_SMALL_NT(KERNEL, DHT, ARG, NTVAL, GRID, SMEM, STREAM)              \
  switch (NTVAL) {                                                                 \
    case 0: RF_ATTN_SMALL_LAUNCH(KERNEL, DHT, 0, ARG, GRID, SMEM, STREAM) break;   \
    case 1: RF_ATTN_SMALL_LAUNCH(KERNEL, DHT, 1, ARG, GRID, SMEM, STREAM) break;   \
    case 2: RF_ATTN_SMALL_LAUNCH(KERNEL, DHT, 2, ARG, GRID, SMEM, STREAM) break;   \
    default: RF_ATTN_SMALL_LAUNCH(KERNEL, DHT, 3, ARG, GRID, SMEM, STREAM) break;  \
  }
#define RF_ATTN_SMALL_DISPATCH(KERNEL, ARG, DHVAL, NTVAL, GRID, SMEM, STREAM)      \
  if ((DHVAL) == 8) RF_ATTN_SMALL_NT(KERNEL, 8, ARG, NTVAL, GRID, SMEM, STREAM)    \
  else RF_ATTN_SMALL_NT(KERNEL, 16, ARG, NTVAL, GRID, SMEM, STREAM)

extern "C" int rf_attention_fwd(const RfAttnParams* p, void* stream) {
  using namespace rf;
  RF_CHECK_ARG(p && p->out, "rf_attention_fwd: null pointer");
  int rc = attn::validate(p, "rf_attention_fwd", true);
  if (rc != RF_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (attention_tc_fwd_eligible(p)) return attention_tc_fwd(p, s);
  if (attn::small_path(p)) {
    const size_t smem = attn::small_fwd_smem(p);
    RF_ATTN_SMALL_DISPATCH(attention_small_fwd_kernel, *p, p->dh, attn::small_nt(p->Lk), p->B * p->H, smem, s)
    RF_LAUNCH_OK();
    return RF_OK;
  }
  if (attn::wide_path(p, false) && attn::fwd_smem(p, attn::wide::THREADS) <= 220 * 1024) {
    const size_t smem_w = attn::fwd_smem(p, attn::wide::THREADS);
    RF_ATTN_WIDE_DISPATCH(attention_fwd_kernel, *p, p->dh, p->B * p->H, smem_w, s)
    RF_LAUNCH_OK();
    return RF_OK;
  }
  const size_t smem = attn::fwd_smem(p);
  RF_CHECK_ARG(smem <= 220 * 1024, "rf_attention_fwd: problem needs %zu B of shared memory (> 220 KiB)", smem);
  RF_ATTN_DISPATCH(attention_fwd_kernel, *p, p->dh, p->B * p->H, smem, s)
  RF_LAUNCH_OK();
  return RF_OK;
}

extern "C" int rf_attention_bwd(const RfAttnBwdParams* p, void* stream) {
  using namespace rf;
  RF_CHECK_ARG(p && p->dout && p->dq && p->dk && p->dv, "rf_attention_bwd: null pointer");
  int rc = attn::validate(&p->f, "rf_attention_bwd", false);
  if (rc != RF_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (attn::small_path(&p->f)) {
    const size_t smem = attn::small_bwd_smem(&p->f);
    RF_ATTN_SMALL_DISPATCH(attention_small_bwd_kernel, *p, p->f.dh, attn::small_nt(p->f.Lk), p->f.B * p->f.H, smem, s)
    RF_LAUNCH_OK();
    return RF_OK;
  }
  if (attn::wide_path(&p->f, true) && attn::bwd_smem(&p->f, attn::wide::THREADS) <= 220 * 1024) {
    const size_t smem_w = attn::bwd_smem(&p->f, attn::wide::THREADS);
    RF_ATTN_WIDE_DISPATCH(attention_bwd_kernel, *p, p->f.dh, p->f.B * p->f.H, smem_w, s)
    RF_LAUNCH_OK();
    return RF_OK;
  }
  const size_t smem = attn::bwd_smem(&p->f);
  RF_CHECK_ARG(smem <= 220 * 1024, "rf_attention_bwd: problem needs %zu B of shared memory (> 220 KiB)", smem);
  RF_ATTN_DISPATCH(attention_bwd_kernel, *p, p->f.dh, p->f.B * p->f.H, smem, s)
  RF_LAUNCH_OK();
  return RF_OK;
}
