// ProbSparse self-attention on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), forward.
// replaces: cross_modal_transformer.py:88-166 (ProbAttention._prob_QK / _get_initial_context / _update_context) for the
// Perceive encoders (head dim 16, L <= 80: frame encoder L = 65, gaze encoder L = 40) -- the kernels that were 27 % of the
// round-1 training step as fp32 FMA code bound by the shared-memory pipe.
//
// One CTA = one sequence x one PAIR of heads (32 of the 128 q/k/v columns = one 128 B-wide TMA box):
//   TMA      : Q, K, V tiles [LP rows x 32 cols] (LP = L rounded up to 16) -> 128B-swizzled shared memory, bit-exact fp32
//   split    : Q = Qhi + Qlo, K = Khi + Klo with hi = the 10-bit-mantissa truncation the tensor core sees, lo = remainder
//   tcgen05  : S = Qhi.Khi^T + Qlo.Khi^T + Qhi.Klo^T  (3xTF32: error ~2^-20, i.e. fp32-level scores -- the top-u selection
//              is a discontinuous function of them) for both heads, accumulators in TMEM (M = 128 lanes = query rows)
//   threads  : thread i <-> query row i reads its score row from TMEM, evaluates the sparsity measure M = max - sum / L_K over
//              the SAMPLED keys (multiplicities from the host-drawn index table, kept as byte counts), ranks it against the
//              other rows (top-u, ties -> lower index, as torch.topk resolves them in the oracle tests), soft-maxes the full
//              score row if selected and stores the probabilities as the A operand of the next product
//   tcgen05  : ctx = P.V (tf32 operands rounded to nearest, fp32 accumulate); one extra row of ones yields sum_j V[j], i.e.
//              the mean(V) every unselected query receives (cross_modal_transformer.py:141-147)
//   threads  : context rows back from TMEM, written [B, L, H, dh] (or [B, H, L, dh])
// The whole chain of one (sequence, head pair) problem never leaves the SM; HBM sees q, k, v once and the context once.
#include "tc_common.cuh"

namespace rf {
namespace attn_tc {

using namespace tc;

constexpr int THREADS = 128;
constexpr int DH = 16;
constexpr int CNT_WORDS = 25;  // byte counts per query row: 100 >= LP keys; an odd word pitch keeps the row reads conflict-free
constexpr int MAX_ROWS = 96;

// Optional in-kernel timeline (profiling hook, rf_debug_attn_stamps): the CTA with blockIdx.x == gridDim.x / 2 records clock64
// at its phase boundaries (thread 0).  One global load per CTA otherwise.
__device__ long long* g_attn_stamps = nullptr;

__device__ __forceinline__ float fast_exp2(float x) {  // MUFU.EX2, 2 ulp; -inf -> 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int LP>
struct Layout {
  static constexpr int TILE = LP * 128;        // one k-block (32 fp32 columns) of LP rows
  static constexpr int NKB = (LP + 31) / 32;   // k-blocks along the key axis (P and V^T operands)
  static constexpr int OFF_Q = 0, OFF_K = TILE, OFF_V = 2 * TILE, OFF_QLO = 3 * TILE, OFF_KLO = 4 * TILE;
  static constexpr int OFF_VT = 5 * TILE;                       // [2 heads][NKB][16 rows x 128 B]
  static constexpr int OFF_P = OFF_VT + 2 * NKB * 2048;         // NKB k-blocks of LP rows; the 128-row A operand reads run on
  static constexpr int OFF_CNT = OFF_P + NKB * TILE;            //   into the following regions (finite garbage rows, never used)
  static constexpr int OFF_M = OFF_CNT + MAX_ROWS * CNT_WORDS * 4;
  static constexpr int OFF_TOP = OFF_M + 2 * 128 * 4;
  static constexpr int OFF_MEAN = OFF_TOP + 2 * 128 * 4;
  static constexpr int OFF_BAR = OFF_MEAN + 2 * DH * 4;
  static constexpr int BYTES = OFF_BAR + 64;
  static constexpr int SMEM_BYTES = BYTES + 1024;               // + alignment slack
  static constexpr int TMEM_COLS = 256;                         // 2 x LP score columns + 2 x 16 context columns <= 192
  static_assert(LP % 16 == 0 && LP <= 96, "LP must be a multiple of 16, at most 96");
  static_assert(OFF_BAR - (OFF_P + (NKB - 1) * TILE) >= 16384, "the last P k-block must be readable as 128 rows");
};

template <int LP>
__global__ void __launch_bounds__(THREADS) attention_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                                                                    const __grid_constant__ CUtensorMap tmV, const RfAttnParams p) {
  using LY = Layout<LP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  float* sQ = reinterpret_cast<float*>(smem + LY::OFF_Q);
  float* sK = reinterpret_cast<float*>(smem + LY::OFF_K);
  float* sV = reinterpret_cast<float*>(smem + LY::OFF_V);
  float* sQlo = reinterpret_cast<float*>(smem + LY::OFF_QLO);
  float* sKlo = reinterpret_cast<float*>(smem + LY::OFF_KLO);
  float* sVt = reinterpret_cast<float*>(smem + LY::OFF_VT);
  float* sP = reinterpret_cast<float*>(smem + LY::OFF_P);
  uint32_t* sCnt = reinterpret_cast<uint32_t*>(smem + LY::OFF_CNT);
  float* sM = reinterpret_cast<float*>(smem + LY::OFF_M);
  int* sTop = reinterpret_cast<int*>(smem + LY::OFF_TOP);
  float* sMean = reinterpret_cast<float*>(smem + LY::OFF_MEAN);
  uint64_t* bar_tma = reinterpret_cast<uint64_t*>(smem + LY::OFF_BAR);
  uint64_t* bar_s = bar_tma + 1;
  uint64_t* bar_pv = bar_tma + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tma + 4);

  const int tid = threadIdx.x, warp = tid >> 5;
  long long* stamps = (tid == 0 && blockIdx.x == gridDim.x / 2) ? g_attn_stamps : nullptr;
  if (stamps) stamps[0] = clock64();
  const int L = p.Lq, u = p.u;
  const int pairs = p.H >> 1;
  const int b = blockIdx.x / pairs, hp = blockIdx.x - b * pairs;
  const int row = tid;              // TMEM lane <-> query row (and key row for the V pass)
  const bool active = row < L;
  const int n_row_warps = (L + 1 + 31) >> 5;  // warps whose lanes hold rows 0..L (row L = the ones row)

  if (tid == 0) {
    prefetch_tensormap(&tmQ);
    prefetch_tensormap(&tmK);
    prefetch_tensormap(&tmV);
    mbar_init(bar_tma, 1);
    mbar_init(bar_s, 1);
    mbar_init(&bar_pv[0], 1);
    mbar_init(&bar_pv[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, LY::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();
  if (stamps) stamps[1] = clock64();  // setup + TMEM allocation done

  if (tid == 0) {
    mbar_arrive_expect_tx(bar_tma, 3 * LY::TILE);
    tma_load_2d(sQ, &tmQ, bar_tma, 32 * hp, b * L);
    tma_load_2d(sK, &tmK, bar_tma, 32 * hp, b * L);
    tma_load_2d(sV, &tmV, bar_tma, 32 * hp, b * L);
  }
  // sampled-key multiplicities of this thread's query row (shared by both heads), while the tiles are in flight
  const bool select = p.forced_top == nullptr;
  if (select && active) {
    uint32_t* mine = sCnt + row * CNT_WORDS;
#pragma unroll
    for (int w = 0; w < CNT_WORDS; ++w) mine[w] = 0u;
    const int group = p.idx_group > 0 ? b / p.idx_group : 0;
    const int* idx = p.idx + (static_cast<long long>(group) * L + row) * p.U;
    unsigned char* bytes = reinterpret_cast<unsigned char*>(mine);
    // all index loads first (independent, ~one L2 round trip in total), then the byte increments
    int samp[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) samp[j] = (j < p.U) ? __ldg(idx + j) : -1;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (samp[j] >= 0) bytes[samp[j]] += 1;
    for (int j = 32; j < p.U; ++j) bytes[__ldg(idx + j)] += 1;
  }
  if (stamps) stamps[2] = clock64();  // count table built
  mbar_wait(bar_tma, 0);
  if (stamps) stamps[3] = clock64();  // tiles landed

  // ---- operand preparation ------------------------------------------------------------------------
  // hi / lo split of Q and K (layout-agnostic: element-wise over the swizzled tiles)
  for (int e = tid; e < LY::TILE / 16; e += THREADS) {
    float4 q = reinterpret_cast<float4*>(sQ)[e], k = reinterpret_cast<float4*>(sK)[e];
    const float4 qh = make_float4(trunc_tf32(q.x), trunc_tf32(q.y), trunc_tf32(q.z), trunc_tf32(q.w));
    const float4 kh = make_float4(trunc_tf32(k.x), trunc_tf32(k.y), trunc_tf32(k.z), trunc_tf32(k.w));
    reinterpret_cast<float4*>(sQ)[e] = qh;
    reinterpret_cast<float4*>(sK)[e] = kh;
    reinterpret_cast<float4*>(sQlo)[e] = make_float4(q.x - qh.x, q.y - qh.y, q.z - qh.z, q.w - qh.w);
    reinterpret_cast<float4*>(sKlo)[e] = make_float4(k.x - kh.x, k.y - kh.y, k.z - kh.z, k.w - kh.w);
  }
  // V^T (B operand of P.V): key row `row` -> column `row` of the [16 x LP] tiles of both heads, rounded to tf32; rows >= L hold
  // the next sequence's keys (or TMA zero fill) and are multiplied by zero probabilities: store zeros so nothing can leak
  if (row < LP) {
    const int kb = row >> 5, jj = row & 31;
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active) v = *reinterpret_cast<const float4*>(sV + row * 32 + ((q4 ^ (row & 7)) << 2));
      const int hh = q4 >> 2, c0 = (q4 & 3) << 2;  // head of the pair, first channel of this chunk
      float* vt = sVt + (hh * LY::NKB + kb) * 512;  // 16 rows x 32 floats
      vt[swz(c0 + 0, jj)] = round_tf32(v.x);
      vt[swz(c0 + 1, jj)] = round_tf32(v.y);
      vt[swz(c0 + 2, jj)] = round_tf32(v.z);
      vt[swz(c0 + 3, jj)] = round_tf32(v.w);
    }
  }
  fence_proxy_async_smem();
  __syncthreads();

  if (stamps) stamps[4] = clock64();  // hi/lo split + V^T done
  // ---- S = Q.K^T for both heads (3xTF32) -------------------------------------------------------------
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = idesc_tf32(128, LP);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const uint32_t d = tmem_base + static_cast<uint32_t>(hh * LP);
      const uint32_t qh = smem_u32(sQ) + hh * 64, kh = smem_u32(sK) + hh * 64;
      const uint32_t ql = smem_u32(sQlo) + hh * 64, kl = smem_u32(sKlo) + hh * 64;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) umma_tf32(d, kmajor_desc(qh + kk * 32), kmajor_desc(kh + kk * 32), idesc, kk);
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) umma_tf32(d, kmajor_desc(ql + kk * 32), kmajor_desc(kh + kk * 32), idesc, 1u);
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) umma_tf32(d, kmajor_desc(qh + kk * 32), kmajor_desc(kl + kk * 32), idesc, 1u);
    }
    umma_commit(bar_s);
  }
  mbar_wait(bar_s, 0);
  tc_fence_after();
  if (stamps) stamps[5] = clock64();  // scores ready

  const float scale = rsqrtf(static_cast<float>(DH));
  int sel_rank[2] = {-1, -1};
  const long long bh0 = static_cast<long long>(b) * p.H + 2 * hp;

#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    float s[LP];
    if (warp < n_row_warps) {
#pragma unroll
      for (int c = 0; c < LP / 16; ++c) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(hh * LP + c * 16), r);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) s[c * 16 + j] = __uint_as_float(r[j]);
      }
    }
    // sparsity measure over the sampled keys (cross_modal_transformer.py:97-100): M = max_j s - sum_j s / L_K.
    // Four independent (max, sum) chains: the loop is latency-bound otherwise (one warp per scheduler).
    if (select) {
      if (active) {
        const uint32_t* mine = sCnt + row * CNT_WORDS;
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j4 = 0; j4 < LP / 4; ++j4) {
          const uint32_t w = mine[j4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int j = j4 * 4 + t;
            const uint32_t c = (w >> (8 * t)) & 0xFFu;
            if (j < L && c != 0u) {
              mx[t] = fmaxf(mx[t], s[j]);
              sum[t] = fmaf(static_cast<float>(c), s[j], sum[t]);
            }
          }
        }
        const float mval = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) - ((sum[0] + sum[1]) + (sum[2] + sum[3])) / p.Lk;
        sM[hh * 128 + row] = mval;
        if (p.measure) p.measure[(bh0 + hh) * L + row] = mval;
      } else if (row < 128) {
        sM[hh * 128 + row] = -INFINITY;  // padding rows never beat a real one (the rank loop reads whole float4 groups)
      }
    } else if (tid < u) {
      sTop[hh * 128 + tid] = p.forced_top[(bh0 + hh) * u + tid];
    }
    __syncthreads();
    if (stamps) stamps[6 + 4 * hh] = clock64();  // measure done
    // top-u: rank = number of rows that beat this one (ties -> lower index first)
    int rank = -1;
    if (active) {
      if (select) {
        const float mi = sM[hh * 128 + row];
        int r0 = 0, r1 = 0, r2 = 0, r3 = 0;
        const float4* m4 = reinterpret_cast<const float4*>(sM + hh * 128);
#pragma unroll 4
        for (int j4 = 0; j4 < (L + 3) >> 2; ++j4) {
          const float4 m = m4[j4];
          const int j = j4 << 2;
          r0 += (m.x > mi) || (m.x == mi && j < row);
          r1 += (m.y > mi) || (m.y == mi && j + 1 < row);
          r2 += (m.z > mi) || (m.z == mi && j + 2 < row);
          r3 += (m.w > mi) || (m.w == mi && j + 3 < row);
        }
        const int r = (r0 + r1) + (r2 + r3);
        if (r < u) rank = r;
      } else {
        for (int r = 0; r < u; ++r)
          if (sTop[hh * 128 + r] == row) rank = r;
      }
    }
    sel_rank[hh] = rank;
    // soft-max of the selected rows (cross_modal_transformer.py:158-162); probabilities rounded to tf32 (nearest)
    float inv = 0.f;
    if (rank >= 0) {
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
      for (int j = 0; j < LP; j += 4) {
        if (j < L) m0 = fmaxf(m0, s[j]);
        if (j + 1 < L) m1 = fmaxf(m1, s[j + 1]);
        if (j + 2 < L) m2 = fmaxf(m2, s[j + 2]);
        if (j + 3 < L) m3 = fmaxf(m3, s[j + 3]);
      }
      const float k2 = scale * 1.4426950408889634f;                       // exp(x * scale) = exp2(x * scale * log2 e)
      const float mraw = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * k2;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int j = 0; j < LP; j += 4) {
        const float e0 = (j < L) ? fast_exp2(fmaf(s[j], k2, -mraw)) : 0.f;
        const float e1 = (j + 1 < L) ? fast_exp2(fmaf(s[j + 1], k2, -mraw)) : 0.f;
        const float e2 = (j + 2 < L) ? fast_exp2(fmaf(s[j + 2], k2, -mraw)) : 0.f;
        const float e3 = (j + 3 < L) ? fast_exp2(fmaf(s[j + 3], k2, -mraw)) : 0.f;
        s[j] = e0; s[j + 1] = e1; s[j + 2] = e2; s[j + 3] = e3;
        a0 += e0; a1 += e1; a2 += e2; a3 += e3;
      }
      inv = 1.f / ((a0 + a1) + (a2 + a3));
    }
    if (stamps) stamps[7 + 4 * hh] = clock64();  // rank + softmax done
    if (hh == 1) {  // the P tile is shared by the two heads: head 0's P.V must have consumed it
      mbar_wait(&bar_pv[0], 0);
      tc_fence_after();
    }
    if (select && rank >= 0) sTop[hh * 128 + rank] = row;
    if (rank >= 0 || row == L) {
#pragma unroll
      for (int c4 = 0; c4 < LP / 4; ++c4) {
        float4 v;
        if (row == L) {
          v = make_float4(c4 * 4 + 0 < L ? 1.f : 0.f, c4 * 4 + 1 < L ? 1.f : 0.f, c4 * 4 + 2 < L ? 1.f : 0.f, c4 * 4 + 3 < L ? 1.f : 0.f);
        } else {
          v = make_float4(round_tf32(s[c4 * 4 + 0] * inv), round_tf32(s[c4 * 4 + 1] * inv), round_tf32(s[c4 * 4 + 2] * inv),
                          round_tf32(s[c4 * 4 + 3] * inv));
        }
        const int kb = c4 >> 3, q16 = c4 & 7;
        *reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(sP) + kb * LY::TILE + row * 128 + ((q16 ^ (row & 7)) << 4)) = v;
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (stamps) stamps[8 + 4 * hh] = clock64();  // P written
    if (p.top && tid < u) p.top[(bh0 + hh) * u + tid] = sTop[hh * 128 + tid];
    // ---- ctx = P.V (+ the ones row: column sums of V) ------------------------------------------------
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = idesc_tf32(128, DH);
      const uint32_t d = tmem_base + static_cast<uint32_t>(2 * LP + hh * DH);
#pragma unroll
      for (int kk = 0; kk < LP / 8; ++kk) {
        const uint32_t a = smem_u32(sP) + (kk >> 2) * LY::TILE + (kk & 3) * 32;
        const uint32_t bb = smem_u32(sVt) + (hh * LY::NKB + (kk >> 2)) * 2048 + (kk & 3) * 32;
        umma_tf32(d, kmajor_desc(a), kmajor_desc(bb), idesc, kk);
      }
      umma_commit(&bar_pv[hh]);
    }
    if (stamps) stamps[9 + 4 * hh] = clock64();  // P.V issued
  }

  // ---- context rows: P.V for the selected queries, mean(V) for the others ------------------------------
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    mbar_wait(&bar_pv[hh], 0);
    tc_fence_after();
    uint32_t r[16];
    if (warp < n_row_warps) {
      tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(2 * LP + hh * DH), r);
      tmem_wait_ld();
    }
    if (row == L) {
      const float inv_l = 1.0f / p.Lk;
#pragma unroll
      for (int c = 0; c < DH; ++c) sMean[hh * DH + c] = __uint_as_float(r[c]) * inv_l;
    }
    __syncthreads();
    if (active) {
      const int h = 2 * hp + hh;
      const long long off = p.out_layout == RF_LAYOUT_BLHD ? ((static_cast<long long>(b) * L + row) * p.H + h) * DH
                                                           : ((static_cast<long long>(b) * p.H + h) * L + row) * DH;
      float4* dst = reinterpret_cast<float4*>(p.out + off);
      if (sel_rank[hh] >= 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          dst[c] = make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]), __uint_as_float(r[4 * c + 2]), __uint_as_float(r[4 * c + 3]));
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) dst[c] = *reinterpret_cast<const float4*>(sMean + hh * DH + 4 * c);
      }
    }
  }

  if (stamps) stamps[14] = clock64();  // contexts written
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, LY::TMEM_COLS);
  if (stamps) stamps[15] = clock64();
}

template <int LP>
static int launch_fwd(const RfAttnParams* p, cudaStream_t stream) {
  using LY = Layout<LP>;
  CUtensorMap tmQ, tmK, tmV;
  const long long rows = static_cast<long long>(p->B) * p->Lq;
  const long long cols = static_cast<long long>(p->H) * p->dh;
  int rc = make_map_f32(&tmQ, p->q, cols, rows, p->q_ls, LP);
  if (rc != RF_OK) return rc;
  rc = make_map_f32(&tmK, p->k, cols, rows, p->k_ls, LP);
  if (rc != RF_OK) return rc;
  rc = make_map_f32(&tmV, p->v, cols, rows, p->v_ls, LP);
  if (rc != RF_OK) return rc;
  RF_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(attention_tc_fwd_kernel<LP>), LY::SMEM_BYTES));
  RF_CUDA_OK(launch_pdl(attention_tc_fwd_kernel<LP>, dim3(p->B * (p->H / 2)), dim3(THREADS), LY::SMEM_BYTES, stream, tmQ, tmK, tmV, *p));
  return RF_OK;
}

}  // namespace attn_tc

// Eligibility of the tensor-core forward: unmasked ProbSparse self-attention, head dim 16, an even number of heads, L <= 79,
// contiguous sequences of rows (batch stride = L x row stride), 16 B-aligned rows, no last-query-only hint.
bool attention_tc_fwd_eligible(const RfAttnParams* p) {
  // Opt-in (read per call): RF_ATTN_TC=1.  The kernel is exact on the selections and tested, but its thread-per-row epilogue
  // keeps only two of the four warp schedulers busy and is latency-bound: 304 us against 224 us for the fp32 FMA kernel on the
  // frame-encoder problem (profiles/r2_attention_tc_microbench.txt), so the FMA kernels stay the default.
  const char* e = getenv("RF_ATTN_TC");
  if (!(e && e[0] == '1')) return false;
  if (p->mode != RF_ATTN_PROB || p->dh != attn_tc::DH || (p->H & 1) || p->Lq != p->Lk || p->Lq > 79 || p->Lq < 8) return false;  // L + 1 rows (the ones row) must fit the LP-row P tile
  if (p->tail_only || p->dropout_p != 0.f) return false;
  if (p->q_bs != static_cast<long long>(p->Lq) * p->q_ls || p->k_bs != static_cast<long long>(p->Lk) * p->k_ls ||
      p->v_bs != static_cast<long long>(p->Lk) * p->v_ls)
    return false;
  if ((p->q_ls | p->k_ls | p->v_ls) & 3) return false;
  if ((reinterpret_cast<uintptr_t>(p->q) | reinterpret_cast<uintptr_t>(p->k) | reinterpret_cast<uintptr_t>(p->v) |
       reinterpret_cast<uintptr_t>(p->out)) & 15)
    return false;
  if (p->U > 100 || p->u > 128) return false;
  return true;
}

}  // namespace rf

extern "C" int rf_debug_attn_stamps(long long* device_buffer) {
  using namespace rf;
  RF_CUDA_OK(cudaMemcpyToSymbol(attn_tc::g_attn_stamps, &device_buffer, sizeof(device_buffer)));
  return RF_OK;
}

namespace rf {

int attention_tc_fwd(const RfAttnParams* p, cudaStream_t stream) {
  if (p->Lq < 48) return attn_tc::launch_fwd<48>(p, stream);
  return attn_tc::launch_fwd<80>(p, stream);
}

}  // namespace rf
