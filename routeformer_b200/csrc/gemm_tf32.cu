// Tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory ->
// tcgen05.mma.kind::tf32 (or kind::f16 for fp16 operands), issued by one elected thread -> fp32 accumulators in
// TMEM -> tcgen05.ld epilogue -> swizzled shared memory -> TMA bulk store / reduce-add.
//
//   C[M,N] (+)= epilogue( sum_k A[m,k] * B[n,k] )
//
// Either operand may be K-major (reduction dim contiguous) or MN-major (the other dim contiguous), so the
// same kernels serve forward (x W^T), dgrad (dy W) and wgrad (dy^T x, split over the long reduction and
// reduce-added into the gradient arena).  fp32 operands stay fp32 in HBM; the TMA descriptor (TFLOAT32 type)
// rounds them to tf32 on the way into shared memory, accumulation is fp32.
//
// Two kernels share the pipeline pieces and the epilogue (RowEpilogue):
//  * gemm_tf32_kernel<BLOCK_N, STAGES>: one 128 x BLOCK_N tile per CTA, 128 threads.
//      warp 0 / lane 0 : TMA producer over a STAGES-deep full/empty mbarrier ring
//      warp 1 / lane 0 : tcgen05.mma issuer, tcgen05.commit releases ring slots and signals the epilogue
//      warp 2          : TMEM allocate / free;   all 4 warps: epilogue, thread t owns accumulator row t
//    3 stages x 2 CTAs per SM for long reductions on big grids, 6 stages x 1 CTA per SM below one wave.
//  * gemm_tf32_persistent_kernel<BLOCK_N>: one CTA per SM loops over tiles (<= 12 k-blocks per tile):
//      producer warp, MMA warp, 8 epilogue warps on a double-buffered TMEM accumulator, and a reducer warp
//      that sums the A tiles column-wise from shared memory (bias gradients ride on the dgrad GEMMs).
// Host side (rf_gemm_tf32): tensor maps, kernel choice, automatic split-K (wgrad; few-tile / long-K calls
// with a linear epilogue), programmatic dependent launch.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"

namespace rf {
namespace gemm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;  // 32 fp32 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 8;    // K of one tcgen05.mma.kind::tf32
constexpr int STAGES_SHALLOW = 3;  // 2 CTAs per SM (large grids: bytes in flight come from CTA count)
constexpr int STAGES_DEEP = 6;     // 1 CTA per SM (grids below one wave: the K loop of each CTA is latency-bound)
constexpr int NUM_THREADS = 128;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;  // 16 KiB
constexpr int GROUP_BYTES = 32 * BLOCK_K * 4;   // one MN-major box: 32 k-rows x 128 B

template <int BLOCK_N, int STAGES = STAGES_SHALLOW>
struct Tile {
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 128 /*barriers*/;
};

struct Args {
  float* C; long long ldc;
  int M, N, K;
  int a_mn, b_mn;
  const float* bias;
  const float* rowadd; int rowadd_period; long long ld_rowadd;
  const float* residual; long long ld_res;
  int act;
  float* preact; long long ld_pre;
  const float* dact_aux; long long ld_aux; int dact;
  int accumulate;
  int kb_total, kb_per_split;
  int split_epilogue;  // split-K of a GEMM with a LINEAR epilogue: bias / rowadd / residual are added by split 0 only
  float* colsum_a;  // persistent kernel: += column sums of A taken from the operand tiles in shared memory
  int f16;  // operands are 16-bit (kind::f16, 64 elements per 128 B k-block row) instead of fp32 read as tf32: 1 = fp16, 2 = bf16
  int c_bf16;  // C holds bf16 (direct stores only)
  int group_in, group_out, row_offset;
  int round_f16;
  int tma_store;  // 1: epilogue stages 128x32 chunks in (swizzled) smem and writes them with TMA bulk stores
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) __trap();  // ~4 s
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (Blackwell version 1).
//  K-major : SWIZZLE_128B (TMA SWIZZLE_128B).  Rows of 128 B (32 tf32 along K); 8-row groups SBO = 1024 B
//            apart; LBO unused (encoded 1).
//  MN-major: for 32-bit operands the only legal layout is SWIZZLE_128B_BASE32B (Swizzle<2,5,2>: 32 B chunks
//            XOR-ed over 4-row atoms; TMA SWIZZLE_128B_ATOM_32B writes exactly this).  128 B = 32 elements
//            along M/N; consecutive k-rows are 128 B apart; SBO = stride between 4-row k atoms (512 B),
//            LBO = stride between 32-element M/N groups (one TMA box = 32 rows x 128 B = 4096 B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (sm_100)
  d |= static_cast<uint64_t>(layout_type) << 61;  // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == RF_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == RF_ACT_GELU) return gelu_erf(v);
  return v;
}

// Optional in-kernel timeline (profiling hook): when rf_debug_gemm_stamps() has installed a buffer, CTA (0,0,0) records
// %globaltimer at its phase boundaries.  One extra global load per CTA otherwise.
__device__ unsigned long long* g_stamps = nullptr;
// Bottleneck probe of the persistent kernel (rf_debug_gemm_probe; TIMING ONLY, results are wrong while a bit is set):
//   1 = the epilogue stages its chunks but issues no bulk store      2 = the epilogue only drains TMEM (no math, no staging, no store)
//   4 = the B operand is loaded for the first tile of a CTA only      8 = the MMA issuer commits without issuing MMAs
__device__ int g_probe = 0;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Epilogue math of one accumulator row (thread <-> row), applied to one 32-column chunk at a time.  Every option is tested
// ONCE per chunk (not per element) and operand rows are fetched as 8 independent 16 B loads.
struct RowEpilogue {
  bool row_ok, al16;
  const float* rowadd_row;
  const float* res_row;
  const float* aux_row;
  float* pre_row;
  float* c_row;

  bool add_linear;  // this CTA adds the bias / positional / residual terms (all CTAs, or split 0 of a split linear epilogue)
  __device__ __forceinline__ void init(const Args& a, int m, int z = 0) {
    row_ok = m < a.M;
    add_linear = !a.split_epilogue || z == 0;
    long long out_row = m;
    if (a.group_in > 0) out_row = static_cast<long long>(m / a.group_in) * a.group_out + (m % a.group_in) + a.row_offset;
    c_row = a.c_bf16 ? reinterpret_cast<float*>(reinterpret_cast<__nv_bfloat16*>(a.C) + out_row * a.ldc) : a.C + out_row * a.ldc;
    rowadd_row = (a.rowadd && add_linear) ? a.rowadd + static_cast<long long>(m % a.rowadd_period) * a.ld_rowadd : nullptr;
    res_row = (a.residual && add_linear) ? a.residual + static_cast<long long>(m) * a.ld_res : nullptr;
    pre_row = a.preact ? a.preact + static_cast<long long>(m) * a.ld_pre : nullptr;
    aux_row = a.dact ? a.dact_aux + static_cast<long long>(m) * a.ld_aux : nullptr;
    al16 = ((reinterpret_cast<uintptr_t>(a.bias) | reinterpret_cast<uintptr_t>(a.rowadd) | reinterpret_cast<uintptr_t>(a.residual) |
             reinterpret_cast<uintptr_t>(a.dact_aux)) & 15) == 0 &&
           ((a.ld_rowadd | a.ld_res | a.ld_aux) & 3) == 0;
  }
  static __device__ __forceinline__ void add_row(float (&x)[32], const float* src, int ncols, bool vec, bool ro) {
    if (vec) {
      float4 t[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) t[q] = ro ? __ldg(reinterpret_cast<const float4*>(src) + q) : reinterpret_cast<const float4*>(src)[q];
#pragma unroll
      for (int q = 0; q < 8; ++q) { x[4 * q] += t[q].x; x[4 * q + 1] += t[q].y; x[4 * q + 2] += t[q].z; x[4 * q + 3] += t[q].w; }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) x[j] += src[j];
    }
  }
  // The per-row side operand (activation-derivative input if any, else the residual) does not depend on the accumulator:
  // the persistent kernel issues its loads BEFORE waiting for the MMAs of the tile, hiding the global-memory round trip.
  __device__ __forceinline__ bool side_prefetch(const Args& a, int nb, int ncols, float4 (&s)[8]) const {
    const float* row = a.dact ? aux_row : res_row;
    if (!row || !row_ok || !(al16 && ncols == 32)) return false;
#pragma unroll
    for (int q = 0; q < 8; ++q) s[q] = reinterpret_cast<const float4*>(row + nb)[q];
    return true;
  }
  __device__ __forceinline__ void apply(const Args& a, float (&x)[32], float (&pre)[32], int nb, int ncols) const {
    float4 none[8];
    apply(a, x, pre, nb, ncols, false, none);
  }
  __device__ __forceinline__ void apply(const Args& a, float (&x)[32], float (&pre)[32], int nb, int ncols, bool have_side,
                                        const float4 (&side)[8]) const {
    const bool vec = al16 && ncols == 32;
    const bool side_is_res = have_side && !a.dact;
    const bool side_is_aux = have_side && a.dact;
    if (row_ok) {
      if (a.bias && add_linear) add_row(x, a.bias + nb, ncols, vec, true);
      if (rowadd_row) add_row(x, rowadd_row + nb, ncols, vec, true);
      if (side_is_res) {
#pragma unroll
        for (int q = 0; q < 8; ++q) { x[4 * q] += side[q].x; x[4 * q + 1] += side[q].y; x[4 * q + 2] += side[q].z; x[4 * q + 3] += side[q].w; }
      } else if (res_row) {
        add_row(x, res_row + nb, ncols, vec, false);
      }
    }
    if (a.preact && a.act != RF_ACT_GELU_SAVE_GRAD) {
#pragma unroll
      for (int j = 0; j < 32; ++j) pre[j] = x[j];
    }
    if (a.act == RF_ACT_RELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.0f);
    } else if (a.act == RF_ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = gelu_erf(x[j]);
    } else if (a.act == RF_ACT_GELU_SAVE_GRAD) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = gelu_erf_with_grad(x[j], pre[j]);
    }
    if (a.dact && row_ok) {
      float t[32];
      if (side_is_aux) {
#pragma unroll
        for (int q = 0; q < 8; ++q) { t[4 * q] = side[q].x; t[4 * q + 1] = side[q].y; t[4 * q + 2] = side[q].z; t[4 * q + 3] = side[q].w; }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = 0.0f;
        add_row(t, aux_row + nb, ncols, vec, false);
      }
      if (a.dact == RF_ACT_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = t[j] > 0.0f ? x[j] : 0.0f;
      } else if (a.dact == RF_DACT_SAVED) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] *= t[j];
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] *= gelu_erf_grad(t[j]);
      }
    }
    if (a.round_f16) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = __half2float(__float2half_rn(x[j]));
    }
  }
  // direct (non-TMA) store of one chunk
  __device__ __forceinline__ void store_direct(const Args& a, const float (&v)[32], const float (&pre)[32], int nb, int ncols) const {
    if (!row_ok) return;
    if (pre_row) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) pre_row[nb + j] = pre[j];
    }
    const bool vec_ok = ((a.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.C) & 15) == 0) && !a.accumulate;
    if (a.c_bf16) {  // c_row was formed with ldc counted in bf16 elements (init)
      __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(c_row) + nb;
      if (ncols == 32 && ((a.ldc & 7) == 0) && ((reinterpret_cast<uintptr_t>(a.C) & 15) == 0) && ((nb & 7) == 0)) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 u;
          __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]), t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
          u.x = *reinterpret_cast<uint32_t*>(&t0); u.y = *reinterpret_cast<uint32_t*>(&t1);
          u.z = *reinterpret_cast<uint32_t*>(&t2); u.w = *reinterpret_cast<uint32_t*>(&t3);
          *reinterpret_cast<uint4*>(crow + j) = u;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncols) crow[j] = __float2bfloat16_rn(v[j]);
      }
      return;
    }
    if (a.accumulate) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) atomicAdd(c_row + nb + j, v[j]);
    } else if (vec_ok && ncols == 32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(c_row + nb + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) c_row[nb + j] = v[j];
    }
  }
  // 128B-swizzled staging of one chunk row for the TMA store (16 B chunk index XOR (row mod 8))
  static __device__ __forceinline__ void stage_row(uint8_t* buf, int row_in_tile, const float (&v)[32]) {
    float* row = reinterpret_cast<float*>(buf + row_in_tile * 128);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int slot = (j ^ (row_in_tile & 7)) << 2;
      *reinterpret_cast<float4*>(row + slot) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
};

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, STAGES == STAGES_SHALLOW ? 2 : 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmP, const Args a) {
  using T = Tile<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * T::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // N-tiles vary fastest in launch order: the CTAs that share one 128-row A panel are resident together, so the panel is read
  // from HBM once and served from L2 to its siblings (with M fastest a 51 MB activation matrix was re-read ~2.5x, ncu r1 v5)
  const int n0 = blockIdx.x * BLOCK_N;
  const int m0 = blockIdx.y * BLOCK_M;
  const int kb_begin = blockIdx.z * a.kb_per_split;
  const int kb_end = min(a.kb_total, kb_begin + a.kb_per_split);
  const int nkb = kb_end - kb_begin;
  unsigned long long* stamps = (blockIdx.x | blockIdx.y | blockIdx.z) == 0 ? g_stamps : nullptr;
  if (stamps && threadIdx.x == 0) stamps[0] = gtime();

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    if (a.tma_store) prefetch_tensormap(&tmC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();     // barriers, tensor-map prefetch and the TMEM allocation above overlap the tail of the previous kernel
  pdl_trigger();
  if (stamps && threadIdx.x == 0) stamps[1] = gtime();

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], T::STAGE_BYTES);
        uint8_t* sa = smem + s * T::STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        const int k0 = (kb_begin + i) * (a.f16 ? 2 * BLOCK_K : BLOCK_K);  // a 128 B row holds 32 fp32 or 64 fp16
        if (!a.a_mn) {
          tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);
        } else {
#pragma unroll
          for (int g = 0; g < BLOCK_M / 32; ++g) tma_load_2d(sa + g * GROUP_BYTES, &tmA, &full_bar[s], m0 + 32 * g, k0);
        }
        if (!a.b_mn) {
          tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);
        } else {
#pragma unroll
          for (int g = 0; g < BLOCK_N / 32; ++g) tma_load_2d(sb + g * GROUP_BYTES, &tmB, &full_bar[s], n0 + 32 * g, k0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D fp32; A/B format tf32 (2), f16 (0) or bf16 (1); major-ness; N >> 3; M >> 4
      const uint32_t fmt = a.f16 == 1 ? 0u : (a.f16 == 2 ? 1u : 2u);
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) |
                             (static_cast<uint32_t>(a.a_mn) << 15) | (static_cast<uint32_t>(a.b_mn) << 16) |
                             (static_cast<uint32_t>(BLOCK_N >> 3) << 17) | (static_cast<uint32_t>(BLOCK_M >> 4) << 24);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (stamps && i == 0) stamps[2] = gtime();
        const uint32_t a_base = smem_u32(smem + s * T::STAGE_BYTES);
        const uint32_t b_base = a_base + A_BYTES;
#pragma unroll
        for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
          const uint64_t adesc = a.a_mn ? make_smem_desc(a_base + kk * 1024, GROUP_BYTES, 512, 1)
                                        : make_smem_desc(a_base + kk * UMMA_K * 4, 16, 1024, 2);
          const uint64_t bdesc = a.b_mn ? make_smem_desc(b_base + kk * 1024, GROUP_BYTES, 512, 1)
                                        : make_smem_desc(b_base + kk * UMMA_K * 4, 16, 1024, 2);
          // one MMA consumes 32 B of K per row either way: 8 tf32 or 16 fp16
          if (a.f16) umma_f16(tmem_base, adesc, bdesc, idesc, (i | kk) != 0 ? 1u : 0u);
          else umma_tf32(tmem_base, adesc, bdesc, idesc, (i | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees the ring slot once these MMAs have read it
      }
      umma_commit(tmem_full_bar);    // accumulator complete
      if (stamps) stamps[3] = gtime();
    }
    __syncwarp();
  }

  // ---------------- epilogue: thread t <-> accumulator row t ---------------------------------
  mbar_wait(tmem_full_bar, 0);
  tc_fence_after();
  if (stamps && threadIdx.x == 64) stamps[4] = gtime();
  const int m = m0 + warp * 32 + lane;
  const bool row_ok = m < a.M;
  long long out_row = m;
  if (a.group_in > 0) out_row = static_cast<long long>(m / a.group_in) * a.group_out + (m % a.group_in) + a.row_offset;
  RowEpilogue re;
  re.init(a, m, blockIdx.z);

  if (a.tma_store) {
    // ---- staged path: 128x32 fp32 chunks -> 128B-swizzled smem (the idle pipeline stages) -> TMA bulk store --------
    // The output (and pre-activation) tiles leave the SM as full 128 B lines written by the TMA engine instead of
    // 32 rows x 16 B per store instruction; TMA also clips rows >= M / cols >= N.
    constexpr int NPAIR = (STAGES * T::STAGE_BYTES) / 32768;  // (out, preact) buffer pairs of 16 KiB each
    const int row_in_tile = warp * 32 + lane;
    int ci = 0;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 32, ++ci) {
      const int nb = n0 + c;
      if (nb >= a.N) break;  // CTA-uniform
      uint8_t* obuf = smem + (ci % NPAIR) * 32768;
      uint8_t* pbuf = obuf + 16384;
      if (ci >= NPAIR) {  // the buffer pair is being re-used: its previous bulk store must have finished reading smem
        if (threadIdx.x == 0) bulk_wait_group_read<NPAIR - 1>();
        __syncthreads();
      }
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(c), r);
      tmem_wait_ld();
      float v[32], pre[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      re.apply(a, v, pre, nb, min(32, a.N - nb));
      float* orow = reinterpret_cast<float*>(obuf + row_in_tile * 128);
      float* prow = reinterpret_cast<float*>(pbuf + row_in_tile * 128);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int slot = (j ^ (row_in_tile & 7)) << 2;  // SWIZZLE_128B: 16 B chunk index XOR (row mod 8)
        *reinterpret_cast<float4*>(orow + slot) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      if (a.preact) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int slot = (j ^ (row_in_tile & 7)) << 2;
          *reinterpret_cast<float4*>(prow + slot) = make_float4(pre[4 * j], pre[4 * j + 1], pre[4 * j + 2], pre[4 * j + 3]);
        }
      }
      fence_proxy_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) {
        if (a.accumulate) tma_reduce_add_2d(&tmC, obuf, nb, m0);  // split-K wgrad: fp32 add performed by the TMA / L2, 128 B lines
        else if (a.group_in > 0) tma_store_3d(&tmC, obuf, nb, 0, m0 / a.group_in);
        else tma_store_2d(&tmC, obuf, nb, m0);
        if (a.preact) tma_store_2d(&tmP, pbuf, nb, m0);
        bulk_commit_group();
      }
    }
    if (threadIdx.x == 0) bulk_wait_group_read<0>();
  } else {
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 32) {
      const int nb = n0 + c;
      if (nb >= a.N) break;  // warp-uniform
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(c), r);
      tmem_wait_ld();
      if (!row_ok) continue;
      const int ncols = min(32, a.N - nb);
      float v[32], pre[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      re.apply(a, v, pre, nb, ncols);
      re.store_direct(a, v, pre, nb, ncols);
    }
  }

  if (stamps && threadIdx.x == 64) stamps[5] = gtime();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, BLOCK_N);
  if (stamps && threadIdx.x == 64) stamps[6] = gtime();
}

// ---------------------------------------------------------------------------------------------
// persistent, warp-specialised variant (default)
// ---------------------------------------------------------------------------------------------
// One CTA per SM loops over output tiles (N-tiles fastest).  Roles: warp 0 = TMA producer running ahead across tile
// boundaries over a 3-stage ring; warp 1 = MMA issuer ping-ponging between two TMEM accumulators; warps 2-9 = epilogue
// (TMEM -> registers -> math -> swizzled smem -> TMA store / reduce-add).  The epilogue of tile j overlaps the operand loads
// and MMAs of tiles j+1, j+2: for the K=128 layers of the frame encoder (4 k-blocks per tile) the loads never drain.
// Ring depth / staging layout (template parameters of the kernel):
//   launches that also write the pre-activation output (FFN1 forward: gelu'(x)) stage (out, preact) pairs of 2 x 16 KiB and run
//   a 3-stage operand ring; every other launch can stage 16 KiB out chunks only, which frees 64 KiB for a 5-stage ring
//   (RF_GEMM_DEEP_RING=1; measured neutral, see the launch code).
constexpr int P_STAGES_PREACT = 3;
constexpr int P_STAGES_PLAIN = 5;
constexpr int P_EPI_WARPS = 8;         // two warps per TMEM lane quarter: each takes one half of the tile's columns
constexpr int P_THREADS = 64 + 32 * P_EPI_WARPS + 32;  // + one warp that sums the A tiles column-wise (bias gradients)
constexpr int P_REDUCER_WARP = 2 + P_EPI_WARPS;

template <int BLOCK_N, int STAGES, int PAIR_BYTES>
struct PTile {
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = RING_BYTES + 4 * PAIR_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void epi_bar_sync(int half) { asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory"); }

template <int BLOCK_N, int STAGES, int PAIR_BYTES>
__global__ void __launch_bounds__(P_THREADS, 1)  // registers are granted per 4 warps: 352 threads count as 384 -> 168 per thread
gemm_tf32_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmP, const Args a,
                            const int tiles_n, const int tiles_m, const int splits) {
  using T = PTile<BLOCK_N, STAGES, PAIR_BYTES>;
  constexpr int P_STAGES = STAGES;
  constexpr int EPI_PAIR_BYTES = PAIR_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* epi = smem + T::RING_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi + 4 * EPI_PAIR_BYTES);
  uint64_t* empty_bar = full_bar + P_STAGES;
  uint64_t* tfull_bar = empty_bar + P_STAGES;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;         // [2] accumulator drained by the 8 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total = tiles_n * tiles_m * splits;
  const int probe = g_probe;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmB);
    if (a.tma_store) prefetch_tensormap(&tmC);
    for (int s = 0; s < P_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], a.colsum_a ? 2 : 1);  // MMA commit (+ the reducer warp when it reads the A tiles)
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], P_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, T::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();     // barriers, tensor-map prefetch and the TMEM allocation above overlap the tail of the previous kernel
  pdl_trigger();

  auto decode = [&](int t, int& m0, int& n0, int& kb_begin, int& nkb) {
    const int nt = t % tiles_n;
    const int r = t / tiles_n;
    const int mt = r % tiles_m;
    const int z = r / tiles_m;  // split index (== kb_begin / kb_per_split)
    m0 = mt * BLOCK_M;
    n0 = nt * BLOCK_N;
    kb_begin = z * a.kb_per_split;
    nkb = min(a.kb_total, kb_begin + a.kb_per_split) - kb_begin;
  };

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int m0, n0, kb_begin, nkb;
        decode(t, m0, n0, kb_begin, nkb);
        const bool skip_b = (probe & 4) && t != static_cast<int>(blockIdx.x);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % P_STAGES;
          const uint32_t ph = (it / P_STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], skip_b ? A_BYTES : T::STAGE_BYTES);
          uint8_t* sa = smem + s * T::STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const int k0 = (kb_begin + i) * BLOCK_K;
          if (!a.a_mn) {
            tma_load_2d(sa, &tmA, &full_bar[s], k0, m0);
          } else {
#pragma unroll
            for (int g = 0; g < BLOCK_M / 32; ++g) tma_load_2d(sa + g * GROUP_BYTES, &tmA, &full_bar[s], m0 + 32 * g, k0);
          }
          if (skip_b) continue;
          if (!a.b_mn) {
            tma_load_2d(sb, &tmB, &full_bar[s], k0, n0);
          } else {
#pragma unroll
            for (int g = 0; g < BLOCK_N / 32; ++g) tma_load_2d(sb + g * GROUP_BYTES, &tmB, &full_bar[s], n0 + 32 * g, k0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(a.a_mn) << 15) |
                             (static_cast<uint32_t>(a.b_mn) << 16) | (static_cast<uint32_t>(BLOCK_N >> 3) << 17) |
                             (static_cast<uint32_t>(BLOCK_M >> 4) << 24);
      int it = 0, j = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++j) {
        int m0, n0, kb_begin, nkb;
        decode(t, m0, n0, kb_begin, nkb);
        const int ab = j & 1;
        const uint32_t aph = (j >> 1) & 1;
        mbar_wait(&tempty_bar[ab], aph ^ 1);  // epilogue has drained this accumulator (passes at once for its first use)
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(ab * BLOCK_N);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % P_STAGES;
          const uint32_t ph = (it / P_STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + s * T::STAGE_BYTES);
          const uint32_t b_base = a_base + A_BYTES;
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
            const uint64_t adesc = a.a_mn ? make_smem_desc(a_base + kk * 1024, GROUP_BYTES, 512, 1)
                                          : make_smem_desc(a_base + kk * UMMA_K * 4, 16, 1024, 2);
            const uint64_t bdesc = a.b_mn ? make_smem_desc(b_base + kk * 1024, GROUP_BYTES, 512, 1)
                                          : make_smem_desc(b_base + kk * UMMA_K * 4, 16, 1024, 2);
            if (!(probe & 8)) umma_tf32(tmem_d, adesc, bdesc, idesc, (i | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[ab]);
      }
    }
    __syncwarp();
  } else if (warp == P_REDUCER_WARP) {
    // ---------------- reducer warp: column sums of the A operand (= bias gradient when A is an output gradient) -----------
    // Every A tile [128 rows x 32 k] passes through shared memory once per N-tile; the N-tile-0 pass is summed over its rows
    // here (lane = (row group of 4, 16 B chunk), rows strided by 4, 128B-swizzle undone) and added to colsum_a with 32 fp32
    // atomics per k-block.  The warp holds the ring slot until it has read it (second arrival on the empty barrier).
    if (a.colsum_a) {
      const int chunk = lane & 7, grp = lane >> 3;
      int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int m0, n0, kb_begin, nkb;
        decode(t, m0, n0, kb_begin, nkb);
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % P_STAGES;
          const uint32_t ph = (it / P_STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          if (n0 == 0) {
            // all 32 row loads are issued back to back and the ring slot is handed back as soon as they are done (the arrive
            // is a release: it orders the loads before it); the additions run after the slot is already free again.  Holding
            // the slot through the arithmetic cost 16 % of the GEMM's bandwidth (3-stage ring, latency-bound recycle time).
            const uint8_t* sa = smem + s * T::STAGE_BYTES;
            float4 v[BLOCK_M / 4];
#pragma unroll
            for (int q = 0; q < BLOCK_M / 4; ++q) {
              const int r = grp + 4 * q;
              v[q] = *reinterpret_cast<const float4*>(sa + r * 128 + ((chunk ^ (r & 7)) << 4));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < BLOCK_M / 4; ++q) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
#pragma unroll
            for (int off = 8; off < 32; off <<= 1) {
              acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
              acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
            }
            if (grp == 0) {
              const int k = (kb_begin + i) * BLOCK_K + 4 * chunk;
              if (k < a.K) atomicAdd(a.colsum_a + k, acc.x);
              if (k + 1 < a.K) atomicAdd(a.colsum_a + k + 1, acc.y);
              if (k + 2 < a.K) atomicAdd(a.colsum_a + k + 2, acc.z);
              if (k + 3 < a.K) atomicAdd(a.colsum_a + k + 3, acc.w);
            }
          } else {
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
          }
        }
      }
    }
  } else {
    // ---------------- epilogue warps 2..9: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 -------------------
    // The K=128 layers are epilogue-bound (bias / GELU / derivative / two outputs per element), so 8 warps share one tile:
    // both halves drain the accumulator concurrently, each with its own staging buffers, named barrier and bulk-store issuer.
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = quarter * 32 + lane;
    const bool elected = threadIdx.x == 64 + half * 128;  // first thread of each half issues that half's bulk stores
    uint8_t* epi_half = epi + half * 2 * EPI_PAIR_BYTES;
    constexpr int HALF_N = BLOCK_N / 2;
    int j = 0, cc = 0;  // tile counter, running chunk counter of this half (staging pair = cc & 1)
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++j) {
      int m0, n0, kb_begin, nkb;
      decode(t, m0, n0, kb_begin, nkb);
      const int ab = j & 1;
      const uint32_t aph = (j >> 1) & 1;
      RowEpilogue re;
      re.init(a, m0 + row_in_tile, kb_begin / a.kb_per_split);
      const int nh = n0 + half * HALF_N;
      int n_chunks = 0;
#pragma unroll
      for (int c = 0; c < HALF_N; c += 32) n_chunks += (nh + c < a.N) ? 1 : 0;
      float4 side[8];  // side operand of the NEXT chunk to process, loaded ahead of its use
      bool have_side = n_chunks > 0 && re.side_prefetch(a, nh, min(32, a.N - nh), side);
      mbar_wait(&tfull_bar[ab], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(ab * BLOCK_N + half * HALF_N);
      if (n_chunks == 0) {  // this half lies entirely beyond N: nothing to read, release the accumulator right away
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[ab]);
      }
#pragma unroll 1
      for (int ci = 0; ci < n_chunks; ++ci) {
        const int c = ci * 32;
        const int nb = nh + c;
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + static_cast<uint32_t>(c), r);
        tmem_wait_ld();
        if (ci == n_chunks - 1) {  // accumulator columns of this warp fully read: hand them back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[ab]);
        }
        if (probe & 2) continue;
        float v[32], pre[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(r[q]);
        const int ncols = min(32, a.N - nb);
        re.apply(a, v, pre, nb, ncols, have_side, side);
        have_side = ci + 1 < n_chunks && re.side_prefetch(a, nb + 32, min(32, a.N - (nb + 32)), side);
        if (a.tma_store) {
          uint8_t* obuf = epi_half + (cc & 1) * EPI_PAIR_BYTES;
          uint8_t* pbuf = obuf + 16384;
          if (cc >= 2) {  // staging pair re-used: its previous bulk store must have finished reading shared memory
            if (elected) bulk_wait_group_read<1>();
            epi_bar_sync(half);
          }
          RowEpilogue::stage_row(obuf, row_in_tile, v);
          if (a.preact) RowEpilogue::stage_row(pbuf, row_in_tile, pre);
          fence_proxy_async_smem();
          epi_bar_sync(half);
          if (elected && !(probe & 1)) {
            if (a.accumulate) tma_reduce_add_2d(&tmC, obuf, nb, m0);
            else if (a.group_in > 0) tma_store_3d(&tmC, obuf, nb, 0, m0 / a.group_in);
            else tma_store_2d(&tmC, obuf, nb, m0);
            if (a.preact) tma_store_2d(&tmP, pbuf, nb, m0);
            bulk_commit_group();
          }
          ++cc;
        } else {
          re.store_direct(a, v, pre, nb, ncols);
        }
      }
    }
    if (a.tma_store && elected) bulk_wait_group_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, T::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

static bool tf32_round_in_tma() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RF_TMA_TF32_ROUND");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static CUtensorMapL2promotion l2_promotion() {  // RF_TMA_L2_PROMO=0..3: none / 64 B / 128 B / 256 B (default)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RF_TMA_L2_PROMO");
    v = (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 3;
  }
  return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
       : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
}

// fp32 tensor map of rank 2 or 3: dim0 = `inner` contiguous elements (box 32 = 128 B), dim1 = rows of pitch ld (box box_rows),
// optional dim2 = groups of pitch ld2 (box box_groups).  `round_tf32`: TFLOAT32 type (operands are rounded RN on load).
static int make_map(CUtensorMap* map, const float* base, long long inner, long long outer, long long ld, int box_rows, bool mn_major,
                    bool round_tf32 = true, long long groups = 0, long long ld2 = 0, int box_groups = 0, int f16 = 0) {
  auto fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return RF_ERR_CUDA;
  }
  const int rank = groups > 0 ? 3 : 2;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer), static_cast<cuuint64_t>(groups)};
  const cuuint64_t esize = f16 ? 2 : 4;
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * esize, static_cast<cuuint64_t>(ld2) * esize};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(128 / esize), static_cast<cuuint32_t>(box_rows), static_cast<cuuint32_t>(box_groups)};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dtype = f16 == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : f16 == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                        : ((round_tf32 && tf32_round_in_tma()) ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
  CUresult r = fn(map, dtype, rank,
                  const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, l2_promotion(),
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): base=%p inner=%lld outer=%lld ld=%lld box_rows=%d groups=%lld", static_cast<int>(r),
              static_cast<const void*>(base), inner, outer, ld, box_rows, groups);
    return RF_ERR_CUDA;
  }
  return RF_OK;
}

template <int BLOCK_N>
static int launch(const RfGemmParams* p, const Args& args, int splits, cudaStream_t stream) {
  using T = Tile<BLOCK_N>;
  CUtensorMap tmA, tmB;
  int rc;
  const int f16 = args.f16;
  if (!p->a_mn_major) rc = make_map(&tmA, p->A, p->K, p->M, p->lda, BLOCK_M, false, true, 0, 0, 0, f16);
  else rc = make_map(&tmA, p->A, p->M, p->K, p->lda, 32, true);
  if (rc != RF_OK) return rc;
  if (!p->b_mn_major) rc = make_map(&tmB, p->B, p->K, p->N, p->ldb, BLOCK_N, false, true, 0, 0, 0, f16);
  else rc = make_map(&tmB, p->B, p->N, p->K, p->ldb, 32, true);
  if (rc != RF_OK) return rc;
  CUtensorMap tmC, tmP;
  memset(&tmC, 0, sizeof(tmC));
  memset(&tmP, 0, sizeof(tmP));
  if (args.tma_store) {
    if (p->out_group_in > 0) {
      const long long n_groups = ceil_div(p->M, p->out_group_in);
      rc = make_map(&tmC, p->C, p->N, p->out_group_out, p->ldc, p->out_group_in, false, false, n_groups,
                    static_cast<long long>(p->out_group_out) * p->ldc, BLOCK_M / p->out_group_in);
    } else {
      rc = make_map(&tmC, p->C, p->N, p->M, p->ldc, BLOCK_M, false, false);
    }
    if (rc != RF_OK) return rc;
    if (p->preact) {
      rc = make_map(&tmP, p->preact, p->N, p->M, p->ld_pre, BLOCK_M, false, false);
      if (rc != RF_OK) return rc;
    }
  }
  static int persistent = -1;
  if (persistent < 0) {
    const char* e = getenv("RF_GEMM_PERSISTENT");
    persistent = (e && e[0] == '0') ? 0 : 1;
  }
  const int tiles_n = ceil_div(p->N, BLOCK_N), tiles_m = ceil_div(p->M, BLOCK_M);
  // Short reductions (<= 12 k-blocks per tile: the K=128/256/384 layers) are epilogue/latency-bound -> persistent kernel whose
  // loads run ahead across tiles.  Long reductions are L2-bandwidth-bound -> two co-resident tile-wise CTAs per SM keep more
  // bytes in flight (2 x 3 stages) and measured 25-35 % faster there (profiles/r1_microbench_gemm_*).
  // (the persistent kernel also for launches of at most one tile per SM -- the M = 2 560 ... 10 240 layers of the gaze / video
  //  encoders, ~200 per step: the lighter tile-wise kernel there measured +0.38 ms per step, its 4 epilogue warps take all columns)
  if (persistent && args.kb_per_split <= 12 && !f16) {
    const long long total = static_cast<long long>(tiles_n) * tiles_m * splits;
    RF_CHECK_ARG(total <= 2147483647LL, "rf_gemm_tf32: too many tiles");
    const int grid = static_cast<int>(total < num_sms() ? total : num_sms());
    // RF_GEMM_DEEP_RING=1: opt-in.  Measured A/B on the training step (profiles/r2_gemm_deep_ring_ab.txt): 19.90 vs 19.98 ms per
    // step, roofline fraction of the HBM-bound launches 0.527 vs 0.529 -- the bytes in flight of the operand ring are NOT what
    // bounds these launches, so the round-1 configuration stays the default.
    static const bool deep_ring = [] { const char* e = getenv("RF_GEMM_DEEP_RING"); return e && e[0] == '1'; }();
    if (p->preact || !deep_ring || !args.tma_store) {
      using PT = PTile<BLOCK_N, P_STAGES_PREACT, 32768>;
      auto kernel = gemm_tf32_persistent_kernel<BLOCK_N, P_STAGES_PREACT, 32768>;
      RF_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), PT::SMEM_BYTES));
      RF_CUDA_OK(launch_pdl(kernel, dim3(grid), dim3(P_THREADS), PT::SMEM_BYTES, stream, tmA, tmB, tmC, tmP, args, tiles_n, tiles_m, splits));
    } else {
      using PT = PTile<BLOCK_N, P_STAGES_PLAIN, 16384>;
      auto kernel = gemm_tf32_persistent_kernel<BLOCK_N, P_STAGES_PLAIN, 16384>;
      RF_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), PT::SMEM_BYTES));
      RF_CUDA_OK(launch_pdl(kernel, dim3(grid), dim3(P_THREADS), PT::SMEM_BYTES, stream, tmA, tmB, tmC, tmP, args, tiles_n, tiles_m, splits));
    }
    return RF_OK;
  }
  if (args.colsum_a) {  // long reductions: the tile-wise kernels have no reducer warp, the column sums take their own pass
    rc = rf_colsum_accumulate(p->A, p->lda, p->M, p->K, p->colsum_a, stream);
    if (rc != RF_OK) return rc;
  }
  dim3 grid(tiles_n, tiles_m, splits);
  RF_CHECK_ARG(grid.y <= 65535, "rf_gemm_tf32: M=%d exceeds 65535 row tiles", p->M);
  static int deep = -1;
  if (deep < 0) {
    const char* e = getenv("RF_GEMM_DEEP");
    deep = (e && e[0] == '0') ? 0 : 1;
  }
  // A grid that does not even fill one wave leaves the HBM/L2 pipes idle while every CTA walks its K loop at the pace of the
  // TMA round trip (measured 0.38 us per k-block with 3 stages): give those CTAs the whole SM's shared memory instead.
  if (deep && static_cast<long long>(tiles_n) * tiles_m * splits <= num_sms() && args.kb_per_split > STAGES_SHALLOW) {
    using TD = Tile<BLOCK_N, STAGES_DEEP>;
    RF_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(gemm_tf32_kernel<BLOCK_N, STAGES_DEEP>), TD::SMEM_BYTES));
    RF_CUDA_OK(launch_pdl(gemm_tf32_kernel<BLOCK_N, STAGES_DEEP>, grid, dim3(NUM_THREADS), TD::SMEM_BYTES, stream, tmA, tmB, tmC, tmP, args));
    return RF_OK;
  }
  RF_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(gemm_tf32_kernel<BLOCK_N, STAGES_SHALLOW>), T::SMEM_BYTES));
  RF_CUDA_OK(launch_pdl(gemm_tf32_kernel<BLOCK_N, STAGES_SHALLOW>, grid, dim3(NUM_THREADS), T::SMEM_BYTES, stream, tmA, tmB, tmC, tmP, args));
  return RF_OK;
}

}  // namespace gemm
}  // namespace rf

extern "C" int rf_debug_gemm_stamps(unsigned long long* device_buffer) {
  using namespace rf;
  RF_CUDA_OK(cudaMemcpyToSymbol(gemm::g_stamps, &device_buffer, sizeof(device_buffer)));
  return RF_OK;
}

extern "C" int rf_debug_gemm_probe(int mode) {
  using namespace rf;
  RF_CUDA_OK(cudaMemcpyToSymbol(gemm::g_probe, &mode, sizeof(mode)));
  return RF_OK;
}

static int split_fill() {  // CTAs (in units of the SM count) that the automatic split of a linear epilogue aims for
  static const int v = [] { const char* e = getenv("RF_GEMM_SPLIT_FILL"); return (e && e[0] >= '1' && e[0] <= '9') ? e[0] - '0' : 1; }();
  return v;
}

extern "C" int rf_gemm_tf32(const RfGemmParams* p, void* stream) {
  using namespace rf;
  RF_CHECK_ARG(p != nullptr, "rf_gemm_tf32: null params");
  RF_CHECK_ARG(p->M > 0 && p->N > 0 && p->K > 0, "rf_gemm_tf32: empty problem M=%d N=%d K=%d", p->M, p->N, p->K);
  RF_CHECK_ARG(p->A && p->B && p->C, "rf_gemm_tf32: null operand");
  RF_CHECK_ARG(p->ab_dtype == RF_F32 || p->ab_dtype == RF_F16 || p->ab_dtype == RF_BF16, "rf_gemm_tf32: ab_dtype must be RF_F32, RF_F16 or RF_BF16");
  const bool f16 = p->ab_dtype != RF_F32;  // 16-bit operands
  RF_CHECK_ARG(p->c_dtype == RF_F32 || p->c_dtype == RF_BF16, "rf_gemm_tf32: c_dtype must be RF_F32 or RF_BF16");
  RF_CHECK_ARG(p->c_dtype == RF_F32 || (f16 && !p->accumulate && !p->preact && p->split_k <= 1),
               "rf_gemm_tf32: a bf16 C needs 16-bit operands, no accumulate / preact / split-K");
  const int pitch_mult = f16 ? 8 : 4;
  RF_CHECK_ARG((p->lda % pitch_mult) == 0 && (p->ldb % pitch_mult) == 0, "rf_gemm_tf32: lda=%lld ldb=%lld must be multiples of %d (TMA 16 B pitch)",
               p->lda, p->ldb, pitch_mult);
  RF_CHECK_ARG(!f16 || (!p->a_mn_major && !p->b_mn_major), "rf_gemm_tf32: 16-bit operands must be K-major");
  RF_CHECK_ARG((reinterpret_cast<uintptr_t>(p->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->B) & 15) == 0,
               "rf_gemm_tf32: A/B must be 16-byte aligned");
  RF_CHECK_ARG(p->lda >= (p->a_mn_major ? p->M : p->K) && p->ldb >= (p->b_mn_major ? p->N : p->K) && p->ldc >= p->N,
               "rf_gemm_tf32: leading dimension smaller than the row length");
  RF_CHECK_ARG(!p->rowadd || p->rowadd_period > 0, "rf_gemm_tf32: rowadd needs rowadd_period > 0");
  RF_CHECK_ARG(!p->dact || p->dact_aux, "rf_gemm_tf32: dact needs dact_aux");
  RF_CHECK_ARG(p->act != RF_ACT_GELU_SAVE_GRAD || p->preact, "rf_gemm_tf32: RF_ACT_GELU_SAVE_GRAD needs the preact buffer");
  RF_CHECK_ARG(!p->colsum_a || (!f16 && !p->a_mn_major && p->split_k <= 1 && !p->accumulate),
               "rf_gemm_tf32: colsum_a needs a K-major fp32 A and no split-K");
  const int kb_total = ceil_div(p->K, f16 ? 2 * gemm::BLOCK_K : gemm::BLOCK_K);
  const bool plain = !p->bias && !p->rowadd && !p->residual && p->act == RF_ACT_NONE && !p->preact && !p->dact && !p->round_f16;
  int splits = p->split_k;
  const bool n64 = p->N <= 64;
  const int tiles = ceil_div(p->M, gemm::BLOCK_M) * ceil_div(p->N, n64 ? 64 : 128);
  if (splits <= 0) {
    splits = 1;
    if (p->accumulate && plain) {
      static const int wfill = [] { const char* e = getenv("RF_GEMM_WGRAD_FILL"); return (e && e[0] >= '1' && e[0] <= '9') ? e[0] - '0' : 2; }();
      const int want = max(1, (wfill * num_sms()) / tiles);   // fill, but do not overflow, one wave of 2 CTAs per SM
      splits = max(1, min(want, kb_total / 4));              // keep >= 4 k-blocks per split
    }
  }
  RF_CHECK_ARG(splits == 1 || (p->accumulate && plain), "rf_gemm_tf32: split_k>1 needs accumulate=1 and a plain epilogue");
  // Few output tiles and a long reduction (the Informer layers at 64 clips per GPU: 14-80 tiles, 26-104 k-blocks): every CTA
  // walks its K loop at the pace of one TMA round trip per stage while most SMs idle.  If the epilogue is linear in the
  // accumulator (bias / positional rows / residual only), split the reduction over more CTAs: the output is zero-filled, split
  // 0 adds the linear terms, and all splits meet in the TMA reduce-add.
  bool split_linear = false;
  if (p->split_k == 0 && !p->accumulate && !f16 && p->act == RF_ACT_NONE && !p->preact && !p->dact && !p->round_f16 &&
      p->out_group_in == 0 && 2 * tiles <= split_fill() * num_sms() && kb_total >= 8) {
    const char* cb = reinterpret_cast<const char*>(p->C);
    const char* ce = cb + (static_cast<long long>(p->M - 1) * p->ldc + p->N) * 4;
    const char* rb = reinterpret_cast<const char*>(p->residual);
    const bool res_aliases = rb && rb < ce && rb + (static_cast<long long>(p->M - 1) * p->ld_res + p->N) * 4 > cb;
    static const bool enabled = [] { const char* e = getenv("RF_GEMM_SPLIT_LINEAR"); return !(e && e[0] == '0'); }();
    const int want = min(split_fill() * num_sms() / tiles, kb_total / 4);
    if (enabled && !res_aliases && want >= 2) {
      splits = want;
      split_linear = true;
    }
  }
  int kb_per_split = ceil_div(kb_total, splits);
  splits = ceil_div(kb_total, kb_per_split);
  if (split_linear && splits < 2) split_linear = false;
  if (split_linear)
    RF_CUDA_OK(cudaMemset2DAsync(p->C, static_cast<size_t>(p->ldc) * 4, 0, static_cast<size_t>(p->N) * 4, p->M, static_cast<cudaStream_t>(stream)));
  gemm::Args a;
  a.C = p->C; a.ldc = p->ldc; a.M = p->M; a.N = p->N; a.K = p->K;
  a.a_mn = p->a_mn_major ? 1 : 0; a.b_mn = p->b_mn_major ? 1 : 0;
  a.bias = p->bias; a.rowadd = p->rowadd; a.rowadd_period = p->rowadd_period; a.ld_rowadd = p->ld_rowadd;
  a.residual = p->residual; a.ld_res = p->ld_res; a.act = p->act; a.preact = p->preact; a.ld_pre = p->ld_pre;
  a.dact_aux = p->dact_aux; a.ld_aux = p->ld_aux; a.dact = p->dact; a.accumulate = (p->accumulate || split_linear) ? 1 : 0;
  a.split_epilogue = split_linear ? 1 : 0;
  a.kb_total = kb_total; a.kb_per_split = kb_per_split; a.f16 = p->ab_dtype == RF_F16 ? 1 : (p->ab_dtype == RF_BF16 ? 2 : 0);
  a.c_bf16 = p->c_dtype == RF_BF16 ? 1 : 0; a.colsum_a = p->colsum_a;
  a.group_in = p->out_group_in; a.group_out = p->out_group_out; a.row_offset = p->out_row_offset;
  a.round_f16 = p->round_f16;
  static int tma_store_enabled = -1;
  if (tma_store_enabled < 0) {
    const char* e = getenv("RF_GEMM_TMA_STORE");
    tma_store_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  const bool c_ok = (p->ldc % 4) == 0 && (reinterpret_cast<uintptr_t>(p->C) & 15) == 0;
  const bool group_ok = p->out_group_in == 0 ||
                        (gemm::BLOCK_M % p->out_group_in == 0 && p->out_row_offset == 0 && p->out_group_out >= p->out_group_in);
  const bool pre_ok = !p->preact || ((p->ld_pre % 4) == 0 && (reinterpret_cast<uintptr_t>(p->preact) & 15) == 0 && p->out_group_in == 0);
  const bool acc_ok = !p->accumulate || (p->out_group_in == 0 && !p->preact);
  a.tma_store = (tma_store_enabled && c_ok && group_ok && pre_ok && acc_ok && !a.c_bf16) ? 1 : 0;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return n64 ? gemm::launch<64>(p, a, splits, s) : gemm::launch<128>(p, a, splits, s);
}
