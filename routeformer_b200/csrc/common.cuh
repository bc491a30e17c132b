// Shared device/host helpers for the routeformer_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/routeformer_b200.h"

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL)
#error "routeformer_b200 kernels must be compiled with -gencode arch=compute_100a,code=sm_100a"
#endif

namespace rf {

// ---- error plumbing: C-ABI entries never throw; they return RF_ERR_* and stash a message ------
void set_error(const char* fmt, ...);

#define RF_CHECK_ARG(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      ::rf::set_error(__VA_ARGS__);             \
      return RF_ERR_INVALID_ARGUMENT;           \
    }                                           \
  } while (0)

#define RF_CUDA_OK(expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::rf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return RF_ERR_CUDA;                                                             \
    }                                                                                 \
  } while (0)

#define RF_LAUNCH_OK()                                                                \
  do {                                                                                \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) {                                                          \
      ::rf::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return RF_ERR_CUDA;                                                             \
    }                                                                                 \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Opt-in to > 48 KiB of dynamic shared memory, once per (device, kernel): cudaFuncSetAttribute is a per-device setting, and
// doing it per launch would break CUDA-graph capture.  Thread-safe (autograd's backward threads launch concurrently with the
// main thread).
inline cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes) {
  static std::mutex mu;
  static const void* done[128][2];
  static int n_done = 0;
  if (bytes <= 48 * 1024) return cudaSuccess;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const void* dkey = reinterpret_cast<const void*>(static_cast<uintptr_t>(dev) + 1);
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < n_done; ++i)
    if (done[i][0] == kernel && done[i][1] == dkey) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) return e;
  if (n_done < 128) { done[n_done][0] = kernel; done[n_done][1] = dkey; ++n_done; }
  return cudaSuccess;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch ------------------------------------------------------------
// A training step is ~800 launches, most of them 5-30 us long: the gap between dependent kernels (the next grid's CTAs are only
// scheduled after the previous grid has drained) is a measurable share of the step.  Kernels launched through launch_pdl may be
// scheduled while their predecessor in the stream is still running; they must call pdl_wait() before touching global memory
// (it returns once the predecessor has completed and flushed) and pdl_trigger() to let their own successor be scheduled.
// RF_PDL=0 disables the attribute (then pdl_wait / pdl_trigger are no-ops by definition).
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("RF_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, i.e. at fp32 rounding level for GELU): branch-free, one reciprocal and
// one exponential, ~3x fewer instructions than erff() in the GEMM epilogues where it runs once per output element.
// e = exp(-x^2) is passed in so that the GELU derivative can share it with its Gaussian term.
__device__ __forceinline__ float erf_as(float x, float e) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));  // MUFU.RCP, 1 ulp
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  return copysignf(fmaf(-p * t, e, 1.0f), x);
}
__device__ __forceinline__ float gelu_erf(float x) {  // exact-erf GELU (F.gelu default): 0.5 x (1 + erf(x / sqrt 2))
  const float z = x * 0.70710678118654752440f;
  return 0.5f * x * (1.0f + erf_as(z, __expf(-z * z)));
}
// gelu(x) and gelu'(x) together: erf and exp are shared
__device__ __forceinline__ float gelu_erf_with_grad(float x, float& grad) {
  const float z = x * 0.70710678118654752440f;
  const float e = __expf(-z * z);
  const float cdf = 0.5f * (1.0f + erf_as(z, e));
  grad = cdf + x * 0.39894228040143267794f * e;
  return x * cdf;
}
__device__ __forceinline__ float gelu_erf_grad(float x) {  // Phi(x) + x phi(x)
  const float z = x * 0.70710678118654752440f;
  const float e = __expf(-z * z);  // exp(-x^2/2)
  return 0.5f * (1.0f + erf_as(z, e)) + x * 0.39894228040143267794f * e;
}

// ---- counter-based RNG for dropout: Philox4x32-10 keyed by `seed`, counter = (index of a 4-element group, call offset).
// Stateless, so the backward pass regenerates the forward mask from (seed, offset) instead of storing it.
__device__ __forceinline__ uint4 philox4x32_10(unsigned long long seed, unsigned long long group, unsigned long long offset) {
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  uint32_t c0 = static_cast<uint32_t>(group), c1 = static_cast<uint32_t>(group >> 32);
  uint32_t c2 = static_cast<uint32_t>(offset), c3 = static_cast<uint32_t>(offset >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// keep-factor (0 or 1/(1-p)) of logical element `index` of the tensor a dropout call covers
__device__ __forceinline__ float dropout_factor(unsigned long long seed, unsigned long long offset, unsigned long long index,
                                                uint32_t threshold, float scale) {
  const uint4 r = philox4x32_10(seed, index >> 2, offset);
  const uint32_t lane = static_cast<uint32_t>(index & 3);
  const uint32_t v = lane == 0 ? r.x : (lane == 1 ? r.y : (lane == 2 ? r.z : r.w));
  return v >= threshold ? scale : 0.0f;
}
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {  // P(u32 < threshold) = p
  const double t = static_cast<double>(p) * 4294967296.0;
  return t >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(t);
}

}  // namespace rf
