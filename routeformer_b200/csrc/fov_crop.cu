// Field-of-view crop / resample (HBM-bound).  See include/routeformer_b200.h (1).
//
// Every output pixel is a bilinear sample of the source frame at
//   sx = ((fw * gx + 2cx - 1 + 1) * W - 1) / 2,  gx = (2 ox + 1) / S - 1   (grid_sample, align_corners=False)
// with zeros outside the frame, then normalised per channel.  One thread produces 4 consecutive output
// pixels of one output row for all 3 channels (the sample positions and weights are channel independent)
// and writes one 16 B (fp32) / 8 B (bf16) vector per channel, either planar [n,3,S,S] or directly in the
// patch-major layout the patch-embedding GEMM consumes as its A operand (no im2col pass).
// Source reads go through the read-only path; a CTA covers 4 consecutive output rows of one frame so the
// 2 x (4+1) source taps of neighbouring threads hit the same L1 lines and DRAM sees each source line once.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace rf {
namespace crop {

template <typename T> __device__ __forceinline__ float load_px(const T* p);
template <> __device__ __forceinline__ float load_px<__half>(const __half* p) { return __half2float(__ldg(p)); }
template <> __device__ __forceinline__ float load_px<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load_px<unsigned char>(const unsigned char* p) { return __ldg(p) * (1.0f / 255.0f); }

__device__ __forceinline__ void store4(float* dst, const float (&v)[4]) {
  *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(__nv_bfloat16* dst, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(dst) = u;
}

__device__ __forceinline__ void store4(__half* dst, const float (&v)[4]) {
  __half2 a = __floats2half2_rn(v[0], v[1]);
  __half2 b = __floats2half2_rn(v[2], v[3]);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(dst) = u;
}

struct Args {
  const void* frames; const int* frame_ids; int n_frames, H, W;
  const float* centers; const float* windows;
  float mean[3], inv_std[3];
  int S, patch; void* out; long long out_ld;
  // host-side constants that keep integer division out of the kernel (it was 35 % of all executed instructions):
  // q = (x * magic) >> 32 is exact for x < 65536 with magic = 2^32 / d + 1
  int quads, G;                     // S / 4, S / patch
  unsigned quads_magic, patch_magic;
};
__device__ __forceinline__ int fast_div(int x, unsigned magic) { return static_cast<int>(__umulhi(static_cast<unsigned>(x), magic)); }

constexpr int ROWS_PER_CTA = 4;

template <typename TS, typename TD>
__global__ void __launch_bounds__(256) fov_crop_kernel(const Args a) {
  const int quads = a.quads;                         // 4-pixel groups per output row
  const int n = blockIdx.y;
  const int local = threadIdx.x;
  const int lrow = fast_div(local, a.quads_magic);
  const int oy = blockIdx.x * ROWS_PER_CTA + lrow;
  const int ox0 = (local - lrow * quads) << 2;
  if (local >= quads * ROWS_PER_CTA || oy >= a.S) return;

  const float cx = __ldg(a.centers + 2 * n), cy = __ldg(a.centers + 2 * n + 1);
  const float fw = __ldg(a.windows + 2 * n), fh = __ldg(a.windows + 2 * n + 1);
  const long long src_frame = a.frame_ids ? __ldg(a.frame_ids + n) : n;
  const TS* src = reinterpret_cast<const TS*>(a.frames) + src_frame * 3ll * a.H * a.W;
  const long long plane = static_cast<long long>(a.H) * a.W;

  const float inv_s = 1.0f / a.S;
  const float gy = (2 * oy + 1) * inv_s - 1.0f;
  const float sy = ((fh * gy + (2.0f * cy - 1.0f) + 1.0f) * a.H - 1.0f) * 0.5f;
  const float fy0 = floorf(sy);
  const int y0 = static_cast<int>(fy0);
  const float wy1 = sy - fy0, wy0 = 1.0f - wy1;
  const bool y0_ok = y0 >= 0 && y0 < a.H, y1_ok = (y0 + 1) >= 0 && (y0 + 1) < a.H;

  float acc[3][4];
  // sample columns of the 4 pixels (channel independent)
  int x0[4];
  float wx1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float gx = (2 * (ox0 + i) + 1) * inv_s - 1.0f;
    const float sx = ((fw * gx + (2.0f * cx - 1.0f) + 1.0f) * a.W - 1.0f) * 0.5f;
    const float fx0 = floorf(sx);
    x0[i] = static_cast<int>(fx0);
    wx1[i] = sx - fx0;
  }
  const bool rows_in = y0_ok && y1_ok;
  const bool mono = fw > 0.0f;  // then sx grows with i and the 4 pixels' taps lie in [x0[0], x0[3]+1]
  const bool cols_in = mono && x0[0] >= 0 && x0[3] + 1 < a.W;
  const bool cols_out = mono && (x0[3] + 1 < 0 || x0[0] >= a.W);
  if ((!y0_ok && !y1_ok) || cols_out) {
    // the whole 4-pixel group samples the zero padding (pad-to-square scene views: ~3/4 of all outputs): constants only
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[c][i] = (0.0f - a.mean[c]) * a.inv_std[c];
  } else if (rows_in && cols_in) {
    // interior: all 16 taps are inside the frame, no per-tap predicates
    const long long r0 = static_cast<long long>(y0) * a.W;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const TS* row0 = src + c * plane + r0;
      const TS* row1 = row0 + a.W;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float v00 = load_px<TS>(row0 + x0[i]), v01 = load_px<TS>(row0 + x0[i] + 1);
        const float v10 = load_px<TS>(row1 + x0[i]), v11 = load_px<TS>(row1 + x0[i] + 1);
        const float wx0 = 1.0f - wx1[i];
        const float sacc = v00 * (wx0 * wy0) + v01 * (wx1[i] * wy0) + v10 * (wx0 * wy1) + v11 * (wx1[i] * wy1);
        acc[c][i] = (sacc - a.mean[c]) * a.inv_std[c];
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float wx0 = 1.0f - wx1[i];
      const bool x0_ok = x0[i] >= 0 && x0[i] < a.W, x1_ok = (x0[i] + 1) >= 0 && (x0[i] + 1) < a.W;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const TS* pl = src + c * plane;
        float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
        if (y0_ok) {
          const TS* row = pl + static_cast<long long>(y0) * a.W;
          if (x0_ok) v00 = load_px<TS>(row + x0[i]);
          if (x1_ok) v01 = load_px<TS>(row + x0[i] + 1);
        }
        if (y1_ok) {
          const TS* row = pl + static_cast<long long>(y0 + 1) * a.W;
          if (x0_ok) v10 = load_px<TS>(row + x0[i]);
          if (x1_ok) v11 = load_px<TS>(row + x0[i] + 1);
        }
        // same association order as ATen's grid_sampler: sum of (value * area weight)
        const float sacc = v00 * (wx0 * wy0) + v01 * (wx1[i] * wy0) + v10 * (wx0 * wy1) + v11 * (wx1[i] * wy1);
        acc[c][i] = (sacc - a.mean[c]) * a.inv_std[c];
      }
    }
  }
  TD* out = reinterpret_cast<TD*>(a.out);
  if (a.patch > 0) {
    const int py = fast_div(oy, a.patch_magic), iy = oy - py * a.patch;
    const int px = fast_div(ox0, a.patch_magic), ix = ox0 - px * a.patch;
    TD* dst = out + (static_cast<long long>(n) * a.G * a.G + py * a.G + px) * a.out_ld + iy * a.patch + ix;
    const int cstride = a.patch * a.patch;
#pragma unroll
    for (int c = 0; c < 3; ++c) store4(dst + c * cstride, acc[c]);
  } else {
    TD* dst = out + (static_cast<long long>(n) * 3 * a.S + oy) * a.S + ox0;
    const long long cstride = static_cast<long long>(a.S) * a.S;
#pragma unroll
    for (int c = 0; c < 3; ++c) store4(dst + c * cstride, acc[c]);
  }
}

template <typename TS>
static int dispatch_out(const RfFovCropParams* p, const Args& a, cudaStream_t s) {
  const int quads = p->out_size / 4;
  dim3 grid(ceil_div(p->out_size, ROWS_PER_CTA), p->n_frames);
  const int threads = ((quads * ROWS_PER_CTA + 31) / 32) * 32;
  if (p->out_dtype == RF_F32) fov_crop_kernel<TS, float><<<grid, threads, 0, s>>>(a);
  else if (p->out_dtype == RF_F16) fov_crop_kernel<TS, __half><<<grid, threads, 0, s>>>(a);
  else fov_crop_kernel<TS, __nv_bfloat16><<<grid, threads, 0, s>>>(a);
  RF_LAUNCH_OK();
  return RF_OK;
}

}  // namespace crop
}  // namespace rf

extern "C" int rf_fov_crop(const RfFovCropParams* p, void* stream) {
  using namespace rf;
  RF_CHECK_ARG(p && p->frames && p->centers && p->windows && p->out, "rf_fov_crop: null pointer");
  RF_CHECK_ARG(p->n_frames > 0 && p->H > 0 && p->W > 0, "rf_fov_crop: empty input");
  RF_CHECK_ARG(p->out_size > 0 && p->out_size % 4 == 0 && p->out_size <= 256, "rf_fov_crop: out_size=%d must be a multiple of 4, <= 256",
               p->out_size);
  RF_CHECK_ARG(p->patch == 0 || (p->patch % 4 == 0 && p->out_size % p->patch == 0), "rf_fov_crop: patch=%d must divide out_size and be a multiple of 4",
               p->patch);
  RF_CHECK_ARG(p->patch == 0 || (p->out_ld >= 3ll * p->patch * p->patch && p->out_ld % 4 == 0), "rf_fov_crop: out_ld too small or not a multiple of 4");
  RF_CHECK_ARG(p->out_dtype == RF_F32 || p->out_dtype == RF_BF16 || p->out_dtype == RF_F16, "rf_fov_crop: out_dtype must be RF_F32, RF_F16 or RF_BF16");
  RF_CHECK_ARG(p->n_frames <= 65535, "rf_fov_crop: at most 65535 frames per call");
  crop::Args a;
  a.frames = p->frames; a.frame_ids = p->frame_ids; a.n_frames = p->n_frames; a.H = p->H; a.W = p->W;
  a.centers = p->centers; a.windows = p->windows;
  for (int c = 0; c < 3; ++c) { a.mean[c] = p->mean[c]; a.inv_std[c] = p->inv_std[c]; }
  a.S = p->out_size; a.patch = p->patch; a.out = p->out; a.out_ld = p->out_ld;
  a.quads = p->out_size / 4;
  a.G = p->patch > 0 ? p->out_size / p->patch : 0;
  a.quads_magic = static_cast<unsigned>((1ull << 32) / static_cast<unsigned>(a.quads) + 1);
  a.patch_magic = p->patch > 0 ? static_cast<unsigned>((1ull << 32) / static_cast<unsigned>(p->patch) + 1) : 0u;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (p->src_dtype) {
    case RF_F16: return crop::dispatch_out<__half>(p, a, s);
    case RF_F32: return crop::dispatch_out<float>(p, a, s);
    case RF_U8: return crop::dispatch_out<unsigned char>(p, a, s);
    default: set_error("rf_fov_crop: unsupported src_dtype %d", p->src_dtype); return RF_ERR_INVALID_ARGUMENT;
  }
}
