// Field-of-view crop / resample (HBM-bound).  See include/routeformer_b200.h (1).
//
// Every output pixel is a bilinear sample of the source frame at
//   sx = ((fw * gx + 2cx - 1 + 1) * W - 1) / 2,  gx = (2 ox + 1) / S - 1   (grid_sample, align_corners=False)
// with zeros outside the frame, then normalised per channel, written either planar [n,3,S,S] or directly in the
// patch-major layout the patch-embedding GEMM consumes as its A operand (no im2col pass).
//
// fov_crop_tiled_kernel (default): the bilinear resample is SEPARABLE, and both the column taps (x0, wx) and the row taps
// (y0, wy) are shared by the three channels.  A CTA owns ROWS_PER_TILE output rows of one frame and runs two passes:
//   V: for every output row and channel, the vertical blend  t[j] = src[y0][j] * wy0 + src[y0+1][j] * wy1  over the source
//      columns the frame's window covers, read from HBM with coalesced 2-element vector loads (a warp reads 128..256
//      contiguous bytes per instruction) and staged in shared memory;
//   H: every thread produces 2 adjacent output pixels per (row, channel) from 2 x 2 shared-memory taps, normalises, and
//      stores one packed 4 B (fp16 / bf16) or 8 B (fp32) word, a warp writing 128 / 256 contiguous bytes.
// ~10 instructions per output element instead of ~35 for the direct 4-tap gather (which was issue-bound at 0.33 of the HBM
// roofline).  Tiles whose rows all fall into the zero padding (3/4 of a pad-to-square scene view) only store constants.
// fov_crop_kernel (the round-1 direct gather) remains for mirrored windows (fw <= 0), which the tiled kernel does not handle.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace rf {
namespace crop {

template <typename T> __device__ __forceinline__ float load_px(const T* p);
template <> __device__ __forceinline__ float load_px<__half>(const __half* p) { return __half2float(__ldg(p)); }
template <> __device__ __forceinline__ float load_px<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float load_px<unsigned char>(const unsigned char* p) { return __ldg(p) * (1.0f / 255.0f); }
// uint8 pixel converted the way the reference's loader does on the host, `astype(float16) / 255.0` (io/dataset.py:1505-1522):
// fp16(v * (1/255)) equals numpy's fp16 division for all 256 values (checked exhaustively), so uint8 frames staged on the
// device (half the H2D bytes) give bit-identical taps
struct u8h { unsigned char v; };
template <> __device__ __forceinline__ float load_px<u8h>(const u8h* p) {
  return __half2float(__float2half_rn(__ldg(&p->v) * (1.0f / 255.0f)));
}

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

// ---- tiled, separable kernel -------------------------------------------------------------------
constexpr int TILE_THREADS = 256;

// two adjacent source elements (element offset even, pointer 2-element aligned) -> float2
template <typename T> __device__ __forceinline__ float2 load_px2(const T* p);
template <> __device__ __forceinline__ float2 load_px2<__half>(const __half* p) {
  const unsigned u = __ldg(reinterpret_cast<const unsigned*>(p));
  return __half22float2(*reinterpret_cast<const __half2*>(&u));
}
template <> __device__ __forceinline__ float2 load_px2<float>(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
template <> __device__ __forceinline__ float2 load_px2<unsigned char>(const unsigned char* p) {
  const unsigned short u = __ldg(reinterpret_cast<const unsigned short*>(p));
  return make_float2((u & 0xFF) * (1.0f / 255.0f), (u >> 8) * (1.0f / 255.0f));
}
template <> __device__ __forceinline__ float2 load_px2<u8h>(const u8h* p) {
  const unsigned short u = __ldg(reinterpret_cast<const unsigned short*>(p));
  return __half22float2(__floats2half2_rn((u & 0xFF) * (1.0f / 255.0f), (u >> 8) * (1.0f / 255.0f)));
}

__device__ __forceinline__ void store2(float* dst, float a, float b) { *reinterpret_cast<float2*>(dst) = make_float2(a, b); }
__device__ __forceinline__ void store2(__half* dst, float a, float b) { *reinterpret_cast<__half2*>(dst) = __floats2half2_rn(a, b); }
__device__ __forceinline__ void store2(__nv_bfloat16* dst, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(a, b);
}

struct TiledArgs {
  Args a;
  int pitch;   // floats per (channel, row) line of the staging buffer: >= W + 2, even
  int vec2;    // 1: W even and the frame base 2-element aligned -> 2-element vector loads in pass V
};

// one bilinear sample with zero padding, taps in ATen's order (mirrored windows only: rare, slow, correct)
template <typename TS>
__device__ __forceinline__ float direct_sample(const TS* pl, float sx, float sy, int H, int W) {
  const float fx0 = floorf(sx), fy0 = floorf(sy);
  const int x0 = static_cast<int>(fx0), y0 = static_cast<int>(fy0);
  const float wx1 = sx - fx0, wx0 = 1.0f - wx1, wy1 = sy - fy0, wy0 = 1.0f - wy1;
  const bool xa = x0 >= 0 && x0 < W, xb = x0 + 1 >= 0 && x0 + 1 < W, ya = y0 >= 0 && y0 < H, yb = y0 + 1 >= 0 && y0 + 1 < H;
  const float v00 = (xa && ya) ? load_px<TS>(pl + static_cast<long long>(y0) * W + x0) : 0.0f;
  const float v01 = (xb && ya) ? load_px<TS>(pl + static_cast<long long>(y0) * W + x0 + 1) : 0.0f;
  const float v10 = (xa && yb) ? load_px<TS>(pl + static_cast<long long>(y0 + 1) * W + x0) : 0.0f;
  const float v11 = (xb && yb) ? load_px<TS>(pl + static_cast<long long>(y0 + 1) * W + x0 + 1) : 0.0f;
  return v00 * (wx0 * wy0) + v01 * (wx1 * wy0) + v10 * (wx0 * wy1) + v11 * (wx1 * wy1);
}

template <typename TS, typename TD, int ROWS>
__global__ void __launch_bounds__(TILE_THREADS) fov_crop_tiled_kernel(const TiledArgs ta) {
  extern __shared__ float tmp[];  // [3][ROWS][pitch]
  const Args& a = ta.a;
  const int n = blockIdx.y;
  const int r0 = blockIdx.x * ROWS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = a.S, H = a.H, W = a.W;

  const float cx = __ldg(a.centers + 2 * n), cy = __ldg(a.centers + 2 * n + 1);
  const float fw = __ldg(a.windows + 2 * n), fh = __ldg(a.windows + 2 * n + 1);
  const long long src_frame = a.frame_ids ? __ldg(a.frame_ids + n) : n;
  const long long plane = static_cast<long long>(H) * W;
  const TS* src = reinterpret_cast<const TS*>(a.frames) + src_frame * 3ll * plane;
  const float inv_s = 1.0f / S;
  TD* out = reinterpret_cast<TD*>(a.out);

  // same expressions as the direct kernel / the oracle: sample position of output column ox / output row oy
  auto sample_x = [&](int ox) { return ((fw * ((2 * ox + 1) * inv_s - 1.0f) + (2.0f * cx - 1.0f) + 1.0f) * W - 1.0f) * 0.5f; };
  auto sample_y = [&](int oy) { return ((fh * ((2 * oy + 1) * inv_s - 1.0f) + (2.0f * cy - 1.0f) + 1.0f) * H - 1.0f) * 0.5f; };

  // ---- this thread's two output columns (pass H) ------------------------------------------------
  const int pairs = S >> 1;
  const int groups = TILE_THREADS / pairs;            // row groups that fit the CTA (S = 256 -> 2, 224 -> 2, 64 -> 8)
  const int pg = tid / pairs, pp = tid - pg * pairs;  // one integer division per thread
  const int ox = pp << 1;
  const bool mirrored = !(fw > 0.0f);                 // CTA-uniform; sx then does not grow with ox: direct gather below
  // Source-column span the window covers, clipped to the frame; the first column is rounded down to an even one (vector
  // loads).  sx grows with ox, so every in-frame tap of every output column lies in [j_lo, j_hi].
  const int j_lo = max(static_cast<int>(floorf(sample_x(0))), 0) & ~1;
  const int j_hi = min(static_cast<int>(floorf(sample_x(S - 1))) + 1, W - 1);
  const int span = j_hi - j_lo + 1;                   // <= 0: the window misses the frame horizontally
  // Column taps: slots (xi, xi + 1) of the staged line with weights (wa, wb).  Slot `span` is zero-filled by pass V, so an
  // in-frame tap a at the last column can read its (out-of-frame, weight 0) neighbour safely.
  int xi[2];
  float wa[2], wb[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float sx = sample_x(ox + i);
    const float fx0 = floorf(sx);
    const int x0 = static_cast<int>(fx0);
    const float w1 = sx - fx0, w0 = 1.0f - w1;
    const bool a_in = x0 >= 0 && x0 < W, b_in = x0 + 1 >= 0 && x0 + 1 < W;
    if (a_in) { xi[i] = x0 - j_lo; wa[i] = w0; wb[i] = b_in ? w1 : 0.0f; }
    else if (b_in) { xi[i] = 0; wa[i] = w1; wb[i] = 0.0f; }   // x0 == -1: column 0 is slot 0 (j_lo == 0)
    else { xi[i] = 0; wa[i] = 0.0f; wb[i] = 0.0f; }
  }

  // ---- rows of this tile (out-of-frame rows: weight 0, row index clamped into the frame) -------
  int yar[ROWS], ybr[ROWS];
  float wy0r[ROWS], wy1r[ROWS];
  bool any_row = false;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const float sy = sample_y(min(r0 + r, S - 1));
    const float fy0 = floorf(sy);
    const int y0 = static_cast<int>(fy0);
    const float w1 = sy - fy0;
    const bool a_in = y0 >= 0 && y0 < H, b_in = y0 + 1 >= 0 && y0 + 1 < H;
    yar[r] = min(max(y0, 0), H - 1);
    ybr[r] = min(max(y0 + 1, 0), H - 1);
    wy0r[r] = a_in ? 1.0f - w1 : 0.0f;
    wy1r[r] = b_in ? w1 : 0.0f;
    any_row = any_row || a_in || b_in;
  }
  const bool constant_tile = !mirrored && (!any_row || span <= 0);  // CTA-uniform: the whole tile samples the zero padding

  // ---- pass V: vertical blend of the window's source columns into shared memory ------------------
  if (!constant_tile && !mirrored) {
    const int npairs = (span + 1) >> 1;
    const int half = (npairs + 1) >> 1;
    for (int item = warp; item < 3 * ROWS * 2; item += TILE_THREADS / 32) {  // (channel, row, half line) per warp
      const int rc = item >> 1, hsel = item & 1;
      const int c = rc / ROWS, r = rc - c * ROWS;
      float* line = tmp + rc * ta.pitch;
      const float w0 = wy0r[r], w1 = wy1r[r];
      const TS* rowa = src + c * plane + static_cast<long long>(yar[r]) * W + j_lo;
      const TS* rowb = src + c * plane + static_cast<long long>(ybr[r]) * W + j_lo;
      const int p_begin = hsel * half, p_end = min(npairs, p_begin + half);
      if (ta.vec2) {
        for (int jp = p_begin + lane; jp < p_end; jp += 32) {
          const int j = jp << 1;
          float2 va, vb;
          if (j + 1 < span) { va = load_px2<TS>(rowa + j); vb = load_px2<TS>(rowb + j); }
          else { va = make_float2(load_px<TS>(rowa + j), 0.0f); vb = make_float2(load_px<TS>(rowb + j), 0.0f); }
          *reinterpret_cast<float2*>(line + j) = make_float2(va.x * w0 + vb.x * w1, va.y * w0 + vb.y * w1);
        }
      } else {
        for (int j = 2 * p_begin + lane; j < min(span, 2 * p_end); j += 32) line[j] = load_px<TS>(rowa + j) * w0 + load_px<TS>(rowb + j) * w1;
      }
      if (hsel == 1 && lane == 0) line[span] = 0.0f;  // (an odd span has already written this slot with the same 0)
    }
  }
  __syncthreads();
  if (pg >= groups) return;

  // ---- pass H: horizontal blend, normalise, store -----------------------------------------------
  long long obase, row_stride, ch_stride;
  int patch_rows_left = 1 << 30;  // patch-major: rows until the tile crosses into the next row of patches
  if (a.patch > 0) {
    const int px = fast_div(ox, a.patch_magic), ix = ox - px * a.patch;
    const int py = fast_div(r0, a.patch_magic), iy = r0 - py * a.patch;
    obase = (static_cast<long long>(n) * a.G * a.G + py * a.G + px) * a.out_ld + iy * a.patch + ix;
    row_stride = a.patch;
    ch_stride = static_cast<long long>(a.patch) * a.patch;
    patch_rows_left = a.patch - iy;
  } else {
    obase = (static_cast<long long>(n) * 3 * S + r0) * S + ox;
    row_stride = S;
    ch_stride = static_cast<long long>(S) * S;
  }
  for (int r = pg; r < ROWS; r += groups) {
    if (r0 + r >= S) break;
    long long o = obase + r * row_stride;
    if (r >= patch_rows_left)  // next row of patches: + G patches, back to row (r - patch_rows_left) of that patch
      o += static_cast<long long>(a.G) * a.out_ld - static_cast<long long>(a.patch) * a.patch;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v0 = 0.0f, v1 = 0.0f;
      if (mirrored) {
        const float sy = sample_y(r0 + r);
        v0 = direct_sample<TS>(src + c * plane, sample_x(ox), sy, H, W);
        v1 = direct_sample<TS>(src + c * plane, sample_x(ox + 1), sy, H, W);
      } else if (!constant_tile) {
        const float* line = tmp + (c * ROWS + r) * ta.pitch;
        v0 = line[xi[0]] * wa[0] + line[xi[0] + 1] * wb[0];
        v1 = line[xi[1]] * wa[1] + line[xi[1] + 1] * wb[1];
      }
      store2(out + o + c * ch_stride, (v0 - a.mean[c]) * a.inv_std[c], (v1 - a.mean[c]) * a.inv_std[c]);
    }
  }
}

// RF_CROP_TILED=1 selects the tiled kernel (read per call so tests can toggle it).  Off by default: its first version is
// latency-bound in pass V and measured 2.6x slower than the direct gather (profiles/r2_bench_crop_micro_v1_tiled_kernel.json).
static bool tiled_enabled() {
  const char* e = getenv("RF_CROP_TILED");
  return e && e[0] == '1';
}

template <typename TS, typename TD>
static int launch_tiled(const RfFovCropParams* p, const Args& a, cudaStream_t s) {
  TiledArgs ta;
  ta.a = a;
  ta.pitch = (p->W + 2 + 1) & ~1;
  ta.vec2 = (p->W % 2 == 0) && (reinterpret_cast<uintptr_t>(p->frames) % (2 * sizeof(TS)) == 0);
  // 8 output rows per tile while the staging lines stay small (more outputs per thread for the same tap setup); wide frames
  // (DR(eye)VE 768, full-resolution 1088) use 4 rows to keep several tiles resident per SM
  if (p->W <= 512) {
    const size_t smem = sizeof(float) * 3 * 8 * ta.pitch;
    RF_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(fov_crop_tiled_kernel<TS, TD, 8>), smem));
    fov_crop_tiled_kernel<TS, TD, 8><<<dim3(ceil_div(p->out_size, 8), p->n_frames), TILE_THREADS, smem, s>>>(ta);
  } else {
    const size_t smem = sizeof(float) * 3 * 4 * ta.pitch;
    RF_CHECK_ARG(smem <= 200 * 1024, "rf_fov_crop: frames of width %d need %zu B of staging memory", p->W, smem);
    RF_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(fov_crop_tiled_kernel<TS, TD, 4>), smem));
    fov_crop_tiled_kernel<TS, TD, 4><<<dim3(ceil_div(p->out_size, 4), p->n_frames), TILE_THREADS, smem, s>>>(ta);
  }
  RF_LAUNCH_OK();
  return RF_OK;
}

template <typename TS>
static int dispatch_out(const RfFovCropParams* p, const Args& a, cudaStream_t s) {
  // tiled kernel: 2 px per thread -> even S with at least one pair row per CTA; a patch no shorter than a tile (one patch-row
  // crossing per tile at most)
  if (tiled_enabled() && p->out_size >= 8 && (p->patch == 0 || p->patch >= 8)) {
    if (p->out_dtype == RF_F32) return launch_tiled<TS, float>(p, a, s);
    if (p->out_dtype == RF_F16) return launch_tiled<TS, __half>(p, a, s);
    return launch_tiled<TS, __nv_bfloat16>(p, a, s);
  }
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
    case RF_U8_F16: return crop::dispatch_out<crop::u8h>(p, a, s);
    default: set_error("rf_fov_crop: unsupported src_dtype %d", p->src_dtype); return RF_ERR_INVALID_ARGUMENT;
  }
}
