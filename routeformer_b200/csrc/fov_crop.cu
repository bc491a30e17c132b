// Field-of-view crop / resample (HBM-bound).  See include/routeformer_b200.h (1).
//
// Every output pixel is a bilinear sample of the source frame at
//   sx = ((fw * gx + 2cx - 1 + 1) * W - 1) / 2,  gx = (2 ox + 1) / S - 1   (grid_sample, align_corners=False)
// with zeros outside the frame, then normalised per channel, written either planar [n,3,S,S] or directly in the
// patch-major layout the patch-embedding GEMM consumes as its A operand (no im2col pass).
//
// fov_crop_walk_kernel (default): the bilinear resample is SEPARABLE and its column taps are shared by the three channels.
// A thread owns two adjacent output columns and walks down the output rows of its CTA's strip with the horizontal blends
// of the current two source rows in registers (no shared-memory staging); see the kernel's comment.
// fov_crop_kernel (the round-1 direct 4-tap gather, issue-bound at 0.3 of the HBM roofline) remains behind RF_CROP_DIRECT=1.
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
// the same in two steps -- the raw element as loaded, converted later -- for kernels that keep loads in flight across other work
template <typename T> struct RawPx { typedef T type; };
template <> struct RawPx<u8h> { typedef unsigned char type; };
template <typename T> __device__ __forceinline__ typename RawPx<T>::type load_raw(const T* p) { return __ldg(p); }
template <> __device__ __forceinline__ unsigned char load_raw<u8h>(const u8h* p) { return __ldg(&p->v); }
template <typename T> __device__ __forceinline__ typename RawPx<T>::type zero_raw() { return static_cast<typename RawPx<T>::type>(0); }
template <> __device__ __forceinline__ __half zero_raw<__half>() { return __ushort_as_half(0); }
template <typename T> __device__ __forceinline__ float cvt_px(typename RawPx<T>::type r);
template <> __device__ __forceinline__ float cvt_px<__half>(__half r) { return __half2float(r); }
template <> __device__ __forceinline__ float cvt_px<float>(float r) { return r; }
template <> __device__ __forceinline__ float cvt_px<unsigned char>(unsigned char r) { return r * (1.0f / 255.0f); }
template <> __device__ __forceinline__ float cvt_px<u8h>(unsigned char r) { return __half2float(__float2half_rn(r * (1.0f / 255.0f))); }

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

// ---- column-walker kernel (default): separable blend, everything in registers ---------------------------------
// packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: two IEEE fp32 operations per issue slot)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float2 upk(f32x2 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

__device__ __forceinline__ void store2(float* dst, float a, float b) { *reinterpret_cast<float2*>(dst) = make_float2(a, b); }
__device__ __forceinline__ void store2(__half* dst, float a, float b) { *reinterpret_cast<__half2*>(dst) = __floats2half2_rn(a, b); }
__device__ __forceinline__ void store2(__nv_bfloat16* dst, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(a, b);
}

constexpr int WALK_ROWS = 32;     // output rows per CTA
constexpr int WALK_THREADS = 128; // one thread per pair of adjacent output columns (S <= 256)
struct __align__(16) WalkRow {    // per output row of the CTA's strip, computed once, read as one broadcast 16 B shared load
  int y0;          // floor of the sample row
  float wy0, wy1;  // weights of source rows y0 / y0 + 1
  int off;         // output offset of the row inside the frame (elements; the thread adds its column / channel part)
};

// The horizontal blend of one source row for the thread's two output columns, all three channels:
//   h[c] = (src[c][y][xa] * wa0 + src[c][y][xa + 1] * wa1,  src[c][y][xb] * wb0 + src[c][y][xb + 1] * wb1)
// Three per-thread variants (fixed for the whole walk):
//   WALK_NARROW: xb - xa in {0, 1} and taps xa .. xa+2 inside the frame (every interior thread of an up-sampling crop): three
//                loads per channel from one address, the second column's weights laid over the three taps (one of them is 0);
//   WALK_WIDE  : all four taps inside the frame, any distance: two addresses per channel, immediate offsets;
//   WALK_EDGE  : taps outside the frame read nothing and contribute exactly zero.
// Addresses are signed 32-bit element offsets from the frame base (taps left of the frame give small negative offsets that
// are never dereferenced).
constexpr int WALK_NARROW = 0, WALK_WIDE = 1, WALK_EDGE = 2;
// base + off * sizeof(T) as ONE instruction (the compiler otherwise re-derives the 64-bit frame base for every address)
template <typename T>
__device__ __forceinline__ T* at_s32(T* base, int off) {
  unsigned long long r;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(off), "n"(sizeof(T)), "l"(reinterpret_cast<unsigned long long>(base)));
  return reinterpret_cast<T*>(r);
}
struct WalkCols {
  int k[3][2];       // per channel: c * plane + xa, c * plane + xb
  bool ok[4];        // taps xa, xa+1, xb, xb+1 inside the frame (EDGE)
  float w[5];        // NARROW: wa0, wa1, wb'0, wb'1, wb'2;  otherwise wa0, wa1, wb0, wb1
};
// raw (unconverted) taps of one source row, all three channels; a row outside the frame is all zeros
template <typename TS, int MODE>
struct WalkTaps { typename RawPx<TS>::type v[3][MODE == WALK_NARROW ? 3 : 4]; };
template <typename TS, int MODE>
__device__ __forceinline__ void walk_load(const TS* __restrict__ src, int y, int H, int W, const WalkCols& q, WalkTaps<TS, MODE>& t) {
  const typename RawPx<TS>::type zero = zero_raw<TS>();
  if (y >= 0 && y < H) {  // CTA-uniform
    const int rowoff = y * W;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const TS* pa = at_s32(src, rowoff + q.k[c][0]);
      if (MODE == WALK_NARROW) {
        t.v[c][0] = load_raw<TS>(pa); t.v[c][1] = load_raw<TS>(pa + 1); t.v[c][2] = load_raw<TS>(pa + 2);
      } else {
        const TS* pb = at_s32(src, rowoff + q.k[c][1]);
        if (MODE == WALK_EDGE) {
          t.v[c][0] = q.ok[0] ? load_raw<TS>(pa) : zero; t.v[c][1] = q.ok[1] ? load_raw<TS>(pa + 1) : zero;
          t.v[c][2] = q.ok[2] ? load_raw<TS>(pb) : zero; t.v[c][3] = q.ok[3] ? load_raw<TS>(pb + 1) : zero;
        } else {
          t.v[c][0] = load_raw<TS>(pa); t.v[c][1] = load_raw<TS>(pa + 1);
          t.v[c][2] = load_raw<TS>(pb); t.v[c][3] = load_raw<TS>(pb + 1);
        }
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int k = 0; k < (MODE == WALK_NARROW ? 3 : 4); ++k) t.v[c][k] = zero;
  }
}
template <typename TS, int MODE>
__device__ __forceinline__ void walk_blend(const WalkTaps<TS, MODE>& t, const WalkCols& q, f32x2 (&h)[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v[4];
#pragma unroll
    for (int k = 0; k < (MODE == WALK_NARROW ? 3 : 4); ++k) v[k] = cvt_px<TS>(t.v[c][k]);
    if (MODE == WALK_NARROW) h[c] = pk(fmaf(v[1], q.w[1], v[0] * q.w[0]), fmaf(v[2], q.w[4], fmaf(v[1], q.w[3], v[0] * q.w[2])));
    else h[c] = pk(fmaf(v[1], q.w[1], v[0] * q.w[0]), fmaf(v[3], q.w[3], v[2] * q.w[2]));
  }
}

// A CTA owns WALK_ROWS output rows of one frame; a thread owns two adjacent output columns and walks DOWN the rows.
// The bilinear resample is separable: the horizontal blend h of a source row (fixed columns and weights for the whole walk)
// is computed once and stays in registers for every output row that uses it -- an up-sampling crop (0.72 source rows per
// output row for the gaze window) needs a new source row for ~3 of 4 output rows and shifts the previous one down; the
// vertical blend, the normalisation and the store are packed f32x2 operations on the column pair.  ~12 instructions per
// output value against ~40 for the direct 4-tap gather, no shared-memory staging and no barrier after the row table; a warp
// reads ~100 contiguous bytes per load instruction and writes 128 (fp16 / bf16) or 256 (fp32).
// The taps of source row cur_y + 2 are always IN FLIGHT while the current output rows are blended and stored (software
// prefetch into registers): a thread's walk is a serial chain of ~24 source rows, and without it every one of them would
// expose a full memory round trip.
// All control flow in the walk is CTA-uniform.  Mirrored windows (fw, fh < 0) take the same code: column taps are per-thread
// constants of any order, and a row that is neither the current nor the next one simply reloads both source rows.
// (A NaN pixel in the source can reach one output column more than in the reference: NARROW multiplies its third tap by 0.)
template <typename TS, typename TD, int MODE>
__device__ __forceinline__ void walk_rows(const Args& a, const WalkRow* __restrict__ rows, int nrows, const TS* __restrict__ src,
                                          const WalkCols& q, TD* __restrict__ out0, long long ch_stride) {
  const int H = a.H, W = a.W;
  f32x2 hA[3], hB[3];
  f32x2 shift[3], istd[3];  // (x - mean) * inv_std as one fused multiply-add
  TD* oc[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    shift[c] = pk(-a.mean[c] * a.inv_std[c], -a.mean[c] * a.inv_std[c]);
    istd[c] = pk(a.inv_std[c], a.inv_std[c]);
    oc[c] = out0 + c * ch_stride;
  }
  int cur_y = -(1 << 30);
  WalkTaps<TS, MODE> nxt;  // taps of source row cur_y + 2
#pragma unroll 1
  for (int r = 0; r < nrows; ++r) {
    const WalkRow ri = rows[r];
    if (ri.y0 != cur_y) {
      if (ri.y0 == cur_y + 1) {
        hA[0] = hB[0]; hA[1] = hB[1]; hA[2] = hB[2];
        walk_blend<TS, MODE>(nxt, q, hB);
      } else {
        WalkTaps<TS, MODE> ta, tb;
        walk_load<TS, MODE>(src, ri.y0, H, W, q, ta);
        walk_load<TS, MODE>(src, ri.y0 + 1, H, W, q, tb);
        walk_blend<TS, MODE>(ta, q, hA);
        walk_blend<TS, MODE>(tb, q, hB);
      }
      cur_y = ri.y0;
      walk_load<TS, MODE>(src, cur_y + 2, H, W, q, nxt);
    }
    const f32x2 wy0 = pk(ri.wy0, ri.wy0), wy1 = pk(ri.wy1, ri.wy1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float2 o = upk(fma2(fma2(hB[c], wy1, mul2(hA[c], wy0)), istd[c], shift[c]));
      store2(oc[c] + ri.off, o.x, o.y);
    }
  }
}

template <typename TS, typename TD>
__global__ void __launch_bounds__(WALK_THREADS) fov_crop_walk_kernel(const Args a) {
  __shared__ WalkRow s_rows[WALK_ROWS];
  const int n = blockIdx.y, r0 = blockIdx.x * WALK_ROWS, tid = threadIdx.x;
  const int S = a.S, H = a.H, W = a.W;
  const int nrows = min(WALK_ROWS, S - r0);
  const float cx = __ldg(a.centers + 2 * n), cy = __ldg(a.centers + 2 * n + 1);
  const float fw = __ldg(a.windows + 2 * n), fh = __ldg(a.windows + 2 * n + 1);
  const long long src_frame = a.frame_ids ? __ldg(a.frame_ids + n) : n;
  const int plane = H * W;
  const TS* src = reinterpret_cast<const TS*>(a.frames) + src_frame * 3ll * plane;
  const float inv_s = 1.0f / S;

  if (tid < nrows) {  // the strip's row table; same expressions as the direct kernel / the oracle
    const int oy = r0 + tid;
    const float gy = (2 * oy + 1) * inv_s - 1.0f;
    const float sy = ((fh * gy + (2.0f * cy - 1.0f) + 1.0f) * H - 1.0f) * 0.5f;
    const float fy0 = floorf(sy);
    WalkRow ri;
    // rows far outside the frame all behave like "both taps out": clamped so that y0 + 1 cannot overflow
    ri.y0 = static_cast<int>(fminf(fmaxf(fy0, -3.0f), static_cast<float>(H)));
    ri.wy1 = sy - fy0;
    ri.wy0 = 1.0f - ri.wy1;
    if (a.patch > 0) {
      const int py = fast_div(oy, a.patch_magic), iy = oy - py * a.patch;
      ri.off = py * a.G * static_cast<int>(a.out_ld) + iy * a.patch;
    } else {
      ri.off = oy * S;
    }
    s_rows[tid] = ri;
  }

  // column taps of the thread's two pixels
  const int ox = 2 * tid;
  int x0[2];
  float w0[2], w1[2];
  WalkCols q;
  bool interior = true;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float gx = (2 * (ox + i) + 1) * inv_s - 1.0f;
    const float sx = ((fw * gx + (2.0f * cx - 1.0f) + 1.0f) * W - 1.0f) * 0.5f;
    const float fx0 = floorf(sx);
    x0[i] = static_cast<int>(fminf(fmaxf(fx0, -2.0f), static_cast<float>(W)));
    w1[i] = sx - fx0;
    w0[i] = 1.0f - w1[i];
    q.ok[2 * i] = x0[i] >= 0 && x0[i] < W;
    q.ok[2 * i + 1] = x0[i] + 1 >= 0 && x0[i] + 1 < W;
    interior = interior && q.ok[2 * i] && q.ok[2 * i + 1];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    q.k[c][0] = c * plane + x0[0];
    q.k[c][1] = c * plane + x0[1];
  }
  const int d = x0[1] - x0[0];
  const bool narrow = interior && (d == 0 || d == 1) && x0[0] + 2 < W;
  q.w[0] = w0[0]; q.w[1] = w1[0];
  if (narrow) {
    q.w[2] = d ? 0.0f : w0[1];
    q.w[3] = d ? w0[1] : w1[1];
    q.w[4] = d ? w1[1] : 0.0f;
  } else {
    q.w[2] = w0[1]; q.w[3] = w1[1]; q.w[4] = 0.0f;
  }
  __syncthreads();
  if (ox >= S) return;

  TD* out = reinterpret_cast<TD*>(a.out);
  long long ch_stride;
  if (a.patch > 0) {
    const int px = fast_div(ox, a.patch_magic), ix = ox - px * a.patch;
    out += (static_cast<long long>(n) * a.G * a.G + px) * a.out_ld + ix;
    ch_stride = static_cast<long long>(a.patch) * a.patch;
  } else {
    out += static_cast<long long>(n) * 3 * S * S + ox;
    ch_stride = static_cast<long long>(S) * S;
  }
  if (narrow) walk_rows<TS, TD, WALK_NARROW>(a, s_rows, nrows, src, q, out, ch_stride);
  else if (interior) walk_rows<TS, TD, WALK_WIDE>(a, s_rows, nrows, src, q, out, ch_stride);
  else walk_rows<TS, TD, WALK_EDGE>(a, s_rows, nrows, src, q, out, ch_stride);
}

// RF_CROP_DIRECT=1 selects the round-1 direct gather (read per call so tests and A/B runs can toggle it).
static bool direct_forced() {
  const char* e = getenv("RF_CROP_DIRECT");
  return e && e[0] == '1';
}

template <typename TS, typename TD>
static int launch_walk(const RfFovCropParams* p, const Args& a, cudaStream_t s) {
  const int threads = ((p->out_size / 2 + 31) / 32) * 32;
  fov_crop_walk_kernel<TS, TD><<<dim3(ceil_div(p->out_size, WALK_ROWS), p->n_frames), threads, 0, s>>>(a);
  RF_LAUNCH_OK();
  return RF_OK;
}

template <typename TS>
static int dispatch_out(const RfFovCropParams* p, const Args& a, cudaStream_t s) {
  // the walker keeps per-frame source and output offsets in signed 32 bits.  Its gain is the re-use of source rows down the
  // output rows; a frame more than twice as tall as the crop is (for the usual windows of ~half a frame or more) down-sampled,
  // every output row needs two new source rows and the direct gather with its independent threads is the faster kernel
  // (full-resolution 1080 x 1088 frames: 0.051 ms direct vs 0.072 ms walker per 128 frames)
  const long long frame_elems = p->patch > 0 ? static_cast<long long>(a.G) * a.G * p->out_ld : 3ll * p->out_size * p->out_size;
  const bool walk_forced = getenv("RF_CROP_WALK") && getenv("RF_CROP_WALK")[0] == '1';
  if (!direct_forced() && frame_elems < (1ll << 31) && 3ll * p->H * p->W + 4 < (1ll << 31) && (p->H <= 2 * p->out_size || walk_forced)) {
    if (p->out_dtype == RF_F32) return launch_walk<TS, float>(p, a, s);
    if (p->out_dtype == RF_F16) return launch_walk<TS, __half>(p, a, s);
    return launch_walk<TS, __nv_bfloat16>(p, a, s);
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
