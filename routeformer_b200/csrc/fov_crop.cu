// Field-of-view crop / resample (HBM-bound).  See include/routeformer_b200.h (1).
//
// Every output pixel is a bilinear sample of the source frame at
//   sx = ((fw * gx + 2cx - 1 + 1) * W - 1) / 2,  gx = (2 ox + 1) / S - 1   (grid_sample, align_corners=False)
// with zeros outside the frame, then normalised per channel, written either planar [n,3,S,S] or directly in the
// patch-major layout the patch-embedding GEMM consumes as its A operand (no im2col pass).
//
// fov_crop_walk4_kernel / fov_crop_walk_kernel (default): the bilinear resample is SEPARABLE and its column taps are shared by the
// three channels.  The source window of a strip of output rows is staged RAW in shared memory with cp.async; a thread owns four
// (walk4) or two adjacent output columns and walks down the output rows of the strip with the horizontal blends of the current
// two source rows in registers; see the kernels' comments.  Frames whose rows are not 4-byte aligned take the same walk from
// global memory.
// fov_crop_kernel (the round-1 direct 4-tap gather, issue-bound at 0.3 of the HBM roofline on the gaze crop) remains for frames
// taller than twice the crop (down-sampling: no row re-use) and behind RF_CROP_DIRECT=1.
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

constexpr int WALK_MAX_ROWS = 32;            // output rows per CTA (28 when S is a multiple of 28 only: strips then match patch rows)
constexpr int WALK_THREADS = 128;            // one thread per pair of adjacent output columns (S <= 256)
constexpr int WALK_STAGE_BYTES = 26 * 1024;  // staged source window per CTA: 8 CTAs of 128 threads per SM
struct __align__(16) WalkRow {    // per output row of the CTA's strip, computed once, read as one broadcast 16 B shared load
  int code;        // (y0 + 4) | flags << 24;  y0 = floor of the sample row, clamped to [-3, H]
  float wy0, wy1;  // weights of source rows y0 / y0 + 1
  int off;         // output offset of the row inside the frame (elements; the thread adds its column / channel part)
};
// flags: what the walk does before it blends this row (relative to the row above it)
constexpr int WALK_NEW = 1;    // y0 differs from the previous row's: source rows must be loaded
constexpr int WALK_SHIFT = 2;  // y0 = previous y0 + 1: the previous lower row becomes the upper row, one new row is loaded
constexpr int WALK_A_IN = 4;   // source row y0 is inside the frame
constexpr int WALK_B_IN = 8;   // source row y0 + 1 is inside the frame
__device__ __forceinline__ int walk_y0(int code) { return (code & 0xFFFFFF) - 4; }

// The horizontal blend of one source row for the thread's two output columns, all three channels:
//   h[c] = (src[c][y][xa] * wa0 + src[c][y][xa + 1] * wa1,  src[c][y][xb] * wb0 + src[c][y][xb + 1] * wb1)
// Three per-thread variants (fixed for the whole walk):
//   WALK_NARROW: xb - xa in {0, 1} and taps xa .. xa+2 inside the frame (every interior thread of an up-sampling crop): three
//                loads per channel from one address, the second column's weights laid over the three taps (one of them is 0);
//   WALK_WIDE  : all four taps inside the frame, any distance: two addresses per channel, immediate offsets;
//   WALK_EDGE  : taps outside the frame read nothing and contribute exactly zero.
// Tap addresses are signed 32-bit element offsets (taps left of the frame give negative offsets that are never dereferenced)
// from the staged window in shared memory (STAGED) or from the frame in global memory.
constexpr int WALK_NARROW = 0, WALK_WIDE = 1, WALK_EDGE = 2;
// base + off * sizeof(T) as ONE instruction (the compiler otherwise re-derives the 64-bit frame base for every address)
template <typename T>
__device__ __forceinline__ T* at_s32(T* base, int off) {
  unsigned long long r;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(off), "n"(sizeof(T)), "l"(reinterpret_cast<unsigned long long>(base)));
  return reinterpret_cast<T*>(r);
}
struct WalkCols {
  int k[2];          // column offsets of xa / xb (elements)
  bool ok[4];        // taps xa, xa+1, xb, xb+1 inside the frame (EDGE)
  float w[5];        // NARROW: wa0, wa1, wb'0, wb'1, wb'2;  otherwise wa0, wa1, wb0, wb1
};
// where the taps come from.  Global: frame base pointer, strides in elements.  STAGED: 32-bit shared-memory address of the
// staged window, strides (and WalkCols::k) in BYTES.
template <typename TS>
struct WalkSrc {
  const typename RawPx<TS>::type* base;
  unsigned sbase;
  int y_lo, row_stride, ch_stride;
};
// ld.shared with an immediate byte offset (a generic pointer into shared memory makes the compiler re-derive the window base)
template <int IMM> __device__ __forceinline__ __half lds_px(unsigned addr, __half) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1+%2];" : "=h"(v) : "r"(addr), "n"(IMM));
  return __ushort_as_half(v);
}
template <int IMM> __device__ __forceinline__ float lds_px(unsigned addr, float) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(IMM));
  return v;
}
template <int IMM> __device__ __forceinline__ unsigned char lds_px(unsigned addr, unsigned char) {
  unsigned v;
  asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(IMM));
  return static_cast<unsigned char>(v);
}
template <typename TS, int MODE, bool STAGED>
__device__ __forceinline__ void walk_row(const WalkSrc<TS>& ws, int rowoff, const WalkCols& q, f32x2 (&h)[3]) {
  constexpr int NT = MODE == WALK_NARROW ? 3 : 4;
  typedef typename RawPx<TS>::type Raw;
  constexpr int ES = sizeof(Raw);
  Raw t[3][NT];
  const Raw zero = zero_raw<TS>();
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int oa = rowoff + c * ws.ch_stride + q.k[0], ob = rowoff + c * ws.ch_stride + q.k[1];
    if (STAGED) {
      const unsigned pa = ws.sbase + oa, pb = ws.sbase + ob;
      if (MODE == WALK_NARROW) {
        t[c][0] = lds_px<0>(pa, zero); t[c][1] = lds_px<ES>(pa, zero); t[c][2] = lds_px<2 * ES>(pa, zero);
      } else if (MODE == WALK_EDGE) {
        t[c][0] = q.ok[0] ? lds_px<0>(pa, zero) : zero; t[c][1] = q.ok[1] ? lds_px<ES>(pa, zero) : zero;
        t[c][2] = q.ok[2] ? lds_px<0>(pb, zero) : zero; t[c][3] = q.ok[3] ? lds_px<ES>(pb, zero) : zero;
      } else {
        t[c][0] = lds_px<0>(pa, zero); t[c][1] = lds_px<ES>(pa, zero);
        t[c][2] = lds_px<0>(pb, zero); t[c][3] = lds_px<ES>(pb, zero);
      }
    } else {
      const Raw* pa = at_s32(ws.base, oa);
      if (MODE == WALK_NARROW) {
        t[c][0] = __ldg(pa); t[c][1] = __ldg(pa + 1); t[c][2] = __ldg(pa + 2);
      } else {
        const Raw* pb = at_s32(ws.base, ob);
        if (MODE == WALK_EDGE) {
          t[c][0] = q.ok[0] ? __ldg(pa) : zero; t[c][1] = q.ok[1] ? __ldg(pa + 1) : zero;
          t[c][2] = q.ok[2] ? __ldg(pb) : zero; t[c][3] = q.ok[3] ? __ldg(pb + 1) : zero;
        } else {
          t[c][0] = __ldg(pa); t[c][1] = __ldg(pa + 1);
          t[c][2] = __ldg(pb); t[c][3] = __ldg(pb + 1);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) v[k] = cvt_px<TS>(t[c][k]);
    if (MODE == WALK_NARROW) h[c] = pk(fmaf(v[1], q.w[1], v[0] * q.w[0]), fmaf(v[2], q.w[4], fmaf(v[1], q.w[3], v[0] * q.w[2])));
    else h[c] = pk(fmaf(v[1], q.w[1], v[0] * q.w[0]), fmaf(v[3], q.w[3], v[2] * q.w[2]));
  }
}

// Output rows [ra, rb) of the strip for one thread (two adjacent output columns, three channels), walking DOWN the rows.
// The bilinear resample is separable: the horizontal blend h of a source row (fixed columns and weights for the whole walk)
// is computed once and stays in registers for every output row that uses it -- an up-sampling crop (0.72 source rows per
// output row for the gaze window) needs a new source row for ~3 of 4 output rows; the vertical blend, the normalisation and
// the store are packed f32x2 operations on the column pair: ~11 instructions per output value against ~40 for the direct
// 4-tap gather.  A warp writes 128 (fp16 / bf16) or 256 (fp32) contiguous bytes per store.
// What happens before a row is blended (nothing / shift + one new source row / two new rows) was decided once per strip and
// sits in the row table (WalkRow::code): all control flow is CTA-uniform.  The two row blends live in h[0] / h[1] and swap
// roles at a shift (the loop body exists twice), so a shift moves no registers.
// Mirrored windows (fw, fh < 0) take the same code: column taps are per-thread constants of any order, and a row that is
// neither the current nor the next one reloads both source rows.
// (A NaN pixel in the source can reach one output column more than in the reference: NARROW multiplies its third tap by 0.)
template <typename TS, typename TD, int MODE, bool STAGED>
__device__ __forceinline__ void walk_rows(const Args& a, const WalkRow* __restrict__ rows, int ra, int rb, const WalkSrc<TS>& ws,
                                          const WalkCols& q, TD* __restrict__ out0, long long ch_stride) {
  f32x2 h[2][3];
  f32x2 shift[3], istd[3];  // (x - mean) * inv_std as one fused multiply-add
  TD* oc[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    shift[c] = pk(-a.mean[c] * a.inv_std[c], -a.mean[c] * a.inv_std[c]);
    istd[c] = pk(a.inv_std[c], a.inv_std[c]);
    oc[c] = out0 + c * ch_stride;
  }
  auto load = [&](bool in_frame, int rowoff, f32x2 (&dst)[3]) {
    // (CTA-uniform; the staged window holds zero lines for the rows outside the frame: no test at all)
    if (STAGED || in_frame) walk_row<TS, MODE, STAGED>(ws, rowoff, q, dst);
    else dst[0] = dst[1] = dst[2] = 0ull;
  };
  auto emit = [&](const WalkRow& e, const f32x2 (&hA)[3], const f32x2 (&hB)[3]) {
    const f32x2 wy0 = pk(e.wy0, e.wy0), wy1 = pk(e.wy1, e.wy1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float2 o = upk(fma2(fma2(hB[c], wy1, mul2(hA[c], wy0)), istd[c], shift[c]));
      store2(oc[c] + e.off, o.x, o.y);
    }
  };
  int r = ra;
  WalkRow e = rows[r];
  {  // the first row of a walk always loads both of its source rows
    const int rowoff = (walk_y0(e.code) - ws.y_lo) * ws.row_stride;
    load((e.code >> 24) & WALK_A_IN, rowoff, h[0]);
    load((e.code >> 24) & WALK_B_IN, rowoff + ws.row_stride, h[1]);
  }
  // one copy of the loop body per assignment of (upper, lower) source row to (h[A], h[B]); a shift leaves to the other copy
#define RF_WALK_BODY(A, B)                                                              \
  for (;;) {                                                                            \
    emit(e, h[A], h[B]);                                                                \
    if (++r >= rb) return;                                                              \
    e = rows[r];                                                                        \
    const int f = e.code >> 24;                                                         \
    if (f & WALK_NEW) {                                                                 \
      const int rowoff = (walk_y0(e.code) - ws.y_lo) * ws.row_stride;                   \
      if (f & WALK_SHIFT) {                                                             \
        load(f & WALK_B_IN, rowoff + ws.row_stride, h[A]);                              \
        break;                                                                          \
      }                                                                                 \
      load(f & WALK_A_IN, rowoff, h[A]);                                                \
      load(f & WALK_B_IN, rowoff + ws.row_stride, h[B]);                                \
    }                                                                                   \
  }
  for (;;) {
    RF_WALK_BODY(0, 1)
    RF_WALK_BODY(1, 0)
  }
#undef RF_WALK_BODY
}
template <typename TS, typename TD, bool STAGED>
__device__ __forceinline__ void walk_dispatch(int mode, const Args& a, const WalkRow* rows, int ra, int rb, const WalkSrc<TS>& ws,
                                              const WalkCols& q, TD* out0, long long ch_stride) {
  if (mode == WALK_NARROW) walk_rows<TS, TD, WALK_NARROW, STAGED>(a, rows, ra, rb, ws, q, out0, ch_stride);
  else if (mode == WALK_WIDE) walk_rows<TS, TD, WALK_WIDE, STAGED>(a, rows, ra, rb, ws, q, out0, ch_stride);
  else walk_rows<TS, TD, WALK_EDGE, STAGED>(a, rows, ra, rb, ws, q, out0, ch_stride);
}

// Four adjacent output columns per thread (two pairs, both NARROW): the same walk with the per-row bookkeeping (row-table entry,
// flags, branches, output addresses) paid once for twelve output values instead of six, and one 8 / 16-byte store per channel.
template <typename TS, typename TD>
__device__ __forceinline__ void walk_rows4(const Args& a, const WalkRow* __restrict__ rows, int ra, int rb, const WalkSrc<TS>& ws,
                                           const WalkCols& qa, const WalkCols& qb, TD* __restrict__ out0, long long ch_stride) {
  f32x2 ha[2][3], hb[2][3];  // pair A / pair B: [upper | lower source row][channel]
  f32x2 shift[3], istd[3];
  TD* oc[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    shift[c] = pk(-a.mean[c] * a.inv_std[c], -a.mean[c] * a.inv_std[c]);
    istd[c] = pk(a.inv_std[c], a.inv_std[c]);
    oc[c] = out0 + c * ch_stride;
  }
  auto load = [&](int rowoff, f32x2 (&da)[3], f32x2 (&db)[3]) {
    walk_row<TS, WALK_NARROW, true>(ws, rowoff, qa, da);
    walk_row<TS, WALK_NARROW, true>(ws, rowoff, qb, db);
  };
  auto emit = [&](const WalkRow& e, const f32x2 (&ua)[3], const f32x2 (&la)[3], const f32x2 (&ub)[3], const f32x2 (&lb)[3]) {
    const f32x2 wy0 = pk(e.wy0, e.wy0), wy1 = pk(e.wy1, e.wy1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float2 o0 = upk(fma2(fma2(la[c], wy1, mul2(ua[c], wy0)), istd[c], shift[c]));
      const float2 o1 = upk(fma2(fma2(lb[c], wy1, mul2(ub[c], wy0)), istd[c], shift[c]));
      const float v[4] = {o0.x, o0.y, o1.x, o1.y};
      store4(oc[c] + e.off, v);
    }
  };
  int r = ra;
  WalkRow e = rows[r];
  {
    const int rowoff = (walk_y0(e.code) - ws.y_lo) * ws.row_stride;
    load(rowoff, ha[0], hb[0]);
    load(rowoff + ws.row_stride, ha[1], hb[1]);
  }
#define RF_WALK4_BODY(A, B)                                                             \
  for (;;) {                                                                            \
    emit(e, ha[A], ha[B], hb[A], hb[B]);                                                \
    if (++r >= rb) return;                                                              \
    e = rows[r];                                                                        \
    const int f = e.code >> 24;                                                         \
    if (f & WALK_NEW) {                                                                 \
      const int rowoff = (walk_y0(e.code) - ws.y_lo) * ws.row_stride;                   \
      if (f & WALK_SHIFT) {                                                             \
        load(rowoff + ws.row_stride, ha[A], hb[A]);                                     \
        break;                                                                          \
      }                                                                                 \
      load(rowoff, ha[A], hb[A]);                                                       \
      load(rowoff + ws.row_stride, ha[B], hb[B]);                                       \
    }                                                                                   \
  }
  for (;;) {
    RF_WALK4_BODY(0, 1)
    RF_WALK4_BODY(1, 0)
  }
#undef RF_WALK4_BODY
}

// column taps of the pair (ox, ox + 1): weights, tap validity, mode; x0 = first tap column of either pixel
template <typename FX, typename FC>
__device__ __forceinline__ int walk_setup_pair(int ox, int W, FX sample_x, FC floor_clamped, WalkCols& q, int (&x0)[2]) {
  float w0[2], w1[2];
  bool interior = true;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float sx = sample_x(ox + i);
    x0[i] = floor_clamped(sx, W);
    w1[i] = sx - floorf(sx);
    w0[i] = 1.0f - w1[i];
    q.ok[2 * i] = x0[i] >= 0 && x0[i] < W;
    q.ok[2 * i + 1] = x0[i] + 1 >= 0 && x0[i] + 1 < W;
    interior = interior && q.ok[2 * i] && q.ok[2 * i + 1];
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
  return narrow ? WALK_NARROW : (interior ? WALK_WIDE : WALK_EDGE);
}

template <int IMM>
__device__ __forceinline__ void cp_async_4(unsigned smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0+%2], [%1+%2], 4;" ::"r"(smem_dst), "l"(gmem_src), "n"(IMM) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// A CTA owns `strip_rows` output rows of one frame.
// STAGED: the source window of the strip -- the rows and columns its taps touch, ~25 rows x 3 channels x 330 B for the gaze
//   crop -- is first copied into shared memory with asynchronous 4-byte copies (cp.async: a warp moves 128 contiguous bytes per
//   instruction, no registers, no conversion; TMA cannot be used because the 652-byte row pitch of the 326-pixel front frames is
//   not a multiple of 16), so the walk itself never waits for global memory.  If the window of the strip does not fit
//   WALK_STAGE_BYTES the strip is processed in chunks of rows.
// !STAGED: the same walk reading its taps from global memory (rows whose pitch or base is not 4-byte aligned, e.g. raw uint8
//   frames of 326 pixels).
template <typename TS, typename TD, bool STAGED>
__global__ void __launch_bounds__(WALK_THREADS, 8) fov_crop_walk_kernel(const Args a, const int strip_rows) {
  typedef typename RawPx<TS>::type Raw;
  extern __shared__ __align__(16) unsigned char walk_stage[];
  __shared__ WalkRow s_rows[WALK_MAX_ROWS];
  const int n = blockIdx.y, r0 = blockIdx.x * strip_rows, tid = threadIdx.x;
  const int S = a.S, H = a.H, W = a.W;
  const int nrows = min(strip_rows, S - r0);
  const float cx = __ldg(a.centers + 2 * n), cy = __ldg(a.centers + 2 * n + 1);
  const float fw = __ldg(a.windows + 2 * n), fh = __ldg(a.windows + 2 * n + 1);
  const long long src_frame = a.frame_ids ? __ldg(a.frame_ids + n) : n;
  const int plane = H * W;
  const Raw* src = reinterpret_cast<const Raw*>(a.frames) + src_frame * 3ll * plane;
  const float inv_s = 1.0f / S;
  // same expressions as the direct kernel / the oracle
  auto sample_x = [&](int ox) { return ((fw * ((2 * ox + 1) * inv_s - 1.0f) + (2.0f * cx - 1.0f) + 1.0f) * W - 1.0f) * 0.5f; };
  auto sample_y = [&](int oy) { return ((fh * ((2 * oy + 1) * inv_s - 1.0f) + (2.0f * cy - 1.0f) + 1.0f) * H - 1.0f) * 0.5f; };
  // positions far outside the frame all behave like "every tap out": clamped so that +1 / +2 cannot overflow
  auto floor_clamped = [](float v, int hi) { return static_cast<int>(fminf(fmaxf(floorf(v), -3.0f), static_cast<float>(hi))); };

  if (tid < nrows) {  // the strip's row table
    const int oy = r0 + tid;
    const float sy = sample_y(oy);
    const int y0 = floor_clamped(sy, H), yp = floor_clamped(sample_y(oy - 1), H);
    const int flags = (y0 != yp ? WALK_NEW : 0) | (y0 == yp + 1 ? WALK_SHIFT : 0) | (y0 >= 0 && y0 < H ? WALK_A_IN : 0) |
                      (y0 + 1 >= 0 && y0 + 1 < H ? WALK_B_IN : 0);
    WalkRow ri;
    ri.code = (y0 + 4) | (flags << 24);
    ri.wy1 = sy - floorf(sy);
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
    const float sx = sample_x(ox + i);
    x0[i] = floor_clamped(sx, W);
    w1[i] = sx - floorf(sx);
    w0[i] = 1.0f - w1[i];
    q.ok[2 * i] = x0[i] >= 0 && x0[i] < W;
    q.ok[2 * i + 1] = x0[i] + 1 >= 0 && x0[i] + 1 < W;
    interior = interior && q.ok[2 * i] && q.ok[2 * i + 1];
  }
  const int d = x0[1] - x0[0];
  const bool narrow = interior && (d == 0 || d == 1) && x0[0] + 2 < W;
  const int mode = narrow ? WALK_NARROW : (interior ? WALK_WIDE : WALK_EDGE);
  q.w[0] = w0[0]; q.w[1] = w1[0];
  if (narrow) {
    q.w[2] = d ? 0.0f : w0[1];
    q.w[3] = d ? w0[1] : w1[1];
    q.w[4] = d ? w1[1] : 0.0f;
  } else {
    q.w[2] = w0[1]; q.w[3] = w1[1]; q.w[4] = 0.0f;
  }

  TD* out = reinterpret_cast<TD*>(a.out);
  long long ch_stride;
  if (a.patch > 0) {
    const int px = fast_div(min(ox, S - 2), a.patch_magic), ix = ox - px * a.patch;
    out += (static_cast<long long>(n) * a.G * a.G + px) * a.out_ld + ix;
    ch_stride = static_cast<long long>(a.patch) * a.patch;
  } else {
    out += static_cast<long long>(n) * 3 * S * S + ox;
    ch_stride = static_cast<long long>(S) * S;
  }
  const bool worker = ox < S;

  if (!STAGED) {
    __syncthreads();
    if (!worker) return;
    WalkSrc<TS> ws;
    ws.base = src; ws.sbase = 0; ws.y_lo = 0; ws.row_stride = W; ws.ch_stride = plane;
    q.k[0] = x0[0]; q.k[1] = x0[1];
    walk_dispatch<TS, TD, false>(mode, a, s_rows, 0, nrows, ws, q, out, ch_stride);
    return;
  }

  // ---- staged: columns [j_lo, j_hi] of the frame that the strip's taps touch, first column on a 4-byte boundary ----
  constexpr int ES = sizeof(Raw), EPW = 4 / ES;  // elements per 4-byte word
  const int xf = floor_clamped(sample_x(0), W), xl = floor_clamped(sample_x(S - 1), W);
  const int j_lo = (max(min(xf, xl), 0) / EPW) * EPW;
  const int j_hi = min(max(xf, xl) + 2, W - 1);
  const int words = j_hi >= j_lo ? (j_hi - j_lo + EPW) / EPW : 0;  // 4-byte words per staged line
  const int pitch_b = 4 * words;                                    // bytes
  const int rows_fit = words > 0 ? WALK_STAGE_BYTES / (3 * pitch_b) : (1 << 20);
  q.k[0] = (x0[0] - j_lo) * ES; q.k[1] = (x0[1] - j_lo) * ES;
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const bool cp0 = lane < words, cp1 = lane + 32 < words, cp2 = lane + 64 < words, cp3 = lane + 96 < words;
  unsigned stage_addr = static_cast<unsigned>(__cvta_generic_to_shared(walk_stage));
  asm volatile("" : "+r"(stage_addr));  // keep it in a register (the compiler would re-derive it from %cluster_ctaid in the walk)
  __syncthreads();  // row table

  int ra = 0;
  while (ra < nrows) {  // chunks of output rows whose source rows fit the staging buffer (CTA-uniform; normally one)
    int lo = walk_y0(s_rows[ra].code), hi = lo, rb;
    {
      const int ye = walk_y0(s_rows[nrows - 1].code);  // sample rows are monotone in the row index: the extremes are at the ends
      const int l2 = min(lo, ye), h2 = max(hi, ye);
      if (h2 + 1 - l2 + 1 <= rows_fit) {
        lo = l2; hi = h2; rb = nrows;
      } else {
        rb = ra + 1;
        while (rb < nrows) {
          const int y = walk_y0(s_rows[rb].code);
          const int nl = min(lo, y), nh = max(hi, y);
          if (nh + 1 - nl + 1 > rows_fit) break;
          lo = nl; hi = nh; ++rb;
        }
      }
    }
    // staged rows [y_lo, y_lo + n_src): y0 is clamped to [-3, H], so at most a few of them lie outside the frame; those become
    // zero lines, and the walk needs no row test
    const int y_lo = lo, n_src = hi + 1 - lo + 1;
    if (words > 0) {
      // one (row, channel) line per warp pass; lane l copies words l, l+32, l+64, l+96 (predicates fixed for the whole kernel)
      for (int c = 0; c < 3; ++c) {
        const unsigned char* g = reinterpret_cast<const unsigned char*>(src + (c * plane + (y_lo + warp) * W + j_lo)) + 4 * lane;
        unsigned sa = stage_addr + static_cast<unsigned>(c * n_src + warp) * pitch_b + 4 * lane;
        for (int yr = warp; yr < n_src; yr += nwarps, g += nwarps * W * ES, sa += nwarps * pitch_b) {
          if (y_lo + yr >= 0 && y_lo + yr < H) {  // warp-uniform
            if (cp0) cp_async_4<0>(sa, g);
            if (cp1) cp_async_4<128>(sa, g);
            if (cp2) cp_async_4<256>(sa, g);
            if (cp3) cp_async_4<384>(sa, g);
            for (int wd = lane + 128; wd < words; wd += 32) cp_async_4<0>(sa + 4 * (wd - lane), g + 4 * (wd - lane));  // very wide windows
          } else {
            for (int wd = lane; wd < words; wd += 32) asm volatile("st.shared.u32 [%0], %1;" ::"r"(sa + 4 * (wd - lane)), "r"(0) : "memory");
          }
        }
      }
      cp_async_wait_all();
    }
    __syncthreads();
    if (worker) {
      WalkSrc<TS> ws;
      ws.base = nullptr; ws.sbase = stage_addr; ws.y_lo = y_lo; ws.row_stride = pitch_b; ws.ch_stride = n_src * pitch_b;
      walk_dispatch<TS, TD, true>(mode, a, s_rows, ra, rb, ws, q, out, ch_stride);
    }
    ra = rb;
    if (ra < nrows) __syncthreads();  // the staging buffer is reused by the next chunk
  }
}

// The staged walk with FOUR adjacent output columns per thread (default when the output is 16-byte aligned).  A CTA is two
// groups of `cols` threads; both share the staged window of the strip, each walks one half of its rows (warp-uniform control).
template <typename TS, typename TD>
__global__ void __launch_bounds__(WALK_THREADS, 6) fov_crop_walk4_kernel(const Args a, const int strip_rows, const int cols) {
  typedef typename RawPx<TS>::type Raw;
  extern __shared__ __align__(16) unsigned char walk_stage[];
  __shared__ WalkRow s_rows[WALK_MAX_ROWS];
  const int n = blockIdx.y, r0 = blockIdx.x * strip_rows, tid = threadIdx.x;
  const int S = a.S, H = a.H, W = a.W;
  const int nrows = min(strip_rows, S - r0);
  const float cx = __ldg(a.centers + 2 * n), cy = __ldg(a.centers + 2 * n + 1);
  const float fw = __ldg(a.windows + 2 * n), fh = __ldg(a.windows + 2 * n + 1);
  const long long src_frame = a.frame_ids ? __ldg(a.frame_ids + n) : n;
  const int plane = H * W;
  const Raw* src = reinterpret_cast<const Raw*>(a.frames) + src_frame * 3ll * plane;
  const float inv_s = 1.0f / S;
  auto sample_x = [&](int ox) { return ((fw * ((2 * ox + 1) * inv_s - 1.0f) + (2.0f * cx - 1.0f) + 1.0f) * W - 1.0f) * 0.5f; };
  auto sample_y = [&](int oy) { return ((fh * ((2 * oy + 1) * inv_s - 1.0f) + (2.0f * cy - 1.0f) + 1.0f) * H - 1.0f) * 0.5f; };
  auto floor_clamped = [](float v, int hi) { return static_cast<int>(fminf(fmaxf(floorf(v), -3.0f), static_cast<float>(hi))); };

  if (tid < nrows) {  // the strip's row table (identical to fov_crop_walk_kernel's)
    const int oy = r0 + tid;
    const float sy = sample_y(oy);
    const int y0 = floor_clamped(sy, H), yp = floor_clamped(sample_y(oy - 1), H);
    const int flags = (y0 != yp ? WALK_NEW : 0) | (y0 == yp + 1 ? WALK_SHIFT : 0) | (y0 >= 0 && y0 < H ? WALK_A_IN : 0) |
                      (y0 + 1 >= 0 && y0 + 1 < H ? WALK_B_IN : 0);
    WalkRow ri;
    ri.code = (y0 + 4) | (flags << 24);
    ri.wy1 = sy - floorf(sy);
    ri.wy0 = 1.0f - ri.wy1;
    if (a.patch > 0) {
      const int py = fast_div(oy, a.patch_magic), iy = oy - py * a.patch;
      ri.off = py * a.G * static_cast<int>(a.out_ld) + iy * a.patch;
    } else {
      ri.off = oy * S;
    }
    s_rows[tid] = ri;
  }

  const int group = tid >= cols ? 1 : 0;
  const int ox = 4 * (tid - group * cols);
  WalkCols qa, qb;
  int xa[2], xb[2];
  const int mode_a = walk_setup_pair(ox, W, sample_x, floor_clamped, qa, xa);
  const int mode_b = walk_setup_pair(ox + 2, W, sample_x, floor_clamped, qb, xb);

  TD* out = reinterpret_cast<TD*>(a.out);
  long long ch_stride;
  if (a.patch > 0) {
    const int px = fast_div(min(ox, S - 4), a.patch_magic), ix = ox - px * a.patch;
    out += (static_cast<long long>(n) * a.G * a.G + px) * a.out_ld + ix;
    ch_stride = static_cast<long long>(a.patch) * a.patch;
  } else {
    out += static_cast<long long>(n) * 3 * S * S + ox;
    ch_stride = static_cast<long long>(S) * S;
  }
  const bool worker = ox < S;

  constexpr int ES = sizeof(Raw), EPW = 4 / ES;
  const int xf = floor_clamped(sample_x(0), W), xl = floor_clamped(sample_x(S - 1), W);
  const int j_lo = (max(min(xf, xl), 0) / EPW) * EPW;
  const int j_hi = min(max(xf, xl) + 2, W - 1);
  const int words = j_hi >= j_lo ? (j_hi - j_lo + EPW) / EPW : 0;
  const int pitch_b = 4 * words;
  const int rows_fit = words > 0 ? WALK_STAGE_BYTES / (3 * pitch_b) : (1 << 20);
  qa.k[0] = (xa[0] - j_lo) * ES; qa.k[1] = (xa[1] - j_lo) * ES;
  qb.k[0] = (xb[0] - j_lo) * ES; qb.k[1] = (xb[1] - j_lo) * ES;
  const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const bool cp0 = lane < words, cp1 = lane + 32 < words, cp2 = lane + 64 < words, cp3 = lane + 96 < words;
  unsigned stage_addr = static_cast<unsigned>(__cvta_generic_to_shared(walk_stage));
  asm volatile("" : "+r"(stage_addr));
  __syncthreads();  // row table

  int ra = 0;
  while (ra < nrows) {  // chunks of output rows whose source rows fit the staging buffer (CTA-uniform; normally one)
    int lo = walk_y0(s_rows[ra].code), hi = lo, rb;
    {
      const int ye = walk_y0(s_rows[nrows - 1].code);
      const int l2 = min(lo, ye), h2 = max(hi, ye);
      if (h2 + 1 - l2 + 1 <= rows_fit) {
        lo = l2; hi = h2; rb = nrows;
      } else {
        rb = ra + 1;
        while (rb < nrows) {
          const int y = walk_y0(s_rows[rb].code);
          const int nl = min(lo, y), nh = max(hi, y);
          if (nh + 1 - nl + 1 > rows_fit) break;
          lo = nl; hi = nh; ++rb;
        }
      }
    }
    const int y_lo = lo, n_src = hi + 1 - lo + 1;
    if (words > 0) {
      for (int c = 0; c < 3; ++c) {
        const unsigned char* g = reinterpret_cast<const unsigned char*>(src + (c * plane + (y_lo + warp) * W + j_lo)) + 4 * lane;
        unsigned sa = stage_addr + static_cast<unsigned>(c * n_src + warp) * pitch_b + 4 * lane;
        for (int yr = warp; yr < n_src; yr += nwarps, g += nwarps * W * ES, sa += nwarps * pitch_b) {
          if (y_lo + yr >= 0 && y_lo + yr < H) {  // warp-uniform
            if (cp0) cp_async_4<0>(sa, g);
            if (cp1) cp_async_4<128>(sa, g);
            if (cp2) cp_async_4<256>(sa, g);
            if (cp3) cp_async_4<384>(sa, g);
            for (int wd = lane + 128; wd < words; wd += 32) cp_async_4<0>(sa + 4 * (wd - lane), g + 4 * (wd - lane));
          } else {
            for (int wd = lane; wd < words; wd += 32) asm volatile("st.shared.u32 [%0], %1;" ::"r"(sa + 4 * (wd - lane)), "r"(0) : "memory");
          }
        }
      }
      cp_async_wait_all();
    }
    __syncthreads();
    if (worker) {
      const int per = (rb - ra + 1) >> 1;  // rows of the chunk per group
      const int ga = ra + group * per, gb = min(ga + per, rb);
      if (ga < gb) {
        WalkSrc<TS> ws;
        ws.base = nullptr; ws.sbase = stage_addr; ws.y_lo = y_lo; ws.row_stride = pitch_b; ws.ch_stride = n_src * pitch_b;
        if (mode_a == WALK_NARROW && mode_b == WALK_NARROW) {
          walk_rows4<TS, TD>(a, s_rows, ga, gb, ws, qa, qb, out, ch_stride);
        } else {  // a pair at the frame border / a down-sampling window: the two-column walk, once per pair
          walk_dispatch<TS, TD, true>(mode_a, a, s_rows, ga, gb, ws, qa, out, ch_stride);
          walk_dispatch<TS, TD, true>(mode_b, a, s_rows, ga, gb, ws, qb, out + 2, ch_stride);
        }
      }
    }
    ra = rb;
    if (ra < nrows) __syncthreads();
  }
}

// RF_CROP_DIRECT=1 selects the round-1 direct gather, RF_CROP_NOSTAGE=1 the walker without shared-memory staging (read per
// call so tests and A/B runs can toggle them).
static bool env_on(const char* name) {
  const char* e = getenv(name);
  return e && e[0] == '1';
}

template <typename TS, typename TD>
static int launch_walk(const RfFovCropParams* p, const Args& a, cudaStream_t s) {
  const int S = p->out_size;
  const int strip = (S % 32 != 0 && S % 28 == 0) ? 28 : WALK_MAX_ROWS;
  const int threads = ((S / 2 + 31) / 32) * 32;
  const dim3 grid(ceil_div(S, strip), p->n_frames);
  const size_t es = sizeof(typename RawPx<TS>::type);
  // cp.async moves 4-byte words: frame base and row pitch must be 4-byte aligned, and two staged rows must fit
  const bool staged = !env_on("RF_CROP_NOSTAGE") && reinterpret_cast<uintptr_t>(p->frames) % 4 == 0 && (p->W * es) % 4 == 0 &&
                      6 * (p->W * es + 8) <= static_cast<size_t>(WALK_STAGE_BYTES);
  // four columns per thread: 16-byte aligned output rows (S and the patch size are multiples of 4 already)
  const bool quad = staged && !env_on("RF_CROP_PAIRS") && reinterpret_cast<uintptr_t>(p->out) % 16 == 0 &&
                    (p->patch == 0 || p->out_ld % 8 == 0);
  if (quad) {
    const int cols = ((S / 4 + 31) / 32) * 32;
    fov_crop_walk4_kernel<TS, TD><<<grid, 2 * cols, WALK_STAGE_BYTES, s>>>(a, strip, cols);
  } else if (staged) {
    RF_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(fov_crop_walk_kernel<TS, TD, true>), WALK_STAGE_BYTES));
    fov_crop_walk_kernel<TS, TD, true><<<grid, threads, WALK_STAGE_BYTES, s>>>(a, strip);
  } else {
    fov_crop_walk_kernel<TS, TD, false><<<grid, threads, 0, s>>>(a, strip);
  }
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
  const bool walk_forced = env_on("RF_CROP_WALK");
  if (!env_on("RF_CROP_DIRECT") && frame_elems < (1ll << 31) && 3ll * p->H * p->W + 4 < (1ll << 31) && (p->H <= 2 * p->out_size || walk_forced)) {
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
