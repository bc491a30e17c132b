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

// ---- strip kernel: shared-memory staged source window, separable blend, register reuse down the rows -------------
constexpr int TILE_THREADS = 256;
constexpr int STRIP_ROWS = 32;        // output rows per CTA
constexpr int STAGE_FLOATS = 15872;   // 62 KiB staging buffer (3 channels x rows x pitch floats)

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
  int vec2;    // 1: W even and the frame base 2-element aligned -> 2-element vector loads while staging
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

// The four staged floats [e, e+3] of one line, e even: two 8 B shared loads
struct Win4 { float f[4]; };
__device__ __forceinline__ Win4 load_win(const float* line, int e) {
  const float2 lo = *reinterpret_cast<const float2*>(line + e), hi = *reinterpret_cast<const float2*>(line + e + 2);
  Win4 w;
  w.f[0] = lo.x; w.f[1] = lo.y; w.f[2] = hi.x; w.f[3] = hi.y;
  return w;
}

// packed fp32 pairs (Blackwell FFMA2 / FMUL2: two IEEE fp32 operations per issue slot)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float2 upk(f32x2 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// the four staged floats [e, e+3] of one line (e even) as two packed pairs: two 8 B shared loads
struct Win { f32x2 lo, hi; };
__device__ __forceinline__ Win load_w(const float* line) {
  Win w;
  w.lo = *reinterpret_cast<const f32x2*>(line);
  w.hi = *reinterpret_cast<const f32x2*>(line + 2);
  return w;
}

struct RowInfo {  // per output row of a chunk, written once per CTA
  int ia, ib;      // float offsets of the staged lines of source rows y0 / y0+1 (channel 0)
  float wy0, wy1;  // their weights (0 outside the frame)
  long long o;     // output offset of (row, this thread's columns excluded, channel 0)
  int y0, pad;
};

__device__ __forceinline__ void store1(float* dst, float a) { *dst = a; }
__device__ __forceinline__ void store1(__half* dst, float a) { *dst = __float2half_rn(a); }
__device__ __forceinline__ void store1(__nv_bfloat16* dst, float a) { *dst = __float2bfloat16_rn(a); }

// Rows [my_a, my_b) of the chunk for one thread.  PAIR: the thread's two adjacent output pixels read the same 4-float window
// (<= 1 source column per output column) and leave as one packed store; otherwise the routine is called once per pixel.
// The source rows of the previous output row stay in registers; all control flow is CTA-uniform.
template <bool PAIR, typename TD>
__device__ __forceinline__ void strip_rows(const float* __restrict__ stage, const RowInfo* __restrict__ rows, int my_a, int my_b, int pitch,
                                           int e0, const float (&wa)[4], const float (&wb)[4], long long col_off, long long ch_stride,
                                           const float (&mean)[3], const float (&inv_std)[3], TD* __restrict__ out) {
  Win A[3], B[3];
  const f32x2 walo = pk(wa[0], wa[1]), wahi = pk(wa[2], wa[3]);
  const f32x2 wblo = pk(wb[0], wb[1]), wbhi = pk(wb[2], wb[3]);
  int cur_y = -(1 << 30);
  for (int r = my_a; r < my_b; ++r) {
    const RowInfo ri = rows[r];
    if (ri.y0 != cur_y) {
      const bool shift = ri.y0 == cur_y + 1;
      const float* la = stage + ri.ia + e0;
      const float* lb = stage + ri.ib + e0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (shift) A[c] = B[c];
        else A[c] = load_w(la + c * pitch);
        B[c] = load_w(lb + c * pitch);
      }
      cur_y = ri.y0;
    }
    const f32x2 wy0 = pk(ri.wy0, ri.wy0), wy1 = pk(ri.wy1, ri.wy1);
    TD* dst = out + ri.o + col_off;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // vertical blend of the window, then the horizontal 4-tap dot product(s)
      const f32x2 tlo = fma2(B[c].lo, wy1, mul2(A[c].lo, wy0)), thi = fma2(B[c].hi, wy1, mul2(A[c].hi, wy0));
      const float2 p0 = upk(fma2(thi, wahi, mul2(tlo, walo)));
      if (PAIR) {
        const float2 p1 = upk(fma2(thi, wbhi, mul2(tlo, wblo)));
        store2(dst + c * ch_stride, ((p0.x + p0.y) - mean[c]) * inv_std[c], ((p1.x + p1.y) - mean[c]) * inv_std[c]);
      } else {
        store1(dst + c * ch_stride, ((p0.x + p0.y) - mean[c]) * inv_std[c]);
      }
    }
  }
}

// A CTA owns STRIP_ROWS output rows of one frame and works through them in chunks whose source rows fit the staging buffer:
//   stage  : the source rows x columns the chunk touches, converted to fp32, loaded with independent coalesced vector loads
//            (every thread issues its whole share before the first use: the pass is bandwidth-, not latency-bound);
//   compute: a thread owns two adjacent output columns and walks DOWN its rows; the column taps are a 4-float window of the
//            staged line with a fixed 4-vector of weights (no indexing in the loop), the two source rows of the previous
//            output row stay in registers (an up-sampling crop needs a new source row for ~2 of 3 output rows), the vertical
//            blend is shared by the two pixels.  ~10 instructions per output element against ~35 for the direct gather.
template <typename TS, typename TD>
__global__ void __launch_bounds__(TILE_THREADS, 2) fov_crop_tiled_kernel(const TiledArgs ta) {
  extern __shared__ __align__(16) float stage[];  // [rows][3][pitch]
  __shared__ RowInfo s_rows[STRIP_ROWS];
  const Args& a = ta.a;
  const int n = blockIdx.y;
  const int r0 = blockIdx.x * STRIP_ROWS;
  const int tid = threadIdx.x;
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

  const int pairs = S >> 1;
  const int groups = TILE_THREADS / pairs;            // row groups that fit the CTA (S = 256 -> 2, 224 -> 2, 64 -> 8)
  const int pg = tid / pairs, pp = tid - pg * pairs;  // one integer division per thread
  const int ox = pp << 1;
  const bool worker = pg < groups;
  const bool mirrored = !(fw > 0.0f) || !(fh > 0.0f);  // CTA-uniform; sample positions then do not grow with the index
  const int r_end = min(r0 + STRIP_ROWS, S);

  // output addressing
  long long obase, row_stride, ch_stride;
  int patch_rows_left = 1 << 30;  // patch-major: rows until the strip crosses into the next row of patches
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
  auto out_offset = [&](int r) {  // r = row inside the strip
    long long o = obase + r * row_stride;
    int left = patch_rows_left, rr = r;
    while (rr >= left) {  // next row(s) of patches: + G patches, back to the first row of that patch
      o += static_cast<long long>(a.G) * a.out_ld - static_cast<long long>(a.patch) * a.patch;
      rr -= a.patch;
    }
    return o;
  };

  if (mirrored) {  // direct gather, any orientation
    if (worker)
      for (int r = r0 + pg; r < r_end; r += groups) {
        const float sy = sample_y(r);
        const long long o = out_offset(r - r0);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float v0 = direct_sample<TS>(src + c * plane, sample_x(ox), sy, H, W);
          const float v1 = direct_sample<TS>(src + c * plane, sample_x(ox + 1), sy, H, W);
          store2(out + o + c * ch_stride, (v0 - a.mean[c]) * a.inv_std[c], (v1 - a.mean[c]) * a.inv_std[c]);
        }
      }
    return;
  }

  // Source-column span the window covers, clipped to the frame, first column rounded down to an even one.  sx grows with ox,
  // so every in-frame tap of every output column lies in [j_lo, j_hi].
  const int j_lo = max(static_cast<int>(floorf(sample_x(0))), 0) & ~1;
  const int j_hi = min(static_cast<int>(floorf(sample_x(S - 1))) + 1, W - 1);
  const int span = j_hi - j_lo + 1;                       // <= 0: the window misses the frame horizontally
  const int pitch = ((max(span, 0) + 1) & ~1) + 4;        // even, >= span + 4: the 4-float windows never leave the line
  const int max_rows = max(span > 0 ? STAGE_FLOATS / (3 * pitch) : 0, 0);

  // Column taps of the two pixels: window start e (even) and a 4-vector of weights over slots e..e+3 of the staged line.
  // narrow (<= 1 source column per output column): both pixels share the window of pixel 0.
  const bool narrow = fw * W <= 0.999f * static_cast<float>(S);  // (margin: the two floors must never differ by 2)
  int e[2];
  float wgt[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float sx = sample_x(ox + i);
    const float fx0 = floorf(sx);
    const int x0 = static_cast<int>(fx0);
    const float w1 = sx - fx0, w0 = 1.0f - w1;
    const int rel = x0 - j_lo;
    const int start = (i == 1 && narrow) ? e[0] : min(max(rel, 0) & ~1, max(pitch - 4, 0));
    e[i] = start;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int col = j_lo + start + k;
      float w = 0.0f;
      if (col >= 0 && col < W) {
        if (col == x0) w = w0;
        else if (col == x0 + 1) w = w1;
      }
      wgt[i][k] = w;
    }
  }

  int ra = r0;
  while (ra < r_end) {  // chunks of rows whose source rows fit the staging buffer (CTA-uniform control flow)
    // rows of the chunk: [ra, rb); source rows [y_lo, y_hi] clipped to the frame
    const int ya0 = static_cast<int>(floorf(sample_y(ra)));
    int rb = ra + 1;
    int y_last = ya0;
    if (max_rows >= 2) {
      // sy grows with the row: extend the chunk while its last source row still fits
      while (rb < r_end) {
        const int yn = static_cast<int>(floorf(sample_y(rb)));
        if (min(yn + 1, H - 1) - max(ya0, 0) + 1 > max_rows) break;
        y_last = yn;
        ++rb;
      }
    }
    const int y_lo = min(max(ya0, 0), H - 1), y_hi = min(max(y_last + 1, 0), H - 1);
    const bool rows_out = (y_last + 1 < 0) || (ya0 >= H);  // every tap row of the chunk is outside the frame
    const bool constant = rows_out || span <= 0 || max_rows < 2;
    const bool fits = max_rows >= 2 || rows_out || span <= 0;
    const int n_rows = y_hi - y_lo + 1;

    if (!constant) {
      // ---- stage: n_rows x 3 lines of `pitch` floats (tail zero-filled) --------------------------------------
      const int npairs = pitch >> 1;
      const int lane = tid & 31, warp = tid >> 5;
      for (int line_id = warp; line_id < n_rows * 3; line_id += TILE_THREADS / 32) {  // one (row, channel) line per warp pass
        const int yr = line_id / 3, c = line_id - yr * 3;
        const TS* row = src + c * plane + static_cast<long long>(y_lo + yr) * W + j_lo;
        float* line = stage + line_id * pitch;
#pragma unroll 4
        for (int jp = lane; jp < npairs; jp += 32) {  // independent loads: the compiler batches four of them per lane
          const int j = jp << 1;
          float2 v = make_float2(0.0f, 0.0f);
          if (j + 1 < span) v = ta.vec2 ? load_px2<TS>(row + j) : make_float2(load_px<TS>(row + j), load_px<TS>(row + j + 1));
          else if (j < span) v.x = load_px<TS>(row + j);
          *reinterpret_cast<float2*>(line + j) = v;
        }
      }
    }
    if (!constant && tid < rb - ra) {  // per-row constants of the chunk, once per CTA
      const int r = ra + tid;
      const float sy = sample_y(r);
      const float fy0 = floorf(sy);
      const int y0 = static_cast<int>(fy0);
      const float w1 = sy - fy0;
      RowInfo ri;
      ri.y0 = y0;
      ri.pad = 0;
      ri.wy0 = (y0 >= 0 && y0 < H) ? 1.0f - w1 : 0.0f;
      ri.wy1 = (y0 + 1 >= 0 && y0 + 1 < H) ? w1 : 0.0f;
      ri.ia = (min(max(y0, y_lo), y_hi) - y_lo) * 3 * pitch;
      ri.ib = (min(max(y0 + 1, y_lo), y_hi) - y_lo) * 3 * pitch;
      ri.o = out_offset(r - r0) - obase;
      s_rows[tid] = ri;
    }
    __syncthreads();

    if (worker) {
      const int per = (rb - ra + groups - 1) / groups;
      const int my_a = ra + pg * per, my_b = min(my_a + per, rb);
      if (!fits) {  // frame too wide for even two staged rows: direct gather for this chunk
        for (int r = my_a; r < my_b; ++r) {
          const float sy = sample_y(r);
          const long long o = out_offset(r - r0);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float v0 = direct_sample<TS>(src + c * plane, sample_x(ox), sy, H, W);
            const float v1 = direct_sample<TS>(src + c * plane, sample_x(ox + 1), sy, H, W);
            store2(out + o + c * ch_stride, (v0 - a.mean[c]) * a.inv_std[c], (v1 - a.mean[c]) * a.inv_std[c]);
          }
        }
      } else if (constant) {
        for (int r = my_a; r < my_b; ++r) {
          const long long o = out_offset(r - r0);
#pragma unroll
          for (int c = 0; c < 3; ++c) store2(out + o + c * ch_stride, (0.0f - a.mean[c]) * a.inv_std[c], (0.0f - a.mean[c]) * a.inv_std[c]);
        }
      } else {
        // ---- compute: walk down the rows, the two source rows of the previous output row stay in registers ----
        const long long col_off = obase;  // (row / patch-row part comes from RowInfo.o, measured from the strip's first row)
        if (narrow) {
          strip_rows<true, TD>(stage, s_rows, my_a - ra, my_b - ra, pitch, e[0], wgt[0], wgt[1], col_off, ch_stride, a.mean, a.inv_std, out);
        } else {  // > 1 source column per output column: each pixel has its own window; one pass per pixel, scalar stores
          strip_rows<false, TD>(stage, s_rows, my_a - ra, my_b - ra, pitch, e[0], wgt[0], wgt[0], col_off, ch_stride, a.mean, a.inv_std, out);
          strip_rows<false, TD>(stage, s_rows, my_a - ra, my_b - ra, pitch, e[1], wgt[1], wgt[1], col_off + 1, ch_stride, a.mean, a.inv_std, out);
        }
      }
    }
    __syncthreads();  // the staging buffer is reused by the next chunk
    ra = rb;
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
  ta.vec2 = (p->W % 2 == 0) && (reinterpret_cast<uintptr_t>(p->frames) % (2 * sizeof(TS)) == 0);
  const size_t smem = sizeof(float) * STAGE_FLOATS;
  RF_CUDA_OK(ensure_dynamic_smem(reinterpret_cast<const void*>(fov_crop_tiled_kernel<TS, TD>), smem));
  fov_crop_tiled_kernel<TS, TD><<<dim3(ceil_div(p->out_size, STRIP_ROWS), p->n_frames), TILE_THREADS, smem, s>>>(ta);
  RF_LAUNCH_OK();
  return RF_OK;
}

template <typename TS>
static int dispatch_out(const RfFovCropParams* p, const Args& a, cudaStream_t s) {
  // strip kernel: 2 px per thread -> even S, at most 256 threads per output row
  if (tiled_enabled() && p->out_size >= 8 && p->out_size <= 2 * TILE_THREADS) {
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
