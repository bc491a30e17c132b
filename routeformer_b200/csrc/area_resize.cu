// Dataset-side video down-scaling on the device.  See include/routeformer_b200.h (10).
//
// Bit-exact re-implementation, for uint8 planes, of what the reference's loader does on the host with
//   cv2.resize(frame, (int(W * f), int(H * f)), interpolation=cv2.INTER_AREA)          (io/dataset.py:1440-1501)
// i.e. of OpenCV's resizeAreaFast_ / resizeArea_ (modules/imgproc/src/resize.cpp), so that raw uint8 frames can be uploaded once
// at camera resolution and scaled where they are consumed:
//   scale = 1 / (dsize / ssize) per axis, in double;
//   * both scales integral: int32 box sums;  2 x 2 -> (s + 2) >> 2;  otherwise cvRound(float(s) * (1.f / area));
//   * otherwise: float32 weighted sums WITHOUT fused multiply-add, in OpenCV's order -- per source row the horizontal cell sum
//     buf = sum_k S[sx_k] * alpha_k (left partial cell, full cells, right partial cell), then down the rows of the destination
//     cell acc = beta * buf (first row) / acc + beta * buf, D = saturate_u8(cvRound(acc)); cell tables as computeResizeAreaTab
//     builds them (double arithmetic, weights stored as float).
// HBM-bound byte work (each source byte is read once, 1 / scale^2 bytes are written): a thread owns one destination column of
// one plane and walks down AREA_ROWS destination rows; its horizontal cell (first column, weights) is computed once, in double,
// exactly as the host code does; the vertical cells of the strip are computed by the first threads of the CTA into shared
// memory.  Adjacent threads read adjacent cells: a warp covers 32 x scale contiguous source bytes per row.
#include <cstdlib>

#include "common.cuh"

namespace rf {
namespace area {

constexpr int AREA_THREADS = 128;
constexpr int AREA_ROWS = 4;  // destination rows per CTA

// One destination cell along one axis (computeResizeAreaTab for a single dx): source index range and the three weights.
struct Cell {
  int first;            // first source index that contributes
  int n_full;           // number of full cells after the optional left partial cell
  float a_left, a_full, a_right;  // weights; a_left / a_right = 0 with has_left / has_right = 0
  int has_left, has_right;
};

__device__ __forceinline__ Cell make_cell(int d, int ssize, double scale) {
  const double f1 = d * scale;
  const double f2 = f1 + scale;
  const double cell = fmin(scale, ssize - f1);
  int s1 = static_cast<int>(ceil(f1)), s2 = static_cast<int>(floor(f2));
  s2 = min(s2, ssize - 1);
  s1 = min(s1, s2);
  Cell c;
  c.has_left = (s1 - f1 > 1e-3) ? 1 : 0;
  c.has_right = (f2 - s2 > 1e-3) ? 1 : 0;
  c.a_left = c.has_left ? static_cast<float>((s1 - f1) / cell) : 0.0f;
  c.a_full = static_cast<float>(1.0 / cell);
  c.a_right = c.has_right ? static_cast<float>(fmin(fmin(f2 - s2, 1.0), cell) / cell) : 0.0f;
  c.first = c.has_left ? s1 - 1 : s1;
  c.n_full = s2 - s1;
  return c;
}

struct Args {
  const unsigned char* src; long long plane_stride, row_pitch;
  int n_planes, H, W;
  unsigned char* dst; int dH, dW;
  double scale_x, scale_y;
  int iscale_x, iscale_y, fast;  // fast: both scales integral (box sums)
};

__device__ __forceinline__ unsigned char saturate_u8(int v) { return static_cast<unsigned char>(min(max(v, 0), 255)); }

// buf = sum_k S[sx_k] * alpha_k in OpenCV's order, float32 multiply then float32 add (never fused)
__device__ __forceinline__ float row_cell_sum(const unsigned char* __restrict__ S, const Cell& c) {
  float buf = 0.0f;
  const unsigned char* p = S + c.first;
  if (c.has_left) buf = __fadd_rn(buf, __fmul_rn(static_cast<float>(__ldg(p++)), c.a_left));
  int k = 0;
  for (; k + 4 <= c.n_full; k += 4) {  // four loads in flight, accumulated in order
    const float u0 = __ldg(p + k), u1 = __ldg(p + k + 1), u2 = __ldg(p + k + 2), u3 = __ldg(p + k + 3);
    buf = __fadd_rn(buf, __fmul_rn(u0, c.a_full));
    buf = __fadd_rn(buf, __fmul_rn(u1, c.a_full));
    buf = __fadd_rn(buf, __fmul_rn(u2, c.a_full));
    buf = __fadd_rn(buf, __fmul_rn(u3, c.a_full));
  }
  for (; k < c.n_full; ++k) buf = __fadd_rn(buf, __fmul_rn(static_cast<float>(__ldg(p + k)), c.a_full));
  if (c.has_right) buf = __fadd_rn(buf, __fmul_rn(static_cast<float>(__ldg(p + c.n_full)), c.a_right));
  return buf;
}

__global__ void __launch_bounds__(AREA_THREADS) area_resize_kernel(const Args a) {
  __shared__ Cell s_rows[AREA_ROWS];
  const int dx = blockIdx.x * AREA_THREADS + threadIdx.x;
  const int dy0 = blockIdx.y * AREA_ROWS;
  const int plane = blockIdx.z;
  const int nrows = min(AREA_ROWS, a.dH - dy0);
  const unsigned char* src = a.src + plane * a.plane_stride;
  unsigned char* dst = a.dst + (static_cast<long long>(plane) * a.dH + dy0) * a.dW;
  if (a.fast) {
    if (dx >= a.dW) return;
    const int area = a.iscale_x * a.iscale_y;
    const float inv_area = __fdiv_rn(1.0f, static_cast<float>(area));
    for (int r = 0; r < nrows; ++r) {
      const unsigned char* S = src + static_cast<long long>(dy0 + r) * a.iscale_y * a.row_pitch + dx * a.iscale_x;
      int sum = 0;
      for (int sy = 0; sy < a.iscale_y; ++sy, S += a.row_pitch)
        for (int sx = 0; sx < a.iscale_x; ++sx) sum += __ldg(S + sx);
      const int v = (a.iscale_x == 2 && a.iscale_y == 2) ? ((sum + 2) >> 2) : __float2int_rn(__fmul_rn(static_cast<float>(sum), inv_area));
      dst[r * a.dW + dx] = saturate_u8(v);
    }
    return;
  }
  if (threadIdx.x < nrows) s_rows[threadIdx.x] = make_cell(dy0 + threadIdx.x, a.H, a.scale_y);
  __syncthreads();
  if (dx >= a.dW) return;
  const Cell cx = make_cell(dx, a.W, a.scale_x);
  for (int r = 0; r < nrows; ++r) {
    const Cell cy = s_rows[r];
    const unsigned char* S = src + static_cast<long long>(cy.first) * a.row_pitch;
    float acc = 0.0f;
    const int n_src = cy.has_left + cy.n_full + cy.has_right;
    for (int j = 0; j < n_src; ++j, S += a.row_pitch) {
      const float beta = (cy.has_left && j == 0) ? cy.a_left : ((j < cy.has_left + cy.n_full) ? cy.a_full : cy.a_right);
      const float buf = row_cell_sum(S, cx);
      // first row of the cell: acc = beta * buf (0 + x == x for the first destination row of an OpenCV stripe as well)
      acc = j == 0 ? __fmul_rn(beta, buf) : __fadd_rn(acc, __fmul_rn(beta, buf));
    }
    dst[r * a.dW + dx] = saturate_u8(__float2int_rn(acc));
  }
}


// ---- word-streamed variant (4-byte aligned planes): loads first, arithmetic after --------------------------------------
// The byte-loop kernel above is latency-bound: a thread issues one dependent 1-byte load per multiply-add (10 % issue
// utilisation, 0.1-0.26 of the HBM roofline).  Here a thread loads, for RB source rows at a time, the NW aligned 32-bit words that
// cover its horizontal cell (<= MAXB bytes), shifts them so that cell byte k sits at a compile-time position (funnel shifts by
// the thread-constant misalignment), and only then runs OpenCV's sum -- same values, same order: byte k is multiplied by the
// thread's weight wk[k] (left partial, full cells, right partial, then zeros: buf + 0 * v == buf exactly for non-negative terms).
template <int MAXB>
struct CellWeights {
  float wk[MAXB];
  int first_word, shift_bits;  // first aligned word of the cell (index into the row), misalignment in bits
};

template <int MAXB>
__device__ __forceinline__ CellWeights<MAXB> make_weights(const Cell& c) {
  CellWeights<MAXB> w;
  const int cnt = c.has_left + c.n_full + c.has_right;
#pragma unroll
  for (int k = 0; k < MAXB; ++k) {
    float v = 0.0f;
    if (k < cnt) v = (k == 0 && c.has_left) ? c.a_left : ((k == cnt - 1 && c.has_right) ? c.a_right : c.a_full);
    w.wk[k] = v;
  }
  w.first_word = c.first >> 2;
  w.shift_bits = (c.first & 3) * 8;
  return w;
}

template <int MAXB, int RB, bool FAST>
__global__ void __launch_bounds__(AREA_THREADS) area_resize_words_kernel(const Args a) {
  constexpr int NW = (MAXB + 3 + 3) / 4;  // aligned words that cover MAXB bytes at any misalignment
  constexpr int WALK = 16;                // destination rows per CTA
  __shared__ Cell s_rows[WALK];
  const int dx = blockIdx.x * AREA_THREADS + threadIdx.x;
  const int dy0 = blockIdx.y * WALK;
  const int plane = blockIdx.z;
  const int nrows = min(WALK, a.dH - dy0);
  const unsigned char* src = a.src + plane * a.plane_stride;
  unsigned char* dst = a.dst + (static_cast<long long>(plane) * a.dH + dy0) * a.dW;
  if (!FAST) {
    if (threadIdx.x < nrows) s_rows[threadIdx.x] = make_cell(dy0 + threadIdx.x, a.H, a.scale_y);
    __syncthreads();
  }
  if (dx >= a.dW) return;
  Cell cx;
  if (FAST) {  // integral scale: the cell is exactly iscale_x bytes, all counted once
    cx.first = dx * a.iscale_x; cx.has_left = 0; cx.has_right = 0; cx.n_full = a.iscale_x; cx.a_left = cx.a_right = 0.0f; cx.a_full = 1.0f;
  } else {
    cx = make_cell(dx, a.W, a.scale_x);
  }
  const CellWeights<MAXB> wx = make_weights<MAXB>(cx);
  unsigned masks[NW];  // FAST: bytes of the (shifted) words that belong to the cell
  if (FAST) {
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const int n = min(max(a.iscale_x - 4 * i, 0), 4);
      masks[i] = n >= 4 ? 0xffffffffu : ((1u << (8 * n)) - 1u);
    }
  }
  const int area = a.iscale_x * a.iscale_y;
  const float inv_area = __fdiv_rn(1.0f, static_cast<float>(area));
  const long long pitch_w = a.row_pitch >> 2;
  const unsigned* base = reinterpret_cast<const unsigned*>(src) + wx.first_word;
  const int wlimit = (a.W >> 2) - wx.first_word;  // words of this row from the cell's first word on (never read past the row)
  // the words of one source row, shifted so that cell byte k is byte (k & 3) of al[k >> 2]
  auto load_row = [&](const unsigned* row, bool ok, unsigned (&al)[NW]) {
    unsigned w[NW + 1];
#pragma unroll
    for (int i = 0; i < NW; ++i) w[i] = (ok && i < wlimit) ? __ldg(row + i) : 0u;
    w[NW] = 0u;
#pragma unroll
    for (int i = 0; i < NW; ++i) al[i] = __funnelshift_r(w[i], w[i + 1], wx.shift_bits);
  };
  // A source row that is the partial LAST row of one destination cell is the partial FIRST row of the next one (the cells tile
  // the axis): its horizontal cell sum is kept instead of being loaded and summed again.
  float buf_last = 0.0f;
  int row_last = -1;
  for (int r = 0; r < nrows; ++r) {
    int first_row, n_src;
    Cell cy;
    if (FAST) {
      first_row = (dy0 + r) * a.iscale_y;
      n_src = a.iscale_y;
    } else {
      cy = s_rows[r];
      first_row = cy.first;
      n_src = cy.has_left + cy.n_full + cy.has_right;
    }
    const unsigned* row = base + static_cast<long long>(first_row) * pitch_w;
    float acc = 0.0f;
    int isum = 0;
    int j_begin = 0;
    if (!FAST && cy.has_left && first_row == row_last) {  // CTA-uniform
      acc = __fmul_rn(cy.a_left, buf_last);
      j_begin = 1;
    }
    for (int j0 = j_begin; j0 < n_src; j0 += RB) {
      unsigned al[RB][NW];
#pragma unroll
      for (int jj = 0; jj < RB; ++jj) load_row(row + static_cast<long long>(j0 + jj) * pitch_w, j0 + jj < n_src, al[jj]);
#pragma unroll
      for (int jj = 0; jj < RB; ++jj) {
        const int j = j0 + jj;
        if (j < n_src) {  // CTA-uniform per destination row
          if (FAST) {
#pragma unroll
            for (int i = 0; i < NW; ++i) isum = __dp4a(al[jj][i] & masks[i], 0x01010101u, static_cast<unsigned>(isum));
          } else {
            float buf = 0.0f;
#pragma unroll
            for (int k = 0; k < MAXB; ++k) {
              const float v = static_cast<float>((al[jj][k >> 2] >> (8 * (k & 3))) & 0xffu);
              buf = __fadd_rn(buf, __fmul_rn(v, wx.wk[k]));
            }
            const float beta = (cy.has_left && j == 0) ? cy.a_left : ((j < cy.has_left + cy.n_full) ? cy.a_full : cy.a_right);
            acc = j == 0 ? __fmul_rn(beta, buf) : __fadd_rn(acc, __fmul_rn(beta, buf));
            buf_last = buf;  // (after the loop: the sum of the cell's last source row)
          }
        }
      }
    }
    row_last = first_row + n_src - 1;
    int v;
    if (FAST) v = (a.iscale_x == 2 && a.iscale_y == 2) ? ((isum + 2) >> 2) : __float2int_rn(__fmul_rn(static_cast<float>(isum), inv_area));
    else v = __float2int_rn(acc);
    dst[r * a.dW + dx] = saturate_u8(v);
  }
}

template <int MAXB, int RB>
static void launch_words(const Args& a, cudaStream_t s) {
  dim3 grid(ceil_div(a.dW, AREA_THREADS), ceil_div(a.dH, 16), a.n_planes);
  if (a.fast) area_resize_words_kernel<MAXB, RB, true><<<grid, AREA_THREADS, 0, s>>>(a);
  else area_resize_words_kernel<MAXB, RB, false><<<grid, AREA_THREADS, 0, s>>>(a);
}

}  // namespace area
}  // namespace rf

extern "C" int rf_area_resize_u8(const RfAreaResizeParams* p, void* stream) {
  using namespace rf;
  RF_CHECK_ARG(p && p->src && p->dst, "rf_area_resize_u8: null pointer");
  RF_CHECK_ARG(p->n_planes > 0 && p->H > 0 && p->W > 0 && p->dH > 0 && p->dW > 0, "rf_area_resize_u8: empty problem");
  RF_CHECK_ARG(p->dH <= p->H && p->dW <= p->W, "rf_area_resize_u8: INTER_AREA path is down-scaling only (%dx%d -> %dx%d)", p->H, p->W, p->dH, p->dW);
  RF_CHECK_ARG(p->src_row_pitch >= p->W && p->src_plane_stride >= static_cast<long long>(p->H - 1) * p->src_row_pitch + p->W,
               "rf_area_resize_u8: source pitch / plane stride smaller than the plane");
  RF_CHECK_ARG(p->n_planes <= 65535 && ceil_div(p->dH, area::AREA_ROWS) <= 65535, "rf_area_resize_u8: at most 65535 planes / row strips per call");
  area::Args a;
  a.src = p->src; a.plane_stride = p->src_plane_stride; a.row_pitch = p->src_row_pitch;
  a.n_planes = p->n_planes; a.H = p->H; a.W = p->W; a.dst = p->dst; a.dH = p->dH; a.dW = p->dW;
  // hal::resize: scale = 1 / inv_scale with inv_scale = dsize / ssize (two roundings, as in OpenCV)
  a.scale_x = 1.0 / (static_cast<double>(p->dW) / static_cast<double>(p->W));
  a.scale_y = 1.0 / (static_cast<double>(p->dH) / static_cast<double>(p->H));
  a.iscale_x = static_cast<int>(nearbyint(a.scale_x));
  a.iscale_y = static_cast<int>(nearbyint(a.scale_y));
  const double eps = 2.220446049250313e-16;
  a.fast = (fabs(a.scale_x - a.iscale_x) < eps && fabs(a.scale_y - a.iscale_y) < eps) ? 1 : 0;
  if (a.fast) RF_CHECK_ARG(p->dW * a.iscale_x == p->W && p->dH * a.iscale_y == p->H, "rf_area_resize_u8: integral scale with a remainder");
  // widest horizontal cell: floor(scale) + 2 bytes (one partial cell on either side); exactly `scale` bytes when scale_x is integral
  const bool x_integral = fabs(a.scale_x - a.iscale_x) < eps && p->dW * a.iscale_x == p->W;
  const int max_bytes = (a.fast || x_integral) ? a.iscale_x : static_cast<int>(a.scale_x) + 2;
  const bool aligned = ((reinterpret_cast<uintptr_t>(p->src) | static_cast<uintptr_t>(p->src_plane_stride) | static_cast<uintptr_t>(p->src_row_pitch) |
                         static_cast<uintptr_t>(p->W)) & 3) == 0;
  static const bool words = [] { const char* e = getenv("RF_AREA_WORDS"); return !(e && e[0] == '0'); }();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (words && aligned && max_bytes <= 13 && ceil_div(p->dH, 16) <= 65535) {
    if (max_bytes <= 4) area::launch_words<4, 4>(a, st);
    else if (max_bytes <= 5) area::launch_words<5, 4>(a, st);
    else if (max_bytes <= 9) area::launch_words<9, 4>(a, st);
    else if (max_bytes <= 10) area::launch_words<10, 4>(a, st);
    else area::launch_words<13, 4>(a, st);
  } else {
    dim3 grid(ceil_div(p->dW, area::AREA_THREADS), ceil_div(p->dH, area::AREA_ROWS), p->n_planes);
    area::area_resize_kernel<<<grid, area::AREA_THREADS, 0, st>>>(a);
  }
  RF_LAUNCH_OK();
  return RF_OK;
}
