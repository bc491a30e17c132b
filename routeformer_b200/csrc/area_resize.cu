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
  dim3 grid(ceil_div(p->dW, area::AREA_THREADS), ceil_div(p->dH, area::AREA_ROWS), p->n_planes);
  area::area_resize_kernel<<<grid, area::AREA_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a);
  RF_LAUNCH_OK();
  return RF_OK;
}
