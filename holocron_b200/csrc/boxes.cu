// Pairwise box operators (xyxy boxes, fp32 math): IoU, GIoU, DIoU penalty, DIoU/CIoU loss, aspect-ratio
// consistency, and the analytic backward of the IoU-based ones.
// Reference: holocron/ops/boxes.py:16-211 (+ torchvision.ops.boxes.box_iou). One launch per operator instead of
// ~30 tiny ATen kernels and 2 MxNx2 temporaries. The forward uses explicit round-to-nearest intrinsics in the
// reference's operation order (no FMA contraction) so the exact-value vectors of the reference's own tests
// (tests/test_ops.py:25-76) hold bit for bit.
//
// NB (reference quirk, reproduced): ciou_loss adds its alpha*v term to a masked COPY (boxes.py:209), so it
// returns exactly the DIoU loss.
#include "common.cuh"

namespace {

enum Mode { M_IOU = 0, M_GIOU = 1, M_PENALTY = 2, M_DIOU_LOSS = 3, M_ARC = 4 };

struct Box { float x1, y1, x2, y2; };

__device__ __forceinline__ Box load_box(const float* p) { return Box{p[0], p[1], p[2], p[3]}; }

__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float dvd(float a, float b) { return __fdiv_rn(a, b); }

__device__ __forceinline__ void inter_union(const Box& a, const Box& b, float& inter, float& uni) {
  const float area_a = mul(sub(a.x2, a.x1), sub(a.y2, a.y1));
  const float area_b = mul(sub(b.x2, b.x1), sub(b.y2, b.y1));
  const float w = fmaxf(sub(fminf(a.x2, b.x2), fmaxf(a.x1, b.x1)), 0.f);
  const float h = fmaxf(sub(fminf(a.y2, b.y2), fmaxf(a.y1, b.y1)), 0.f);
  inter = mul(w, h);
  uni = sub(add(area_a, area_b), inter);
}

__device__ __forceinline__ float penalty(const Box& a, const Box& b) {
  const float cw = sub(fmaxf(a.x2, b.x2), fminf(a.x1, b.x1));
  const float ch = sub(fmaxf(a.y2, b.y2), fminf(a.y1, b.y1));
  const float c2 = add(mul(cw, cw), mul(ch, ch));
  const float dx = sub(add(a.x1, a.x2), add(b.x1, b.x2));
  const float dy = sub(add(a.y1, a.y2), add(b.y1, b.y2));
  const float r2 = dvd(add(mul(dx, dx), mul(dy, dy)), 4.f);
  return dvd(r2, c2);
}

__device__ __forceinline__ float pair_value(int mode, const Box& a, const Box& b) {
  float inter, uni;
  switch (mode) {
    case M_IOU:
      inter_union(a, b, inter, uni);
      return dvd(inter, uni);
    case M_GIOU: {
      inter_union(a, b, inter, uni);
      const float ew = fmaxf(sub(fmaxf(a.x2, b.x2), fminf(a.x1, b.x1)), 0.f);
      const float eh = fmaxf(sub(fmaxf(a.y2, b.y2), fminf(a.y1, b.y1)), 0.f);
      const float area = mul(ew, eh);
      return sub(dvd(inter, uni), dvd(sub(area, uni), area));
    }
    case M_PENALTY: return penalty(a, b);
    case M_DIOU_LOSS:
      inter_union(a, b, inter, uni);
      return add(sub(1.f, dvd(inter, uni)), penalty(a, b));
    case M_ARC: {
      const float va = atanf(dvd(sub(a.x2, a.x1), sub(a.y2, a.y1)));
      const float vb = atanf(dvd(sub(b.x2, b.x1), sub(b.y2, b.y1)));
      const float d = sub(va, vb);
      return mul(mul(d, d), 0.40528473456935105f);  // 4 / pi^2 rounded to fp32
    }
  }
  return 0.f;
}

// Tile = kRows rows of boxes1 x (blockDim.x * 4) columns of boxes2. A thread keeps its 4 boxes2 in registers, walks the
// rows (boxes1 row = one broadcast 16-byte load) and writes 4 consecutive outputs per row - one 128-bit store when the
// output row is 16-byte aligned. The first version did a 64-bit division and 8 scalar loads per PAIR (0.07-0.10 of the HBM
// rate on 4096 x 4096). kMode is a template parameter so the per-pair switch is gone as well.
constexpr int kRows = 16;

template <int kMode>
__global__ void __launch_bounds__(128) pairwise_kernel(const float* __restrict__ b1, const float* __restrict__ b2,
                                                       float* __restrict__ out, int M, int N) {
  const int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (j0 >= N) return;
  const int i0 = blockIdx.y * kRows;
  const int i1 = min(i0 + kRows, M);
  Box b[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 v = __ldg((const float4*)b2 + min(j0 + q, N - 1));
    b[q] = Box{v.x, v.y, v.z, v.w};
  }
  const bool vec = (N & 3) == 0 && j0 + 3 < N && ((size_t)out & 15) == 0;
  for (int i = i0; i < i1; ++i) {
    const float4 va = __ldg((const float4*)b1 + i);
    const Box a{va.x, va.y, va.z, va.w};
    float r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) r[q] = pair_value(kMode, a, b[q]);
    float* o = out + (size_t)i * N + j0;
    if (vec) {
      *(float4*)o = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (j0 + q < N) o[q] = r[q];
    }
  }
}

// degenerate-box flag for box_giou's AssertionError (any x2 < x1 or y2 < y1)
__global__ void degenerate_kernel(const float* __restrict__ b, int n, int* flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (b[4 * i + 2] < b[4 * i] || b[4 * i + 3] < b[4 * i + 1])) atomicOr(flag, 1);
}

// ---- backward ---------------------------------------------------------------------------------------
// sub-gradients follow PyTorch: binary max/min split ties evenly, clamp(min=0) passes the gradient at 0.
__device__ __forceinline__ void dmax(float a, float b, float& da, float& db) {
  da = a > b ? 1.f : (a == b ? 0.5f : 0.f);
  db = 1.f - da;
}
__device__ __forceinline__ void dmin(float a, float b, float& da, float& db) {
  da = a < b ? 1.f : (a == b ? 0.5f : 0.f);
  db = 1.f - da;
}

// accumulates g * d value / d (a, b) into ga[4], gb[4]
__device__ __forceinline__ void pair_grad(int mode, const Box& a, const Box& b, float g, float* ga, float* gb) {
  const float wa = a.x2 - a.x1, ha = a.y2 - a.y1, wb = b.x2 - b.x1, hb_ = b.y2 - b.y1;
  const float ltx = fmaxf(a.x1, b.x1), lty = fmaxf(a.y1, b.y1), rbx = fminf(a.x2, b.x2), rby = fminf(a.y2, b.y2);
  const float wr = rbx - ltx, hr = rby - lty;
  const float w = fmaxf(wr, 0.f), h = fmaxf(hr, 0.f);
  const float inter = w * h;
  const float uni = wa * ha + wb * hb_ - inter;
  // coefficient of d inter, d area_a, d area_b in d value, plus enclosing-box terms
  float c_inter = 0.f, c_area = 0.f;  // d value = c_inter * d inter + c_area * (d area_a + d area_b) + ...
  float g_cw = 0.f, g_ch = 0.f;       // d value / d (enclosing width / height, unclamped)
  float g_dx = 0.f, g_dy = 0.f;       // d value / d (centre differences dx, dy)
  const float cwr = fmaxf(a.x2, b.x2) - fminf(a.x1, b.x1), chr = fmaxf(a.y2, b.y2) - fminf(a.y1, b.y1);
  if (mode == M_IOU || mode == M_GIOU || mode == M_DIOU_LOSS) {
    // iou = inter / uni, uni = area_a + area_b - inter
    const float s = (mode == M_DIOU_LOSS) ? -1.f : 1.f;
    c_inter += s * (uni + inter) / (uni * uni);
    c_area += s * (-inter / (uni * uni));
  }
  if (mode == M_GIOU) {
    // giou = iou - 1 + uni / area_c,  area_c = clamp(cw) * clamp(ch)
    const float cw = fmaxf(cwr, 0.f), ch = fmaxf(chr, 0.f);
    const float area_c = cw * ch;
    c_area += 1.f / area_c;
    c_inter += -1.f / area_c;
    const float g_area_c = -uni / (area_c * area_c);
    g_cw += g_area_c * ch * (cwr >= 0.f ? 1.f : 0.f);
    g_ch += g_area_c * cw * (chr >= 0.f ? 1.f : 0.f);
  }
  if (mode == M_PENALTY || mode == M_DIOU_LOSS) {
    const float c2 = cwr * cwr + chr * chr;
    const float dx = (a.x1 + a.x2) - (b.x1 + b.x2), dy = (a.y1 + a.y2) - (b.y1 + b.y2);
    const float r2 = (dx * dx + dy * dy) * 0.25f;
    g_dx += 0.5f * dx / c2;
    g_dy += 0.5f * dy / c2;
    const float g_c2 = -r2 / (c2 * c2);
    g_cw += g_c2 * 2.f * cwr;
    g_ch += g_c2 * 2.f * chr;
  }
  // chain to coordinates
  const float gi_w = c_inter * h * (wr >= 0.f ? 1.f : 0.f);  // d / d (rbx - ltx)
  const float gi_h = c_inter * w * (hr >= 0.f ? 1.f : 0.f);
  float da, db;
  float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // ax1 ay1 ax2 ay2 bx1 by1 bx2 by2
  dmax(a.x1, b.x1, da, db); t[0] -= gi_w * da; t[4] -= gi_w * db;   // ltx
  dmax(a.y1, b.y1, da, db); t[1] -= gi_h * da; t[5] -= gi_h * db;   // lty
  dmin(a.x2, b.x2, da, db); t[2] += gi_w * da; t[6] += gi_w * db;   // rbx
  dmin(a.y2, b.y2, da, db); t[3] += gi_h * da; t[7] += gi_h * db;   // rby
  // areas
  t[0] += c_area * (-ha); t[2] += c_area * ha; t[1] += c_area * (-wa); t[3] += c_area * wa;
  t[4] += c_area * (-hb_); t[6] += c_area * hb_; t[5] += c_area * (-wb); t[7] += c_area * wb;
  // enclosing box
  dmax(a.x2, b.x2, da, db); t[2] += g_cw * da; t[6] += g_cw * db;
  dmin(a.x1, b.x1, da, db); t[0] -= g_cw * da; t[4] -= g_cw * db;
  dmax(a.y2, b.y2, da, db); t[3] += g_ch * da; t[7] += g_ch * db;
  dmin(a.y1, b.y1, da, db); t[1] -= g_ch * da; t[5] -= g_ch * db;
  // centres
  t[0] += g_dx; t[2] += g_dx; t[4] -= g_dx; t[6] -= g_dx;
  t[1] += g_dy; t[3] += g_dy; t[5] -= g_dy; t[7] -= g_dy;
#pragma unroll
  for (int k = 0; k < 4; ++k) { ga[k] += g * t[k]; gb[k] += g * t[4 + k]; }
}

// thread r < M: gradient row of boxes1[r]; thread M + c: gradient row of boxes2[c]   (deterministic, no atomics)
__global__ void pairwise_bwd_kernel(const float* __restrict__ b1, const float* __restrict__ b2,
                                    const float* __restrict__ gout, float* __restrict__ g1, float* __restrict__ g2, int M,
                                    int N, int mode) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M + N) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, dump[4] = {0.f, 0.f, 0.f, 0.f};
  if (t < M) {
    if (!g1) return;
    const Box a = load_box(b1 + 4 * t);
    for (int j = 0; j < N; ++j) pair_grad(mode, a, load_box(b2 + 4 * j), gout[(size_t)t * N + j], acc, dump);
#pragma unroll
    for (int k = 0; k < 4; ++k) g1[4 * t + k] = acc[k];
  } else {
    if (!g2) return;
    const int c = t - M;
    const Box b = load_box(b2 + 4 * c);
    for (int i = 0; i < M; ++i) pair_grad(mode, load_box(b1 + 4 * i), b, gout[(size_t)i * N + c], dump, acc);
#pragma unroll
    for (int k = 0; k < 4; ++k) g2[4 * c + k] = acc[k];
  }
}

}  // namespace

extern "C" {

// mode: 0 IoU, 1 GIoU, 2 DIoU penalty (rho^2/c^2), 3 DIoU loss (= the reference's ciou_loss too), 4 aspect-ratio
// consistency. boxes: fp32 [M,4] / [N,4] xyxy contiguous; out: fp32 [M,N].
int hb_box_pairwise(const float* boxes1, const float* boxes2, float* out, int M, int N, int mode, void* stream) {
  const long long total = (long long)M * N;
  if (total == 0) return 0;
  if ((((size_t)boxes1) | ((size_t)boxes2)) & 15) return (int)cudaErrorMisalignedAddress;   // [*, 4] fp32 rows: float4 loads
  const dim3 grid((N + 4 * 128 - 1) / (4 * 128), (M + kRows - 1) / kRows);
  if (grid.y > 65535) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case M_IOU: pairwise_kernel<M_IOU><<<grid, 128, 0, st>>>(boxes1, boxes2, out, M, N); break;
    case M_GIOU: pairwise_kernel<M_GIOU><<<grid, 128, 0, st>>>(boxes1, boxes2, out, M, N); break;
    case M_PENALTY: pairwise_kernel<M_PENALTY><<<grid, 128, 0, st>>>(boxes1, boxes2, out, M, N); break;
    case M_DIOU_LOSS: pairwise_kernel<M_DIOU_LOSS><<<grid, 128, 0, st>>>(boxes1, boxes2, out, M, N); break;
    case M_ARC: pairwise_kernel<M_ARC><<<grid, 128, 0, st>>>(boxes1, boxes2, out, M, N); break;
    default: return (int)cudaErrorInvalidValue;
  }
  HB_LAUNCH_CHECK();
  return 0;
}

// flag (device int, pre-zeroed) is set to 1 if any box has x2 < x1 or y2 < y1
int hb_box_degenerate(const float* boxes, int n, int* flag, void* stream) {
  if (n == 0) return 0;
  degenerate_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(boxes, n, flag);
  HB_LAUNCH_CHECK();
  return 0;
}

// gradients of sum(gout * op(boxes1, boxes2)) for modes 0-3; g1 [M,4] / g2 [N,4] may be NULL
int hb_box_pairwise_bwd(const float* boxes1, const float* boxes2, const float* gout, float* g1, float* g2, int M, int N,
                        int mode, void* stream) {
  if (mode == M_ARC) return (int)cudaErrorInvalidValue;
  if (M + N == 0) return 0;
  pairwise_bwd_kernel<<<(M + N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(boxes1, boxes2, gout, g1, g2, M, N, mode);
  HB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
