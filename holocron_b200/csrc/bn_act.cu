// Training-mode BatchNorm + branch-sum + activation, fused for NHWC bf16 activations.
//
// The unit being fused is the reference's "conv -> BatchNorm2d -> act" sequence
// (holocron/models/utils.py:28-86 conv_sequence) and the RepVGG block
//   out = act( BN3(conv3x3(x)) + BN1(conv1x1(x)) [+ BNid(x)] )        (models/classification/repvgg.py:71-73)
// generalised to B <= 3 normalised branches u_b of identical shape [M, C] plus an optional un-normalised
// residual. The reference runs one cuDNN/ATen kernel per BN, add and activation (>= 8 HBM passes per
// RepBlock); here the whole thing is
//   stats      : per-channel sum / sum-of-squares PARTIALS, normally produced by whoever wrote the tensor (the epilogue of
//                the tensor-core convolution, conv_fprop.cu / conv_rows.cu, or the forward pass below for a block's own
//                output); bn_stats_partials_kernel is the stand-alone pass for tensors that come without partials
//   finalize   : C-sized: fixed-order fp64 sum of the partials (deterministic: no floating-point atomics anywhere),
//                mean, rstd, scale = gamma*rstd, shift = beta - mean*scale, running-stat update
//   forward    : one pass: out = act(sum_b (scale_b * u_b + shift_b) + residual)
//   bwd reduce : one pass: sum(dz), sum(dz * xhat_b)    with dz = dOut * act'(z), z recomputed (not stored)
//   bwd apply  : one pass: du_b = scale_b * (dz - mean(dz) - xhat_b * mean(dz*xhat_b)), dresidual = dz
// Every thread owns 8 consecutive channels (one 128-bit bf16 vector) of a row; rows are walked with a stride
// that keeps the thread's channel group fixed, so per-channel parameters live in registers.
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "act.cuh"

namespace {

using namespace hb;

constexpr int kThreads = 256;
constexpr int kMaxBranches = 3;

struct Branches {
  const __nv_bfloat16* u[kMaxBranches];
  int n;
};

// thread geometry shared by all kernels: tx = channel group inside the block's channel slab, ty = row lane
struct Geo {
  int cg_total;  // C / 8
  int cg_t;      // channel groups per block (<= 32)
  int rows_t;    // row lanes per block
  __host__ static Geo make(int C) {
    Geo g;
    g.cg_total = C / 8;
    // balanced channel slabs: 38 groups -> 2 x 19 rather than 32 + 6 (the ragged last slab kept 80 % of its block's
    // threads idle on ReXNet's widths: 300, 366, 432, 576, ... channels)
    const int slabs = (g.cg_total + 31) / 32;
    g.cg_t = (g.cg_total + slabs - 1) / slabs;
    g.rows_t = kThreads / g.cg_t;
    return g;
  }
};

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float* f) {
  Vec16<__nv_bfloat16> v = ld16_stream(p);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = __bfloat162float(v.v[j]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float* f) {
  Vec16<__nv_bfloat16> v;
#pragma unroll
  for (int j = 0; j < 8; ++j) v.v[j] = __float2bfloat16_rn(f[j]);
  st16(p, v);
}

// ---------------------------------------------------------------------------------------------------
// stand-alone statistics pass: parts[blockIdx.x][c] = (sum, sum of squares) of this block's rows of u [M, C]
// grid = (row blocks = slots, channel slabs)
__global__ void __launch_bounds__(kThreads) bn_stats_partials_kernel(const __nv_bfloat16* __restrict__ u, int M, int C, Geo g,
                                                                     float* __restrict__ parts) {
  __shared__ float red[2][kThreads * 8];
  const int tx = threadIdx.x % g.cg_t, ty = threadIdx.x / g.cg_t;
  const int cg = blockIdx.y * g.cg_t + tx;
  const bool active = ty < g.rows_t && cg < g.cg_total;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
  if (active) {
    const size_t row_stride = (size_t)gridDim.x * g.rows_t;
    size_t m = (size_t)blockIdx.x * g.rows_t + ty;
    // two rows in flight per trip
    for (; m + row_stride < (size_t)M; m += 2 * row_stride) {
      float a[8], b[8];
      load8(u + m * C + cg * 8, a);
      load8(u + (m + row_stride) * C + cg * 8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += a[j] + b[j]; q[j] += a[j] * a[j] + b[j] * b[j]; }
    }
    if (m < (size_t)M) {
      float a[8];
      load8(u + m * C + cg * 8, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += a[j]; q[j] += a[j] * a[j]; }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { red[0][threadIdx.x * 8 + j] = s[j]; red[1][threadIdx.x * 8 + j] = q[j]; }
  __syncthreads();
  // threads 0 .. cg_t*8-1 each own one channel of the slab: sum over row lanes in a fixed order
  const int nch = g.cg_t * 8;
  for (int ch = threadIdx.x; ch < nch; ch += kThreads) {
    const int ctx = ch / 8, j = ch % 8;
    const int gcg = blockIdx.y * g.cg_t + ctx;
    if (gcg >= g.cg_total) continue;
    double a = 0.0, b = 0.0;
    for (int r = 0; r < g.rows_t; ++r) {
      a += (double)red[0][(r * g.cg_t + ctx) * 8 + j];
      b += (double)red[1][(r * g.cg_t + ctx) * 8 + j];
    }
    *reinterpret_cast<float2*>(parts + ((size_t)blockIdx.x * C + gcg * 8 + j) * 2) = make_float2((float)a, (float)b);
  }
}

// ---------------------------------------------------------------------------------------------------
// finalize: per branch b and channel c
struct FinalizeParams {
  const float* parts[kMaxBranches];   // [slots_b][C][2] (sum, sum of squares) partials
  int slots[kMaxBranches];
  const float* gamma[kMaxBranches];
  const float* beta[kMaxBranches];
  float* running_mean[kMaxBranches];  // may be null
  float* running_var[kMaxBranches];
  long long* num_batches_tracked[kMaxBranches];  // int64 scalar per branch, may be null
  float* mean;   // [B][C] out
  float* rstd;   // [B][C] out
  float* scale;  // [B][C] out
  float* shift;  // [B][C] out
  int B, C, M;
  int C_logical;  // channels >= C_logical are padding: scale = shift = 0, no parameter / running-stat access
  float eps, momentum;
};
// block = 8 channels (threadIdx.x: one 64-byte run of the [slot][C][2] partials) x 32 slot lanes (threadIdx.y). A lane adds
// slots L, L + 32, ... with four independent loads in flight, the 32 lane sums are then combined 8-by-8 in lane order: a
// fixed summation order (run-to-run identical) whose dependent-load chain is slots / 128 long. (32 channels x 8 lanes: slots /
// 8 dependent L2 round trips and C / 32 blocks - two blocks for a 48-channel layer - made each of these ~10-20 us.)
constexpr int kFinCh = 8, kFinLanes = 32;

__global__ void __launch_bounds__(kFinCh * kFinLanes) bn_finalize_kernel(FinalizeParams p) {
  __shared__ double red[2][kFinLanes][kFinCh];
  const int c = blockIdx.x * kFinCh + threadIdx.x;
  const int b = blockIdx.y;
  double s = 0.0, q = 0.0;
  if (c < p.C_logical) {
    const float* pp = p.parts[b] + (size_t)c * 2;
    const size_t row = (size_t)p.C * 2;
    const int n = p.slots[b];
    int k = threadIdx.y;
    for (; k + 3 * kFinLanes < n; k += 4 * kFinLanes) {
      float2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float2*>(pp + (size_t)(k + u * kFinLanes) * row);
#pragma unroll
      for (int u = 0; u < 4; ++u) { s += (double)v[u].x; q += (double)v[u].y; }
    }
    for (; k < n; k += kFinLanes) {
      const float2 v = *reinterpret_cast<const float2*>(pp + (size_t)k * row);
      s += (double)v.x; q += (double)v.y;
    }
  }
  red[0][threadIdx.y][threadIdx.x] = s;
  red[1][threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y < 4) {
    s = 0.0; q = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { s += red[0][threadIdx.y * 8 + j][threadIdx.x]; q += red[1][threadIdx.y * 8 + j][threadIdx.x]; }
  }
  __syncthreads();
  if (threadIdx.y < 4) { red[0][threadIdx.y][threadIdx.x] = s; red[1][threadIdx.y][threadIdx.x] = q; }
  __syncthreads();
  if (threadIdx.y != 0 || c >= p.C) return;
  const size_t o = (size_t)b * p.C + c;
  if (c >= p.C_logical) {
    p.mean[o] = 0.f; p.rstd[o] = 0.f; p.scale[o] = 0.f; p.shift[o] = 0.f;
    return;
  }
  s = 0.0; q = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j) { s += red[0][j][threadIdx.x]; q += red[1][j][threadIdx.x]; }
  const double mean = s / p.M;
  double var = q / p.M - mean * mean;
  if (var < 0) var = 0;
  const float rstd = (float)(1.0 / sqrt(var + (double)p.eps));
  const float g = p.gamma[b] ? p.gamma[b][c] : 1.f;
  const float be = p.beta[b] ? p.beta[b][c] : 0.f;
  const float sc = g * rstd;
  p.mean[o] = (float)mean;
  p.rstd[o] = rstd;
  p.scale[o] = sc;
  p.shift[o] = be - (float)mean * sc;
  if (c == 0 && p.num_batches_tracked[b]) *p.num_batches_tracked[b] += 1;
  if (p.running_mean[b]) {
    const double unbiased = p.M > 1 ? var * ((double)p.M / (double)(p.M - 1)) : var;
    p.running_mean[b][c] = (1.f - p.momentum) * p.running_mean[b][c] + p.momentum * (float)mean;
    p.running_var[b][c] = (1.f - p.momentum) * p.running_var[b][c] + p.momentum * (float)unbiased;
  }
}

// eval-mode affine from running statistics: scale = gamma / sqrt(var + eps), shift = beta - mean * scale
__global__ void bn_eval_affine_kernel(const float* gamma, const float* beta, const float* rmean, const float* rvar,
                                      float eps, int C, int C_logical, float* scale, float* shift, float* mean,
                                      float* rstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (c >= C_logical) {
    scale[c] = 0.f; shift[c] = 0.f;
    if (mean) mean[c] = 0.f;
    if (rstd) rstd[c] = 0.f;
    return;
  }
  const float r = 1.f / sqrtf(rvar[c] + eps);
  const float sc = (gamma ? gamma[c] : 1.f) * r;
  scale[c] = sc;
  shift[c] = (beta ? beta[c] : 0.f) - rmean[c] * sc;
  if (mean) mean[c] = rmean[c];
  if (rstd) rstd[c] = r;
}

// ---------------------------------------------------------------------------------------------------
// forward: out = act(sum_b scale_b*u_b + shift_b (+ residual))
struct FwdParams {
  Branches br;
  const float* scale;  // [B][C]
  const float* shift;  // [B][C]
  const __nv_bfloat16* residual;
  __nv_bfloat16* out;
  int M, C, act;
  float slope;
  int res_after;  // 1: out = act(z) + residual (ResNet-style shortcut after the activation); 0: act(z + residual)
  float* out_stats;  // optional [gridDim.x][C][2]: (sum, sum of squares) partials of the bf16 OUTPUT - the statistics the
                     // identity-branch BatchNorm of the NEXT RepVGG block needs, produced while the data is in registers
};
using RawVec = Vec16<__nv_bfloat16>;

__device__ __forceinline__ void unpack8(const RawVec& v, float* f) {
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = __bfloat162float(v.v[j]);
}

__device__ __forceinline__ void lds8(const float* p, float* f);

// ---- per-thread cp.async ring ---------------------------------------------------------------------
// Every thread streams ITS OWN 16-byte vectors (one per input tensor and row) global -> shared with cp.async, kDepth rows
// ahead of the row it is computing on; nothing else reads those slots, so the ring needs no barrier at all. This puts
// kDepth * (#inputs) * 16 B per thread in flight without holding them in registers: the register-staged version of these
// kernels had ~36 KB per SM in flight and sat at 2.2 - 3.4 TB/s with long-scoreboard stalls (profiles/r01_bn_ncu.md);
// HBM needs ~60 KB per SM (44 GB/s per SM x ~1.3 us loaded latency).
// rows in flight per thread as a function of the number of streamed tensors NT: what matters is BYTES in flight per SM
// (depth * NT * 16 B * 256 threads * resident blocks). With depth 3 the single-branch units (ReXNet / Darknet / UNet3+ /
// YOLOv4: one input tensor) had only ~49 KB per SM in flight and ran at 1.6 TB/s; deeper rings for fewer tensors.
// depth + 1 slots: keep the slot count a power of two (the slot index is k % slots on a 64-bit row counter)
__host__ __device__ constexpr int ring_depth(int nt) { return nt <= 2 ? 7 : 3; }

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ RawVec lds16(uint32_t saddr) {
  RawVec r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.raw.x), "=r"(r.raw.y), "=r"(r.raw.z), "=r"(r.raw.w) : "r"(saddr));
  return r;
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Walks the rows m = m0, m0 + stride, ... < M of one thread. NT input tensors; `src(t)` gives tensor t's base pointer.
template <int NT>
struct RowRing {
  static constexpr int kDepth = ring_depth(NT);
  static constexpr int kSlots = kDepth + 1;   // the slot refilled at step k is the one consumed at step k-1
  uint32_t base;       // shared address of this thread's slot 0 / tensor 0
  size_t m0, stride, M;
  size_t col_off;      // element offset of the thread's 8 channels inside a row
  int C;
  const __nv_bfloat16* src[NT > 0 ? NT : 1];
  // slot s, tensor t of this thread
  __device__ __forceinline__ uint32_t addr(int s, int t) const { return base + (uint32_t)((s * NT + t) * kThreads * 16); }
  __device__ __forceinline__ bool valid(size_t k) const { return m0 + k * stride < M; }
  __device__ __forceinline__ size_t off(size_t k) const { return (m0 + k * stride) * C + col_off; }
  __device__ __forceinline__ void issue(size_t k) const {
    if (valid(k)) {
      const int s = (int)(k % kSlots);
      const size_t o = off(k);
#pragma unroll
      for (int t = 0; t < NT; ++t)
        if (src[t]) cp_async16(addr(s, t), src[t] + o);
    }
    cp_async_commit();   // always: keeps the group count uniform
  }
  __device__ __forceinline__ void prologue() const {
#pragma unroll
    for (int k = 0; k < kDepth; ++k) issue(k);
  }
  // row k has landed
  __device__ __forceinline__ void wait() const { cp_async_wait<kDepth - 1>(); }
};

// forward: out = act(sum_b scale_b*u_b + shift_b (+ residual))
// kStats: also accumulate the output statistics (separate instantiation: the plain kernel keeps its register budget; with
// the 16 extra accumulators the 3-branch kernel would drop from 3 to 2 resident blocks per SM, so the statistics variant is
// compiled for 3 blocks explicitly)
template <int NB, bool kStats>
__global__ void __launch_bounds__(kThreads, 3) bn_act_fwd_kernel(FwdParams p, Geo g) {
  extern __shared__ __align__(16) uint8_t ring_smem[];
  const int tx = threadIdx.x % g.cg_t, ty = threadIdx.x / g.cg_t;
  const int cg = blockIdx.y * g.cg_t + tx;
  const bool active = ty < g.rows_t && cg < g.cg_total;
  float os[8], oq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { os[j] = 0.f; oq[j] = 0.f; }
  // folded per-channel constants of this block's channel slab live in shared memory (read as float4 pairs per row): with
  // them in registers the statistics variant of the 3-branch kernel spilled inside the streaming loop
  __shared__ __align__(16) float k_sc[kMaxBranches][256];
  __shared__ __align__(16) float k_sh[256];
  {
    const int nch = g.cg_t * 8;
    for (int ch = threadIdx.x; ch < nch; ch += kThreads) {
      const int c = blockIdx.y * nch + ch;
      float shv = 0.f;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        k_sc[b][ch] = c < p.C ? p.scale[(size_t)b * p.C + c] : 0.f;
        shv += c < p.C ? p.shift[(size_t)b * p.C + c] : 0.f;
      }
      k_sh[ch] = shv;
    }
    __syncthreads();
  }
  if (!active && !kStats) return;
  if (active) {
    const bool has_res = p.residual != nullptr;
    RowRing<NB + 1> ring;
    ring.base = smem_addr(ring_smem) + threadIdx.x * 16;
    ring.m0 = (size_t)blockIdx.x * g.rows_t + ty;
    ring.stride = (size_t)gridDim.x * g.rows_t;
    ring.M = (size_t)p.M;
    ring.col_off = (size_t)cg * 8;
    ring.C = p.C;
#pragma unroll
    for (int b = 0; b < NB; ++b) ring.src[b] = p.br.u[b];
    ring.src[NB] = p.residual;
    ring.prologue();
    for (size_t k = 0; ring.valid(k); ++k) {
      ring.wait();
      const int s = (int)(k % RowRing<NB + 1>::kSlots);
      RawVec u[NB > 0 ? NB : 1], r;
#pragma unroll
      for (int b = 0; b < NB; ++b) u[b] = lds16(ring.addr(s, b));
      if (has_res) r = lds16(ring.addr(s, NB));
      ring.issue(k + RowRing<NB + 1>::kDepth);
      float z[8];
      lds8(&k_sh[tx * 8], z);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        float f[8], sc[8];
        lds8(&k_sc[b][tx * 8], sc);
        unpack8(u[b], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] = fmaf(sc[j], f[j], z[j]);
      }
      float rr[8];
      if (has_res) unpack8(r, rr);
      if (has_res && !p.res_after) {
        if (p.act == ACT_FRELU) {
#pragma unroll
          for (int j = 0; j < 8; ++j) z[j] = max_nan(z[j], rr[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) z[j] += rr[j];
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] = act_fwd(p.act, z[j], p.slope);
      if (has_res && p.res_after) {
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] += rr[j];
      }
      Vec16<__nv_bfloat16> ov;
#pragma unroll
      for (int j = 0; j < 8; ++j) ov.v[j] = __float2bfloat16_rn(z[j]);
      st16(p.out + ring.off(k), ov);
      if (kStats) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float f = __bfloat162float(ov.v[j]);   // statistics of what the consumer will read
          os[j] += f; oq[j] = fmaf(f, f, oq[j]);
        }
      }
    }
  }
  if (!kStats) return;
  // block partial of the output statistics: the ring memory is free now (every cp.async group has been waited for)
  cp_async_wait<0>();
  __syncthreads();
  float* red = reinterpret_cast<float*>(ring_smem);   // [2][kThreads * 8] floats = 16 KB <= smallest ring
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[threadIdx.x * 8 + j] = active ? os[j] : 0.f;
    red[kThreads * 8 + threadIdx.x * 8 + j] = active ? oq[j] : 0.f;
  }
  __syncthreads();
  const int nch = g.cg_t * 8;
  for (int ch = threadIdx.x; ch < nch; ch += kThreads) {
    const int ctx = ch / 8, j = ch % 8;
    const int gcg = blockIdx.y * g.cg_t + ctx;
    if (gcg >= g.cg_total) continue;
    double a = 0.0, b = 0.0;
    for (int r = 0; r < g.rows_t; ++r) {
      a += (double)red[(r * g.cg_t + ctx) * 8 + j];
      b += (double)red[kThreads * 8 + (r * g.cg_t + ctx) * 8 + j];
    }
    *reinterpret_cast<float2*>(p.out_stats + ((size_t)blockIdx.x * p.C + gcg * 8 + j) * 2) = make_float2((float)a, (float)b);
  }
}

// ---------------------------------------------------------------------------------------------------
// backward
struct BwdParams {
  Branches br;
  const __nv_bfloat16* dout;
  const __nv_bfloat16* residual;
  const float* scale;  // [B][C]
  const float* shift;
  const float* mean;
  const float* rstd;
  double* sums;        // [1 + B][C]: sum dz, sum dz*u_b (final, written by bn_bwd_finalize_kernel)
  double* part;        // [row blocks][1 + B][C]: per-block partials of the above (pass 1), summed in a fixed order
  __nv_bfloat16* du[kMaxBranches];
  __nv_bfloat16* dres;  // may be null
  int M, C, act;
  float slope;
  int train;  // 1: batch statistics (full BN backward); 0: running statistics (du = scale * dz)
  int res_after;
};

// Per-channel constants of the block's channel slab in shared memory, read as float4 pairs (8 channels per thread).
//   z    = sum_b scale_b * u_b + shift
//   du_b = scale_b * (dz - mean(dz) - xhat_b * mean(dz * xhat_b))  =  scale_b * dz + cu_b * u_b + c0_b
// with cu_b = -scale_b * rstd_b * mdzx_b, c0_b = -scale_b * mdz - cu_b * mean_b and
// mdzx_b = mean(dz * xhat_b) = rstd_b * (sum(dz * u_b) / M - mean_b * mdz) from the first pass' sums.
struct SlabConsts {
  float scale[kMaxBranches][256];
  float cu[kMaxBranches][256];
  float c0[kMaxBranches][256];
  float shift[256];   // sum over branches
};

template <int NB>
__device__ __forceinline__ void load_slab_consts(SlabConsts& k, const BwdParams& p, const Geo& g, bool with_means) {
  const int nch = g.cg_t * 8;
  const double invM = 1.0 / (double)p.M;
  for (int ch = threadIdx.x; ch < nch; ch += kThreads) {
    const int c = blockIdx.y * nch + ch;
    const bool ok = c < p.C;
    float sh = 0.f;
    const double mdz = (with_means && p.train && ok) ? p.sums[c] * invM : 0.0;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const size_t o = (size_t)b * p.C + c;
      const float sc = ok ? p.scale[o] : 0.f;
      k.scale[b][ch] = sc;
      sh += ok ? p.shift[o] : 0.f;
      float cu = 0.f, c0 = 0.f;
      if (with_means && p.train && ok) {
        const double mean = (double)p.mean[o], rstd = (double)p.rstd[o];
        const double mdzx = rstd * (p.sums[(size_t)(1 + b) * p.C + c] * invM - mean * mdz);
        const double cud = -(double)sc * rstd * mdzx;
        cu = (float)cud;
        c0 = (float)(-(double)sc * mdz - cud * mean);
      }
      k.cu[b][ch] = cu;
      k.c0[b][ch] = c0;
    }
    k.shift[ch] = sh;
  }
  __syncthreads();
}

__device__ __forceinline__ void lds8(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// dz = d out / d z (z = normalised branch sum [+ residual]); dr = gradient reaching the residual input
template <int NB>
__device__ __forceinline__ void recompute_dz(const BwdParams& p, const SlabConsts& k, int ch0, const RawVec* uv,
                                             const RawVec& rv, const RawVec& dv, float (*u)[8], float* dz, float* dr) {
  float z[8];
  lds8(&k.shift[ch0], z);
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    float sc[8];
    lds8(&k.scale[b][ch0], sc);
    unpack8(uv[b], u[b]);
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = fmaf(sc[j], u[b][j], z[j]);
  }
  float d[8];
  unpack8(dv, d);
  if (p.residual && p.res_after) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { dz[j] = d[j] * act_grad(p.act, z[j], p.slope); dr[j] = d[j]; }
    return;
  }
  if (p.residual) {
    float r[8];
    unpack8(rv, r);
    if (p.act == ACT_FRELU) {
      // binary max: the gradient goes to the larger argument, ties are split evenly (PyTorch semantics)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float gate = z[j] > r[j] ? 1.f : (z[j] == r[j] ? 0.5f : 0.f);
        dz[j] = d[j] * gate;
        dr[j] = d[j] - dz[j];
      }
      return;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] += r[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { dz[j] = d[j] * act_grad(p.act, z[j], p.slope); dr[j] = dz[j]; }
}

template <int NB>
__device__ __forceinline__ void init_bwd_ring(RowRing<NB + 2>& ring, const BwdParams& p, const Geo& g, uint8_t* ring_smem,
                                              int ty, int cg) {
  ring.base = smem_addr(ring_smem) + threadIdx.x * 16;
  ring.m0 = (size_t)blockIdx.x * g.rows_t + ty;
  ring.stride = (size_t)gridDim.x * g.rows_t;
  ring.M = (size_t)p.M;
  ring.col_off = (size_t)cg * 8;
  ring.C = p.C;
#pragma unroll
  for (int b = 0; b < NB; ++b) ring.src[b] = p.br.u[b];
  ring.src[NB] = (p.residual && !p.res_after) ? p.residual : nullptr;   // its value only matters inside act()
  ring.src[NB + 1] = p.dout;
}

// pass 1: part[blk][0][c] = sum_m dz, part[blk][1+b][c] = sum_m dz * u_b over the rows of this block (no atomics)
template <int NB>
__global__ void __launch_bounds__(kThreads, 2) bn_act_bwd_reduce_kernel(BwdParams p, Geo g) {
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  SlabConsts& k = *reinterpret_cast<SlabConsts*>(dyn_smem);
  float* red = reinterpret_cast<float*>(dyn_smem + sizeof(SlabConsts));   // [kThreads * 8]
  uint8_t* ring_smem = dyn_smem + sizeof(SlabConsts) + kThreads * 8 * sizeof(float);
  load_slab_consts<NB>(k, p, g, false);
  const int tx = threadIdx.x % g.cg_t, ty = threadIdx.x / g.cg_t;
  const int cg = blockIdx.y * g.cg_t + tx;
  const bool active = ty < g.rows_t && cg < g.cg_total;
  float acc[1 + NB][8];
#pragma unroll
  for (int i = 0; i < 1 + NB; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  if (active) {
    RowRing<NB + 2> ring;
    init_bwd_ring<NB>(ring, p, g, ring_smem, ty, cg);
    ring.prologue();
    for (size_t kk = 0; ring.valid(kk); ++kk) {
      ring.wait();
      const int s = (int)(kk % RowRing<NB + 2>::kSlots);
      RawVec uv[NB > 0 ? NB : 1], rv, dv;
#pragma unroll
      for (int b = 0; b < NB; ++b) uv[b] = lds16(ring.addr(s, b));
      if (ring.src[NB]) rv = lds16(ring.addr(s, NB));
      dv = lds16(ring.addr(s, NB + 1));
      ring.issue(kk + RowRing<NB + 2>::kDepth);
      float u[NB > 0 ? NB : 1][8], dz[8], dr[8];
      recompute_dz<NB>(p, k, tx * 8, uv, rv, dv, u, dz, dr);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[0][j] += dz[j];
#pragma unroll
      for (int b = 0; b < NB; ++b) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[1 + b][j] = fmaf(dz[j], u[b][j], acc[1 + b][j]);
      }
    }
  }
  const int nch = g.cg_t * 8;
#pragma unroll
  for (int i = 0; i < 1 + NB; ++i) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = acc[i][j];
    __syncthreads();
    for (int ch = threadIdx.x; ch < nch; ch += kThreads) {
      const int ctx = ch / 8, j = ch % 8;
      const int gcg = blockIdx.y * g.cg_t + ctx;
      if (gcg >= g.cg_total) continue;
      double a = 0.0;
      for (int r = 0; r < g.rows_t; ++r) a += (double)red[(r * g.cg_t + ctx) * 8 + j];
      p.part[((size_t)blockIdx.x * (1 + NB) + i) * p.C + gcg * 8 + j] = a;
    }
  }
}

// pass 2: du_b = scale_b * dz + cu_b * u_b + c0_b, dres = dr
template <int NB>
__global__ void __launch_bounds__(kThreads, 2) bn_act_bwd_apply_kernel(BwdParams p, Geo g) {
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  SlabConsts& k = *reinterpret_cast<SlabConsts*>(dyn_smem);
  uint8_t* ring_smem = dyn_smem + sizeof(SlabConsts);
  load_slab_consts<NB>(k, p, g, true);
  const int tx = threadIdx.x % g.cg_t, ty = threadIdx.x / g.cg_t;
  const int cg = blockIdx.y * g.cg_t + tx;
  if (ty >= g.rows_t || cg >= g.cg_total) return;
  RowRing<NB + 2> ring;
  init_bwd_ring<NB>(ring, p, g, ring_smem, ty, cg);
  ring.prologue();
  for (size_t kk = 0; ring.valid(kk); ++kk) {
    ring.wait();
    const int s = (int)(kk % RowRing<NB + 2>::kSlots);
    RawVec uv[NB > 0 ? NB : 1], rv, dv;
#pragma unroll
    for (int b = 0; b < NB; ++b) uv[b] = lds16(ring.addr(s, b));
    if (ring.src[NB]) rv = lds16(ring.addr(s, NB));
    dv = lds16(ring.addr(s, NB + 1));
    ring.issue(kk + RowRing<NB + 2>::kDepth);
    const size_t off = ring.off(kk);
    float u[NB > 0 ? NB : 1][8], dz[8], dr[8];
    recompute_dz<NB>(p, k, tx * 8, uv, rv, dv, u, dz, dr);
    if (p.dres) store8(p.dres + off, dr);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      if (p.du[b]) {
        float sc[8], cu[8], c0[8], d[8];
        lds8(&k.scale[b][tx * 8], sc);
        lds8(&k.cu[b][tx * 8], cu);
        lds8(&k.c0[b][tx * 8], c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = fmaf(sc[j], dz[j], fmaf(cu[j], u[b][j], c0[j]));
        store8(p.du[b] + off, d);
      }
    }
  }
}

// Between the two passes: sums[i][c] = sum over the row blocks of part[blk][i][c] (one warp per channel, fixed order), then
//   dgamma_b = sum dz*xhat_b = rstd_b * (sum dz*u_b - mean_b * sum dz),   dbeta_b = sum dz     (evaluated in fp64)
// written to dgamma/dbeta ([B][C] fp32, optional) and / or ACCUMULATED into the parameters' gradient buffers gacc/bacc
// (`.grad` storage of the BatchNorm weight / bias: what autograd's AccumulateGrad would do with one more kernel each).
struct BwdFinalizeParams {
  const double* part; double* sums; const float* mean; const float* rstd;
  float* dgamma; float* dbeta;
  float* gacc[kMaxBranches]; float* bacc[kMaxBranches];
  int nblocks, B, C, C_logical;
};
// same block shape as bn_finalize_kernel: 8 channels x 32 row-block lanes, four independent rows in flight, fixed order
__global__ void __launch_bounds__(kFinCh * kFinLanes) bn_bwd_finalize_kernel(BwdFinalizeParams p) {
  __shared__ double red[1 + kMaxBranches][kFinLanes][kFinCh];
  const int c = blockIdx.x * kFinCh + threadIdx.x;
  double acc[1 + kMaxBranches];
#pragma unroll
  for (int i = 0; i <= kMaxBranches; ++i) acc[i] = 0.0;
  if (c < p.C) {
    const size_t row = (size_t)(1 + p.B) * p.C;
    int k = threadIdx.y;
    for (; k + 3 * kFinLanes < p.nblocks; k += 4 * kFinLanes) {
      double v[4][1 + kMaxBranches];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i <= kMaxBranches; ++i)
          if (i <= p.B) v[u][i] = p.part[(size_t)(k + u * kFinLanes) * row + (size_t)i * p.C + c];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i <= kMaxBranches; ++i)
          if (i <= p.B) acc[i] += v[u][i];
    }
    for (; k < p.nblocks; k += kFinLanes) {
#pragma unroll
      for (int i = 0; i <= kMaxBranches; ++i)
        if (i <= p.B) acc[i] += p.part[(size_t)k * row + (size_t)i * p.C + c];
    }
  }
#pragma unroll
  for (int i = 0; i <= kMaxBranches; ++i) red[i][threadIdx.y][threadIdx.x] = acc[i];
  __syncthreads();
  if (threadIdx.y < 4) {
#pragma unroll
    for (int i = 0; i <= kMaxBranches; ++i) {
      double t = 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) t += red[i][threadIdx.y * 8 + j][threadIdx.x];
      acc[i] = t;
    }
  }
  __syncthreads();
  if (threadIdx.y < 4) {
#pragma unroll
    for (int i = 0; i <= kMaxBranches; ++i) red[i][threadIdx.y][threadIdx.x] = acc[i];
  }
  __syncthreads();
  if (threadIdx.y != 0 || c >= p.C) return;
  double tot[1 + kMaxBranches];
#pragma unroll
  for (int i = 0; i <= kMaxBranches; ++i) {
    double t = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) t += red[i][j][threadIdx.x];
    tot[i] = t;
  }
  for (int i = 0; i <= p.B; ++i) p.sums[(size_t)i * p.C + c] = tot[i];
  for (int b = 0; b < p.B; ++b) {
    const size_t o = (size_t)b * p.C + c;
    const float dg = (float)((double)p.rstd[o] * (tot[1 + b] - (double)p.mean[o] * tot[0]));
    const float db = (float)tot[0];
    if (p.dgamma) p.dgamma[o] = dg;
    if (p.dbeta) p.dbeta[o] = db;
    if (c < p.C_logical) {
      if (p.gacc[b]) p.gacc[b][c] += dg;
      if (p.bacc[b]) p.bacc[b][c] += db;
    }
  }
}

// persistent grid: `per_sm` blocks per SM, grid-stride over rows
inline dim3 make_grid(const Geo& g, int M, int z, int per_sm = 4, int min_rows = 4) {
  const int slabs = (g.cg_total + g.cg_t - 1) / g.cg_t;
  long long row_blocks = ((long long)M + g.rows_t - 1) / g.rows_t;
  long long cap = (HB_NUM_SMS * per_sm) / slabs;
  if (cap < 1) cap = 1;
  long long want = (row_blocks + min_rows - 1) / min_rows;   // at least ~min_rows rows per lane when there is enough work
  if (want < 1) want = 1;
  if (want > cap) want = cap;
  return dim3((unsigned)want, (unsigned)slabs, (unsigned)z);
}

inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

template <typename K>
inline cudaError_t allow_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// Resident blocks per SM of one kernel instantiation with SMEM dynamic bytes, asked from the runtime once. The grids are
// persistent (every block walks M / gridDim.x rows), so a grid of 4 blocks per SM of a kernel that only fits 3 runs as one
// full wave plus a one-third-occupied second wave: bn_act_fwd_kernel<2, 1> streamed at 3.3 TB/s with 592 blocks where the
// 444-block <3, 1> reached 4.8 TB/s (profiles/r02_launches_repvgg_a0_b256.csv).
template <typename K>
inline int resident_blocks(K kernel, size_t smem) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kThreads, smem) != cudaSuccess || n < 1) n = 1;
  return n;
}
#define HB_BN_OCC(KERNEL, NB, SMEM, OUT)                                                     \
  {                                                                                          \
    static int cached_occ_ = 0;                                                              \
    if (!cached_occ_) {                                                                      \
      if (allow_smem(KERNEL<NB>, 200 * 1024) != cudaSuccess) return (int)cudaErrorInvalidValue; \
      cached_occ_ = resident_blocks(KERNEL<NB>, SMEM);                                       \
    }                                                                                        \
    OUT = cached_occ_;                                                                       \
  }
#define HB_BN_OCC_DISPATCH(KERNEL, B, SMEM, OUT)                                             \
  switch (B) {                                                                               \
    case 0: HB_BN_OCC(KERNEL, 0, SMEM, OUT) break;                                           \
    case 1: HB_BN_OCC(KERNEL, 1, SMEM, OUT) break;                                           \
    case 2: HB_BN_OCC(KERNEL, 2, SMEM, OUT) break;                                           \
    default: HB_BN_OCC(KERNEL, 3, SMEM, OUT) break;                                          \
  }

#define HB_BN_LAUNCH(KERNEL, NB, GRID, SMEM, ST, ...)                                        \
  {                                                                                          \
    static bool ready = false;                                                               \
    if (!ready) {                                                                            \
      if (allow_smem(KERNEL<NB>, 200 * 1024) != cudaSuccess) return (int)cudaErrorInvalidValue; \
      ready = true;                                                                          \
    }                                                                                        \
    KERNEL<NB><<<GRID, kThreads, SMEM, ST>>>(__VA_ARGS__);                                   \
  }
#define HB_BN_DISPATCH(KERNEL, B, GRID, SMEM, ST, ...)                                       \
  switch (B) {                                                                               \
    case 0: HB_BN_LAUNCH(KERNEL, 0, GRID, SMEM, ST, __VA_ARGS__) break;                      \
    case 1: HB_BN_LAUNCH(KERNEL, 1, GRID, SMEM, ST, __VA_ARGS__) break;                      \
    case 2: HB_BN_LAUNCH(KERNEL, 2, GRID, SMEM, ST, __VA_ARGS__) break;                      \
    default: HB_BN_LAUNCH(KERNEL, 3, GRID, SMEM, ST, __VA_ARGS__) break;                     \
  }

inline size_t ring_bytes(int tensors) { return (size_t)(ring_depth(tensors) + 1) * tensors * kThreads * 16; }

}  // namespace

extern "C" {

// Stand-alone statistics pass over u [M, C] bf16 -> parts float [*slots][C][2] (capacity hb_bn_stat_slots_max()).
int hb_bn_stats_partials_bf16(const void* u, int M, int C, float* parts, int* slots, void* stream) {
  if (C % 8 != 0 || !slots) return (int)cudaErrorInvalidValue;
  Geo g = Geo::make(C);
  static const int min_rows = env_int("HB_BN_STATS_ROWS", 16);
  const dim3 grid = make_grid(g, M, 1, 4, min_rows);
  *slots = (int)grid.x;
  bn_stats_partials_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)u, M, C, g, parts);
  HB_LAUNCH_CHECK();
  return 0;
}

int hb_bn_stat_slots_max(void) { return HB_NUM_SMS * 8; }

int hb_bn_finalize(const float* const* parts, const int* slots, const float* const* gamma, const float* const* beta,
                   float* const* running_mean, float* const* running_var, long long* const* num_batches_tracked,
                   float* mean, float* rstd, float* scale, float* shift, int B, int C, int C_logical, int M, float eps,
                   float momentum, void* stream) {
  if (B < 1 || B > kMaxBranches || !parts || !slots) return (int)cudaErrorInvalidValue;
  FinalizeParams p{};
  for (int b = 0; b < B; ++b) {
    if (!parts[b] || slots[b] < 1) return (int)cudaErrorInvalidValue;
    p.parts[b] = parts[b];
    p.slots[b] = slots[b];
    p.gamma[b] = gamma ? gamma[b] : nullptr;
    p.beta[b] = beta ? beta[b] : nullptr;
    p.running_mean[b] = running_mean ? running_mean[b] : nullptr;
    p.running_var[b] = running_var ? running_var[b] : nullptr;
    p.num_batches_tracked[b] = num_batches_tracked ? num_batches_tracked[b] : nullptr;
  }
  p.mean = mean; p.rstd = rstd; p.scale = scale; p.shift = shift;
  p.B = B; p.C = C; p.M = M; p.C_logical = C_logical; p.eps = eps; p.momentum = momentum;
  bn_finalize_kernel<<<dim3((C + kFinCh - 1) / kFinCh, B), dim3(kFinCh, kFinLanes), 0, (cudaStream_t)stream>>>(p);
  HB_LAUNCH_CHECK();
  return 0;
}

int hb_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                      float eps, int C, int C_logical, float* scale, float* shift, float* mean, float* rstd,
                      void* stream) {
  bn_eval_affine_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, running_mean, running_var, eps,
                                                                            C, C_logical, scale, shift, mean, rstd);
  HB_LAUNCH_CHECK();
  return 0;
}

int hb_bn_act_fwd_bf16(const void* u0, const void* u1, const void* u2, int B, const float* scale, const float* shift,
                       const void* residual, void* out, int M, int C, int act, float slope, int res_after,
                       float* out_stats, int* out_stat_slots, void* stream) {
  if (C % 8 != 0 || B < 0 || B > kMaxBranches) return (int)cudaErrorInvalidValue;
  FwdParams p{};
  p.br = Branches{{(const __nv_bfloat16*)u0, (const __nv_bfloat16*)u1, (const __nv_bfloat16*)u2}, B};
  p.scale = scale; p.shift = shift; p.residual = (const __nv_bfloat16*)residual; p.out = (__nv_bfloat16*)out;
  p.M = M; p.C = C; p.act = act; p.slope = slope; p.res_after = res_after;
  if (out_stats && !out_stat_slots) return (int)cudaErrorInvalidValue;
  p.out_stats = out_stats;
  Geo g = Geo::make(C);
  static const int per_sm_env = env_int("HB_BN_CAP_FWD", 0);
  static const bool use_occ = env_int("HB_BN_USE_OCC", 1) != 0, occ_debug = env_int("HB_BN_DEBUG", 0) != 0;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = ring_bytes(B + 1);
  // grid = min(4, resident blocks of this instantiation) per SM: registers (launch bound 3) and the ring (48 - 64 KB) decide
#define HB_FWD_GO(NBV, STATS)                                                                              \
  {                                                                                                        \
    static int occ = 0;                                                                                    \
    if (!occ) {                                                                                            \
      if (allow_smem(bn_act_fwd_kernel<NBV, STATS>, 200 * 1024) != cudaSuccess) return (int)cudaErrorInvalidValue; \
      occ = resident_blocks(bn_act_fwd_kernel<NBV, STATS>, smem);                                          \
    }                                                                                                      \
    if (occ_debug) fprintf(stderr, "[hb] bn_act_fwd_kernel<%d,%d> smem %zu: %d resident blocks/SM\n", NBV, (int)STATS, smem, occ); \
    const int fixed = (B + (residual != nullptr) >= 3) ? 3 : 4;                                            \
    const int per_sm = per_sm_env > 0 ? per_sm_env : (use_occ ? (occ < 4 ? occ : 4) : fixed);              \
    const dim3 grid = make_grid(g, M, 1, per_sm);                                                          \
    if (out_stat_slots) *out_stat_slots = (int)grid.x;                                                     \
    bn_act_fwd_kernel<NBV, STATS><<<grid, kThreads, smem, st>>>(p, g);                                     \
  }
#define HB_FWD_CASE(NBV)                                                                                   \
  case NBV:                                                                                                \
    if (out_stats) HB_FWD_GO(NBV, true) else HB_FWD_GO(NBV, false)                                         \
    break;
  switch (B) {
    HB_FWD_CASE(0)
    HB_FWD_CASE(1)
    HB_FWD_CASE(2)
    default:
    HB_FWD_CASE(3)
  }
#undef HB_FWD_GO
#undef HB_FWD_CASE
  HB_LAUNCH_CHECK();
  return 0;
}

// Backward. scratch: double [hb_bn_bwd_scratch_doubles(M, C, B)], no initialisation needed: [1+B][C] final sums followed
// by the per-block partials of the reduction pass. du_b / dres may be NULL when not needed.
// dgamma/dbeta: fp32 [B][C] outputs (optional); gamma_grad_acc / beta_grad_acc: optional HOST arrays of B device pointers
// (entries may be NULL) to fp32 [C_logical] gradient buffers that dgamma_b / dbeta_b are ADDED to.
size_t hb_bn_bwd_scratch_doubles(int M, int C, int B) {
  Geo g = Geo::make(C);
  const dim3 grid = make_grid(g, M, 1, 3);   // upper bound of the reduction pass' row blocks (cap <= 3 per SM)
  return (size_t)(1 + B) * C * (1 + (size_t)grid.x);
}

int hb_bn_act_bwd_bf16(const void* dout, const void* u0, const void* u1, const void* u2, int B, const float* scale,
                       const float* shift, const float* mean, const float* rstd, const void* residual, double* scratch,
                       void* du0, void* du1, void* du2, void* dres, float* dgamma, float* dbeta,
                       float* const* gamma_grad_acc, float* const* beta_grad_acc, int C_logical, int M, int C, int act,
                       float slope, int train, int res_after, void* stream) {
  if (C % 8 != 0 || B < 0 || B > kMaxBranches) return (int)cudaErrorInvalidValue;
  BwdParams p{};
  p.br = Branches{{(const __nv_bfloat16*)u0, (const __nv_bfloat16*)u1, (const __nv_bfloat16*)u2}, B};
  p.dout = (const __nv_bfloat16*)dout; p.residual = (const __nv_bfloat16*)residual;
  p.scale = scale; p.shift = shift; p.mean = mean; p.rstd = rstd;
  p.sums = scratch; p.part = scratch + (size_t)(1 + B) * C;
  p.du[0] = (__nv_bfloat16*)du0; p.du[1] = (__nv_bfloat16*)du1; p.du[2] = (__nv_bfloat16*)du2;
  p.dres = (__nv_bfloat16*)dres;
  p.M = M; p.C = C; p.act = act; p.slope = slope; p.train = train; p.res_after = res_after;
  Geo g = Geo::make(C);
  cudaStream_t st = (cudaStream_t)stream;
  static const int cap_red_env = env_int("HB_BN_CAP_RED", 0), cap_app_env = env_int("HB_BN_CAP_APPLY", 0);
  static const bool use_occ = env_int("HB_BN_USE_OCC", 1) != 0, occ_debug = env_int("HB_BN_DEBUG", 0) != 0;
  // one branch (Darknet / ReXNet / UNet blocks): ~70 registers and a 48 KB ring -> three resident blocks per SM
  int cap_red = cap_red_env > 0 ? cap_red_env : (B <= 1 ? 3 : 2);
  if (cap_red > 3) cap_red = 3;   // hb_bn_bwd_scratch_doubles sizes the partials for <= 3 blocks per SM
  const int cap_app = cap_app_env > 0 ? cap_app_env : (B <= 2 ? 3 : 2);   // further limited by the measured occupancy
  const bool want_params = (dgamma && dbeta) || gamma_grad_acc || beta_grad_acc;
  if (train || want_params) {
    const size_t smem = sizeof(SlabConsts) + kThreads * 8 * sizeof(float) + ring_bytes(B + 2);
    int occ = 1;
    HB_BN_OCC_DISPATCH(bn_act_bwd_reduce_kernel, B, smem, occ)
    if (occ_debug) fprintf(stderr, "[hb] bn_act_bwd_reduce_kernel<%d> smem %zu: %d resident blocks/SM (cap %d)\n", B, smem, occ, cap_red);
    const dim3 grid = make_grid(g, M, 1, (use_occ && occ < cap_red) ? occ : cap_red);
    HB_BN_DISPATCH(bn_act_bwd_reduce_kernel, B, grid, smem, st, p, g)
    HB_LAUNCH_CHECK();
    BwdFinalizeParams f{};
    f.part = p.part; f.sums = p.sums; f.mean = mean; f.rstd = rstd; f.dgamma = dgamma; f.dbeta = dbeta;
    for (int b = 0; b < B; ++b) {
      f.gacc[b] = gamma_grad_acc ? gamma_grad_acc[b] : nullptr;
      f.bacc[b] = beta_grad_acc ? beta_grad_acc[b] : nullptr;
    }
    f.nblocks = (int)grid.x; f.B = B; f.C = C; f.C_logical = C_logical > 0 ? C_logical : C;
    bn_bwd_finalize_kernel<<<(C + kFinCh - 1) / kFinCh, dim3(kFinCh, kFinLanes), 0, st>>>(f);
    HB_LAUNCH_CHECK();
  }
  {
    const size_t smem = sizeof(SlabConsts) + ring_bytes(B + 2);
    int occ = 1;
    HB_BN_OCC_DISPATCH(bn_act_bwd_apply_kernel, B, smem, occ)
    if (occ_debug) fprintf(stderr, "[hb] bn_act_bwd_apply_kernel<%d> smem %zu: %d resident blocks/SM (cap %d)\n", B, smem, occ, cap_app);
    const dim3 grid = make_grid(g, M, 1, (use_occ && occ < cap_app) ? occ : cap_app);
    HB_BN_DISPATCH(bn_act_bwd_apply_kernel, B, grid, smem, st, p, g)
    HB_LAUNCH_CHECK();
  }
  return 0;
}

}  // extern "C"
