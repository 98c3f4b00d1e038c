// Weight-gradient of a 2-D convolution on the sm_100a tensor cores.
//
//   dW[co, r, s, ci] = sum_m dY[m, co] * X[pix(m) + (r, s), ci]        (m over all N*Ho*Wo output pixels)
//
// GEMM view per filter tap: D[M = co][N = ci] += A[co, m] * B[ci, m] with the reduction (K) dimension being the
// output pixels. Both operands are "MN-major" in shared memory (the pixel index is the slow one), which
// tcgen05.mma supports directly for 16-bit types, so dY ([M_total, Cout] row-major) and the im2col view of X
// are loaded by TMA exactly as they sit in HBM - no transposes:
//   A stage = 64 pixels x 128 co   (two 64-wide TMA boxes, 128B-swizzled rows = pixels)
//   B stage = per tap: 64 pixels x Cin-tile (im2col TMA: padding / stride / row wrap handled in hardware)
// One CTA owns (co tile, ci tile, tap group, pixel range); all taps of the group accumulate in separate TMEM
// column ranges so the dY tile is loaded once per tap group. Results are reduced across pixel ranges with
// fp32 red.global.add (or stored directly when there is a single range).
// This replaces cuDNN's wgrad behind autograd for nn.Conv2d in the reference
// (holocron/models/utils.py:71, models/classification/repvgg.py:55-62).
#include <cstdlib>
#include "common.cuh"
#include "tc_common.cuh"
#include "tmap.cuh"

namespace {

using namespace tc;

constexpr int kBKpix = 64;     // pixels (reduction) per stage
constexpr int kThreads = 192;  // producer, MMA, 4 epilogue warps
constexpr int kTmemCols = 512;
constexpr int kChunkBytes = kBKpix * 128;  // one 64px x 64ch box = 8 KiB
constexpr int kABytes = 2 * kChunkBytes;   // 128 co

struct WgradParams {
  int m_total, Ho, Wo, stride, pad, dil, R, S, Cin, Cout;
  int ci_tile;         // Cin tile (<= 256), multiple of 16 after rounding
  int ci_chunks;       // ceil(ci_tile / 64)
  int ci_cols;         // TMEM columns per tap (ci_tile rounded up to 32)
  int taps_per_group;  // taps accumulated concurrently in TMEM
  int num_tap_groups, num_co_tiles, num_ci_tiles, k_splits;
  int kblocks_total;   // ceil(m_total / 64)
  int stages, stage_bytes;
  int use_atomics;     // 0: single pixel range, store; 1: atomics; 2: per-range partials in `ws` (reduced by a 2nd kernel)
  float* dw;           // [Cout, R, S, Cin] fp32
  float* ws;           // [k_splits][Cout*R*S*Cin] fp32 partial sums (mode 2)
  long long dw_elems;
};

__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* acc_full = empty_bar + p.stages;  // [1]
  uint64_t* acc_empty = acc_full + 1;         // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmDY);
    prefetch_tmap(&tmX);
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int RS = p.R * p.S;
  const int num_units = p.num_co_tiles * p.num_ci_tiles * p.num_tap_groups * p.k_splits;
  const int b_tap_bytes = p.ci_chunks * kChunkBytes;

  // unit decode: k_split fastest so neighbouring CTAs share the same filter slab / write target
  auto decode = [&](int unit, int& co_t, int& ci_t, int& tg, int& ks) {
    ks = unit % p.k_splits; unit /= p.k_splits;
    tg = unit % p.num_tap_groups; unit /= p.num_tap_groups;
    ci_t = unit % p.num_ci_tiles; co_t = unit / p.num_ci_tiles;
  };
  auto kb_range = [&](int ks, int& kb0, int& kb1) {
    const int per = (p.kblocks_total + p.k_splits - 1) / p.k_splits;
    kb0 = ks * per;
    kb1 = min(kb0 + per, p.kblocks_total);
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        int co_t, ci_t, tg, ks, kb0, kb1;
        decode(unit, co_t, ci_t, tg, ks);
        kb_range(ks, kb0, kb1);
        const int tap0 = tg * p.taps_per_group;
        const int ntaps = min(p.taps_per_group, RS - tap0);
        const uint32_t tx = kABytes + ntaps * b_tap_bytes;
        for (int kb = kb0; kb < kb1; ++kb) {
          const int m0 = kb * kBKpix;
          const int q0 = m0 % p.Wo, p0 = (m0 / p.Wo) % p.Ho, n0 = m0 / (p.Wo * p.Ho);
          const int base_w = q0 * p.stride - p.pad, base_h = p0 * p.stride - p.pad;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * p.stage_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], tx);
          tma_load_2d(&tmDY, &full_bar[stage], sa, co_t * 128, m0);
          tma_load_2d(&tmDY, &full_bar[stage], sa + kChunkBytes, co_t * 128 + 64, m0);
          for (int t = 0; t < ntaps; ++t) {
            const int tap = tap0 + t, r = tap / p.S, s = tap % p.S;
            uint8_t* sb = sa + kABytes + t * b_tap_bytes;
            for (int c = 0; c < p.ci_chunks; ++c)
              tma_load_im2col_4d(&tmX, &full_bar[stage], sb + c * kChunkBytes, ci_t * p.ci_tile + c * 64, base_w, base_h,
                                 n0, (uint16_t)(s * p.dil), (uint16_t)(r * p.dil));
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const int n_mma = (p.ci_tile + 15) & ~15;
      const uint32_t idesc = make_idesc_bf16(128, n_mma, 1, 1);
      const uint32_t dhi = desc_hi(1024, kLayoutSW128);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        int co_t, ci_t, tg, ks, kb0, kb1;
        decode(unit, co_t, ci_t, tg, ks);
        kb_range(ks, kb0, kb1);
        const int tap0 = tg * p.taps_per_group;
        const int ntaps = min(p.taps_per_group, RS - tap0);
        mbar_wait(acc_empty, (it & 1) ^ 1);
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          // 16 pixels = two 8-row swizzle atoms (SBO 1024 B); 64-channel chunks are LBO = 8 KiB apart
          const uint32_t a_lo = desc_lo(smem_u32(smem + (size_t)stage * p.stage_bytes), kChunkBytes);
          uint32_t b_lo = a_lo + (kABytes >> 4);
          const uint32_t acc0 = kb > kb0 ? 1u : 0u;
          for (int t = 0; t < ntaps; ++t) {
#pragma unroll
            for (int k = 0; k < kBKpix / 16; ++k)
              umma_f16_lh(tmem_base + t * p.ci_cols, a_lo + k * (2048 >> 4), dhi, b_lo + k * (2048 >> 4), dhi, idesc,
                          acc0 | (uint32_t)k);
            b_lo += (uint32_t)b_tap_bytes >> 4;
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(acc_full);
      }
    }
  } else {
    const int quarter = warp & 3;
    int it = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
      int co_t, ci_t, tg, ks, kb0, kb1;
      decode(unit, co_t, ci_t, tg, ks);
      kb_range(ks, kb0, kb1);
      const int tap0 = tg * p.taps_per_group;
      const int ntaps = min(p.taps_per_group, RS - tap0);
      mbar_wait(acc_full, it & 1);
      tc_fence_after();
      const int co = co_t * 128 + quarter * 32 + lane;
      const bool co_ok = co < p.Cout;
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16);
      if (kb1 <= kb0 && p.use_atomics == 2 && co_ok) {
        // empty pixel range: this unit's slice of the partial buffer must still read as zero
        for (int t = 0; t < ntaps; ++t)
          for (int c = 0; c < p.ci_tile && ci_t * p.ci_tile + c < p.Cin; ++c)
            p.ws[(size_t)ks * p.dw_elems + ((size_t)co * RS + tap0 + t) * p.Cin + ci_t * p.ci_tile + c] = 0.f;
      }
      if (kb1 > kb0) {
        for (int t = 0; t < ntaps; ++t) {
          const int tap = tap0 + t;
          for (int c = 0; c < p.ci_tile; c += 16) {
            uint32_t v[16];
            tmem_ld_x16(tbase + t * p.ci_cols + c, v);
            tmem_ld_wait();
            const int ci = ci_t * p.ci_tile + c;
            if (co_ok && ci < p.Cin) {
              float* base = p.use_atomics == 2 ? p.ws + (size_t)ks * p.dw_elems : p.dw;
              float* dst = base + ((size_t)co * RS + tap) * p.Cin + ci;
              const int nvalid = min(16, p.Cin - ci);
              if (p.use_atomics == 1) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (j < nvalid) atomicAdd(dst + j, __uint_as_float(v[j]));
              } else {
                if (nvalid == 16) {
                  float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    d4[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                        __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (j < nvalid) dst[j] = __uint_as_float(v[j]);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

// dw[i] = sum_k ws[k][i] in a fixed order (deterministic). blockDim = (32, 8): threadIdx.x walks float4 columns,
// threadIdx.y takes the slices k = y, y+8, ...; the 8 partial sums are combined through shared memory in y order.
// (A single thread per column walking up to 148 slices serially took 23 us per launch: latency, not bandwidth.)
// Elements [0, n_first) go to dw, [n_first, n) to dw2 (two gradient tensors filled by one launch; n_first % 4 == 0);
// accumulate != 0: dw += sum instead of dw = sum (gradient accumulation straight into the parameter's .grad storage, what
// autograd's AccumulateGrad would do with one more element-wise kernel per parameter and step).
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw, long long n,
                                                           int k_splits, float* __restrict__ dw2, long long n_first,
                                                           int accumulate) {
  __shared__ float4 part[8][32];
  const long long i4 = ((long long)blockIdx.x * 32 + threadIdx.x) * 4;
  const int y = threadIdx.y;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i4 + 3 < n) {
#pragma unroll 4
    for (int k = y; k < k_splits; k += 8) {
      const float4 v = *reinterpret_cast<const float4*>(ws + (size_t)k * n + i4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  } else if (i4 < n) {
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = y; k < k_splits; k += 8)
      for (int j = 0; j < 4 && i4 + j < n; ++j) t[j] += ws[(size_t)k * n + i4 + j];
    acc = make_float4(t[0], t[1], t[2], t[3]);
  }
  part[y][threadIdx.x] = acc;
  __syncthreads();
  if (y == 0 && i4 < n) {
    float4 r = part[0][threadIdx.x];
#pragma unroll
    for (int j = 1; j < 8; ++j) {
      const float4 v = part[j][threadIdx.x];
      r.x += v.x; r.y += v.y; r.z += v.z; r.w += v.w;
    }
    float* dst = i4 < n_first ? dw + i4 : dw2 + (i4 - n_first);
    if (i4 + 3 < n) {
      if (accumulate) {
        const float4 o = *reinterpret_cast<const float4*>(dst);
        r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
      }
      *reinterpret_cast<float4*>(dst) = r;
    } else {
      const float t[4] = {r.x, r.y, r.z, r.w};
      for (int j = 0; j < 4 && i4 + j < n; ++j) dst[j] = accumulate ? dst[j] + t[j] : t[j];
    }
  }
}

inline void launch_wgrad_reduce(const float* ws, float* dw, long long n, int slices, cudaStream_t st, float* dw2 = nullptr,
                                long long n_first = -1, int accumulate = 0) {
  const unsigned blocks = (unsigned)((n / 4 + 32) / 32);
  if (n_first < 0 || !dw2) { n_first = n; dw2 = dw; }
  wgrad_reduce_kernel<<<blocks, dim3(32, 8), 0, st>>>(ws, dw, n, slices, dw2, n_first, accumulate);
}

struct WgradPlan {
  WgradParams p;
  size_t ws_bytes;
};

// shared planning of the decomposition (also used by the workspace-size query)
int plan_wgrad(WgradPlan& plan, int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int dil,
               int num_ctas) {
  WgradParams& p = plan.p;
  const int Ho = (H + 2 * pad - dil * (R - 1) - 1) / stride + 1;
  const int Wo = (W + 2 * pad - dil * (S - 1) - 1) / stride + 1;
  const long long m_ll = (long long)N * Ho * Wo;
  if (Ho <= 0 || Wo <= 0 || m_ll > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
  p.m_total = (int)m_ll; p.Ho = Ho; p.Wo = Wo; p.stride = stride; p.pad = pad; p.dil = dil;
  p.R = R; p.S = S; p.Cin = Cin; p.Cout = Cout;
  const int RS = R * S;
  int ci_tile = Cin;
  if (Cin > 256) {
    ci_tile = 256;
    for (int c = 256; c >= 64; c -= 64) if (Cin % c == 0) { ci_tile = c; break; }
  }
  p.ci_tile = ci_tile;
  p.ci_chunks = (ci_tile + 63) / 64;
  p.ci_cols = (ci_tile + 15) & ~15;
  p.taps_per_group = kTmemCols / p.ci_cols;
  if (p.taps_per_group > RS) p.taps_per_group = RS;
  auto stage_bytes_for = [&](int taps) { return kABytes + taps * p.ci_chunks * kChunkBytes; };
  while (p.taps_per_group > 1 && 2 * stage_bytes_for(p.taps_per_group) > 200 * 1024) --p.taps_per_group;
  p.stage_bytes = stage_bytes_for(p.taps_per_group);
  p.stages = (200 * 1024) / p.stage_bytes;
  if (p.stages > 6) p.stages = 6;
  if (p.stages < 2) return (int)cudaErrorInvalidValue;
  p.num_tap_groups = (RS + p.taps_per_group - 1) / p.taps_per_group;
  p.num_co_tiles = (Cout + 127) / 128;
  p.num_ci_tiles = (Cin + ci_tile - 1) / ci_tile;
  p.kblocks_total = (p.m_total + kBKpix - 1) / kBKpix;
  const int base_units = p.num_co_tiles * p.num_ci_tiles * p.num_tap_groups;
  const int ctas = num_ctas > 0 ? num_ctas : HB_NUM_SMS;
  int k_splits = (2 * ctas + base_units - 1) / base_units;   // aim at ~2 units per CTA
  if (base_units >= ctas) k_splits = 1;
  const int max_splits = (p.kblocks_total + 7) / 8;          // at least 8 K blocks (512 pixels) per unit
  if (k_splits > max_splits) k_splits = max_splits;
  if (k_splits < 1) k_splits = 1;
  p.k_splits = k_splits;
  p.dw_elems = (long long)Cout * RS * Cin;
  plan.ws_bytes = k_splits > 1 ? (size_t)k_splits * p.dw_elems * sizeof(float) : 0;
  return 0;
}

}  // namespace

// conv_wgrad_rows.cu: row-window variant for stride-1 3x3 layers with few channels
size_t hb_wgrad_rows_workspace_bytes(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int dil,
                                     int num_ctas, int has_b1);
int hb_wgrad_rows_try(const void* x, const void* dy, const void* dy1, float* ws, size_t ws_bytes, int N, int H, int W, int Cin,
                      int Cout, int num_ctas, cudaStream_t stream, int* slices_out);

extern "C" {

// Bytes of fp32 scratch hb_conv2d_wgrad_bf16 wants for this shape (0 when a single pixel range is used). With a
// workspace the per-range partial sums are written with plain stores and reduced in a fixed order (deterministic);
// without one they are accumulated with fp32 atomics.
size_t hb_conv2d_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int dil,
                                       int num_ctas) {
  WgradPlan plan{};
  if (plan_wgrad(plan, N, H, W, Cin, Cout, R, S, stride, pad, dil, num_ctas)) return 0;
  const size_t rows = hb_wgrad_rows_workspace_bytes(N, H, W, Cin, Cout, R, S, stride, pad, dil, num_ctas, 0);
  return rows > plan.ws_bytes ? rows : plan.ws_bytes;
}

// dW (fp32, [Cout,R,S,Cin]) = wgrad(x [N,H,W,Cin] bf16, dy [N,Ho,Wo,Cout] bf16). Overwrites dW.
// Requirements: Cin % 8 == 0, Cout % 8 == 0, 16-byte aligned pointers.
static int wgrad_impl(const void* x, const void* dy, float* dw, float* workspace, size_t workspace_bytes, int N, int H,
                      int W, int Cin, int Cout, int R, int S, int stride, int pad, int dil, int num_ctas, void* stream,
                      int accumulate) {
  if (Cin % 8 != 0 || Cout % 8 != 0) return (int)cudaErrorInvalidValue;
  if (!hb::aligned16(x) || !hb::aligned16(dy) || !hb::aligned16(dw)) return (int)cudaErrorMisalignedAddress;
  cudaStream_t st = (cudaStream_t)stream;
  static const bool rows_enabled = getenv("HB_DISABLE_WGRAD_ROWS") == nullptr;
  if (rows_enabled && R == 3 && S == 3 && stride == 1 && pad == 1 && dil == 1 && workspace && hb::aligned16(workspace)) {
    int slices = 0;
    const int rc = hb_wgrad_rows_try(x, dy, nullptr, workspace, workspace_bytes, N, H, W, Cin, Cout, num_ctas, st, &slices);
    if (rc == 0) {
      const long long n = (long long)Cout * 9 * Cin;
      launch_wgrad_reduce(workspace, dw, n, slices, st, nullptr, -1, accumulate);
      HB_LAUNCH_CHECK();
      return 0;
    }
    if (rc == -2) return (int)cudaErrorLaunchFailure;
  }
  WgradPlan plan{};
  if (int rc = plan_wgrad(plan, N, H, W, Cin, Cout, R, S, stride, pad, dil, num_ctas)) return rc;
  WgradParams& p = plan.p;
  // accumulation happens in the fixed-order reduction kernel: it needs the workspace path (more than one pixel range)
  if (accumulate && !(p.k_splits > 1 && workspace && workspace_bytes >= plan.ws_bytes && hb::aligned16(workspace)))
    return (int)cudaErrorNotSupported;
  const int RS = R * S;
  const int k_splits = p.k_splits;
  const int base_units = p.num_co_tiles * p.num_ci_tiles * p.num_tap_groups;
  const int ctas = num_ctas > 0 ? num_ctas : HB_NUM_SMS;
  p.dw = dw;
  p.ws = workspace;
  if (k_splits == 1) {
    p.use_atomics = 0;
  } else if (workspace && workspace_bytes >= plan.ws_bytes && hb::aligned16(workspace)) {
    p.use_atomics = 2;
  } else {
    p.use_atomics = 1;
    cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)Cout * RS * Cin * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
  }

  CUtensorMap tmDY, tmX;
  {
    uint64_t dims[2] = {(uint64_t)Cout, (uint64_t)p.m_total};
    uint64_t strides[1] = {(uint64_t)Cout * 2};
    uint32_t box[2] = {64, (uint32_t)kBKpix};
    int rc = tmap::encode_tiled_bf16(&tmDY, dy, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = tmap::encode_im2col_bf16(&tmX, x, N, H, W, Cin, pad, pad, R, S, dil, stride, 64, kBKpix,
                                  CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  const size_t smem_bytes = (size_t)p.stages * p.stage_bytes + (2 * p.stages + 2) * sizeof(uint64_t) + 16 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int num_units = base_units * k_splits;
  int grid = ctas < num_units ? ctas : num_units;
  conv_wgrad_kernel<<<grid, kThreads, smem_bytes, st>>>(tmDY, tmX, p);
  HB_LAUNCH_CHECK();
  if (p.use_atomics == 2) {
    const long long n = p.dw_elems;
    launch_wgrad_reduce(workspace, dw, n, k_splits, st, nullptr, -1, accumulate);
    HB_LAUNCH_CHECK();
  }
  return 0;
}

int hb_conv2d_wgrad_bf16(const void* x, const void* dy, float* dw, float* workspace, size_t workspace_bytes, int N, int H,
                         int W, int Cin, int Cout, int R, int S, int stride, int pad, int dil, int num_ctas, void* stream) {
  return wgrad_impl(x, dy, dw, workspace, workspace_bytes, N, H, W, Cin, Cout, R, S, stride, pad, dil, num_ctas, stream, 0);
}

// dW += wgrad(x, dy): the fixed-order reduction adds onto the existing contents of dw (e.g. the parameter's .grad view in
// the flat gradient bucket). Returns cudaErrorNotSupported (801) without touching dw when the shape runs as a single
// pixel range (no reduction pass to fold the addition into): compute into a scratch tensor and add then.
int hb_conv2d_wgrad_acc_bf16(const void* x, const void* dy, float* dw, float* workspace, size_t workspace_bytes, int N, int H,
                             int W, int Cin, int Cout, int R, int S, int stride, int pad, int dil, int num_ctas,
                             void* stream) {
  return wgrad_impl(x, dy, dw, workspace, workspace_bytes, N, H, W, Cin, Cout, R, S, stride, pad, dil, num_ctas, stream, 1);
}

// Both weight gradients of a stride-1 RepVGG block (3x3 pad-1 branch and 1x1 branch over the same input,
// models/classification/repvgg.py:55-73) in one pass over x: dw = [dW3 (Cout,3,3,Cin) | dW1 (Cout,Cin)] fp32, overwritten.
// workspace: hb_repvgg_wgrad_workspace_bytes(...) bytes; a size of 0 means the shape does not fit the row-window scheme
// (call hb_conv2d_wgrad_bf16 twice then); hb_repvgg_wgrad_bf16 returns cudaErrorNotSupported in that case.
size_t hb_repvgg_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout, int num_ctas) {
  return hb_wgrad_rows_workspace_bytes(N, H, W, Cin, Cout, 3, 3, 1, 1, 1, num_ctas, 1);
}

static int repvgg_wgrad_impl(const void* x, const void* dy3, const void* dy1, float* dw, float* dw1, float* workspace,
                             size_t workspace_bytes, int N, int H, int W, int Cin, int Cout, int num_ctas, void* stream,
                             int accumulate) {
  if (Cin % 8 != 0 || Cout % 8 != 0) return (int)cudaErrorInvalidValue;
  if (!hb::aligned16(x) || !hb::aligned16(dy3) || !hb::aligned16(dy1) || !hb::aligned16(dw) || !hb::aligned16(workspace) ||
      !hb::aligned16(dw1))
    return (int)cudaErrorMisalignedAddress;
  cudaStream_t st = (cudaStream_t)stream;
  int slices = 0;
  const int rc = hb_wgrad_rows_try(x, dy3, dy1, workspace, workspace_bytes, N, H, W, Cin, Cout, num_ctas, st, &slices);
  if (rc == -1) return (int)cudaErrorNotSupported;
  if (rc != 0) return (int)cudaErrorLaunchFailure;
  const long long n = (long long)Cout * 10 * Cin;
  launch_wgrad_reduce(workspace, dw, n, slices, st, dw1, dw1 ? (long long)Cout * 9 * Cin : -1, accumulate);
  HB_LAUNCH_CHECK();
  return 0;
}

int hb_repvgg_wgrad_bf16(const void* x, const void* dy3, const void* dy1, float* dw, float* workspace, size_t workspace_bytes,
                         int N, int H, int W, int Cin, int Cout, int num_ctas, void* stream) {
  return repvgg_wgrad_impl(x, dy3, dy1, dw, nullptr, workspace, workspace_bytes, N, H, W, Cin, Cout, num_ctas, stream, 0);
}

// dW3 [Cout,3,3,Cin] += ..., dW1 [Cout,Cin] += ...: the two gradients ADDED to separate destination buffers (the .grad
// storage of the two branch filters), one pass over x, one reduction launch.
int hb_repvgg_wgrad_acc_bf16(const void* x, const void* dy3, const void* dy1, float* dw3, float* dw1, float* workspace,
                             size_t workspace_bytes, int N, int H, int W, int Cin, int Cout, int num_ctas, void* stream) {
  if (!dw1) return (int)cudaErrorInvalidValue;
  return repvgg_wgrad_impl(x, dy3, dy1, dw3, dw1, workspace, workspace_bytes, N, H, W, Cin, Cout, num_ctas, stream, 1);
}

}  // extern "C"
