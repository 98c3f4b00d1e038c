// Implicit-GEMM 2-D convolution (forward; also the data-gradient pass with transformed weights) on
// the sm_100a tensor cores: NHWC bf16 activations, KRSC bf16 filters, fp32 accumulation in TMEM.
//
// GEMM view:  Y[m, co] = sum_{r,s,ci} X[pix(m) + (r,s), ci] * Wt[co, r, s, ci]
//   M = N*Ho*Wo output pixels (tile 128 = the 128 TMEM lanes), N = Cout (tile BN <= 256 TMEM columns),
//   K = R*S*Cin walked tap by tap in 64-channel blocks.
// Replaces the cuDNN calls behind nn.Conv2d in holocron.models.utils.conv_sequence
// (reference holocron/models/utils.py:28-86) and RepBlock (models/classification/repvgg.py:55-73).
//
// Pipeline (one persistent CTA per SM, 320 threads):
//   warp 0      TMA producer: im2col-mode loads of the activation tile (hardware handles padding, stride,
//               row/image wrap; out-of-range channels are zero-filled) + tiled loads of the filter slab,
//               both landing 128B-swizzled in a multi-stage smem ring (mbarrier complete_tx).
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma (M=128, N=BN, K=16) x4 per stage,
//               tcgen05.commit releases the smem stage / publishes the accumulator.
//   warps 2-9   two epilogue warpgroups taking alternate tiles (one per TMEM accumulator): tcgen05.ld the accumulator
//               (double-buffered in TMEM, so the epilogue of tile i
//               overlaps the MMAs of tile i+1), fuse bias / activation, stage the bf16 tile in shared memory
//               (bank-conflict-free padded rows) and write it out with fully coalesced 128-bit stores
//               (+ residual add in that pass). The first version stored one 96-byte row per thread straight
//               from registers: 32 L1TEX wavefronts per store instruction made L1TEX the busiest unit
//               (profiles/r01_conv_fprop_ncu_full.md).
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "tmap.cuh"

namespace {

using namespace tc;

constexpr int kBM = 128;          // output pixels per tile
constexpr int kBK = 64;           // channels per K block (one 128-byte swizzle row)
constexpr int kUmmaK = 16;        // K per tcgen05.mma for 16-bit inputs
constexpr int kThreads = 320;     // producer warp, MMA warp, 2 x 4 epilogue warps (alternate tiles)
constexpr int kTmemCols = 512;    // two accumulators of up to 256 columns
constexpr int kABytes = kBM * kBK * 2;  // 16 KiB

struct FpropParams {
  int m_total;      // N*Ho*Wo
  int Ho, Wo;
  int stride, pad_h, pad_w, dil;
  int R, S;
  int Cin, Cout;
  int BN;           // Cout tile
  int num_m_tiles, num_n_tiles;
  int cblocks;      // ceil(Cin / 64)
  int stages;
  int b_stage_bytes;  // BN*128 rounded up to 1024
  int out_pitch;      // bytes per row of the smem output staging tile (BN*2 + 16)
  int a_mode;         // 0: plain 2-D [M, C] matrix (1x1 s1 p0), 1: im2col
  int act;            // 0 none, 1 relu
  __nv_bfloat16* y;
  const float* bias;              // [Cout] or null
  const __nv_bfloat16* residual;  // [M, Cout] or null (same addressing as y)
  // output addressing: dense rows (scatter == 0) or output pixel (n, i, j) of the Ho x Wo grid written to pixel
  // (i*o_step + o_a, j*o_step + o_b) of an OH x OW image (the parity classes of a strided data gradient)
  int scatter, OH, OW, o_step, o_a, o_b;
};

__global__ void __launch_bounds__(kThreads, 1)
conv_fprop_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FpropParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A | B] then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = kABytes + p.b_stage_bytes;
  // per epilogue group: [128][out_pitch] output staging tile + [128] row offset table
  uint8_t* sout0 = smem + (size_t)p.stages * stage_bytes;
  const size_t sout_bytes = (((size_t)kBM * p.out_pitch + 15) & ~(size_t)15) + kBM * sizeof(long long);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sout0 + 2 * sout_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tmem_full = empty_bar + p.stages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, kTmemCols); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int kblocks = p.R * p.S * p.cblocks;
  const uint32_t tx_bytes = kABytes + p.BN * 128;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.num_n_tiles, n_tile = tile % p.num_n_tiles;
        const int m0 = m_tile * kBM;
        const int q0 = m0 % p.Wo, p0 = (m0 / p.Wo) % p.Ho, n0 = m0 / (p.Wo * p.Ho);
        const int base_w = q0 * p.stride - p.pad_w, base_h = p0 * p.stride - p.pad_h;
        for (int tap = 0; tap < p.R * p.S; ++tap) {
          const int r = tap / p.S, s = tap % p.S;
          for (int cb = 0; cb < p.cblocks; ++cb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + (size_t)stage * stage_bytes;
            uint8_t* sb = sa + kABytes;
            mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            if (p.a_mode == 1)
              tma_load_im2col_4d(&tmA, &full_bar[stage], sa, cb * kBK, base_w, base_h, n0,
                                 (uint16_t)(s * p.dil), (uint16_t)(r * p.dil));
            else
              tma_load_2d(&tmA, &full_bar[stage], sa, cb * kBK, m0);
            tma_load_3d(&tmB, &full_bar[stage], sb, cb * kBK, tap, n_tile * p.BN);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kBM, p.BN, 0, 0);
      const uint32_t dhi = desc_hi(1024, kLayoutSW128);
      const uint32_t a_lo0 = desc_lo(smem_u32(smem), 16);
      const uint32_t stage_lo = (uint32_t)stage_bytes >> 4;
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * stage_lo;
          const uint32_t b_lo = a_lo + (kABytes >> 4);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k)
            umma_f16_lh(d_tmem, a_lo + 2 * k, dhi, b_lo + 2 * k, dhi, idesc, (uint32_t)(kb | k));
          umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs have read it
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);      // accumulator complete
      }
    }
  } else {
    // ================= epilogue (warps 2..5 take the even tiles of this CTA, warps 6..9 the odd ones) =================
    // One tile's epilogue is a chain of latencies (tcgen05.ld -> convert -> st.shared -> barrier -> ld.shared ->
    // st.global); two groups on the two TMEM accumulators overlap two of those chains.
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int group = (warp - 2) >> 2;
    const int et = (threadIdx.x - 64) & 127;   // 0..127 inside the group
    uint8_t* sout = sout0 + (size_t)group * sout_bytes;
    long long* row_off = reinterpret_cast<long long*>(sout + (((size_t)kBM * p.out_pitch + 15) & ~(size_t)15));
    for (int tile = blockIdx.x + group * gridDim.x, it = group; tile < num_tiles; tile += 2 * gridDim.x, it += 2) {
      const int m_tile = tile / p.num_n_tiles, n_tile = tile % p.num_n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row_in_tile = quarter * 32 + lane;
      const int col_base = n_tile * p.BN;
      const uint32_t taddr = tmem_base + acc * 256 + ((uint32_t)(quarter * 32) << 16);
      const int rows_valid = min(kBM, p.m_total - m_tile * kBM);
      const int ncols_valid = min(p.BN, p.Cout - col_base);          // multiple of 16
      uint8_t* srow = sout + (size_t)row_in_tile * p.out_pitch;
      // column groups of <= 64 accumulator columns go through a small fixed-size staging tile (keeps the smem for
      // pipeline stages whatever BN is); each group is written out as 128-byte row segments
      for (int g0 = 0; g0 < p.BN; g0 += 64) {
        const int gw = min(64, p.BN - g0);
        if (group == 0) asm volatile("bar.sync 1, 128;" ::: "memory");   // staging tile free
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        if (g0 == 0) {
          // element offset of this thread's output row (read by every thread in the copy-out below)
          const long long m = (long long)m_tile * kBM + et;
          if (!p.scatter) {
            row_off[et] = m * p.Cout;
          } else {
            const int j = (int)(m % p.Wo), i = (int)((m / p.Wo) % p.Ho), n = (int)(m / ((long long)p.Wo * p.Ho));
            row_off[et] = (((long long)n * p.OH + (long long)i * p.o_step + p.o_a) * p.OW + (long long)j * p.o_step + p.o_b) * p.Cout;
          }
        }
        for (int c = 0; c < gw; c += 32) {
          uint32_t v[32];
          const bool two = (c + 16) < gw;
          tmem_ld_x16(taddr + g0 + c, v);
          if (two) tmem_ld_x16(taddr + g0 + c + 16, v + 16);
          tmem_ld_wait();
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h == 0 || two) {
              float f[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[h * 16 + j]);
              const int col = col_base + g0 + c + h * 16;
              if (p.bias && col < p.Cout) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] += __ldg(p.bias + col + j);
              }
              if (p.act == 1 && !p.residual) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.0f);
              }
              uint4 o[2];
              __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
              for (int j = 0; j < 8; ++j) ob[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
              uint4* sp = reinterpret_cast<uint4*>(srow + (c + h * 16) * 2);
              sp[0] = o[0];
              sp[1] = o[1];
            }
          }
        }
        if (g0 + 64 >= p.BN) {   // last TMEM read of this accumulator: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (group == 0) asm volatile("bar.sync 1, 128;" ::: "memory");   // staged group visible to the whole group
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        const int chunks_per_row = gw / 8;
        const int total_chunks = rows_valid * chunks_per_row;
        int r = et / chunks_per_row, c8 = et - r * chunks_per_row;
        const int dr = 128 / chunks_per_row, dc = 128 - dr * chunks_per_row;
        if (!p.residual) {
          for (int ch = et; ch < total_chunks; ch += 128) {
            if (g0 + c8 * 8 < ncols_valid) {
              const uint4 val = *reinterpret_cast<const uint4*>(sout + (size_t)r * p.out_pitch + c8 * 16);
              *reinterpret_cast<uint4*>(p.y + (size_t)row_off[r] + col_base + g0 + c8 * 8) = val;
            }
            r += dr; c8 += dc;
            if (c8 >= chunks_per_row) { c8 -= chunks_per_row; ++r; }
          }
        } else {
          // residual add: the global loads of 4 trips are issued before the first one is consumed (one dependent
          // global load per trip made this pass latency-bound: +100 % on the 1x1 data-gradient launches)
          for (int ch = et; ch < total_chunks; ch += 4 * 128) {
            size_t off[4];
            int sidx[4];
            uint4 rv[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              ok[u] = (ch + u * 128 < total_chunks) && (g0 + c8 * 8 < ncols_valid);
              sidx[u] = r * p.out_pitch + c8 * 16;
              off[u] = ok[u] ? (size_t)row_off[r] + col_base + g0 + c8 * 8 : 0;
              if (ok[u]) rv[u] = *reinterpret_cast<const uint4*>(p.residual + off[u]);
              r += dr; c8 += dc;
              if (c8 >= chunks_per_row) { c8 -= chunks_per_row; ++r; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (!ok[u]) continue;
              uint4 val = *reinterpret_cast<const uint4*>(sout + sidx[u]);
              __nv_bfloat162* a = reinterpret_cast<__nv_bfloat162*>(&val);
              const __nv_bfloat162* b = reinterpret_cast<const __nv_bfloat162*>(&rv[u]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 fa = __bfloat1622float2(a[j]), fb = __bfloat1622float2(b[j]);
                fa.x += fb.x; fa.y += fb.y;
                if (p.act == 1) { fa.x = fmaxf(fa.x, 0.f); fa.y = fmaxf(fa.y, 0.f); }
                a[j] = __floats2bfloat162_rn(fa.x, fa.y);
              }
              *reinterpret_cast<uint4*>(p.y + off[u]) = val;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

}  // namespace

// conv_rows.cu: shared-memory-reuse kernel for stride-1 3x3 layers whose filter fits in shared memory
int hb_conv_rows_try(const void* x, const void* w, void* y, const float* bias, const void* residual, int N, int H, int W,
                     int Cin, int Cout, int act, int num_ctas, cudaStream_t stream, int nextra = 0,
                     const void* const* xe = nullptr, const void* const* we = nullptr);

namespace {

struct FpropArgs {
  const void* x; const void* w; void* y; const float* bias; const void* residual;
  int N, H, W, Cin, Cout, R, S, stride;
  int pad_h, pad_w, pad_after_h, pad_after_w, dil, act, num_ctas;
  int Ho, Wo;                      // output grid walked by the GEMM rows
  int scatter, OH, OW, o_step, o_a, o_b;
  cudaStream_t stream;
};

int fprop_launch(const FpropArgs& a) {
  const int Cin = a.Cin, Cout = a.Cout, R = a.R, S = a.S;
  const long long m_total_ll = (long long)a.N * a.Ho * a.Wo;
  if (a.Ho <= 0 || a.Wo <= 0 || m_total_ll <= 0 || m_total_ll > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
  FpropParams p{};
  p.m_total = (int)m_total_ll;
  p.Ho = a.Ho; p.Wo = a.Wo;
  p.stride = a.stride; p.pad_h = a.pad_h; p.pad_w = a.pad_w; p.dil = a.dil;
  p.R = R; p.S = S; p.Cin = Cin; p.Cout = Cout;
  p.scatter = a.scatter; p.OH = a.OH; p.OW = a.OW; p.o_step = a.o_step; p.o_a = a.o_a; p.o_b = a.o_b;
  // Cout tile: whole Cout when it fits the 256 accumulator columns, else the largest multiple of 16
  // <= 256 that divides Cout (falls back to 256 with a masked tail).
  int BN = Cout;
  if (Cout > 256) {
    BN = 256;
    for (int c = 256; c >= 64; c -= 16) if (Cout % c == 0) { BN = c; break; }
  }
  p.BN = BN;
  p.num_m_tiles = (p.m_total + kBM - 1) / kBM;
  p.num_n_tiles = (Cout + BN - 1) / BN;
  p.cblocks = (Cin + kBK - 1) / kBK;
  p.b_stage_bytes = ((BN * 128) + 1023) & ~1023;
  p.out_pitch = (BN < 64 ? BN : 64) * 2 + 16;
  const int out_bytes = (2 * (((kBM * p.out_pitch + 15) & ~15) + kBM * 8) + 1023) & ~1023;   // per group: staging + offsets
  const int stage_bytes = kABytes + p.b_stage_bytes;
  int stages = (204 * 1024 - out_bytes) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return (int)cudaErrorInvalidValue;
  p.stages = stages;
  p.a_mode = (R == 1 && S == 1 && a.stride == 1 && a.pad_h == 0 && a.pad_w == 0 && a.pad_after_h == 0 &&
              a.pad_after_w == 0) ? 0 : 1;
  p.act = a.act;
  p.y = (__nv_bfloat16*)a.y;
  p.bias = a.bias;
  p.residual = (const __nv_bfloat16*)a.residual;

  CUtensorMap tmA, tmB;
  int rc;
  if (p.a_mode == 0) {
    uint64_t dims[2] = {(uint64_t)Cin, (uint64_t)p.m_total};
    uint64_t strides[1] = {(uint64_t)Cin * 2};
    uint32_t box[2] = {kBK, kBM};
    rc = tmap::encode_tiled_bf16(&tmA, a.x, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
  } else {
    rc = tmap::encode_im2col_bf16(&tmA, a.x, a.N, a.H, a.W, Cin, a.pad_h, a.pad_w, R, S, a.dil, a.stride, kBK, kBM,
                                  CU_TENSOR_MAP_SWIZZLE_128B, a.pad_after_h, a.pad_after_w);
  }
  if (rc) return rc;
  {
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)(R * S), (uint64_t)Cout};
    uint64_t strides[2] = {(uint64_t)Cin * 2, (uint64_t)R * S * Cin * 2};
    uint32_t box[3] = {kBK, 1, (uint32_t)BN};
    rc = tmap::encode_tiled_bf16(&tmB, a.w, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }

  const size_t smem_bytes = (size_t)stages * stage_bytes + out_bytes + (2 * stages + 4) * sizeof(uint64_t) + 16 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_fprop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  int grid = a.num_ctas > 0 ? a.num_ctas : HB_NUM_SMS;
  if (grid > num_tiles) grid = num_tiles;
  conv_fprop_kernel<<<grid, kThreads, smem_bytes, a.stream>>>(tmA, tmB, p);
  HB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

// Forward convolution, NHWC bf16.  x: [N,H,W,Cin]  w: [Cout,R,S,Cin]  y: [N,Ho,Wo,Cout]
//   bias: fp32 [Cout] or NULL; residual: bf16 [N,Ho,Wo,Cout] or NULL; act: 0 none, 1 relu.
// Requirements: Cin % 8 == 0, Cout % 16 == 0, all pointers 16-byte aligned.
int hb_conv2d_fprop_bf16(const void* x, const void* w, void* y, const float* bias, const void* residual, int N, int H,
                         int W, int Cin, int Cout, int R, int S, int stride, int pad, int dil, int act, int num_ctas,
                         void* stream) {
  if (Cin % 8 != 0 || Cout % 16 != 0) return (int)cudaErrorInvalidValue;
  if (!hb::aligned16(x) || !hb::aligned16(w) || !hb::aligned16(y)) return (int)cudaErrorMisalignedAddress;
  const int Ho = (H + 2 * pad - dil * (R - 1) - 1) / stride + 1;
  const int Wo = (W + 2 * pad - dil * (S - 1) - 1) / stride + 1;
  if (Ho <= 0 || Wo <= 0) return (int)cudaErrorInvalidValue;
  if ((long long)N * Ho * Wo > 0x7fffffffLL) return (int)cudaErrorInvalidValue;

  if (R == 3 && S == 3 && stride == 1 && pad == 1 && dil == 1) {
    static const bool rows_enabled = getenv("HB_DISABLE_CONV_ROWS") == nullptr;
    if (rows_enabled) {
      const int rc = hb_conv_rows_try(x, w, y, bias, residual, N, H, W, Cin, Cout, act, num_ctas, (cudaStream_t)stream);
      if (rc == 0) return 0;
      if (rc == -2) return (int)cudaErrorLaunchFailure;
    }
  }
  FpropArgs a{};
  a.x = x; a.w = w; a.y = y; a.bias = bias; a.residual = residual;
  a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.R = R; a.S = S; a.stride = stride;
  a.pad_h = a.pad_w = a.pad_after_h = a.pad_after_w = pad; a.dil = dil; a.act = act; a.num_ctas = num_ctas;
  a.Ho = Ho; a.Wo = Wo; a.stream = (cudaStream_t)stream;
  return fprop_launch(a);
}

// Data gradient of a stride-2 3x3 pad-1 convolution WITHOUT zero insertion: the four (row, column) parity classes of dx
// are four small stride-1 correlations over dy (1, 2, 2 and 4 taps), each written straight to its sub-grid of dx:
//   dx[2i+a, 2j+b] = sum_{t,u} dy[i+t, j+u] * wcls_ab[t, u]       t < 1+a, u < 1+b
// 9/4 of the multiply-adds per pixel instead of 9, dy read at its own resolution, no H x W scratch tensors.
//   dy: [N,Ho,Wo,C] bf16; wcls: the class filters from hb_pack_dgrad_s2_weights; dx: [N,H,W,Cd] bf16 (every element written).
//   dy1/wd1 (optional): output gradient and [Cd][C] filter of a parallel 1x1 stride-2 branch (RepVGG), whose data gradient
//   only touches class (0,0); it is written first and the 3x3 class accumulates onto it.
int hb_conv2d_dgrad_s2_bf16(const void* dy, const void* wcls, const void* dy1, const void* wd1, void* dx, int N, int H, int W,
                            int Ho, int Wo, int C, int Cd, int num_ctas, void* stream) {
  if (C % 8 != 0 || Cd % 16 != 0) return (int)cudaErrorInvalidValue;
  if (!hb::aligned16(dy) || !hb::aligned16(wcls) || !hb::aligned16(dx)) return (int)cudaErrorMisalignedAddress;
  if (Ho != (H - 1) / 2 + 1 || Wo != (W - 1) / 2 + 1 || H < 2 || W < 2) return (int)cudaErrorInvalidValue;
  const __nv_bfloat16* wc = (const __nv_bfloat16*)wcls;
  size_t woff = 0;
  for (int a = 0; a < 2; ++a) {
    for (int b = 0; b < 2; ++b) {
      const int Hc = (H - a + 1) / 2, Wc = (W - b + 1) / 2;   // pixels of this parity class
      const int R = 1 + a, S = 1 + b;
      FpropArgs f{};
      f.x = dy; f.w = wc + woff; f.y = dx; f.bias = nullptr; f.residual = nullptr;
      f.N = N; f.H = Ho; f.W = Wo; f.Cin = C; f.Cout = Cd; f.R = R; f.S = S; f.stride = 1;
      f.pad_h = 0; f.pad_w = 0;
      f.pad_after_h = a ? Hc + 1 - Ho : 0;   // 1 when the last odd row reaches dy row Ho (zero), else 0
      f.pad_after_w = b ? Wc + 1 - Wo : 0;
      f.dil = 1; f.act = 0; f.num_ctas = num_ctas;
      f.Ho = Hc; f.Wo = Wc;
      f.scatter = 1; f.OH = H; f.OW = W; f.o_step = 2; f.o_a = a; f.o_b = b;
      f.stream = (cudaStream_t)stream;
      if (a == 0 && b == 0 && dy1 && wd1) {
        FpropArgs g = f;
        g.x = dy1; g.w = wd1;
        if (int rc = fprop_launch(g)) return rc;
        f.residual = dx;   // accumulate onto the 1x1 branch's contribution
      }
      if (Hc > 0 && Wc > 0) {
        if (int rc = fprop_launch(f)) return rc;
      }
      woff += (size_t)Cd * R * S * C;
    }
  }
  return 0;
}

}  // extern "C"
