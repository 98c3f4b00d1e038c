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
#include "holocron_b200.h"

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
  int ksteps_last;  // 16-channel UMMA steps of the last channel block (Cin = 48 -> 3 instead of 4 zero-padded ones)
  // extra K blocks issued after the R*S*cblocks main ones:
  //   e_mode 1: a second source xe [M, Ce] (rows = the output rows) with filter we [Cout,1,1,Ce] accumulated into the
  //             SAME accumulator (K extension: dX = dgrad3x3(dY3) + dgrad1x1(dY1) in one kernel);
  //   e_mode 2: the centre tap of the main source once more with a second filter w2 [Cout,1,1,Cin] into a SECOND
  //             accumulator (dual output: y = conv RxS, y2 = conv 1x1 of the same input, same stride - RepVGG forward).
  int e_mode, e_cblocks, e_ksteps_last;
  int nout;         // 1, or 2 in dual mode
  int stages;
  int b_stage_bytes;  // BN*128 rounded up to 1024
  int out_pitch;      // bytes per row of the smem output staging tile (BN*2 + 16)
  int a_mode;         // 0: plain 2-D [M, C] matrix (1x1 s1 p0), 1: im2col
  int act;            // 0 none, 1 relu
  __nv_bfloat16* y;
  __nv_bfloat16* y2;              // dual mode: second output (same addressing as y)
  // optional per-channel statistics of the bf16 OUTPUT (what the BatchNorm that follows normalises): float
  // [slots][Cout][2] = (sum, sum of squares) partials, slot = (blockIdx.x / num_n_tiles) * 2 + epilogue group; every
  // (slot, channel) is written exactly once and the consumer adds the slots in a fixed order (deterministic)
  float* stats;
  float* stats2;
  const float* bias;              // [Cout] or null
  const __nv_bfloat16* residual;  // [M, Cout] or null (same addressing as y)
  // patch normalisation in the epilogue (NormConv2d, reference nn/functional.py:322-413): with the per-patch statistics of
  // the im2col rows, y = rstd[m] * (acc - mean[m] * wsum[co]) (+ bias) == sum_k ((patch_k - mean) * rstd) * w[co, k]
  const float* norm_mean;         // [M] or null
  const float* norm_rstd;         // [M]
  const float* norm_wsum;         // [Cout] = sum over (r, s, ci) of the bf16 filter
  // output addressing: dense rows (scatter == 0) or output pixel (n, i, j) of the Ho x Wo grid written to pixel
  // (i*o_step + o_a, j*o_step + o_b) of an OH x OW image (the parity classes of a strided data gradient)
  int scatter, OH, OW, o_step, o_a, o_b;
};

// kStats: the epilogue also accumulates the output-column statistics (separate instantiation: the plain kernel carries
// neither the 16 accumulator registers nor the extra shared-memory pass in its instruction stream)
template <bool kStats>
__global__ void __launch_bounds__(kThreads, 1)
conv_fprop_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, const FpropParams p) {
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
    if (p.e_mode) { prefetch_tmap(&tmB2); if (p.e_mode == 1) prefetch_tmap(&tmA2); }
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_ptr, kTmemCols); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // Tile walk: CTA b always works on Cout tile  b % num_n_tiles  (gridDim.x is a multiple of num_n_tiles) and takes the
  // pixel tiles  b / num_n_tiles + k * (gridDim.x / num_n_tiles): CTAs that run side by side share their A tile in L2, and
  // the per-CTA column statistics cover a fixed channel range.
  const int n_tile = blockIdx.x % p.num_n_tiles;
  const int m_first = blockIdx.x / p.num_n_tiles, m_step = gridDim.x / p.num_n_tiles;
  const int kblocks = p.R * p.S * p.cblocks;
  const uint32_t tx_bytes = kABytes + p.BN * 128;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int m_tile = m_first; m_tile < p.num_m_tiles; m_tile += m_step) {
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
        for (int cb = 0; cb < p.e_cblocks; ++cb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          uint8_t* sb = sa + kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          if (p.e_mode == 1)
            tma_load_2d(&tmA2, &full_bar[stage], sa, cb * kBK, m0);
          else if (p.a_mode == 1)
            tma_load_im2col_4d(&tmA, &full_bar[stage], sa, cb * kBK, base_w, base_h, n0,
                               (uint16_t)((p.S / 2) * p.dil), (uint16_t)((p.R / 2) * p.dil));
          else
            tma_load_2d(&tmA, &full_bar[stage], sa, cb * kBK, m0);
          tma_load_3d(&tmB2, &full_bar[stage], sb, cb * kBK, 0, n_tile * p.BN);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
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
      for (int m_tile = m_first; m_tile < p.num_m_tiles; m_tile += m_step, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        int cb = 0;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * stage_lo;
          const uint32_t b_lo = a_lo + (kABytes >> 4);
          const bool last_cb = ++cb == p.cblocks;
          const int ks = last_cb ? p.ksteps_last : kBK / kUmmaK;
          if (last_cb) cb = 0;
          // fully unrolled with a predicate per step: the single issuing thread must not pay loop overhead per MMA (a
          // runtime-trip-count loop here cost ~20 % on every layer: the kernel is issue-rate sensitive)
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k)
            if (k < ks) umma_f16_lh(d_tmem, a_lo + 2 * k, dhi, b_lo + 2 * k, dhi, idesc, (uint32_t)(kb | k));
          umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs have read it
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        // extra K blocks: same accumulator (e_mode 1) or the second accumulator, BN columns further (e_mode 2)
        const uint32_t d_extra = d_tmem + (p.e_mode == 2 ? (uint32_t)p.BN : 0u);
        for (int ecb = 0; ecb < p.e_cblocks; ++ecb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * stage_lo;
          const uint32_t b_lo = a_lo + (kABytes >> 4);
          const int ks = (ecb == p.e_cblocks - 1) ? p.e_ksteps_last : kBK / kUmmaK;
          const uint32_t acc_first = p.e_mode == 2 ? (uint32_t)ecb : 1u;
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k)
            if (k < ks) umma_f16_lh(d_extra, a_lo + 2 * k, dhi, b_lo + 2 * k, dhi, idesc, acc_first | (uint32_t)k);
          umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);      // accumulator(s) complete
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
    // column groups of <= 64 accumulator columns go through a small fixed-size staging tile (keeps the smem for pipeline
    // stages whatever BN is): gpo groups per output, nout outputs, at most 4 groups in total (host-checked)
    const int gpo = (p.BN + 63) >> 6;
    const int ngroups = p.nout * gpo;
    // column statistics: this thread's running (sum0, sum1, sumsq0, sumsq1) of one column PAIR of each group over a fixed
    // subset of the tile rows (rows rg, rg + rgs, ...), kept in registers across all tiles of the CTA
    float st[4][4];
#pragma unroll
    for (int gi = 0; gi < 4; ++gi) { st[gi][0] = st[gi][1] = st[gi][2] = st[gi][3] = 0.f; }
    const int col_base = n_tile * p.BN;
    const int ncols_valid = min(p.BN, p.Cout - col_base);          // multiple of 16
    for (int m_tile = m_first + group * m_step, it = group; m_tile < p.num_m_tiles; m_tile += 2 * m_step, it += 2) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row_in_tile = quarter * 32 + lane;
      const uint32_t taddr0 = tmem_base + acc * 256 + ((uint32_t)(quarter * 32) << 16);
      const int rows_valid = min(kBM, p.m_total - m_tile * kBM);
      uint8_t* srow = sout + (size_t)row_in_tile * p.out_pitch;
      float nm = 0.f, nr = 1.f;
      if (p.norm_mean) {
        const int m_row = m_tile * kBM + row_in_tile;
        if (m_row < p.m_total) { nm = __ldg(p.norm_mean + m_row); nr = __ldg(p.norm_rstd + m_row); }
      }
#pragma unroll
      for (int gi = 0; gi < 4; ++gi) {
        if (gi >= ngroups) break;
        const int o = gi >= gpo ? 1 : 0;                 // output index (dual mode: 0 = RxS conv, 1 = 1x1 conv)
        const int g0 = (gi - o * gpo) * 64;
        const int gw = min(64, p.BN - g0);
        const uint32_t taddr = taddr0 + (uint32_t)(o * p.BN);
        __nv_bfloat16* yo = o ? p.y2 : p.y;
        const bool first = o == 0;
        if (group == 0) asm volatile("bar.sync 1, 128;" ::: "memory");   // staging tile free
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        if (gi == 0) {
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
              if (first && p.norm_mean && col < p.Cout) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = nr * (f[j] - nm * __ldg(p.norm_wsum + col + j));
              }
              if (first && p.bias && col < p.Cout) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] += __ldg(p.bias + col + j);
              }
              if (first && p.act == 1 && !p.residual) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = hb::relu_nan(f[j]);
              }
              uint4 ov[2];
              __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(ov);
#pragma unroll
              for (int j = 0; j < 8; ++j) ob[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
              uint4* sp = reinterpret_cast<uint4*>(srow + (c + h * 16) * 2);
              sp[0] = ov[0];
              sp[1] = ov[1];
            }
          }
        }
        if (gi == ngroups - 1) {   // last TMEM read of this accumulator set: hand it back to the MMA warp
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
        const __nv_bfloat16* resid = first ? p.residual : nullptr;
        if (!resid) {
          for (int ch = et; ch < total_chunks; ch += 128) {
            if (g0 + c8 * 8 < ncols_valid) {
              const uint4 val = *reinterpret_cast<const uint4*>(sout + (size_t)r * p.out_pitch + c8 * 16);
              *reinterpret_cast<uint4*>(yo + (size_t)row_off[r] + col_base + g0 + c8 * 8) = val;
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
              if (ok[u]) rv[u] = *reinterpret_cast<const uint4*>(resid + off[u]);
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
                if (p.act == 1) { fa.x = hb::relu_nan(fa.x); fa.y = hb::relu_nan(fa.y); }
                a[j] = __floats2bfloat162_rn(fa.x, fa.y);
              }
              *reinterpret_cast<uint4*>(yo + off[u]) = val;
            }
          }
        }
        if (kStats && (o ? (p.stats2 != nullptr) : (p.stats != nullptr))) {
          // per-channel sum / sum of squares of the staged (bf16-rounded) tile: thread = (column pair, row subset)
          const int npairs = gw >> 1, rgs = 128 / npairs;
          const int rg = et / npairs, pr = et - rg * npairs;
          if (rg < rgs && g0 + pr * 2 < ncols_valid) {
            const uint8_t* sp = sout + pr * 4;
            float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
            for (int rr = rg; rr < rows_valid; rr += rgs) {
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(sp + (size_t)rr * p.out_pitch));
              s0 += f.x; s1 += f.y; q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1);
            }
            st[gi][0] += s0; st[gi][1] += s1; st[gi][2] += q0; st[gi][3] += q1;
          }
        }
      }
    }
    // ---- statistics: fold the row subsets in a fixed order and write this (CTA, group)'s partial (every slot is written,
    // zeros included, so the consumer can add all slots without a memset)
    if (kStats && (p.stats || p.stats2)) {
      const int slot = (blockIdx.x / p.num_n_tiles) * 2 + group;
#pragma unroll
      for (int gi = 0; gi < 4; ++gi) {
        if (gi >= ngroups) break;
        const int o = gi >= gpo ? 1 : 0;
        float* so = o ? p.stats2 : p.stats;
        if (!so) continue;
        const int g0 = (gi - o * gpo) * 64;
        const int gw = min(64, p.BN - g0);
        const int npairs = gw >> 1, rgs = 128 / npairs;
        if (group == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        float4* scratch = reinterpret_cast<float4*>(sout);           // [128] float4, 2 KB <= staging tile
        scratch[et] = make_float4(st[gi][0], st[gi][1], st[gi][2], st[gi][3]);
        if (group == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        if (et < gw && g0 + et < ncols_valid) {
          const int pr = et >> 1, hi = et & 1;
          float sv = 0.f, qv = 0.f;
          for (int rg = 0; rg < rgs; ++rg) {
            const float4 v = scratch[rg * npairs + pr];
            sv += hi ? v.y : v.x;
            qv += hi ? v.w : v.z;
          }
          float2* dst = reinterpret_cast<float2*>(so + ((size_t)slot * p.Cout + col_base + g0 + et) * 2);
          *dst = make_float2(sv, qv);
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
                     const void* const* xe = nullptr, const void* const* we = nullptr, float* stats = nullptr,
                     int* stat_slots = nullptr);

namespace {

struct FpropArgs {
  const void* x; const void* w; void* y; const float* bias; const void* residual;
  int N, H, W, Cin, Cout, R, S, stride;
  int pad_h, pad_w, pad_after_h, pad_after_w, dil, act, num_ctas;
  int Ho, Wo;                      // output grid walked by the GEMM rows
  int scatter, OH, OW, o_step, o_a, o_b;
  // K extension (same accumulator): xe [M, Ce] bf16 rows aligned with the output rows, we [Cout,1,1,Ce]
  const void* xe; const void* we; int Ce;
  // dual output: w2 [Cout,1,1,Cin] applied to the centre tap of x -> y2
  const void* w2; void* y2;
  float* stats; float* stats2; int* stat_slots;
  const float* norm_mean; const float* norm_rstd; const float* norm_wsum;
  cudaStream_t stream;
};

int fprop_launch(const FpropArgs& a) {
  const int Cin = a.Cin, Cout = a.Cout, R = a.R, S = a.S;
  const long long m_total_ll = (long long)a.N * a.Ho * a.Wo;
  if (a.Ho <= 0 || a.Wo <= 0 || m_total_ll <= 0 || m_total_ll > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
  const bool dual = a.w2 != nullptr;
  const bool kext = a.xe != nullptr;
  if (dual && kext) return (int)cudaErrorInvalidValue;
  if (dual && (!a.y2 || a.scatter || !(R & 1) || !(S & 1))) return (int)cudaErrorInvalidValue;
  if (kext && (!a.we || a.Ce % 8 != 0 || a.scatter || !hb::aligned16(a.xe) || !hb::aligned16(a.we)))
    return (int)cudaErrorInvalidValue;
  FpropParams p{};
  p.m_total = (int)m_total_ll;
  p.Ho = a.Ho; p.Wo = a.Wo;
  p.stride = a.stride; p.pad_h = a.pad_h; p.pad_w = a.pad_w; p.dil = a.dil;
  p.R = R; p.S = S; p.Cin = Cin; p.Cout = Cout;
  p.scatter = a.scatter; p.OH = a.OH; p.OW = a.OW; p.o_step = a.o_step; p.o_a = a.o_a; p.o_b = a.o_b;
  // Cout tile: whole Cout when it fits the accumulator columns (256, or 128 per output in dual mode), else the largest
  // multiple of 16 below that limit that divides Cout (falls back to the limit with a masked tail).
  const int bn_max = dual ? 128 : 256;
  int BN = Cout;
  if (Cout > bn_max) {
    BN = bn_max;
    for (int c = bn_max; c >= 64; c -= 16) if (Cout % c == 0) { BN = c; break; }
  }
  p.BN = BN;
  p.nout = dual ? 2 : 1;
  p.num_m_tiles = (p.m_total + kBM - 1) / kBM;
  p.num_n_tiles = (Cout + BN - 1) / BN;
  p.cblocks = (Cin + kBK - 1) / kBK;
  p.ksteps_last = ((Cin - (p.cblocks - 1) * kBK) + kUmmaK - 1) / kUmmaK;
  p.e_mode = kext ? 1 : (dual ? 2 : 0);
  const int ce = kext ? a.Ce : (dual ? Cin : 0);
  p.e_cblocks = (ce + kBK - 1) / kBK;
  p.e_ksteps_last = p.e_cblocks ? ((ce - (p.e_cblocks - 1) * kBK) + kUmmaK - 1) / kUmmaK : 0;
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
  p.y2 = (__nv_bfloat16*)a.y2;
  p.stats = a.stats; p.stats2 = dual ? a.stats2 : nullptr;
  p.bias = a.bias;
  p.residual = (const __nv_bfloat16*)a.residual;
  p.norm_mean = a.norm_mean; p.norm_rstd = a.norm_rstd; p.norm_wsum = a.norm_wsum;
  if (p.norm_mean && (!p.norm_rstd || !p.norm_wsum || a.scatter)) return (int)cudaErrorInvalidValue;

  CUtensorMap tmA, tmB, tmA2, tmB2;
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
  tmA2 = tmA; tmB2 = tmB;   // unused slots alias the main maps (never dereferenced by the kernel)
  if (kext) {
    uint64_t dims[2] = {(uint64_t)a.Ce, (uint64_t)p.m_total};
    uint64_t strides[1] = {(uint64_t)a.Ce * 2};
    uint32_t box[2] = {kBK, kBM};
    if ((rc = tmap::encode_tiled_bf16(&tmA2, a.xe, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  }
  if (p.e_mode) {
    uint64_t dims[3] = {(uint64_t)ce, 1, (uint64_t)Cout};
    uint64_t strides[2] = {(uint64_t)ce * 2, (uint64_t)ce * 2};
    uint32_t box[3] = {kBK, 1, (uint32_t)BN};
    if ((rc = tmap::encode_tiled_bf16(&tmB2, kext ? a.we : a.w2, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)))
      return rc;
  }

  const size_t smem_bytes = (size_t)stages * stage_bytes + out_bytes + (2 * stages + 4) * sizeof(uint64_t) + 16 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_fprop_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_fprop_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  // grid: a multiple of num_n_tiles (every CTA keeps one Cout tile), at most one CTA per SM / per tile
  int grid = a.num_ctas > 0 ? a.num_ctas : HB_NUM_SMS;
  if (grid > HB_NUM_SMS * 4) grid = HB_NUM_SMS * 4;
  int per_n = grid / p.num_n_tiles;
  if (per_n < 1) per_n = 1;
  if (per_n > p.num_m_tiles) per_n = p.num_m_tiles;
  grid = per_n * p.num_n_tiles;
  if (a.stat_slots) *a.stat_slots = 2 * per_n;
  if (p.stats || p.stats2) conv_fprop_kernel<true><<<grid, kThreads, smem_bytes, a.stream>>>(tmA, tmB, tmA2, tmB2, p);
  else conv_fprop_kernel<false><<<grid, kThreads, smem_bytes, a.stream>>>(tmA, tmB, tmA2, tmB2, p);
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
  hb_conv_args a{};
  a.x = x; a.w = w; a.y = y; a.bias = bias; a.residual = residual;
  a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.R = R; a.S = S; a.stride = stride; a.pad = pad; a.dil = dil;
  a.act = act; a.num_ctas = num_ctas;
  return hb_conv2d_fused_bf16(&a, nullptr, stream);
}

int hb_conv_stat_slots_max(void) { return 2 * HB_NUM_SMS * 4; }

// General form (see include/holocron_b200.h): K extension, dual output and output-column statistics.
int hb_conv2d_fused_bf16(const hb_conv_args* c, int* stat_slots, void* stream) {
  if (!c) return (int)cudaErrorInvalidValue;
  const int N = c->N, H = c->H, W = c->W, Cin = c->Cin, Cout = c->Cout, R = c->R, S = c->S;
  const int stride = c->stride, pad = c->pad, dil = c->dil;
  if (Cin % 8 != 0 || Cout % 16 != 0) return (int)cudaErrorInvalidValue;
  if (!hb::aligned16(c->x) || !hb::aligned16(c->w) || !hb::aligned16(c->y)) return (int)cudaErrorMisalignedAddress;
  if (c->w2 && (!hb::aligned16(c->w2) || !hb::aligned16(c->y2))) return (int)cudaErrorMisalignedAddress;
  const int Ho = (H + 2 * pad - dil * (R - 1) - 1) / stride + 1;
  const int Wo = (W + 2 * pad - dil * (S - 1) - 1) / stride + 1;
  if (Ho <= 0 || Wo <= 0) return (int)cudaErrorInvalidValue;
  if ((long long)N * Ho * Wo > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
  if (c->w2 && pad != (R / 2) * dil) return (int)cudaErrorInvalidValue;   // centre tap == the 1x1 pad-0 conv's input
  if ((c->stats || c->stats2) && !stat_slots) return (int)cudaErrorInvalidValue;

  if (R == 3 && S == 3 && stride == 1 && pad == 1 && dil == 1 && !c->xe && !c->w2 && !c->norm_mean) {
    static const bool rows_enabled = getenv("HB_DISABLE_CONV_ROWS") == nullptr;
    if (rows_enabled) {
      const int rc = hb_conv_rows_try(c->x, c->w, c->y, c->bias, c->residual, N, H, W, Cin, Cout, c->act, c->num_ctas,
                                      (cudaStream_t)stream, 0, nullptr, nullptr, c->stats, stat_slots);
      if (rc == 0) return 0;
      if (rc == -2) return (int)cudaErrorLaunchFailure;
    }
  }
  FpropArgs a{};
  a.x = c->x; a.w = c->w; a.y = c->y; a.bias = c->bias; a.residual = c->residual;
  a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.R = R; a.S = S; a.stride = stride;
  a.pad_h = a.pad_w = a.pad_after_h = a.pad_after_w = pad; a.dil = dil; a.act = c->act; a.num_ctas = c->num_ctas;
  a.Ho = Ho; a.Wo = Wo; a.stream = (cudaStream_t)stream;
  a.xe = c->xe; a.we = c->we; a.Ce = c->Ce; a.w2 = c->w2; a.y2 = c->y2;
  a.stats = c->stats; a.stats2 = c->stats2; a.stat_slots = stat_slots;
  a.norm_mean = c->norm_mean; a.norm_rstd = c->norm_rstd; a.norm_wsum = c->norm_wsum;
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
