// Hardware probe (development aid, not on the product path): does tcgen05.mma accept a 128B-swizzled K-major A operand
// whose start address is offset by an arbitrary number of 128-byte rows inside a TMA-written buffer?
// Answers how the shifted-window (smem halo reuse) convolution can address its filter taps.
//   out[128 x 64] = A[shift : shift+128, 0:64] * B[64 x 64]^T     A: [rows x 64] bf16 in gmem, B: [64 x 64] bf16
// mode 0: base_offset field = 0;  mode 1: base_offset = (start_addr >> 7) & 7
#include "common.cuh"
#include "tc_common.cuh"
#include "tmap.cuh"

namespace {
using namespace tc;

__global__ void __launch_bounds__(128, 1) shift_probe_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB, float* out, int shift,
                                                             int mode, int rows) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                       // rows x 128 B
  uint8_t* sb = smem + 64 * 1024;           // 64 x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 80 * 1024);
  uint64_t* done = bar + 1;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_ptr, 64);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, rows * 128 + 64 * 128);
    for (int r0 = 0; r0 < rows; r0 += 128) tma_load_2d(&tmA, bar, sa + r0 * 128, 0, r0);
    tma_load_2d(&tmB, bar, sb, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sa) + shift * 128;
    const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    for (int k = 0; k < 4; ++k) {
      uint64_t ad = make_smem_desc(a_addr + k * 32, 16, 1024, kLayoutSW128);
      if (mode == 1) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;
      const uint64_t bd = make_smem_desc(smem_u32(sb) + k * 32, 16, 1024, kLayoutSW128);
      umma_f16(tmem, ad, bd, idesc, k > 0);
    }
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  for (int c = 0; c < 64; c += 16) {
    uint32_t v[16];
    tmem_ld_x16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}
}  // namespace

extern "C" int hb_dev_umma_shift_probe(const void* a, const void* b, float* out, int rows, int shift, int mode, void* stream) {
  if (rows > 512 || rows % 128 != 0 || shift + 128 > rows) return (int)cudaErrorInvalidValue;
  CUtensorMap tmA, tmB;
  uint64_t dimsA[2] = {64, (uint64_t)rows}, dimsB[2] = {64, 64}, strides[1] = {128};
  uint32_t boxA[2] = {64, 128}, boxB[2] = {64, 64};
  int rc = tmap::encode_tiled_bf16(&tmA, a, 2, dimsA, strides, boxA, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = tmap::encode_tiled_bf16(&tmB, b, 2, dimsB, strides, boxB, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  cudaFuncSetAttribute(shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  shift_probe_kernel<<<1, 128, 90 * 1024, (cudaStream_t)stream>>>(tmA, tmB, out, shift, mode, rows);
  HB_LAUNCH_CHECK();
  return 0;
}

// ---- probe 2: tcgen05.mma issue/execute rate as a function of N (M = 128, K = 16, bf16) -----------------------------
// Every CTA (one per SM) issues `count` back-to-back MMAs on the same (uninitialised) smem operands and waits for the
// commit. Reports nothing itself; time it with events from the host.
namespace {
__global__ void __launch_bounds__(128, 1) mma_rate_probe_kernel(int N, int count, int distinct_acc) {
  extern __shared__ __align__(1024) uint8_t smem_raw2[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw2) + 1023) & ~uintptr_t(1023));
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 64 * 1024);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { tc::mbar_init(done, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc(tmem_ptr, 512);
  tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 0);
    const uint32_t dhi = tc::desc_hi(1024, tc::kLayoutSW128);
    const uint32_t a_lo = tc::desc_lo(tc::smem_u32(smem), 16);
    const uint32_t b_lo = tc::desc_lo(tc::smem_u32(smem + 16 * 1024), 16);
    for (int i = 0; i < count; ++i) {
      const uint32_t d = tmem + (distinct_acc ? (uint32_t)((i & 1) * 256) : 0u);
      tc::umma_f16_lh(d, a_lo + 2 * (i & 3), dhi, b_lo + 2 * (i & 3), dhi, idesc, 1u);
    }
    tc::umma_commit(done);
    tc::mbar_wait(done, 0);
  }
  tc::tc_fence_before(); __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}
}  // namespace

// ---- probe 3: the same with the A window starting `shift_rows` 128-byte rows into the (128B-swizzled) buffer, like the
// tap windows of conv_rows.cu: does a start that is not a multiple of the 1024-byte swizzle atom cost fetch bandwidth?
namespace {
__global__ void __launch_bounds__(128, 1) mma_shift_rate_probe_kernel(int N, int count, int shift_rows, int cycle) {
  extern __shared__ __align__(1024) uint8_t smem_raw3[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw3) + 1023) & ~uintptr_t(1023));
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { tc::mbar_init(done, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc(tmem_ptr, 512);
  tc::fence_proxy_async();
  tc::tc_fence_before(); __syncthreads(); tc::tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 0);
    const uint32_t dhi = tc::desc_hi(1024, tc::kLayoutSW128);
    const uint32_t a_lo = tc::desc_lo(tc::smem_u32(smem), 16);
    const uint32_t b_lo = tc::desc_lo(tc::smem_u32(smem + 64 * 1024), 16);
    for (int i = 0; i < count; ++i) {
      // cycle > 0: walk `cycle` different windows (shift_rows apart) like consecutive taps; else always the same window
      const int w = cycle > 0 ? (i / 4) % cycle : 1;
      tc::umma_f16_lh(tmem, a_lo + (uint32_t)(w * shift_rows * 8) + 2 * (i & 3), dhi, b_lo + 2 * (i & 3), dhi, idesc, 1u);
    }
    tc::umma_commit(done);
    tc::mbar_wait(done, 0);
  }
  tc::tc_fence_before(); __syncthreads();
  if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc(tmem, 512); }
}
}  // namespace

extern "C" int hb_dev_mma_shift_rate_probe(int N, int count, int shift_rows, int cycle, int ctas, void* stream) {
  cudaFuncSetAttribute(mma_shift_rate_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  mma_shift_rate_probe_kernel<<<ctas, 128, 98 * 1024, (cudaStream_t)stream>>>(N, count, shift_rows, cycle);
  HB_LAUNCH_CHECK();
  return 0;
}

extern "C" int hb_dev_mma_rate_probe(int N, int count, int distinct_acc, int ctas, void* stream) {
  cudaFuncSetAttribute(mma_rate_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  mma_rate_probe_kernel<<<ctas, 128, 70 * 1024, (cudaStream_t)stream>>>(N, count, distinct_acc);
  HB_LAUNCH_CHECK();
  return 0;
}
