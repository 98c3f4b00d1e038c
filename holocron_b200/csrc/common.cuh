// Shared device helpers for the holocron_b200 sm_100a kernels.
// Everything here is header-only; each .cu translation unit includes it.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#define HB_DTYPE_F32 0
#define HB_DTYPE_BF16 1
#define HB_DTYPE_F16 2

#define HB_NUM_SMS 148

#include <atomic>
extern std::atomic<long long> g_hb_launches;  // defined in runtime.cu

// after every kernel launch: count it and surface launch-configuration errors as the return code
#define HB_LAUNCH_CHECK()                          \
  do {                                             \
    g_hb_launches.fetch_add(1, std::memory_order_relaxed); \
    cudaError_t e__ = cudaGetLastError();          \
    if (e__ != cudaSuccess) return (int)e__;       \
  } while (0)

namespace hb {

// ---- scalar conversions -------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// ---- 128-bit vector container -------------------------------------------------------------
template <typename T> struct Vec16 {
  static constexpr int N = 16 / sizeof(T);
  union { uint4 raw; T v[N]; };
};

template <typename T> __device__ __forceinline__ Vec16<T> ld16(const T* p) {
  Vec16<T> r; r.raw = *reinterpret_cast<const uint4*>(p); return r;
}
// streaming (read-once) 128-bit load that does not allocate in L1 (coherent path: safe for in-place ops)
template <typename T> __device__ __forceinline__ Vec16<T> ld16_stream(const T* p) {
  Vec16<T> r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.raw.x), "=r"(r.raw.y), "=r"(r.raw.z), "=r"(r.raw.w) : "l"(p));
  return r;
}
template <typename T> __device__ __forceinline__ void st16(T* p, const Vec16<T>& v) {
  *reinterpret_cast<uint4*>(p) = v.raw;
}

__host__ __device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---- NaN-propagating clamps -------------------------------------------------------------
// torch.relu / clamp / max propagate NaN; CUDA's fmaxf / fminf return the non-NaN operand, which would silently launder a NaN
// activation into 0 at the first ReLU - and with it the reference trainer's NaN-loss detection (trainer/core.py:153-159).
__device__ __forceinline__ float relu_nan(float z) { return z < 0.f ? 0.f : z; }
__device__ __forceinline__ float clamp_nan(float z, float lo, float hi) { return z < lo ? lo : (z > hi ? hi : z); }
__device__ __forceinline__ float max_nan(float a, float b) { return (a > b || a != a) ? a : b; }

// ---- reductions ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; result valid in thread 0 (and broadcast to all when kBroadcast).
// `scratch` must hold >= 32 elements of T in shared memory.
template <typename T, bool kBroadcast = false>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect scratch reuse across consecutive calls
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    T t = lane < nwarps ? scratch[lane] : T(0);
    t = warp_sum(t);
    if (lane == 0) scratch[0] = t;
  }
  if (kBroadcast) { __syncthreads(); return scratch[0]; }
  return (threadIdx.x == 0) ? scratch[0] : T(0);
}

// grid sizing for streaming passes: enough CTAs for >= 2 waves but capped at a multiple of the SM count
__host__ inline int stream_grid(size_t work_items, int per_block, int max_waves = 8) {
  size_t need = (work_items + per_block - 1) / per_block;
  size_t cap = (size_t)HB_NUM_SMS * max_waves;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

}  // namespace hb
