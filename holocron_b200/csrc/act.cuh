// Activation codes shared by the fused BatchNorm / gate kernels: forward value and derivative w.r.t. the pre-activation.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

namespace hb {

// ACT_FRELU: out = max(z, residual) with z the normalised branch sum (funnel activation, reference activation.py:58-82)
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_RELU6 = 2, ACT_SILU = 3, ACT_LEAKY = 4, ACT_MISH = 5, ACT_HARDMISH = 6, ACT_FRELU = 7 };

__device__ __forceinline__ float act_fwd(int act, float z, float slope) {
  switch (act) {
    case ACT_RELU: return relu_nan(z);
    case ACT_RELU6: return clamp_nan(z, 0.f, 6.f);
    case ACT_SILU: return z / (1.f + __expf(-z));
    case ACT_LEAKY: return z > 0.f ? z : z * slope;
    case ACT_MISH: {
      // x * tanh(softplus(x)); softplus threshold 20 as in torch
      float sp = z > 20.f ? z : log1pf(__expf(z));
      return z * tanhf(sp);
    }
    case ACT_HARDMISH: return (0.5f * z) * clamp_nan(z + 2.f, 0.f, 2.f);
    default: return z;
  }
}
__device__ __forceinline__ float act_grad(int act, float z, float slope) {
  switch (act) {
    case ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case ACT_RELU6: return (z > 0.f && z < 6.f) ? 1.f : 0.f;
    case ACT_SILU: {
      float s = 1.f / (1.f + __expf(-z));
      return s * (1.f + z * (1.f - s));
    }
    case ACT_LEAKY: return z > 0.f ? 1.f : slope;
    case ACT_MISH: {
      float sp = z > 20.f ? z : log1pf(__expf(z));
      float t = tanhf(sp);
      float sg = 1.f / (1.f + __expf(-z));
      return t + z * (1.f - t * t) * sg;
    }
    case ACT_HARDMISH: {
      float t = z + 2.f;
      float c = fminf(fmaxf(t, 0.f), 2.f);
      return 0.5f * c + ((t >= 0.f && t <= 2.f) ? 0.5f * z : 0.f);
    }
    default: return 1.f;
  }
}

}  // namespace hb
