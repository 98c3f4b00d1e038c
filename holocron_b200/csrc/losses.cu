// Classification / segmentation losses: focal, poly-1 (hard + soft targets) and dice.
// Reference: holocron/nn/functional.py:59-113 (focal_loss), :540-613 (poly_loss), :503-537 (dice_loss).
//
// Logits are [N, K, S] (S = product of the spatial dims, possibly 1); a "position" is one (n, s) pair.
// The reference runs log_softmax + transpose/flatten/gather + boolean-mask indexing + mean (~10 kernels, 4-6
// passes over N*K and a host sync for the mask). Here: ONE pass over the logits produces the per-position
// loss and per-CTA partial sums (fp64, combined in a fixed order -> deterministic), and the backward pass
// recomputes the softmax and writes dlogits in one more pass.
//   S == 1 : one warp per position, lanes stride over the K classes (coalesced rows)
//   S  > 1 : one thread per position, consecutive threads = consecutive s (coalesced for every class k)
#include "common.cuh"

namespace {

using namespace hb;

constexpr int kThreads = 256;

enum LossKind { FOCAL = 0, POLY = 1 };

struct LossParams {
  const void* x;            // [N, K, S]
  const long long* target;  // [N, S] hard targets (int64)         (hard)
  const void* soft;         // [N, K, S] soft targets, same dtype  (soft)
  const float* weight;      // [K] or null
  float* loss_pos;          // [N*S] per-position loss
  double* partials;         // [grid][2]: sum of valid losses, number of valid positions
  int N, K, S;
  int ignore_index;         // honoured only when 0 <= ignore_index < K (reference quirk)
  int kind;
  float gamma, eps;
};

template <typename T, typename Acc>
__device__ __forceinline__ float lse_thread(Acc x_at, int K) {
  float mx = -INFINITY;
  for (int k = 0; k < K; ++k) mx = fmaxf(mx, x_at(k));
  float s = 0.f;
  for (int k = 0; k < K; ++k) s += expf(x_at(k) - mx);
  return mx + logf(s);
}

__device__ __forceinline__ float hard_loss(const LossParams& p, float logpt, float w) {
  const float pt = expf(logpt);
  if (p.kind == FOCAL) {
    const float mod = p.gamma == 0.f ? 1.f : powf(fmaxf(1.f - pt, 0.f), p.gamma);
    return -mod * (w * logpt);
  }
  return w * (-logpt + p.eps * (1.f - pt));
}
// d loss / d logpt
__device__ __forceinline__ float hard_dloss(const LossParams& p, float logpt, float w) {
  const float pt = expf(logpt);
  if (p.kind == FOCAL) {
    const float om = fmaxf(1.f - pt, 0.f);
    if (p.gamma == 0.f) return -w;
    const float mod = powf(om, p.gamma);
    const float dmod = om > 0.f ? p.gamma * powf(om, p.gamma - 1.f) * pt : 0.f;  // -(d mod / d logpt)
    return -w * (mod - dmod * logpt);
  }
  return w * (-1.f - p.eps * pt);
}

// ---- hard targets, forward ------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) hard_fwd_kernel(LossParams p) {
  __shared__ double red[32];
  const T* x = (const T*)p.x;
  const long long P = (long long)p.N * p.S;
  double lsum = 0.0, lcnt = 0.0;
  const bool ign = p.ignore_index >= 0 && p.ignore_index < p.K;
  if (p.S == 1) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (kThreads / 32);
    for (long long pos = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); pos < P; pos += warps) {
      const T* row = x + pos * p.K;
      float mx = -INFINITY;
      for (int k = lane; k < p.K; k += 32) mx = fmaxf(mx, to_f(row[k]));
      mx = warp_max(mx);
      float s = 0.f;
      for (int k = lane; k < p.K; k += 32) s += expf(to_f(row[k]) - mx);
      s = warp_sum(s);
      if (lane == 0) {
        const long long t = p.target[pos];
        float l = NAN;
        if (t >= 0 && t < p.K) {
          const float logpt = to_f(row[t]) - (mx + logf(s));
          l = hard_loss(p, logpt, p.weight ? p.weight[t] : 1.f);
        }
        p.loss_pos[pos] = l;
        if (!(ign && t == p.ignore_index)) { lsum += l; lcnt += 1.0; }
      }
    }
  } else {
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long pos = (long long)blockIdx.x * kThreads + threadIdx.x; pos < P; pos += stride) {
      const long long n = pos / p.S, s = pos % p.S;
      const T* base = x + n * p.K * p.S + s;
      auto x_at = [&](int k) { return to_f(base[(long long)k * p.S]); };
      const float lse = lse_thread<T>(x_at, p.K);
      const long long t = p.target[pos];
      float l = NAN;
      if (t >= 0 && t < p.K) l = hard_loss(p, x_at((int)t) - lse, p.weight ? p.weight[t] : 1.f);
      p.loss_pos[pos] = l;
      if (!(ign && t == p.ignore_index)) { lsum += l; lcnt += 1.0; }
    }
  }
  lsum = block_sum<double>(lsum, red);
  lcnt = block_sum<double>(lcnt, red);
  if (threadIdx.x == 0) { p.partials[2 * blockIdx.x] = lsum; p.partials[2 * blockIdx.x + 1] = lcnt; }
}

// out[0] = sum, out[1] = count, out[2] = mean  (fixed summation order)
__global__ void finalize_kernel(const double* partials, int n, float* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0, c = 0.0;
    for (int i = 0; i < n; ++i) { s += partials[2 * i]; c += partials[2 * i + 1]; }
    out[0] = (float)s;
    out[1] = (float)c;
    out[2] = (float)(s / c);
  }
}

// ---- hard targets, backward -----------------------------------------------------------------------
struct LossBwdParams {
  LossParams f;
  const float* gout;    // reduction none: [N*S]; else 1 element
  const float* fwd_out; // {sum, count, mean} from the forward (count used for 'mean')
  void* dx;             // [N, K, S]
  int reduction;        // 0 none, 1 mean, 2 sum
};

template <typename T>
__global__ void __launch_bounds__(kThreads) hard_bwd_kernel(LossBwdParams b) {
  const LossParams& p = b.f;
  const T* x = (const T*)p.x;
  T* dx = (T*)b.dx;
  const long long P = (long long)p.N * p.S;
  const bool ign = p.ignore_index >= 0 && p.ignore_index < p.K;
  const float gscale = b.reduction == 1 ? b.gout[0] / b.fwd_out[1] : (b.reduction == 2 ? b.gout[0] : 0.f);
  if (p.S == 1) {
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (kThreads / 32);
    for (long long pos = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); pos < P; pos += warps) {
      const T* row = x + pos * p.K;
      float mx = -INFINITY;
      for (int k = lane; k < p.K; k += 32) mx = fmaxf(mx, to_f(row[k]));
      mx = warp_max(mx);
      float s = 0.f;
      for (int k = lane; k < p.K; k += 32) s += expf(to_f(row[k]) - mx);
      s = warp_sum(s);
      const float lse = mx + logf(s);
      const long long t = p.target[pos];
      float g = b.reduction == 0 ? b.gout[pos] : ((ign && t == p.ignore_index) ? 0.f : gscale);
      float c = 0.f;
      if (t >= 0 && t < p.K) c = g * hard_dloss(p, to_f(row[t]) - lse, p.weight ? p.weight[t] : 1.f);
      for (int k = lane; k < p.K; k += 32) {
        const float pk = expf(to_f(row[k]) - lse);
        dx[pos * p.K + k] = from_f<T>(c * ((k == t ? 1.f : 0.f) - pk));
      }
    }
  } else {
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long pos = (long long)blockIdx.x * kThreads + threadIdx.x; pos < P; pos += stride) {
      const long long n = pos / p.S, s = pos % p.S;
      const long long off = n * p.K * p.S + s;
      auto x_at = [&](int k) { return to_f(x[off + (long long)k * p.S]); };
      const float lse = lse_thread<T>(x_at, p.K);
      const long long t = p.target[pos];
      float g = b.reduction == 0 ? b.gout[pos] : ((ign && t == p.ignore_index) ? 0.f : gscale);
      float c = 0.f;
      if (t >= 0 && t < p.K) c = g * hard_dloss(p, x_at((int)t) - lse, p.weight ? p.weight[t] : 1.f);
      for (int k = 0; k < p.K; ++k) {
        const float pk = expf(x_at(k) - lse);
        dx[off + (long long)k * p.S] = from_f<T>(c * ((k == t ? 1.f : 0.f) - pk));
      }
    }
  }
}

// ---- poly loss with soft targets ------------------------------------------------------------------
// per position: L = sum_{k valid} w_k * (-z_k + eps * (1 - exp(z_k))),  z_k = log_softmax(x)_k * t_k
template <typename T, bool kBackward>
__global__ void __launch_bounds__(kThreads) poly_soft_kernel(LossBwdParams b) {
  __shared__ double red[32];
  const LossParams& p = b.f;
  const T* x = (const T*)p.x;
  const T* tg = (const T*)p.soft;
  T* dx = (T*)b.dx;
  const long long P = (long long)p.N * p.S;
  const bool ign = p.ignore_index >= 0 && p.ignore_index < p.K;
  double lsum = 0.0;
  const long long stride = (long long)gridDim.x * kThreads;
  for (long long pos = (long long)blockIdx.x * kThreads + threadIdx.x; pos < P; pos += stride) {
    const long long n = pos / p.S, s = pos % p.S;
    const long long off = n * p.K * p.S + s;
    auto x_at = [&](int k) { return to_f(x[off + (long long)k * p.S]); };
    const float lse = lse_thread<T>(x_at, p.K);
    if (!kBackward) {
      float l = 0.f;
      for (int k = 0; k < p.K; ++k) {
        if (ign && k == p.ignore_index) continue;
        const float z = (x_at(k) - lse) * to_f(tg[off + (long long)k * p.S]);
        l += (p.weight ? p.weight[k] : 1.f) * (-z + p.eps * (1.f - expf(z)));
      }
      p.loss_pos[pos] = l;
      lsum += l;
    } else {
      const float g = b.reduction == 0 ? b.gout[pos] : (b.reduction == 1 ? b.gout[0] / (float)P : b.gout[0]);
      // dL/dx_j = c_j t_j - p_j * sum_k c_k t_k,   c_k = w_k * valid_k * (-1 - eps * exp(z_k))
      float tot = 0.f;
      for (int k = 0; k < p.K; ++k) {
        if (ign && k == p.ignore_index) continue;
        const float tk = to_f(tg[off + (long long)k * p.S]);
        const float z = (x_at(k) - lse) * tk;
        tot += (p.weight ? p.weight[k] : 1.f) * (-1.f - p.eps * expf(z)) * tk;
      }
      for (int k = 0; k < p.K; ++k) {
        const float lp = x_at(k) - lse;
        float ck = 0.f;
        if (!(ign && k == p.ignore_index)) {
          const float tk = to_f(tg[off + (long long)k * p.S]);
          ck = (p.weight ? p.weight[k] : 1.f) * (-1.f - p.eps * expf(lp * tk)) * tk;
        }
        dx[off + (long long)k * p.S] = from_f<T>(g * (ck - expf(lp) * tot));
      }
    }
  }
  if (!kBackward) {
    lsum = block_sum<double>(lsum, red);
    if (threadIdx.x == 0) { p.partials[2 * blockIdx.x] = lsum; p.partials[2 * blockIdx.x + 1] = 0.0; }
  }
}

__global__ void finalize_soft_kernel(const double* partials, int n, double P, float* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += partials[2 * i];
    out[0] = (float)s;
    out[1] = (float)P;
    out[2] = (float)(s / P);
  }
}

// ---- dice -----------------------------------------------------------------------------------------
// sums[k] = {sum_{n,s} x*t, sum_{n,s} (x + gamma*t)}; grid = (blocks per class, K)
template <typename T>
__global__ void __launch_bounds__(kThreads) dice_sums_kernel(const T* __restrict__ x, const T* __restrict__ t, int N, int K,
                                                             long long S, float gamma, double* sums) {
  __shared__ double red[32];
  const int k = blockIdx.y;
  double a = 0.0, c = 0.0;
  float fa = 0.f, fc = 0.f;
  const long long per_class = (long long)N * S;
  const long long stride = (long long)gridDim.x * kThreads;
  int cnt = 0;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < per_class; i += stride) {
    const long long n = i / S, s = i % S;
    const long long off = (n * K + k) * S + s;
    const float xv = to_f(x[off]), tv = to_f(t[off]);
    fa = fmaf(xv, tv, fa);
    fc += xv + gamma * tv;
    if (++cnt == 256) { a += fa; c += fc; fa = fc = 0.f; cnt = 0; }  // bounded fp32 partials
  }
  a += fa; c += fc;
  a = block_sum<double>(a, red);
  c = block_sum<double>(c, red);
  if (threadIdx.x == 0) { atomicAdd(&sums[2 * k], a); atomicAdd(&sums[2 * k + 1], c); }
}

// loss = 1 - (1 + 1/gamma) * sum_k w_k * dice_k / sum_k w_k,  dice_k = (gamma*I_k + eps) / (C_k + eps)
// also emits coef[k] = {d loss / d I_k', d loss / d C_k} pieces used by the backward: for element (k):
//   dloss/dx = -(1+1/gamma) * wn_k * (gamma * t * (C_k+eps) - (gamma*I_k+eps)) / (C_k+eps)^2
__global__ void dice_finalize_kernel(const double* sums, const float* weight, int K, float gamma, float eps, float* out,
                                     float* coef /*[K][2]*/) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double wsum = 0.0, acc = 0.0;
  for (int k = 0; k < K; ++k) {
    const double w = weight ? (double)weight[k] : 1.0;
    const double inter = (double)gamma * sums[2 * k] + (double)eps;
    const double card = sums[2 * k + 1] + (double)eps;
    acc += w * inter / card;
    wsum += w;
  }
  const double f = 1.0 + 1.0 / (double)gamma;
  out[0] = (float)(1.0 - f * acc / wsum);
  for (int k = 0; k < K; ++k) {
    const double w = (weight ? (double)weight[k] : 1.0) / wsum;
    const double inter = (double)gamma * sums[2 * k] + (double)eps;
    const double card = sums[2 * k + 1] + (double)eps;
    coef[2 * k] = (float)(-f * w * (double)gamma / card);       // multiplies t
    coef[2 * k + 1] = (float)(f * w * inter / (card * card));   // constant term
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) dice_bwd_kernel(const T* __restrict__ t, const float* __restrict__ coef,
                                                            const float* __restrict__ gout, T* __restrict__ dx, int N,
                                                            int K, long long S) {
  const long long total = (long long)N * K * S;
  const long long stride = (long long)gridDim.x * kThreads;
  const float g = gout[0];
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += stride) {
    const int k = (int)((i / S) % K);
    dx[i] = from_f<T>(g * fmaf(coef[2 * k], to_f(t[i]), coef[2 * k + 1]));
  }
}

int grid_for(long long work, int per_block) {
  long long g = (work + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > HB_NUM_SMS * 8) g = HB_NUM_SMS * 8;
  return (int)g;
}

}  // namespace

extern "C" {

// Upper bound of the number of partial-sum pairs a forward launch writes (size `partials` as 2*this doubles).
int hb_loss_max_partials(void) { return HB_NUM_SMS * 8; }

// kind: 0 focal, 1 poly. fwd_out: float[3] = {sum, valid count, mean}. loss_pos: float[N*S].
int hb_cls_loss_hard_fwd(const void* x, const long long* target, const float* weight, float* loss_pos, double* partials,
                         float* fwd_out, int N, int K, int S, int ignore_index, int kind, float gamma, float eps,
                         int dtype, void* stream) {
  LossParams p{};
  p.x = x; p.target = target; p.weight = weight; p.loss_pos = loss_pos; p.partials = partials;
  p.N = N; p.K = K; p.S = S; p.ignore_index = ignore_index; p.kind = kind; p.gamma = gamma; p.eps = eps;
  const long long P = (long long)N * S;
  if (P == 0) return 0;
  const int grid = grid_for(P, S == 1 ? kThreads / 32 : kThreads);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case HB_DTYPE_F32: hard_fwd_kernel<float><<<grid, kThreads, 0, st>>>(p); break;
    case HB_DTYPE_BF16: hard_fwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>(p); break;
    case HB_DTYPE_F16: hard_fwd_kernel<__half><<<grid, kThreads, 0, st>>>(p); break;
    default: return (int)cudaErrorInvalidValue;
  }
  HB_LAUNCH_CHECK();
  finalize_kernel<<<1, 32, 0, st>>>(partials, grid, fwd_out);
  HB_LAUNCH_CHECK();
  return 0;
}

// reduction: 0 none (gout[N*S]), 1 mean, 2 sum (gout[1]). dx has the dtype/shape of x.
int hb_cls_loss_hard_bwd(const void* x, const long long* target, const float* weight, const float* gout,
                         const float* fwd_out, void* dx, int N, int K, int S, int ignore_index, int kind, float gamma,
                         float eps, int reduction, int dtype, void* stream) {
  LossBwdParams b{};
  b.f.x = x; b.f.target = target; b.f.weight = weight;
  b.f.N = N; b.f.K = K; b.f.S = S; b.f.ignore_index = ignore_index; b.f.kind = kind; b.f.gamma = gamma; b.f.eps = eps;
  b.gout = gout; b.fwd_out = fwd_out; b.dx = dx; b.reduction = reduction;
  const long long P = (long long)N * S;
  if (P == 0) return 0;
  const int grid = grid_for(P, S == 1 ? kThreads / 32 : kThreads);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case HB_DTYPE_F32: hard_bwd_kernel<float><<<grid, kThreads, 0, st>>>(b); break;
    case HB_DTYPE_BF16: hard_bwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>(b); break;
    case HB_DTYPE_F16: hard_bwd_kernel<__half><<<grid, kThreads, 0, st>>>(b); break;
    default: return (int)cudaErrorInvalidValue;
  }
  HB_LAUNCH_CHECK();
  return 0;
}

int hb_poly_soft_fwd(const void* x, const void* soft, const float* weight, float* loss_pos, double* partials,
                     float* fwd_out, int N, int K, int S, int ignore_index, float eps, int dtype, void* stream) {
  LossBwdParams b{};
  b.f.x = x; b.f.soft = soft; b.f.weight = weight; b.f.loss_pos = loss_pos; b.f.partials = partials;
  b.f.N = N; b.f.K = K; b.f.S = S; b.f.ignore_index = ignore_index; b.f.kind = POLY; b.f.eps = eps;
  const long long P = (long long)N * S;
  if (P == 0) return 0;
  const int grid = grid_for(P, kThreads);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case HB_DTYPE_F32: poly_soft_kernel<float, false><<<grid, kThreads, 0, st>>>(b); break;
    case HB_DTYPE_BF16: poly_soft_kernel<__nv_bfloat16, false><<<grid, kThreads, 0, st>>>(b); break;
    case HB_DTYPE_F16: poly_soft_kernel<__half, false><<<grid, kThreads, 0, st>>>(b); break;
    default: return (int)cudaErrorInvalidValue;
  }
  HB_LAUNCH_CHECK();
  finalize_soft_kernel<<<1, 32, 0, st>>>(partials, grid, (double)P, fwd_out);
  HB_LAUNCH_CHECK();
  return 0;
}

int hb_poly_soft_bwd(const void* x, const void* soft, const float* weight, const float* gout, void* dx, int N, int K,
                     int S, int ignore_index, float eps, int reduction, int dtype, void* stream) {
  LossBwdParams b{};
  b.f.x = x; b.f.soft = soft; b.f.weight = weight;
  b.f.N = N; b.f.K = K; b.f.S = S; b.f.ignore_index = ignore_index; b.f.kind = POLY; b.f.eps = eps;
  b.gout = gout; b.dx = dx; b.reduction = reduction;
  const long long P = (long long)N * S;
  if (P == 0) return 0;
  const int grid = grid_for(P, kThreads);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case HB_DTYPE_F32: poly_soft_kernel<float, true><<<grid, kThreads, 0, st>>>(b); break;
    case HB_DTYPE_BF16: poly_soft_kernel<__nv_bfloat16, true><<<grid, kThreads, 0, st>>>(b); break;
    case HB_DTYPE_F16: poly_soft_kernel<__half, true><<<grid, kThreads, 0, st>>>(b); break;
    default: return (int)cudaErrorInvalidValue;
  }
  HB_LAUNCH_CHECK();
  return 0;
}

// sums: double[2K] scratch (zeroed here); out: float[1]; coef: float[2K] (kept for the backward)
int hb_dice_fwd(const void* x, const void* target, const float* weight, double* sums, float* out, float* coef, int N,
                int K, long long S, float gamma, float eps, int dtype, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * K, st);
  if (e != cudaSuccess) return (int)e;
  const long long per_class = (long long)N * S;
  int gx = grid_for(per_class, kThreads * 8);
  if ((long long)gx * K > HB_NUM_SMS * 16) gx = (HB_NUM_SMS * 16 + K - 1) / K;
  if (gx < 1) gx = 1;
  dim3 grid(gx, K);
  switch (dtype) {
    case HB_DTYPE_F32: dice_sums_kernel<float><<<grid, kThreads, 0, st>>>((const float*)x, (const float*)target, N, K, S, gamma, sums); break;
    case HB_DTYPE_BF16: dice_sums_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)target, N, K, S, gamma, sums); break;
    case HB_DTYPE_F16: dice_sums_kernel<__half><<<grid, kThreads, 0, st>>>((const __half*)x, (const __half*)target, N, K, S, gamma, sums); break;
    default: return (int)cudaErrorInvalidValue;
  }
  HB_LAUNCH_CHECK();
  dice_finalize_kernel<<<1, 32, 0, st>>>(sums, weight, K, gamma, eps, out, coef);
  HB_LAUNCH_CHECK();
  return 0;
}

int hb_dice_bwd(const void* target, const float* coef, const float* gout, void* dx, int N, int K, long long S, int dtype,
                void* stream) {
  const long long total = (long long)N * K * S;
  if (total == 0) return 0;
  const int grid = grid_for(total, kThreads * 4);
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case HB_DTYPE_F32: dice_bwd_kernel<float><<<grid, kThreads, 0, st>>>((const float*)target, coef, gout, (float*)dx, N, K, S); break;
    case HB_DTYPE_BF16: dice_bwd_kernel<__nv_bfloat16><<<grid, kThreads, 0, st>>>((const __nv_bfloat16*)target, coef, gout, (__nv_bfloat16*)dx, N, K, S); break;
    case HB_DTYPE_F16: dice_bwd_kernel<__half><<<grid, kThreads, 0, st>>>((const __half*)target, coef, gout, (__half*)dx, N, K, S); break;
    default: return (int)cudaErrorInvalidValue;
  }
  HB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
