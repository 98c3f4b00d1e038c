// Multi-tensor optimizer steps: AdaBelief, LAMB, TAdam, AdamP, Adan, AdEMAMix, LARS, RaLars and the Lookahead weight
// synchronisation (reference holocron/optim/{adabelief,lamb,tadam,adamp,adan,ademamix,lars,ralars,wrapper}.py).
//
// The reference loops over parameter tensors in Python and issues ~9-14 ATen kernels per tensor (plus, for LAMB,
// two host synchronisations per tensor). Here a whole parameter group is updated by 1 (AdaBelief) or 2-3
// (LAMB / TAdam: a per-tensor reduction has to complete before the update) launches: a device-resident table
// describes every tensor (pointers + numel) and a chunk list maps each CTA to a 4096-element slice of one tensor,
// so the kernels are pure 128-bit-vectorised HBM streams (AdaBelief: 28 B/param algorithmic traffic).
// All state is fp32; per-tensor reductions are accumulated in fp64 atomics (order-insensitive at fp32 precision).
#include "common.cuh"

namespace {

using namespace hb;

constexpr int kThreads = 256;
constexpr int kChunk = 4096;  // elements per CTA

struct TensorMeta {
  float* p;
  const float* g;
  float* m;
  float* v;
  float* vmax;  // amsgrad state or null
  float* aux;   // TAdam: W_t (1 element); LAMB / RaLars: local_lr out (1 element); Adan: prev_grad (full tensor); else null
  float* ext;   // Adan: exp_avg_delta; AdEMAMix: exp_avg_slow; else null
  long long numel;
};

struct Hyper {
  float lr, beta1, beta2, eps, wd;
  float bc1, bc2;       // bias corrections 1 - beta^step (host-computed) ...
  const int* step_dev;  // ... or, when non-null, computed on device from *step_dev (CUDA-graph friendly)
  int amsgrad;
  float clip_lo, clip_hi;  // LAMB
  float dof;               // TAdam (< 0: use numel)
  float delta;             // AdamP
  float beta3, alpha;      // Adan / AdEMAMix
  float bc3;               // Adan: 1 - beta3^step
  float momentum, dampening;  // LARS
  int nesterov, first;        // LARS: `first` = the momentum buffers of this launch's tensors do not exist yet
  int mode;                   // RaLars: 0 rectified (x r_t), 1 plain Adam ratio, 2 unadapted momentum
  float r_t;                  // RaLars variance rectification
  // optional device control block of a captured training step (train_ctl.cu): {lr, beta1, skip, ...}. When given, the
  // learning rate (and beta1 when >= 0) are read from it and the whole update is skipped while skip != 0
  const float* ctl;
};

// applies the control block to a by-value copy of the hyper-parameters; returns false when the update must be skipped
__device__ __forceinline__ bool apply_ctl(Hyper& h) {
  if (!h.ctl) return true;
  if (reinterpret_cast<const int*>(h.ctl)[2] != 0) return false;
  h.lr = h.ctl[0];
  if (h.ctl[1] >= 0.f) h.beta1 = h.ctl[1];
  return true;
}

__device__ __forceinline__ void bias_corrections(const Hyper& h, float& bc1, float& bc2) {
  if (h.step_dev) {
    const double s = (double)(*h.step_dev);
    bc1 = (float)(1.0 - pow((double)h.beta1, s));
    bc2 = (float)(1.0 - pow((double)h.beta2, s));
  } else {
    bc1 = h.bc1; bc2 = h.bc2;
  }
}

// Generic chunk walker: calls f(i) for each element index of this CTA's chunk, 4 at a time when aligned.
template <typename F4, typename F1>
__device__ __forceinline__ void for_chunk(const TensorMeta& t, int chunk, bool vec_ok, F4 f4, F1 f1) {
  const long long base = (long long)chunk * kChunk;
  const long long end = min(base + (long long)kChunk, t.numel);
  if (vec_ok) {
    const long long end4 = base + ((end - base) & ~3LL);
    for (long long i = base + threadIdx.x * 4; i < end4; i += kThreads * 4) f4(i);
    for (long long i = end4 + threadIdx.x; i < end; i += kThreads) f1(i);
  } else {
    for (long long i = base + threadIdx.x; i < end; i += kThreads) f1(i);
  }
}

__device__ __forceinline__ bool meta_vec_ok(const TensorMeta& t) {
  return aligned16(t.p) && aligned16(t.g) && aligned16(t.m) && aligned16(t.v) && (t.vmax == nullptr || aligned16(t.vmax));
}

// ---------------------------------------------------------------------------------------------------
// AdaBelief  (reference adabelief.py:121-167; NB no +eps inside the belief EMA)
__device__ __forceinline__ void adabelief_elem(float& p, float g, float& m, float& s, float* smax, const Hyper& h,
                                               float step_size, float inv_sqrt_bc2) {
  if (h.wd != 0.f) g = fmaf(h.wd, p, g);
  m = fmaf(1.f - h.beta1, g, h.beta1 * m);
  const float r = g - m;
  s = fmaf(1.f - h.beta2, r * r, h.beta2 * s);
  float sec = s;
  if (smax) { *smax = fmaxf(*smax, s); sec = *smax; }
  const float denom = sqrtf(sec) * inv_sqrt_bc2 + h.eps;
  p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(kThreads) adabelief_kernel(const TensorMeta* __restrict__ metas,
                                                             const int2* __restrict__ chunks, Hyper h) {
  if (!apply_ctl(h)) return;
  const int2 c = chunks[blockIdx.x];
  const TensorMeta t = metas[c.x];
  float bc1, bc2;
  bias_corrections(h, bc1, bc2);
  const float step_size = h.lr / bc1;
  const float inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  const bool ams = h.amsgrad && t.vmax;
  for_chunk(t, c.y, meta_vec_ok(t),
      [&](long long i) {
        float4 p = *reinterpret_cast<float4*>(t.p + i);
        const float4 g = *reinterpret_cast<const float4*>(t.g + i);
        float4 m = *reinterpret_cast<float4*>(t.m + i);
        float4 s = *reinterpret_cast<float4*>(t.v + i);
        float4 x = ams ? *reinterpret_cast<float4*>(t.vmax + i) : make_float4(0, 0, 0, 0);
        adabelief_elem(p.x, g.x, m.x, s.x, ams ? &x.x : nullptr, h, step_size, inv_sqrt_bc2);
        adabelief_elem(p.y, g.y, m.y, s.y, ams ? &x.y : nullptr, h, step_size, inv_sqrt_bc2);
        adabelief_elem(p.z, g.z, m.z, s.z, ams ? &x.z : nullptr, h, step_size, inv_sqrt_bc2);
        adabelief_elem(p.w, g.w, m.w, s.w, ams ? &x.w : nullptr, h, step_size, inv_sqrt_bc2);
        *reinterpret_cast<float4*>(t.p + i) = p;
        *reinterpret_cast<float4*>(t.m + i) = m;
        *reinterpret_cast<float4*>(t.v + i) = s;
        if (ams) *reinterpret_cast<float4*>(t.vmax + i) = x;
      },
      [&](long long i) {
        float p = t.p[i], m = t.m[i], s = t.v[i];
        float x = ams ? t.vmax[i] : 0.f;
        adabelief_elem(p, t.g[i], m, s, ams ? &x : nullptr, h, step_size, inv_sqrt_bc2);
        t.p[i] = p; t.m[i] = m; t.v[i] = s;
        if (ams) t.vmax[i] = x;
      });
}

// ---------------------------------------------------------------------------------------------------
// LAMB (reference lamb.py:79-137): no bias correction; update = m/(sqrt(v)+eps) + wd*p;
// local_lr = clamp(||p||, lo, hi) / ||update||  (1 when either norm is 0)
__global__ void __launch_bounds__(kThreads) lamb_moments_kernel(const TensorMeta* __restrict__ metas,
                                                                const int2* __restrict__ chunks, Hyper h,
                                                                double* __restrict__ norms /*[T][2]*/) {
  __shared__ double red[32];
  const int2 c = chunks[blockIdx.x];
  const TensorMeta t = metas[c.x];
  float pn = 0.f, un = 0.f;   // <= 16 elements per thread: fp32 partials, fp64 from the block reduction on
  auto one = [&](float p, float g, float& m, float& v) {
    m = fmaf(1.f - h.beta1, g, h.beta1 * m);
    v = fmaf(1.f - h.beta2, g * g, h.beta2 * v);
    float u = m / (sqrtf(v) + h.eps);
    if (h.wd != 0.f) u = fmaf(h.wd, p, u);
    pn = fmaf(p, p, pn);
    un = fmaf(u, u, un);
  };
  for_chunk(t, c.y, meta_vec_ok(t),
      [&](long long i) {
        const float4 p = *reinterpret_cast<const float4*>(t.p + i);
        const float4 g = *reinterpret_cast<const float4*>(t.g + i);
        float4 m = *reinterpret_cast<float4*>(t.m + i);
        float4 v = *reinterpret_cast<float4*>(t.v + i);
        one(p.x, g.x, m.x, v.x); one(p.y, g.y, m.y, v.y); one(p.z, g.z, m.z, v.z); one(p.w, g.w, m.w, v.w);
        *reinterpret_cast<float4*>(t.m + i) = m;
        *reinterpret_cast<float4*>(t.v + i) = v;
      },
      [&](long long i) {
        float m = t.m[i], v = t.v[i];
        one(t.p[i], t.g[i], m, v);
        t.m[i] = m; t.v[i] = v;
      });
  const double pd = block_sum<double>((double)pn, red);
  const double ud = block_sum<double>((double)un, red);
  if (threadIdx.x == 0) {
    atomicAdd(&norms[2 * c.x + 0], pd);
    atomicAdd(&norms[2 * c.x + 1], ud);
  }
}

__global__ void __launch_bounds__(kThreads) lamb_apply_kernel(const TensorMeta* __restrict__ metas,
                                                              const int2* __restrict__ chunks, Hyper h,
                                                              const double* __restrict__ norms) {
  const int2 c = chunks[blockIdx.x];
  const TensorMeta t = metas[c.x];
  const float p_norm = (float)sqrt(norms[2 * c.x + 0]);
  const float u_norm = (float)sqrt(norms[2 * c.x + 1]);
  const float phi = fminf(fmaxf(p_norm, h.clip_lo), h.clip_hi);
  const float local_lr = (phi == 0.f || u_norm == 0.f) ? 1.f : phi / u_norm;
  if (c.y == 0 && threadIdx.x == 0 && t.aux) *t.aux = local_lr;
  const float a = h.lr * local_lr;
  auto one = [&](float p, float m, float v) {
    float u = m / (sqrtf(v) + h.eps);
    if (h.wd != 0.f) u = fmaf(h.wd, p, u);
    return p - a * u;
  };
  for_chunk(t, c.y, meta_vec_ok(t),
      [&](long long i) {
        float4 p = *reinterpret_cast<float4*>(t.p + i);
        const float4 m = *reinterpret_cast<const float4*>(t.m + i);
        const float4 v = *reinterpret_cast<const float4*>(t.v + i);
        p.x = one(p.x, m.x, v.x); p.y = one(p.y, m.y, v.y); p.z = one(p.z, m.z, v.z); p.w = one(p.w, m.w, v.w);
        *reinterpret_cast<float4*>(t.p + i) = p;
      },
      [&](long long i) { t.p[i] = one(t.p[i], t.m[i], t.v[i]); });
}

// ---------------------------------------------------------------------------------------------------
// AdamP (reference adamp.py:144-191): Adam moments with bias correction; when the gradient is (nearly) orthogonal to
// the weight tensor, cos(p, g) < delta / sqrt(numel), the radial component of the update is projected out:
//   pt = (m / bc1) / (sqrt(v) / sqrt(bc2) + eps);  pt -= <p / (||p|| + eps), pt> * p / (||p|| + eps);  p -= lr * pt
// Pass 1 updates the moments and reduces <p,g>, ||p||^2, ||g||^2, <p,pt> per tensor; pass 2 applies (40 B / parameter).
__device__ __forceinline__ float adamp_pt(float m, float v, float vmax_or_neg, float bc1, float inv_sqrt_bc2, float eps) {
  const float sec = vmax_or_neg >= 0.f ? vmax_or_neg : v;
  return (m / bc1) / (sqrtf(sec) * inv_sqrt_bc2 + eps);
}

__global__ void __launch_bounds__(kThreads) adamp_moments_kernel(const TensorMeta* __restrict__ metas,
                                                                 const int2* __restrict__ chunks, Hyper h,
                                                                 double* __restrict__ sums /*[T][4]*/) {
  __shared__ double red[32];
  if (!apply_ctl(h)) return;
  const int2 c = chunks[blockIdx.x];
  const TensorMeta t = metas[c.x];
  float bc1, bc2;
  bias_corrections(h, bc1, bc2);
  const float inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  const bool ams = h.amsgrad && t.vmax;
  float s_pg = 0.f, s_pp = 0.f, s_gg = 0.f, s_ppt = 0.f;
  auto one = [&](float p, float g, float& m, float& v, float& x) {
    if (h.wd != 0.f) g = fmaf(h.wd, p, g);
    m = fmaf(1.f - h.beta1, g, h.beta1 * m);
    v = fmaf(1.f - h.beta2, g * g, h.beta2 * v);
    if (ams) x = fmaxf(x, v);
    const float pt = adamp_pt(m, v, ams ? x : -1.f, bc1, inv_sqrt_bc2, h.eps);
    s_pg = fmaf(p, g, s_pg); s_pp = fmaf(p, p, s_pp); s_gg = fmaf(g, g, s_gg); s_ppt = fmaf(p, pt, s_ppt);
  };
  for_chunk(t, c.y, meta_vec_ok(t),
      [&](long long i) {
        const float4 p = *reinterpret_cast<const float4*>(t.p + i);
        const float4 g = *reinterpret_cast<const float4*>(t.g + i);
        float4 m = *reinterpret_cast<float4*>(t.m + i);
        float4 v = *reinterpret_cast<float4*>(t.v + i);
        float4 x = ams ? *reinterpret_cast<float4*>(t.vmax + i) : make_float4(0, 0, 0, 0);
        one(p.x, g.x, m.x, v.x, x.x); one(p.y, g.y, m.y, v.y, x.y); one(p.z, g.z, m.z, v.z, x.z); one(p.w, g.w, m.w, v.w, x.w);
        *reinterpret_cast<float4*>(t.m + i) = m;
        *reinterpret_cast<float4*>(t.v + i) = v;
        if (ams) *reinterpret_cast<float4*>(t.vmax + i) = x;
      },
      [&](long long i) {
        float m = t.m[i], v = t.v[i], x = ams ? t.vmax[i] : 0.f;
        one(t.p[i], t.g[i], m, v, x);
        t.m[i] = m; t.v[i] = v;
        if (ams) t.vmax[i] = x;
      });
  const double a = block_sum<double>((double)s_pg, red), b = block_sum<double>((double)s_pp, red);
  const double cc = block_sum<double>((double)s_gg, red), d = block_sum<double>((double)s_ppt, red);
  if (threadIdx.x == 0) {
    atomicAdd(&sums[4 * c.x + 0], a); atomicAdd(&sums[4 * c.x + 1], b);
    atomicAdd(&sums[4 * c.x + 2], cc); atomicAdd(&sums[4 * c.x + 3], d);
  }
}

__global__ void __launch_bounds__(kThreads) adamp_apply_kernel(const TensorMeta* __restrict__ metas,
                                                               const int2* __restrict__ chunks, Hyper h,
                                                               const double* __restrict__ sums) {
  if (!apply_ctl(h)) return;
  const int2 c = chunks[blockIdx.x];
  const TensorMeta t = metas[c.x];
  float bc1, bc2;
  bias_corrections(h, bc1, bc2);
  const float inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  const bool ams = h.amsgrad && t.vmax;
  // F.cosine_similarity clamps each norm at 1e-8 (torch eps default)
  const float pn = (float)sqrt(sums[4 * c.x + 1]), gn = (float)sqrt(sums[4 * c.x + 2]);
  const float cosv = (float)sums[4 * c.x + 0] / (fmaxf(pn, 1e-8f) * fmaxf(gn, 1e-8f));
  const bool project = cosv < h.delta / sqrtf((float)t.numel);
  const float inv = 1.f / (pn + h.eps);
  const float k = project ? (float)sums[4 * c.x + 3] * inv * inv : 0.f;   // <p_hat, pt> / (||p|| + eps)
  auto one = [&](float p, float m, float v, float x) {
    float pt = adamp_pt(m, v, ams ? x : -1.f, bc1, inv_sqrt_bc2, h.eps);
    pt = fmaf(-k, p, pt);
    return fmaf(-h.lr, pt, p);
  };
  for_chunk(t, c.y, meta_vec_ok(t),
      [&](long long i) {
        float4 p = *reinterpret_cast<float4*>(t.p + i);
        const float4 m = *reinterpret_cast<const float4*>(t.m + i);
        const float4 v = *reinterpret_cast<const float4*>(t.v + i);
        const float4 x = ams ? *reinterpret_cast<const float4*>(t.vmax + i) : make_float4(0, 0, 0, 0);
        p.x = one(p.x, m.x, v.x, x.x); p.y = one(p.y, m.y, v.y, x.y); p.z = one(p.z, m.z, v.z, x.z); p.w = one(p.w, m.w, v.w, x.w);
        *reinterpret_cast<float4*>(t.p + i) = p;
      },
      [&](long long i) { t.p[i] = one(t.p[i], t.m[i], t.v[i], ams ? t.vmax[i] : 0.f); });
}

// ---------------------------------------------------------------------------------------------------
// TAdam (reference tadam.py:160-212)
__global__ void __launch_bounds__(kThreads) tadam_reduce_kernel(const TensorMeta* __restrict__ metas,
                                                                const int2* __restrict__ chunks, Hyper h,
                                                                double* __restrict__ sums /*[T]*/) {
  __shared__ double red[32];
  const int2 c = chunks[blockIdx.x];
  const TensorMeta t = metas[c.x];
  float acc = 0.f;
  auto one = [&](float p, float g, float m, float v) {
    if (h.wd != 0.f) g = fmaf(h.wd, p, g);
    const float d = g - m;
    acc += (d * d) / (v + h.eps);
  };
  const bool need_p = h.wd != 0.f;
  for_chunk(t, c.y, meta_vec_ok(t),
      [&](long long i) {
        const float4 g = *reinterpret_cast<const float4*>(t.g + i);
        const float4 m = *reinterpret_cast<const float4*>(t.m + i);
        const float4 v = *reinterpret_cast<const float4*>(t.v + i);
        const float4 p = need_p ? *reinterpret_cast<const float4*>(t.p + i) : make_float4(0, 0, 0, 0);
        one(p.x, g.x, m.x, v.x); one(p.y, g.y, m.y, v.y); one(p.z, g.z, m.z, v.z); one(p.w, g.w, m.w, v.w);
      },
      [&](long long i) { one(need_p ? t.p[i] : 0.f, t.g[i], t.m[i], t.v[i]); });
  const double tot = block_sum<double>((double)acc, red);
  if (threadIdx.x == 0) atomicAdd(&sums[c.x], tot);
}

__device__ __forceinline__ float tadam_wt(const TensorMeta& t, const Hyper& h, double sum) {
  const float n = (float)t.numel;
  const float dof = h.dof < 0.f ? n : h.dof;
  return (dof + n) / ((float)sum + dof);
}

__global__ void __launch_bounds__(kThreads) tadam_apply_kernel(const TensorMeta* __restrict__ metas,
                                                               const int2* __restrict__ chunks, Hyper h,
                                                               const double* __restrict__ sums) {
  const int2 c = chunks[blockIdx.x];
  const TensorMeta t = metas[c.x];
  float bc1, bc2;
  bias_corrections(h, bc1, bc2);
  const float step_size = h.lr / bc1;
  const float inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  const float w = tadam_wt(t, h, sums[c.x]);
  const float W = *t.aux;
  const float a = W / (W + w);
  const bool ams = h.amsgrad && t.vmax;
  auto one = [&](float& p, float g, float& m, float& v, float& x) {
    if (h.wd != 0.f) g = fmaf(h.wd, p, g);
    m = m * a + (w * g) / (W + w);
    v = fmaf(1.f - h.beta2, g * g, h.beta2 * v);
    float sec = v;
    if (ams) { x = fmaxf(x, v); sec = x; }
    const float denom = sqrtf(sec) * inv_sqrt_bc2 + h.eps;
    p = p - step_size * (m / denom);
  };
  for_chunk(t, c.y, meta_vec_ok(t),
      [&](long long i) {
        float4 p = *reinterpret_cast<float4*>(t.p + i);
        const float4 g = *reinterpret_cast<const float4*>(t.g + i);
        float4 m = *reinterpret_cast<float4*>(t.m + i);
        float4 v = *reinterpret_cast<float4*>(t.v + i);
        float4 x = ams ? *reinterpret_cast<float4*>(t.vmax + i) : make_float4(0, 0, 0, 0);
        one(p.x, g.x, m.x, v.x, x.x); one(p.y, g.y, m.y, v.y, x.y); one(p.z, g.z, m.z, v.z, x.z); one(p.w, g.w, m.w, v.w, x.w);
        *reinterpret_cast<float4*>(t.p + i) = p;
        *reinterpret_cast<float4*>(t.m + i) = m;
        *reinterpret_cast<float4*>(t.v + i) = v;
        if (ams) *reinterpret_cast<float4*>(t.vmax + i) = x;
      },
      [&](long long i) {
        float p = t.p[i], m = t.m[i], v = t.v[i], x = ams ? t.vmax[i] : 0.f;
        one(p, t.g[i], m, v, x);
        t.p[i] = p; t.m[i] = m; t.v[i] = v;
        if (ams) t.vmax[i] = x;
      });
}

// W_t <- W_t * (2 beta1 - 1) / beta1 + w_t   (after every CTA of the apply pass has read the old W_t)
__global__ void tadam_wt_update_kernel(const TensorMeta* __restrict__ metas, int T, Hyper h,
                                       const double* __restrict__ sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T) return;
  const TensorMeta t = metas[i];
  const float w = tadam_wt(t, h, sums[i]);
  *t.aux = *t.aux * ((2.f * h.beta1 - 1.f) / h.beta1) + w;
}

// ---------------------------------------------------------------------------------------------------
// Adan (reference adan.py:145-199). Quirks kept: `prev_grad` is read but never written by the reference (it stays at its
// initial zeros, so delta_grad == grad unless a loaded state says otherwise); the update mixes beta2 * exp_avg_sq / bc2
// (not 1 - beta2); with weight decay the parameter is divided by (1 + wd * lr) after the step.
// m = exp_avg, v = exp_avg_sq (EMA of gradient differences), ext = exp_avg_delta (EMA of squares), vmax = its running max.
__global__ void __launch_bounds__(kThreads) adan_kernel(const TensorMeta* __restrict__ metas, const int2* __restrict__ chunks,
                                                        Hyper h) {
  if (!apply_ctl(h)) return;
  const int2 c = chunks[blockIdx.x];
  const TensorMeta t = metas[c.x];
  float bc1, bc2, bc3 = h.bc3;
  bias_corrections(h, bc1, bc2);
  if (h.step_dev) bc3 = (float)(1.0 - pow((double)h.beta3, (double)(*h.step_dev)));
  const float inv_sqrt_bc3 = 1.f / sqrtf(bc3);
  const bool ams = h.amsgrad && t.vmax;
  const float shrink = 1.f + h.wd * h.lr;
  auto one = [&](float& p, float g, float pg, float& m, float& v, float& n, float& x) {
    if (h.wd != 0.f) g = fmaf(h.wd, p, g);
    m = fmaf(1.f - h.beta1, g, h.beta1 * m);
    const float dg = g - pg;
    v = fmaf(1.f - h.beta2, dg, h.beta2 * v);
    const float tmp = fmaf(h.beta2, dg, g);
    n = fmaf(1.f - h.beta3, tmp * tmp, h.beta3 * n);
    float sec = n;
    if (ams) { x = fmaxf(x, n); sec = x; }
    const float denom = sqrtf(sec) * inv_sqrt_bc3 + h.eps;
    const float pt = (m / bc1 + h.beta2 * v / bc2) / denom;
    p = fmaf(-h.lr, pt, p);
    if (h.wd != 0.f) p = p / shrink;
  };
  const bool vec = meta_vec_ok(t) && aligned16(t.aux) && aligned16(t.ext);
  for_chunk(t, c.y, vec,
      [&](long long i) {
        float4 p = *reinterpret_cast<float4*>(t.p + i);
        const float4 g = *reinterpret_cast<const float4*>(t.g + i);
        const float4 pg = *reinterpret_cast<const float4*>(t.aux + i);
        float4 m = *reinterpret_cast<float4*>(t.m + i);
        float4 v = *reinterpret_cast<float4*>(t.v + i);
        float4 n = *reinterpret_cast<float4*>(t.ext + i);
        float4 x = ams ? *reinterpret_cast<float4*>(t.vmax + i) : make_float4(0, 0, 0, 0);
        one(p.x, g.x, pg.x, m.x, v.x, n.x, x.x); one(p.y, g.y, pg.y, m.y, v.y, n.y, x.y);
        one(p.z, g.z, pg.z, m.z, v.z, n.z, x.z); one(p.w, g.w, pg.w, m.w, v.w, n.w, x.w);
        *reinterpret_cast<float4*>(t.p + i) = p;
        *reinterpret_cast<float4*>(t.m + i) = m;
        *reinterpret_cast<float4*>(t.v + i) = v;
        *reinterpret_cast<float4*>(t.ext + i) = n;
        if (ams) *reinterpret_cast<float4*>(t.vmax + i) = x;
      },
      [&](long long i) {
        float p = t.p[i], m = t.m[i], v = t.v[i], n = t.ext[i], x = ams ? t.vmax[i] : 0.f;
        one(p, t.g[i], t.aux[i], m, v, n, x);
        t.p[i] = p; t.m[i] = m; t.v[i] = v; t.ext[i] = n;
        if (ams) t.vmax[i] = x;
      });
}

// ---------------------------------------------------------------------------------------------------
// AdEMAMix (reference ademamix.py:138-176): fast EMA m1 (bias-corrected), slow EMA m2 (beta3, not corrected), Adam second
// moment; p -= lr * (m1 / bc1 + alpha * m2) / (sqrt(nu) / sqrt(bc2) + eps). m = exp_avg, ext = exp_avg_slow, v = exp_avg_sq.
__global__ void __launch_bounds__(kThreads) ademamix_kernel(const TensorMeta* __restrict__ metas,
                                                            const int2* __restrict__ chunks, Hyper h) {
  if (!apply_ctl(h)) return;
  const int2 c = chunks[blockIdx.x];
  const TensorMeta t = metas[c.x];
  float bc1, bc2;
  bias_corrections(h, bc1, bc2);
  const float inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  auto one = [&](float& p, float g, float& m1, float& m2, float& nu) {
    if (h.wd != 0.f) g = fmaf(h.wd, p, g);
    m1 = fmaf(1.f - h.beta1, g, h.beta1 * m1);
    nu = fmaf(1.f - h.beta2, g * g, h.beta2 * nu);
    m2 = fmaf(1.f - h.beta3, g, h.beta3 * m2);
    const float denom = sqrtf(nu) * inv_sqrt_bc2 + h.eps;
    p = fmaf(-h.lr, fmaf(h.alpha, m2, m1 / bc1) / denom, p);
  };
  const bool vec = meta_vec_ok(t) && aligned16(t.ext);
  for_chunk(t, c.y, vec,
      [&](long long i) {
        float4 p = *reinterpret_cast<float4*>(t.p + i);
        const float4 g = *reinterpret_cast<const float4*>(t.g + i);
        float4 m1 = *reinterpret_cast<float4*>(t.m + i);
        float4 m2 = *reinterpret_cast<float4*>(t.ext + i);
        float4 nu = *reinterpret_cast<float4*>(t.v + i);
        one(p.x, g.x, m1.x, m2.x, nu.x); one(p.y, g.y, m1.y, m2.y, nu.y);
        one(p.z, g.z, m1.z, m2.z, nu.z); one(p.w, g.w, m1.w, m2.w, nu.w);
        *reinterpret_cast<float4*>(t.p + i) = p;
        *reinterpret_cast<float4*>(t.m + i) = m1;
        *reinterpret_cast<float4*>(t.ext + i) = m2;
        *reinterpret_cast<float4*>(t.v + i) = nu;
      },
      [&](long long i) {
        float p = t.p[i], m1 = t.m[i], m2 = t.ext[i], nu = t.v[i];
        one(p, t.g[i], m1, m2, nu);
        t.p[i] = p; t.m[i] = m1; t.ext[i] = m2; t.v[i] = nu;
      });
}

// ---------------------------------------------------------------------------------------------------
// Per-tensor sums of squares of two streams (fp64 atomics per CTA): norms[2t] += sum a^2, norms[2t+1] += sum b^2
template <typename FA, typename FB>
__device__ __forceinline__ void two_norms(const TensorMeta& t, int chunk, bool vec, FA a_at4, FB b_at4, double* norms, int ti,
                                          double* red) {
  float an = 0.f, bn = 0.f;
  for_chunk(t, chunk, vec,
      [&](long long i) {
        const float4 a = a_at4(i, true), b = b_at4(i, true);
        an = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, fmaf(a.w, a.w, an))));
        bn = fmaf(b.x, b.x, fmaf(b.y, b.y, fmaf(b.z, b.z, fmaf(b.w, b.w, bn))));
      },
      [&](long long i) {
        const float4 a = a_at4(i, false), b = b_at4(i, false);
        an = fmaf(a.x, a.x, an);
        bn = fmaf(b.x, b.x, bn);
      });
  const double ad = block_sum<double>((double)an, red);
  const double bd = block_sum<double>((double)bn, red);
  if (threadIdx.x == 0) { atomicAdd(&norms[2 * ti], ad); atomicAdd(&norms[2 * ti + 1], bd); }
}

// LARS (reference lars.py:91-135): local_lr = ||p|| / (||g|| + wd ||p||) (1 when either is 0; `scale_clip` is stored but never
// applied by the reference); d_p = g + wd * p is written back INTO THE GRADIENT like the reference's in-place add_;
// SGD momentum with dampening / Nesterov, the first buffer being a copy of d_p. m = momentum_buffer (may be null).
__global__ void __launch_bounds__(kThreads) lars_norms_kernel(const TensorMeta* __restrict__ metas,
                                                              const int2* __restrict__ chunks, double* __restrict__ norms) {
  __shared__ double red[32];
  const int2 c = chunks[blockIdx.x];
  const TensorMeta t = metas[c.x];
  const bool vec = aligned16(t.p) && aligned16(t.g);
  two_norms(t, c.y, vec,
      [&](long long i, bool v4) { return v4 ? *reinterpret_cast<const float4*>(t.p + i) : make_float4(t.p[i], 0, 0, 0); },
      [&](long long i, bool v4) { return v4 ? *reinterpret_cast<const float4*>(t.g + i) : make_float4(t.g[i], 0, 0, 0); },
      norms, c.x, red);
}

__global__ void __launch_bounds__(kThreads) lars_apply_kernel(const TensorMeta* __restrict__ metas,
                                                              const int2* __restrict__ chunks, Hyper h,
                                                              const double* __restrict__ norms) {
  const int2 c = chunks[blockIdx.x];
  const TensorMeta t = metas[c.x];
  const float p_norm = (float)sqrt(norms[2 * c.x + 0]);
  float denom = (float)sqrt(norms[2 * c.x + 1]);
  if (h.wd != 0.f) denom = fmaf(h.wd, p_norm, denom);
  const float local_lr = (p_norm == 0.f || denom == 0.f) ? 1.f : p_norm / denom;
  const float a = h.lr * local_lr;
  float* gw = const_cast<float*>(t.g);
  const bool mom = h.momentum != 0.f && t.m != nullptr;
  auto one = [&](float& p, float& g, float& b) {
    if (h.wd != 0.f) g = fmaf(h.wd, p, g);
    float d = g;
    if (mom) {
      b = h.first ? g : fmaf(h.momentum, b, (1.f - h.dampening) * g);
      d = h.nesterov ? fmaf(h.momentum, b, g) : b;
    }
    p = fmaf(-a, d, p);
  };
  const bool vec = aligned16(t.p) && aligned16(t.g) && (!mom || aligned16(t.m));
  for_chunk(t, c.y, vec,
      [&](long long i) {
        float4 p = *reinterpret_cast<float4*>(t.p + i);
        float4 g = *reinterpret_cast<const float4*>(t.g + i);
        float4 b = (mom && !h.first) ? *reinterpret_cast<float4*>(t.m + i) : make_float4(0, 0, 0, 0);
        one(p.x, g.x, b.x); one(p.y, g.y, b.y); one(p.z, g.z, b.z); one(p.w, g.w, b.w);
        *reinterpret_cast<float4*>(t.p + i) = p;
        if (h.wd != 0.f) *reinterpret_cast<float4*>(gw + i) = g;
        if (mom) *reinterpret_cast<float4*>(t.m + i) = b;
      },
      [&](long long i) {
        float p = t.p[i], g = t.g[i], b = (mom && !h.first) ? t.m[i] : 0.f;
        one(p, g, b);
        t.p[i] = p;
        if (h.wd != 0.f) gw[i] = g;
        if (mom) t.m[i] = b;
      });
}

// ---------------------------------------------------------------------------------------------------
// RaLars (reference ralars.py:56-140): RAdam update (rectified / plain Adam ratio / unadapted momentum, chosen on the host
// from the SMA length) + wd * p, scaled by the LARS trust ratio clamp(||p||, *scale_clip) / ||update||.
__device__ __forceinline__ float ralars_update(float p, float m, float v, const Hyper& h, float bc1, float bc2) {
  float u;
  if (h.mode == 2) u = m / bc1;
  else u = h.r_t * ((m / bc1) / (sqrtf(v / bc2) + h.eps));
  if (h.wd != 0.f) u = fmaf(h.wd, p, u);
  return u;
}

__global__ void __launch_bounds__(kThreads) ralars_moments_kernel(const TensorMeta* __restrict__ metas,
                                                                  const int2* __restrict__ chunks, Hyper h,
                                                                  double* __restrict__ norms) {
  __shared__ double red[32];
  const int2 c = chunks[blockIdx.x];
  const TensorMeta t = metas[c.x];
  float pn = 0.f, un = 0.f;
  auto one = [&](float p, float g, float& m, float& v) {
    m = fmaf(1.f - h.beta1, g, h.beta1 * m);
    v = fmaf(1.f - h.beta2, g * g, h.beta2 * v);
    const float u = ralars_update(p, m, v, h, h.bc1, h.bc2);
    pn = fmaf(p, p, pn);
    un = fmaf(u, u, un);
  };
  for_chunk(t, c.y, meta_vec_ok(t),
      [&](long long i) {
        const float4 p = *reinterpret_cast<const float4*>(t.p + i);
        const float4 g = *reinterpret_cast<const float4*>(t.g + i);
        float4 m = *reinterpret_cast<float4*>(t.m + i);
        float4 v = *reinterpret_cast<float4*>(t.v + i);
        one(p.x, g.x, m.x, v.x); one(p.y, g.y, m.y, v.y); one(p.z, g.z, m.z, v.z); one(p.w, g.w, m.w, v.w);
        *reinterpret_cast<float4*>(t.m + i) = m;
        *reinterpret_cast<float4*>(t.v + i) = v;
      },
      [&](long long i) {
        float m = t.m[i], v = t.v[i];
        one(t.p[i], t.g[i], m, v);
        t.m[i] = m; t.v[i] = v;
      });
  const double pd = block_sum<double>((double)pn, red);
  const double ud = block_sum<double>((double)un, red);
  if (threadIdx.x == 0) { atomicAdd(&norms[2 * c.x + 0], pd); atomicAdd(&norms[2 * c.x + 1], ud); }
}

__global__ void __launch_bounds__(kThreads) ralars_apply_kernel(const TensorMeta* __restrict__ metas,
                                                                const int2* __restrict__ chunks, Hyper h,
                                                                const double* __restrict__ norms) {
  const int2 c = chunks[blockIdx.x];
  const TensorMeta t = metas[c.x];
  const float p_norm = (float)sqrt(norms[2 * c.x + 0]);
  const float u_norm = (float)sqrt(norms[2 * c.x + 1]);
  const float phi = fminf(fmaxf(p_norm, h.clip_lo), h.clip_hi);
  const float local_lr = (phi == 0.f || u_norm == 0.f) ? 1.f : phi / u_norm;
  if (c.y == 0 && threadIdx.x == 0 && t.aux) *t.aux = local_lr;
  const float a = h.lr * local_lr;
  for_chunk(t, c.y, meta_vec_ok(t),
      [&](long long i) {
        float4 p = *reinterpret_cast<float4*>(t.p + i);
        const float4 m = *reinterpret_cast<const float4*>(t.m + i);
        const float4 v = *reinterpret_cast<const float4*>(t.v + i);
        p.x = fmaf(-a, ralars_update(p.x, m.x, v.x, h, h.bc1, h.bc2), p.x);
        p.y = fmaf(-a, ralars_update(p.y, m.y, v.y, h, h.bc1, h.bc2), p.y);
        p.z = fmaf(-a, ralars_update(p.z, m.z, v.z, h, h.bc1, h.bc2), p.z);
        p.w = fmaf(-a, ralars_update(p.w, m.w, v.w, h, h.bc1, h.bc2), p.w);
        *reinterpret_cast<float4*>(t.p + i) = p;
      },
      [&](long long i) { t.p[i] = fmaf(-a, ralars_update(t.p[i], t.m[i], t.v[i], h, h.bc1, h.bc2), t.p[i]); });
}

// ---------------------------------------------------------------------------------------------------
// Lookahead.sync_params (reference wrapper.py:122-135): slow += rate * (fast - slow) (skipped when rate == 0); fast = slow.
// p = fast weights, m = slow weights.
__global__ void __launch_bounds__(kThreads) lookahead_sync_kernel(const TensorMeta* __restrict__ metas,
                                                                  const int2* __restrict__ chunks, float rate) {
  const int2 c = chunks[blockIdx.x];
  const TensorMeta t = metas[c.x];
  auto one = [&](float f, float s) { return rate > 0.f ? fmaf(rate, f - s, s) : s; };
  for_chunk(t, c.y, aligned16(t.p) && aligned16(t.m),
      [&](long long i) {
        const float4 f = *reinterpret_cast<const float4*>(t.p + i);
        float4 s = *reinterpret_cast<float4*>(t.m + i);
        s.x = one(f.x, s.x); s.y = one(f.y, s.y); s.z = one(f.z, s.z); s.w = one(f.w, s.w);
        *reinterpret_cast<float4*>(t.m + i) = s;
        *reinterpret_cast<float4*>(t.p + i) = s;
      },
      [&](long long i) {
        const float s = one(t.p[i], t.m[i]);
        t.m[i] = s; t.p[i] = s;
      });
}

__global__ void step_increment_kernel(int* step, const int* ctl) { if (!ctl || ctl[2] == 0) *step += 1; }

Hyper make_hyper(float lr, float b1, float b2, float eps, float wd, int step, const int* step_dev, int amsgrad) {
  Hyper h{};
  h.lr = lr; h.beta1 = b1; h.beta2 = b2; h.eps = eps; h.wd = wd;
  h.bc1 = (float)(1.0 - pow((double)b1, (double)step));
  h.bc2 = (float)(1.0 - pow((double)b2, (double)step));
  h.step_dev = step_dev;
  h.amsgrad = amsgrad;
  h.dof = -1.f;
  h.delta = 0.1f;
  h.beta3 = 0.f; h.alpha = 0.f; h.bc3 = 1.f;
  h.momentum = 0.f; h.dampening = 0.f; h.nesterov = 0; h.first = 0;
  h.mode = 0; h.r_t = 1.f;
  h.ctl = nullptr;
  return h;
}

}  // namespace

extern "C" {

// metas: device array of T records {p, g, m, v, vmax, aux, ext, numel} (8 x 8 bytes); chunks: device int2[num_chunks]
// {tensor index, chunk index} with chunk = 4096 elements.
int hb_optim_chunk_elems(void) { return kChunk; }

int hb_adabelief_step(const void* metas, const void* chunks, int num_chunks, float lr, float beta1, float beta2,
                      float eps, float weight_decay, int amsgrad, int step, const int* step_dev, const void* ctl,
                      void* stream) {
  if (num_chunks <= 0) return 0;
  Hyper h = make_hyper(lr, beta1, beta2, eps, weight_decay, step, step_dev, amsgrad);
  h.ctl = (const float*)ctl;
  adabelief_kernel<<<num_chunks, kThreads, 0, (cudaStream_t)stream>>>((const TensorMeta*)metas, (const int2*)chunks, h);
  HB_LAUNCH_CHECK();
  return 0;
}

// scratch: device double[2*T], zeroed here. local_lr lands in each tensor's aux slot (if non-null).
int hb_lamb_step(const void* metas, const void* chunks, int num_chunks, int T, float lr, float beta1, float beta2,
                 float eps, float weight_decay, float clip_lo, float clip_hi, double* scratch, void* stream) {
  if (num_chunks <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  Hyper h = make_hyper(lr, beta1, beta2, eps, weight_decay, 1, nullptr, 0);
  h.clip_lo = clip_lo; h.clip_hi = clip_hi;
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * T, st);
  if (e != cudaSuccess) return (int)e;
  lamb_moments_kernel<<<num_chunks, kThreads, 0, st>>>((const TensorMeta*)metas, (const int2*)chunks, h, scratch);
  HB_LAUNCH_CHECK();
  lamb_apply_kernel<<<num_chunks, kThreads, 0, st>>>((const TensorMeta*)metas, (const int2*)chunks, h, scratch);
  HB_LAUNCH_CHECK();
  return 0;
}

// scratch: device double[T], zeroed here. aux = W_t (1-element fp32 state) per tensor. dof < 0 -> numel.
int hb_tadam_step(const void* metas, const void* chunks, int num_chunks, int T, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int amsgrad, float dof, int step, const int* step_dev, double* scratch,
                  void* stream) {
  if (num_chunks <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  Hyper h = make_hyper(lr, beta1, beta2, eps, weight_decay, step, step_dev, amsgrad);
  h.dof = dof;
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(double) * T, st);
  if (e != cudaSuccess) return (int)e;
  tadam_reduce_kernel<<<num_chunks, kThreads, 0, st>>>((const TensorMeta*)metas, (const int2*)chunks, h, scratch);
  HB_LAUNCH_CHECK();
  tadam_apply_kernel<<<num_chunks, kThreads, 0, st>>>((const TensorMeta*)metas, (const int2*)chunks, h, scratch);
  HB_LAUNCH_CHECK();
  tadam_wt_update_kernel<<<(T + 127) / 128, 128, 0, st>>>((const TensorMeta*)metas, T, h, scratch);
  HB_LAUNCH_CHECK();
  return 0;
}

// scratch: device double[4*T], zeroed here. delta: projection threshold (reference default 0.1).
int hb_adamp_step(const void* metas, const void* chunks, int num_chunks, int T, float lr, float beta1, float beta2, float eps,
                  float weight_decay, int amsgrad, float delta, int step, const int* step_dev, const void* ctl,
                  double* scratch, void* stream) {
  if (num_chunks <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  Hyper h = make_hyper(lr, beta1, beta2, eps, weight_decay, step, step_dev, amsgrad);
  h.delta = delta;
  h.ctl = (const float*)ctl;
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(double) * 4 * T, st);
  if (e != cudaSuccess) return (int)e;
  adamp_moments_kernel<<<num_chunks, kThreads, 0, st>>>((const TensorMeta*)metas, (const int2*)chunks, h, scratch);
  HB_LAUNCH_CHECK();
  adamp_apply_kernel<<<num_chunks, kThreads, 0, st>>>((const TensorMeta*)metas, (const int2*)chunks, h, scratch);
  HB_LAUNCH_CHECK();
  return 0;
}

// Adan: aux = prev_grad, ext = exp_avg_delta, vmax = max_exp_avg_delta (amsgrad). One launch, 40 B / parameter.
int hb_adan_step(const void* metas, const void* chunks, int num_chunks, float lr, float beta1, float beta2, float beta3,
                 float eps, float weight_decay, int amsgrad, int step, const int* step_dev, const void* ctl, void* stream) {
  if (num_chunks <= 0) return 0;
  Hyper h = make_hyper(lr, beta1, beta2, eps, weight_decay, step, step_dev, amsgrad);
  h.beta3 = beta3;
  h.bc3 = (float)(1.0 - pow((double)beta3, (double)step));
  h.ctl = (const float*)ctl;
  adan_kernel<<<num_chunks, kThreads, 0, (cudaStream_t)stream>>>((const TensorMeta*)metas, (const int2*)chunks, h);
  HB_LAUNCH_CHECK();
  return 0;
}

// AdEMAMix: ext = exp_avg_slow. One launch, 36 B / parameter.
int hb_ademamix_step(const void* metas, const void* chunks, int num_chunks, float lr, float beta1, float beta2, float beta3,
                     float alpha, float eps, float weight_decay, int step, const int* step_dev, const void* ctl,
                     void* stream) {
  if (num_chunks <= 0) return 0;
  Hyper h = make_hyper(lr, beta1, beta2, eps, weight_decay, step, step_dev, 0);
  h.beta3 = beta3; h.alpha = alpha;
  h.ctl = (const float*)ctl;
  ademamix_kernel<<<num_chunks, kThreads, 0, (cudaStream_t)stream>>>((const TensorMeta*)metas, (const int2*)chunks, h);
  HB_LAUNCH_CHECK();
  return 0;
}

// LARS: m = momentum buffer (null when momentum == 0); first != 0: the buffers are being created by this step.
// scratch: device double[2*T], zeroed here. With weight decay the gradient tensors are overwritten by g + wd * p.
int hb_lars_step(const void* metas, const void* chunks, int num_chunks, int T, float lr, float momentum, float dampening,
                 float weight_decay, int nesterov, int first, double* scratch, void* stream) {
  if (num_chunks <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  Hyper h = make_hyper(lr, 0.f, 0.f, 0.f, weight_decay, 1, nullptr, 0);
  h.momentum = momentum; h.dampening = dampening; h.nesterov = nesterov; h.first = first;
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * T, st);
  if (e != cudaSuccess) return (int)e;
  lars_norms_kernel<<<num_chunks, kThreads, 0, st>>>((const TensorMeta*)metas, (const int2*)chunks, scratch);
  HB_LAUNCH_CHECK();
  lars_apply_kernel<<<num_chunks, kThreads, 0, st>>>((const TensorMeta*)metas, (const int2*)chunks, h, scratch);
  HB_LAUNCH_CHECK();
  return 0;
}

// RaLars: mode 0 rectified (update x r_t), 1 plain Adam ratio, 2 unadapted momentum; aux = local_lr out (1 element).
// scratch: device double[2*T], zeroed here.
int hb_ralars_step(const void* metas, const void* chunks, int num_chunks, int T, float lr, float beta1, float beta2, float eps,
                   float weight_decay, float clip_lo, float clip_hi, int mode, float r_t, int step, double* scratch,
                   void* stream) {
  if (num_chunks <= 0) return 0;
  if (mode < 0 || mode > 2) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  Hyper h = make_hyper(lr, beta1, beta2, eps, weight_decay, step, nullptr, 0);
  h.clip_lo = clip_lo; h.clip_hi = clip_hi;
  h.mode = mode; h.r_t = mode == 0 ? r_t : 1.f;
  cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * T, st);
  if (e != cudaSuccess) return (int)e;
  ralars_moments_kernel<<<num_chunks, kThreads, 0, st>>>((const TensorMeta*)metas, (const int2*)chunks, h, scratch);
  HB_LAUNCH_CHECK();
  ralars_apply_kernel<<<num_chunks, kThreads, 0, st>>>((const TensorMeta*)metas, (const int2*)chunks, h, scratch);
  HB_LAUNCH_CHECK();
  return 0;
}

// Lookahead.sync_params: p = fast weights, m = slow weights
int hb_lookahead_sync(const void* metas, const void* chunks, int num_chunks, float sync_rate, void* stream) {
  if (num_chunks <= 0) return 0;
  lookahead_sync_kernel<<<num_chunks, kThreads, 0, (cudaStream_t)stream>>>((const TensorMeta*)metas, (const int2*)chunks,
                                                                          sync_rate);
  HB_LAUNCH_CHECK();
  return 0;
}

// ctl (optional): control block of a captured training step - the counter only advances when the update is not skipped
int hb_step_increment(int* step_dev, const void* ctl, void* stream) {
  step_increment_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev, (const int*)ctl);
  HB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
