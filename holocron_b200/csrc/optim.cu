// Multi-tensor optimizer steps: AdaBelief, LAMB, TAdam (reference holocron/optim/{adabelief,lamb,tadam}.py).
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
  float* aux;   // TAdam: W_t (1 element); LAMB: local_lr out (1 element); else null
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
  h.ctl = nullptr;
  return h;
}

}  // namespace

extern "C" {

// metas: device array of T records {p, g, m, v, vmax, aux, numel} (7 x 8 bytes); chunks: device int2[num_chunks]
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

// ctl (optional): control block of a captured training step - the counter only advances when the update is not skipped
int hb_step_increment(int* step_dev, const void* ctl, void* stream) {
  step_increment_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev, (const int*)ctl);
  HB_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
