// Host-step glue that the reference runs as Python loops over ~1700 tensors
// (experiment/experiment_manager.py:78-80 Adamax, :346-350 L2 norm) and the importance-weighted
// bound's logsumexp (boilr test_procedure, call site evaluate.py:30), as single flat kernels.
#include "common.cuh"

// ---- Adamax over a flat parameter arena (torch.optim.Adamax semantics) ----
// step_count: device int64, incremented here so the captured graph advances on replay.
// 16-byte accesses (four parameters per thread and iteration: the scalar version kept ~32 KB per SM in flight, short of what
// HBM3e needs, 85-95 us for the 14.5 M parameters of the CIFAR-15 model) and, with L2, the sum of squares of the UPDATED
// parameters for the global L2 norm of experiment_manager.py:346-350 in the same pass (it used to be a second read of the arena).
template <bool L2>
__global__ void __launch_bounds__(256) adamax_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                     float* __restrict__ u, long long n, float lr, float b1, float b2, float eps,
                                                     float wd, const long long* __restrict__ step_count, float grad_scale,
                                                     const float* __restrict__ hyper, double* __restrict__ l2_acc) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[32];
  if (hyper) { lr = hyper[0]; wd = hyper[1]; }     // device-resident learning rate / weight decay: a captured graph follows them
  const double t = (double)(*step_count);
  const float clr = lr / (float)(1.0 - pow((double)b1, t));
  float ss = 0.f;
  auto upd = [&](float gi, float pi, float& mi, float& ui) {
    gi *= grad_scale;
    if (wd != 0.f) gi += wd * pi;
    mi = b1 * mi + (1.f - b1) * gi;
    ui = fmaxf(b2 * ui, fabsf(gi) + eps);
    const float pn = pi - clr * (mi / ui);
    if (L2) ss += pn * pn;
    return pn;
  };
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i], u4 = reinterpret_cast<float4*>(u)[i];
    p4.x = upd(g4.x, p4.x, m4.x, u4.x);
    p4.y = upd(g4.y, p4.y, m4.y, u4.y);
    p4.z = upd(g4.z, p4.z, m4.z, u4.z);
    p4.w = upd(g4.w, p4.w, m4.w, u4.w);
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(u)[i] = u4;
    reinterpret_cast<float4*>(p)[i] = p4;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float mi = m[i], ui = u[i];
    p[i] = upd(g[i], p[i], mi, ui);
    m[i] = mi;
    u[i] = ui;
  }
  if (L2) {
    ss = block_sum(ss, red);
    if (threadIdx.x == 0) atomicAdd(l2_acc, (double)ss);
  }
}
__global__ void step_inc_kernel(long long* s) {
  pdl_wait();
  pdl_launch(); *s += 1; }
__global__ void sqrt_finalize_kernel(double* acc, float* out);

static int adamax_launch(float* p, const float* g, float* exp_avg, float* exp_inf, long long n, float lr, float beta1, float beta2,
                         float eps, float weight_decay, long long* step_count_dev, float grad_scale, const float* hyper_dev,
                         double* l2_acc, float* l2_out, cudaStream_t stream) {
  LVAE_REQUIRE(p && g && exp_avg && exp_inf && step_count_dev && n > 0, "adamax_step: bad args");
  LVAE_REQUIRE((((size_t)p | (size_t)g | (size_t)exp_avg | (size_t)exp_inf) & 15) == 0, "adamax_step: the arenas must be 16-byte aligned");
  lvae_launch(step_inc_kernel, 1, 1, 0, stream, step_count_dev);
  LVAE_COUNT_LAUNCH();
  int grid = (int)min((long long)8 * lvae_num_sms(), ((n >> 2) + 255) / 256 + 1);
  if (l2_acc)
    lvae_launch(adamax_kernel<true>, grid, 256, 0, stream, p, g, exp_avg, exp_inf, n, lr, beta1, beta2, eps, weight_decay,
                step_count_dev, grad_scale, hyper_dev, l2_acc);
  else
    lvae_launch(adamax_kernel<false>, grid, 256, 0, stream, p, g, exp_avg, exp_inf, n, lr, beta1, beta2, eps, weight_decay,
                step_count_dev, grad_scale, hyper_dev, l2_acc);
  LVAE_COUNT_LAUNCH();
  if (l2_acc) {
    lvae_launch(sqrt_finalize_kernel, 1, 1, 0, stream, l2_acc, l2_out);
    LVAE_COUNT_LAUNCH();
  }
  LVAE_CHECK_LAUNCH("adamax_step");
  return LVAE_OK;
}

LVAE_API int lvae_adamax_step(float* p, const float* g, float* exp_avg, float* exp_inf, long long n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, long long* step_count_dev,
                              float grad_scale, const float* hyper_dev, cudaStream_t stream) {
  return adamax_launch(p, g, exp_avg, exp_inf, n, lr, beta1, beta2, eps, weight_decay, step_count_dev, grad_scale, hyper_dev,
                       nullptr, nullptr, stream);
}

// The same step, and l2_out[0] = sqrt(sum p^2) of the updated parameters (what lvae_l2_norm would return right after it).
// l2_acc: device double scratch, zero on entry, cleared again on exit.
LVAE_API int lvae_adamax_step_l2(float* p, const float* g, float* exp_avg, float* exp_inf, long long n, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, long long* step_count_dev,
                                 float grad_scale, const float* hyper_dev, double* l2_acc, float* l2_out, cudaStream_t stream) {
  LVAE_REQUIRE(l2_acc && l2_out, "adamax_step_l2: null pointer");
  return adamax_launch(p, g, exp_avg, exp_inf, n, lr, beta1, beta2, eps, weight_decay, step_count_dev, grad_scale, hyper_dev,
                       l2_acc, l2_out, stream);
}

// ---- global L2 norm of a flat arena: out[0] = sqrt(sum p^2) ----
__global__ void sumsq_kernel(const float* __restrict__ p, long long n, double* __restrict__ acc) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[32];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = p[i];
    s += v * v;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(acc, (double)s);
}
__global__ void sqrt_finalize_kernel(double* acc, float* out) {
  pdl_wait();
  pdl_launch();
  *out = (float)sqrt(*acc);
  *acc = 0.0;
}

// acc: device double scratch (zero on entry; cleared again on exit)
LVAE_API int lvae_l2_norm(const float* p, long long n, double* acc, float* out, cudaStream_t stream) {
  LVAE_REQUIRE(p && acc && out && n > 0, "l2_norm: bad args");
  int grid = (int)min((long long)4 * lvae_num_sms(), (n + 255) / 256);
  lvae_launch(sumsq_kernel, grid, 256, 0, stream, p, n, acc);
  LVAE_COUNT_LAUNCH();
  lvae_launch(sqrt_finalize_kernel, 1, 1, 0, stream, acc, out);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("l2_norm");
  return LVAE_OK;
}

// ---- importance-weighted bound: streaming logsumexp over samples, mergeable across ranks ----
// state (B,2) = (running max m, running sum s of exp(elbo - m)); elbo = ll - kl per image.
__global__ void iw_update_kernel(const float* __restrict__ ll, const float* __restrict__ kl, float* __restrict__ state,
                                 int B, int first) {
  pdl_wait();
  pdl_launch();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float e = ll[i] - kl[i];
  if (first) {
    state[2 * i] = e;
    state[2 * i + 1] = 1.f;
  } else {
    float m = state[2 * i], s = state[2 * i + 1];
    float nm = fmaxf(m, e);
    state[2 * i] = nm;
    state[2 * i + 1] = s * expf(m - nm) + expf(e - nm);
  }
}

LVAE_API int lvae_iw_lse_update(const float* ll, const float* kl, float* state, int B, int first, cudaStream_t stream) {
  LVAE_REQUIRE(ll && kl && state && B > 0, "iw_lse_update: bad args");
  lvae_launch(iw_update_kernel, cdiv(B, 256), 256, 0, stream, ll, kl, state, B, first);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("iw_lse_update");
  return LVAE_OK;
}

// states: (R,B,2) gathered from R ranks (R = 1 on one GPU); out[b] = logsumexp over all K samples - log K
__global__ void iw_combine_kernel(const float* __restrict__ states, float* __restrict__ out, int R, int B, float logK) {
  pdl_wait();
  pdl_launch();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float m = -INFINITY;
  for (int r = 0; r < R; ++r) m = fmaxf(m, states[((long long)r * B + i) * 2]);
  float s = 0.f;
  for (int r = 0; r < R; ++r) s += states[((long long)r * B + i) * 2 + 1] * expf(states[((long long)r * B + i) * 2] - m);
  out[i] = m + logf(s) - logK;
}

LVAE_API int lvae_iw_lse_combine(const float* states, float* out, int R, int B, int K_total, cudaStream_t stream) {
  LVAE_REQUIRE(states && out && R > 0 && B > 0 && K_total > 0, "iw_lse_combine: bad args");
  lvae_launch(iw_combine_kernel, cdiv(B, 256), 256, 0, stream, states, out, R, B, logf((float)K_total));
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("iw_lse_combine");
  return LVAE_OK;
}
