// Fused output log-likelihoods (forward + backward) and their samplers.
//   * Bernoulli on probabilities with BCE's -100 log clamp (lib/likelihoods.py:62,385-388)
//   * 10-component discretized mixture of logistics (lib/likelihoods.py:226-230,291-382),
//     per-pixel logsumexp kept in registers, parameters staged through shared memory so the
//     (B,H,W,100) NHWC rows are read/written fully coalesced.
//   * samplers: Bernoulli (likelihoods.py:73-75), DMoL (lib/stochastic.py:141-206, likelihoods.py:221-225)
// x is the user's image tensor: (B,C,H,W) NCHW fp32 in [0,1].  params are NHWC fp32.
#include "common.cuh"
#include <stdlib.h>

// =========================================================================================
// Bernoulli
// =========================================================================================
__global__ void __launch_bounds__(256) bernoulli_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ x,
                                                            float* __restrict__ prob, float* __restrict__ ll,
                                                            int hw, int C) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[32];
  const int b = blockIdx.x, n = hw * C;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int pix = i / C, c = i - pix * C;
    float p = sigmoidf_(logits[(long long)b * n + i]);
    prob[(long long)b * n + i] = p;
    if (x) {
      float xv = x[((long long)b * C + c) * hw + pix];
      float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.f - p), -100.f);
      s += xv * lp + (1.f - xv) * l1p;
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0 && ll) ll[b] = s;
}

// dlogit = g_ll[b] * (x - p) * p(1-p) / max(p(1-p), 1e-12)   (ATen BCE backward times sigmoid')
__global__ void bernoulli_bwd_kernel(const float* __restrict__ prob, const float* __restrict__ x,
                                     const float* __restrict__ g_ll, const float* __restrict__ g_prob,
                                     float* __restrict__ dlogits, int B, int hw, int C) {
  pdl_wait();
  pdl_launch();
  long long total = (long long)B * hw * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long t = i / C;
    int pix = (int)(t % hw);
    int b = (int)(t / hw);
    float p = prob[i], xv = x[((long long)b * C + c) * hw + pix];
    float pq = p * (1.f - p);
    float d = g_ll[b] * (xv - p) / fmaxf(pq, 1e-12f) * pq;
    if (g_prob) d += g_prob[i] * pq;
    dlogits[i] = d;
  }
}

LVAE_API int lvae_bernoulli_fwd(const float* logits, const float* x, float* prob, float* ll, int B, int hw, int C,
                                cudaStream_t stream) {
  LVAE_REQUIRE(logits && prob && B > 0 && hw > 0 && C > 0, "bernoulli_fwd: bad args");
  LVAE_REQUIRE((x == nullptr) == (ll == nullptr), "bernoulli_fwd: x and ll go together");
  lvae_launch(bernoulli_fwd_kernel, B, 256, 0, stream, logits, x, prob, ll, hw, C);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("bernoulli_fwd");
  return LVAE_OK;
}

LVAE_API int lvae_bernoulli_bwd(const float* prob, const float* x, const float* g_ll, const float* g_prob,
                                float* dlogits, int B, int hw, int C, cudaStream_t stream) {
  LVAE_REQUIRE(prob && x && g_ll && dlogits, "bernoulli_bwd: bad args");
  long long n = (long long)B * hw * C;
  int grid = (int)min((long long)4 * lvae_num_sms(), (n + 255) / 256);
  lvae_launch(bernoulli_bwd_kernel, grid, 256, 0, stream, prob, x, g_ll, g_prob, dlogits, B, hw, C);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("bernoulli_bwd");
  return LVAE_OK;
}

__global__ void bernoulli_sample_kernel(const float* prob, float* out, long long n, int hw, int C,
                                        const PhiloxState* rng, unsigned long long stream_id) {
  pdl_wait();
  pdl_launch();
  // out is NCHW like every image the module API hands back
  PhiloxState st = *rng;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float4 u = philox_uniform4(st, stream_id, (unsigned long long)i);
    int c = (int)(i % C);
    long long t = i / C;
    int pix = (int)(t % hw);
    long long b = t / hw;
    out[(b * C + c) * hw + pix] = (u.x - 5.9e-8f) < prob[i] ? 1.f : 0.f;
  }
}

LVAE_API int lvae_bernoulli_sample(const float* prob, float* out_nchw, int B, int hw, int C, const void* rng_state,
                                   unsigned long long stream_id, cudaStream_t stream) {
  LVAE_REQUIRE(prob && out_nchw && rng_state, "bernoulli_sample: bad args");
  long long n = (long long)B * hw * C;
  int grid = (int)min((long long)4 * lvae_num_sms(), (n + 255) / 256);
  lvae_launch(bernoulli_sample_kernel, grid, 256, 0, stream, prob, out_nchw, n, hw, C, (const PhiloxState*)rng_state, stream_id);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("bernoulli_sample");
  return LVAE_OK;
}

// =========================================================================================
// Discretized mixture of logistics, 10 components, 3 colour channels
// =========================================================================================
constexpr int DM_M = 10, DM_P = 100, DM_PITCH = 107, DM_TILE = 72;   // pitch % 32 = 11: the three pixel rows of a warp hit disjoint banks
#define LOG_127_5 4.8481163519437300f

// Single-MUFU exponentials / logarithms (ex2.approx, lg2.approx; relative error 2^-22) instead of the ~10-25-instruction
// exact expf / logf / log1pf / tanhf: the kernel is issue-bound on those (1376 / 2120 SASS instructions with the exact
// functions, 704 / 1080 with these).  Validated on the B200 against the fp64 oracle inside the fp32 bounds of the parity
// tests (1e-4 on ll, 1e-3 on the gradient: tests/test_kernels_gpu.py::test_dmol, tests/test_model_gpu.py) and 1.7x / 1.5x
// faster (forward 76 -> 45 us, backward 114 -> 76 us for the CIFAR batch of 256; profiles/ab_r02_summary.txt).
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float dm_exp(float v) { return exp_t<true>(v); }
__device__ __forceinline__ float dm_log(float v) { return lg2_approx(v) * 0.6931471805599453f; }
__device__ __forceinline__ float dm_softplus(float v) { return v > 20.f ? v : dm_log(1.f + dm_exp(v)); }
__device__ __forceinline__ float dm_tanh(float v) { return 1.f - __fdividef(2.f, 1.f + dm_exp(2.f * v)); }
__device__ __forceinline__ float dm_sigmoid(float v) { return sigmoid_t<true>(v); }

// log-prob of one sub-pixel under one logistic; optionally its derivatives wrt the centred value
// and the (clamped) log-scale.  Mirrors likelihoods.py:331-375.
__device__ __forceinline__ float dmol_term(float x, float cen, float ls, float& d_cen, float& d_ls, bool want_grad) {
  float inv = dm_exp(-ls);
  float plus_in = inv * (cen + (1.f / 255.f));
  float min_in = inv * (cen - (1.f / 255.f));
  if (x < -0.999f) {
    if (want_grad) { float om = 1.f - dm_sigmoid(plus_in); d_cen = inv * om; d_ls = -plus_in * om; }
    return plus_in - dm_softplus(plus_in);
  }
  if (x > 0.999f) {
    if (want_grad) { float s = dm_sigmoid(min_in); d_cen = -inv * s; d_ls = min_in * s; }
    return -dm_softplus(min_in);
  }
  // both logistic CDFs from ONE reciprocal: cp = 1 / (1 + e_p), cm = 1 / (1 + e_m), r = 1 / ((1 + e_p)(1 + e_m)).  The exponents
  // are clamped to +-30 (sigmoid is 0 / 1 to fp32 resolution beyond that) so that the product cannot overflow.
  const float e_p = dm_exp(-fminf(fmaxf(plus_in, -30.f), 30.f)), e_m = dm_exp(-fminf(fmaxf(min_in, -30.f), 30.f));
  const float r = __fdividef(1.f, (1.f + e_p) * (1.f + e_m));
  const float cp = r * (1.f + e_m), cm = r * (1.f + e_p);
  const float delta = r * (e_m - e_p);
  if (delta > 1e-5f) {
    if (want_grad) {
      float dpv = cp * (1.f - cp), dmv = cm * (1.f - cm);
      float dd = fmaxf(delta, 1e-12f);
      d_cen = inv * __fdividef(dpv - dmv, dd);
      d_ls = -__fdividef(plus_in * dpv - min_in * dmv, dd);
    }
    return dm_log(fmaxf(delta, 1e-12f));
  }
  float mid = inv * cen;
  if (want_grad) { float w = 1.f - 2.f * dm_sigmoid(mid); d_cen = inv * w; d_ls = -mid * w - 1.f; }
  return mid - ls - 2.f * dm_softplus(mid) - LOG_127_5;
}

// BWD = false: ll[b] += sum over this CTA's pixels.  BWD = true: dl = g_ll[b] * d ll / d l.
// A CTA stages DM_TILE pixels x 100 parameters in shared memory (coalesced 400-byte rows), then TEN lanes work on one
// pixel -- a warp covers three pixels, lanes 30 and 31 idle along -- lane m owning mixture component m: its three per-colour
// log-probabilities (and, backward, its ten parameter gradients, written back over the staged parameters) are independent
// of the other components; the two logsumexps over the components are guarded 4-step shuffle trees inside the 10-lane
// segment plus one broadcast.  History: one thread per pixel left the MUFU-heavy per-component math (30 logistic terms) in
// one long dependent chain per thread (265 us for the CIFAR batch-256 backward); 16 lanes per pixel with 6 of them idle ran
// the same arithmetic with 10x the parallelism but wasted 37 % of the issue slots of an issue-bound kernel (76 us); this
// layout wastes 6 %.
constexpr int DM_THREADS = 256, DM_SEG = 10, DM_PIX_PER_PASS = (DM_THREADS / 32) * 3;
template <bool BWD>
__global__ void __launch_bounds__(DM_THREADS) dmol_kernel(const float* __restrict__ l, const float* __restrict__ x,
                                                          float* __restrict__ ll, const float* __restrict__ g_ll,
                                                          float* __restrict__ dl, int hw, __nv_bfloat16* __restrict__ dl_lp) {
  pdl_wait();
  pdl_launch();
  extern __shared__ float sm[];  // DM_TILE * DM_PITCH (+32 for the reduction)
  float* red = sm + DM_TILE * DM_PITCH;
  const int b = blockIdx.y;
  const int pix0 = blockIdx.x * DM_TILE;
  const int npix = min(DM_TILE, hw - pix0);
  const long long base = ((long long)b * hw + pix0) * DM_P;
  // coalesced stage-in (rows are contiguous in NHWC); base is a multiple of 4 floats
  {
    const float4* src = reinterpret_cast<const float4*>(l + base);
    int nq = npix * DM_P / 4;
    for (int i = threadIdx.x; i < nq; i += DM_THREADS) {
      float4 v = src[i];
      int e = i * 4;
      float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int r = (e + j) / DM_P, cidx = (e + j) - r * DM_P;
        sm[r * DM_PITCH + cidx] = vv[j];
      }
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int seg = lane / DM_SEG;                            // pixel of this lane within its warp (3 = the two idle lanes)
  const int m = lane - seg * DM_SEG;                        // mixture component of this lane
  const int seg_base = seg * DM_SEG;
  const int slot = (threadIdx.x >> 5) * 3 + min(seg, 2);    // pixel slot within a pass
  const bool comp = seg < 3;
  const int mm = m;
  float pix_ll = 0.f;
  // reductions over the 10 lanes of a segment: guarded shuffle-down tree (offsets 8, 4, 2, 1 leave the result in the
  // segment's first lane), then a broadcast from that lane.  All 32 lanes take part in every shuffle.
  auto gmax = [&](float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float t = __shfl_down_sync(0xffffffffu, v, o);
      if (m + o < DM_SEG) v = fmaxf(v, t);
    }
    return __shfl_sync(0xffffffffu, v, seg_base);
  };
  auto gsum = [&](float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      const float t = __shfl_down_sync(0xffffffffu, v, o);
      if (m + o < DM_SEG) v += t;
    }
    return __shfl_sync(0xffffffffu, v, seg_base);
  };
  for (int r = slot; r < ((npix + DM_PIX_PER_PASS - 1) / DM_PIX_PER_PASS) * DM_PIX_PER_PASS; r += DM_PIX_PER_PASS) {
    const bool live = comp && r < npix;                     // whole warps stay in the loop together (shuffles)
    float* L = sm + (live ? r : 0) * DM_PITCH;
    const int pix = pix0 + (live ? r : 0);
    float xc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) xc[c] = x[((long long)b * 3 + c) * hw + pix] * 2.f - 1.f;
    // log_softmax of the mixture logits
    const float logit = comp ? L[mm] : -INFINITY;
    const float mx = gmax(logit);
    const float se = gsum(comp ? dm_exp(logit - mx) : 0.f);
    const float lse0 = mx + dm_log(se);
    // this component's log-probability of the three sub-pixels
    const float c0r = L[10 + 20 + mm], c1r = L[40 + 20 + mm], c2r = L[70 + 20 + mm];
    const float k0 = dm_tanh(c0r), k1 = dm_tanh(c1r), k2 = dm_tanh(c2r);
    const float mu[3] = {L[10 + mm], L[40 + mm] + k0 * xc[0], L[70 + mm] + k1 * xc[0] + k2 * xc[1]};
    float lsr[3], dc[3], dls[3];
    float S = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      lsr[c] = L[10 + 30 * c + 10 + mm];
      S += dmol_term(xc[c], xc[c] - mu[c], fmaxf(lsr[c], -7.f), dc[c], dls[c], BWD);
    }
    const float v = comp ? S + (logit - lse0) : -INFINITY;
    const float vmax = gmax(v);
    const float sv = gsum(comp ? dm_exp(v - vmax) : 0.f);
    const float lse = vmax + dm_log(sv);
    if (live && m == 0) pix_ll += lse;
    if (BWD && live && comp) {
      const float g = g_ll[b];
      const float rm = dm_exp(v - lse);               // responsibility
      const float gS = g * rm;
      float dmu[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        dmu[c] = -gS * dc[c];                               // cen = x - mu
        L[10 + 30 * c + 10 + m] = lsr[c] >= -7.f ? gS * dls[c] : 0.f;   // clamp(min=-7) backward
      }
      L[m] = g * (rm - dm_exp(logit - lse0));
      L[10 + m] = dmu[0];
      L[40 + m] = dmu[1];
      L[70 + m] = dmu[2];
      L[10 + 20 + m] = dmu[1] * xc[0] * (1.f - k0 * k0);
      L[40 + 20 + m] = dmu[2] * xc[0] * (1.f - k1 * k1);
      L[70 + 20 + m] = dmu[2] * xc[1] * (1.f - k2 * k2);
    }
  }
  if (!BWD) {
    float s = block_sum(pix_ll, red);
    if (threadIdx.x == 0) atomicAdd(ll + b, s);
  } else {
    __syncthreads();
    if (dl_lp) {
      // bf16, 128 channels per pixel (100 gradients + 28 zeros): the operand layout of the tcgen05 dgrad / wgrad of the head conv
      __nv_bfloat16* dstp = dl_lp + ((long long)b * hw + pix0) * 128;
      for (int i = threadIdx.x; i < npix * 32; i += DM_THREADS) {       // quads of 4 channels
        int rr = i >> 5, cq = (i & 31) * 4;
        float v0 = cq < DM_P ? sm[rr * DM_PITCH + cq] : 0.f, v1 = cq + 1 < DM_P ? sm[rr * DM_PITCH + cq + 1] : 0.f;
        float v2 = cq + 2 < DM_P ? sm[rr * DM_PITCH + cq + 2] : 0.f, v3 = cq + 3 < DM_P ? sm[rr * DM_PITCH + cq + 3] : 0.f;
        st4<__nv_bfloat16>(dstp + rr * 128 + cq, make_float4(v0, v1, v2, v3));
      }
    } else {
      float4* dst = reinterpret_cast<float4*>(dl + base);
      int nq = npix * DM_P / 4;
      for (int i = threadIdx.x; i < nq; i += DM_THREADS) {
        int e = i * 4;
        float vv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int rr = (e + j) / DM_P, cidx = (e + j) - rr * DM_P;
          vv[j] = sm[rr * DM_PITCH + cidx];
        }
        dst[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
      }
    }
  }
}

static const size_t DMOL_SMEM = (DM_TILE * DM_PITCH + 32) * sizeof(float);

static void dmol_init() {
  static bool done = false;
  if (!done) {
    cudaFuncSetAttribute(dmol_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DMOL_SMEM);
    cudaFuncSetAttribute(dmol_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DMOL_SMEM);
    done = true;
  }
}

// ll must be zeroed by the caller (partial sums are accumulated with atomics, 16 per image at 32x32)
LVAE_API int lvae_dmol_fwd(const float* l, const float* x, float* ll, int B, int hw, cudaStream_t stream) {
  LVAE_REQUIRE(l && x && ll && B > 0 && hw > 0, "dmol_fwd: bad args");
  dmol_init();
  dim3 grid(cdiv(hw, DM_TILE), B);
  lvae_launch(dmol_kernel<false>, grid, DM_THREADS, DMOL_SMEM, stream, l, x, ll, nullptr, nullptr, hw, nullptr);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("dmol_fwd");
  return LVAE_OK;
}

// dl: fp32 (B,hw,100), or -- when dl_bf16_128 != NULL -- bf16 (B,hw,128) zero-padded (dl may then be NULL)
LVAE_API int lvae_dmol_bwd(const float* l, const float* x, const float* g_ll, float* dl, void* dl_bf16_128, int B, int hw,
                           cudaStream_t stream) {
  LVAE_REQUIRE(l && x && g_ll && (dl || dl_bf16_128) && B > 0 && hw > 0, "dmol_bwd: bad args");
  dmol_init();
  dim3 grid(cdiv(hw, DM_TILE), B);
  lvae_launch(dmol_kernel<true>, grid, DM_THREADS, DMOL_SMEM, stream, l, x, nullptr, g_ll, dl, hw, (__nv_bfloat16*)dl_bf16_128);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("dmol_bwd");
  return LVAE_OK;
}

// DMoL sampler: Gumbel-max over the mixture logits, logistic draw per colour, autoregressive means,
// clamp to [-1,1], rescale to [0,1].  One thread per pixel; out is (B,3,H,W) NCHW.
__global__ void dmol_sample_kernel(const float* __restrict__ l, float* __restrict__ out, long long npix, int hw,
                                   const PhiloxState* rng, unsigned long long stream_id) {
  pdl_wait();
  pdl_launch();
  PhiloxState st = *rng;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    const float* L = l + i * DM_P;
    float u[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 t = philox_uniform4(st, stream_id, (unsigned long long)(i * 4 + q));
      u[q * 4] = t.x; u[q * 4 + 1] = t.y; u[q * 4 + 2] = t.z; u[q * 4 + 3] = t.w;
    }
    int sel = 0;
    float best = -INFINITY;
#pragma unroll
    for (int m = 0; m < DM_M; ++m) {
      float uu = 1e-5f + u[m] * (1.f - 2e-5f);
      float s = L[m] - logf(-logf(uu));
      if (s > best) { best = s; sel = m; }
    }
    float xs[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float uu = 1e-5f + u[10 + c] * (1.f - 2e-5f);
      float mean = L[10 + 30 * c + sel], ls = fmaxf(L[10 + 30 * c + 10 + sel], -7.f);
      xs[c] = mean + expf(ls) * (logf(uu) - logf(1.f - uu));
    }
    float k0 = tanhf(L[10 + 20 + sel]), k1 = tanhf(L[40 + 20 + sel]), k2 = tanhf(L[70 + 20 + sel]);
    float x0 = fminf(fmaxf(xs[0], -1.f), 1.f);
    float x1 = fminf(fmaxf(xs[1] + k0 * x0, -1.f), 1.f);
    float x2 = fminf(fmaxf(xs[2] + k1 * x0 + k2 * x1, -1.f), 1.f);
    long long b = i / hw;
    int pix = (int)(i - b * hw);
    out[(b * 3 + 0) * hw + pix] = fminf(fmaxf((x0 + 1.f) * 0.5f, 0.f), 1.f);
    out[(b * 3 + 1) * hw + pix] = fminf(fmaxf((x1 + 1.f) * 0.5f, 0.f), 1.f);
    out[(b * 3 + 2) * hw + pix] = fminf(fmaxf((x2 + 1.f) * 0.5f, 0.f), 1.f);
  }
}

LVAE_API int lvae_dmol_sample(const float* l, float* out_nchw, int B, int hw, const void* rng_state,
                              unsigned long long stream_id, cudaStream_t stream) {
  LVAE_REQUIRE(l && out_nchw && rng_state, "dmol_sample: bad args");
  long long n = (long long)B * hw;
  int grid = (int)min((long long)4 * lvae_num_sms(), (n + 127) / 128);
  lvae_launch(dmol_sample_kernel, grid, 128, 0, stream, l, out_nchw, n, hw, (const PhiloxState*)rng_state, stream_id);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("dmol_sample");
  return LVAE_OK;
}
