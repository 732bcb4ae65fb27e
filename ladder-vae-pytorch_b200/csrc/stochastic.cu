// Fused diagonal-Gaussian stochastic block (everything between the conv_in_* outputs and the
// conv_out input of NormalStochasticBlock2d, lib/stochastic.py:45-96 and kl_normal_mc :209-226):
// reparameterised sample (external eps or Philox), log p(z), log q(z), MC or analytic KL per
// sample, analytic KL per pixel.  One CTA per sample; a group of G lanes owns one pixel so the
// per-pixel (channel) reduction is a warp shuffle and the per-sample one a block reduction.
// Pure HBM-bound: 20 B per latent element forward, 40 B backward.
//
// Layout: q, p are (B, hw, 2Z) fp32 rows = [mu(Z) | logvar(Z)] (NHWC conv outputs);
// p may be batch-broadcast (p_bstride = 0: learned top prior, lvae_layers.py:131-136).
#include "common.cuh"

#define HALF_LOG_2PI 0.91893853320467274178f

struct StochArgs {
  const float* q;       // may be null (generation: sample from p)
  const float* p;
  long long p_bstride;  // 0 or hw*2Z
  const float* eps;     // (B,hw,Z) or null -> Philox
  const float* forced;  // (B,hw,Z) forced latent or null
  const PhiloxState* rng;
  unsigned long long stream_id;
  float* z;             // (B,hw,Z)
  void* z_lp;           // optional bf16 copy of z (input of conv_out on the bf16 path), row pitch z_lp_pitch >= Z, zero padded
  int z_lp_pitch;
  float* kl_sample;     // (B) or null when q == null
  float* kl_spatial;    // (B,hw) or null
  float* logp;          // (B)
  float* logq;          // (B) or null
  int B, hw, Z;
  int use_mode, analytical;
};

template <int VEC>
__global__ void __launch_bounds__(256) stoch_fwd_kernel(StochArgs a) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[32];
  const int b = blockIdx.x;
  const int ZV = a.Z / VEC;
  int G = 1;
  while (G < ZV && G < 32) G <<= 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int ppw = 32 / G;                 // pixels per warp per iteration
  const int gl = lane % G, gp = lane / G;
  PhiloxState st;
  if (!a.eps && !a.forced && !a.use_mode) st = *a.rng;
  const float* qb = a.q ? a.q + (long long)b * a.hw * 2 * a.Z : nullptr;
  const float* pb = a.p + (long long)b * a.p_bstride;
  float s_kl = 0.f, s_lp = 0.f, s_lq = 0.f;
  for (int pix0 = warp * ppw; pix0 < a.hw; pix0 += nwarp * ppw) {
    int pix = pix0 + gp;
    bool pvalid = pix < a.hw;
    float kls = 0.f;
    if (pvalid) {
      for (int cv = gl; cv < ZV; cv += G) {
        int c = cv * VEC;
        long long row = (long long)pix * 2 * a.Z;
        long long zi = ((long long)b * a.hw + pix) * a.Z + c;
        float mq[VEC], lq[VEC], mp[VEC], lp[VEC], e[VEC], zz[VEC];
        if (VEC == 4) {
          float4 t;
          t = *reinterpret_cast<const float4*>(pb + row + c); mp[0] = t.x; mp[1] = t.y; mp[2] = t.z; mp[3] = t.w;
          t = *reinterpret_cast<const float4*>(pb + row + a.Z + c); lp[0] = t.x; lp[1] = t.y; lp[2] = t.z; lp[3] = t.w;
          if (qb) {
            t = *reinterpret_cast<const float4*>(qb + row + c); mq[0] = t.x; mq[1] = t.y; mq[2] = t.z; mq[3] = t.w;
            t = *reinterpret_cast<const float4*>(qb + row + a.Z + c); lq[0] = t.x; lq[1] = t.y; lq[2] = t.z; lq[3] = t.w;
          }
          if (a.forced) { t = *reinterpret_cast<const float4*>(a.forced + zi); zz[0] = t.x; zz[1] = t.y; zz[2] = t.z; zz[3] = t.w; }
          else if (a.eps) { t = *reinterpret_cast<const float4*>(a.eps + zi); e[0] = t.x; e[1] = t.y; e[2] = t.z; e[3] = t.w; }
          else if (!a.use_mode) { t = philox_normal4(st, a.stream_id, (unsigned long long)(zi >> 2)); e[0] = t.x; e[1] = t.y; e[2] = t.z; e[3] = t.w; }
        } else {
          mp[0] = pb[row + c]; lp[0] = pb[row + a.Z + c];
          if (qb) { mq[0] = qb[row + c]; lq[0] = qb[row + a.Z + c]; }
          if (a.forced) zz[0] = a.forced[zi];
          else if (a.eps) e[0] = a.eps[zi];
          else if (!a.use_mode) {
            float4 t = philox_normal4(st, a.stream_id, (unsigned long long)(zi >> 2));
            int k = (int)(zi & 3);
            e[0] = k == 0 ? t.x : (k == 1 ? t.y : (k == 2 ? t.z : t.w));
          }
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          float smu = qb ? mq[j] : mp[j], slv = qb ? lq[j] : lp[j];
          float zv;
          if (a.forced) zv = zz[j];
          else if (a.use_mode) zv = smu;
          else zv = smu + expf(slv * 0.5f) * e[j];
          zz[j] = zv;
          // log N(z; mu, exp(lv/2)) in torch.distributions' form
          float sp = expf(lp[j] * 0.5f);
          float dp = zv - mp[j];
          float logp = -(dp * dp) / (2.f * sp * sp) - logf(sp) - HALF_LOG_2PI;
          s_lp += logp;
          if (qb) {
            float sq = expf(lq[j] * 0.5f);
            float dq = zv - mq[j];
            float logq = -(dq * dq) / (2.f * sq * sq) - logf(sq) - HALF_LOG_2PI;
            s_lq += logq;
            float r = sq / sp, vr = r * r;
            float t1 = (mq[j] - mp[j]) / sp;
            float kl_an = 0.5f * (vr + t1 * t1 - 1.f - logf(vr));
            kls += kl_an;
            s_kl += a.analytical ? kl_an : (logq - logp);
          }
        }
        if (VEC == 4) {
          *reinterpret_cast<float4*>(a.z + zi) = make_float4(zz[0], zz[1], zz[2], zz[3]);
          if (a.z_lp) st4<__nv_bfloat16>((__nv_bfloat16*)a.z_lp + ((long long)b * a.hw + pix) * a.z_lp_pitch + c, make_float4(zz[0], zz[1], zz[2], zz[3]));
        } else {
          a.z[zi] = zz[0];
          if (a.z_lp) ((__nv_bfloat16*)a.z_lp)[((long long)b * a.hw + pix) * a.z_lp_pitch + c] = __float2bfloat16(zz[0]);
        }
      }
    }
    // zero padding of the low-precision copy (channels Z .. pitch)
    if (pvalid && a.z_lp) {
      __nv_bfloat16* zr = (__nv_bfloat16*)a.z_lp + ((long long)b * a.hw + pix) * a.z_lp_pitch;
      for (int cz = a.Z + gl; cz < a.z_lp_pitch; cz += G) zr[cz] = __float2bfloat16(0.f);
    }
    // per-pixel channel reduction inside the lane group
    for (int o = G >> 1; o > 0; o >>= 1) kls += __shfl_xor_sync(0xffffffffu, kls, o);
    if (pvalid && gl == 0 && a.kl_spatial) a.kl_spatial[(long long)b * a.hw + pix] = kls;
  }
  float v;
  v = block_sum(s_lp, red);
  if (threadIdx.x == 0) a.logp[b] = v;
  if (a.q) {
    v = block_sum(s_lq, red);
    if (threadIdx.x == 0 && a.logq) a.logq[b] = v;
    v = block_sum(s_kl, red);
    if (threadIdx.x == 0 && a.kl_sample) a.kl_sample[b] = v;
  }
}

LVAE_API int lvae_stoch_fwd(const float* q, const float* p, int p_broadcast, const float* eps, const float* forced,
                            const void* rng_state, unsigned long long stream_id, float* z, void* z_bf16, int z_bf16_pitch,
                            float* kl_sample, float* kl_spatial, float* logp, float* logq, int B, int hw, int Z,
                            int use_mode, int analytical, cudaStream_t stream) {
  LVAE_REQUIRE(p && z && logp && B > 0 && hw > 0 && Z > 0, "stoch_fwd: bad args");
  LVAE_REQUIRE(eps || forced || use_mode || rng_state, "stoch_fwd: need eps, forced latent, mode, or an RNG state");
  LVAE_REQUIRE(!z_bf16 || (z_bf16_pitch >= Z && z_bf16_pitch % 4 == 0), "stoch_fwd: bad low-precision pitch");
  StochArgs a{q, p, p_broadcast ? 0LL : (long long)hw * 2 * Z, eps, forced, (const PhiloxState*)rng_state, stream_id,
              z, z_bf16, z_bf16_pitch, kl_sample, kl_spatial, logp, logq, B, hw, Z, use_mode, analytical};
  if (Z % 4 == 0) lvae_launch(stoch_fwd_kernel<4>, B, 256, 0, stream, a);
  else lvae_launch(stoch_fwd_kernel<1>, B, 256, 0, stream, a);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("stoch_fwd");
  return LVAE_OK;
}

// ------------------------------------------------------------------------------------------
// backward.  Inputs: upstream grads g_z (B,hw,Z) [null = 0], g_kl (B), g_logp (B), g_logq (B),
// g_kls (B,hw) [any may be null = 0].  Outputs dq, dp (B,hw,2Z) (dp is per sample; the caller
// sums over the batch when p was broadcast).  eps is never needed: sigma_q*eps == z - mu_q.
// ------------------------------------------------------------------------------------------
struct StochBwdArgs {
  const float* q; const float* p; long long p_bstride; const float* z;
  const float* g_z; const float* g_kl; const float* g_logp; const float* g_logq; const float* g_kls;
  float* dq; float* dp;
  int B, hw, Z, analytical, z_is_sample;  // z_is_sample: 1 rsample, 0 forced latent (no dz/dq path), 2 mode (dz/dmu only)
};

__global__ void __launch_bounds__(256) stoch_bwd_kernel(StochBwdArgs a) {
  pdl_wait();
  pdl_launch();
  const int b = blockIdx.x;
  const float gkl = a.g_kl ? a.g_kl[b] : 0.f, glp = a.g_logp ? a.g_logp[b] : 0.f, glq = a.g_logq ? a.g_logq[b] : 0.f;
  const float* qb = a.q + (long long)b * a.hw * 2 * a.Z;
  const float* pb = a.p + (long long)b * a.p_bstride;
  float* dqb = a.dq + (long long)b * a.hw * 2 * a.Z;
  float* dpb = a.dp + (long long)b * a.hw * 2 * a.Z;
  const int n = a.hw * a.Z;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int pix = i / a.Z, c = i - pix * a.Z;
    long long row = (long long)pix * 2 * a.Z;
    float mq = qb[row + c], lq = qb[row + a.Z + c], mp = pb[row + c], lp = pb[row + a.Z + c];
    float zv = a.z[(long long)b * n + i];
    float gz = a.g_z ? a.g_z[(long long)b * n + i] : 0.f;
    float gks = a.g_kls ? a.g_kls[(long long)b * a.hw + pix] : 0.f;
    float ivq = expf(-lq), ivp = expf(-lp);     // 1/sigma^2
    float aq = zv - mq, ap = zv - mp;
    // coefficients on log q(z) and log p(z)
    float cq = glq + (a.analytical ? 0.f : gkl);
    float cp = glp - (a.analytical ? 0.f : gkl);
    // d/dz of (cq*logq + cp*logp) plus upstream
    float gzt = gz - cq * aq * ivq - cp * ap * ivp;
    // direct parameter terms
    float dmq = cq * aq * ivq, dlq = cq * (0.5f * aq * aq * ivq - 0.5f);
    float dmp = cp * ap * ivp, dlp = cp * (0.5f * ap * ap * ivp - 0.5f);
    // analytic KL terms (always for kl_spatial, for kl_sample when analytical)
    float ga = gks + (a.analytical ? gkl : 0.f);
    if (ga != 0.f) {
      float vr = expf(lq - lp), dm = mq - mp;
      dmq += ga * dm * ivp;
      dmp -= ga * dm * ivp;
      dlq += ga * 0.5f * (vr - 1.f);
      dlp += ga * 0.5f * (1.f - vr - dm * dm * ivp);
    }
    // reparameterisation path z = mu_q + sigma_q*eps
    if (a.z_is_sample == 1) { dmq += gzt; dlq += gzt * 0.5f * aq; }
    else if (a.z_is_sample == 2) { dmq += gzt; }
    dqb[row + c] = dmq; dqb[row + a.Z + c] = dlq;
    dpb[row + c] = dmp; dpb[row + a.Z + c] = dlp;
  }
}

LVAE_API int lvae_stoch_bwd(const float* q, const float* p, int p_broadcast, const float* z, const float* g_z,
                            const float* g_kl, const float* g_logp, const float* g_logq, const float* g_kls,
                            float* dq, float* dp, int B, int hw, int Z, int analytical, int z_kind,
                            cudaStream_t stream) {
  LVAE_REQUIRE(q && p && z && dq && dp && B > 0, "stoch_bwd: bad args");
  StochBwdArgs a{q, p, p_broadcast ? 0LL : (long long)hw * 2 * Z, z, g_z, g_kl, g_logp, g_logq, g_kls, dq, dp,
                 B, hw, Z, analytical, z_kind};
  lvae_launch(stoch_bwd_kernel, B, 256, 0, stream, a);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("stoch_bwd");
  return LVAE_OK;
}
