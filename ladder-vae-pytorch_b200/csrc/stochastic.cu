// Fused diagonal-Gaussian stochastic block (everything between the conv_in_* outputs and the
// conv_out input of NormalStochasticBlock2d, lib/stochastic.py:45-96 and kl_normal_mc :209-226):
// reparameterised sample (external eps or Philox), log p(z), log q(z), MC or analytic KL per
// sample, analytic KL per pixel -- and (lvae_kl_bookkeeping) the free-bits clamp and the KL / log p
// bookkeeping of LadderVAE.forward (models/lvae.py:192-198,301-302) over all layers in one launch.
// Pure HBM-bound: 20 B per latent element forward, 40 B backward.
//
// Layout: q, p are (B, hw, 2Z) fp32 rows = [mu(Z) | logvar(Z)] (NHWC conv outputs);
// p may be batch-broadcast (p_bstride = 0: learned top prior, lvae_layers.py:131-136).
//
// Round-2 redesign (the round-1 kernel ran at 22 % / 42 % of the copy bandwidth at 16x16, issue-bound):
//  * grid = (pixel chunks, samples) instead of one CTA per sample, so that small batches of large latents (CelebA: 64 x
//    32x32) still fill 148 SMs; the per-sample sums of a multi-chunk launch are combined by the LAST CTA of each sample
//    in chunk order (deterministic: no floating-point atomics), through a caller-provided workspace;
//  * the hot case (Z = 32, q present, reparameterised sample) has its own software-pipelined kernel (stoch_fwd_z32_kernel);
//  * 4 MUFU per latent element instead of ~275 instructions: sigma_q = ex2(lv_q * log2e/2), 1/sigma_p^2 = ex2(-lv_p *
//    log2e), the variance ratio is sigma_q^2 / sigma_p^2 (no third exponential, no divisions), log sigma = lv/2 (the
//    reference's log(exp(lv/2)) round trip, stochastic.py:45-46, differs from it by <= 1 ulp), log q(z) of a reparameterised
//    sample is -eps^2/2 - lv_q/2 - log(2 pi)/2, Box-Muller on lg2/sqrt/sin/cos.approx.  All inside the 1e-4 (forward) and
//    1e-3 (gradient) fp32 bounds of tests/test_kernels_gpu.py::test_stochastic_core against the fp64 oracle;
//  * the backward reads and writes 16 bytes per thread; dL/dz arrives as the fp32 tensor the conv_out data gradient wrote
//    (its tcgen05 epilogue stores fp32 for this one consumer: no pad / slice / cast passes in between).
#include "common.cuh"
#include <stdlib.h>

#define HALF_LOG_2PI 0.91893853320467274178f
#define LOG2E 1.4426950408889634f

namespace {

__device__ __forceinline__ float lg2a(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sqrta(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sina(float x) { float y; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float cosa(float x) { float y; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// 4 standard normals (Box-Muller on the approximate MUFU functions: the argument of sin / cos stays inside [-pi, pi],
// where their absolute error is 2^-21; the noise only has to be N(0,1), tests/test_kernels_gpu.py checks its moments)
__device__ __forceinline__ float4 philox_normal4_fast(const PhiloxState& st, unsigned long long stream, unsigned long long idx) {
  float4 u = philox_uniform4(st, stream, idx);
  const float r0 = sqrta(-1.3862943611198906f * lg2a(u.x)), r1 = sqrta(-1.3862943611198906f * lg2a(u.z));   // sqrt(-2 ln u)
  const float a0 = 6.283185307179586f * u.y - 3.141592653589793f, a1 = 6.283185307179586f * u.w - 3.141592653589793f;
  return make_float4(r0 * cosa(a0), r0 * sina(a0), r1 * cosa(a1), r1 * sina(a1));
}

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
  int chunk_pix, nchunk;   // pixels per CTA, CTAs per sample
  float* ws_part;          // (B, nchunk, 3) partial sums (nchunk > 1)
  unsigned int* ws_cnt;    // (B) arrival tickets, zero between launches
};

constexpr int ST_THREADS = 256;

// Generic kernel: any Z, forced latent / mode / sampling from the prior.  TRAIN = q present, Philox noise, no forced latent, no
// mode: a template flag so that this instantiation carries none of the other cases' branches and fits 64 registers (4 CTAs
// per SM; the all-cases one needs 112).
template <int VEC, bool TRAIN>
__global__ void __launch_bounds__(ST_THREADS, TRAIN ? 4 : 2) stoch_fwd_kernel(StochArgs a) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[3][ST_THREADS / 32];
  __shared__ unsigned int s_last;
  if (TRAIN) { a.forced = nullptr; a.eps = nullptr; a.use_mode = 0; }
  const int b = blockIdx.y;
  const int ZV = a.Z / VEC;
  int G = 1;
  while (G < ZV && G < 32) G <<= 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = ST_THREADS >> 5;
  const int ppw = 32 / G;                 // pixels per warp per iteration
  const int gl = lane % G, gp = lane / G;
  const bool sampled = !a.eps && !a.forced && !a.use_mode;
  const bool reparam = !a.forced && !a.use_mode;       // z = mu + sigma * e with e known exactly
  PhiloxState st;
  if (sampled) st = *a.rng;
  const float* qb = (TRAIN || a.q) ? a.q + (long long)b * a.hw * 2 * a.Z : nullptr;
  const float* pb = a.p + (long long)b * a.p_bstride;
  const int pix_begin = blockIdx.x * a.chunk_pix;
  const int pix_end = min(a.hw, pix_begin + a.chunk_pix);
  float s_kl = 0.f, s_lp = 0.f, s_lq = 0.f;
  for (int pix0 = pix_begin + warp * ppw; pix0 < pix_end; pix0 += nwarp * ppw) {
    const int pix = pix0 + gp;
    const bool pvalid = pix < pix_end;
    float kls = 0.f;
    if (pvalid) {
      for (int cv = gl; cv < ZV; cv += G) {
        const int c = cv * VEC;
        const long long row = (long long)pix * 2 * a.Z;
        const long long zi = ((long long)b * a.hw + pix) * a.Z + c;
        float mq[VEC], lq[VEC], mp[VEC], lp[VEC], e[VEC], zz[VEC];
        if (VEC == 4) {
          float4 t;
          t = __ldg(reinterpret_cast<const float4*>(pb + row + c)); mp[0] = t.x; mp[1] = t.y; mp[2] = t.z; mp[3] = t.w;
          t = __ldg(reinterpret_cast<const float4*>(pb + row + a.Z + c)); lp[0] = t.x; lp[1] = t.y; lp[2] = t.z; lp[3] = t.w;
          if (qb) {
            t = __ldg(reinterpret_cast<const float4*>(qb + row + c)); mq[0] = t.x; mq[1] = t.y; mq[2] = t.z; mq[3] = t.w;
            t = __ldg(reinterpret_cast<const float4*>(qb + row + a.Z + c)); lq[0] = t.x; lq[1] = t.y; lq[2] = t.z; lq[3] = t.w;
          }
          if (a.forced) { t = __ldg(reinterpret_cast<const float4*>(a.forced + zi)); zz[0] = t.x; zz[1] = t.y; zz[2] = t.z; zz[3] = t.w; }
          else if (a.eps) { t = __ldg(reinterpret_cast<const float4*>(a.eps + zi)); e[0] = t.x; e[1] = t.y; e[2] = t.z; e[3] = t.w; }
          else if (!a.use_mode) { t = philox_normal4_fast(st, a.stream_id, (unsigned long long)(zi >> 2)); e[0] = t.x; e[1] = t.y; e[2] = t.z; e[3] = t.w; }
        } else {
          mp[0] = pb[row + c]; lp[0] = pb[row + a.Z + c];
          if (qb) { mq[0] = qb[row + c]; lq[0] = qb[row + a.Z + c]; }
          if (a.forced) zz[0] = a.forced[zi];
          else if (a.eps) e[0] = a.eps[zi];
          else if (!a.use_mode) {
            float4 t = philox_normal4_fast(st, a.stream_id, (unsigned long long)(zi >> 2));
            int k = (int)(zi & 3);
            e[0] = k == 0 ? t.x : (k == 1 ? t.y : (k == 2 ? t.z : t.w));
          }
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float smu = qb ? mq[j] : mp[j], slv = qb ? lq[j] : lp[j];
          const float ss = ex2_approx(slv * (0.5f * LOG2E));            // sigma of the sampling distribution
          float zv;
          if (a.forced) zv = zz[j];
          else if (a.use_mode) zv = smu;
          else zv = fmaf(ss, e[j], smu);
          zz[j] = zv;
          // log N(z; mu, sigma) = -(z - mu)^2 / (2 sigma^2) - log sigma - log(2 pi)/2, log sigma = lv / 2
          const float ivp = qb ? ex2_approx(-lp[j] * LOG2E) : 0.f;      // 1 / sigma_p^2
          const float dp = zv - mp[j];
          float logp;
          if (qb) logp = -0.5f * dp * dp * ivp - 0.5f * lp[j] - HALF_LOG_2PI;
          else logp = (reparam ? -0.5f * e[j] * e[j] : (a.use_mode ? 0.f : -0.5f * dp * dp * ex2_approx(-lp[j] * LOG2E))) - 0.5f * lp[j] - HALF_LOG_2PI;
          s_lp += logp;
          if (qb) {
            float logq;
            if (reparam) logq = -0.5f * e[j] * e[j] - 0.5f * lq[j] - HALF_LOG_2PI;
            else {
              const float dq = zv - mq[j];
              logq = -0.5f * dq * dq * ex2_approx(-lq[j] * LOG2E) - 0.5f * lq[j] - HALF_LOG_2PI;
            }
            s_lq += logq;
            const float vr = ss * ss * ivp;                              // sigma_q^2 / sigma_p^2
            const float dm = mq[j] - mp[j];
            const float kl_an = 0.5f * (vr + dm * dm * ivp - 1.f - (lq[j] - lp[j]));
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
      if (VEC == 4) { for (int cz = a.Z + 4 * gl; cz < a.z_lp_pitch; cz += 4 * G) *reinterpret_cast<uint2*>(zr + cz) = make_uint2(0u, 0u); }
      else { for (int cz = a.Z + gl; cz < a.z_lp_pitch; cz += G) zr[cz] = __float2bfloat16(0.f); }
    }
    // per-pixel channel reduction inside the lane group
    for (int o = G >> 1; o > 0; o >>= 1) kls += __shfl_xor_sync(0xffffffffu, kls, o);
    if (pvalid && gl == 0 && a.kl_spatial) a.kl_spatial[(long long)b * a.hw + pix] = kls;
  }
  // the three per-sample sums in one block reduction (results valid in thread 0)
  float v_lp = warp_sum(s_lp), v_lq = warp_sum(s_lq), v_kl = warp_sum(s_kl);
  if (lane == 0) { red[0][warp] = v_lp; red[1][warp] = v_lq; red[2][warp] = v_kl; }
  __syncthreads();
  if (warp == 0) {
    v_lp = warp_sum(lane < nwarp ? red[0][lane] : 0.f);
    v_lq = warp_sum(lane < nwarp ? red[1][lane] : 0.f);
    v_kl = warp_sum(lane < nwarp ? red[2][lane] : 0.f);
  }
  if (a.nchunk == 1) {
    if (threadIdx.x == 0) {
      a.logp[b] = v_lp;
      if (a.q && a.logq) a.logq[b] = v_lq;
      if (a.q && a.kl_sample) a.kl_sample[b] = v_kl;
    }
    return;
  }
  // several CTAs per sample: publish this chunk's partial sums; the last CTA of the sample adds them up in chunk order
  if (threadIdx.x == 0) {
    float* part = a.ws_part + ((long long)b * a.nchunk + blockIdx.x) * 3;
    part[0] = v_lp; part[1] = v_lq; part[2] = v_kl;
    __threadfence();
    s_last = atomicAdd(a.ws_cnt + b, 1u);
  }
  __syncthreads();
  if (s_last != (unsigned int)(a.nchunk - 1)) return;
  if (threadIdx.x == 0) {
    __threadfence();
    const volatile float* part = a.ws_part + (long long)b * a.nchunk * 3;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    for (int c = 0; c < a.nchunk; ++c) { t0 += part[3 * c]; t1 += part[3 * c + 1]; t2 += part[3 * c + 2]; }
    a.logp[b] = t0;
    if (a.q && a.logq) a.logq[b] = t1;
    if (a.q && a.kl_sample) a.kl_sample[b] = t2;
    a.ws_cnt[b] = 0u;                      // re-armed for the next launch (stream order)
  }
}



// ------------------------------------------------------------------------------------------
// The hot case as its own kernel (q present, Z = 32, reparameterised sample from Philox or external eps): every thread owns
// one 16-byte channel group of PASSES pixels of one (sample, chunk) work item and software-pipelines its loads: the four
// 16-byte loads of pass i+1 are issued before pass i is computed, so loads stay in flight during the ~500 instructions of
// compute per pass (the generic kernel issues its loads and then computes with nothing in flight: ncu issue slots 42 %,
// DRAM 21 %; more CTAs per SM alone took it from 17.0 to 13.3 us at 16x16).
// ------------------------------------------------------------------------------------------
template <bool PHILOX, int PASSES>
__global__ void __launch_bounds__(ST_THREADS, 4) stoch_fwd_z32_kernel(StochArgs a) {
  pdl_wait();
  pdl_launch();
  __shared__ float red[3][ST_THREADS / 32];
  __shared__ unsigned int s_last;
  constexpr int Z = 32, PPP = ST_THREADS / 8;             // pixels per pass
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gl = lane & 7, px = threadIdx.x >> 3;
  const int b = blockIdx.y;
  const int pix_begin = blockIdx.x * a.chunk_pix;
  const int pix_end = min(a.hw, pix_begin + a.chunk_pix);
  PhiloxState st;
  if (PHILOX) st = *a.rng;
  const float* qb = a.q + (long long)b * a.hw * 2 * Z;
  const float* pb = a.p + (long long)b * a.p_bstride;
  float4 mq4, lq4, mp4, lp4, e4;
  auto load = [&](int pix) {
    const long long row = (long long)pix * 2 * Z + 4 * gl;
    mq4 = __ldg(reinterpret_cast<const float4*>(qb + row));
    lq4 = __ldg(reinterpret_cast<const float4*>(qb + row + Z));
    mp4 = __ldg(reinterpret_cast<const float4*>(pb + row));
    lp4 = __ldg(reinterpret_cast<const float4*>(pb + row + Z));
    if (!PHILOX) e4 = __ldg(reinterpret_cast<const float4*>(a.eps + ((long long)b * a.hw + pix) * Z + 4 * gl));
  };
  float s_kl = 0.f, s_lp = 0.f, s_lq = 0.f;
  int pix = pix_begin + px;
  if (pix < pix_end) load(pix);
#pragma unroll
  for (int ps = 0; ps < PASSES; ++ps, pix += PPP) {
    const bool pvalid = pix < pix_end;                  // warp-uniform per 4-pixel group; whole warps stay in the loop (shuffles)
    const float mq[4] = {mq4.x, mq4.y, mq4.z, mq4.w}, lq[4] = {lq4.x, lq4.y, lq4.z, lq4.w};
    const float mp[4] = {mp4.x, mp4.y, mp4.z, mp4.w}, lp[4] = {lp4.x, lp4.y, lp4.z, lp4.w};
    float e[4] = {e4.x, e4.y, e4.z, e4.w};
    if (ps + 1 < PASSES && pix + PPP < pix_end) load(pix + PPP);      // next pass in flight during this one's math
    float kls = 0.f;
    if (pvalid) {
      const long long zi = ((long long)b * a.hw + pix) * Z + 4 * gl;
      if (PHILOX) {
        const float4 t = philox_normal4_fast(st, a.stream_id, (unsigned long long)(zi >> 2));
        e[0] = t.x; e[1] = t.y; e[2] = t.z; e[3] = t.w;
      }
      float zz[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float ss = ex2_approx(lq[j] * (0.5f * LOG2E));
        const float zv = fmaf(ss, e[j], mq[j]);
        zz[j] = zv;
        const float ivp = ex2_approx(-lp[j] * LOG2E);
        const float dp = zv - mp[j];
        const float logp = -0.5f * dp * dp * ivp - 0.5f * lp[j] - HALF_LOG_2PI;
        const float logq = -0.5f * e[j] * e[j] - 0.5f * lq[j] - HALF_LOG_2PI;
        s_lp += logp;
        s_lq += logq;
        const float vr = ss * ss * ivp;
        const float dm = mq[j] - mp[j];
        const float kl_an = 0.5f * (vr + dm * dm * ivp - 1.f - (lq[j] - lp[j]));
        kls += kl_an;
        s_kl += a.analytical ? kl_an : (logq - logp);
      }
      *reinterpret_cast<float4*>(a.z + zi) = make_float4(zz[0], zz[1], zz[2], zz[3]);
      if (a.z_lp) {
        __nv_bfloat16* zr = (__nv_bfloat16*)a.z_lp + ((long long)b * a.hw + pix) * a.z_lp_pitch;
        st4<__nv_bfloat16>(zr + 4 * gl, make_float4(zz[0], zz[1], zz[2], zz[3]));
        for (int cz = Z + 4 * gl; cz < a.z_lp_pitch; cz += 32) *reinterpret_cast<uint2*>(zr + cz) = make_uint2(0u, 0u);
      }
    }
    kls += __shfl_xor_sync(0xffffffffu, kls, 4);
    kls += __shfl_xor_sync(0xffffffffu, kls, 2);
    kls += __shfl_xor_sync(0xffffffffu, kls, 1);
    if (pvalid && gl == 0 && a.kl_spatial) a.kl_spatial[(long long)b * a.hw + pix] = kls;
  }
  float v_lp = warp_sum(s_lp), v_lq = warp_sum(s_lq), v_kl = warp_sum(s_kl);
  if (lane == 0) { red[0][warp] = v_lp; red[1][warp] = v_lq; red[2][warp] = v_kl; }
  __syncthreads();
  if (warp == 0) {
    v_lp = warp_sum(lane < ST_THREADS / 32 ? red[0][lane] : 0.f);
    v_lq = warp_sum(lane < ST_THREADS / 32 ? red[1][lane] : 0.f);
    v_kl = warp_sum(lane < ST_THREADS / 32 ? red[2][lane] : 0.f);
  }
  if (a.nchunk == 1) {
    if (threadIdx.x == 0) {
      a.logp[b] = v_lp;
      if (a.logq) a.logq[b] = v_lq;
      if (a.kl_sample) a.kl_sample[b] = v_kl;
    }
    return;
  }
  if (threadIdx.x == 0) {
    float* part = a.ws_part + ((long long)b * a.nchunk + blockIdx.x) * 3;
    part[0] = v_lp; part[1] = v_lq; part[2] = v_kl;
    __threadfence();
    s_last = atomicAdd(a.ws_cnt + b, 1u);
    if (s_last == (unsigned int)(a.nchunk - 1)) {
      __threadfence();
      const volatile float* pp = a.ws_part + (long long)b * a.nchunk * 3;
      float t0 = 0.f, t1 = 0.f, t2 = 0.f;
      for (int c = 0; c < a.nchunk; ++c) { t0 += pp[3 * c]; t1 += pp[3 * c + 1]; t2 += pp[3 * c + 2]; }
      a.logp[b] = t0;
      if (a.logq) a.logq[b] = t1;
      if (a.kl_sample) a.kl_sample[b] = t2;
      a.ws_cnt[b] = 0u;
    }
  }
}

}  // namespace

// workspace of a multi-chunk launch: B arrival tickets, then B * 64 * 3 floats of partial sums.  Zero-initialised ONCE by
// the caller and used with ONE batch size B (the tickets must stay zero between launches); launches that share it must be
// stream-ordered
LVAE_API long long lvae_stoch_ws_bytes(int B) { return (long long)B * (64 * 3 * 4 + 4); }

LVAE_API int lvae_stoch_fwd(const float* q, const float* p, int p_broadcast, const float* eps, const float* forced,
                            const void* rng_state, unsigned long long stream_id, float* z, void* z_bf16, int z_bf16_pitch,
                            float* kl_sample, float* kl_spatial, float* logp, float* logq, int B, int hw, int Z,
                            int use_mode, int analytical, void* ws, cudaStream_t stream) {
  LVAE_REQUIRE(p && z && logp && B > 0 && hw > 0 && Z > 0, "stoch_fwd: bad args");
  LVAE_REQUIRE(eps || forced || use_mode || rng_state, "stoch_fwd: need eps, forced latent, mode, or an RNG state");
  LVAE_REQUIRE(!z_bf16 || (z_bf16_pitch >= Z && z_bf16_pitch % 4 == 0), "stoch_fwd: bad low-precision pitch");
  StochArgs a{q, p, p_broadcast ? 0LL : (long long)hw * 2 * Z, eps, forced, (const PhiloxState*)rng_state, stream_id,
              z, z_bf16, z_bf16_pitch, kl_sample, kl_spatial, logp, logq, B, hw, Z, use_mode, analytical, hw, 1, nullptr, nullptr};
  // pixels one CTA covers per pass of its 8 warps; a chunk = one pass (every thread handles one 16-byte group of channels:
  // all loads of the launch are independent and in flight at once), unless that would need more than 64 chunks per
  // sample or there is no workspace for the cross-CTA sums
  const int vec = Z % 4 == 0 ? 4 : 1;
  int G = 1;
  while (G < Z / vec && G < 32) G <<= 1;
  const int per_pass = (ST_THREADS / 32) * (32 / G);
  if (ws && hw > per_pass) {
    int chunk = per_pass;
    while ((hw + chunk - 1) / chunk > 64) chunk *= 2;
    a.chunk_pix = chunk;
    a.nchunk = (hw + chunk - 1) / chunk;
    a.ws_cnt = (unsigned int*)ws;
    a.ws_part = (float*)ws + B;
  }
  // hot case: Z = 32, q present, reparameterised sample -> the software-pipelined kernel, PASSES pixels per thread.  Measured
  // at batch 256 (us, 16x16 / 8x8): generic 17.1 / 5.9, PASSES 1: 15.2 / 5.7, 2: 13.3 / 4.6, 4: 12.3 / 5.0
  if (q && !forced && !use_mode && Z == 32) {
    const int ppp = ST_THREADS / 8;
    int passes = hw >= 8 * ppp ? 4 : (hw >= 2 * ppp ? 2 : 1);
    if (!ws || hw <= ppp) passes = (hw + ppp - 1) / ppp;             // one CTA per sample
    if (passes != 1 && passes != 2 && passes != 4) passes = passes > 4 ? 0 : (passes == 3 ? 4 : passes);
    if (passes && (hw + ppp * passes - 1) / (ppp * passes) <= 64 && (ws || hw <= ppp * passes)) {
      a.chunk_pix = ppp * passes;
      a.nchunk = (hw + a.chunk_pix - 1) / a.chunk_pix;
      if (a.nchunk > 1) { a.ws_cnt = (unsigned int*)ws; a.ws_part = (float*)ws + B; }
      dim3 g(a.nchunk, B);
#define LVAE_Z32(PH, PS) lvae_launch(stoch_fwd_z32_kernel<PH, PS>, g, ST_THREADS, 0, stream, a)
      if (eps) { if (passes == 1) LVAE_Z32(false, 1); else if (passes == 2) LVAE_Z32(false, 2); else LVAE_Z32(false, 4); }
      else { if (passes == 1) LVAE_Z32(true, 1); else if (passes == 2) LVAE_Z32(true, 2); else LVAE_Z32(true, 4); }
#undef LVAE_Z32
      LVAE_COUNT_LAUNCH();
      LVAE_CHECK_LAUNCH("stoch_fwd (z32)");
      return LVAE_OK;
    }
  }
  dim3 grid(a.nchunk, B);
  const bool train = q && !eps && !forced && !use_mode;
  if (vec == 4 && train) lvae_launch(stoch_fwd_kernel<4, true>, grid, ST_THREADS, 0, stream, a);
  else if (vec == 4) lvae_launch(stoch_fwd_kernel<4, false>, grid, ST_THREADS, 0, stream, a);
  else lvae_launch(stoch_fwd_kernel<1, false>, grid, ST_THREADS, 0, stream, a);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("stoch_fwd");
  return LVAE_OK;
}

// ------------------------------------------------------------------------------------------
// backward.  Inputs: upstream grads g_z (B,hw,Z) [null = 0], g_kl (B), g_logp (B), g_logq (B), g_kls (B,hw) [any may be
// null = 0].  Outputs dq, dp (B,hw,2Z) (dp is per sample; the
// caller sums over the batch when p was broadcast).  eps is never needed: sigma_q*eps == z - mu_q.
// ------------------------------------------------------------------------------------------
namespace {

struct StochBwdArgs {
  const float* q; const float* p; long long p_bstride; const float* z;
  const float* g_z;
  const float* g_kl; const float* g_logp; const float* g_logq; const float* g_kls;
  float* dq; float* dp;
  __nv_bfloat16* dq_lp; __nv_bfloat16* dp_lp;   // optional bf16 copies (same layout): the operands of the tcgen05 dgrad / wgrad of conv_in_q / conv_in_p
  int B, hw, Z, analytical, z_is_sample;  // z_is_sample: 1 rsample, 0 forced latent (no dz/dq path), 2 mode (dz/dmu only)
  long long nvec;                          // B * hw * Z / VEC
};

__device__ __forceinline__ uint2 pack_bf16x4(float a, float b, float c, float d) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}

template <int VEC>
__global__ void __launch_bounds__(256) stoch_bwd_kernel(StochBwdArgs a) {
  pdl_wait();
  pdl_launch();
  const int ZV = a.Z / VEC;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.nvec; i += (long long)gridDim.x * blockDim.x) {
    const long long pixg = i / ZV;                         // global pixel index b * hw + pix
    const int c = (int)(i - pixg * ZV) * VEC;
    const int b = (int)(pixg / a.hw);
    const int pix = (int)(pixg - (long long)b * a.hw);
    const long long rowq = pixg * 2 * a.Z;
    const long long rowp = (long long)b * a.p_bstride + (long long)pix * 2 * a.Z;
    float mq[VEC], lq[VEC], mp[VEC], lp[VEC], zv[VEC], gz[VEC];
    if (VEC == 4) {
      float4 t;
      t = __ldg(reinterpret_cast<const float4*>(a.q + rowq + c)); mq[0] = t.x; mq[1] = t.y; mq[2] = t.z; mq[3] = t.w;
      t = __ldg(reinterpret_cast<const float4*>(a.q + rowq + a.Z + c)); lq[0] = t.x; lq[1] = t.y; lq[2] = t.z; lq[3] = t.w;
      t = __ldg(reinterpret_cast<const float4*>(a.p + rowp + c)); mp[0] = t.x; mp[1] = t.y; mp[2] = t.z; mp[3] = t.w;
      t = __ldg(reinterpret_cast<const float4*>(a.p + rowp + a.Z + c)); lp[0] = t.x; lp[1] = t.y; lp[2] = t.z; lp[3] = t.w;
      t = __ldg(reinterpret_cast<const float4*>(a.z + pixg * a.Z + c)); zv[0] = t.x; zv[1] = t.y; zv[2] = t.z; zv[3] = t.w;
      gz[0] = gz[1] = gz[2] = gz[3] = 0.f;
      if (a.g_z) { t = __ldg(reinterpret_cast<const float4*>(a.g_z + pixg * a.Z + c)); gz[0] = t.x; gz[1] = t.y; gz[2] = t.z; gz[3] = t.w; }
    } else {
      mq[0] = a.q[rowq + c]; lq[0] = a.q[rowq + a.Z + c]; mp[0] = a.p[rowp + c]; lp[0] = a.p[rowp + a.Z + c];
      zv[0] = a.z[pixg * a.Z + c];
      gz[0] = a.g_z ? a.g_z[pixg * a.Z + c] : 0.f;
    }
    const float gkl = a.g_kl ? __ldg(a.g_kl + b) : 0.f, glp = a.g_logp ? __ldg(a.g_logp + b) : 0.f;
    const float glq = a.g_logq ? __ldg(a.g_logq + b) : 0.f;
    const float gks = a.g_kls ? __ldg(a.g_kls + pixg) : 0.f;
    // coefficients on log q(z) and log p(z)
    const float cq = glq + (a.analytical ? 0.f : gkl);
    const float cp = glp - (a.analytical ? 0.f : gkl);
    const float ga = gks + (a.analytical ? gkl : 0.f);      // coefficient on the analytic KL
    float dmq[VEC], dlq[VEC], dmp[VEC], dlp[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float ivq = ex2_approx(-lq[j] * LOG2E), ivp = ex2_approx(-lp[j] * LOG2E);     // 1/sigma^2
      const float aq = zv[j] - mq[j], ap = zv[j] - mp[j];
      // d/dz of (cq*logq + cp*logp) plus upstream
      const float gzt = gz[j] - cq * aq * ivq - cp * ap * ivp;
      // direct parameter terms
      float tmq = cq * aq * ivq, tlq = cq * (0.5f * aq * aq * ivq - 0.5f);
      float tmp_ = cp * ap * ivp, tlp = cp * (0.5f * ap * ap * ivp - 0.5f);
      // analytic KL terms (always for kl_spatial, for kl_sample when analytical)
      if (ga != 0.f) {
        const float vr = ex2_approx((lq[j] - lp[j]) * LOG2E), dm = mq[j] - mp[j];
        tmq += ga * dm * ivp;
        tmp_ -= ga * dm * ivp;
        tlq += ga * 0.5f * (vr - 1.f);
        tlp += ga * 0.5f * (1.f - vr - dm * dm * ivp);
      }
      // reparameterisation path z = mu_q + sigma_q*eps
      if (a.z_is_sample == 1) { tmq += gzt; tlq += gzt * 0.5f * aq; }
      else if (a.z_is_sample == 2) { tmq += gzt; }
      dmq[j] = tmq; dlq[j] = tlq; dmp[j] = tmp_; dlp[j] = tlp;
    }
    if (VEC == 4) {
      *reinterpret_cast<float4*>(a.dq + rowq + c) = make_float4(dmq[0], dmq[1], dmq[2], dmq[3]);
      *reinterpret_cast<float4*>(a.dq + rowq + a.Z + c) = make_float4(dlq[0], dlq[1], dlq[2], dlq[3]);
      *reinterpret_cast<float4*>(a.dp + rowq + c) = make_float4(dmp[0], dmp[1], dmp[2], dmp[3]);
      *reinterpret_cast<float4*>(a.dp + rowq + a.Z + c) = make_float4(dlp[0], dlp[1], dlp[2], dlp[3]);
      if (a.dq_lp) {
        *reinterpret_cast<uint2*>(a.dq_lp + rowq + c) = pack_bf16x4(dmq[0], dmq[1], dmq[2], dmq[3]);
        *reinterpret_cast<uint2*>(a.dq_lp + rowq + a.Z + c) = pack_bf16x4(dlq[0], dlq[1], dlq[2], dlq[3]);
      }
      if (a.dp_lp) {
        *reinterpret_cast<uint2*>(a.dp_lp + rowq + c) = pack_bf16x4(dmp[0], dmp[1], dmp[2], dmp[3]);
        *reinterpret_cast<uint2*>(a.dp_lp + rowq + a.Z + c) = pack_bf16x4(dlp[0], dlp[1], dlp[2], dlp[3]);
      }
    } else {
      a.dq[rowq + c] = dmq[0]; a.dq[rowq + a.Z + c] = dlq[0];
      a.dp[rowq + c] = dmp[0]; a.dp[rowq + a.Z + c] = dlp[0];
      if (a.dq_lp) { a.dq_lp[rowq + c] = __float2bfloat16(dmq[0]); a.dq_lp[rowq + a.Z + c] = __float2bfloat16(dlq[0]); }
      if (a.dp_lp) { a.dp_lp[rowq + c] = __float2bfloat16(dmp[0]); a.dp_lp[rowq + a.Z + c] = __float2bfloat16(dlp[0]); }
    }
  }
}

}  // namespace

LVAE_API int lvae_stoch_bwd_ex(const float* q, const float* p, int p_broadcast, const float* z, const float* g_z,
                               const float* g_kl, const float* g_logp, const float* g_logq, const float* g_kls, float* dq,
                               float* dp, void* dq_bf16, void* dp_bf16, int B, int hw, int Z, int analytical, int z_kind,
                               cudaStream_t stream);

LVAE_API int lvae_stoch_bwd(const float* q, const float* p, int p_broadcast, const float* z, const float* g_z,
                            const float* g_kl, const float* g_logp, const float* g_logq, const float* g_kls, float* dq,
                            float* dp, int B, int hw, int Z, int analytical, int z_kind, cudaStream_t stream) {
  return lvae_stoch_bwd_ex(q, p, p_broadcast, z, g_z, g_kl, g_logp, g_logq, g_kls, dq, dp, nullptr, nullptr, B, hw, Z, analytical,
                           z_kind, stream);
}

// dq_bf16 / dp_bf16 (optional): bf16 copies of dq / dp in the same (B,hw,2Z) layout, rounded to nearest even like a cast of
// the fp32 result -- the tcgen05 dgrad / wgrad of conv_in_q / conv_in_p then read them without a conversion pass.
LVAE_API int lvae_stoch_bwd_ex(const float* q, const float* p, int p_broadcast, const float* z, const float* g_z,
                               const float* g_kl, const float* g_logp, const float* g_logq, const float* g_kls, float* dq,
                               float* dp, void* dq_bf16, void* dp_bf16, int B, int hw, int Z, int analytical, int z_kind,
                               cudaStream_t stream) {
  LVAE_REQUIRE(q && p && z && dq && dp && B > 0, "stoch_bwd: bad args");
  const int vec = Z % 4 == 0 ? 4 : 1;
  StochBwdArgs a{q, p, p_broadcast ? 0LL : (long long)hw * 2 * Z, z, g_z,
                 g_kl, g_logp, g_logq, g_kls, dq, dp, (__nv_bfloat16*)dq_bf16, (__nv_bfloat16*)dp_bf16,
                 B, hw, Z, analytical, z_kind, (long long)B * hw * Z / vec};
  const int grid = (int)min((long long)lvae_num_sms() * 8, (a.nvec + 255) / 256);
  if (vec == 4) lvae_launch(stoch_bwd_kernel<4>, grid, 256, 0, stream, a);
  else lvae_launch(stoch_bwd_kernel<1>, grid, 256, 0, stream, a);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("stoch_bwd");
  return LVAE_OK;
}

// ------------------------------------------------------------------------------------------
// KL / free-bits bookkeeping of LadderVAE.forward (models/lvae.py:192-198) and the log p(z) total of topdown_pass
// (:301-302) over the (L,B) matrices whose rows the stochastic kernels of the L layers wrote:
//   kl_sep[b] = sum_l kl[l,b];  kl = mean_b kl_sep;  kl_avg_layerwise[l] = mean_b kl[l,b];
//   kl_loss = sum_l mean_b max(kl[l,b], free_bits)   (boilr free_bits_kl: per-sample, per-layer clamp, then the batch
//   mean; free_bits < 1e-6 switches the clamp off);  logp = sum_l mean_b logp[l,b].
// coef[l,b] = d kl_loss / d kl[l,b] = (kl[l,b] >= free_bits or clamp off) / B is kept for the backward.
// One CTA: the matrices are a few thousand floats; per-layer sums are warp-shuffle + block reductions in a fixed order.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kl_bookkeeping_kernel(const float* __restrict__ kl, const float* __restrict__ lp,
                                                             int L, int B, float free_bits, float* __restrict__ kl_sep,
                                                             float* __restrict__ scalars, float* __restrict__ kl_avg,
                                                             float* __restrict__ coef) {
  pdl_wait();
  pdl_launch();
  // One warp per layer (layers warp, warp + 8, ...): lane-strided partial sums over the batch and one shuffle tree per statistic,
  // no block-wide barrier inside the layer loop (36-45 of them in a row made this one-CTA launch 25-60 us long); the eight
  // warps' totals meet once at the end, in a fixed order.
  __shared__ float tot[8][3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float invB = 1.f / (float)B;
  const bool clamp = free_bits >= 1e-6f;
  float tot_kl = 0.f, tot_loss = 0.f, tot_lp = 0.f;
  for (int l = warp; l < L; l += 8) {
    float s = 0.f, sf = 0.f, sp = 0.f;
    for (int b = lane; b < B; b += 32) {
      const float v = kl[(long long)l * B + b];
      const bool pass = !clamp || v >= free_bits;          // torch.clamp(min): gradient 1 where kl >= free_bits
      s += v;
      sf += pass ? v : free_bits;
      if (coef) coef[(long long)l * B + b] = pass ? invB : 0.f;
      if (lp) sp += lp[(long long)l * B + b];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      sf += __shfl_xor_sync(0xffffffffu, sf, o);
      sp += __shfl_xor_sync(0xffffffffu, sp, o);
    }
    if (lane == 0) kl_avg[l] = s * invB;
    tot_kl += s * invB; tot_loss += sf * invB; tot_lp += sp * invB;
  }
  if (lane == 0) { tot[warp][0] = tot_kl; tot[warp][1] = tot_loss; tot[warp][2] = tot_lp; }
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float s = 0.f;
    for (int l = 0; l < L; ++l) s += kl[(long long)l * B + b];
    kl_sep[b] = s;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += tot[w][threadIdx.x];
    scalars[threadIdx.x] = t;
  }
}

// scalars: float[3] = {kl, kl_loss, logp}; logp_rows / coef may be NULL
LVAE_API int lvae_kl_bookkeeping(const float* kl_rows, const float* logp_rows, int L, int B, float free_bits, float* kl_sep,
                                 float* scalars, float* kl_avg_layerwise, float* coef, cudaStream_t stream) {
  LVAE_REQUIRE(kl_rows && kl_sep && scalars && kl_avg_layerwise && L > 0 && B > 0, "kl_bookkeeping: bad args");
  lvae_launch(kl_bookkeeping_kernel, 1, 256, 0, stream, kl_rows, logp_rows, L, B, free_bits, kl_sep, scalars, kl_avg_layerwise, coef);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("kl_bookkeeping");
  return LVAE_OK;
}

// g_kl[l,b] = g_loss * coef[l,b] + g_klmean / B + g_sep[b] + g_avg[l] / B;  g_lp[l,b] = g_logp / B
// g_scalars: device float[3] = upstream gradients of {kl, kl_loss, logp} (entries of absent gradients are 0)
__global__ void kl_bookkeeping_bwd_kernel(const float* __restrict__ coef, const float* __restrict__ g_scalars,
                                          const float* __restrict__ g_sep, const float* __restrict__ g_avg, int L, int B,
                                          float* __restrict__ g_kl, float* __restrict__ g_lp) {
  pdl_wait();
  pdl_launch();
  const float invB = 1.f / (float)B;
  const float gk = g_scalars[0] * invB, gl = g_scalars[1], gp = g_scalars[2] * invB;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L * B; i += gridDim.x * blockDim.x) {
    const int l = i / B, b = i - l * B;
    g_kl[i] = gl * coef[i] + gk + (g_sep ? g_sep[b] : 0.f) + (g_avg ? g_avg[l] * invB : 0.f);
    if (g_lp) g_lp[i] = gp;
  }
}

LVAE_API int lvae_kl_bookkeeping_bwd(const float* coef, const float* g_scalars, const float* g_kl_sep, const float* g_kl_avg,
                                     int L, int B, float* g_kl_rows, float* g_logp_rows, cudaStream_t stream) {
  LVAE_REQUIRE(coef && g_scalars && g_kl_rows && L > 0 && B > 0, "kl_bookkeeping_bwd: bad args");
  lvae_launch(kl_bookkeeping_bwd_kernel, cdiv((long long)L * B, 256), 256, 0, stream, coef, g_scalars, g_kl_sep, g_kl_avg, L, B,
              g_kl_rows, g_logp_rows);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("kl_bookkeeping_bwd");
  return LVAE_OK;
}
