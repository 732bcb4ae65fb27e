// Shared device/host helpers for the lvae_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#define LVAE_API extern "C" __attribute__((visibility("default")))

// ---- error plumbing (C-ABI: int return code, message via lvae_last_error) ----
void lvae_set_error(const char* fmt, ...);
#define LVAE_OK 0
#define LVAE_ERR_ARG 1
#define LVAE_ERR_CUDA 2
#define LVAE_ERR_UNSUPPORTED 3

#define LVAE_REQUIRE(cond, ...)                      \
  do {                                               \
    if (!(cond)) {                                   \
      lvae_set_error(__VA_ARGS__);                   \
      return LVAE_ERR_ARG;                           \
    }                                                \
  } while (0)

#define LVAE_CHECK_LAUNCH(name)                                               \
  do {                                                                        \
    cudaError_t e__ = cudaGetLastError();                                     \
    if (e__ != cudaSuccess) {                                                 \
      lvae_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return LVAE_ERR_CUDA;                                                   \
    }                                                                         \
  } while (0)

// ---- programmatic dependent launch (PDL) ----
// Every kernel of this library is launched with programmaticStreamSerializationAllowed: its CTAs may be
// scheduled while the previous kernel in the stream is still draining.  A kernel must not touch memory
// written by its predecessor before pdl_wait(); it calls pdl_launch() right after, so the NEXT kernel can
// only start once this one has seen its predecessor complete (anything two or more kernels upstream is
// therefore safe to read in a prologue: packed weights, parameters, tensor maps).
extern int g_lvae_pdl;
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t lvae_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                      Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_lvae_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// launch counter (bench.py reports gpu_launches from it)
extern unsigned long long g_lvae_launches;
#define LVAE_COUNT_LAUNCH() (++g_lvae_launches)

// ---- activation ids (nonlin strings of models/lvae.py:64-69) ----
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_ELU = 3, ACT_SELU = 4 };

#define SELU_ALPHA 1.6732632423543772848170429916717f
#define SELU_SCALE 1.0507009873554804934193349852946f

// FAST (bf16 activations): ex2.approx-based exponentials, ~2 instructions instead of the ~40 of expm1f / expf;
// their 1e-7 absolute error is far below bf16 rounding.  The fp32 parity mode keeps the exact functions.
__device__ __forceinline__ float ex2_approx(float x) {        // one MUFU.EX2, no range fix-ups (flushes denormals)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <bool FAST> __device__ __forceinline__ float exp_t(float v) { return FAST ? ex2_approx(v * 1.4426950408889634f) : expf(v); }
template <bool FAST> __device__ __forceinline__ float expm1_t(float v) { return FAST ? ex2_approx(v * 1.4426950408889634f) - 1.f : expm1f(v); }
template <bool FAST>
__device__ __forceinline__ float act_fwd_t(float v, int act) {
  switch (act) {
    case ACT_RELU: return v > 0.f ? v : 0.f;
    case ACT_LEAKY: return v > 0.f ? v : 0.01f * v;
    case ACT_ELU: return v > 0.f ? v : expm1_t<FAST>(v);
    case ACT_SELU: return SELU_SCALE * (v > 0.f ? v : SELU_ALPHA * expm1_t<FAST>(v));
    default: return v;
  }
}
// derivative wrt the pre-activation v
template <bool FAST>
__device__ __forceinline__ float act_bwd_t(float v, int act) {
  switch (act) {
    case ACT_RELU: return v > 0.f ? 1.f : 0.f;
    case ACT_LEAKY: return v > 0.f ? 1.f : 0.01f;
    case ACT_ELU: return v > 0.f ? 1.f : exp_t<FAST>(v);
    case ACT_SELU: return v > 0.f ? SELU_SCALE : SELU_SCALE * SELU_ALPHA * exp_t<FAST>(v);
    default: return 1.f;
  }
}
__device__ __forceinline__ float act_fwd(float v, int act) { return act_fwd_t<false>(v, act); }
__device__ __forceinline__ float act_bwd(float v, int act) { return act_bwd_t<false>(v, act); }

template <bool FAST> __device__ __forceinline__ float sigmoid_t(float v) {
  return FAST ? __fdividef(1.f, 1.f + ex2_approx(-1.4426950408889634f * v)) : 1.f / (1.f + expf(-v));
}
__device__ __forceinline__ float sigmoidf_(float v) { return sigmoid_t<false>(v); }
// ONE MUFU instead of two (ex2 + rcp): sigmoid(v) = 0.5 tanh(v / 2) + 0.5 on tanh.approx (relative error 2^-11, i.e. an absolute
// error <= 2.5e-4 on the gate value -- below the 2^-9 rounding of the bf16 tensor it is multiplied into).  For the gate passes of
// the bf16 convolution epilogues, whose 128 x 64 sigmoids + ELUs per tile make them MUFU-bound.
__device__ __forceinline__ float sigmoid_tanh_approx(float v) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * v));
  return fmaf(t, 0.5f, 0.5f);
}
// torch softplus (beta 1, threshold 20)
__device__ __forceinline__ float softplusf_(float v) { return v > 20.f ? v : log1pf(expf(v)); }

// ---- typed element IO (activations are fp32 or bf16, math is fp32) ----
template <typename T> __device__ __forceinline__ float ld1(const T* p);
template <> __device__ __forceinline__ float ld1<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st1(T* p, float v);
template <> __device__ __forceinline__ void st1<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st1<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

template <typename T> __device__ __forceinline__ float4 ld4(const T* p);
template <> __device__ __forceinline__ float4 ld4<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x), b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T> __device__ __forceinline__ void st4(T* p, float4 v);
template <> __device__ __forceinline__ void st4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

// ---- reductions ----
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum of one float; result valid in thread 0. smem: >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* smem) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = lane < nw ? smem[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;
}

// ---- Philox4x32-10 counter RNG (stateless; seed/offset live in device memory so that a
//      captured CUDA graph draws fresh noise on every replay) ----
struct PhiloxState { unsigned long long seed, offset; };

__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// 4 uniform(0,1] floats for (stream, index)
__device__ __forceinline__ float4 philox_uniform4(const PhiloxState& st, unsigned long long stream, unsigned long long idx) {
  unsigned long long c = idx + st.offset;
  uint4 ctr = make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)stream, (uint32_t)(stream >> 32));
  uint2 key = make_uint2((uint32_t)st.seed, (uint32_t)(st.seed >> 32));
  uint4 r = philox4x32(ctr, key);
  const float s = 2.3283064365386963e-10f;  // 2^-32
  return make_float4((r.x + 1.0f) * s * 0.99999994f + 0.f, (r.y + 1.0f) * s * 0.99999994f,
                     (r.z + 1.0f) * s * 0.99999994f, (r.w + 1.0f) * s * 0.99999994f);
}
// 4 standard normals (Box-Muller)
__device__ __forceinline__ float4 philox_normal4(const PhiloxState& st, unsigned long long stream, unsigned long long idx) {
  float4 u = philox_uniform4(st, stream, idx);
  float r0 = sqrtf(-2.f * logf(fmaxf(u.x, 1e-37f))), r1 = sqrtf(-2.f * logf(fmaxf(u.z, 1e-37f)));
  float s0, c0, s1, c1;
  sincospif(2.f * u.y, &s0, &c0);
  sincospif(2.f * u.w, &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

static inline int lvae_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}
static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
