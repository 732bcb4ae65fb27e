// Generic implicit-GEMM convolution on CUDA cores (fp32 accumulate), NHWC activations.
//
// This is the exact-fp32 path (parity runs, odd shapes: 5x5 stem, 64->100 / 64->1 heads,
// strided / transposed 3x3).  The bf16 tcgen05 kernels in conv_tcgen05.cu take over the
// dominant 64-channel shapes.  Replaces the cuDNN calls behind nn.Conv2d / nn.ConvTranspose2d
// at lib/nn.py:83-87,118, lib/stochastic.py:25-27, models/lvae_layers.py:263-276,350,
// models/lvae.py:75 and lib/likelihoods.py:55,199 (forward, dgrad and wgrad).
#include "common.cuh"

struct ConvArgs {
  const void* x;        // (B,Hi,Wi,C1)
  const void* x2;       // optional (B,Hi,Wi,C2): channel-concatenated second input (MergeLayer)
  const void* wp;       // packed weights [kh*kw*(C1+C2)][ldw]
  const float* bias;    // [N] or null
  const float* in_scale;   // (B, C1+C2) per-sample channel scale on the input, or null
  const float* out_scale;  // (B, N) per-sample channel scale on the output (Dropout2d), or null
  const void* res;      // optional residual (B,Ho,Wo,N) added after scaling
  void* y;              // (B,Ho,Wo,N)
  int B, Hi, Wi, C1, C2, Ho, Wo, N, ldw;
  int kh, kw, stride, pad, mode;  // mode 0: iy = oy*stride - pad + ky ; mode 1: iy = (oy + pad - ky)/stride
  // mode 1 with stride 2: output pixels are enumerated parity-class-major ((oy&1, ox&1) constant per CTA), so a CTA
  // skips the filter taps that never meet its class (3/4 of a 3x3 kernel's work)
  int cls;
};

constexpr int CG_BM = 128, CG_BN = 64, CG_BK = 32, CG_THREADS = 256;
constexpr int CG_APITCH = CG_BK + 4;

// output pixel (b, oy, ox) of GEMM row m
__device__ __forceinline__ void decode_pixel(const ConvArgs& a, long long M, long long m, int& b, int& oy, int& ox) {
  if (a.cls) {
    const long long Mq = M >> 2;
    const int c = (int)(m / Mq);
    const int r = (int)(m - (long long)c * Mq);
    const int Hq = a.Ho >> 1, Wq = a.Wo >> 1;
    b = r / (Hq * Wq);
    const int r2 = r - b * (Hq * Wq);
    oy = 2 * (r2 / Wq) + (c >> 1);
    ox = 2 * (r2 % Wq) + (c & 1);
  } else {
    const int hw = a.Ho * a.Wo;
    b = (int)(m / hw);
    const int r = (int)(m - (long long)b * hw);
    oy = r / a.Wo;
    ox = r - oy * a.Wo;
  }
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(CG_THREADS) conv_gather_kernel(ConvArgs a) {
  pdl_wait();
  pdl_launch();
  __shared__ __align__(16) float As[CG_BM][CG_APITCH];
  __shared__ __align__(16) float Bs[CG_BK][CG_BN];
  const int t = threadIdx.x;
  const int Cin = a.C1 + a.C2;
  const int K = a.kh * a.kw * Cin;
  const long long M = (long long)a.B * a.Ho * a.Wo;
  const long long m0 = (long long)blockIdx.x * CG_BM;
  const int n0 = blockIdx.y * CG_BN;
  const T* x = (const T*)a.x;
  const T* x2 = (const T*)a.x2;
  const T* wp = (const T*)a.wp;

  // --- A-load bookkeeping: this thread loads rows (t/8 + 32 i), k-quad (t%8) ---
  const int kq = t & 7;
  int pb[4], py[4], px[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + (t >> 3) + 32 * i;
    if (m < M) {
      decode_pixel(a, M, m, pb[i], py[i], px[i]);
    } else {
      pb[i] = -1; py[i] = 0; px[i] = 0;
    }
  }
  // --- B-load bookkeeping: rows (t/16 + 16 i), col-quad (t%16) ---
  const int bq = t & 15;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int tm = t >> 4, tn = t & 15;
  float4 ra[4], rb[2];

  auto load_tiles = [&](int k0) {
    // A: gathered input patch
    int k = k0 + kq * 4;
    if (VEC) {
      int tap = k / Cin, ci = k - tap * Cin;
      int ky = tap / a.kw, kx = tap - ky * a.kw;
      bool kvalid = k < K;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kvalid && pb[i] >= 0) {
          int iy, ix;
          bool ok;
          if (a.mode == 0) {
            iy = py[i] * a.stride - a.pad + ky;
            ix = px[i] * a.stride - a.pad + kx;
            ok = iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi;
          } else {
            int ty = py[i] + a.pad - ky, tx = px[i] + a.pad - kx;
            ok = ty >= 0 && tx >= 0 && (ty % a.stride) == 0 && (tx % a.stride) == 0;
            iy = ty / a.stride;
            ix = tx / a.stride;
            ok = ok && iy < a.Hi && ix < a.Wi;
          }
          if (ok) {
            long long pix = ((long long)pb[i] * a.Hi + iy) * a.Wi + ix;
            v = ci < a.C1 ? ld4<T>(x + pix * a.C1 + ci) : ld4<T>(x2 + pix * a.C2 + (ci - a.C1));
            if (a.in_scale) {
              float4 s = *reinterpret_cast<const float4*>(a.in_scale + (long long)pb[i] * Cin + ci);
              v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
            }
          }
        }
        ra[i] = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float vv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int kk = k + j;
          float v = 0.f;
          if (kk < K && pb[i] >= 0) {
            int tap = kk / Cin, ci = kk - tap * Cin;
            int ky = tap / a.kw, kx = tap - ky * a.kw;
            int iy, ix;
            bool ok;
            if (a.mode == 0) {
              iy = py[i] * a.stride - a.pad + ky;
              ix = px[i] * a.stride - a.pad + kx;
              ok = iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi;
            } else {
              int ty = py[i] + a.pad - ky, tx = px[i] + a.pad - kx;
              ok = ty >= 0 && tx >= 0 && (ty % a.stride) == 0 && (tx % a.stride) == 0;
              iy = ty / a.stride;
              ix = tx / a.stride;
              ok = ok && iy < a.Hi && ix < a.Wi;
            }
            if (ok) {
              long long pix = ((long long)pb[i] * a.Hi + iy) * a.Wi + ix;
              v = ci < a.C1 ? ld1<T>(x + pix * a.C1 + ci) : ld1<T>(x2 + pix * a.C2 + (ci - a.C1));
              if (a.in_scale) v *= a.in_scale[(long long)pb[i] * Cin + ci];
            }
          }
          vv[j] = v;
        }
        ra[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
      }
    }
    // B: packed weights, rows k, ldw is a multiple of 4
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int kr = k0 + (t >> 4) + 16 * i;
      int n = n0 + bq * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kr < K && n < a.ldw) v = ld4<T>(wp + (long long)kr * a.ldw + n);
      rb[i] = v;
    }
  };

  // parity-class mode: the CTA's class and the taps that can reach it
  const int ccls = a.cls ? (int)(m0 / (M >> 2)) : 0;
  auto next_valid = [&](int k0) {
    if (a.cls) {
      while (k0 < K) {
        const int tap = k0 / Cin, ky = tap / a.kw, kx = tap - ky * a.kw;
        if ((((ccls >> 1) + a.pad + ky) & 1) == 0 && (((ccls & 1) + a.pad + kx) & 1) == 0) break;
        k0 += CG_BK;
      }
    }
    return k0;
  };
  int k0 = next_valid(0);
  if (k0 < K) load_tiles(k0);
  while (k0 < K) {
    const int kn = next_valid(k0 + CG_BK);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(&As[(t >> 3) + 32 * i][kq * 4]) = ra[i];
#pragma unroll
    for (int i = 0; i < 2; ++i) *reinterpret_cast<float4*>(&Bs[(t >> 4) + 16 * i][bq * 4]) = rb[i];
    __syncthreads();
    if (kn < K) load_tiles(kn);
#pragma unroll
    for (int k4 = 0; k4 < CG_BK; k4 += 4) {
      float4 b4[4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) b4[kk] = *reinterpret_cast<const float4*>(&Bs[k4 + kk][tn * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 a4 = *reinterpret_cast<const float4*>(&As[tm * 8 + i][k4]);
        acc[i][0] += a4.x * b4[0].x; acc[i][1] += a4.x * b4[0].y; acc[i][2] += a4.x * b4[0].z; acc[i][3] += a4.x * b4[0].w;
        acc[i][0] += a4.y * b4[1].x; acc[i][1] += a4.y * b4[1].y; acc[i][2] += a4.y * b4[1].z; acc[i][3] += a4.y * b4[1].w;
        acc[i][0] += a4.z * b4[2].x; acc[i][1] += a4.z * b4[2].y; acc[i][2] += a4.z * b4[2].z; acc[i][3] += a4.z * b4[2].w;
        acc[i][0] += a4.w * b4[3].x; acc[i][1] += a4.w * b4[3].y; acc[i][2] += a4.w * b4[3].z; acc[i][3] += a4.w * b4[3].w;
      }
    }
    k0 = kn;
  }

  // --- epilogue: (acc + bias) * out_scale + res ---
  T* y = (T*)a.y;
  const T* res = (const T*)a.res;
  const int n = n0 + tn * 4;
  if (n >= a.N) return;
  float bv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bv[j] = (a.bias && n + j < a.N) ? a.bias[n + j] : 0.f;
  const bool vec_out = (a.N & 3) == 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    long long m = m0 + tm * 8 + i;
    if (m >= M) break;
    int b, oy, ox;
    decode_pixel(a, M, m, b, oy, ox);
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j] = acc[i][j] + bv[j];
      if (a.out_scale && n + j < a.N) o[j] *= a.out_scale[(long long)b * a.N + n + j];
    }
    long long off = (((long long)b * a.Ho + oy) * a.Wo + ox) * a.N + n;
    if (vec_out) {
      if (res) {
        float4 r = ld4<T>(res + off);
        o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
      }
      st4<T>(y + off, make_float4(o[0], o[1], o[2], o[3]));
    } else {
      for (int j = 0; j < 4 && n + j < a.N; ++j) {
        float v = o[j];
        if (res) v += ld1<T>(res + off + j);
        st1<T>(y + off + j, v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// The stem: Conv2d(c -> 64, 5x5, stride 2, padding 2) on the c = 1 / 3 channel image (models/lvae.py:75).  K = 25 c is far too
// short for the 128x64x32 tiling above, whose scalar gather (c is not a multiple of 4) decodes (tap, channel) per element:
// 99 us for the (256,32,32,3) CIFAR batch, 6 TFLOP/s.  Here a CTA owns a 16x16 tile of output pixels of one image: the 35x35
// input patch under it (zero-filled outside the image = the padding) and all 25 c x 64 weights sit in shared memory as fp32,
// a thread owns two pixels (8 rows apart) x 32 output channels, and the weights reach the FMAs as warp-wide broadcasts
// (one LDS.128 per 4 x 2 FMAs).  The products are summed in the same order as in conv_gather_kernel (tap-major, channel
// inner, one fmaf each), so the result is the same to the bit.
// ------------------------------------------------------------------------------------------
constexpr int ST_TILE = 16, ST_PATCH = 2 * ST_TILE + 3;

template <typename T, int C>
__global__ void __launch_bounds__(256) conv_stem5x5s2_kernel(const T* __restrict__ x, const T* __restrict__ wp,
                                                             const float* __restrict__ bias, T* __restrict__ y, int Hi, int Wi,
                                                             int Ho, int Wo, int ldw, int tiles_x, int tiles_per_img) {
  pdl_wait();
  pdl_launch();
  constexpr int K = 25 * C;
  constexpr int XPITCH = ST_PATCH * C + 1;
  __shared__ __align__(16) float ws[K][64];
  __shared__ float xs[ST_PATCH][XPITCH];
  const int t = threadIdx.x;
  const int b = blockIdx.x / tiles_per_img;
  const int tr = blockIdx.x - b * tiles_per_img;
  const int oy0 = (tr / tiles_x) * ST_TILE, ox0 = (tr - (tr / tiles_x) * tiles_x) * ST_TILE;
  // staging: the loads of a thread are issued in batches of eight before the first shared-memory store (one round trip to L2
  // per batch instead of one per element: the 19 + 15 dependent iterations of the plain loops were two thirds of the kernel)
  const int iy0 = 2 * oy0 - 2, ix0 = 2 * ox0 - 2;
  const T* xb = x + (size_t)b * Hi * Wi * C;
  for (int base = t; base < K * 64; base += 8 * 256) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * 256;
      v[u] = i < K * 64 ? ld1<T>(wp + (size_t)(i >> 6) * ldw + (i & 63)) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * 256;
      if (i < K * 64) ws[i >> 6][i & 63] = v[u];
    }
  }
  for (int base = t; base < ST_PATCH * ST_PATCH * C; base += 8 * 256) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * 256;
      const int r = i / (ST_PATCH * C), e = i - r * (ST_PATCH * C);
      const int iy = iy0 + r, ix = ix0 + e / C;
      v[u] = 0.f;
      if (i < ST_PATCH * ST_PATCH * C && iy >= 0 && iy < Hi && ix >= 0 && ix < Wi)
        v[u] = ld1<T>(xb + ((size_t)iy * Wi + ix) * C + (e - (e / C) * C));
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base + u * 256;
      const int r = i / (ST_PATCH * C), e = i - r * (ST_PATCH * C);
      if (i < ST_PATCH * ST_PATCH * C) xs[r][e] = v[u];
    }
  }
  __syncthreads();
  const int pp = t & 127, n0 = (t >> 7) * 32;           // the channel half is warp-uniform: weight reads are broadcasts
  const int ly = pp >> 4, lx = pp & 15;
  float acc[2][32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[0][j] = acc[1][j] = 0.f;
#pragma unroll 1
  for (int ky = 0; ky < 5; ++ky) {
    const float* r0 = &xs[2 * ly + ky][2 * lx * C];
    const float* r1 = &xs[2 * (ly + 8) + ky][2 * lx * C];
#pragma unroll
    for (int kc = 0; kc < 5 * C; ++kc) {                // kc = kx * C + ci: five consecutive input pixels of the row
      const float a0 = r0[kc], a1 = r1[kc];
      const float* wr = &ws[ky * 5 * C + kc][n0];
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 w4 = *reinterpret_cast<const float4*>(wr + 4 * j4);
        acc[0][4 * j4 + 0] = fmaf(a0, w4.x, acc[0][4 * j4 + 0]); acc[1][4 * j4 + 0] = fmaf(a1, w4.x, acc[1][4 * j4 + 0]);
        acc[0][4 * j4 + 1] = fmaf(a0, w4.y, acc[0][4 * j4 + 1]); acc[1][4 * j4 + 1] = fmaf(a1, w4.y, acc[1][4 * j4 + 1]);
        acc[0][4 * j4 + 2] = fmaf(a0, w4.z, acc[0][4 * j4 + 2]); acc[1][4 * j4 + 2] = fmaf(a1, w4.z, acc[1][4 * j4 + 2]);
        acc[0][4 * j4 + 3] = fmaf(a0, w4.w, acc[0][4 * j4 + 3]); acc[1][4 * j4 + 3] = fmaf(a1, w4.w, acc[1][4 * j4 + 3]);
      }
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int oy = oy0 + ly + 8 * h, ox = ox0 + lx;
    if (oy < Ho && ox < Wo) {
      T* yp = y + (((size_t)b * Ho + oy) * Wo + ox) * 64 + n0;
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        float4 o = make_float4(acc[h][4 * j4], acc[h][4 * j4 + 1], acc[h][4 * j4 + 2], acc[h][4 * j4 + 3]);
        if (bias) {
          const float4 bv = *reinterpret_cast<const float4*>(bias + n0 + 4 * j4);
          o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
        }
        st4<T>(yp + 4 * j4, o);
      }
    }
  }
}

template <typename T>
static bool launch_stem(const void* x, const void* wp, const float* bias, void* y, int B, int Hi, int Wi, int C, int Ho, int Wo,
                        int ldw, cudaStream_t stream) {
  const int tiles_x = cdiv(Wo, ST_TILE), tiles_per_img = tiles_x * cdiv(Ho, ST_TILE);
  const long long grid = (long long)B * tiles_per_img;
  if (grid >= (1LL << 31)) return false;
  if (C == 1)
    lvae_launch(conv_stem5x5s2_kernel<T, 1>, (int)grid, 256, 0, stream, (const T*)x, (const T*)wp, bias, (T*)y, Hi, Wi, Ho, Wo, ldw, tiles_x, tiles_per_img);
  else if (C == 3)
    lvae_launch(conv_stem5x5s2_kernel<T, 3>, (int)grid, 256, 0, stream, (const T*)x, (const T*)wp, bias, (T*)y, Hi, Wi, Ho, Wo, ldw, tiles_x, tiles_per_img);
  else
    return false;
  return true;
}

LVAE_API int lvae_conv2d_gather(const void* x, const void* x2, const void* wp, const float* bias,
                                const float* in_scale, const float* out_scale, const void* res, void* y,
                                int B, int Hi, int Wi, int C1, int C2, int Ho, int Wo, int N, int ldw,
                                int kh, int kw, int stride, int pad, int mode, int dtype, cudaStream_t stream) {
  LVAE_REQUIRE(x && wp && y, "conv2d_gather: null pointer");
  LVAE_REQUIRE(B > 0 && Hi > 0 && Wi > 0 && C1 > 0 && C2 >= 0 && Ho > 0 && Wo > 0 && N > 0, "conv2d_gather: bad shape");
  LVAE_REQUIRE((C2 == 0) == (x2 == nullptr), "conv2d_gather: x2/C2 mismatch");
  LVAE_REQUIRE(ldw % 4 == 0 && ldw >= N, "conv2d_gather: ldw must be a multiple of 4 and >= N");
  LVAE_REQUIRE(mode == 0 || mode == 1, "conv2d_gather: bad mode");
  LVAE_REQUIRE(dtype == 0 || dtype == 1, "conv2d_gather: dtype must be 0 (f32) or 1 (bf16)");
  long long M = (long long)B * Ho * Wo;
  const int cls = (mode == 1 && stride == 2 && Ho % 2 == 0 && Wo % 2 == 0 && (M / 4) % CG_BM == 0 && (C1 + C2) % CG_BK == 0) ? 1 : 0;
  if (mode == 0 && kh == 5 && kw == 5 && stride == 2 && pad == 2 && !x2 && N == 64 && !in_scale && !out_scale && !res) {
    const bool done = dtype == 0 ? launch_stem<float>(x, wp, bias, y, B, Hi, Wi, C1, Ho, Wo, ldw, stream)
                                 : launch_stem<__nv_bfloat16>(x, wp, bias, y, B, Hi, Wi, C1, Ho, Wo, ldw, stream);
    if (done) {
      LVAE_COUNT_LAUNCH();
      LVAE_CHECK_LAUNCH("conv2d_gather (stem)");
      return LVAE_OK;
    }
  }
  ConvArgs a{x, x2, wp, bias, in_scale, out_scale, res, y, B, Hi, Wi, C1, C2, Ho, Wo, N, ldw, kh, kw, stride, pad, mode, cls};
  dim3 grid(cdiv(M, CG_BM), cdiv(N, CG_BN));
  bool vec = (C1 % 4 == 0) && (C2 % 4 == 0);
  if (dtype == 0) {
    if (vec) lvae_launch(conv_gather_kernel<float, true>, grid, CG_THREADS, 0, stream, a);
    else lvae_launch(conv_gather_kernel<float, false>, grid, CG_THREADS, 0, stream, a);
  } else {
    if (vec) lvae_launch(conv_gather_kernel<__nv_bfloat16, true>, grid, CG_THREADS, 0, stream, a);
    else lvae_launch(conv_gather_kernel<__nv_bfloat16, false>, grid, CG_THREADS, 0, stream, a);
  }
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("conv2d_gather");
  return LVAE_OK;
}

// ------------------------------------------------------------------------------------------
// wgrad: dW[o][i][ky][kx] += sum_{b,oy,ox} dz[b,oy,ox,o] * u[b, oy*stride-pad+ky, ox*stride-pad+kx, i]
// (u may be the channel concat of u1,u2).  For ConvTranspose2d the caller swaps roles
// (u = dy, dz = x), which yields the (Cin,Cout,kh,kw) layout directly.
// dbias[o] += sum dz[...,o] (only when dbias != null).
// ------------------------------------------------------------------------------------------
struct WgradArgs {
  const void* u; const void* u2; const void* dz;
  const float* in_scale;   // (B, I) on u
  const float* out_scale;  // (B, O) on dz
  float* dw; float* dbias;
  int B, Hi, Wi, C1, C2, Ho, Wo, O, kh, kw, stride, pad;
  int m_per_cta;
};

constexpr int WG_BK = 64, WG_BN = 64, WG_BM = 32, WG_THREADS = 256;

template <typename T, bool VEC>
__global__ void __launch_bounds__(WG_THREADS) conv_wgrad_kernel(WgradArgs a) {
  pdl_wait();
  pdl_launch();
  __shared__ __align__(16) float As[WG_BM][WG_BK + 4];
  __shared__ __align__(16) float Ds[WG_BM][WG_BN + 4];
  const int t = threadIdx.x;
  const int I = a.C1 + a.C2;
  const int K = a.kh * a.kw * I;
  const long long M = (long long)a.B * a.Ho * a.Wo;
  const int k0 = blockIdx.x * WG_BK, n0 = blockIdx.y * WG_BN;
  const long long mbeg = (long long)blockIdx.z * a.m_per_cta;
  const long long mend = min(M, mbeg + (long long)a.m_per_cta);
  const T* u = (const T*)a.u;
  const T* u2 = (const T*)a.u2;
  const T* dz = (const T*)a.dz;
  const int hw = a.Ho * a.Wo;

  // load mapping: pixel rows (t/16 + 16 i), quad (t%16)
  const int q = t & 15;
  // k decode for this thread's A quad is fixed over the whole loop
  int kk = k0 + q * 4;
  int tapv = 0, civ = 0, kyv = 0, kxv = 0;
  if (VEC) {
    tapv = kk / I; civ = kk - tapv * I; kyv = tapv / a.kw; kxv = tapv - kyv * a.kw;
  }
  // scalar path (channel counts that are not a multiple of 4: the 3-channel stem, whose weight gradient closes the backward
  // pass): the (ky, kx, ci) of this thread's four k are fixed over the whole pixel loop too -- decoded once, not per element
  int ky4[4] = {0, 0, 0, 0}, kx4[4] = {0, 0, 0, 0}, ci4[4] = {0, 0, 0, 0};
  if (!VEC) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kk + j;
      if (k < K) {
        const int tap = k / I;
        ci4[j] = k - tap * I;
        ky4[j] = tap / a.kw;
        kx4[j] = tap - ky4[j] * a.kw;
      } else {
        ci4[j] = -1;
      }
    }
  }
  const int tk = t >> 4, tn = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  const bool do_bias = a.dbias != nullptr && blockIdx.x == 0;

  for (long long mc = mbeg; mc < mend; mc += WG_BM) {
    float4 ra[2], rd[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      long long m = mc + (t >> 4) + 16 * i;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vd = va;
      if (m < mend) {
        int b = (int)(m / hw);
        int r = (int)(m - (long long)b * hw);
        int oy = r / a.Wo, ox = r - oy * a.Wo;
        // dz quad
        int n = n0 + q * 4;
        if (n < a.O) {
          if ((a.O & 3) == 0) {
            vd = ld4<T>(dz + m * a.O + n);
          } else {
            float tmp[4] = {0.f, 0.f, 0.f, 0.f};
            for (int j = 0; j < 4 && n + j < a.O; ++j) tmp[j] = ld1<T>(dz + m * a.O + n + j);
            vd = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
          }
          if (a.out_scale) {
            float s[4] = {1.f, 1.f, 1.f, 1.f};
            for (int j = 0; j < 4 && n + j < a.O; ++j) s[j] = a.out_scale[(long long)b * a.O + n + j];
            vd.x *= s[0]; vd.y *= s[1]; vd.z *= s[2]; vd.w *= s[3];
          }
        }
        // gathered u quad
        if (VEC) {
          if (kk < K) {
            int iy = oy * a.stride - a.pad + kyv, ix = ox * a.stride - a.pad + kxv;
            if (iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi) {
              long long pix = ((long long)b * a.Hi + iy) * a.Wi + ix;
              va = civ < a.C1 ? ld4<T>(u + pix * a.C1 + civ) : ld4<T>(u2 + pix * a.C2 + (civ - a.C1));
              if (a.in_scale) {
                float4 s = *reinterpret_cast<const float4*>(a.in_scale + (long long)b * I + civ);
                va.x *= s.x; va.y *= s.y; va.z *= s.z; va.w *= s.w;
              }
            }
          }
        } else {
          float tmp[4] = {0.f, 0.f, 0.f, 0.f};
          const int iy0 = oy * a.stride - a.pad, ix0 = ox * a.stride - a.pad;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ci = ci4[j];
            const int iy = iy0 + ky4[j], ix = ix0 + kx4[j];
            if (ci >= 0 && iy >= 0 && iy < a.Hi && ix >= 0 && ix < a.Wi) {
              const long long pix = ((long long)b * a.Hi + iy) * a.Wi + ix;
              float v = ci < a.C1 ? ld1<T>(u + pix * a.C1 + ci) : ld1<T>(u2 + pix * a.C2 + (ci - a.C1));
              if (a.in_scale) v *= a.in_scale[(long long)b * I + ci];
              tmp[j] = v;
            }
          }
          va = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
        }
      }
      ra[i] = va;
      rd[i] = vd;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      *reinterpret_cast<float4*>(&As[(t >> 4) + 16 * i][q * 4]) = ra[i];
      *reinterpret_cast<float4*>(&Ds[(t >> 4) + 16 * i][q * 4]) = rd[i];
    }
    __syncthreads();
#pragma unroll
    for (int p = 0; p < WG_BM; ++p) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[p][tk * 4]);
      float4 d4 = *reinterpret_cast<const float4*>(&Ds[p][tn * 4]);
      acc[0][0] += a4.x * d4.x; acc[0][1] += a4.x * d4.y; acc[0][2] += a4.x * d4.z; acc[0][3] += a4.x * d4.w;
      acc[1][0] += a4.y * d4.x; acc[1][1] += a4.y * d4.y; acc[1][2] += a4.y * d4.z; acc[1][3] += a4.y * d4.w;
      acc[2][0] += a4.z * d4.x; acc[2][1] += a4.z * d4.y; acc[2][2] += a4.z * d4.z; acc[2][3] += a4.z * d4.w;
      acc[3][0] += a4.w * d4.x; acc[3][1] += a4.w * d4.y; acc[3][2] += a4.w * d4.z; acc[3][3] += a4.w * d4.w;
    }
    if (do_bias && tk == 0) {
#pragma unroll
      for (int p = 0; p < WG_BM; ++p) {
        float4 d4 = *reinterpret_cast<const float4*>(&Ds[p][tn * 4]);
        bsum[0] += d4.x; bsum[1] += d4.y; bsum[2] += d4.z; bsum[3] += d4.w;
      }
    }
  }
  // scatter-accumulate into (O, I, kh, kw)
  const int taps = a.kh * a.kw;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int k = k0 + tk * 4 + i;
    if (k >= K) continue;
    int tap = k / I, ci = k - tap * I;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int o = n0 + tn * 4 + j;
      if (o < a.O) atomicAdd(a.dw + ((long long)o * I + ci) * taps + tap, acc[i][j]);
    }
  }
  if (do_bias && tk == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int o = n0 + tn * 4 + j;
      if (o < a.O) atomicAdd(a.dbias + o, bsum[j]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Weight gradient of the stem (Conv2d c -> 64, 5x5, stride 2, padding 2; c = 1 / 3): the LAST kernel of the backward pass,
// nothing overlaps it, and the scalar path of conv_wgrad_kernel takes ~90-130 us for the CIFAR batch.  Same tiling as the
// forward stem kernel: a CTA walks over 8x16 tiles of output pixels with the 19x35 input patch and the 128x64 dY tile in
// shared memory; thread (ky, ci, o-quad) keeps dW[o..o+3][ci][ky][0..4] (20 accumulators) in registers over ALL its tiles,
// reads a whole patch row once per output row (35 broadcast LDS for 16 pixels) and one LDS.128 of dY per pixel; 16 spare
// threads sum dY for the bias gradient.  One pass of float atomics per CTA at the end (like conv_wgrad_kernel's K-splits).
// ------------------------------------------------------------------------------------------
constexpr int SW_ROWS = 8, SW_PR = 2 * SW_ROWS + 3;

template <typename T, int C>
__global__ void __launch_bounds__(256) wgrad_stem5x5s2_kernel(const T* __restrict__ u, const T* __restrict__ dz, float* __restrict__ dw,
                                                              float* __restrict__ dbias, int Hi, int Wi, int Ho, int Wo, int tiles_x,
                                                              int tiles_per_img, int n_tiles) {
  pdl_wait();
  pdl_launch();
  constexpr int XPITCH = ST_PATCH * C + 1;
  __shared__ float xs[SW_PR][XPITCH];
  __shared__ __align__(16) float ds[SW_ROWS * ST_TILE][64];
  const int t = threadIdx.x;
  const int kyc = t >> 4, oq = t & 15;                  // kyc = ky * C + ci
  const bool active = kyc < 5 * C, is_bias = kyc == 5 * C && dbias != nullptr;
  const int ky = kyc / C, ci = kyc - ky * C;
  float acc[5][4];
#pragma unroll
  for (int kx = 0; kx < 5; ++kx)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[kx][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_img;
    const int tr = tile - b * tiles_per_img;
    const int oy0 = (tr / tiles_x) * SW_ROWS, ox0 = (tr - (tr / tiles_x) * tiles_x) * ST_TILE;
    const int iy0 = 2 * oy0 - 2, ix0 = 2 * ox0 - 2;
    const T* ub = u + (size_t)b * Hi * Wi * C;
    __syncthreads();                                    // the previous tile has been consumed
    {
      // all 8 + 8 staging loads of a thread are in flight before the first shared-memory store (SW_PR * ST_PATCH * C <= 8 * 256,
      // SW_ROWS * ST_TILE * 16 = 8 * 256): one round trip per tile instead of sixteen dependent ones
      static_assert(SW_PR * ST_PATCH * C <= 8 * 256 && SW_ROWS * ST_TILE * 16 == 8 * 256, "staging assumes eight items per thread");
      float xv8[8];
      float4 dv8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = t + u * 256;
        const int r = i / (ST_PATCH * C), e = i - r * (ST_PATCH * C);
        const int iy = iy0 + r, ix = ix0 + e / C;
        xv8[u] = 0.f;
        if (i < SW_PR * ST_PATCH * C && iy >= 0 && iy < Hi && ix >= 0 && ix < Wi)
          xv8[u] = ld1<T>(ub + ((size_t)iy * Wi + ix) * C + (e - (e / C) * C));
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = t + u * 256;
        const int p = i >> 4, c4 = (i & 15) * 4;
        const int oy = oy0 + (p >> 4), ox = ox0 + (p & 15);
        dv8[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (oy < Ho && ox < Wo) dv8[u] = ld4<T>(dz + (((size_t)b * Ho + oy) * Wo + ox) * 64 + c4);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = t + u * 256;
        const int r = i / (ST_PATCH * C), e = i - r * (ST_PATCH * C);
        if (i < SW_PR * ST_PATCH * C) xs[r][e] = xv8[u];
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = t + u * 256;
        *reinterpret_cast<float4*>(&ds[i >> 4][(i & 15) * 4]) = dv8[u];
      }
    }
    __syncthreads();
    if (active) {
#pragma unroll 1
      for (int r = 0; r < SW_ROWS; ++r) {
        const float* xr = &xs[2 * r + ky][ci];
        float xv[ST_PATCH];
#pragma unroll
        for (int i = 0; i < ST_PATCH; ++i) xv[i] = xr[i * C];
#pragma unroll
        for (int ox = 0; ox < ST_TILE; ++ox) {
          const float4 d4 = *reinterpret_cast<const float4*>(&ds[r * ST_TILE + ox][4 * oq]);
#pragma unroll
          for (int kx = 0; kx < 5; ++kx) {
            const float a = xv[2 * ox + kx];
            acc[kx][0] = fmaf(a, d4.x, acc[kx][0]); acc[kx][1] = fmaf(a, d4.y, acc[kx][1]);
            acc[kx][2] = fmaf(a, d4.z, acc[kx][2]); acc[kx][3] = fmaf(a, d4.w, acc[kx][3]);
          }
        }
      }
    } else if (is_bias) {
      for (int p = 0; p < SW_ROWS * ST_TILE; ++p) {
        const float4 d4 = *reinterpret_cast<const float4*>(&ds[p][4 * oq]);
        bsum[0] += d4.x; bsum[1] += d4.y; bsum[2] += d4.z; bsum[3] += d4.w;
      }
    }
  }
  if (active) {
#pragma unroll
    for (int kx = 0; kx < 5; ++kx)
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(dw + ((size_t)(4 * oq + j) * C + ci) * 25 + ky * 5 + kx, acc[kx][j]);
  } else if (is_bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(dbias + 4 * oq + j, bsum[j]);
  }
}

template <typename T>
static bool launch_stem_wgrad(const void* u, const void* dz, float* dw, float* dbias, int B, int Hi, int Wi, int C, int Ho, int Wo,
                              cudaStream_t stream) {
  const int tiles_x = cdiv(Wo, ST_TILE), tiles_per_img = tiles_x * cdiv(Ho, SW_ROWS);
  const long long n_tiles = (long long)B * tiles_per_img;
  if (n_tiles >= (1LL << 31) || (C != 1 && C != 3)) return false;
  const int grid = (int)min(n_tiles, (long long)lvae_num_sms());
  if (C == 1)
    lvae_launch(wgrad_stem5x5s2_kernel<T, 1>, grid, 256, 0, stream, (const T*)u, (const T*)dz, dw, dbias, Hi, Wi, Ho, Wo, tiles_x, tiles_per_img, (int)n_tiles);
  else
    lvae_launch(wgrad_stem5x5s2_kernel<T, 3>, grid, 256, 0, stream, (const T*)u, (const T*)dz, dw, dbias, Hi, Wi, Ho, Wo, tiles_x, tiles_per_img, (int)n_tiles);
  return true;
}

LVAE_API int lvae_conv2d_wgrad(const void* u, const void* u2, const void* dz, const float* in_scale,
                               const float* out_scale, float* dw, float* dbias, int B, int Hi, int Wi,
                               int C1, int C2, int Ho, int Wo, int O, int kh, int kw, int stride, int pad,
                               int dtype, cudaStream_t stream) {
  LVAE_REQUIRE(u && dz && dw, "conv2d_wgrad: null pointer");
  LVAE_REQUIRE((C2 == 0) == (u2 == nullptr), "conv2d_wgrad: u2/C2 mismatch");
  LVAE_REQUIRE(dtype == 0 || dtype == 1, "conv2d_wgrad: dtype must be 0 (f32) or 1 (bf16)");
  if (kh == 5 && kw == 5 && stride == 2 && pad == 2 && !u2 && O == 64 && !in_scale && !out_scale) {
    const bool done = dtype == 0 ? launch_stem_wgrad<float>(u, dz, dw, dbias, B, Hi, Wi, C1, Ho, Wo, stream)
                                 : launch_stem_wgrad<__nv_bfloat16>(u, dz, dw, dbias, B, Hi, Wi, C1, Ho, Wo, stream);
    if (done) {
      LVAE_COUNT_LAUNCH();
      LVAE_CHECK_LAUNCH("conv2d_wgrad (stem)");
      return LVAE_OK;
    }
  }
  int I = C1 + C2, K = kh * kw * I;
  long long M = (long long)B * Ho * Wo;
  int ktiles = cdiv(K, WG_BK), ntiles = cdiv(O, WG_BN);
  // split the pixel reduction so that the grid is ~4 waves, chunks are multiples of WG_BM
  long long target = 4LL * lvae_num_sms();
  long long splits = target / ((long long)ktiles * ntiles);
  if (splits < 1) splits = 1;
  long long max_splits = (M + 4 * WG_BM - 1) / (4 * WG_BM);
  if (splits > max_splits) splits = max_splits;
  int m_per = (int)(((M + splits - 1) / splits + WG_BM - 1) / WG_BM * WG_BM);
  splits = (M + m_per - 1) / m_per;
  WgradArgs a{u, u2, dz, in_scale, out_scale, dw, dbias, B, Hi, Wi, C1, C2, Ho, Wo, O, kh, kw, stride, pad, m_per};
  dim3 grid(ktiles, ntiles, (unsigned)splits);
  bool vec = (C1 % 4 == 0) && (C2 % 4 == 0);
  if (dtype == 0) {
    if (vec) lvae_launch(conv_wgrad_kernel<float, true>, grid, WG_THREADS, 0, stream, a);
    else lvae_launch(conv_wgrad_kernel<float, false>, grid, WG_THREADS, 0, stream, a);
  } else {
    if (vec) lvae_launch(conv_wgrad_kernel<__nv_bfloat16, true>, grid, WG_THREADS, 0, stream, a);
    else lvae_launch(conv_wgrad_kernel<__nv_bfloat16, false>, grid, WG_THREADS, 0, stream, a);
  }
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("conv2d_wgrad");
  return LVAE_OK;
}

// ------------------------------------------------------------------------------------------
// Weight packing: torch (O,I,kh,kw) fp32 -> GEMM-ready [K][ld] rows (fp32 or bf16).
//   mode 0: dst[(tap*I + i)*ld + o] = w[o][i][tap]   (conv forward / ConvTranspose dgrad)
//   mode 1: dst[(tap*O + o)*ld + i] = w[o][i][tap]   (conv dgrad / ConvTranspose forward)
//   mode 2/3: bf16 K-major operand tiles for the tcgen05 kernel (forward / dgrad), see below
// ------------------------------------------------------------------------------------------
struct LvaePackDesc {
  const float* src;
  void* dst;
  int O, I, taps, mode, ld, dtype;
};

__global__ void pack_weights_kernel(const LvaePackDesc* descs, int n) {
  pdl_wait();
  pdl_launch();
  for (int d = blockIdx.y; d < n; d += gridDim.y) {
    LvaePackDesc p = descs[d];
    if (p.mode >= 2) {
      // tcgen05 operand layout: [tap][k-block of 64][Npad rows][64] bf16, K-major rows of 128 bytes
      //   mode 2 (forward):  row n = output channel o, k = input channel i   -> w[o][i][tap]
      //   mode 3 (dgrad):    row n = input channel i,  k = output channel o  -> w[o][i][tap]
      const int nreal = p.mode == 2 ? p.O : p.I, kreal = p.mode == 2 ? p.I : p.O;
      const int KB = (kreal + 63) / 64, Npad = (nreal + 15) / 16 * 16;
      // A thread keeps its column c and walks over rows (tap, k-block, n) with a fixed stride: the row is decoded once and
      // then advanced with carries -- the two runtime divisions per element made this launch instruction-bound (~100 us for
      // the 29 M elements of the CIFAR-15 model's 1118 packs, at the head of every step).
      const int rows = p.taps * KB * Npad, rstep = gridDim.x * (blockDim.x >> 6);
      const int c = threadIdx.x & 63;
      int r = blockIdx.x * (blockDim.x >> 6) + (threadIdx.x >> 6);
      int nn = r % Npad, t2 = r / Npad;
      int kb = t2 % KB, tap = t2 / KB;
      const int dn = rstep % Npad, dt = rstep / Npad;
      __nv_bfloat16* dst = (__nv_bfloat16*)p.dst;
      for (; r < rows; r += rstep) {
        const int kk = kb * 64 + c;
        float v = 0.f;
        if (nn < nreal && kk < kreal) {
          const int o = p.mode == 2 ? nn : kk, i = p.mode == 2 ? kk : nn;
          v = p.src[(o * p.I + i) * p.taps + tap];
        }
        dst[r * 64 + c] = __float2bfloat16(v);
        nn += dn;
        kb += dt;
        if (nn >= Npad) { nn -= Npad; ++kb; }
        while (kb >= KB) { kb -= KB; ++tap; }
      }
      continue;
    }
    int rows = p.taps * (p.mode == 0 ? p.I : p.O);
    int total = rows * p.ld;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
      int r = idx / p.ld, c = idx - r * p.ld;
      int tap, o, i;
      if (p.mode == 0) { tap = r / p.I; i = r - tap * p.I; o = c; }
      else { tap = r / p.O; o = r - tap * p.O; i = c; }
      float v = 0.f;
      if (o < p.O && i < p.I) v = p.src[((long long)o * p.I + i) * p.taps + tap];
      if (p.dtype == 0) ((float*)p.dst)[idx] = v;
      else ((__nv_bfloat16*)p.dst)[idx] = __float2bfloat16(v);
    }
  }
}

// Consumers (conv_tc_kernel and friends) TMA-load packed weights in their prologue, BEFORE griddepcontrol.wait, on the
// invariant of common.cuh that weights were written two or more kernels upstream.  This separator makes the invariant hold
// for whoever launches right after a pack: its dependents can only start once it has seen the pack kernel complete.
__global__ void pack_fence_kernel() {
  pdl_wait();
  pdl_launch();
}

// descs: DEVICE array of n descriptors (built once by the host; weights are re-packed every step)
LVAE_API int lvae_pack_weights(const void* descs_dev, int n, cudaStream_t stream) {
  LVAE_REQUIRE(descs_dev && n > 0, "pack_weights: bad args");
  dim3 grid(8, n < 65535 ? n : 65535);
  lvae_launch(pack_weights_kernel, grid, 256, 0, stream, (const LvaePackDesc*)descs_dev, n);
  LVAE_COUNT_LAUNCH();
  lvae_launch(pack_fence_kernel, 1, 32, 0, stream);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("pack_weights");
  return LVAE_OK;
}

LVAE_API int lvae_pack_desc_size() { return (int)sizeof(LvaePackDesc); }

// per-channel sum over (B,H,W) of dy * scale[b,c]  -> out[c] (+=); used for ConvTranspose dbias
template <typename T>
__global__ void colsum_kernel(const T* dy, const float* scale, float* out, long long M, int C, int hw) {
  pdl_wait();
  pdl_launch();
  // grid.x strides over pixel rows, each thread owns channel (threadIdx.x % C) when blockDim % C == 0
  int c = threadIdx.x % C, rpb = blockDim.x / C, r0 = threadIdx.x / C;
  float s = 0.f;
  for (long long m = (long long)blockIdx.x * rpb + r0; m < M; m += (long long)gridDim.x * rpb) {
    float v = ld1<T>(dy + m * C + c);
    if (scale) v *= scale[(m / hw) * C + c];
    s += v;
  }
  atomicAdd(out + c, s);
}

LVAE_API int lvae_colsum(const void* dy, const float* scale, float* out, int B, int HW, int C, int dtype,
                         cudaStream_t stream) {
  LVAE_REQUIRE(dy && out && C > 0 && C <= 1024, "colsum: bad args");
  int threads = (256 / C) * C;
  if (threads == 0) threads = C;
  long long M = (long long)B * HW;
  int rpb = threads / C;
  int grid = (int)min((long long)2 * lvae_num_sms(), (M + rpb - 1) / rpb);
  if (dtype == 0) lvae_launch(colsum_kernel<float>, grid, threads, 0, stream, (const float*)dy, scale, out, M, C, HW);
  else lvae_launch(colsum_kernel<__nv_bfloat16>, grid, threads, 0, stream, (const __nv_bfloat16*)dy, scale, out, M, C, HW);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("colsum");
  return LVAE_OK;
}

// ------------------------------------------------------------------------------------------
// Narrow-output 3x3 convolution: 64 bf16 channels -> N <= 4 outputs per pixel (the Bernoulli head's parameter_net,
// lib/likelihoods.py:61, 64 -> 1 at 28x28).  A 128x64 GEMM tile would waste 63/64 of its columns there; this is a
// bandwidth problem: 8 lanes own one pixel (16 bytes = 8 channels each, so a pixel row is one coalesced 128-byte read),
// the taps' weights sit in shared memory as fp32, partial dot products meet in a 3-step lane butterfly.
// w: torch layout (N, 64, 3, 3) fp32; y: (B,H,W,N) fp32 or bf16.
// ------------------------------------------------------------------------------------------
template <typename TO>
__global__ void __launch_bounds__(256) conv3x3_narrow_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, TO* __restrict__ y, long long M,
                                                             int H, int W, int N) {
  pdl_wait();
  pdl_launch();
  __shared__ float sw[4][9][64];
  for (int i = threadIdx.x; i < N * 9 * 64; i += blockDim.x) {
    const int n = i / 576, r = i - n * 576, ci = r / 9, t = r - ci * 9;      // torch (n, ci, ky, kx) order
    sw[n][t][ci] = w[i];
  }
  __syncthreads();
  const int sub = threadIdx.x & 7;                                            // which 8 channels of the pixel
  const int hw = H * W;
  for (long long m = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3; m < ((M + 31) & ~31LL);
       m += ((long long)gridDim.x * blockDim.x) >> 3) {
    const bool live = m < M;
    const long long mm = live ? m : 0;
    const int b = (int)(mm / hw);
    const int r = (int)(mm - (long long)b * hw);
    const int py = r / W, px = r - py * W;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int iy = py + t / 3 - 1, ix = px + t % 3 - 1;
      if (live && iy >= 0 && iy < H && ix >= 0 && ix < W) {
        const uint4 u = *reinterpret_cast<const uint4*>(x + (((long long)b * H + iy) * W + ix) * 64 + sub * 8);
        const uint32_t wd[4] = {u.x, u.y, u.z, u.w};
        float v[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) { v[2 * j] = __uint_as_float(wd[j] << 16); v[2 * j + 1] = __uint_as_float(wd[j] & 0xFFFF0000u); }
        for (int n = 0; n < N; ++n) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[n] = fmaf(v[j], sw[n][t][sub * 8 + j], acc[n]);
        }
      }
    }
    for (int n = 0; n < N; ++n) {
      float a = acc[n];
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      if (live && sub == 0) st1<TO>(y + m * N + n, a + (bias ? bias[n] : 0.f));
    }
  }
}

// N = 1 (the Bernoulli head): column-strip formulation.  A group of 8 lanes (8 channels each) owns 8 consecutive output pixels
// of up to NR_ROWS consecutive image rows.  It walks DOWN the strip: every input row (10 pixels = 10 coalesced 128-byte rows,
// loaded unconditionally from clamped coordinates and zeroed afterwards, all ten in flight) feeds the three output rows it
// touches, whose partial sums live in three rotating accumulator sets; the 9 x 8 weights of the lane's channel slice sit in
// registers, and a finished output row leaves through a 7-shuffle transposing reduction (one pixel's sum per lane).  Each input
// pixel is fetched (R + 2) / R x 10 / 8 = 1.6 times (R = 7) instead of 3.75 times with one output row per group, and the taps
// of an output pixel are added in the same order as before (dy, then x, then channel), so results are unchanged to the bit.
// The input may be a WINDOW of a larger NHWC tensor (row / image pitch in elements): the centred crop in front of the
// likelihood (models/lvae.py:143) then costs no pass of its own.
constexpr int NR_ROWS = 8;

template <typename TO>
__global__ void __launch_bounds__(256) conv3x3_narrow1_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, TO* __restrict__ y, int B, int H,
                                                              int W, int row_pitch, long long img_pitch, int chunks, int R) {
  pdl_wait();
  pdl_launch();
  const int sub = threadIdx.x & 7;                          // which 8 channels of a pixel
  float wr[9][8];                                           // torch layout (1, 64, 3, 3): w[ci * 9 + t]
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[t][j] = __ldg(w + (sub * 8 + j) * 9 + t);
  const float b0 = bias ? __ldg(bias) : 0.f;
  const int segs_per_row = (W + 7) >> 3;
  const int n_unit = B * chunks * segs_per_row;             // unit = (image, row chunk, 8-pixel column segment), segment fastest
  const int group = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 3), n_groups = (int)((gridDim.x * blockDim.x) >> 3);
  // all 32 lanes of a warp run the same number of iterations and rows (the shuffles below are warp-wide)
  const int iters = (n_unit + n_groups - 1) / n_groups;
  for (int itn = 0; itn < iters; ++itn) {
    const int unit = group + itn * n_groups;
    const bool live = unit < n_unit;
    const int un = live ? unit : 0;
    const int t1 = un / segs_per_row;
    const int x0 = (un - t1 * segs_per_row) << 3;
    const int b = t1 / chunks;
    const int y0 = (t1 - b * chunks) * R, y1 = min(H, y0 + R);
    const __nv_bfloat16* xb = x + (long long)b * img_pitch + sub * 8;
    float a0[8], a1[8], a2[8];                              // output rows iy + 1, iy, iy - 1 while input row iy is processed
#pragma unroll
    for (int o = 0; o < 8; ++o) a0[o] = a1[o] = a2[o] = 0.f;
    // the next input row is requested before the current one is consumed (rows outside the image or the chunk are fetched from
    // the clamped row and ignored): with one CTA of 8 warps per SM (160+ registers) nothing else hides the load latency
    uint4 u[10], nx[10];
    int xoff[10];
#pragma unroll
    for (int hx = 0; hx < 10; ++hx) xoff[hx] = min(max(x0 + hx - 1, 0), W - 1) * 64;
    {
      const __nv_bfloat16* rp = xb + (long long)min(max(y0 - 1, 0), H - 1) * row_pitch;
#pragma unroll
      for (int hx = 0; hx < 10; ++hx) u[hx] = *reinterpret_cast<const uint4*>(rp + xoff[hx]);
    }
#pragma unroll 1
    for (int k = 0; k < R + 2; ++k) {
      const int iy = y0 - 1 + k;
      const bool rowin = live && iy >= 0 && iy < H && iy <= y1;
      // which of the three output rows this input row feeds lie inside the chunk
      const bool do0 = iy + 1 < y1, do1 = iy >= y0 && iy < y1, do2 = iy - 1 >= y0 && iy - 1 < y1;
      {
        const __nv_bfloat16* rp = xb + (long long)min(max(iy + 1, 0), H - 1) * row_pitch;
#pragma unroll
        for (int hx = 0; hx < 10; ++hx) nx[hx] = *reinterpret_cast<const uint4*>(rp + xoff[hx]);
      }
      if (rowin) {
#pragma unroll
        for (int hx = 0; hx < 10; ++hx) {
          const int ix = x0 + hx - 1;
          const bool ok = ix >= 0 && ix < W;
          const uint32_t wd[4] = {ok ? u[hx].x : 0u, ok ? u[hx].y : 0u, ok ? u[hx].z : 0u, ok ? u[hx].w : 0u};
          float v[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) { v[2 * j] = __uint_as_float(wd[j] << 16); v[2 * j + 1] = __uint_as_float(wd[j] & 0xFFFF0000u); }
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const int o = hx - dx;                           // output pixel x0 + o reads input x0 + o + dx - 1 = x0 + hx - 1
            if (o >= 0 && o < 8) {
              if (do0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) a0[o] = fmaf(v[j], wr[dx][j], a0[o]);
              }
              if (do1) {
#pragma unroll
                for (int j = 0; j < 8; ++j) a1[o] = fmaf(v[j], wr[3 + dx][j], a1[o]);
              }
              if (do2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) a2[o] = fmaf(v[j], wr[6 + dx][j], a2[o]);
              }
            }
          }
        }
      }
      // output row iy - 1 is complete: transposing reduction over the 8 lanes of the group (lane `sub` ends with the sum of
      // output pixel `sub`), then the accumulator sets rotate
      float r[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) r[o] = a2[o];
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const float keep = (sub & 4) ? r[o + 4] : r[o], give = (sub & 4) ? r[o] : r[o + 4];
        r[o] = keep + __shfl_xor_sync(0xffffffffu, give, 4);
      }
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        const float keep = (sub & 2) ? r[o + 2] : r[o], give = (sub & 2) ? r[o] : r[o + 2];
        r[o] = keep + __shfl_xor_sync(0xffffffffu, give, 2);
      }
      {
        const float keep = (sub & 1) ? r[1] : r[0], give = (sub & 1) ? r[0] : r[1];
        r[0] = keep + __shfl_xor_sync(0xffffffffu, give, 1);
      }
      if (live && do2 && x0 + sub < W) st1<TO>(y + ((long long)b * H + (iy - 1)) * W + x0 + sub, r[0] + b0);
#pragma unroll
      for (int o = 0; o < 8; ++o) { a2[o] = a1[o]; a1[o] = a0[o]; a0[o] = 0.f; }
#pragma unroll
      for (int hx = 0; hx < 10; ++hx) u[hx] = nx[hx];
    }
  }
}

// x may be a window of a larger (B, Hs, Ws, 64) tensor: row_pitch / img_pitch = elements between consecutive rows / images of
// the window (0 = dense: W * 64 and H * W * 64); x points at the window's first pixel.
LVAE_API int lvae_conv3x3_narrow_ex(const void* x, const float* w, const float* bias, void* y, int B, int H, int W, int N,
                                    int out_f32, int row_pitch, long long img_pitch, cudaStream_t stream);

LVAE_API int lvae_conv3x3_narrow(const void* x, const float* w, const float* bias, void* y, int B, int H, int W, int N,
                                 int out_f32, cudaStream_t stream) {
  return lvae_conv3x3_narrow_ex(x, w, bias, y, B, H, W, N, out_f32, 0, 0, stream);
}

LVAE_API int lvae_conv3x3_narrow_ex(const void* x, const float* w, const float* bias, void* y, int B, int H, int W, int N,
                                    int out_f32, int row_pitch, long long img_pitch, cudaStream_t stream) {
  LVAE_REQUIRE(x && w && y && B > 0 && H > 0 && W > 0 && N >= 1 && N <= 4, "conv3x3_narrow: bad args (1 <= N <= 4)");
  if (!row_pitch) row_pitch = W * 64;
  if (!img_pitch) img_pitch = (long long)H * W * 64;
  const bool dense = row_pitch == W * 64 && img_pitch == (long long)H * W * 64;
  LVAE_REQUIRE(dense || N == 1, "conv3x3_narrow: a windowed input needs N == 1");
  LVAE_REQUIRE(row_pitch % 8 == 0 && img_pitch % 8 == 0 && ((size_t)x & 15) == 0, "conv3x3_narrow: the window must keep 16-byte alignment");
  if (N == 1 && (long long)B * H * ((W + 7) / 8) < (1LL << 30)) {
    const int chunks = cdiv(H, NR_ROWS), R = cdiv(H, chunks);        // H = 28: four chunks of 7 rows
    const long long groups = (long long)B * chunks * ((W + 7) / 8);
    const int grid = (int)min((long long)8 * lvae_num_sms(), (groups * 8 + 255) / 256);
    if (out_f32) lvae_launch(conv3x3_narrow1_kernel<float>, grid, 256, 0, stream, (const __nv_bfloat16*)x, w, bias, (float*)y, B, H, W, row_pitch, img_pitch, chunks, R);
    else lvae_launch(conv3x3_narrow1_kernel<__nv_bfloat16>, grid, 256, 0, stream, (const __nv_bfloat16*)x, w, bias, (__nv_bfloat16*)y, B, H, W, row_pitch, img_pitch, chunks, R);
    LVAE_COUNT_LAUNCH();
    LVAE_CHECK_LAUNCH("conv3x3_narrow");
    return LVAE_OK;
  }
  const long long M = (long long)B * H * W;
  const int grid = (int)min((long long)8 * lvae_num_sms(), (M * 8 + 255) / 256);
  if (out_f32) lvae_launch(conv3x3_narrow_kernel<float>, grid, 256, 0, stream, (const __nv_bfloat16*)x, w, bias, (float*)y, M, H, W, N);
  else lvae_launch(conv3x3_narrow_kernel<__nv_bfloat16>, grid, 256, 0, stream, (const __nv_bfloat16*)x, w, bias, (__nv_bfloat16*)y, M, H, W, N);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("conv3x3_narrow");
  return LVAE_OK;
}
