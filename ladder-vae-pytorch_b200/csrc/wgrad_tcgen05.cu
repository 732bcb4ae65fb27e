// bf16 weight-gradient GEMM on the 5th-generation tensor cores (sm_100a):
//     dW[tap][ci][co] = sum_pixels X[pixel + offset(tap)][ci] * dY[pixel][co]
// for the stride-1 "same" convolutions (3x3 64->64, 1x1 64->128 gate, 1x1 merge over two inputs).
//
// The reduction dimension is the pixel index, so both operands are "MN-major": a TMA tile of 128
// pixels x 64 channels (128-byte swizzled rows) is, read column-wise, a 64 x 128 operand whose K
// runs over the rows.  One tcgen05.mma (M = 128, N = co, K = 16 pixels) consumes TWO activation
// tiles stacked along M (two filter taps, or a tap and an all-ones tile whose accumulator rows are
// the bias gradient), so a 3x3 convolution needs five accumulators of 64 TMEM columns each.
// Every CTA reduces its share of the pixel tiles into TMEM, then writes ONE fp32 partial to a
// workspace; lvae_wgrad_reduce sums the partials and scatters into the (O, I, kh, kw) gradient.
// Shifted 4-D TMA boxes implement im2col, out-of-bounds zero fill implements the padding.
#include "common.cuh"
#include <cuda.h>

namespace {

constexpr int WG_THREADS_TC = 192;
constexpr int WG_TILE = 128;                 // pixels per tile (K of the GEMM)
constexpr int WG_BLK_BYTES = WG_TILE * 64 * 2;   // 16 KB
constexpr int WG_MAX_PAIRS = 5;
constexpr int WG_PAIR_SLOTS = 4;
constexpr int WG_B_SLOTS = 2;

struct WgParams {
  int M_total, H, W;
  int n_pairs;               // accumulators
  int n_bblk;                // dY channel blocks of 64 (1 or 2)  -> N = 64 * n_bblk
  int tmem_cols;
  int tiles_per_cta;
  int dy_c0;                 // first dY channel of this launch (dY may be wider than N: split launches)
  float* ws;                 // (grid, n_pairs, 128, N) fp32 partials
  // A-block table: 2 per pair.  src: 0 = x, 1 = x2, 2 = ones tile, 3 = unused (zero rows, never read back)
  int8_t a_src[2 * WG_MAX_PAIRS], a_dx[2 * WG_MAX_PAIRS], a_dy[2 * WG_MAX_PAIRS];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// MN-major, 128B-swizzled operand: 64 MN elements per 128-byte row, 8 K-rows per 1024-byte atom (SBO),
// next block of 64 MN elements LBO bytes further.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(WG_THREADS_TC, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmX2,
                const __grid_constant__ CUtensorMap tmDY, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sPair = smem;                                                   // WG_PAIR_SLOTS x 32 KB
  uint8_t* sB = sPair + WG_PAIR_SLOTS * 2 * WG_BLK_BYTES;                  // WG_B_SLOTS x n_bblk x 16 KB
  uint8_t* sOnes = sB + WG_B_SLOTS * p.n_bblk * WG_BLK_BYTES;              // 16 KB of bf16 1.0
  uint64_t* bars = (uint64_t*)(sOnes + WG_BLK_BYTES);
  // barriers: [0,4) pair full, [4,8) pair empty, [8,10) b full, [10,12) b empty, 12: accumulators done
  uint32_t* tmem_slot = (uint32_t*)(bars + 16);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.M_total + WG_TILE - 1) / WG_TILE;
  const int tile_beg = blockIdx.x * p.tiles_per_cta;
  const int tile_end = min(n_tiles, tile_beg + p.tiles_per_cta);
  const int N = 64 * p.n_bblk;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 13; ++i) mbar_init(BAR(i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // all-ones tile (any swizzle of a constant tile is the same tile)
  for (int i = threadIdx.x; i < WG_BLK_BYTES / 4; i += WG_THREADS_TC) ((uint32_t*)sOnes)[i] = 0x3F803F80u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core (async proxy)
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch();

  if (warp == 0) {
    {
      // ===================== TMA producer (whole warp loops, one elected lane issues) =====================
      const int hw = p.H * p.W;
      int ps = 0, bs = 0;
      uint32_t pph = 0, bph = 0;
      for (int tile = tile_beg; tile < tile_end; ++tile) {
        int p0 = tile * WG_TILE;
        int n0 = p0 / hw;
        int rem = p0 - n0 * hw;
        int h0 = rem / p.W;
        int w0 = rem - h0 * p.W;
        mbar_wait(BAR(10 + bs), bph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(BAR(8 + bs), (uint32_t)(p.n_bblk * WG_BLK_BYTES));
          for (int j = 0; j < p.n_bblk; ++j)
            tma_load_4d(smem_u32(sB + (bs * p.n_bblk + j) * WG_BLK_BYTES), &tmDY, BAR(8 + bs), p.dy_c0 + 64 * j, w0, h0, n0);
        }
        __syncwarp();
        if (++bs == WG_B_SLOTS) { bs = 0; bph ^= 1; }
        for (int pr = 0; pr < p.n_pairs; ++pr) {
          mbar_wait(BAR(4 + ps), pph ^ 1);
          if (elect_one()) {
            int nload = 0;
            for (int h = 0; h < 2; ++h) nload += p.a_src[2 * pr + h] < 2 ? 1 : 0;
            mbar_expect_tx(BAR(ps), (uint32_t)(nload * WG_BLK_BYTES));
            for (int h = 0; h < 2; ++h) {
              int src = p.a_src[2 * pr + h];
              if (src < 2)
                tma_load_4d(smem_u32(sPair + (ps * 2 + h) * WG_BLK_BYTES), src ? &tmX2 : &tmX, BAR(ps), 0,
                            w0 + p.a_dx[2 * pr + h], h0 + p.a_dy[2 * pr + h], n0);
            }
          }
          __syncwarp();
          if (++ps == WG_PAIR_SLOTS) { ps = 0; pph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {
      // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      // D fp32, A/B bf16, both MN-major, M = 128, N = 64 * n_bblk
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int ps = 0, bs = 0;
      uint32_t pph = 0, bph = 0;
      for (int tile = tile_beg; tile < tile_end; ++tile) {
        mbar_wait(BAR(8 + bs), bph);
        tc_fence_after();
        const uint32_t b_addr = smem_u32(sB + bs * p.n_bblk * WG_BLK_BYTES);
        for (int pr = 0; pr < p.n_pairs; ++pr) {
          mbar_wait(BAR(ps), pph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sPair + ps * 2 * WG_BLK_BYTES);
          // second 64-row block: the slot's own second half, or the constant ones tile
          const uint32_t a2 = p.a_src[2 * pr + 1] == 2 ? smem_u32(sOnes) : a_addr + WG_BLK_BYTES;
          const uint32_t lbo_a = a2 - a_addr;
          const uint32_t d_tmem = tmem_u + (uint32_t)(pr * N);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < WG_TILE / 16; ++k) {
              // 16 pixels = two 8-row swizzle atoms = 2048 bytes along K; the ones tile is constant, so the
              // same LBO works for it at every k (a_k + LBO stays inside its 16 KB)
              umma_bf16(d_tmem, umma_desc_mn_sw128(a_addr + k * 2048, lbo_a), umma_desc_mn_sw128(b_addr + k * 2048, WG_BLK_BYTES),
                        idesc, (uint32_t)((tile != tile_beg) || k != 0));
            }
            umma_commit(BAR(4 + ps));
          }
          __syncwarp();
          if (++ps == WG_PAIR_SLOTS) { ps = 0; pph ^= 1; }
        }
        if (elect_one()) umma_commit(BAR(10 + bs));
        __syncwarp();
        if (++bs == WG_B_SLOTS) { bs = 0; bph ^= 1; }
      }
      if (elect_one()) umma_commit(BAR(12));
      __syncwarp();
    }
  } else {
    // ===================== epilogue: TMEM -> fp32 partial in the workspace =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    mbar_wait(BAR(12), 0);
    tc_fence_after();
    float* wsc = p.ws + (size_t)blockIdx.x * p.n_pairs * 128 * N;
    for (int pr = 0; pr < p.n_pairs; ++pr) {
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(pr * N);
      float* dst = wsc + ((size_t)pr * 128 + row) * N;
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
        if (tile_beg >= tile_end) {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;          // CTA without work: contribute zeros
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          reinterpret_cast<float4*>(dst + c0)[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                               __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// sum the per-CTA partials and accumulate into the torch-layout gradient: block blk (= pair*2 + half) of kind
//   tap t, input-channel offset ci0:   dw[(co*I + ci0 + ci)*taps + t] += sum      (ci = row % 64)
//   ones:                              dbias[co] += sum (row 0 of the block only)
struct RedParams {
  const float* ws; float* dw; float* dbias;
  int n_cta, n_pairs, N, I, taps, I_real, N_real;
  int8_t kind[2 * WG_MAX_PAIRS];     // 0 tap block, 1 ones, 2 unused
  int8_t tap[2 * WG_MAX_PAIRS], ci0_blk[2 * WG_MAX_PAIRS];
};

__global__ void wgrad_reduce_kernel(RedParams p) {
  pdl_wait();
  pdl_launch();
  // grid.y splits the partials: each thread sums its slice of CTAs (independent loads, 4 in flight), then adds
  const int per = p.n_pairs * 128 * p.N;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per) return;
  const int co = idx % p.N;
  const int row = (idx / p.N) % 128;
  const int pr = idx / (p.N * 128);
  const int blk = pr * 2 + (row >> 6);
  const int kind = p.kind[blk];
  if (kind == 2 || (kind == 1 && ((row & 63) != 0 || !p.dbias))) return;
  if (kind == 0 && p.ci0_blk[blk] * 64 + (row & 63) >= p.I_real) return;      // zero-padded input channels
  if (co >= p.N_real) return;                                                   // zero-padded output channels
  const int chunk = (p.n_cta + gridDim.y - 1) / gridDim.y;
  const int c_beg = blockIdx.y * chunk, c_end = min(p.n_cta, c_beg + chunk);
  if (c_beg >= c_end) return;
  const float* src = p.ws + idx;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int c = c_beg;
  for (; c + 4 <= c_end; c += 4) {
    s0 += __ldg(src + (size_t)c * per);
    s1 += __ldg(src + (size_t)(c + 1) * per);
    s2 += __ldg(src + (size_t)(c + 2) * per);
    s3 += __ldg(src + (size_t)(c + 3) * per);
  }
  for (; c < c_end; ++c) s0 += __ldg(src + (size_t)c * per);
  const float s = (s0 + s1) + (s2 + s3);
  float* dst = kind == 1 ? p.dbias + co
                         : p.dw + ((size_t)co * p.I_real + p.ci0_blk[blk] * 64 + (row & 63)) * p.taps + p.tap[blk];
  if (gridDim.y == 1) *dst += s;
  else atomicAdd(dst, s);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn wg_get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

int make_act_map(EncodeTiledFn enc, CUtensorMap* tm, const void* ptr, int B, int H, int W, int C, int bw, int bh, int bn) {
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return (int)enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)ptr, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace

// Workspace floats needed by lvae_conv2d_wgrad_tc for these shapes (0 if unsupported).
LVAE_API long long lvae_wgrad_tc_workspace(int B, int H, int W, int N, int ksize, int two_inputs) {
  if (!(N == 64 || N == 128) || !(ksize == 1 || ksize == 3)) return 0;
  int ablk = ksize * ksize * (two_inputs ? 2 : 1) + 1;
  int n_pairs = (ablk + 1) / 2;
  if (n_pairs * N > 512) return 0;
  return (long long)lvae_num_sms() * n_pairs * 128 * N;
}

// x, x2: (B,H,W,64) bf16 (x2 optional); dy: (B,H,W,N) bf16, N in {64,128}, already multiplied by any Dropout2d mask.
// dw: (N, I_real, k, k) fp32 (+=); I_real <= 64 * inputs (x may carry zero-padded channels beyond I_real, 0 = no padding);
// dy may carry zero-padded channels beyond N_real (0 = none): dw is then (N_real, I_real, k, k) and dbias [N_real].
// dY is (B,H,W,dyC) (dyC = 0 means N); this launch uses its channels [dy_c0, dy_c0 + N).
// dbias: fp32 (+=) or NULL.  ws: workspace (see above).
LVAE_API int lvae_conv2d_wgrad_tc(const void* x, const void* x2, const void* dy, float* dw, float* dbias, float* ws,
                                  int B, int H, int W, int N, int ksize, int I_real, int N_real, int dyC, int dy_c0,
                                  cudaStream_t stream) {
  LVAE_REQUIRE(x && dy && dw && ws, "conv2d_wgrad_tc: null pointer");
  LVAE_REQUIRE((N == 64 || N == 128) && (ksize == 1 || ksize == 3), "conv2d_wgrad_tc: N must be 64 or 128, ksize 1 or 3");
  LVAE_REQUIRE((W & (W - 1)) == 0 && (H & (H - 1)) == 0 && W <= 128, "conv2d_wgrad_tc: H and W must be powers of two (W <= 128)");
  EncodeTiledFn enc = wg_get_encode();
  if (!enc) { lvae_set_error("conv2d_wgrad_tc: cuTensorMapEncodeTiled unavailable"); return LVAE_ERR_CUDA; }
  const int inputs = x2 ? 2 : 1, taps = ksize * ksize;
  LVAE_REQUIRE(taps * inputs + 1 + ((taps * inputs) % 2 == 0 ? 1 : 0) <= 2 * WG_MAX_PAIRS,
               "conv2d_wgrad_tc: too many operand blocks (3x3 over two inputs is not supported)");
  WgParams p{};
  RedParams rp{};
  int nblk = 0;
  for (int t = 0; t < taps; ++t) {
    int oy = t / ksize - ksize / 2, ox = t % ksize - ksize / 2;
    for (int s = 0; s < inputs; ++s, ++nblk) {
      p.a_src[nblk] = (int8_t)s; p.a_dx[nblk] = (int8_t)ox; p.a_dy[nblk] = (int8_t)oy;
      rp.kind[nblk] = 0; rp.tap[nblk] = (int8_t)t; rp.ci0_blk[nblk] = (int8_t)s;
    }
  }
  // the ones block must be the SECOND half of a pair (its half of the descriptor is not advanced along K)
  if (nblk % 2 == 0) { p.a_src[nblk] = 0; p.a_dx[nblk] = 0; p.a_dy[nblk] = 0; rp.kind[nblk] = 2; ++nblk; }
  p.a_src[nblk] = 2; rp.kind[nblk] = 1; ++nblk;
  p.n_pairs = nblk / 2;
  LVAE_REQUIRE(p.n_pairs <= WG_MAX_PAIRS, "conv2d_wgrad_tc: too many operand blocks");
  p.n_bblk = N / 64;
  LVAE_REQUIRE(p.n_pairs * N <= 512, "conv2d_wgrad_tc: accumulators do not fit TMEM");
  p.tmem_cols = 32;
  while (p.tmem_cols < p.n_pairs * N) p.tmem_cols *= 2;
  p.M_total = B * H * W; p.H = H; p.W = W;
  const int n_tiles = (p.M_total + WG_TILE - 1) / WG_TILE;
  int grid = n_tiles < lvae_num_sms() ? n_tiles : lvae_num_sms();
  // fewer, fatter CTAs for small problems: every CTA costs one partial in the reduction
  if (n_tiles <= lvae_num_sms() && n_tiles >= 8) grid = (n_tiles + 1) / 2;
  p.tiles_per_cta = (n_tiles + grid - 1) / grid;
  grid = (n_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  p.ws = ws;
  const int bw = W;
  int bh = 1;
  while (bh * 2 <= H && bh * 2 * bw <= WG_TILE) bh *= 2;
  const int bn = WG_TILE / (bw * bh);
  CUtensorMap tmX, tmX2, tmDY;
  int r = make_act_map(enc, &tmX, x, B, H, W, 64, bw, bh, bn);
  if (!r) r = make_act_map(enc, &tmX2, x2 ? x2 : x, B, H, W, 64, bw, bh, bn);
  if (dyC <= 0) dyC = N;
  LVAE_REQUIRE(dy_c0 % 64 == 0 && dy_c0 + N <= dyC, "conv2d_wgrad_tc: bad dY channel window");
  p.dy_c0 = dy_c0;
  if (!r) r = make_act_map(enc, &tmDY, dy, B, H, W, dyC, bw, bh, bn);
  if (r) { lvae_set_error("conv2d_wgrad_tc: tensor map encode failed: %d", r); return LVAE_ERR_CUDA; }
  const size_t smem = 1024 + (size_t)WG_PAIR_SLOTS * 2 * WG_BLK_BYTES + (size_t)WG_B_SLOTS * p.n_bblk * WG_BLK_BYTES + WG_BLK_BYTES + 256;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { lvae_set_error("conv2d_wgrad_tc: cannot raise dynamic smem: %s", cudaGetErrorString(e)); return LVAE_ERR_CUDA; }
    attr = true;
  }
  lvae_launch(wgrad_tc_kernel, grid, WG_THREADS_TC, smem, stream, tmX, tmX2, tmDY, p);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("conv2d_wgrad_tc");
  rp.ws = ws; rp.dw = dw; rp.dbias = dbias; rp.n_cta = grid; rp.n_pairs = p.n_pairs; rp.N = N; rp.I = 64 * inputs; rp.taps = taps; rp.I_real = I_real > 0 ? I_real : 64 * inputs; rp.N_real = N_real > 0 ? N_real : N;
  const int per = p.n_pairs * 128 * N;
  const int ysplit = grid >= 64 ? 8 : (grid >= 16 ? 4 : 1);
  lvae_launch(wgrad_reduce_kernel, dim3((per + 255) / 256, ysplit), 256, 0, stream, rp);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("wgrad_reduce");
  return LVAE_OK;
}
