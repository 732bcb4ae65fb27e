// Backward of the gate of a gated residual block as ONE kernel (opt-in, LVAE_GATE_BWD_CHAIN=1 on the Python side):
//     dh  = [ dout * sigmoid(g) * act'(a) | dout * act(a) * sigmoid(g) (1 - sigmoid(g)) ]      backward of lib/nn.py:121-126
//     dc2 = (dh . Wg) * mask2                                                                   data gradient of the 1x1 gate conv
//                                                                                               (lib/nn.py:118), with the Dropout2d
//                                                                                               mask of the conv that produced c2
// The 1x1 data gradient is tile-local, so the elementwise gate backward is its operand producer: eight warps read dout and
// h = [a | g] with coalesced 16-byte loads, compute dh in registers, store it (the gate conv's weight gradient needs it) and
// write it, 128B-swizzled, as the K-major A operand (two 64-channel k-blocks) of eight tcgen05.mma (M = 128 pixels, N = 64,
// K = 128); the same warps then drain the TMEM accumulator, apply the mask and TMA-store dc2.  One launch instead of two
// (lvae_gate_bwd + lvae_conv2d_tc) and dh is never re-read.  The operand tile and the accumulator are double buffered: the
// warps produce tile i+1 while the tensor pipe works on tile i.
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int GD_THREADS = 288;            // warp 0: MMA issuer (+ weights TMA), warps 1..8: operand producers / epilogue
constexpr int GD_BM = 128;
constexpr int GD_KBLK_BYTES = GD_BM * 128;         // 16 KB: 128 pixels x 64 channels bf16
constexpr int GD_W_BYTES = 2 * 64 * 128;           // two k-blocks of [64 rows (N)][64 (K)] bf16

struct GdParams {
  const __nv_bfloat16* dout;    // (M,64)
  const __nv_bfloat16* h;       // (M,128) = [a | g]
  __nv_bfloat16* dh;            // (M,128)
  const float* scale;           // (B,64) Dropout2d mask folded into dc2, or null
  int M_total, hw, act;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
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
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) h2[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
  return u;
}

__global__ void __launch_bounds__(GD_THREADS, 1)
gate_dgrad_tc_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmY, const GdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem;                                  // 16 KB: k-block 0 (d a), k-block 1 (d g): [64][64] bf16 each
  uint8_t* sA = sW + GD_W_BYTES;                       // 2 buffers x 2 k-blocks x 16 KB
  uint8_t* sOut = sA + 4 * GD_KBLK_BYTES;              // 16 KB: staged dc2 tile
  uint64_t* bars = (uint64_t*)(sOut + GD_KBLK_BYTES);
  // barriers: 0 weights, 1..2 a_ready[2], 3..4 a_free[2], 5..6 acc_full[2], 7..8 acc_empty[2]
  uint32_t* tmem_slot = (uint32_t*)(bars + 9);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.M_total + GD_BM - 1) / GD_BM;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    mbar_init(BAR(0), 1);
    mbar_init(BAR(1), 1); mbar_init(BAR(2), 1);        // thread 32, after the producers' barrier
    mbar_init(BAR(3), 1); mbar_init(BAR(4), 1);        // tcgen05.commit
    mbar_init(BAR(5), 1); mbar_init(BAR(6), 1);        // tcgen05.commit
    mbar_init(BAR(7), 8); mbar_init(BAR(8), 8);        // one arrive per epilogue warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    if (elect_one()) {                                 // packed many kernels ago: safe ahead of pdl_wait
      mbar_expect_tx(BAR(0), GD_W_BYTES);
      tma_load_2d(smem_u32(sW), &tmW, BAR(0), 0, 0);
      tma_load_2d(smem_u32(sW + 64 * 128), &tmW, BAR(0), 0, 64);
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;               // columns [0,64) / [64,128): the two accumulators
  pdl_wait();
  pdl_launch();

  if (warp == 0) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(GD_BM >> 4) << 24);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    mbar_wait(BAR(0), 0);
    tc_fence_after();
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(BAR(7 + buf), (use & 1) ^ 1);          // epilogue drained this accumulator
      mbar_wait(BAR(1 + buf), use & 1);                // operand tile produced (and fenced for the async proxy)
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t adesc = umma_desc_k_sw128(smem_u32(sA + (buf * 2 + kb) * GD_KBLK_BYTES));
          const uint64_t bdesc = umma_desc_k_sw128(smem_u32(sW + kb * 64 * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_u + (uint32_t)(buf * 64), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
        }
        umma_commit(BAR(3 + buf));                     // operand buffer free again
        umma_commit(BAR(5 + buf));                     // accumulator ready
      }
      __syncwarp();
    }
  } else {
    // ===================== operand producers + epilogue (8 warps) =====================
    const int t = threadIdx.x - 32;                    // 0..255
    const int ew = warp - 1;                           // 0..7
    const int quad = warp & 3;                         // TMEM lane quadrant this warp may read (warp id % 4)
    const int half = ew >> 2;                          // warps 1..4 -> columns [0,32), warps 5..8 -> [32,64)
    const int row = quad * 32 + lane;

    // produce the operand tile of this CTA's j-th tile: item = (row, 16-byte chunk of 8 channels), four items per thread
    auto produce = [&](int j, int tile) {
      const int buf = j & 1;
      const uint32_t use = (uint32_t)(j >> 1);
      const long long m0 = (long long)tile * GD_BM;
      uint4 vd[4], va[4], vg[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int item = t + 256 * i;
        const int r = item >> 3, q = item & 7;
        const long long m = m0 + r;
        if (m < p.M_total) {
          vd[i] = __ldg(reinterpret_cast<const uint4*>(p.dout + m * 64) + q);
          va[i] = __ldg(reinterpret_cast<const uint4*>(p.h + m * 128) + q);
          vg[i] = __ldg(reinterpret_cast<const uint4*>(p.h + m * 128 + 64) + q);
        } else {
          vd[i] = va[i] = vg[i] = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      mbar_wait(BAR(3 + buf), (use & 1) ^ 1);          // the MMAs that read this buffer two tiles ago have retired
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int item = t + 256 * i;
        const int r = item >> 3, q = item & 7;
        const long long m = m0 + r;
        float d[8], a[8], g[8], da[8], dg[8];
        unpack8(vd[i], d); unpack8(va[i], a); unpack8(vg[i], g);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float s = sigmoid_t<true>(g[e]);
          da[e] = d[e] * s * act_bwd_t<true>(a[e], p.act);
          dg[e] = d[e] * act_fwd_t<true>(a[e], p.act) * s * (1.f - s);
        }
        const uint4 ua = pack8(da), ug = pack8(dg);
        const int off = r * 128 + ((q ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(sA + (buf * 2 + 0) * GD_KBLK_BYTES + off) = ua;
        *reinterpret_cast<uint4*>(sA + (buf * 2 + 1) * GD_KBLK_BYTES + off) = ug;
        if (m < p.M_total) {
          *(reinterpret_cast<uint4*>(p.dh + m * 128) + q) = ua;
          *(reinterpret_cast<uint4*>(p.dh + m * 128 + 64) + q) = ug;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (t == 0) mbar_arrive(BAR(1 + buf));
    };

    int it = 0;
    if ((int)blockIdx.x < n_tiles) produce(0, blockIdx.x);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      if (tile + (int)gridDim.x < n_tiles) produce(it + 1, tile + gridDim.x);     // while the tensor pipe works on tile it
      // ---- epilogue of tile it: dc2 = acc * mask -> bf16 -> staged tile -> TMA store ----
      const long long m = (long long)tile * GD_BM + row;
      const int b = m < p.M_total ? (int)(m / p.hw) : 0;
      const float* scale_row = p.scale ? p.scale + (long long)b * 64 : nullptr;
      mbar_wait(BAR(5 + buf), use & 1);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32_nowait(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * 64 + 32 * half), r);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(BAR(7 + buf));
      uint4 packed[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[8 * q + e]);
        if (scale_row) {
          const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale_row + 32 * half + 8 * q));
          const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale_row + 32 * half + 8 * q + 4));
          f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w; f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
        }
        packed[q] = pack8(f);
      }
      if (t == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous tile's store has read sOut
      asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(sOut + row * 128 + (((4 * half + q) ^ (row & 7)) << 4)) = packed[q];
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (t == 0) {
        tma_store_2d(&tmY, smem_u32(sOut), 0, tile * GD_BM);      // rows past M_total are clipped by the tensor map
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_gd() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

}  // namespace

// dout (P,64), h (P,128) = [a | g], dh (P,128), dc2 (P,64): bf16, P = B * hw pixels.  wpb: the gate conv's packed dgrad
// weights, two [64][64] k-blocks (rows = input channel of the forward conv, K = its output channels 0..63 / 64..127;
// lvae_pack_weights mode 3).  scale: (B,64) Dropout2d mask multiplied into dc2, or NULL.
LVAE_API int lvae_gate_bwd_dgrad_tc(const void* dout, const void* h, const void* wpb, const float* scale, void* dh, void* dc2,
                                    int B, int hw, int act, cudaStream_t stream) {
  LVAE_REQUIRE(dout && h && wpb && dh && dc2 && B > 0 && hw > 0, "gate_bwd_dgrad_tc: bad args");
  EncodeTiledFn enc = get_encode_gd();
  if (!enc) { lvae_set_error("gate_bwd_dgrad_tc: cuTensorMapEncodeTiled unavailable"); return LVAE_ERR_CUDA; }
  GdParams p{};
  p.dout = (const __nv_bfloat16*)dout; p.h = (const __nv_bfloat16*)h; p.dh = (__nv_bfloat16*)dh; p.scale = scale;
  p.M_total = B * hw; p.hw = hw; p.act = act;
  CUtensorMap tmW, tmY;
  const cuuint32_t one[2] = {1, 1};
  {
    cuuint64_t wdim[2] = {64, 128};
    cuuint64_t wstr[1] = {128};
    cuuint32_t wbox[2] = {64, 64};
    CUresult r = enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)wpb, wdim, wstr, wbox, one, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t ydim[2] = {64, (cuuint64_t)p.M_total};
    cuuint32_t ybox[2] = {64, (cuuint32_t)GD_BM};
    if (r == CUDA_SUCCESS)
      r = enc(&tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dc2, ydim, wstr, ybox, one, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { lvae_set_error("gate_bwd_dgrad_tc: tensor map encode failed: %d", (int)r); return LVAE_ERR_CUDA; }
  }
  const size_t smem = 1024 + GD_W_BYTES + 5 * GD_KBLK_BYTES + 1024;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gate_dgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { lvae_set_error("gate_bwd_dgrad_tc: cannot raise dynamic smem: %s", cudaGetErrorString(e)); return LVAE_ERR_CUDA; }
    attr = true;
  }
  const int n_tiles = (p.M_total + GD_BM - 1) / GD_BM;
  const int grid = n_tiles < lvae_num_sms() ? n_tiles : lvae_num_sms();
  lvae_launch(gate_dgrad_tc_kernel, grid, GD_THREADS, smem, stream, tmW, tmY, p);
  LVAE_COUNT_LAUNCH();
  LVAE_CHECK_LAUNCH("gate_bwd_dgrad_tc");
  return LVAE_OK;
}
