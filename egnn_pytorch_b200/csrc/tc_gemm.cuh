// tcgen05 GEMM for the per-node contractions of the layer (bf16 operands, fp32 accumulation in TMEM):
//     out[r, n] = act( scale * (sum_k A[r,k] W[n,k] + bias[n]) ) + R[r, n]
// used for  A' = 0.5 (h W1_i^T + b1)  (fp32 out),  B' = 0.5 h W1_j^T  (bf16 out)   [split of egnn_pytorch.py:287]
//           h1 = SiLU([LN(h) | m_i] Wn1^T + bn1),  h' = h1 Wn2^T + bn2 + h                   [egnn_pytorch.py:335-337]
//
// 128x128 output tile per CTA, BK = 64, two shared-memory stages filled with cp.async straight into the
// UMMA no-swizzle K-major core-matrix layout (8 rows x 16 B per core matrix), one elected thread issues
// tcgen05.mma (M=128, N=128, K=16) x4 per stage, completion tracked with tcgen05.commit -> mbarrier,
// epilogue reads the accumulator with tcgen05.ld (thread == row).
#pragma once

#include <cuda_bf16.h>
#include "tc_common.cuh"

namespace egnn {

struct TcGemmArgs {
  const __nv_bfloat16* A; int lda;    // [M][K]
  const __nv_bfloat16* W; int ldw;    // [Nv][K]
  const float* bias;                  // [Nv] | null
  const __nv_bfloat16* R; int ldr;    // residual [M][Nv] | null
  void* out; int ldo; int out_f32;
  int M, Nv, Nout, K;
  float scale; int act;               // act: 0 none, 1 SiLU
};

constexpr int GEMM_BM = 128, GEMM_BN = 128, GEMM_BK = 64;
constexpr int GEMM_TILE_BYTES = GEMM_BM * GEMM_BK * 2;            // 16 KB per operand per stage
constexpr int GEMM_SMEM_BYTES = 2 * 2 * GEMM_TILE_BYTES + 1024;   // 2 stages x (A, W) + alignment slack

__global__ void __launch_bounds__(128, 1) tc_gemm_kernel(const TcGemmArgs g) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t mma_done[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * GEMM_BM, n0 = blockIdx.x * GEMM_BN;
  const uint32_t s_base = tc::smem_u32(smem);

  if (tid == 0) { tc::mbar_init(&mma_done[0], 1); tc::mbar_init(&mma_done[1], 1); tc::mbar_fence_init(); }
  if (warp == 0) tc::tmem_alloc<128>(&tmem_base_s);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  const int nk = (g.K + GEMM_BK - 1) / GEMM_BK;
  auto load_stage = [&](int kt) {
    const int st = kt & 1;
    const uint32_t sA = s_base + st * 2 * GEMM_TILE_BYTES, sW = sA + GEMM_TILE_BYTES;
    const int rowA = m0 + tid, rowW = n0 + tid;
#pragma unroll
    for (int kc = 0; kc < 8; ++kc) {
      const int k = kt * GEMM_BK + kc * 8;
      const bool kin = k < g.K;
      const bool va = kin && rowA < g.M, vw = kin && rowW < g.Nv;
      const __nv_bfloat16* pa = va ? g.A + (size_t)rowA * g.lda + k : g.A;
      const __nv_bfloat16* pw = vw ? g.W + (size_t)rowW * g.ldw + k : g.W;
      tc::cp_async16(sA + kc * 2048 + tid * 16, pa, va ? 16u : 0u);
      tc::cp_async16(sW + kc * 2048 + tid * 16, pw, vw ? 16u : 0u);
    }
    tc::cp_async_commit();
  };

  constexpr uint32_t IDESC = tc::idesc_bf16_f32(128, 128);
  load_stage(0);
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) {
      // stage (kt+1)&1 was last read by the MMAs of iteration kt-1
      if (kt >= 1) tc::mbar_wait(&mma_done[(kt + 1) & 1], ((kt - 1) >> 1) & 1);
      load_stage(kt + 1);
      tc::cp_async_wait<1>();
    } else {
      tc::cp_async_wait<0>();
    }
    tc::fence_proxy_async_smem();          // cp.async (generic proxy) -> tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      const uint32_t sA = s_base + (kt & 1) * 2 * GEMM_TILE_BYTES, sW = sA + GEMM_TILE_BYTES;
#pragma unroll
      for (int kk = 0; kk < GEMM_BK / 16; ++kk) {
        constexpr uint32_t lbo = 2048u, sbo = 128u;   // K-adjacent / M-adjacent core matrices of the tile
        const uint64_t da = tc::smem_desc_kmajor_noswizzle(sA + kk * 4096, lbo, sbo);
        const uint64_t dw = tc::smem_desc_kmajor_noswizzle(sW + kk * 4096, lbo, sbo);
        tc::mma_ss(tmem, da, dw, IDESC, (kt > 0 || kk > 0) ? 1u : 0u);
      }
      tc::mma_commit(&mma_done[kt & 1]);
    }
  }
  // the last commit covers every MMA issued before it
  tc::mbar_wait(&mma_done[(nk - 1) & 1], ((nk - 1) >> 1) & 1);
  tc::tc_fence_after();

  const int row = m0 + warp * 32 + lane;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
  for (int cb = 0; cb < GEMM_BN / 32; ++cb) {
    uint32_t r[32];
    tc::tmem_ld32(trow + cb * 32, r);
    tc::tmem_wait_ld();
    const int col0 = n0 + cb * 32;
    if (row < g.M && col0 < g.Nout) {
      float v[32];
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const int col = col0 + q;
        float x = 0.f;
        if (col < g.Nv) {
          x = __uint_as_float(r[q]);
          if (g.bias) x += g.bias[col];
          x *= g.scale;
          if (g.act == 1) x = x / (1.0f + __expf(-x));
          if (g.R) x += __bfloat162float(g.R[(size_t)row * g.ldr + col]);
        }
        v[q] = x;
      }
      if (g.out_f32) {
        float* o = static_cast<float*>(g.out) + (size_t)row * g.ldo + col0;
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          if (col0 + q + 4 <= g.Nout) *reinterpret_cast<float4*>(o + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
          else for (int z = 0; z < 4; ++z) if (col0 + q + z < g.Nout) o[q + z] = v[q + z];
        }
      } else {
        __nv_bfloat16* o = static_cast<__nv_bfloat16*>(g.out) + (size_t)row * g.ldo + col0;
#pragma unroll
        for (int q = 0; q < 32; q += 8) {
          if (col0 + q + 8 <= g.Nout) {
            uint4 pk;
            pk.x = tc::pack_bf16x2(v[q], v[q + 1]); pk.y = tc::pack_bf16x2(v[q + 2], v[q + 3]);
            pk.z = tc::pack_bf16x2(v[q + 4], v[q + 5]); pk.w = tc::pack_bf16x2(v[q + 6], v[q + 7]);
            *reinterpret_cast<uint4*>(o + q) = pk;
          } else {
            for (int z = 0; z < 8; ++z) if (col0 + q + z < g.Nout) o[q + z] = __float2bfloat16(v[q + z]);
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<128>(tmem);
}

}  // namespace egnn
