// The fused edge step on 5th-gen tensor cores (dense all-pairs, bf16 operands, fp32 accumulation).
//
// Reference semantics: egnn_pytorch.py:232-233 (x_i - x_j, squared distance), :282-287 (edge MLP on
// [h_i | h_j | d]), :289-290 (gate), :292-333 (masks, coors MLP, clamp, both sums over j).
//
// Per pair (i, j) the split form needs  hidden[c] = SiLU(A_i[c] + B_j[c] + wd[c] d_ij), c < H, and
// m_pre = hidden . W2^T  (H -> 16).  One CTA owns TI = 4 query rows i and walks all j in blocks of 512:
//   * 4 compute warpgroups (128 threads each) -- 4 warps per SM sub-partition, which is what it takes
//     to keep the MUFU pipe fed (round-1 profile: 2 warps/SMSP left it 54 % idle on fixed-latency
//     stalls).  A warpgroup owns 128 neighbours j.  For each hidden chunk of 64 channels and each row i
//     it produces the 64 bf16 hidden values of its 128 pairs in registers (fp32 math, one MUFU.TANH per
//     value) and stores them with tcgen05.st into a TMEM slot laid out as the MMA A operand (lane = pair,
//     32 columns = 64 bf16): the O(N^2 H) hidden tensor only ever exists 8 KB at a time, in TMEM;
//   * the MMAs are issued by the compute warps themselves (a 17th warp would cut the register budget from
//     128 to 96): after storing round n, warp n%4 of the warpgroup waits for the other three to arrive on
//     the slot's `full` mbarrier and one lane issues tcgen05.mma.kind::f16 (M=128 pairs, N=16, K=16) x4
//     with A from TMEM and B = the W2 slab from shared memory, accumulating m_pre[i] (128 x 16 fp32) in
//     TMEM across all chunks; tcgen05.commit releases the slot through its `empty` mbarrier;
//   * after the last chunk each thread reads its own pair's 16 accumulators back (tcgen05.ld), applies
//     SiLU / gate / coors MLP / mask / clamp in fp32 and the warp reduces over j with shuffles; per-row
//     sums  sum_j m_ij  and  sum_j w_ij (x_i - x_j)  live in shared memory and are written once per row.
// W2 (packed in UMMA core-matrix order), the A_i rows and wd are staged once per CTA with TMA bulk copies
// (cp.async.bulk -> UBLKCP) onto an mbarrier.
//
// Thread <-> data mappings inside a compute warp:
//   "pair" mapping     (geometry, epilogue, tcgen05.ld 32x32b): lane l owns pair row 32*wq + l of the tile;
//   "fragment" mapping (hidden production, tcgen05.st 16x256b): lane (lr = l/4, lq = l%4) owns rows
//     lr + 8*rho (rho = 0..3) of the warp's 32-row quadrant and, in every 16-channel K-slab, channels
//     4*lq .. 4*lq+3.  Lanes sharing lq read the same A'/wd words (4 distinct addresses per warp instead
//     of a 32-way broadcast, which cost one shared-memory wavefront per 4 bytes in the first version).
#pragma once

#include <cuda_bf16.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace egnn {

constexpr int TP_TI = 4;          // query rows per CTA
constexpr int TP_KC = 64;         // hidden channels per chunk (= 32 TMEM columns, 4 MMAs)
constexpr int TP_SLOTS = 2;       // A-operand slots per warpgroup
constexpr int TP_WG = 4;          // compute warpgroups
constexpr int TP_CWARPS = TP_WG * 4;
constexpr int TP_THREADS = TP_WG * 128;      // 16 warps = 4 per SM sub-partition, 128 registers each
constexpr int TP_JB = TP_WG * 128;    // neighbours per block
constexpr int TP_WGCOLS = 128;        // TMEM columns per warpgroup: TI*16 accumulators + SLOTS*32 operand
constexpr int TP_EPI_FLOATS = 64 * 16 + 64 + 64 + 16 + 16 + 4;   // W3 | b3 | w4 | b2 | gate_w | gate_b, b4, scale, 0
static_assert(TP_TI * 16 + TP_SLOTS * 32 == TP_WGCOLS && TP_WG * TP_WGCOLS == 512, "TMEM budget");

struct TcPairArgs {
  int B, N, Hp, ldn, dim;          // ldn: row stride of node_in (bf16 elements)
  uint32_t flags; int has_mask; float clamp;
  const float* Atab;               // [M][Hp]  0.5 (h W1_i^T + b1)
  const __nv_bfloat16* Btab;       // [M][Hp]  0.5 h W1_j^T
  const float* wdh;                // [Hp]     0.5 W1[:, 2dim]
  const __nv_bfloat16* w2p;        // [Hp/16][2][2][8][8]  W2 in core-matrix order
  const float* epi;                // TP_EPI_FLOATS
  const float* coors;              // [B][N][3]
  const uint8_t* mask;             // [B][N] | null
  __nv_bfloat16* m_out;            // node_in + dim (stride ldn) | null
  float* coors_out;                // [B][N][3] | null
};

inline size_t tc_pair_smem_bytes(int Hp) {
  size_t n = 0;
  n += (size_t)Hp * 32;                       // W2 slabs
  n += (size_t)TP_TI * Hp * 4;                // A rows (fp32)
  n += (size_t)Hp * 4;                        // wd
  n += (size_t)TP_EPI_FLOATS * 4;             // epilogue constants
  n += (size_t)TP_CWARPS * TP_TI * 20 * 4;    // per-warp partial sums
  n += (size_t)TP_TI * 4 * 4 + 64;            // x_i, mask_i, tmem pointer
  n += (size_t)TP_TI * TP_JB * 4;             // d_ij of the current j-block
  n += 32 * 8;                                // mbarriers
  return n + 128;
}

__global__ void __launch_bounds__(TP_THREADS, 1) tc_pair_kernel(const TcPairArgs a) {
  // carve the dynamic shared memory directly (no integer round trip) so every access stays in the
  // shared state space (LDS/STS, not generic LD/ST); nothing here needs more than 128-byte alignment
  extern __shared__ __align__(128) unsigned char sm[];
  const int Hp = a.Hp, N = a.N;
  unsigned char* w2s = sm;                                                    // Hp*32 bytes
  float* As = reinterpret_cast<float*>(w2s + (size_t)Hp * 32);                // [TI][Hp]
  float* wds = As + (size_t)TP_TI * Hp;                                       // [Hp]
  float* epi = wds + Hp;                                                      // constants
  float* part = epi + TP_EPI_FLOATS;                                          // [16 warps][TI][20]
  float* xis = part + TP_CWARPS * TP_TI * 20;                                 // [TI][4]
  uint32_t* mki = reinterpret_cast<uint32_t*>(xis + TP_TI * 4);               // [TI] (+ tmem ptr at [15])
  float* dsm = reinterpret_cast<float*>(mki + 16);                            // [TI][TP_JB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(dsm + TP_TI * TP_JB);
  uint64_t* full = bars;                          // [WG][SLOTS]
  uint64_t* empty = bars + TP_WG * TP_SLOTS;      // [WG][SLOTS]
  uint64_t* accdone = empty + TP_WG * TP_SLOTS;   // [WG]
  uint64_t* ldbar = accdone + TP_WG;              // staging barrier

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.y, i0 = blockIdx.x * TP_TI;
  const int rows_valid = min(TP_TI, N - i0);
  const int nchunks = Hp / TP_KC;
  const int njb = (N + TP_JB - 1) / TP_JB;
  const bool upd_feats = a.flags & EGNN_FLAG_UPDATE_FEATS, upd_coors = a.flags & EGNN_FLAG_UPDATE_COORS;

  // ---------------- setup
  if (tid == 0) {
    for (int x = 0; x < TP_WG * TP_SLOTS; ++x) { tc::mbar_init(&full[x], 128); tc::mbar_init(&empty[x], 1); }
    for (int x = 0; x < TP_WG; ++x) tc::mbar_init(&accdone[x], 1);
    tc::mbar_init(ldbar, 1);
    tc::mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc<512>(&mki[15]);
  for (int x = tid; x < TP_EPI_FLOATS; x += TP_THREADS) epi[x] = a.epi[x];
  for (int x = tid; x < TP_CWARPS * TP_TI * 20; x += TP_THREADS) part[x] = 0.f;
  if (tid < TP_TI) {
    const bool v = tid < rows_valid;
    const size_t node = (size_t)b * N + (v ? i0 + tid : i0);
    xis[tid * 4 + 0] = a.coors[node * 3 + 0]; xis[tid * 4 + 1] = a.coors[node * 3 + 1]; xis[tid * 4 + 2] = a.coors[node * 3 + 2];
    xis[tid * 4 + 3] = 0.f;
    mki[tid] = v && (a.has_mask ? a.mask[node] != 0 : true);
  }
  // rows beyond the graph: zero A (their pairs are discarded anyway, keep them finite)
  for (int x = tid + rows_valid * Hp; x < TP_TI * Hp; x += TP_THREADS) As[x] = 0.f;
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = mki[15];

  if (tid == 0) {
    // TMA bulk staging: W2 slabs, the A_i rows of this CTA (contiguous in the table), wd
    const uint32_t w2_bytes = (uint32_t)Hp * 32, as_bytes = (uint32_t)rows_valid * Hp * 4, wd_bytes = (uint32_t)Hp * 4;
    tc::mbar_arrive_expect_tx(ldbar, w2_bytes + as_bytes + wd_bytes);
    auto bulk = [&](uint32_t dst, const unsigned char* src, uint32_t bytes) {
      for (uint32_t o = 0; o < bytes; o += 16384) tc::tma_bulk_g2s(dst + o, src + o, min(16384u, bytes - o), ldbar);
    };
    bulk(tc::smem_u32(w2s), reinterpret_cast<const unsigned char*>(a.w2p), w2_bytes);
    bulk(tc::smem_u32(As), reinterpret_cast<const unsigned char*>(a.Atab + ((size_t)b * N + i0) * Hp), as_bytes);
    bulk(tc::smem_u32(wds), reinterpret_cast<const unsigned char*>(a.wdh), wd_bytes);
  }

  {
    // =========================================================== compute warpgroups
    const int g = warp >> 2, wq = warp & 3, t128 = tid & 127;
    const int lr = lane >> 2, lq = lane & 3;
    const uint32_t tm_wg = tmem + g * TP_WGCOLS + ((uint32_t)(wq * 32) << 16);    // this warp's lane quadrant
    float* mypart = part + (size_t)warp * TP_TI * 20;
    float* dwg = dsm + g * 128;                       // d_ij of this warpgroup's tile: dwg[i * TP_JB + pair]
    constexpr uint32_t IDESC = tc::idesc_bf16_f32(128, 16);
    const uint32_t w2a = tc::smem_u32(w2s);           // W2 slab s at +512*s: K-adjacent core matrices 256 B apart, N-adjacent 128 B
    tc::mbar_wait(ldbar, 0);
    uint32_t n = 0;                                   // rounds this warpgroup has produced
    for (int jb = 0; jb < njb; ++jb) {
      if (jb * TP_JB + g * 128 >= N) break;           // this warpgroup's tile lies beyond the graph
      // ---- pair mapping: geometry of (i, j) for the rows i
      const int j = jb * TP_JB + g * 128 + t128;
      const bool jv = j < N;
      const size_t nodej = (size_t)b * N + (jv ? j : N - 1);
      const float xj0 = a.coors[nodej * 3 + 0], xj1 = a.coors[nodej * 3 + 1], xj2 = a.coors[nodej * 3 + 2];
      const bool mask_j = jv && (a.has_mask ? a.mask[nodej] != 0 : true);
      __syncwarp();                                   // previous block's readers of dwg are done
#pragma unroll
      for (int i = 0; i < TP_TI; ++i) {
        const float r0 = xis[i * 4 + 0] - xj0, r1 = xis[i * 4 + 1] - xj1, r2 = xis[i * 4 + 2] - xj2;
        dwg[i * TP_JB + t128] = r0 * r0 + r1 * r1 + r2 * r2;
      }
      __syncwarp();
      // ---- fragment mapping: B' rows of this lane's 4 pairs
      const uint2* Bp[4];
#pragma unroll
      for (int rho = 0; rho < 4; ++rho) {
        const int jr = jb * TP_JB + g * 128 + wq * 32 + lr + 8 * rho;
        Bp[rho] = reinterpret_cast<const uint2*>(a.Btab + ((size_t)b * N + min(jr, N - 1)) * Hp + 4 * lq);
      }
      uint2 Bc[4][4];                                 // [rho][slab] B' of the current chunk (bf16 x4 each)
#pragma unroll
      for (int rho = 0; rho < 4; ++rho)
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) Bc[rho][sl] = __ldg(Bp[rho] + sl * 4);

      int pend_c = 0, pend_i = 0;
      uint32_t pend_n = 0;
      bool pend_valid = false;
      auto issue_pending = [&]() {
        if (!pend_valid) return;
        pend_valid = false;
        if ((pend_n & 3u) != (uint32_t)wq) return;      // rotating duty: warp (round % 4) of the warpgroup issues
        const uint32_t pslot = pend_n & (TP_SLOTS - 1);
        tc::mbar_wait(&full[g * TP_SLOTS + pslot], (pend_n / TP_SLOTS) & 1);
        tc::tc_fence_after();
        if (lane == 0) {
          const uint32_t tm_g = tmem + g * TP_WGCOLS;
#pragma unroll
          for (int kk = 0; kk < TP_KC / 16; ++kk) {
            const uint64_t bd = tc::smem_desc_kmajor_noswizzle(w2a + (uint32_t)(pend_c * 4 + kk) * 512, 256u, 128u);
            tc::mma_ts(tm_g + pend_i * 16, tm_g + TP_TI * 16 + pslot * 32 + kk * 8, bd, IDESC, (pend_c > 0 || kk > 0) ? 1u : 0u);
          }
          tc::mma_commit(&empty[g * TP_SLOTS + pslot]);
          if (pend_c + 1 == nchunks && pend_i + 1 == TP_TI) tc::mma_commit(&accdone[g]);
        }
        __syncwarp();
      };
      for (int c = 0; c < nchunks; ++c) {
        float4 wdr[4];                                // wd of this lane's 16 channels of the chunk
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) wdr[sl] = *reinterpret_cast<const float4*>(wds + c * TP_KC + sl * 16 + lq * 4);
        // One round = the 64 hidden channels of chunk c for row i and this warp's 32 pairs.  In the last round
        // of a chunk (`reload`), every B' register is re-filled for chunk c+1 right after its last use, so the
        // L2 latency is covered by the rest of that round without a second register buffer.
        // MMA issue is deferred by half a round: the MMAs of round n-1 are issued (by warp (n-1)%4 of the
        // warpgroup) in the middle of round n, when the other three warps have long arrived, and the slot of
        // round n is only waited for right before its first tcgen05.st.
        auto round = [&](int i, bool reload) {
          const uint32_t slot = n & (TP_SLOTS - 1);
          float dr[4];
#pragma unroll
          for (int rho = 0; rho < 4; ++rho) dr[rho] = dwg[i * TP_JB + wq * 32 + lr + 8 * rho];
          const float* Ai = As + (size_t)i * Hp + c * TP_KC + lq * 4;
          float4 avr[4];
#pragma unroll
          for (int sl = 0; sl < 4; ++sl) avr[sl] = *reinterpret_cast<const float4*>(Ai + sl * 16);
          const uint32_t ta = tm_wg + TP_TI * 16 + slot * 32;
#pragma unroll
          for (int half = 0; half < 2; ++half) {       // rows (lr, lr+8), then (lr+16, lr+24)
            uint32_t hp[16];
#pragma unroll
            for (int sl = 0; sl < 4; ++sl) {
              const float4 av = avr[sl];
              const float4 wv = wdr[sl];
#pragma unroll
              for (int r2 = 0; r2 < 2; ++r2) {
                const int rho = half * 2 + r2;
                const uint2 bb = Bc[rho][sl];
                const float d = dr[rho];
                const float2 dd = make_float2(d, d);
                const float2 z01 = tc::ffma2(make_float2(wv.x, wv.y), dd, make_float2(av.x, av.y));   // wd*d + A'
                const float2 z23 = tc::ffma2(make_float2(wv.z, wv.w), dd, make_float2(av.z, av.w));
                const float2 y01 = make_float2(tc::add_bf16_lo(bb.x, z01.x), tc::add_bf16_hi(bb.x, z01.y));  // + B'
                const float2 y23 = make_float2(tc::add_bf16_lo(bb.y, z23.x), tc::add_bf16_hi(bb.y, z23.y));
                const float2 h01 = tc::ffma2(y01, make_float2(tc::tanh_fast(y01.x), tc::tanh_fast(y01.y)), y01);  // y + y tanh y
                const float2 h23 = tc::ffma2(y23, make_float2(tc::tanh_fast(y23.x), tc::tanh_fast(y23.y)), y23);
                // 16x256b fragment: regs {0,1} of a slab -> row lr (+16), regs {2,3} -> row lr+8 (+24); even k low
                hp[sl * 4 + r2 * 2 + 0] = tc::pack_bf16x2(h01.x, h01.y);
                hp[sl * 4 + r2 * 2 + 1] = tc::pack_bf16x2(h23.x, h23.y);
                if (reload) Bc[rho][sl] = __ldg(Bp[rho] + (c + 1) * 16 + sl * 4);
              }
            }
            if (half == 0) {
              issue_pending();                          // round n-1's MMAs, if this warp has the duty
              tc::mbar_wait(&empty[g * TP_SLOTS + slot], ((n / TP_SLOTS) & 1) ^ 1);
              tc::tc_fence_after();
            }
            tc::tmem_st_16x256b_x4(ta + ((uint32_t)(half * 16) << 16), hp);
          }
          tc::tmem_wait_st();
          tc::tc_fence_before();
          tc::mbar_arrive(&full[g * TP_SLOTS + slot]);
          pend_c = c; pend_i = i; pend_n = n; pend_valid = true;
          ++n;
        };
#pragma unroll 1
        for (int i = 0; i < TP_TI - 1; ++i) round(i, false);
        if (c + 1 < nchunks) round(TP_TI - 1, true);
        else round(TP_TI - 1, false);
      }
      issue_pending();                                  // last round of the block: also signals accdone

      // ---- epilogue of this j-block: accumulators back to the owning thread (pair mapping)
      tc::mbar_wait(&accdone[g], jb & 1);
      tc::tc_fence_after();
      const float* W3 = epi; const float* b3 = epi + 1024; const float* w4 = b3 + 64;
      const float* b2 = w4 + 64; const float* gw = b2 + 16; const float* sc = gw + 16;   // sc: gate_b, b4, scale
      float m[TP_TI][16];                              // m_ij of this thread's pair for the TI rows
#pragma unroll
      for (int i = 0; i < TP_TI; ++i) {
        uint32_t r[16];
        tc::tmem_ld16(tm_wg + i * 16, r);
        tc::tmem_wait_ld();
#pragma unroll
        for (int o = 0; o < 16; ++o) m[i][o] = tc::silu_half_arg(0.5f * (__uint_as_float(r[o]) + b2[o]));    // :183
        if (a.flags & EGNN_FLAG_SOFT_EDGES) {                                                               // :289-290
          float z = sc[0];
#pragma unroll
          for (int o = 0; o < 16; ++o) z = fmaf(gw[o], m[i][o], z);
          const float gate = 0.5f + 0.5f * tc::tanh_fast(0.5f * z);
#pragma unroll
          for (int o = 0; o < 16; ++o) m[i][o] *= gate;
        }
      }
      float wgt[TP_TI];
#pragma unroll
      for (int i = 0; i < TP_TI; ++i) wgt[i] = 0.f;
      if (upd_coors) {                                                                                      // :302-315
#pragma unroll
        for (int i = 0; i < TP_TI; ++i) wgt[i] = sc[1];
        // hidden unit u outermost: one W3 row (4 x LDS.128) serves all TI rows of this pair
#pragma unroll 2
        for (int u = 0; u < 64; ++u) {
          const float4* w3 = reinterpret_cast<const float4*>(W3 + u * 16);
          const float4 wa = w3[0], wb = w3[1], wc = w3[2], wd4 = w3[3];
          const float bu = b3[u], w4u = w4[u];
#pragma unroll
          for (int i = 0; i < TP_TI; ++i) {
            float tt = bu;
            tt = fmaf(wa.x, m[i][0], tt); tt = fmaf(wa.y, m[i][1], tt); tt = fmaf(wa.z, m[i][2], tt); tt = fmaf(wa.w, m[i][3], tt);
            tt = fmaf(wb.x, m[i][4], tt); tt = fmaf(wb.y, m[i][5], tt); tt = fmaf(wb.z, m[i][6], tt); tt = fmaf(wb.w, m[i][7], tt);
            tt = fmaf(wc.x, m[i][8], tt); tt = fmaf(wc.y, m[i][9], tt); tt = fmaf(wc.z, m[i][10], tt); tt = fmaf(wc.w, m[i][11], tt);
            tt = fmaf(wd4.x, m[i][12], tt); tt = fmaf(wd4.y, m[i][13], tt); tt = fmaf(wd4.z, m[i][14], tt); tt = fmaf(wd4.w, m[i][15], tt);
            wgt[i] = fmaf(w4u, tc::silu_half_arg(0.5f * tt), wgt[i]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < TP_TI; ++i) {
        const bool pm = jv && (mki[i] != 0) && (a.has_mask ? mask_j : true);
        float w = wgt[i];
        if (upd_coors) {
          if (!pm) w = 0.f;                                                                                 // :309
          if (a.flags & EGNN_FLAG_CLAMP) w = fminf(fmaxf(w, -a.clamp), a.clamp);                            // :313
          if (a.flags & EGNN_FLAG_NORM_COORS) w *= sc[2] / fmaxf(sqrtf(dwg[i * TP_JB + t128]), 1e-8f);      // :74-77
        }
        float v[20];
        v[16] = w * (xis[i * 4 + 0] - xj0); v[17] = w * (xis[i * 4 + 1] - xj1); v[18] = w * (xis[i * 4 + 2] - xj2);
        v[19] = pm ? 1.f : 0.f;
#pragma unroll
        for (int o = 0; o < 16; ++o) v[o] = pm ? m[i][o] : 0.f;                                             // :322
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
#pragma unroll
          for (int o = 0; o < 20; ++o) v[o] += __shfl_xor_sync(0xffffffffu, v[o], off);
        if (lane == 0) {
#pragma unroll
          for (int o = 0; o < 20; ++o) mypart[i * 20 + o] += v[o];
        }
      }
      tc::tc_fence_before();
    }
  }

  // ---------------- per-row outputs
  tc::tc_fence_before();
  __syncthreads();
  if (tid < TP_TI * 20) {
    const int i = tid / 20, o = tid % 20;
    if (i < rows_valid) {
      float s = 0.f;
#pragma unroll
      for (int wv = 0; wv < TP_CWARPS; ++wv) s += part[(size_t)wv * TP_TI * 20 + i * 20 + o];
      const size_t node = (size_t)b * N + i0 + i;
      if (o < 16) {
        if (upd_feats) {
          float inv = 1.f;
          if (a.flags & EGNN_FLAG_POOL_MEAN) {
            float cnt = 0.f;
            for (int wv = 0; wv < TP_CWARPS; ++wv) cnt += part[(size_t)wv * TP_TI * 20 + i * 20 + 19];
            inv = a.has_mask ? (cnt > 0.f ? 1.f / cnt : 0.f) : 1.f / (float)N;                            // :325-330
          }
          a.m_out[node * a.ldn + o] = __float2bfloat16(s * inv);
        }
      } else if (o < 19) {
        if (upd_coors) a.coors_out[node * 3 + (o - 16)] = xis[i * 4 + (o - 16)] + s;                      // :315
      }
    }
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

}  // namespace egnn
