// Backward of the EGNN layer (SIMT fp32 / fp64), SURVEY.md section 8(f) rank 1.
//
// The reference trains through autograd over the materialised [B,N,J,2E] tensors (egnn_pytorch.py:224-341).
// Here the forward is RECOMPUTED per pair in the split form of simt_kernels.cuh and differentiated by hand:
//
//   bwd1  thread per (i, slot) pair: recompute m_ij, differentiate the pair epilogue (coordinate MLP, CoorsNorm,
//         clamp, masks, gate, pooling) and leave per pair  g_pre2[m] = dL/d(W2 hid + b2), the scalar channels f_q,
//         coef = w_ij * scale and the CoorsNorm part of dL/d(dist)  in a [pairs][R] record; the small parameter
//         gradients (coors_mlp, edge_gate, b2, coors_norm.scale) are reduced per CTA.
//   bwd2  thread per hidden channel h, CTA = (block of i rows) x (128 channels): for every pair recompute
//         pre1[h], hid[h]; g_hid = W2[:,h] . g_pre2; g_pre1 = g_hid * silu'(pre1).  Row sums give dL/dA_i, column
//         sums dL/dB_j, and  dL/dW2, dL/dWq, dL/dTab  accumulate in registers / shared memory; dL/df_q per pair is
//         reduced over h and added to the record.
//   bwd3  thread per pair: dL/d(dist) -> dL/d(rel) -> coordinates (x_i +, x_j -), dL/d(edges).
//
// The per-node GEMMs around them (tables A/B, node MLP, LayerNorm) are differentiated in egnn_backward.cu.
// Nothing of size O(pairs * H) is stored; the record is O(pairs * (m + 2Q + 2)).
#pragma once

#include "common.cuh"
#include "simt_kernels.cuh"

namespace egnn {

template <typename T> __device__ __forceinline__ void atomic_add_t(T* p, T v) { atomicAdd(p, v); }

struct RecLayout {
  int gpre2, f, gf, coef, gdn, R;
};
inline RecLayout rec_layout(const Dims& s, int MP) {
  RecLayout r;
  r.gpre2 = 0;
  r.f = MP;
  r.gf = MP + s.Q;
  r.coef = MP + 2 * s.Q;
  r.gdn = r.coef + 1;
  r.R = round_up_i(r.gdn + 1, 4);
  return r;
}

template <typename T>
struct BwdArgs {
  Dims s;
  SimtPackLayout L;
  RecLayout rl;
  uint32_t flags;
  int has_mask;
  int TS;                     // bwd1 / bwd3: slots per row group
  int TI2;                    // bwd2: rows per CTA
  T clamp;
  const T* P; int ldP;        // forward tables [M][2*Hp]: A | B
  const T* coors;
  const T* edges;
  const uint8_t* labels;
  const uint8_t* mask;
  const int32_t* nbr_idx;
  const uint8_t* nbr_ok;
  const T* packed;
  const T* g_node_in; int ld_g;   // [M][dim+m]; dL/dm_i = columns dim..dim+m  (null when !update_feats)
  const T* g_coors_out;           // [B,N,C]
  const T* pre2;                  // optional: W2 silu(pre1) per pair, row-major [B,N,J][MP]: saved by the forward
                                  // (EgnnLayerIO.pre2_out) or, dense only, recomputed by the register-tiled forward
                                  // kernel; null = bwd1 recomputes it itself
  T* rec;                         // [pairs][R]
  T* gpk;                         // gradient accumulators in SimtPackLayout order (zeroed by the caller)
  T* gP;                          // [M][2*Hp]: dL/dA | dL/dB (zeroed by the caller)
  T* g_coors;                     // [B,N,C], pre-loaded with g_coors_out
  T* g_edges;                     // [B,N,N,edge_dim] | null
  DropCfg drop;                   // the forward's dropout configuration (masks are regenerated, never stored)
};

template <typename T> __device__ __forceinline__ T dsilu_from(T x, T sg) { return sg * (T(1) + x * (T(1) - sg)); }

// sigmoid for the bwd2 kernels: ex2.approx.ftz directly (no range fix-up: +-inf / 0 are the right limits here).
__device__ __forceinline__ float sigmoid_bw(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  return __fdividef(1.0f, 1.0f + e);
}
__device__ __forceinline__ double sigmoid_bw(double x) { return 1.0 / (1.0 + exp(-x)); }

// Record index of pair (b, i, slot).  kNN: row-major (b, i, slot).  Dense: (b, j, i) -- "column-major" -- so that the
// records of 32 consecutive rows i for one neighbour j are contiguous, which is what a bwd2 CTA streams.
template <bool KNN>
__device__ __forceinline__ size_t rec_index(int b, int N, int J, int i, int slot) {
  return KNN ? ((size_t)b * N + i) * J + slot : ((size_t)b * N + slot) * N + i;
}

__device__ __forceinline__ void cp_async_elem(void* smem_dst, const float* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_elem(void* smem_dst, const double* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// =====================================================================================
// bwd1
// =====================================================================================
template <typename T>
inline size_t bwd1_smem_bytes(const Dims& s, const SimtPackLayout& L, bool knn, bool soft) {
  const int U = 4 * s.m;
  size_t n = 0;
  n += (size_t)PAIR_CH * L.MP;                 // W2s
  n += (size_t)s.Q * PAIR_CH;                  // wqs
  if (!knn) n += (size_t)PAIR_CH * 33;         // Bs
  if (s.Q > 1) n += (size_t)s.Q * PAIR_THREADS;  // fs
  n += (size_t)U * L.MP + 2 * U + 2 * L.MP + 4;   // w3s, b3s, w4s, misc
  n += (size_t)(soft ? 3 : 2) * PAIR_THREADS * L.MP;   // mms, gp2s, aux
  n += PAIR_THREADS;                           // gw0s
  n += (size_t)PAIR_THREADS * (U + 1);         // tt
  return round_up(n * sizeof(T), 16) + 16;
}

template <typename T, int MP, bool KNN>
__global__ void __launch_bounds__(PAIR_THREADS)
pair_bwd1_kernel(const BwdArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Dims& s = a.s;
  const int tid = threadIdx.x;
  const int TS = a.TS, TI = PAIR_THREADS / TS;
  const int g = tid / TS, sl = tid % TS;
  const int b = blockIdx.y;
  const int i_raw = blockIdx.x * TI + g;
  const bool row_valid = i_raw < s.N;
  const int i = row_valid ? i_raw : 0;
  const int J = KNN ? s.k : s.N;
  const int U = 4 * s.m, UP = U + 1;
  const int qd = 2 * s.F;
  const bool upd_feats = a.flags & EGNN_FLAG_UPDATE_FEATS;
  const bool upd_coors = a.flags & EGNN_FLAG_UPDATE_COORS;
  const bool soft = a.flags & EGNN_FLAG_SOFT_EDGES;
  const bool normc = a.flags & EGNN_FLAG_NORM_COORS;
  const bool clampf = a.flags & EGNN_FLAG_CLAMP;

  T* W2s = reinterpret_cast<T*>(smem_raw);                 // [CH][MP]
  T* wqs = W2s + PAIR_CH * MP;                             // [Q][CH]
  T* Bs = wqs + s.Q * PAIR_CH;                             // [CH][33]      (dense only)
  T* fs = Bs + (KNN ? 0 : PAIR_CH * 33);                   // [Q][128]      (Q > 1 only)
  T* w3s = fs + (s.Q > 1 ? s.Q * PAIR_THREADS : 0);        // [U][MP]
  T* b3s = w3s + U * MP;                                   // [U]
  T* w4s = b3s + U;                                        // [U]
  T* misc = w4s + U;                                       // b2[MP] | gate_w[MP] | gate_b, b4, scale, 0
  T* mms = misc + 2 * MP + 4;                              // [128][MP]  m_ij (after the gate)
  T* gp2s = mms + PAIR_THREADS * MP;                       // [128][MP]  g_pre2
  T* aux = gp2s + PAIR_THREADS * MP;                       // [128][MP]  g_z * s2 (soft edges)
  T* gw0s = aux + (soft ? PAIR_THREADS * MP : 0);          // [128]
  T* tt = gw0s + PAIR_THREADS;                             // [128][U+1] coors_mlp pre-activations

  const T* pk = a.packed;
  for (int x = tid; x < U * MP; x += PAIR_THREADS) w3s[x] = upd_coors ? pk[a.L.w3 + x] : T(0);
  for (int x = tid; x < U; x += PAIR_THREADS) {
    b3s[x] = upd_coors ? pk[a.L.b3 + x] : T(0);
    w4s[x] = upd_coors ? pk[a.L.w4 + x] : T(0);
  }
  for (int x = tid; x < 2 * MP + 4; x += PAIR_THREADS) misc[x] = pk[a.L.misc + x];
  // (visibility: the __syncthreads at the top of the chunk loop)

  const size_t node_i = (size_t)b * s.N + i;
  const T* xi = a.coors + node_i * s.C;
  const bool mask_i = a.has_mask ? (a.mask[node_i] != 0) : true;
  const T* Arow = a.P + node_i * a.ldP;
  T gxo[PAIR_CMAX];
#pragma unroll
  for (int c = 0; c < PAIR_CMAX; ++c) gxo[c] = (c < s.C) ? a.g_coors_out[node_i * s.C + c] : T(0);

  // pooling factor (egnn_pytorch.py:325-333): 1, 1/J, or 1/count of valid pairs of this row
  T inv = T(1);
  if (upd_feats && (a.flags & EGNN_FLAG_POOL_MEAN)) {
    if (a.has_mask) {
      T cnt = T(0);
      for (int s0 = 0; s0 < J; s0 += TS) {
        const int sidx = s0 + sl;
        bool pv = row_valid && sidx < J;
        int j = 0;
        bool ok = true;
        if (KNN) {
          if (pv) {
            const size_t o = node_i * s.k + sidx;
            j = a.nbr_idx[o];
            ok = a.nbr_ok ? a.nbr_ok[o] != 0 : true;
            if (j < 0) { j = 0; pv = false; }
          }
        } else {
          j = pv ? sidx : 0;
        }
        if (pv && mask_i && a.mask[(size_t)b * s.N + j] != 0 && ok) cnt += T(1);
      }
      for (int off = TS >> 1; off > 0; off >>= 1) cnt += shfl_xor_t<T>(cnt, off);
      inv = cnt > T(0) ? T(1) / cnt : T(0);
    } else {
      inv = T(1) / T(J);
    }
  }

  // GEMM-phase role: coors_mlp.0 row u, over the pairs [p_lo, p_hi)
  const int PG = U > 0 ? PAIR_THREADS / U : 1;
  const int role_u = tid % (U > 0 ? U : 1), role_pg = tid / (U > 0 ? U : 1);
  const bool role_on = upd_coors && role_pg < PG;
  const int p_per = PAIR_THREADS / (PG > 0 ? PG : 1);
  T accW3[MP];
#pragma unroll
  for (int o = 0; o < MP; ++o) accW3[o] = T(0);
  T accb3 = T(0), accw4 = T(0);
  T acc_col_b2 = T(0), acc_col_gw = T(0);          // threads tid < MP: column sums
  T acc_gb = T(0), acc_b4 = T(0), acc_cs = T(0);   // per-thread scalars

  for (int s0 = 0; s0 < J; s0 += TS) {
    const int sidx = s0 + sl;
    const bool pair_exists = row_valid && sidx < J;
    bool pair_valid = pair_exists;
    int j = 0;
    bool ok = true;
    if (KNN) {
      if (pair_valid) {
        const size_t o = node_i * s.k + sidx;
        j = a.nbr_idx[o];
        ok = a.nbr_ok ? a.nbr_ok[o] != 0 : true;
        if (j < 0) { j = 0; pair_valid = false; }
      }
    } else {
      j = pair_valid ? sidx : 0;
    }
    T rel[PAIR_CMAX];
    T d = T(0);
    {
      const T* xj = a.coors + ((size_t)b * s.N + j) * s.C;
#pragma unroll
      for (int c = 0; c < PAIR_CMAX; ++c) {
        rel[c] = T(0);
        if (c < s.C) { rel[c] = xi[c] - xj[c]; d = sq_acc<T>(rel[c], d); }
      }
    }
    if (s.Q > 1) {
      for (int q = 0; q < s.Q; ++q) {
        T f;
        if (q < s.F) f = sin(d / T(1 << q));
        else if (q < 2 * s.F) f = cos(d / T(1 << (q - s.F)));
        else if (q == qd) f = d;
        else f = a.edges[((node_i) * s.N + j) * s.edge_dim + (q - s.Qd)];
        fs[q * PAIR_THREADS + tid] = f;
      }
    }
    int lab = 0;
    if (a.labels) lab = a.labels[node_i * s.N + j];
    const T* Brow = a.P + ((size_t)b * s.N + j) * a.ldP + s.Hp;
    const T* tabrow = pk + a.L.tab + (size_t)lab * s.Hp;

    Pk2<T> accp[MP / 2];
#pragma unroll
    for (int o = 0; o < MP / 2; ++o) accp[o] = Pk2<T>::make(T(0), T(0));

    // ---- forward recompute of W2 silu(pre1) (identical to pair_kernel), unless the caller already did it
    if (a.pre2) {
      __syncthreads();                               // tiles of the previous iteration fully consumed
      if (pair_valid) {                              // (empty slots were never stored: keep their zeros)
        const T* src = a.pre2 + (node_i * (size_t)J + sidx) * MP;
#pragma unroll
        for (int o = 0; o < MP; o += 4) {
          Vec4<T> v;
          v.load_g(src + o);
          accp[o / 2] = Pk2<T>::make(v.v[0], v.v[1]);
          accp[o / 2 + 1] = Pk2<T>::make(v.v[2], v.v[3]);
        }
      }
    }
    for (int c0 = 0; c0 < (a.pre2 ? 0 : s.Hp); c0 += PAIR_CH) {
      const int cn = min(PAIR_CH, s.Hp - c0);
      __syncthreads();
      for (int x = tid; x < cn * MP; x += PAIR_THREADS) W2s[x] = pk[a.L.w2t + (size_t)c0 * MP + x];
      for (int x = tid; x < s.Q * cn; x += PAIR_THREADS) {
        int q = x / cn, cc = x % cn;
        wqs[q * PAIR_CH + cc] = pk[a.L.wq + (size_t)q * s.Hp + c0 + cc];
      }
      if (!KNN) {
        const int cc = tid % PAIR_CH, jj0 = tid / PAIR_CH;
        for (int jj = jj0; jj < 32; jj += PAIR_THREADS / PAIR_CH) {
          T v = T(0);
          if (cc < cn && s0 + jj < s.N) v = a.P[((size_t)b * s.N + s0 + jj) * a.ldP + s.Hp + c0 + cc];
          Bs[cc * 33 + jj] = v;
        }
      }
      __syncthreads();
      for (int cc = 0; cc < cn; cc += 4) {
        Vec4<T> av, wd;
        av.load_g(Arow + c0 + cc);
        wd.load(wqs + qd * PAIR_CH + cc);
        T pre[4];
        if (KNN) {
          Vec4<T> bv;
          bv.load_g(Brow + c0 + cc);
#pragma unroll
          for (int u = 0; u < 4; ++u) pre[u] = av.v[u] + bv.v[u];
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) pre[u] = av.v[u] + Bs[(cc + u) * 33 + sl];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) pre[u] = fma_t(wd.v[u], d, pre[u]);
        if (s.Q > 1) {
          for (int q = 0; q < s.Q; ++q) {
            if (q == qd) continue;
            const T f = fs[q * PAIR_THREADS + tid];
            Vec4<T> wv;
            wv.load(wqs + q * PAIR_CH + cc);
#pragma unroll
            for (int u = 0; u < 4; ++u) pre[u] = fma_t(wv.v[u], f, pre[u]);
          }
        }
        if (a.labels) {
          Vec4<T> tv;
          tv.load_g(tabrow + c0 + cc);
#pragma unroll
          for (int u = 0; u < 4; ++u) pre[u] += tv.v[u];
        }
        if (a.drop.thr) {                                // edge_mlp Dropout, same mask as the forward
          const unsigned long long pkey = (((unsigned long long)b * s.N + i) * s.N + j) * s.Hp + c0 + cc;
#pragma unroll
          for (int u = 0; u < 4; ++u) pre[u] *= (T)drop_mul(a.drop, 0u, pkey + u);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const T hv = silu_acc<T>(pre[u]);             // egnn_pytorch.py:181
          const Pk2<T> hdn = Pk2<T>::make(hv, hv);
          const T* w2 = W2s + (cc + u) * MP;
#pragma unroll
          for (int v4 = 0; v4 < MP; v4 += 4) {
            Vec4<T> wv;
            wv.load(w2 + v4);
            accp[v4 / 2].fma(hdn, Pk2<T>::make(wv.v[0], wv.v[1]));
            accp[v4 / 2 + 1].fma(hdn, Pk2<T>::make(wv.v[2], wv.v[3]));
          }
        }
      }
    }
    T acc[MP];
#pragma unroll
    for (int o = 0; o < MP; o += 2) { acc[o] = accp[o / 2].lo(); acc[o + 1] = accp[o / 2].hi(); }

    // ---- pair epilogue, forward
    T mm[MP];
#pragma unroll
    for (int o = 0; o < MP; ++o) mm[o] = silu_acc<T>(acc[o] + misc[o]);      // s2
    T gate = T(1);
    if (soft) {
      T z = misc[2 * MP + 0];
#pragma unroll
      for (int o = 0; o < MP; ++o) z = fma_t(misc[MP + o], mm[o], z);
      gate = sigmoid_acc<T>(z);
#pragma unroll
      for (int o = 0; o < MP; ++o) mm[o] *= gate;
    }
    bool pm = pair_valid;
    if (a.has_mask) {
      const bool mask_j = a.mask[(size_t)b * s.N + j] != 0;
      pm = pm && mask_i && mask_j && (KNN ? ok : true);
    }

    // ---- backward through the coordinate branch (egnn_pytorch.py:302-315 reversed)
    T gmm[MP];
#pragma unroll
    for (int o = 0; o < MP; ++o) gmm[o] = T(0);
    T coef = T(0), gdn = T(0), gw0 = T(0);
    if (upd_coors) {
      T w0 = misc[2 * MP + 1];
      for (int u = 0; u < U; ++u) {
        T t = b3s[u];
        const T* w3 = w3s + u * MP;
#pragma unroll
        for (int o = 0; o < MP; o += 4) {
          Vec4<T> wv;
          wv.load(w3 + o);
#pragma unroll
          for (int z = 0; z < 4; ++z) t = fma_t(wv.v[z], mm[o + z], t);
        }
        if (a.drop.thr) {                                // coors_mlp Dropout: a dropped unit is stored as NaN (silu(0) = 0
          const T f = (T)drop_mul(a.drop, 1u, (((unsigned long long)b * s.N + i) * s.N + j) * U + u);   // adds nothing)
          t = f == T(0) ? T(NAN) : t * f;
        }
        tt[tid * UP + u] = t;
        if (t == t) w0 = fma_t(w4s[u], silu_acc<T>(t), w0);
      }
      const T w1 = pm ? w0 : T(0);
      bool inside = true;
      T w2 = w1;
      if (clampf) {
        inside = (w1 >= -a.clamp) && (w1 <= a.clamp);
        w2 = w1 < -a.clamp ? -a.clamp : (w1 > a.clamp ? a.clamp : w1);
      }
      if (!pair_valid) w2 = T(0);
      T scale = T(1), den = T(1), nrm = T(0);
      const T cs = misc[2 * MP + 2];
      if (normc) {
        nrm = sqrt(d);
        den = nrm > T(1e-8) ? nrm : T(1e-8);
        scale = cs / den;
      }
      coef = w2 * scale;
      T gcoef = T(0);
      if (pair_valid) {
#pragma unroll
        for (int c = 0; c < PAIR_CMAX; ++c) gcoef = fma_t(gxo[c], rel[c], gcoef);
      }
      const T gw2 = gcoef * scale;
      if (normc) {
        const T gscale = gcoef * w2;
        acc_cs += gscale / den;
        const T gnrm = (nrm >= T(1e-8)) ? -gscale * cs / (den * den) : T(0);
        gdn = (nrm > T(0)) ? gnrm / (T(2) * nrm) : T(0);
      }
      gw0 = (pm && inside) ? gw2 : T(0);
      acc_b4 += gw0;
      for (int u = 0; u < U; ++u) {
        const T t = tt[tid * UP + u];
        if (t != t) continue;                            // dropped unit: no gradient
        const T sg = sigmoid_acc<T>(t);
        const T gt = gw0 * w4s[u] * dsilu_from<T>(t, sg) * (a.drop.thr ? (T)a.drop.inv_keep : T(1));
        const T* w3 = w3s + u * MP;
#pragma unroll
        for (int o = 0; o < MP; o += 4) {
          Vec4<T> wv;
          wv.load(w3 + o);
#pragma unroll
          for (int z = 0; z < 4; ++z) gmm[o + z] = fma_t(wv.v[z], gt, gmm[o + z]);
        }
      }
    }
    gw0s[tid] = gw0;
    // ---- pooled message (egnn_pytorch.py:319-333 reversed)
    if (upd_feats && pm) {
      const T* gmi = a.g_node_in + node_i * a.ld_g + s.dim;
#pragma unroll
      for (int o = 0; o < MP; ++o)
        if (o < s.m) gmm[o] = fma_t(gmi[o], inv, gmm[o]);
    }
#pragma unroll
    for (int o = 0; o < MP; ++o) mms[tid * MP + o] = mm[o];
    // ---- gate and the second SiLU (egnn_pytorch.py:287-290 reversed); s2 and silu'(pre2) recomputed from acc
    T gz = T(0);
    if (soft) {
      T ggt = T(0);
#pragma unroll
      for (int o = 0; o < MP; ++o) ggt = fma_t(gmm[o], silu_acc<T>(acc[o] + misc[o]), ggt);
      gz = ggt * gate * (T(1) - gate);
      acc_gb += gz;
    }
#pragma unroll
    for (int o = 0; o < MP; ++o) {
      const T p2 = acc[o] + misc[o];
      const T sg = sigmoid_acc<T>(p2);
      const T s2 = p2 * sg;
      T gs2 = gmm[o];
      if (soft) {
        gs2 = fma_t(gz, misc[MP + o], gmm[o] * gate);
        aux[tid * MP + o] = gz * s2;
      }
      const T gp2 = gs2 * dsilu_from<T>(p2, sg);
      gp2s[tid * MP + o] = gp2;
      gmm[o] = gp2;                                   // reuse as the value written to the record
    }
    if (pair_exists) {
      T* r = a.rec + rec_index<KNN>(b, s.N, J, i, sidx) * a.rl.R;
#pragma unroll
      for (int o = 0; o < MP; ++o) r[a.rl.gpre2 + o] = gmm[o];
      if (s.Q > 1) {
        for (int q = 0; q < s.Q; ++q) r[a.rl.f + q] = fs[q * PAIR_THREADS + tid];
      } else {
        r[a.rl.f] = d;
      }
      r[a.rl.coef] = coef;
      r[a.rl.gdn] = gdn;
    }
    __syncthreads();
    // ---- CTA-level parameter gradients of this tile of 128 pairs
    if (role_on) {
      const T w4u = w4s[role_u];
      for (int p = role_pg * p_per; p < (role_pg + 1) * p_per; ++p) {
        const T g0 = gw0s[p];
        if (g0 == T(0)) continue;
        const T t = tt[p * UP + role_u];
        if (t != t) continue;                            // dropped unit
        const T sg = sigmoid_acc<T>(t);
        const T gt = g0 * w4u * dsilu_from<T>(t, sg) * (a.drop.thr ? (T)a.drop.inv_keep : T(1));
        accb3 += gt;
        accw4 = fma_t(g0, t * sg, accw4);
        const T* mrow = mms + p * MP;
#pragma unroll
        for (int o = 0; o < MP; o += 4) {
          Vec4<T> mv;
          mv.load(mrow + o);
#pragma unroll
          for (int z = 0; z < 4; ++z) accW3[o + z] = fma_t(gt, mv.v[z], accW3[o + z]);
        }
      }
    }
    if (tid < MP) {
      for (int p = 0; p < PAIR_THREADS; ++p) {
        acc_col_b2 += gp2s[p * MP + tid];
        if (soft) acc_col_gw += aux[p * MP + tid];
      }
    }
    // (the next iteration's first __syncthreads orders these reads before the tiles are rewritten)
  }

  // ---- flush
  T* gpk = a.gpk;
  if (role_on) {
#pragma unroll
    for (int o = 0; o < MP; ++o)
      if (o < s.m) atomic_add_t<T>(gpk + a.L.w3 + role_u * MP + o, accW3[o]);
    atomic_add_t<T>(gpk + a.L.b3 + role_u, accb3);
    atomic_add_t<T>(gpk + a.L.w4 + role_u, accw4);
  }
  if (tid < MP) {
    atomic_add_t<T>(gpk + a.L.misc + tid, acc_col_b2);
    if (soft) atomic_add_t<T>(gpk + a.L.misc + MP + tid, acc_col_gw);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    acc_gb += shfl_xor_t<T>(acc_gb, off);
    acc_b4 += shfl_xor_t<T>(acc_b4, off);
    acc_cs += shfl_xor_t<T>(acc_cs, off);
  }
  if ((tid & 31) == 0) {
    if (soft) atomic_add_t<T>(gpk + a.L.misc + 2 * MP + 0, acc_gb);
    if (upd_coors) atomic_add_t<T>(gpk + a.L.misc + 2 * MP + 1, acc_b4);
    if (normc) atomic_add_t<T>(gpk + a.L.misc + 2 * MP + 2, acc_cs);
  }
}

// =====================================================================================
// bwd2, neighbour-list form: thread = hidden channel, CTA = TI2 rows x 128 channels.  One step = up to 32 slots of
// one row: their records are contiguous (rec_index<true>) and are prefetched one step ahead with cp.async together
// with the neighbour indices; the 32 gathered B rows are loaded up front (32 independent L2 requests per thread),
// dL/dA_i is a register, dL/dB_j one fire-and-forget atomic per (pair, channel) -- the scatter is inherent to a
// neighbour list -- and dL/df_q is reduced over the channels through a [32][128] shared tile.
// =====================================================================================
constexpr int BW2_TH = 128;       // channels per CTA
constexpr int BW2_PB = 32;        // pairs per step
constexpr int BW2_MAXLAB = 16;    // label rows kept in shared memory

template <typename T>
inline size_t bwd2_knn_smem_bytes(const Dims& s, int R) {
  const int NL = s.label_dim > 0 ? s.num_labels : 0;
  size_t n = 0;
  n += (size_t)2 * s.Q * BW2_TH;            // wqs, gwqs
  n += (size_t)2 * NL * BW2_TH;             // tabs, gtabs
  n += (size_t)2 * BW2_PB * R;              // recs (double buffered)
  n += (size_t)BW2_PB * BW2_TH;             // gps
  return round_up(n * sizeof(T), 16) + 4 * BW2_PB * sizeof(int) + 16;
}

template <typename T, int MP, int QR, bool DROP>
__global__ void __launch_bounds__(BW2_TH)
pair_bwd2_knn_kernel(const BwdArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Dims& s = a.s;
  const int tid = threadIdx.x, lane = tid & 31;
  const int b = blockIdx.z;
  const int i0 = blockIdx.x * a.TI2;
  const int h0 = blockIdx.y * BW2_TH;
  const int hh = h0 + tid;
  const bool hv = hh < s.Hp;
  const int N = s.N, R = a.rl.R, Q = s.Q, k = s.k;
  const int NL = s.label_dim > 0 ? s.num_labels : 0;
  const int nrows = min(a.TI2, N - i0);
  const int nch = ceil_div(k, BW2_PB);
  const int nsteps = nrows * nch;

  T* wqs = reinterpret_cast<T*>(smem_raw);       // [Q][128]
  T* gwqs = wqs + Q * BW2_TH;                    // [Q][128]
  T* tabs = gwqs + Q * BW2_TH;                   // [NL][128]
  T* gtabs = tabs + NL * BW2_TH;                 // [NL][128]
  T* recs = gtabs + NL * BW2_TH;                 // [2][32][R]
  T* gps = recs + 2 * BW2_PB * R;                // [32][128]
  int* nb = reinterpret_cast<int*>(smem_raw + round_up((size_t)(2 * Q * BW2_TH + 2 * NL * BW2_TH + 2 * BW2_PB * R +
                                                                 BW2_PB * BW2_TH) * sizeof(T), 16));      // [2][32] neighbour
  int* lb = nb + 2 * BW2_PB;                                                                              // [2][32] label

  const T* pk = a.packed;
  for (int q = 0; q < Q; ++q) {
    wqs[q * BW2_TH + tid] = hv ? pk[a.L.wq + (size_t)q * s.Hp + hh] : T(0);
    gwqs[q * BW2_TH + tid] = T(0);
  }
  for (int l = 0; l < NL; ++l) {
    tabs[l * BW2_TH + tid] = hv ? pk[a.L.tab + (size_t)l * s.Hp + hh] : T(0);
    gtabs[l * BW2_TH + tid] = T(0);
  }
  Pk2<T> w2p[MP / 2], gW2p[MP / 2];
#pragma unroll
  for (int o = 0; o < MP; o += 2) {
    w2p[o / 2] = Pk2<T>::make(hv ? pk[a.L.w2t + (size_t)hh * MP + o] : T(0), hv ? pk[a.L.w2t + (size_t)hh * MP + o + 1] : T(0));
    gW2p[o / 2] = Pk2<T>::make(T(0), T(0));
  }
  const T wq0 = hv ? pk[a.L.wq + (size_t)(2 * s.F) * s.Hp + hh] : T(0);
  T gwq0 = T(0);
  constexpr bool SIMPLE = QR == 1;          // distance channel only, no label table
  constexpr bool QREG = QR > 1;             // Q <= QR: per-pair scalar channels and their weights live in registers
  constexpr int QN = QREG ? QR : 1;
  T wqr[QN], gwqr[QN];
#pragma unroll
  for (int q = 0; q < QN; ++q) {
    wqr[q] = (QREG && q < Q && hv) ? pk[a.L.wq + (size_t)q * s.Hp + hh] : T(0);
    gwqr[q] = T(0);
  }

  auto prefetch = [&](int t, int buf) {
    const int row = t / nch, c0 = (t % nch) * BW2_PB, kc = min(BW2_PB, k - c0);
    const size_t node = (size_t)b * N + i0 + row;
    const T* src = a.rec + rec_index<true>(b, N, k, i0 + row, c0) * R;
    T* dst = recs + buf * BW2_PB * R;
    for (int x = tid; x < kc * R; x += BW2_TH) cp_async_elem(dst + x, src + x);
    if (tid < BW2_PB) {
      const int j = tid < kc ? a.nbr_idx[node * k + c0 + tid] : -1;
      nb[buf * BW2_PB + tid] = j;
      lb[buf * BW2_PB + tid] = (NL && j >= 0) ? a.labels[node * N + j] : 0;
    }
    cp_async_commit_group();
  };
  if (nsteps > 0) prefetch(0, 0);
  cp_async_wait_all();
  __syncthreads();

  int cur_row = -1;
  T Ai = T(0), gA = T(0);
  for (int t = 0; t < nsteps; ++t) {
    const int cur = t & 1;
    const int row = t / nch, c0 = (t % nch) * BW2_PB, kc = min(BW2_PB, k - c0);
    if (t + 1 < nsteps) prefetch(t + 1, cur ^ 1);
    if (row != cur_row) {
      if (cur_row >= 0 && hv) a.gP[((size_t)b * N + i0 + cur_row) * a.ldP + hh] = gA;
      gA = T(0);
      cur_row = row;
      Ai = hv ? a.P[((size_t)b * N + i0 + row) * a.ldP + hh] : T(0);
    }
    const int* nbc = nb + cur * BW2_PB;
    const T* rb = recs + cur * BW2_PB * R;
    T bjv[BW2_PB];
#pragma unroll
    for (int p = 0; p < BW2_PB; ++p) {
      const int j = nbc[p];
      bjv[p] = (j >= 0 && hv) ? a.P[((size_t)b * N + j) * a.ldP + s.Hp + hh] : T(0);
    }
#pragma unroll
    for (int p = 0; p < BW2_PB; ++p) {
      if (p >= kc) break;                              // uniform
      const int j = nbc[p];
      T gp = T(0);
      if (j >= 0) {                                    // uniform
        const T* r = rb + p * R;
        T pre = Ai + bjv[p];
        int lab = 0;
        T fq[QN];
        if (SIMPLE) {
          pre = fma_t(wq0, r[MP], pre);
        } else {
          if (QREG) {
#pragma unroll
            for (int q = 0; q < QN; ++q)
              if (q < Q) { fq[q] = r[MP + q]; pre = fma_t(wqr[q], fq[q], pre); }
          } else {
            for (int q = 0; q < Q; ++q) pre = fma_t(wqs[q * BW2_TH + tid], r[MP + q], pre);
          }
          if (NL) { lab = lb[cur * BW2_PB + p]; pre += tabs[lab * BW2_TH + tid]; }
        }
        T fdrop = T(1);
        if (DROP) {                                      // edge_mlp Dropout: same mask as the forward (own instantiation: the
                                                         // key arithmetic costs 50 registers when unrolled over the rows)
          fdrop = (T)drop_mul(a.drop, 0u, (((unsigned long long)b * N + i0 + row) * N + j) * s.Hp + hh);
          pre *= fdrop;
        }
        const T sg = sigmoid_bw(pre);
        const T a1 = pre * sg;
        Pk2<T> ga1p = Pk2<T>::make(T(0), T(0));
        const Pk2<T> a1p = Pk2<T>::make(a1, a1);
#pragma unroll
        for (int o = 0; o < MP; o += 4) {
          Vec4<T> gv;
          gv.load(r + o);
          const Pk2<T> g01 = Pk2<T>::make(gv.v[0], gv.v[1]), g23 = Pk2<T>::make(gv.v[2], gv.v[3]);
          ga1p.fma(w2p[o / 2], g01);
          ga1p.fma(w2p[o / 2 + 1], g23);
          gW2p[o / 2].fma(a1p, g01);
          gW2p[o / 2 + 1].fma(a1p, g23);
        }
        const T ga1 = ga1p.lo() + ga1p.hi();
        gp = ga1 * dsilu_from<T>(pre, sg);
        if (DROP) gp *= fdrop;
        gA += gp;
        if (hv) atomic_add_t<T>(a.gP + ((size_t)b * N + j) * a.ldP + s.Hp + hh, gp);
        if (SIMPLE) {
          gwq0 = fma_t(r[MP], gp, gwq0);
        } else {
          if (QREG) {
#pragma unroll
            for (int q = 0; q < QN; ++q)
              if (q < Q) gwqr[q] = fma_t(fq[q], gp, gwqr[q]);
          } else {
            for (int q = 0; q < Q; ++q) gwqs[q * BW2_TH + tid] = fma_t(r[MP + q], gp, gwqs[q * BW2_TH + tid]);
          }
          if (NL) gtabs[lab * BW2_TH + tid] += gp;
        }
      }
      gps[p * BW2_TH + tid] = SIMPLE ? wq0 * gp : gp;      // SIMPLE: the tile holds Wq[h] * g_pre1 already
    }
    __syncthreads();                                   // gps complete
    {
      const int p = tid >> 2, qt = tid & 3;
      const bool live = p < kc && nbc[p] >= 0;
      const T* grow = gps + p * BW2_TH + qt * 32;
      for (int q = 0; q < Q; ++q) {
        const T* wrow = wqs + q * BW2_TH + qt * 32;
        T v = T(0);
        if (live && SIMPLE) {
#pragma unroll
          for (int kx = 0; kx < 8; ++kx) {
            Vec4<T> t;
            t.load(grow + 4 * ((kx + lane) & 7));
            v += (t.v[0] + t.v[1]) + (t.v[2] + t.v[3]);
          }
        } else if (live) {
#pragma unroll 8
          for (int kx = 0; kx < 32; ++kx) {
            const int kk = (kx + lane) & 31;
            v = fma_t(wrow[kk], grow[kk], v);
          }
        }
        v += shfl_xor_t<T>(v, 1);
        v += shfl_xor_t<T>(v, 2);
        if (qt == 0 && live) atomic_add_t<T>(a.rec + rec_index<true>(b, N, k, i0 + row, c0 + p) * R + a.rl.gf + q, v);
      }
    }
    cp_async_wait_all();
    __syncthreads();                                   // next records landed; gps free
  }
  if (hv) {
    if (cur_row >= 0) a.gP[((size_t)b * N + i0 + cur_row) * a.ldP + hh] = gA;
#pragma unroll
    for (int o = 0; o < MP; o += 2) {
      atomic_add_t<T>(a.gpk + a.L.w2t + (size_t)hh * MP + o, gW2p[o / 2].lo());
      atomic_add_t<T>(a.gpk + a.L.w2t + (size_t)hh * MP + o + 1, gW2p[o / 2].hi());
    }
    if (SIMPLE) {
      atomic_add_t<T>(a.gpk + a.L.wq + (size_t)(2 * s.F) * s.Hp + hh, gwq0);
    } else {
      if (QREG) {
#pragma unroll
        for (int q = 0; q < QN; ++q)
          if (q < Q) atomic_add_t<T>(a.gpk + a.L.wq + (size_t)q * s.Hp + hh, gwqr[q]);
      } else {
        for (int q = 0; q < Q; ++q) atomic_add_t<T>(a.gpk + a.L.wq + (size_t)q * s.Hp + hh, gwqs[q * BW2_TH + tid]);
      }
      for (int l = 0; l < NL; ++l) atomic_add_t<T>(a.gpk + a.L.tab + (size_t)l * s.Hp + hh, gtabs[l * BW2_TH + tid]);
    }
  }
}

// =====================================================================================
// bwd2, dense all-pairs specialisation: CTA = 32 rows x 128 channels, one neighbour j per step.  The 32 records of
// (i0..i0+31, j) are contiguous (rec_index<false>) and are prefetched one step ahead with cp.async; row sums dL/dA live
// in registers (the pair loop is unrolled over the 32 rows), the column sum dL/dB_j is one atomic per step, and
// dL/df_q is reduced over the 128 channels through a [32][128] shared tile instead of per-pair shuffles.
// SIMPLE = only the distance channel (Q == 1) and no label table: the common EGNN(dim) configuration.
// =====================================================================================
constexpr int BW2_ROWS = 32;

template <typename T>
inline size_t bwd2_dense_smem_bytes(const Dims& s, int R) {
  const int NL = s.label_dim > 0 ? s.num_labels : 0;
  size_t n = 0;
  n += (size_t)BW2_ROWS * BW2_TH;           // As
  n += (size_t)2 * s.Q * BW2_TH;            // wqs, gwqs
  n += (size_t)2 * NL * BW2_TH;             // tabs, gtabs
  n += (size_t)2 * BW2_ROWS * R;            // recs (double buffered)
  n += (size_t)BW2_ROWS * BW2_TH;           // gps
  return round_up(n * sizeof(T), 16) + 2 * BW2_ROWS * sizeof(int) + 16;
}

template <typename T, int MP, int QR, bool DROP>
__global__ void __launch_bounds__(BW2_TH)
pair_bwd2_dense_kernel(const BwdArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const Dims& s = a.s;
  const int tid = threadIdx.x, lane = tid & 31;
  const int b = blockIdx.z;
  const int i0 = blockIdx.x * BW2_ROWS;
  const int h0 = blockIdx.y * BW2_TH;
  const int hh = h0 + tid;
  const bool hv = hh < s.Hp;
  const int N = s.N, R = a.rl.R, Q = s.Q;
  const int NL = s.label_dim > 0 ? s.num_labels : 0;
  const int nrows = min(BW2_ROWS, N - i0);

  T* As = reinterpret_cast<T*>(smem_raw);        // [32][128]
  T* wqs = As + BW2_ROWS * BW2_TH;               // [Q][128]
  T* gwqs = wqs + Q * BW2_TH;                    // [Q][128]
  T* tabs = gwqs + Q * BW2_TH;                   // [NL][128]
  T* gtabs = tabs + NL * BW2_TH;                 // [NL][128]
  T* recs = gtabs + NL * BW2_TH;                 // [2][32][R]
  T* gps = recs + 2 * BW2_ROWS * R;              // [32][128]
  int* labs = reinterpret_cast<int*>(smem_raw + round_up((size_t)(2 * BW2_ROWS * BW2_TH + 2 * Q * BW2_TH + 2 * NL * BW2_TH +
                                                                   2 * BW2_ROWS * R) * sizeof(T), 16));   // [2][32]

  const T* pk = a.packed;
  for (int r = 0; r < BW2_ROWS; ++r)
    As[r * BW2_TH + tid] = (r < nrows && hv) ? a.P[((size_t)b * N + i0 + r) * a.ldP + hh] : T(0);
  for (int q = 0; q < Q; ++q) {
    wqs[q * BW2_TH + tid] = hv ? pk[a.L.wq + (size_t)q * s.Hp + hh] : T(0);
    gwqs[q * BW2_TH + tid] = T(0);
  }
  for (int l = 0; l < NL; ++l) {
    tabs[l * BW2_TH + tid] = hv ? pk[a.L.tab + (size_t)l * s.Hp + hh] : T(0);
    gtabs[l * BW2_TH + tid] = T(0);
  }
  for (int x = tid; x < 2 * BW2_ROWS * R; x += BW2_TH) recs[x] = T(0);     // rows >= nrows stay zero
  if (tid < 2 * BW2_ROWS) labs[tid] = 0;
  T gA[BW2_ROWS];
  Pk2<T> w2p[MP / 2], gW2p[MP / 2];
#pragma unroll
  for (int o = 0; o < MP; o += 2) {
    w2p[o / 2] = Pk2<T>::make(hv ? pk[a.L.w2t + (size_t)hh * MP + o] : T(0), hv ? pk[a.L.w2t + (size_t)hh * MP + o + 1] : T(0));
    gW2p[o / 2] = Pk2<T>::make(T(0), T(0));
  }
#pragma unroll
  for (int p = 0; p < BW2_ROWS; ++p) gA[p] = T(0);
  const T wq0 = hv ? pk[a.L.wq + (size_t)(2 * s.F) * s.Hp + hh] : T(0);   // SIMPLE: the distance column
  T gwq0 = T(0);
  constexpr bool SIMPLE = QR == 1;          // distance channel only, no label table
  constexpr bool QREG = QR > 1;             // Q <= QR: per-pair scalar channels and their weights live in registers
  constexpr int QN = QREG ? QR : 1;
  T wqr[QN], gwqr[QN];
#pragma unroll
  for (int q = 0; q < QN; ++q) {
    wqr[q] = (QREG && q < Q && hv) ? pk[a.L.wq + (size_t)q * s.Hp + hh] : T(0);
    gwqr[q] = T(0);
  }
  __syncthreads();

  auto prefetch = [&](int j, int buf) {
    const T* src = a.rec + rec_index<false>(b, N, N, i0, j) * R;
    T* dst = recs + buf * BW2_ROWS * R;
    for (int x = tid; x < nrows * R; x += BW2_TH) cp_async_elem(dst + x, src + x);
    if (!SIMPLE && NL && tid < nrows) labs[buf * BW2_ROWS + tid] = a.labels[((size_t)b * N + i0 + tid) * N + j];
    cp_async_commit_group();
  };
  prefetch(0, 0);
  T bj_next = hv ? a.P[((size_t)b * N) * a.ldP + s.Hp + hh] : T(0);
  cp_async_wait_all();
  __syncthreads();

  for (int j = 0; j < N; ++j) {
    const int cur = j & 1;
    if (j + 1 < N) prefetch(j + 1, cur ^ 1);
    const T bj = bj_next;
    if (j + 1 < N) bj_next = hv ? a.P[((size_t)b * N + j + 1) * a.ldP + s.Hp + hh] : T(0);
    const T* rb = recs + cur * BW2_ROWS * R;
    T gB = T(0);
#pragma unroll
    for (int p = 0; p < BW2_ROWS; ++p) {
      const T* r = rb + p * R;
      T pre = As[p * BW2_TH + tid] + bj;
      int lab = 0;
      T fq[QN];
      if (SIMPLE) {
        pre = fma_t(wq0, r[MP], pre);
      } else {
        if (QREG) {
#pragma unroll
          for (int q = 0; q < QN; ++q)
            if (q < Q) { fq[q] = r[MP + q]; pre = fma_t(wqr[q], fq[q], pre); }
        } else {
          for (int q = 0; q < Q; ++q) pre = fma_t(wqs[q * BW2_TH + tid], r[MP + q], pre);
        }
        if (NL) { lab = labs[cur * BW2_ROWS + p]; pre += tabs[lab * BW2_TH + tid]; }
      }
      T fdrop = T(1);
      if (DROP) {                                        // edge_mlp Dropout: same mask as the forward
        fdrop = (T)drop_mul(a.drop, 0u, (((unsigned long long)b * N + i0 + p) * N + j) * s.Hp + hh);
        pre *= fdrop;
      }
      const T sg = sigmoid_bw(pre);
      const T a1 = pre * sg;
      Pk2<T> ga1p = Pk2<T>::make(T(0), T(0));
      const Pk2<T> a1p = Pk2<T>::make(a1, a1);
#pragma unroll
      for (int o = 0; o < MP; o += 4) {
        Vec4<T> gv;
        gv.load(r + o);
        const Pk2<T> g01 = Pk2<T>::make(gv.v[0], gv.v[1]), g23 = Pk2<T>::make(gv.v[2], gv.v[3]);
        ga1p.fma(w2p[o / 2], g01);
        ga1p.fma(w2p[o / 2 + 1], g23);
        gW2p[o / 2].fma(a1p, g01);
        gW2p[o / 2 + 1].fma(a1p, g23);
      }
      const T ga1 = ga1p.lo() + ga1p.hi();
      T gp = ga1 * dsilu_from<T>(pre, sg);
      if (DROP) gp *= fdrop;
      gA[p] += gp;
      gB += gp;
      gps[p * BW2_TH + tid] = SIMPLE ? wq0 * gp : gp;      // SIMPLE: the tile holds Wq[h] * g_pre1 already
      if (SIMPLE) {
        gwq0 = fma_t(r[MP], gp, gwq0);
      } else {
        if (QREG) {
#pragma unroll
          for (int q = 0; q < QN; ++q)
            if (q < Q) gwqr[q] = fma_t(fq[q], gp, gwqr[q]);
        } else {
          for (int q = 0; q < Q; ++q) gwqs[q * BW2_TH + tid] = fma_t(r[MP + q], gp, gwqs[q * BW2_TH + tid]);
        }
        if (NL) gtabs[lab * BW2_TH + tid] += gp;
      }
    }
    if (hv) atomic_add_t<T>(a.gP + ((size_t)b * N + j) * a.ldP + s.Hp + hh, gB);
    __syncthreads();                                   // gps complete
    {
      // dL/df_q(i0+p, j) = sum_h Wq[q][h] gp[p][h]: thread = (row p, quarter of the channels); lanes start at
      // rotated offsets so that the 32 lanes of a warp hit 32 different banks
      const int p = tid >> 2, qt = tid & 3;
      const T* grow = gps + p * BW2_TH + qt * 32;
      for (int q = 0; q < Q; ++q) {
        const T* wrow = wqs + q * BW2_TH + qt * 32;
        T v = T(0);
        if (SIMPLE) {                                   // plain row sum, 4 values per load, rotated start per lane
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            Vec4<T> t;
            t.load(grow + 4 * ((k + lane) & 7));
            v += (t.v[0] + t.v[1]) + (t.v[2] + t.v[3]);
          }
        } else {
#pragma unroll 8
          for (int k = 0; k < 32; ++k) {
            const int kk = (k + lane) & 31;
            v = fma_t(wrow[kk], grow[kk], v);
          }
        }
        v += shfl_xor_t<T>(v, 1);
        v += shfl_xor_t<T>(v, 2);
        if (qt == 0 && p < nrows) atomic_add_t<T>(a.rec + rec_index<false>(b, N, N, i0 + p, j) * R + a.rl.gf + q, v);
      }
    }
    cp_async_wait_all();
    __syncthreads();                                   // next records landed; gps free
  }
  if (hv) {
#pragma unroll
    for (int p = 0; p < BW2_ROWS; ++p)
      if (p < nrows) a.gP[((size_t)b * N + i0 + p) * a.ldP + hh] = gA[p];
#pragma unroll
    for (int o = 0; o < MP; o += 2) {
      atomic_add_t<T>(a.gpk + a.L.w2t + (size_t)hh * MP + o, gW2p[o / 2].lo());
      atomic_add_t<T>(a.gpk + a.L.w2t + (size_t)hh * MP + o + 1, gW2p[o / 2].hi());
    }
    if (SIMPLE) {
      atomic_add_t<T>(a.gpk + a.L.wq + (size_t)(2 * s.F) * s.Hp + hh, gwq0);
    } else {
      if (QREG) {
#pragma unroll
        for (int q = 0; q < QN; ++q)
          if (q < Q) atomic_add_t<T>(a.gpk + a.L.wq + (size_t)q * s.Hp + hh, gwqr[q]);
      } else {
        for (int q = 0; q < Q; ++q) atomic_add_t<T>(a.gpk + a.L.wq + (size_t)q * s.Hp + hh, gwqs[q * BW2_TH + tid]);
      }
      for (int l = 0; l < NL; ++l) atomic_add_t<T>(a.gpk + a.L.tab + (size_t)l * s.Hp + hh, gtabs[l * BW2_TH + tid]);
    }
  }
}

// =====================================================================================
// bwd3: dL/d(dist) -> coordinates and edges.  Same thread <-> pair mapping as bwd1.
// =====================================================================================
template <typename T, bool KNN>
__global__ void __launch_bounds__(PAIR_THREADS)
pair_bwd3_kernel(const BwdArgs<T> a) {
  const Dims& s = a.s;
  const int tid = threadIdx.x;
  const int TS = a.TS, TI = PAIR_THREADS / TS;
  const int g = tid / TS, sl = tid % TS;
  const int b = blockIdx.y;
  const int i_raw = blockIdx.x * TI + g;
  const bool row_valid = i_raw < s.N;
  const int i = row_valid ? i_raw : 0;
  const int J = KNN ? s.k : s.N;
  const int qd = 2 * s.F;
  const size_t node_i = (size_t)b * s.N + i;
  const T* xi = a.coors + node_i * s.C;
  T gxo[PAIR_CMAX], gxi[PAIR_CMAX];
#pragma unroll
  for (int c = 0; c < PAIR_CMAX; ++c) {
    gxo[c] = (c < s.C) ? a.g_coors_out[node_i * s.C + c] : T(0);
    gxi[c] = T(0);
  }
  for (int s0 = 0; s0 < J; s0 += TS) {
    const int sidx = s0 + sl;
    if (!(row_valid && sidx < J)) continue;
    int j = KNN ? a.nbr_idx[node_i * s.k + sidx] : sidx;
    if (j < 0) continue;
    const T* r = a.rec + rec_index<KNN>(b, s.N, J, i, sidx) * a.rl.R;
    const T* xj = a.coors + ((size_t)b * s.N + j) * s.C;
    T rel[PAIR_CMAX];
    T d = T(0);
#pragma unroll
    for (int c = 0; c < PAIR_CMAX; ++c) {
      rel[c] = T(0);
      if (c < s.C) { rel[c] = xi[c] - xj[c]; d = sq_acc<T>(rel[c], d); }
    }
    T gd = r[a.rl.gf + qd] + r[a.rl.gdn];
    for (int q = 0; q < s.F; ++q) {                        // fourier_encode_dist :34-41 reversed
      const T sc = T(1 << q);
      gd += (r[a.rl.gf + q] * cos(d / sc) - r[a.rl.gf + s.F + q] * sin(d / sc)) / sc;
    }
    if (a.g_edges) {
      T* ge = a.g_edges + (node_i * s.N + j) * s.edge_dim;
      for (int e = 0; e < s.edge_dim; ++e) {
        if (KNN) atomic_add_t<T>(ge + e, r[a.rl.gf + s.Qd + e]);
        else ge[e] = r[a.rl.gf + s.Qd + e];
      }
    }
    if (j == i) continue;        // x_i - x_i: the two contributions cancel exactly (and would be 1/eps-sized under CoorsNorm)
    const T coef = r[a.rl.coef];
    T* gxj = a.g_coors + ((size_t)b * s.N + j) * s.C;
#pragma unroll
    for (int c = 0; c < PAIR_CMAX; ++c) {
      if (c < s.C) {
        const T gr = fma_t(coef, gxo[c], T(2) * gd * rel[c]);
        gxi[c] += gr;
        atomic_add_t<T>(gxj + c, -gr);
      }
    }
  }
  for (int off = TS >> 1; off > 0; off >>= 1) {
#pragma unroll
    for (int c = 0; c < PAIR_CMAX; ++c) gxi[c] += shfl_xor_t<T>(gxi[c], off);
  }
  if (sl == 0 && row_valid) {
#pragma unroll
    for (int c = 0; c < PAIR_CMAX; ++c)
      if (c < s.C) atomic_add_t<T>(a.g_coors + node_i * s.C + c, gxi[c]);
  }
}

// =====================================================================================
// Per-node pieces
// =====================================================================================
// C[r, c] += sum_k A(r, k) B(k, c),  A(r,k) = A[r*ars + k*aks],  B(k,c) = B[k*bks + c*bcs]; K split over gridDim.z.
// Always accumulates with atomics: the caller zero-fills C or wants the sum.
template <typename T>
__global__ void __launch_bounds__(256)
gemm_acc_kernel(const T* __restrict__ A, long ars, long aks, const T* __restrict__ Bm, long bks, long bcs,
                T* __restrict__ Cm, long ldc, int Mr, int Nc, int K, int kper) {
  __shared__ T As[16][64 + 4];
  __shared__ T Bs[16][64 + 4];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int kbeg = blockIdx.z * kper, kend = min(K, kbeg + kper);
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = T(0);
  const bool a_rows_fast = ars == 1, b_cols_fast = bcs == 1;
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + r * 256;
      {
        const int row = a_rows_fast ? idx % 64 : idx / 16, kk = a_rows_fast ? idx / 64 : idx % 16;
        T v = T(0);
        if (m0 + row < Mr && k0 + kk < kend) v = A[(long)(m0 + row) * ars + (long)(k0 + kk) * aks];
        As[kk][row] = v;
      }
      {
        const int col = b_cols_fast ? idx % 64 : idx / 16, kk = b_cols_fast ? idx / 64 : idx % 16;
        T v = T(0);
        if (n0 + col < Nc && k0 + kk < kend) v = Bm[(long)(k0 + kk) * bks + (long)(n0 + col) * bcs];
        Bs[kk][col] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      T av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = As[kk][ty * 4 + i]; bv[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma_t(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
    if (r >= Mr) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c < Nc) atomic_add_t<T>(Cm + (long)r * ldc + c, acc[i][j]);
    }
  }
}

// out[c] += sum_r X[r*ld + c]
template <typename T>
__global__ void colsum_acc_kernel(const T* __restrict__ X, long ld, int rows, int cols, T* __restrict__ out) {
  __shared__ T part[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  T s = T(0);
  if (c < cols)
    for (int r = blockIdx.y * 8 + threadIdx.y; r < rows; r += gridDim.y * 8) s += X[(long)r * ld + c];
  part[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    T t = T(0);
#pragma unroll
    for (int y = 0; y < 8; ++y) t += part[y][threadIdx.x];
    atomic_add_t<T>(out + c, t);
  }
}

// g[x] = g[x] * silu'(drop(pre[x])) * drop'   (node_mlp: Linear -> Dropout -> SiLU, egnn_pytorch.py:197-199; x = row * cols + col
// is the element index the forward GEMM epilogue hashed)
template <typename T>
__global__ void dsilu_mul_kernel(T* __restrict__ g, const T* __restrict__ pre, size_t n, DropCfg drop) {
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (size_t)gridDim.x * blockDim.x) {
    T p = pre[x], f = T(1);
    if (drop.thr) { f = (T)drop_mul(drop, 2u, (unsigned long long)x); p *= f; }
    g[x] *= dsilu_from<T>(p, sigmoid_acc<T>(p)) * f;
  }
}

// LayerNorm backward (or identity) of the first `dim` columns of g_node_in, one warp per row:
//   g_feats[row] += d(node_norm)/dh . g_normed;  gyx[row] = g_normed * xhat  (for dL/dgamma by column sum)
template <typename T>
__global__ void ln_bwd_kernel(const T* __restrict__ h, const T* __restrict__ gamma, const T* __restrict__ g_node_in,
                              int ld_g, T* __restrict__ g_feats, T* __restrict__ gyx, int dim, int M, int do_norm) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (row >= M) return;
  const T* x = h + (size_t)row * dim;
  const T* gy = g_node_in + (size_t)row * ld_g;
  T* go = g_feats + (size_t)row * dim;
  if (!do_norm) {
    for (int c = lane; c < dim; c += 32) go[c] += gy[c];
    return;
  }
  T sm = T(0);
  for (int c = lane; c < dim; c += 32) sm += x[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sm += shfl_xor_t<T>(sm, o);
  const T mu = sm / T(dim);
  T v = T(0);
  for (int c = lane; c < dim; c += 32) { const T t = x[c] - mu; v += t * t; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += shfl_xor_t<T>(v, o);
  const T rstd = T(1) / sqrt(v / T(dim) + T(1e-5));
  T m1 = T(0), m2 = T(0);
  for (int c = lane; c < dim; c += 32) {
    const T xh = (x[c] - mu) * rstd, gg = gy[c] * gamma[c];
    m1 += gg;
    m2 += gg * xh;
    gyx[(size_t)row * dim + c] = gy[c] * xh;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { m1 += shfl_xor_t<T>(m1, o); m2 += shfl_xor_t<T>(m2, o); }
  m1 /= T(dim);
  m2 /= T(dim);
  for (int c = lane; c < dim; c += 32) {
    const T xh = (x[c] - mu) * rstd, gg = gy[c] * gamma[c];
    go[c] += rstd * (gg - m1 - xh * m2);
  }
}

// Accumulators in SimtPackLayout order -> gradients shaped like the parameters (all targets zero-filled before).
template <typename T>
__global__ void unpack_grads_kernel(Dims s, SimtPackLayout L, uint32_t flags, const T* __restrict__ gpk,
                                    const T* __restrict__ W1, const T* __restrict__ emb, EgnnLayerWeightGrads g) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int MP = L.MP, U = 4 * s.m;
  T* gW1 = static_cast<T*>((g.edge_w1));
  T* gW2 = static_cast<T*>((g.edge_w2));
  for (size_t x = t0; x < (size_t)s.m * s.H; x += stride) {
    const int o = (int)(x / s.H), c = (int)(x % s.H);
    gW2[x] = gpk[L.w2t + (size_t)c * MP + o];
  }
  for (size_t x = t0; x < (size_t)s.Q * s.H; x += stride) {
    const int q = (int)(x / s.H), c = (int)(x % s.H);
    gW1[(size_t)c * s.E + 2 * s.dim + q] = gpk[L.wq + (size_t)q * s.Hp + c];
  }
  if (s.label_dim > 0) {
    T* gemb = static_cast<T*>((g.label_emb));
    for (size_t x = t0; x < (size_t)s.label_dim * s.H; x += stride) {     // dL/dW1[c, label col a] = sum_l gTab[l,c] emb[l,a]
      const int a_ = (int)(x / s.H), c = (int)(x % s.H);
      T acc = T(0);
      for (int l = 0; l < s.num_labels; ++l) acc += gpk[L.tab + (size_t)l * s.Hp + c] * emb[(size_t)l * s.label_dim + a_];
      gW1[(size_t)c * s.E + 2 * s.dim + s.Q + a_] = acc;
    }
    if (gemb)
      for (size_t x = t0; x < (size_t)s.num_labels * s.label_dim; x += stride) {
        const int l = (int)(x / s.label_dim), a_ = (int)(x % s.label_dim);
        T acc = T(0);
        for (int c = 0; c < s.H; ++c) acc += gpk[L.tab + (size_t)l * s.Hp + c] * W1[(size_t)c * s.E + 2 * s.dim + s.Q + a_];
        gemb[x] = acc;
      }
  }
  T* gb2 = static_cast<T*>((g.edge_b2));
  for (size_t x = t0; x < (size_t)s.m; x += stride) gb2[x] = gpk[L.misc + x];
  if (flags & EGNN_FLAG_SOFT_EDGES) {
    T* ggw = static_cast<T*>((g.gate_w));
    for (size_t x = t0; x < (size_t)s.m; x += stride) ggw[x] = gpk[L.misc + MP + x];
    if (t0 == 0) static_cast<T*>((g.gate_b))[0] = gpk[L.misc + 2 * MP + 0];
  }
  if (flags & EGNN_FLAG_UPDATE_COORS) {
    T* gW3 = static_cast<T*>((g.coors_w1));
    T* gb3 = static_cast<T*>((g.coors_b1));
    T* gW4 = static_cast<T*>((g.coors_w2));
    for (size_t x = t0; x < (size_t)U * s.m; x += stride) {
      const int u = (int)(x / s.m), o = (int)(x % s.m);
      gW3[x] = gpk[L.w3 + (size_t)u * MP + o];
    }
    for (size_t x = t0; x < (size_t)U; x += stride) { gb3[x] = gpk[L.b3 + x]; gW4[x] = gpk[L.w4 + x]; }
    if (t0 == 0) static_cast<T*>((g.coors_b2))[0] = gpk[L.misc + 2 * MP + 1];
    if ((flags & EGNN_FLAG_NORM_COORS) && t0 == 0)
      static_cast<T*>((g.coors_scale))[0] = gpk[L.misc + 2 * MP + 2];
  }
}

}  // namespace egnn
