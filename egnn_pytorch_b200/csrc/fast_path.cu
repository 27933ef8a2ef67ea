// bf16 tensor-core path of the layer (EGNN_DTYPE_BF16): parameter packing, workspace layout and the
// per-layer launch sequence   node tables (tcgen05 GEMM x2) -> fused edge kernel (tc_pair.cuh)
//                             -> LayerNorm/concat -> node MLP (tcgen05 GEMM x2).
// Option sets the tensor-core kernels do not cover return EGNN_ERR_UNSUPPORTED; the binding then runs
// the fp32 SIMT kernels (never a CPU path).
#include "fast_path.h"
#include "profile.h"
#include "tc_gemm.cuh"
#include "tc_pair.cuh"
#include "tc_knn.cuh"

namespace egnn {

int knn_select_dispatch(int32_t dtype, int B, int N, int C, int k, const void* coors, const uint8_t* mask,
                        const uint8_t* adj, int adj_batched, double valid_radius, int32_t* out_idx,
                        uint8_t* out_ok, cudaStream_t st);

namespace {

struct FastDims {
  Dims s;
  int Hp;      // H rounded up to 64 (hidden chunks of the fused kernel)
  int Kn;      // dim + m rounded up to 8 (K of the first node GEMM)
};

// layout of the packed-parameter buffer (byte offsets, 256-aligned)
struct FastPack {
  size_t w1i, w1j, b1, wdh, weh, w2p, epi, wn1, bn1, wn2, bn2, lng, lnb, total;
};

FastDims fast_dims(const EgnnLayerDesc& d) {
  FastDims f;
  f.s = make_dims(d);
  f.Hp = round_up_i(f.s.H, 64);
  f.Kn = round_up_i(f.s.dim + f.s.m, 8);
  return f;
}

FastPack fast_pack_layout(const FastDims& f) {
  FastPack p;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += round_up(bytes, 256); return r; };
  const int d = f.s.dim;
  p.w1i = take((size_t)f.Hp * d * 2);
  p.w1j = take((size_t)f.Hp * d * 2);
  p.b1 = take((size_t)f.Hp * 4);
  p.wdh = take((size_t)f.Hp * 4);
  p.weh = take((size_t)TK_QE * f.Hp * 4);
  p.w2p = take((size_t)f.Hp * 32);
  p.epi = take((size_t)TP_EPI_FLOATS * 4);
  p.wn1 = take((size_t)2 * d * f.Kn * 2);
  p.bn1 = take((size_t)2 * d * 4);
  p.wn2 = take((size_t)d * 2 * d * 2);
  p.bn2 = take((size_t)d * 4);
  p.lng = take((size_t)d * 4);
  p.lnb = take((size_t)d * 4);
  p.total = o;
  return p;
}

int fast_supported(const EgnnLayerDesc& d) {
  const FastDims f = fast_dims(d);
  if (d.label_dim != 0 || d.fourier != 0) return EGNN_ERR_UNSUPPORTED;
  if (d.k == 0) {                                                  // dense all-pairs: tc_pair_kernel
    if (d.edge_dim != 0) return EGNN_ERR_UNSUPPORTED;
    if (tc_pair_smem_bytes(f.Hp) > 226 * 1024) return EGNN_ERR_UNSUPPORTED;
  } else {                                                         // neighbour lists: tc_knn_kernel
    if (d.k > 32 || d.edge_dim > TK_QE) return EGNN_ERR_UNSUPPORTED;
    if (tc_knn_smem_bytes(f.Hp, d.edge_dim > 0) > 226 * 1024) return EGNN_ERR_UNSUPPORTED;
  }
  if (d.C != 3 || d.m_dim != 16) return EGNN_ERR_UNSUPPORTED;
  if (d.dim % 8 != 0) return EGNN_ERR_UNSUPPORTED;                 // 16-byte rows for cp.async
  if (!(d.row_begin == 0 && (d.row_end == 0 || d.row_end == d.N))) return EGNN_ERR_UNSUPPORTED;
  return EGNN_OK;
}

__device__ __forceinline__ float bf(const void* p, size_t i) { return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i]); }

__global__ void fast_pack_kernel(FastDims f, FastPack L, EgnnLayerWeights w, uint32_t flags, unsigned char* out) {
  const Dims& s = f.s;
  const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int d = s.dim, H = s.H, E = s.E, Hp = f.Hp;
  __nv_bfloat16* w1i = reinterpret_cast<__nv_bfloat16*>(out + L.w1i);
  __nv_bfloat16* w1j = reinterpret_cast<__nv_bfloat16*>(out + L.w1j);
  for (size_t x = t0; x < (size_t)Hp * d; x += stride) {
    const int c = (int)(x / d), k = (int)(x % d);
    const __nv_bfloat16 z = __float2bfloat16(0.f);
    w1i[x] = c < H ? static_cast<const __nv_bfloat16*>(w.edge_w1)[(size_t)c * E + k] : z;
    w1j[x] = c < H ? static_cast<const __nv_bfloat16*>(w.edge_w1)[(size_t)c * E + d + k] : z;
  }
  float* b1 = reinterpret_cast<float*>(out + L.b1);
  float* wdh = reinterpret_cast<float*>(out + L.wdh);
  for (size_t c = t0; c < (size_t)Hp; c += stride) {
    b1[c] = c < (size_t)H ? bf(w.edge_b1, c) : 0.f;
    wdh[c] = c < (size_t)H ? 0.5f * bf(w.edge_w1, c * E + 2 * d) : 0.f;      // the d_ij column of W1, pre-halved
  }
  float* weh = reinterpret_cast<float*>(out + L.weh);                          // continuous edge columns, pre-halved
  for (size_t x = t0; x < (size_t)TK_QE * Hp; x += stride) {
    const int q = (int)(x / Hp);
    const size_t c = x % Hp;
    weh[x] = (c < (size_t)H && q < s.edge_dim) ? 0.5f * bf(w.edge_w1, c * E + 2 * d + 1 + q) : 0.f;
  }
  // W2 [16][H] -> UMMA K-major core matrices: [slab = c/16][kc = (c%16)/8][nc = n/8][r = n%8][e = c%8]
  __nv_bfloat16* w2p = reinterpret_cast<__nv_bfloat16*>(out + L.w2p);
  for (size_t x = t0; x < (size_t)Hp * 16; x += stride) {
    const int e = (int)(x & 7), r = (int)((x >> 3) & 7), nc = (int)((x >> 6) & 1), kc = (int)((x >> 7) & 1);
    const int slab = (int)(x >> 8);
    const int n = nc * 8 + r, c = slab * 16 + kc * 8 + e;
    w2p[x] = c < H ? static_cast<const __nv_bfloat16*>(w.edge_w2)[(size_t)n * H + c] : __float2bfloat16(0.f);
  }
  // epilogue constants (fp32): W3[64][16] | b3[64] | w4[64] | b2[16] | gate_w[16] | gate_b, b4, scale, 0
  float* epi = reinterpret_cast<float*>(out + L.epi);
  const bool uc = flags & EGNN_FLAG_UPDATE_COORS, soft = flags & EGNN_FLAG_SOFT_EDGES;
  for (size_t x = t0; x < (size_t)TP_EPI_FLOATS; x += stride) {
    float v = 0.f;
    if (x < 1024) v = uc ? bf(w.coors_w1, x) : 0.f;
    else if (x < 1088) v = uc ? bf(w.coors_b1, x - 1024) : 0.f;
    else if (x < 1152) v = uc ? bf(w.coors_w2, x - 1088) : 0.f;
    else if (x < 1168) v = bf(w.edge_b2, x - 1152);
    else if (x < 1184) v = soft ? bf(w.gate_w, x - 1168) : 0.f;
    else if (x == 1184) v = soft ? bf(w.gate_b, 0) : 0.f;
    else if (x == 1185) v = uc ? bf(w.coors_b2, 0) : 0.f;
    else if (x == 1186) v = (flags & EGNN_FLAG_NORM_COORS) ? bf(w.coors_scale, 0) : 1.f;
    epi[x] = v;
  }
  if (flags & EGNN_FLAG_UPDATE_FEATS) {
    __nv_bfloat16* wn1 = reinterpret_cast<__nv_bfloat16*>(out + L.wn1);
    const int Kin = d + s.m;
    for (size_t x = t0; x < (size_t)2 * d * f.Kn; x += stride) {
      const int n = (int)(x / f.Kn), k = (int)(x % f.Kn);
      wn1[x] = k < Kin ? static_cast<const __nv_bfloat16*>(w.node_w1)[(size_t)n * Kin + k] : __float2bfloat16(0.f);
    }
    __nv_bfloat16* wn2 = reinterpret_cast<__nv_bfloat16*>(out + L.wn2);
    for (size_t x = t0; x < (size_t)d * 2 * d; x += stride) wn2[x] = static_cast<const __nv_bfloat16*>(w.node_w2)[x];
    float* bn1 = reinterpret_cast<float*>(out + L.bn1);
    for (size_t x = t0; x < (size_t)2 * d; x += stride) bn1[x] = bf(w.node_b1, x);
    float* bn2 = reinterpret_cast<float*>(out + L.bn2);
    float* lng = reinterpret_cast<float*>(out + L.lng);
    float* lnb = reinterpret_cast<float*>(out + L.lnb);
    const bool nf = flags & EGNN_FLAG_NORM_FEATS;
    for (size_t x = t0; x < (size_t)d; x += stride) {
      bn2[x] = bf(w.node_b2, x);
      lng[x] = nf ? bf(w.norm_g, x) : 1.f;
      lnb[x] = nf ? bf(w.norm_b, x) : 0.f;
    }
  }
}

// node_in[row, 0:dim] = LayerNorm(h[row]) | h[row] (bf16), pad columns [dim+m, Kn) = 0; one warp per row.
__global__ void ln_concat_bf16_kernel(const __nv_bfloat16* __restrict__ h, const float* __restrict__ g,
                                      const float* __restrict__ bta, __nv_bfloat16* __restrict__ node_in, int Kn, int dim,
                                      int m, int M, int do_norm) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (row >= M) return;
  const __nv_bfloat16* x = h + (size_t)row * dim;
  __nv_bfloat16* y = node_in + (size_t)row * Kn;
  for (int c = dim + m + lane; c < Kn; c += 32) y[c] = __float2bfloat16(0.f);
  if (!do_norm) {
    for (int c = lane; c < dim; c += 32) y[c] = x[c];
    return;
  }
  float s = 0.f;
  for (int c = lane; c < dim; c += 32) s += __bfloat162float(x[c]);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s / dim;
  float v = 0.f;
  for (int c = lane; c < dim; c += 32) { const float t = __bfloat162float(x[c]) - mu; v += t * t; }
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float rstd = rsqrtf(v / dim + 1e-5f);
  for (int c = lane; c < dim; c += 32) y[c] = __float2bfloat16((__bfloat162float(x[c]) - mu) * rstd * g[c] + bta[c]);
}

struct FastWs { size_t Atab, Btab, node_in, h1, nbr_idx, nbr_ok, total; };
FastWs fast_ws_layout(const FastDims& f, uint32_t flags) {
  FastWs w;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += round_up(bytes, 256); return r; };
  const bool uf = flags & EGNN_FLAG_UPDATE_FEATS;
  w.Atab = take((size_t)f.s.M * f.Hp * 4);
  w.Btab = take((size_t)f.s.M * f.Hp * 2);
  w.node_in = take(uf ? (size_t)f.s.M * f.Kn * 2 : 0);
  w.h1 = take(uf ? (size_t)f.s.M * 2 * f.s.dim * 2 : 0);
  w.nbr_idx = take((size_t)f.s.M * f.s.k * sizeof(int32_t));
  w.nbr_ok = take((size_t)f.s.M * f.s.k);
  w.total = o;
  return w;
}

int launch_tc_gemm(const TcGemmArgs& g, cudaStream_t st) {
  static bool attr_set[64] = {false};
  int dev = 0;
  EGNN_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    EGNN_CUDA_TRY(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    attr_set[dev] = true;
  }
  dim3 grid(ceil_div(g.Nout, GEMM_BN), ceil_div(g.M, GEMM_BM));
  tc_gemm_kernel<<<grid, 128, GEMM_SMEM_BYTES, st>>>(g);
  EGNN_LAUNCH_CHECK();
  count_launch();
  return EGNN_OK;
}

}  // namespace

int debug_gemm_bf16(int M, int N, int K, const void* A, const void* W, const float* bias, float scale, int act,
                    void* out, int out_f32, cudaStream_t st) {
  if (!A || !W || !out) return EGNN_ERR_NULL;
  if (M <= 0 || N <= 0 || K <= 0 || K % 8 != 0 || N % 8 != 0) return EGNN_ERR_SHAPE;
  TcGemmArgs g{};
  g.A = static_cast<const __nv_bfloat16*>(A); g.lda = K; g.W = static_cast<const __nv_bfloat16*>(W); g.ldw = K;
  g.bias = bias; g.R = nullptr; g.ldr = 0; g.out = out; g.ldo = N; g.out_f32 = out_f32;
  g.M = M; g.Nv = N; g.Nout = N; g.K = K; g.scale = scale; g.act = act;
  return launch_tc_gemm(g, st);
}

int fast_packed_bytes(const EgnnLayerDesc& d, size_t* out) {
  EGNN_TRY(fast_supported(d));
  *out = fast_pack_layout(fast_dims(d)).total;
  return EGNN_OK;
}

int fast_pack_weights(const EgnnLayerDesc& d, const EgnnLayerWeights& w, void* packed, size_t bytes, cudaStream_t st) {
  EGNN_TRY(fast_supported(d));
  const FastDims f = fast_dims(d);
  const FastPack L = fast_pack_layout(f);
  if (bytes < L.total) return EGNN_ERR_WORKSPACE;
  fast_pack_kernel<<<296, 256, 0, st>>>(f, L, w, d.flags, static_cast<unsigned char*>(packed));
  EGNN_LAUNCH_CHECK();
  return EGNN_OK;
}

int fast_workspace_bytes(const EgnnLayerDesc& d, size_t* out) {
  EGNN_TRY(fast_supported(d));
  *out = fast_ws_layout(fast_dims(d), d.flags).total + 256;
  return EGNN_OK;
}

int fast_forward(const EgnnLayerDesc& d, const EgnnLayerWeights& w, const void* packed, const EgnnLayerIO& io,
                 void* ws, size_t ws_bytes, cudaStream_t st) {
  (void)w;
  EGNN_TRY(fast_supported(d));
  const FastDims f = fast_dims(d);
  const Dims& s = f.s;
  const FastPack L = fast_pack_layout(f);
  const FastWs wl = fast_ws_layout(f, d.flags);
  if (ws_bytes < wl.total) return EGNN_ERR_WORKSPACE;
  const unsigned char* pk = static_cast<const unsigned char*>(packed);
  unsigned char* base = static_cast<unsigned char*>(ws);
  float* Atab = reinterpret_cast<float*>(base + wl.Atab);
  __nv_bfloat16* Btab = reinterpret_cast<__nv_bfloat16*>(base + wl.Btab);
  __nv_bfloat16* node_in = reinterpret_cast<__nv_bfloat16*>(base + wl.node_in);
  __nv_bfloat16* h1 = reinterpret_cast<__nv_bfloat16*>(base + wl.h1);
  const __nv_bfloat16* feats = static_cast<const __nv_bfloat16*>(io.feats);
  const bool uf = d.flags & EGNN_FLAG_UPDATE_FEATS, uc = d.flags & EGNN_FLAG_UPDATE_COORS;

  {  // per-node tables, pre-halved for the tanh form of SiLU:  A' = 0.5 (h W1_i^T + b1),  B' = 0.5 h W1_j^T
    StageTimer tm(st, STAGE_NODE_PRE);
    TcGemmArgs g{};
    g.A = feats; g.lda = s.dim; g.K = s.dim; g.M = s.M; g.Nv = f.Hp; g.Nout = f.Hp; g.scale = 0.5f; g.act = 0;
    g.W = reinterpret_cast<const __nv_bfloat16*>(pk + L.w1i); g.ldw = s.dim;
    g.bias = reinterpret_cast<const float*>(pk + L.b1); g.R = nullptr; g.ldr = 0;
    g.out = Atab; g.ldo = f.Hp; g.out_f32 = 1;
    EGNN_TRY(launch_tc_gemm(g, st));
    g.W = reinterpret_cast<const __nv_bfloat16*>(pk + L.w1j); g.bias = nullptr;
    g.out = Btab; g.out_f32 = 0;
    EGNN_TRY(launch_tc_gemm(g, st));
  }
  if (s.k == 0) {  // fused edge kernel, dense all-pairs
    StageTimer tm(st, STAGE_PAIR);
    TcPairArgs a{};
    a.B = s.B; a.N = s.N; a.Hp = f.Hp; a.ldn = f.Kn; a.dim = s.dim;
    a.flags = d.flags; a.has_mask = io.mask != nullptr; a.clamp = (float)d.clamp;
    a.Atab = Atab; a.Btab = Btab;
    a.wdh = reinterpret_cast<const float*>(pk + L.wdh);
    a.w2p = reinterpret_cast<const __nv_bfloat16*>(pk + L.w2p);
    a.epi = reinterpret_cast<const float*>(pk + L.epi);
    a.coors = static_cast<const float*>(io.coors);
    a.mask = io.mask;
    a.m_out = uf ? node_in + s.dim : nullptr;
    a.coors_out = uc ? static_cast<float*>(io.coors_out) : nullptr;
    const size_t smem = tc_pair_smem_bytes(f.Hp);
    static size_t smem_set[64] = {0};
    int dev = 0;
    EGNN_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 64 && smem_set[dev] < smem) {
      EGNN_CUDA_TRY(cudaFuncSetAttribute(tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smem_set[dev] = smem;
    }
    dim3 grid(ceil_div(s.N, TP_TI), s.B);
    tc_pair_kernel<<<grid, TP_THREADS, smem, st>>>(a);
    EGNN_LAUNCH_CHECK();
    count_launch();
  } else {         // neighbour lists: distance + top-k select, then the gathered fused edge kernel
    int32_t* nbr_idx = reinterpret_cast<int32_t*>(base + wl.nbr_idx);
    uint8_t* nbr_ok = base + wl.nbr_ok;
    if (io.nbr_idx) {                                  // edge-list mode: the caller's lists, no ranking
      nbr_idx = const_cast<int32_t*>(io.nbr_idx);
      nbr_ok = nullptr;
    } else {
      StageTimer tm(st, STAGE_SELECT);
      const double vr = (d.flags & EGNN_FLAG_ONLY_SPARSE) ? 0.0 : d.valid_radius;
      EGNN_TRY(knn_select_dispatch(EGNN_DTYPE_F32, s.B, s.N, 3, s.k, io.coors, io.mask, io.adj,
                                   (d.flags & EGNN_FLAG_ADJ_BATCHED) ? 1 : 0, vr, nbr_idx, nbr_ok, st));
      count_launch();
    }
    StageTimer tm(st, STAGE_PAIR);
    TcKnnArgs a{};
    a.B = s.B; a.N = s.N; a.Hp = f.Hp; a.ldn = f.Kn; a.dim = s.dim; a.k = s.k; a.edge_dim = s.edge_dim;
    a.flags = d.flags; a.has_mask = io.mask != nullptr; a.clamp = (float)d.clamp;
    a.Atab = Atab; a.Btab = Btab;
    a.wdh = reinterpret_cast<const float*>(pk + L.wdh);
    a.weh = reinterpret_cast<const float*>(pk + L.weh);
    a.w2p = reinterpret_cast<const __nv_bfloat16*>(pk + L.w2p);
    a.epi = reinterpret_cast<const float*>(pk + L.epi);
    a.coors = static_cast<const float*>(io.coors);
    a.edges = static_cast<const __nv_bfloat16*>(io.edges);
    a.mask = io.mask;
    a.nbr_idx = nbr_idx; a.nbr_ok = nbr_ok;
    a.m_out = uf ? node_in + s.dim : nullptr;
    a.coors_out = uc ? static_cast<float*>(io.coors_out) : nullptr;
    const bool ed = s.edge_dim > 0;
    const size_t smem = tc_knn_smem_bytes(f.Hp, ed);
    static size_t smem_set[2][64] = {{0}};
    int dev = 0;
    EGNN_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 64 && smem_set[ed][dev] < smem) {
      if (ed) EGNN_CUDA_TRY(cudaFuncSetAttribute(tc_knn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      else EGNN_CUDA_TRY(cudaFuncSetAttribute(tc_knn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smem_set[ed][dev] = smem;
    }
    dim3 grid(ceil_div(s.N, TK_ROWS), s.B);
    if (ed) tc_knn_kernel<true><<<grid, TK_THREADS, smem, st>>>(a);
    else tc_knn_kernel<false><<<grid, TK_THREADS, smem, st>>>(a);
    EGNN_LAUNCH_CHECK();
    count_launch();
  }
  StageTimer post(st, STAGE_NODE_POST);
  if (uf) {  // h' = node_mlp([LN(h) | m_i]) + h
    ln_concat_bf16_kernel<<<ceil_div(s.M * 32, 256), 256, 0, st>>>(
        feats, reinterpret_cast<const float*>(pk + L.lng), reinterpret_cast<const float*>(pk + L.lnb), node_in, f.Kn,
        s.dim, s.m, s.M, (d.flags & EGNN_FLAG_NORM_FEATS) ? 1 : 0);
    EGNN_LAUNCH_CHECK();
    count_launch();
    TcGemmArgs g{};
    g.A = node_in; g.lda = f.Kn; g.K = f.Kn; g.M = s.M; g.Nv = 2 * s.dim; g.Nout = 2 * s.dim; g.scale = 1.f; g.act = 1;
    g.W = reinterpret_cast<const __nv_bfloat16*>(pk + L.wn1); g.ldw = f.Kn;
    g.bias = reinterpret_cast<const float*>(pk + L.bn1);
    g.out = h1; g.ldo = 2 * s.dim; g.out_f32 = 0;
    EGNN_TRY(launch_tc_gemm(g, st));
    g.A = h1; g.lda = 2 * s.dim; g.K = 2 * s.dim; g.Nv = s.dim; g.Nout = s.dim; g.act = 0;
    g.W = reinterpret_cast<const __nv_bfloat16*>(pk + L.wn2); g.ldw = 2 * s.dim;
    g.bias = reinterpret_cast<const float*>(pk + L.bn2);
    g.R = feats; g.ldr = s.dim;
    g.out = io.feats_out; g.ldo = s.dim; g.out_f32 = 0;
    EGNN_TRY(launch_tc_gemm(g, st));
  } else if (io.feats_out != io.feats) {
    EGNN_CUDA_TRY(cudaMemcpyAsync(io.feats_out, io.feats, (size_t)s.M * s.dim * 2, cudaMemcpyDeviceToDevice, st));
  }
  if (!uc && io.coors_out != io.coors)
    EGNN_CUDA_TRY(cudaMemcpyAsync(io.coors_out, io.coors, (size_t)s.M * 3 * 4, cudaMemcpyDeviceToDevice, st));
  return EGNN_OK;
}

}  // namespace egnn
