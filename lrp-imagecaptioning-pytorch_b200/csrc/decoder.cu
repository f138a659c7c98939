// Decoder relevance (fp32), batched over Q = (image, target word) requests.
//
//   gridTD : ExplainGridTDAttention.explain_caption_wordt   models/gridTDmodel.py:1014-1135
//   AoA    : ExplainAOAAttention.explain_caption_wordt      models/aoamodel.py:1064-1156 (+ lrp_mha :812-862)
//   tune   : get_lrp_weight_step                            models/gridTDmodel.py:549-578, aoamodel.py:597-626
//
// The reference walks each vector with lrp_linear_eps (gridTDmodel.py:744-765), materialising
// `weight * input` (out x in) and `eye(H)` on every call.  Here the same arithmetic is in closed form:
//   ident(r,x,z) = x * r / stab(z)                     (weight = eye)
//   lin(r,x,z,W) = x * ((r / stab(z)) @ W)              (GEMM over all requests at once)
// The LSTM chain is sequential in the step index i (t..0) but independent across requests, so every
// step is: one warp-friendly element-wise kernel + one (Q x H) @ (H x in) GEMM + one element-wise kernel.
#include "lrpx_common.cuh"

namespace lrpx {

// ------------------------------------------------------------------------------------------------
// SGEMM  C[M,N] = A[M,K] @ B[K,N]  (row-major), 64x64x16 tiles, 4x4 per thread, fused epilogues.
// ------------------------------------------------------------------------------------------------
enum { GE_STORE = 0, GE_FEAT = 1, GE_AOA_PROJ = 2 };

struct GemmEpi {
  // GE_FEAT: out[m][n] = feat[b][p][n] * (acc + add_q[q][n])      rows m = q*P + p
  // GE_AOA_PROJ: out[m][n] = (A[b][p][n] * (acc + add_q[q][n])) / stab(A_pre[b][p][n])
  const float* x0;       // feat / A          (B,P,N)
  const float* x1;       // A_pre             (B,P,N)
  const float* add_q;    // (Q,N) or null
  const int32_t* req_img;
  int P;
};

template <int EPI>
__global__ void __launch_bounds__(256) sgemm_nn_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                       float* __restrict__ C, int M, int N, int K, int lda, int ldb,
                                                       int ldc, GemmEpi e) {
  constexpr int TBM = 64, TBN = 64, TBK = 16;
  __shared__ float As[TBK][TBM + 4];
  __shared__ float Bs[TBK][TBN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * TBN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TBK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {           // A tile: 64 x 16
      int idx = tid + 256 * j;
      int r = idx / TBK, c = idx % TBK;
      int m = m0 + r, k = k0 + c;
      As[c][r] = (m < M && k < K) ? A[(size_t)m * lda + k] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {           // B tile: 16 x 64
      int idx = tid + 256 * j;
      int r = idx / TBN, c = idx % TBN;
      int k = k0 + r, n = n0 + c;
      Bs[r][c] = (k < K && n < N) ? Bm[(size_t)k * ldb + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TBK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    int q = 0, pp = 0, b = 0;
    if (EPI != GE_STORE) {
      q = m / e.P;
      pp = m % e.P;
      b = e.req_img[q];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx + 16 * j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (EPI == GE_FEAT) {
        float add = e.add_q ? e.add_q[(size_t)q * N + n] : 0.f;
        v = e.x0[((size_t)b * e.P + pp) * N + n] * (v + add);
      } else if (EPI == GE_AOA_PROJ) {
        size_t o = ((size_t)b * e.P + pp) * N + n;
        float add = e.add_q ? e.add_q[(size_t)q * N + n] : 0.f;
        v = (e.x0[o] * (v + add)) / stab(e.x1[o]);
      }
      C[(size_t)m * ldc + n] = v;
    }
  }
}

template <int EPI>
static int sgemm(const float* A, const float* B, float* C, int M, int N, int K, const GemmEpi& e, cudaStream_t st) {
  if (M == 0) return LRPX_OK;
  dim3 grid(ceil_div(N, 64), ceil_div(M, 64));
  sgemm_nn_kernel<EPI><<<grid, 256, 0, st>>>(A, B, C, M, N, K, K, N, N, e);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error("sgemm launch failed: %s", cudaGetErrorString(err));
    return LRPX_E_CUDA;
  }
  return LRPX_OK;
}

// ------------------------------------------------------------------------------------------------
// Error-compensated tensor-core GEMM (LRPX_DEC_TC_GEMM):  x = hi + lo with hi = bf16(x), lo = bf16(x - hi);
//   a*w ~= a_hi*w_hi + a_hi*w_lo + a_lo*w_hi   (the dropped lo*lo term is ~2^-16 relative)
// evaluated as ONE bf16 GEMM with the K dimension concatenated three times, fp32 accumulation in TMEM:
//   A' = [a_hi | a_hi | a_lo]  (M x 3K),   W' = [w_hi | w_lo | w_hi]  (N x 3K, K-major = the transposed weight)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16(x);
  lo = __float2bfloat16(x - __bfloat162float(hi));
}
// a producer kernel writes its GEMM operand both as fp32 (CUDA-core GEMM) and, when `a3` is given, directly as the
// split row [hi | hi | lo] of the tensor-core GEMM (saves the separate split pass over the operand)
__device__ __forceinline__ void put_operand(float* u, __nv_bfloat16* a3, size_t row, int K, int k, float x) {
  u[row * K + k] = x;
  if (a3) {
    __nv_bfloat16 hi, lo;
    split_bf16(x, hi, lo);
    __nv_bfloat16* o = a3 + row * 2 * K;
    o[k] = hi; o[K + k] = lo;
  }
}
// rows x K fp32 (row pitch lda) -> rows x 2K bf16 [hi | lo]; the GEMM reads a row as K = [hi | lo | hi] (a_phys wrap of
// lrpx_tc_conv: the third block group re-reads the first), so hi is stored once: 2/3 of the bytes of [hi | hi | lo]
__global__ void split3_act_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long rows, int K,
                                  int lda) {
  long long total = rows * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i / K;
    int k = (int)(i - r * K);
    __nv_bfloat16 hi, lo;
    split_bf16(x[r * lda + k], hi, lo);
    __nv_bfloat16* o = out + r * 2 * K;
    o[k] = hi; o[K + k] = lo;
  }
}
// W (K x N fp32 row-major, i.e. [k][n]) -> N x 3K bf16 [hi | hi | lo] of W^T:  a*w ~ a_hi*w_hi + a_lo*w_hi + a_hi*w_lo
__global__ void split3_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int K, int N) {
  long long total = (long long)K * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int n = (int)(i % N);
    int k = (int)(i / N);
    __nv_bfloat16 hi, lo;
    split_bf16(w[i], hi, lo);
    __nv_bfloat16* o = out + (size_t)n * 3 * K;
    o[k] = hi; o[K + k] = hi; o[2 * K + k] = lo;
  }
}
// GE_FEAT / GE_AOA_PROJ epilogues applied to a plain GEMM result in place
template <int EPI>
__global__ void gemm_epilogue_kernel(float* __restrict__ C, long long M, int N, GemmEpi e) {
  long long total = M * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long m = i / N;
    int n = (int)(i - m * N);
    int q = (int)(m / e.P), pp = (int)(m % e.P);
    int b = e.req_img[q];
    size_t o = ((size_t)b * e.P + pp) * N + n;
    float add = e.add_q ? e.add_q[(size_t)q * N + n] : 0.f;
    float v = e.x0[o] * (C[i] + add);
    if (EPI == GE_AOA_PROJ) v = v / stab(e.x1[o]);
    C[i] = v;
  }
}

static inline int ew_grid(long long total) {
  long long g = (total + 255) / 256, cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
static bool tc_shape_ok(int N, int K) { return K % 64 == 0 && N % 32 == 0 && (N <= 256 || N % 256 == 0); }

// C[M,N] = A[M,K] @ W[K,N] (+ epilogue): tensor cores when `wt3` (prepared W') is given, CUDA cores otherwise
// a3_ready: the producer kernel already wrote the split operand [hi | lo] into a3 (no fp32 A exists then)
template <int EPI>
static int gemm_any(const float* A, const float* W, const __nv_bfloat16* wt3, __nv_bfloat16* a3, float* C, int M, int N,
                    int K, const GemmEpi& e, cudaStream_t st, bool a3_ready = false) {
  if (M == 0) return LRPX_OK;
  if (!wt3) return sgemm<EPI>(A, W, C, M, N, K, e, st);
  if (!a3_ready) split3_act_kernel<<<ew_grid((long long)M * K), 256, 0, st>>>(A, a3, M, K, K);
  lrpx_tc_conv_args g{};
  g.cin = 3 * K; g.a_phys = 2 * K; g.ncol = N; g.ksize = 1;
  g.a = a3; g.wt = wt3; g.out = C;
  if (EPI == GE_STORE) {          // one PF "block" of M rows: the STORE_F32 epilogue writes every in-range row
    g.n_img = 1; g.h = 0; g.w = M - 1;
    g.epilogue = LRPX_TC_EPI_STORE_F32;
    return lrpx_tc_conv(&g, st);
  }
  // projector rules fused into the GEMM's epilogue: rows are (request, pixel) = one PF "block" of P rows per request
  g.n_img = M / e.P; g.h = 0; g.w = e.P - 1;
  g.epilogue = EPI == GE_FEAT ? LRPX_TC_EPI_FEAT : LRPX_TC_EPI_FEAT_DIV;
  g.x = e.x0; g.x1 = e.x1; g.bias = e.add_q; g.row_img = e.req_img;
  return lrpx_tc_conv(&g, st);
}
static __nv_bfloat16* prep_weight3(const float* W, __nv_bfloat16* dst, int K, int N, cudaStream_t st) {
  split3_weight_kernel<<<ew_grid((long long)K * N), 256, 0, st>>>(W, dst, K, N);
  return dst;
}

// fc rule for the target word only (one-hot relevance): gridTDmodel.py:1033-1059
//   r_sum = s_in * W_fc[word] * logit/stab(logit);  r_a = a * r_sum / stab(s_in), r_b likewise
__device__ __forceinline__ void fc_split(float a, float b, float wrow, float coef, float& ra, float& rb) {
  float s_in = a + b;
  float r_sum = s_in * wrow * coef;
  float d = stab(s_in);
  ra = a * r_sum / d;
  rb = b * r_sum / d;
}

// ------------------------------------------------------------------------------------------------
// gridTD
// ------------------------------------------------------------------------------------------------
struct GridWs {
  float *r_h2, *r_c2, *r_c1, *r_cth, *r_glob, *u, *v, *uctx, *coefavg, *wproj;
  // LRPX_DEC_TC_GEMM: split activations (largest GEMM) and prepared weights, bf16
  __nv_bfloat16 *a3, *w3_g2, *w3_g1, *w3_glob, *w3_proj;
  __nv_bfloat16* a3u;      // split form of u written by the step kernels themselves (null: CUDA-core GEMMs)
};

__global__ void grid_init_kernel(lrpx_gridtd_args a, GridWs w) {
  int q = blockIdx.x;
  int b = a.req_img[q], t = a.req_t[q], word = a.req_word[q];
  float logit = a.pred[((size_t)b * a.T + t) * a.V + word];
  float coef = logit / stab(logit);
  const float* h2 = a.h2 + ((size_t)b * (a.T + 1) + t + 1) * a.H;
  const float* ch = a.ctx_hat + ((size_t)b * a.T + t) * a.H;
  const float* wr = a.W_fc + (size_t)word * a.H;
  for (int j = threadIdx.x; j < a.H; j += blockDim.x) {
    float rh, rc;
    fc_split(h2[j], ch[j], wr[j], coef, rh, rc);
    size_t o = (size_t)q * a.H + j;
    w.r_h2[o] = rh;
    w.r_cth[o] = rc;
    w.r_c2[o] = 0.f;
    w.r_c1[o] = 0.f;
  }
  for (int j = threadIdx.x; j < a.E; j += blockDim.x) w.r_glob[(size_t)q * a.E + j] = 0.f;
  for (int j = threadIdx.x; j < a.T; j += blockDim.x) {
    a.r_words[(size_t)q * a.T + j] = 0.f;
    if (a.r_words_raw) a.r_words_raw[(size_t)q * a.T + j] = 0.f;
  }
}

// LanguageLSTM cell rule (:1061-1069) -> u = r_g2 / stab(g2)
__global__ void grid_cell2_kernel(lrpx_gridtd_args a, GridWs w, int i) {
  int q = blockIdx.x;
  int b = a.req_img[q], t = a.req_t[q];
  bool active = i <= t;
  size_t bi = ((size_t)b * a.T + i) * a.H, bi1 = ((size_t)b * (a.T + 1) + i + 1) * a.H,
         bi0 = ((size_t)b * (a.T + 1) + i) * a.H;
  for (int j = threadIdx.x; j < a.H; j += blockDim.x) {
    size_t o = (size_t)q * a.H + j;
    if (!active) { put_operand(w.u, w.a3u, q, a.H, j, 0.f); continue; }
    float rc2 = w.r_c2[o] + w.r_h2[o];
    float d = stab(a.c2[bi1 + j]);
    float g = a.g2[bi + j];
    float r_g = a.i2[bi + j] * tanhf(g) * rc2 / d;
    w.r_c2[o] = a.f2[bi + j] * a.c2[bi0 + j] * rc2 / d;
    put_operand(w.u, w.a3u, q, a.H, j, r_g / stab(g));
  }
}

// after v = u @ W_g2 : slices of xh2 (:1070-1084), attention split, AdaLSTM cell rule (:1096-1105)
__global__ void grid_post2_kernel(lrpx_gridtd_args a, GridWs w, int i) {
  int q = blockIdx.x;
  int b = a.req_img[q], t = a.req_t[q];
  bool active = i <= t;
  const int H = a.H;
  size_t bi = ((size_t)b * a.T + i) * H, bi1 = ((size_t)b * (a.T + 1) + i + 1) * H,
         bi0 = ((size_t)b * (a.T + 1) + i) * H;
  const float* x2 = a.x2 + ((size_t)b * a.T + i) * 2 * H;
  const float* vq = w.v + (size_t)q * 3 * H;
  float beta = a.beta[(size_t)b * a.T + i];
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    size_t o = (size_t)q * H + j;
    if (!active) { put_operand(w.u, w.a3u, q, H, j, 0.f); continue; }
    float rx_ctx = x2[j] * vq[j];                 // xh2[:H]   = ctx_hat_i
    float rx_h1 = x2[H + j] * vq[H + j];          // xh2[H:2H] = h1_{i+1}
    float rx_h2 = a.h2[bi0 + j] * vq[2 * H + j];  // xh2[2H:]  = h2_i
    float rcth = (i == t ? w.r_cth[o] : 0.f) + rx_ctx;
    float cth = a.ctx_hat[bi + j];
    float dct = stab(cth);
    float r_s = beta * a.st[bi + j] * rcth / dct;
    float cx = a.ctx[bi + j];
    float r_ctx = cx * (1.f - beta) * rcth / dct;
    w.uctx[((size_t)q * a.T + i) * H + j] = r_ctx / stab(cx);
    float rc1 = w.r_c1[o] + r_s + rx_h1;
    float d = stab(a.c1[bi1 + j]);
    float g = a.g1[bi + j];
    float r_g = a.i1[bi + j] * tanhf(g) * rc1 / d;
    w.r_c1[o] = a.f1[bi + j] * a.c1[bi0 + j] * rc1 / d;
    put_operand(w.u, w.a3u, q, H, j, r_g / stab(g));
    w.r_h2[o] = rx_h2;
  }
}

// after v = u @ W_g1 : slices of xh1 (:1106-1115)
__global__ void grid_post1_kernel(lrpx_gridtd_args a, GridWs w, int i) {
  int q = blockIdx.x;
  int b = a.req_img[q], t = a.req_t[q];
  if (i > t) return;
  const int H = a.H, E = a.E;
  const float* x1 = a.x1 + ((size_t)b * a.T + i) * (H + 2 * E);
  const float* vq = w.v + (size_t)q * (2 * H + 2 * E);
  float wsum = 0.f;
  for (int k = threadIdx.x; k < H + 2 * E; k += blockDim.x) {
    float rx = x1[k] * vq[k];
    if (k < H) w.r_h2[(size_t)q * H + k] += rx;
    else if (k < H + E) w.r_glob[(size_t)q * E + (k - H)] += rx;
    else wsum += rx;
  }
  // r_h1[i] = rx1[H+2E:] is dead: it is overwritten at step i-1 before being read (:1075 vs :1110)
  for (int o = 16; o; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  __shared__ float sm[32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sm[wid] = wsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < (blockDim.x + 31) / 32; ++k) s += sm[k];
    float* dst = a.r_words_raw ? a.r_words_raw : a.r_words;
    dst[(size_t)q * a.T + i] = s;
    if (a.r_words_raw) a.r_words[(size_t)q * a.T + i] = s;
  }
}

// u_g = r_glob / stab(glob_pre)   (:1116-1119)
__global__ void grid_glob_kernel(lrpx_gridtd_args a, GridWs w) {
  int q = blockIdx.x, b = a.req_img[q];
  for (int j = threadIdx.x; j < a.E; j += blockDim.x)
    w.u[(size_t)q * a.E + j] = w.r_glob[(size_t)q * a.E + j] / stab(a.glob_pre[(size_t)b * a.E + j]);
}
// coefavg = avg * v / stab(avg) / P  (mean-pool rule, :1121-1124; r_avg = avg * v)
__global__ void grid_avg_kernel(lrpx_gridtd_args a, GridWs w) {
  int q = blockIdx.x, b = a.req_img[q];
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float av = a.avg[(size_t)b * a.C + c];
    float r_avg = av * w.v[(size_t)q * a.C + c];
    w.coefavg[(size_t)q * a.C + c] = r_avg / stab(av) / (float)a.P;
  }
}
// wproj[q][p][h] = A * (sum_i alpha_i[p] * uctx_i[h]) / stab(A_pre)     (:1091-1095 then :1125-1128)
__global__ void grid_attn_kernel(lrpx_gridtd_args a, GridWs w) {
  int q = blockIdx.y, p = blockIdx.x;
  int b = a.req_img[q], t = a.req_t[q];
  const float* al = a.alpha + (size_t)b * a.T * a.P + p;
  size_t o0 = ((size_t)b * a.P + p) * a.H;
  for (int h = threadIdx.x; h < a.H; h += blockDim.x) {
    float acc = 0.f;
    for (int i = t; i >= 0; --i) acc += al[(size_t)i * a.P] * w.uctx[((size_t)q * a.T + i) * a.H + h] * a.A[o0 + h];
    w.wproj[((size_t)q * a.P + p) * a.H + h] = acc / stab(a.A_pre[o0 + h]);
  }
}
// Same rule, one block per request and one thread per hidden unit.  The request's alpha rows ((t+1) x P) and uctx
// rows ((t+1) x H) sit in shared memory; a thread walks the pixels four at a time: per step i one LDS (uctx) + one
// LDS.128 (four alphas, broadcast) feed four FMAs, i = t..0 like the reference loop, then
// wproj = A * acc / stab(A_pre).  (The per-(request, pixel) form above re-reads uctx 196 times and spends ~4
// instructions per multiply-add: 1.4 ms per 1216 requests; this form is bound by its 0.7 GB of output.)
// SPLIT: the result leaves directly as the split bf16 operand [hi | lo] of the tensor-core projector GEMM.
template <bool SPLIT>
__global__ void __launch_bounds__(512) grid_attn_rows_kernel(lrpx_gridtd_args a, GridWs w) {
  extern __shared__ __align__(16) float att_s[];          // alpha[(t+1)][P4] | uctx[(t+1)][H]
  const int q = blockIdx.x;
  const int b = a.req_img[q], t = a.req_t[q];
  const int H = a.H, P = a.P, P4 = (P + 3) & ~3;
  float* al_s = att_s;
  float* u_s = att_s + (size_t)(t + 1) * P4;
  for (int k = threadIdx.x; k < (t + 1) * P4; k += blockDim.x) {
    const int i = k / P4, p = k - i * P4;
    al_s[k] = p < P ? a.alpha[((size_t)b * a.T + i) * P + p] : 0.f;
  }
  for (int k = threadIdx.x; k < (t + 1) * H; k += blockDim.x) u_s[k] = w.uctx[(size_t)q * a.T * H + k];
  __syncthreads();
  if ((H & 3) == 0) {
    // four hidden units per thread: per step one LDS.128 of uctx and one of the alphas feed 16 FMAs, A / A_pre come in
    // 16-byte loads and the results leave as 8- or 16-byte stores (one unit per thread meant 2-byte stores of the split
    // operand: 0.26 of the copy bandwidth in profiles/r1_hbm_kernels.md)
    for (int h4 = threadIdx.x * 4; h4 < H; h4 += blockDim.x * 4) {
      for (int p0 = 0; p0 < P; p0 += 4) {
        float4 Av[4], Ap[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool ok = p0 + k < P;
          const size_t o0 = ((size_t)b * P + (ok ? p0 + k : 0)) * H + h4;
          Av[k] = ok ? __ldg(reinterpret_cast<const float4*>(a.A + o0)) : make_float4(0.f, 0.f, 0.f, 0.f);
          Ap[k] = ok ? __ldg(reinterpret_cast<const float4*>(a.A_pre + o0)) : make_float4(1.f, 1.f, 1.f, 1.f);
        }
        float acc[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[k][j] = 0.f;
        for (int i = t; i >= 0; --i) {
          const float4 uv = *reinterpret_cast<const float4*>(u_s + i * H + h4);
          const float4 a4 = *reinterpret_cast<const float4*>(al_s + i * P4 + p0);
          const float al[4] = {a4.x, a4.y, a4.z, a4.w}, uu[4] = {uv.x, uv.y, uv.z, uv.w};
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[k][j] = fmaf(al[k], uu[j], acc[k][j]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (p0 + k >= P) break;
          const size_t row = (size_t)q * P + p0 + k;
          const float av[4] = {Av[k].x, Av[k].y, Av[k].z, Av[k].w}, ap[4] = {Ap[k].x, Ap[k].y, Ap[k].z, Ap[k].w};
          if (SPLIT) {
            uint32_t hi2[2], lo2[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              __nv_bfloat16 h0, l0, h1, l1;
              split_bf16(__fdividef(acc[k][2 * j] * av[2 * j], stab(ap[2 * j])), h0, l0);
              split_bf16(__fdividef(acc[k][2 * j + 1] * av[2 * j + 1], stab(ap[2 * j + 1])), h1, l1);
              hi2[j] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
              lo2[j] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
            __nv_bfloat16* o = w.a3 + row * 2 * H + h4;
            *reinterpret_cast<uint2*>(o) = make_uint2(hi2[0], hi2[1]);
            *reinterpret_cast<uint2*>(o + H) = make_uint2(lo2[0], lo2[1]);
          } else {
            *reinterpret_cast<float4*>(w.wproj + row * H + h4) =
                make_float4(acc[k][0] * av[0] / stab(ap[0]), acc[k][1] * av[1] / stab(ap[1]),
                            acc[k][2] * av[2] / stab(ap[2]), acc[k][3] * av[3] / stab(ap[3]));
          }
        }
      }
    }
    return;
  }
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    for (int p0 = 0; p0 < P; p0 += 4) {
      float Av[4], Ap[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool ok = p0 + k < P;
        const size_t o0 = ((size_t)b * P + (ok ? p0 + k : 0)) * H + h;
        Av[k] = ok ? __ldg(a.A + o0) : 0.f;
        Ap[k] = ok ? __ldg(a.A_pre + o0) : 1.f;
      }
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i = t; i >= 0; --i) {
        const float uv = u_s[i * H + h];
        const float4 a4 = *reinterpret_cast<const float4*>(al_s + i * P4 + p0);
        acc[0] = fmaf(a4.x, uv, acc[0]);
        acc[1] = fmaf(a4.y, uv, acc[1]);
        acc[2] = fmaf(a4.z, uv, acc[2]);
        acc[3] = fmaf(a4.w, uv, acc[3]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (p0 + k >= P) break;
        const size_t row = (size_t)q * P + p0 + k;
        if (SPLIT) {
          // the quotient is split into bf16 hi + lo (16 mantissa bits): the 2-ulp division is far below that
          const float r = __fdividef(acc[k] * Av[k], stab(Ap[k]));
          __nv_bfloat16 hi, lo;
          split_bf16(r, hi, lo);
          __nv_bfloat16* o = w.a3 + row * 2 * H;
          o[h] = hi; o[H + h] = lo;
        } else {
          w.wproj[row * H + h] = acc[k] * Av[k] / stab(Ap[k]);
        }
      }
    }
  }
}
// r_words / max|r_words|   (:1129-1132)
__global__ void words_norm_kernel(float* r_words, const int32_t* req_t, int T) {
  int q = blockIdx.x;
  int t = req_t[q];
  float* r = r_words + (size_t)q * T;
  float m = 0.f;
  for (int i = threadIdx.x; i <= t; i += 32) m = fmaxf(m, fabsf(r[i]));
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (m > 0.f)
    for (int i = threadIdx.x; i <= t; i += 32) r[i] = r[i] / m;
}

static size_t align_up(size_t x) { return (x + 63) & ~(size_t)63; }

static size_t grid_carve(const lrpx_gridtd_args* a, float* base, GridWs* w) {
  size_t off = 0;
  auto take = [&](size_t n) {
    float* p = base ? base + off : nullptr;
    off += align_up(n);
    return p;
  };
  size_t Q = a->Q, H = a->H, E = a->E;
  size_t nmax = 3 * H > 2 * H + 2 * E ? 3 * H : 2 * H + 2 * E;
  if ((size_t)a->C > nmax) nmax = a->C;
  GridWs t;
  t.r_h2 = take(Q * H); t.r_c2 = take(Q * H); t.r_c1 = take(Q * H); t.r_cth = take(Q * H);
  t.r_glob = take(Q * E);
  t.u = take(Q * (H > E ? H : E));
  t.v = take(Q * nmax);
  t.uctx = take(Q * a->T * H);
  t.coefavg = take(Q * a->C);
  t.wproj = take(Q * a->P * H);
  t.a3 = t.w3_g2 = t.w3_g1 = t.w3_glob = t.w3_proj = nullptr;
  t.a3u = nullptr;
  if (a->flags & LRPX_DEC_TC_GEMM) {
    auto take16 = [&](size_t n) { return reinterpret_cast<__nv_bfloat16*>(take((n + 1) / 2)); };   // n bf16 elements
    size_t rows_max = Q * a->P;
    t.a3 = take16(rows_max * 3 * H > Q * 3 * E ? rows_max * 3 * H : Q * 3 * E);
    t.w3_g2 = take16((size_t)3 * H * 3 * H);
    t.w3_g1 = take16((size_t)(2 * H + 2 * E) * 3 * H);
    t.w3_glob = take16((size_t)a->C * 3 * E);
    t.w3_proj = take16((size_t)a->C * 3 * H);
    t.a3u = take16(Q * 3 * H);
  }
  if (w) *w = t;
  return off * sizeof(float);
}

// ------------------------------------------------------------------------------------------------
// AoA
// ------------------------------------------------------------------------------------------------
struct AoaWs {
  float *r_h, *r_glob, *u, *v, *uval, *addq, *wval, *w2;
  __nv_bfloat16 *a3, *w3_aoa, *w3_g, *w3_v, *w3_proj;
};

__global__ void aoa_init_kernel(lrpx_aoa_args a, AoaWs w) {
  int q = blockIdx.x;
  int b = a.req_img[q], t = a.req_t[q], word = a.req_word[q];
  float logit = a.pred[((size_t)b * a.T + t) * a.V + word];
  float coef = logit / stab(logit);
  const float* h = a.h + ((size_t)b * (a.T + 1) + t + 1) * a.H;
  const float* ca = a.caoa + ((size_t)b * a.T + t) * a.H;
  const float* cl = a.caoa_lin + ((size_t)b * a.T + t) * a.H;
  const float* wr = a.W_fc + (size_t)word * a.H;
  for (int j = threadIdx.x; j < a.H; j += blockDim.x) {
    float rh, rc;
    fc_split(h[j], ca[j], wr[j], coef, rh, rc);               // :1092-1104
    w.r_h[(size_t)q * a.H + j] = rh;
    w.u[(size_t)q * a.H + j] = rc / stab(cl[j]);              // lin() through decoder_aoa_linear (:1107-1110)
    w.r_glob[(size_t)q * a.H + j] = 0.f;
  }
  for (int j = threadIdx.x; j < a.T; j += blockDim.x) {
    a.r_words[(size_t)q * a.T + j] = 0.f;
    if (a.r_words_raw) a.r_words_raw[(size_t)q * a.T + j] = 0.f;
  }
}
// r_ctx = ctx * v ;  uval = r_ctx / stab(ctx) on the chosen head, 0 elsewhere (lrp_mha :848-860, Q5)
__global__ void aoa_ctx_kernel(lrpx_aoa_args a, AoaWs w) {
  int q = blockIdx.x;
  int b = a.req_img[q], t = a.req_t[q], head = a.req_head[q];
  int dk = a.H / a.num_head;
  const float* cx = a.ctx + ((size_t)b * a.T + t) * a.H;
  for (int j = threadIdx.x; j < a.H; j += blockDim.x) {
    float r_ctx = cx[j] * w.v[(size_t)q * a.H + j];
    w.uval[(size_t)q * a.H + j] = (j / dk == head) ? r_ctx / stab(cx[j]) : 0.f;
  }
}
// LSTM chain without cell carry (Q4, :1115-1124): r_g = ident(r_h[i+1], i*tanh(g), c[i+1])
__global__ void aoa_cell_kernel(lrpx_aoa_args a, AoaWs w, int i) {
  int q = blockIdx.x;
  int b = a.req_img[q], t = a.req_t[q];
  bool active = i <= t;
  size_t bi = ((size_t)b * a.T + i) * a.H, bi1 = ((size_t)b * (a.T + 1) + i + 1) * a.H;
  for (int j = threadIdx.x; j < a.H; j += blockDim.x) {
    size_t o = (size_t)q * a.H + j;
    if (!active) { w.u[o] = 0.f; continue; }
    float g = a.g[bi + j];
    float r_g = a.i[bi + j] * tanhf(g) * w.r_h[o] / stab(a.c[bi1 + j]);
    w.u[o] = r_g / stab(g);
  }
}
__global__ void aoa_post_kernel(lrpx_aoa_args a, AoaWs w, int i) {
  int q = blockIdx.x;
  int b = a.req_img[q], t = a.req_t[q];
  if (i > t) return;
  const int H = a.H, E = a.E;
  const float* x = a.x + ((size_t)b * a.T + i) * (E + H);
  const float* hh = a.h + ((size_t)b * (a.T + 1) + i) * H;
  const float* vq = w.v + (size_t)q * (E + 2 * H);
  float wsum = 0.f;
  for (int k = threadIdx.x; k < E + 2 * H; k += blockDim.x) {
    float xv = k < E + H ? x[k] : hh[k - E - H];
    float rx = xv * vq[k];
    if (k < E) wsum += rx;                                            // :1130
    else if (k < E + H) w.r_glob[(size_t)q * H + (k - E)] += rx;      // :1133
    else w.r_h[(size_t)q * H + (k - E - H)] = rx;                     // :1129
  }
  for (int o = 16; o; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  __shared__ float sm[32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sm[wid] = wsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < (blockDim.x + 31) / 32; ++k) s += sm[k];
    a.r_words[(size_t)q * a.T + i] = s;
    if (a.r_words_raw) a.r_words_raw[(size_t)q * a.T + i] = s;
  }
}
// wval = r_val / stab(value), r_val = value * alpha[head][p] * uval   (:1112-1113, :1141-1144)
// addq = r_glob / stab(glob) / P                                     (:1136-1139)
__global__ void aoa_val_kernel(lrpx_aoa_args a, AoaWs w) {
  int q = blockIdx.y, p = blockIdx.x;
  int b = a.req_img[q], t = a.req_t[q], head = a.req_head[q];
  float al = a.alpha[(((size_t)b * a.T + t) * a.num_head + head) * a.P + p];
  size_t o0 = ((size_t)b * a.P + p) * a.H;
  for (int h = threadIdx.x; h < a.H; h += blockDim.x) {
    float val = a.value[o0 + h];
    w.wval[((size_t)q * a.P + p) * a.H + h] = val * al * w.uval[(size_t)q * a.H + h] / stab(val);
    if (p == 0) w.addq[(size_t)q * a.H + h] = w.r_glob[(size_t)q * a.H + h] / stab(a.glob[(size_t)b * a.H + h]) / (float)a.P;
  }
}

static size_t aoa_carve(const lrpx_aoa_args* a, float* base, AoaWs* w) {
  size_t off = 0;
  auto take = [&](size_t n) {
    float* p = base ? base + off : nullptr;
    off += align_up(n);
    return p;
  };
  size_t Q = a->Q, H = a->H;
  AoaWs t;
  t.r_h = take(Q * H); t.r_glob = take(Q * H); t.u = take(Q * H);
  t.v = take(Q * (a->E + 2 * H));
  t.uval = take(Q * H); t.addq = take(Q * H);
  t.wval = take(Q * a->P * H);
  t.w2 = take(Q * a->P * H);
  t.a3 = t.w3_aoa = t.w3_g = t.w3_v = t.w3_proj = nullptr;
  if (a->flags & LRPX_DEC_TC_GEMM) {
    auto take16 = [&](size_t n) { return reinterpret_cast<__nv_bfloat16*>(take((n + 1) / 2)); };
    t.a3 = take16(Q * a->P * 3 * H);
    t.w3_aoa = take16(H * 3 * H);
    t.w3_g = take16((size_t)(a->E + 2 * H) * 3 * H);
    t.w3_v = take16(H * 3 * H);
    t.w3_proj = take16((size_t)a->C * 3 * H);
  }
  if (w) *w = t;
  return off * sizeof(float);
}

// ------------------------------------------------------------------------------------------------
// Adaptive attention (single AdaLSTM): ExplainAdaptiveAttention.explain_caption_wordt, adaptiveattention.py:679-771
// ------------------------------------------------------------------------------------------------
struct AdaWs {
  float *r_h, *r_c, *r_glob, *u, *v, *uctx, *coefavg, *wproj;
  __nv_bfloat16 *a3, *w3_g, *w3_glob, *w3_proj;
  __nv_bfloat16* a3u;      // split form of u written by the cell kernel itself (null: CUDA-core GEMM)
};

// fc rule on the target row, split into h / ctx_hat, ctx_hat into context / sentinel (:700-724)
__global__ void ada_init_kernel(lrpx_adaptive_args a, AdaWs w) {
  int q = blockIdx.x;
  int b = a.req_img[q], t = a.req_t[q], word = a.req_word[q];
  float logit = a.pred[((size_t)b * a.T + t) * a.V + word];
  float coef = logit / stab(logit);
  const float* h = a.h + ((size_t)b * (a.T + 1) + t + 1) * a.H;
  size_t bi = ((size_t)b * a.T + t) * a.H;
  const float* wr = a.W_fc + (size_t)word * a.H;
  float beta = a.beta[(size_t)b * a.T + t];
  for (int j = threadIdx.x; j < a.H; j += blockDim.x) {
    float rh, rcth;
    float cth = a.ctx_hat[bi + j];
    fc_split(h[j], cth, wr[j], coef, rh, rcth);
    float dct = stab(cth);
    float cx = a.ctx[bi + j];
    float r_ctx = (1.f - beta) * cx * rcth / dct;
    size_t o = (size_t)q * a.H + j;
    w.r_h[o] = rh;
    w.r_c[o] = beta * a.st[bi + j] * rcth / dct;          // r_ct[t+1] = r_st (:724)
    w.uctx[o] = r_ctx / stab(cx);
  }
  for (int j = threadIdx.x; j < a.E; j += blockDim.x) w.r_glob[(size_t)q * a.E + j] = 0.f;
  for (int j = threadIdx.x; j < a.T; j += blockDim.x) {
    a.r_words[(size_t)q * a.T + j] = 0.f;
    if (a.r_words_raw) a.r_words_raw[(size_t)q * a.T + j] = 0.f;
  }
}

// cell rule (:726-734) -> u = r_g / stab(tanh g)   (the reference passes tanh(gt) as forward_output, :737)
__global__ void ada_cell_kernel(lrpx_adaptive_args a, AdaWs w, int i) {
  int q = blockIdx.x;
  int b = a.req_img[q], t = a.req_t[q];
  bool active = i <= t;
  size_t bi = ((size_t)b * a.T + i) * a.H, bi1 = ((size_t)b * (a.T + 1) + i + 1) * a.H,
         bi0 = ((size_t)b * (a.T + 1) + i) * a.H;
  for (int j = threadIdx.x; j < a.H; j += blockDim.x) {
    size_t o = (size_t)q * a.H + j;
    if (!active) { put_operand(w.u, w.a3u, q, a.H, j, 0.f); continue; }
    float rc = w.r_c[o] + w.r_h[o];
    float d = stab(a.c[bi1 + j]);
    float tg = tanhf(a.g[bi + j]);
    float r_g = a.i[bi + j] * tg * rc / d;
    w.r_c[o] = a.f[bi + j] * a.c[bi0 + j] * rc / d;
    put_operand(w.u, w.a3u, q, a.H, j, r_g / stab(tg));
  }
}

// after v = u @ W_g : slices of xht = [emb | glob | h_i] (:739-742)
__global__ void ada_post_kernel(lrpx_adaptive_args a, AdaWs w, int i) {
  int q = blockIdx.x;
  int b = a.req_img[q], t = a.req_t[q];
  if (i > t) return;
  const int H = a.H, E = a.E;
  const float* x = a.x + ((size_t)b * a.T + i) * 2 * E;
  const float* hp = a.h + ((size_t)b * (a.T + 1) + i) * H;
  const float* vq = w.v + (size_t)q * (2 * E + H);
  float wsum = 0.f;
  for (int k = threadIdx.x; k < 2 * E + H; k += blockDim.x) {
    if (k < E) wsum += x[k] * vq[k];
    else if (k < 2 * E) { if (i == t) w.r_glob[(size_t)q * E + (k - E)] = x[k] * vq[k]; }   // only the explained step (:740)
    else w.r_h[(size_t)q * H + (k - 2 * E)] = hp[k - 2 * E] * vq[k];
  }
  for (int o = 16; o; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
  __shared__ float sm[32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sm[wid] = wsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < (blockDim.x + 31) / 32; ++k) s += sm[k];
    float* dst = a.r_words_raw ? a.r_words_raw : a.r_words;
    dst[(size_t)q * a.T + i] = s;
    if (a.r_words_raw) a.r_words[(size_t)q * a.T + i] = s;
  }
}

// u_g = r_glob / stab(avg @ W_glob^T)   (:743-746, forward_output=False: no bias)
__global__ void ada_glob_kernel(lrpx_adaptive_args a, AdaWs w) {
  int q = blockIdx.x, b = a.req_img[q];
  for (int j = threadIdx.x; j < a.E; j += blockDim.x)
    w.u[(size_t)q * a.E + j] = w.r_glob[(size_t)q * a.E + j] / stab(a.z_glob[(size_t)b * a.E + j]);
}
// coefavg = avg * v / stab(avg) / P  (mean rule :752-755 applied to r_avg = avg * v)
__global__ void ada_avg_kernel(lrpx_adaptive_args a, AdaWs w) {
  int q = blockIdx.x, b = a.req_img[q];
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float av = a.avg[(size_t)b * a.C + c];
    float r_avg = av * w.v[(size_t)q * a.C + c];
    w.coefavg[(size_t)q * a.C + c] = r_avg / stab(av) / (float)a.P;
  }
}
// attention weighted-sum rule at the explained step only (:756-759), already divided for the projector rule (:760-763):
//   wproj[q][p][h] = A[p][h] * alpha_t[p] * uctx[h] / stab(z_proj[p][h]);  block = (request, pixel group of 8)
template <bool SPLIT>
__global__ void __launch_bounds__(256) ada_attn_kernel(lrpx_adaptive_args a, AdaWs w) {
  const int H = a.H, P = a.P;
  const int ng = (P + 7) / 8;
  const int q = blockIdx.x / ng;
  const int b = a.req_img[q], t = a.req_t[q];
  const int p0 = (blockIdx.x - q * ng) * 8, p1 = min(P, p0 + 8);
  const float* al = a.alpha + ((size_t)b * a.T + t) * P;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const float uv = w.uctx[(size_t)q * H + h];
    for (int p = p0; p < p1; ++p) {
      const size_t o0 = ((size_t)b * P + p) * H + h;
      const size_t row = (size_t)q * P + p;
      const float num = __ldg(a.A + o0) * al[p] * uv, den = stab(__ldg(a.z_proj + o0));
      if (SPLIT) {
        __nv_bfloat16 hi, lo;
        split_bf16(__fdividef(num, den), hi, lo);       // bf16 hi + lo keeps 16 mantissa bits: above the 2-ulp division
        __nv_bfloat16* o = w.a3 + row * 2 * H;
        o[h] = hi; o[H + h] = lo;
      } else {
        w.wproj[row * H + h] = num / den;
      }
    }
  }
}

static size_t ada_carve(const lrpx_adaptive_args* a, float* base, AdaWs* w) {
  size_t off = 0;
  auto take = [&](size_t n) {
    float* p = base ? base + off : nullptr;
    off += align_up(n);
    return p;
  };
  size_t Q = a->Q, H = a->H, E = a->E;
  size_t nmax = 2 * E + H > (size_t)a->C ? 2 * E + H : (size_t)a->C;
  AdaWs t;
  t.r_h = take(Q * H); t.r_c = take(Q * H);
  t.r_glob = take(Q * E);
  t.u = take(Q * (H > E ? H : E));
  t.v = take(Q * nmax);
  t.uctx = take(Q * H);
  t.coefavg = take(Q * a->C);
  t.wproj = take(Q * a->P * H);
  t.a3 = t.w3_g = t.w3_glob = t.w3_proj = t.a3u = nullptr;
  if (a->flags & LRPX_DEC_TC_GEMM) {
    auto take16 = [&](size_t n) { return reinterpret_cast<__nv_bfloat16*>(take((n + 1) / 2)); };   // n bf16 elements
    size_t rows_max = Q * a->P;
    t.a3 = take16(rows_max * 3 * H > Q * 3 * E ? rows_max * 3 * H : Q * 3 * E);
    t.w3_g = take16((size_t)(2 * E + H) * 3 * H);
    t.w3_glob = take16((size_t)a->C * 3 * E);
    t.w3_proj = take16((size_t)a->C * 3 * H);
    t.a3u = take16(Q * 3 * H);
  }
  if (w) *w = t;
  return off * sizeof(float);
}

// ------------------------------------------------------------------------------------------------
// lrp_tune weights: one block per sample
// ------------------------------------------------------------------------------------------------
__global__ void fc_lrp_weights_kernel(const float* __restrict__ logits, const float* __restrict__ h,
                                      const float* __restrict__ ctx, const float* __restrict__ W_fc,
                                      const uint8_t* __restrict__ is_stop, float* __restrict__ w_ctx,
                                      float* __restrict__ w_h, int32_t* __restrict__ argmax_out, int V, int H) {
  int b = blockIdx.x;
  const float* lg = logits + (size_t)b * V;
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  __shared__ float s_m[2][32];
  // argmax, first index wins on ties (torch.argmax, gridTDmodel.py:555); NaN counts as the maximum (torch's order:
  // a diverged tuning run with NaN logits returns the first NaN's index there and must not index out of bounds here)
  auto better = [](float v, int j, float bv, int bi) {
    const bool vn = v != v, bn = bv != bv;
    if (vn || bn) return vn && (!bn || j < bi);
    return v > bv || (v == bv && j < bi);
  };
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int j = threadIdx.x; j < V; j += blockDim.x) {
    float v = lg[j];
    if (better(v, j, bv, bi)) { bv = v; bi = j; }
  }
  for (int o = 16; o; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) / 32;
  if (lane == 0) { s_val[wid] = bv; s_idx[wid] = bi; }
  __syncthreads();
  if (wid == 0) {
    bv = lane < nw ? s_val[lane] : -INFINITY;
    bi = lane < nw ? s_idx[lane] : 0x7fffffff;
    for (int o = 16; o; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_val[0] = bv; s_idx[0] = (bi >= 0 && bi < V) ? bi : 0; }
  }
  __syncthreads();
  int word = s_idx[0];
  float logit = s_val[0];
  if (argmax_out && threadIdx.x == 0) argmax_out[b] = word;
  bool stop = is_stop[word] != 0;                                   // Q15
  float coef = logit / stab(logit);
  const float* wr = W_fc + (size_t)word * H;
  // pass 1: max |r|
  float mh = 0.f, mc = 0.f;
  if (!stop)
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
      float rh, rc;
      fc_split(h[(size_t)b * H + j], ctx[(size_t)b * H + j], wr[j], coef, rh, rc);
      mh = fmaxf(mh, fabsf(rh));
      mc = fmaxf(mc, fabsf(rc));
    }
  for (int o = 16; o; o >>= 1) {
    mh = fmaxf(mh, __shfl_xor_sync(0xffffffffu, mh, o));
    mc = fmaxf(mc, __shfl_xor_sync(0xffffffffu, mc, o));
  }
  if (lane == 0) { s_m[0][wid] = mh; s_m[1][wid] = mc; }
  __syncthreads();
  mh = 0.f; mc = 0.f;
  for (int k = 0; k < nw; ++k) { mh = fmaxf(mh, s_m[0][k]); mc = fmaxf(mc, s_m[1][k]); }
  if (mh == 0.f) mh = 1.f;                                          // utils.py:59
  if (mc == 0.f) mc = 1.f;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    float rh = 0.f, rc = 0.f;
    if (!stop) fc_split(h[(size_t)b * H + j], ctx[(size_t)b * H + j], wr[j], coef, rh, rc);
    w_h[(size_t)b * H + j] = rh / mh + 1.f;
    w_ctx[(size_t)b * H + j] = rc / mc + 1.f;
  }
}

}  // namespace lrpx

using namespace lrpx;

#define RUN(expr)                 \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != LRPX_OK) return rc__; \
  } while (0)

extern "C" {

size_t lrpx_gridtd_decoder_workspace_bytes(const lrpx_gridtd_args* a) {
  if (!a) return 0;
  return grid_carve(a, nullptr, nullptr);
}

int lrpx_gridtd_decoder_lrp_f32(const lrpx_gridtd_args* a, void* workspace, size_t workspace_bytes, void* stream) {
  LRPX_CHECK_ARG(a, "null args");
  LRPX_CHECK_ARG(a->B > 0 && a->T > 0 && a->H > 0 && a->E > 0 && a->P > 0 && a->C > 0 && a->V > 0 && a->Q >= 0,
                 "bad dimensions");
  if (a->Q == 0) return LRPX_OK;
  LRPX_CHECK_ARG(a->feat && a->avg && a->A_pre && a->A && a->glob_pre && a->x1 && a->x2 && a->h1 && a->c1 && a->h2 &&
                     a->c2 && a->g1 && a->i1 && a->f1 && a->g2 && a->i2 && a->f2 && a->st && a->ctx && a->ctx_hat &&
                     a->alpha && a->beta && a->pred && a->W_g1 && a->W_g2 && a->W_fc && a->W_glob && a->W_proj &&
                     a->req_img && a->req_t && a->req_word && a->r_feat && a->r_words,
                 "null pointer in args");
  GridWs w;
  size_t need = grid_carve(a, (float*)workspace, &w);
  LRPX_CHECK_ARG(workspace && workspace_bytes >= need, "workspace too small");
  cudaStream_t st = as_stream(stream);
  const int Q = a->Q, H = a->H, E = a->E, T = a->T;
  int nt = H >= 256 ? 256 : 128;
  GemmEpi none{};
  // tensor-core GEMMs where the shape allows it (per GEMM), CUDA cores otherwise
  const bool tc = (a->flags & LRPX_DEC_TC_GEMM) != 0;
  const __nv_bfloat16* w3_g2 = (tc && tc_shape_ok(3 * H, H)) ? prep_weight3(a->W_g2, w.w3_g2, H, 3 * H, st) : nullptr;
  const __nv_bfloat16* w3_g1 =
      (tc && tc_shape_ok(2 * H + 2 * E, H)) ? prep_weight3(a->W_g1, w.w3_g1, H, 2 * H + 2 * E, st) : nullptr;
  const __nv_bfloat16* w3_glob = (tc && tc_shape_ok(a->C, E)) ? prep_weight3(a->W_glob, w.w3_glob, E, a->C, st) : nullptr;
  const __nv_bfloat16* w3_proj = (tc && tc_shape_ok(a->C, H)) ? prep_weight3(a->W_proj, w.w3_proj, H, a->C, st) : nullptr;
  // the step kernels write the split operand themselves when both step GEMMs run on the tensor cores
  const bool fused_split = w3_g2 && w3_g1;
  if (!fused_split) w.a3u = nullptr;
  __nv_bfloat16* a3_step = fused_split ? w.a3u : w.a3;
  grid_init_kernel<<<Q, nt, 0, st>>>(*a, w);
  cudaMemsetAsync(w.uctx, 0, (size_t)Q * T * H * sizeof(float), st);
  for (int i = T - 1; i >= 0; --i) {
    grid_cell2_kernel<<<Q, nt, 0, st>>>(*a, w, i);
    RUN(gemm_any<GE_STORE>(w.u, a->W_g2, w3_g2, a3_step, w.v, Q, 3 * H, H, none, st, fused_split));
    grid_post2_kernel<<<Q, nt, 0, st>>>(*a, w, i);
    RUN(gemm_any<GE_STORE>(w.u, a->W_g1, w3_g1, a3_step, w.v, Q, 2 * H + 2 * E, H, none, st, fused_split));
    grid_post1_kernel<<<Q, 256, 0, st>>>(*a, w, i);
  }
  grid_glob_kernel<<<Q, 128, 0, st>>>(*a, w);
  RUN(gemm_any<GE_STORE>(w.u, a->W_glob, w3_glob, w.a3, w.v, Q, a->C, E, none, st));
  grid_avg_kernel<<<Q, 128, 0, st>>>(*a, w);
  GemmEpi fe{a->feat, nullptr, w.coefavg, a->req_img, a->P};
  const size_t att_smem = (size_t)T * (((a->P + 3) & ~3) + H) * sizeof(float);
  if (att_smem <= 160 * 1024) {
    int at = H >= 512 ? 512 : (H >= 256 ? 256 : 128);
    if ((H & 3) == 0) at = H / 4 >= 512 ? 512 : ((H / 4 + 31) & ~31);        // four hidden units per thread
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(grid_attn_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      cudaFuncSetAttribute(grid_attn_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      attr_done = true;
    }
    if (w3_proj) grid_attn_rows_kernel<true><<<Q, at, att_smem, st>>>(*a, w);
    else grid_attn_rows_kernel<false><<<Q, at, att_smem, st>>>(*a, w);
    RUN(gemm_any<GE_FEAT>(w.wproj, a->W_proj, w3_proj, w.a3, a->r_feat, Q * a->P, a->C, H, fe, st, w3_proj != nullptr));
  } else {
    grid_attn_kernel<<<dim3(a->P, Q), nt, 0, st>>>(*a, w);
    RUN(gemm_any<GE_FEAT>(w.wproj, a->W_proj, w3_proj, w.a3, a->r_feat, Q * a->P, a->C, H, fe, st));
  }
  words_norm_kernel<<<Q, 32, 0, st>>>(a->r_words, a->req_t, T);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

size_t lrpx_aoa_decoder_workspace_bytes(const lrpx_aoa_args* a) {
  if (!a) return 0;
  return aoa_carve(a, nullptr, nullptr);
}

int lrpx_aoa_decoder_lrp_f32(const lrpx_aoa_args* a, void* workspace, size_t workspace_bytes, void* stream) {
  LRPX_CHECK_ARG(a, "null args");
  LRPX_CHECK_ARG(a->B > 0 && a->T > 0 && a->H > 0 && a->E > 0 && a->P > 0 && a->C > 0 && a->V > 0 && a->Q >= 0 &&
                     a->num_head > 0 && a->H % a->num_head == 0,
                 "bad dimensions");
  if (a->Q == 0) return LRPX_OK;
  LRPX_CHECK_ARG(a->feat && a->A_pre && a->A && a->glob && a->value && a->x && a->h && a->c && a->g && a->i &&
                     a->ctx && a->caoa && a->caoa_lin && a->alpha && a->pred && a->W_g && a->W_fc && a->W_aoa &&
                     a->W_v && a->W_proj && a->req_img && a->req_t && a->req_word && a->req_head && a->r_feat &&
                     a->r_words,
                 "null pointer in args");
  AoaWs w;
  size_t need = aoa_carve(a, (float*)workspace, &w);
  LRPX_CHECK_ARG(workspace && workspace_bytes >= need, "workspace too small");
  cudaStream_t st = as_stream(stream);
  const int Q = a->Q, H = a->H, E = a->E, T = a->T;
  int nt = H >= 256 ? 256 : 128;
  GemmEpi none{};
  const bool tc = (a->flags & LRPX_DEC_TC_GEMM) != 0;
  const __nv_bfloat16* w3_aoa = (tc && tc_shape_ok(H, H)) ? prep_weight3(a->W_aoa, w.w3_aoa, H, H, st) : nullptr;
  const __nv_bfloat16* w3_g = (tc && tc_shape_ok(E + 2 * H, H)) ? prep_weight3(a->W_g, w.w3_g, H, E + 2 * H, st) : nullptr;
  const __nv_bfloat16* w3_v = (tc && tc_shape_ok(H, H)) ? prep_weight3(a->W_v, w.w3_v, H, H, st) : nullptr;
  const __nv_bfloat16* w3_proj = (tc && tc_shape_ok(a->C, H)) ? prep_weight3(a->W_proj, w.w3_proj, H, a->C, st) : nullptr;
  aoa_init_kernel<<<Q, nt, 0, st>>>(*a, w);
  RUN(gemm_any<GE_STORE>(w.u, a->W_aoa, w3_aoa, w.a3, w.v, Q, H, H, none, st));
  aoa_ctx_kernel<<<Q, nt, 0, st>>>(*a, w);
  for (int i = T - 1; i >= 0; --i) {
    aoa_cell_kernel<<<Q, nt, 0, st>>>(*a, w, i);
    RUN(gemm_any<GE_STORE>(w.u, a->W_g, w3_g, w.a3, w.v, Q, E + 2 * H, H, none, st));
    aoa_post_kernel<<<Q, 256, 0, st>>>(*a, w, i);
  }
  aoa_val_kernel<<<dim3(a->P, Q), nt, 0, st>>>(*a, w);
  GemmEpi pe{a->A, a->A_pre, w.addq, a->req_img, a->P};
  RUN(gemm_any<GE_AOA_PROJ>(w.wval, a->W_v, w3_v, w.a3, w.w2, Q * a->P, H, H, pe, st));
  GemmEpi fe{a->feat, nullptr, nullptr, a->req_img, a->P};
  RUN(gemm_any<GE_FEAT>(w.w2, a->W_proj, w3_proj, w.a3, a->r_feat, Q * a->P, a->C, H, fe, st));
  words_norm_kernel<<<Q, 32, 0, st>>>(a->r_words, a->req_t, T);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

size_t lrpx_adaptive_decoder_workspace_bytes(const lrpx_adaptive_args* a) {
  if (!a) return 0;
  return ada_carve(a, nullptr, nullptr);
}

int lrpx_adaptive_decoder_lrp_f32(const lrpx_adaptive_args* a, void* workspace, size_t workspace_bytes, void* stream) {
  LRPX_CHECK_ARG(a, "null args");
  LRPX_CHECK_ARG(a->B > 0 && a->T > 0 && a->H > 0 && a->E > 0 && a->P > 0 && a->C > 0 && a->V > 0 && a->Q >= 0,
                 "bad dimensions");
  if (a->Q == 0) return LRPX_OK;
  LRPX_CHECK_ARG(a->feat && a->avg && a->z_proj && a->A && a->z_glob && a->x && a->h && a->c && a->g && a->i && a->f &&
                     a->st && a->ctx && a->ctx_hat && a->alpha && a->beta && a->pred && a->W_g && a->W_fc && a->W_glob &&
                     a->W_proj && a->req_img && a->req_t && a->req_word && a->r_feat && a->r_words,
                 "null pointer in args");
  AdaWs w;
  size_t need = ada_carve(a, (float*)workspace, &w);
  LRPX_CHECK_ARG(workspace && workspace_bytes >= need, "workspace too small");
  cudaStream_t st = as_stream(stream);
  const int Q = a->Q, H = a->H, E = a->E, T = a->T;
  int nt = H >= 256 ? 256 : 128;
  GemmEpi none{};
  const bool tc = (a->flags & LRPX_DEC_TC_GEMM) != 0;
  const __nv_bfloat16* w3_g = (tc && tc_shape_ok(2 * E + H, H)) ? prep_weight3(a->W_g, w.w3_g, H, 2 * E + H, st) : nullptr;
  const __nv_bfloat16* w3_glob = (tc && tc_shape_ok(a->C, E)) ? prep_weight3(a->W_glob, w.w3_glob, E, a->C, st) : nullptr;
  const __nv_bfloat16* w3_proj = (tc && tc_shape_ok(a->C, H)) ? prep_weight3(a->W_proj, w.w3_proj, H, a->C, st) : nullptr;
  if (!w3_g) w.a3u = nullptr;
  ada_init_kernel<<<Q, nt, 0, st>>>(*a, w);
  for (int i = T - 1; i >= 0; --i) {
    ada_cell_kernel<<<Q, nt, 0, st>>>(*a, w, i);
    RUN(gemm_any<GE_STORE>(w.u, a->W_g, w3_g, w.a3u, w.v, Q, 2 * E + H, H, none, st, w3_g != nullptr));
    ada_post_kernel<<<Q, 256, 0, st>>>(*a, w, i);
  }
  ada_glob_kernel<<<Q, 128, 0, st>>>(*a, w);
  RUN(gemm_any<GE_STORE>(w.u, a->W_glob, w3_glob, w.a3, w.v, Q, a->C, E, none, st));
  ada_avg_kernel<<<Q, 128, 0, st>>>(*a, w);
  GemmEpi fe{a->feat, nullptr, w.coefavg, a->req_img, a->P};
  const unsigned ag = (unsigned)((a->P + 7) / 8) * (unsigned)Q;
  if (w3_proj) ada_attn_kernel<true><<<ag, nt, 0, st>>>(*a, w);
  else ada_attn_kernel<false><<<ag, nt, 0, st>>>(*a, w);
  RUN(gemm_any<GE_FEAT>(w.wproj, a->W_proj, w3_proj, w.a3, a->r_feat, Q * a->P, a->C, H, fe, st, w3_proj != nullptr));
  words_norm_kernel<<<Q, 32, 0, st>>>(a->r_words, a->req_t, T);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

int lrpx_fc_lrp_weights_f32(const float* logits, const float* h, const float* ctx, const float* W_fc,
                            const uint8_t* is_stop, float* w_ctx, float* w_h, int32_t* argmax_out, int B, int V, int H,
                            void* stream) {
  LRPX_CHECK_ARG(logits && h && ctx && W_fc && is_stop && w_ctx && w_h && B >= 0 && V > 0 && H > 0, "bad argument");
  if (B == 0) return LRPX_OK;
  fc_lrp_weights_kernel<<<B, 256, 0, as_stream(stream)>>>(logits, h, ctx, W_fc, is_stop, w_ctx, w_h, argmax_out, V, H);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

}  // extern "C"
