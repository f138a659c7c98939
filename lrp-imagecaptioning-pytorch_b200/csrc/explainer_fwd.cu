// Fused element-wise steps of the explainer's teacher-forced forward (the producer of the saved state the decoder
// relevance kernels consume), fp32.
//
//   ExplainGridTDAttention.get_hidden_parameters   models/gridTDmodel.py:933-1012
//     adalstm_forward / language_lstm_forward      :773-797   (hand-rolled LSTM cells returning g, i, f)
//     AdaptiveAttention.forward                    :61-103    (adaptive attention with the sentinel)
//
// The reference runs ~30 small tensor ops per time step.  Here a step is: one GEMM over the concatenated
// recurrent inputs (library GEMM on the host side), lrpx_lstm_cell_f32, one GEMM for the two attention
// projections, lrpx_adaptive_attention_f32, one GEMM, lrpx_lstm_cell_f32.  Outputs go straight into the
// (B, T, .) saved-state tensors (row strides) and into the staging rows of the next GEMM.
#include "lrpx_common.cuh"

namespace lrpx {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// z: (B, >=4H) pre-activations in torch's gate order i, f, g, o (gridTDmodel.py:777-783)
__global__ void lstm_cell_kernel(lrpx_lstm_cell_args a) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.H) return;
  const float* z = a.z + (size_t)b * a.ldz;
  const float zi = z[j], zf = z[a.H + j], zg = z[2 * a.H + j], zo = z[3 * a.H + j];
  const float i = sigmoidf_(zi), f = sigmoidf_(zf);
  const float c = f * a.c_prev[(size_t)b * a.ld_cprev + j] + i * tanhf(zg);
  const float tc = tanhf(c);
  const float h = sigmoidf_(zo) * tc;
  a.h[(size_t)b * a.ld_state + j] = h;
  a.c[(size_t)b * a.ld_state + j] = c;
  a.g[(size_t)b * a.ld_gate + j] = zg;
  a.i[(size_t)b * a.ld_gate + j] = i;
  a.f[(size_t)b * a.ld_gate + j] = f;
  if (a.h_copy0) a.h_copy0[(size_t)b * a.ld_copy0 + j] = h;
  if (a.h_copy1) a.h_copy1[(size_t)b * a.ld_copy1 + j] = h;
  if (a.h_copy2) a.h_copy2[(size_t)b * a.ld_copy2 + j] = h;
  if (a.gate_pre) {      // sentinel  s = sigmoid(x_gate(x) + h_gate(h_old)) * tanh(c_new)   (:982-983)
    const float s = sigmoidf_(a.gate_pre[(size_t)b * a.ld_gate_pre + j]) * tc;
    a.s[(size_t)b * a.ld_gate + j] = s;
    if (a.s_copy) a.s_copy[(size_t)b * a.ld_s_copy + j] = s;
  }
}

// One block per image.  AdaptiveAttention.attend (gridTDmodel.py:80-103):
//   z[p]   = sum_k w_h[k] * tanh(img_proj[b][p][k] + hproj[b][p])     (sic: the reference adds the h projection
//            along the pixel axis — its bmm with a ones matrix, :81-87 — which is only shape-valid for P == K, Q19)
//   alpha  = softmax_p z;  ctx = sum_p alpha[p] * A[b][p][:]
//   zs     = sum_k w_h[k] * tanh(sproj[b][k] + hproj[b][k]);  beta = softmax([z ; zs])[-1]
//   ctx_hat = beta * s + (1 - beta) * ctx
__global__ void __launch_bounds__(512) adaptive_attention_kernel(lrpx_ada_attention_args a) {
  extern __shared__ float sm[];            // z[P] | hproj[K] | sproj[K] | w_h[K] | red[32]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int P = a.P, K = a.K, H = a.H;
  float* z = sm;
  float* hp = z + P;
  float* sp = hp + K;
  float* wh = sp + K;
  float* red = wh + K;
  for (int k = tid; k < K; k += blockDim.x) {
    hp[k] = a.hs_proj[(size_t)b * a.ld_hs + k];
    sp[k] = a.hs_proj[(size_t)b * a.ld_hs + K + k];
    wh[k] = a.w_h[k];
  }
  __syncthreads();
  const float* ip = a.img_proj + (size_t)b * P * K;
  for (int p = warp; p <= P; p += nwarp) {         // row P is the sentinel
    float acc = 0.f;
    if (p < P) {
      const float hpp = hp[p];      // Q19: the h projection is broadcast along the PIXEL axis here (:81-87)
      for (int k = lane; k < K; k += 32) acc += wh[k] * tanhf(ip[(size_t)p * K + k] + hpp);
    } else {
      for (int k = lane; k < K; k += 32) acc += wh[k] * tanhf(sp[k] + hp[k]);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      if (p < P) z[p] = acc; else red[31] = acc;
    }
  }
  __syncthreads();
  const float zs = red[31];
  __syncthreads();
  // max over the P pixel scores (the sentinel joins for the second softmax)
  float m = -INFINITY;
  for (int p = tid; p < P; p += blockDim.x) m = fmaxf(m, z[p]);
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = -INFINITY;
  for (int w = 0; w < nwarp; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int p = tid; p < P; p += blockDim.x) sum += expf(z[p] - m);
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < nwarp; ++w) sum += red[w];
  __syncthreads();
  // alpha = softmax over the pixels;  beta from the softmax over pixels + sentinel
  const float m2 = fmaxf(m, zs);
  const float sum2 = sum * expf(m - m2) + expf(zs - m2);
  const float beta = expf(zs - m2) / sum2;
  for (int p = tid; p < P; p += blockDim.x) {
    const float al = expf(z[p] - m) / sum;
    z[p] = al;
    a.alpha[(size_t)b * a.ld_alpha + p] = al;
  }
  if (tid == 0) a.beta[(size_t)b * a.ld_beta] = beta;
  __syncthreads();
  const float* A = a.A + (size_t)b * P * H;
  for (int h = tid; h < H; h += blockDim.x) {
    float acc = 0.f;
    for (int p = 0; p < P; ++p) acc += A[(size_t)p * H + h] * z[p];
    const float s = a.s[(size_t)b * a.ld_s + h];
    const float ch = beta * s + (1.f - beta) * acc;
    a.ctx[(size_t)b * a.ld_out + h] = acc;
    a.ctx_hat[(size_t)b * a.ld_out + h] = ch;
    if (a.ctx_hat_copy) a.ctx_hat_copy[(size_t)b * a.ld_copy + h] = ch;
  }
}

}  // namespace lrpx

using namespace lrpx;

extern "C" int lrpx_lstm_cell_f32(const lrpx_lstm_cell_args* a, void* stream) {
  LRPX_CHECK_ARG(a && a->B > 0 && a->H > 0, "bad shape");
  LRPX_CHECK_ARG(a->z && a->c_prev && a->h && a->c && a->g && a->i && a->f, "null pointer");
  LRPX_CHECK_ARG(!a->gate_pre || a->s, "gate_pre needs the sentinel output s");
  dim3 grid(ceil_div(a->H, 128), a->B);
  lstm_cell_kernel<<<grid, 128, 0, as_stream(stream)>>>(*a);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

extern "C" int lrpx_adaptive_attention_f32(const lrpx_ada_attention_args* a, void* stream) {
  LRPX_CHECK_ARG(a && a->B > 0 && a->P > 0 && a->K > 0 && a->H > 0, "bad shape");
  LRPX_CHECK_ARG(a->P == a->K, "the reference's attention is only defined for P == n_pixel (gridTDmodel.py:81-87)");
  LRPX_CHECK_ARG(a->A && a->img_proj && a->hs_proj && a->w_h && a->s && a->ctx && a->ctx_hat && a->alpha && a->beta,
                 "null pointer");
  size_t smem = (size_t)(a->P + 3 * a->K + 32) * sizeof(float);
  LRPX_CHECK_ARG(smem <= 48 * 1024, "P + 3K too large for one block");
  adaptive_attention_kernel<<<a->B, 512, smem, as_stream(stream)>>>(*a);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}
