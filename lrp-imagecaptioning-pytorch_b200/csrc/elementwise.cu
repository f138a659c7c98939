// Element-wise and pooling relevance rules (HBM-bound), fp32 NCHW — parity path.
// Reference: LRPtools/lrp_modules.py:39-54 (ReLU), :172-195 (Pool2d), :197-246 (BatchNorm),
//            :256-280 (Add), LRPtools/utils.py:55-64 (normalize_relevance).
#include "lrpx_common.cuh"
#include <math.h>

namespace lrpx {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------- max-pool forward + argmax
__global__ void maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t* __restrict__ idx,
                                   lrpx_pool_shape s, int oh, int ow, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int q = (int)(i % ow);
    int p = (int)((i / ow) % oh);
    long long nc = i / ((long long)ow * oh);
    const float* plane = x + nc * (long long)s.h * s.w;
    int h0 = p * s.stride_h - s.pad_h, w0 = q * s.stride_w - s.pad_w;
    int hs = max(h0, 0), ws = max(w0, 0);
    int he = min(h0 + s.kh, s.h), we = min(w0 + s.kw, s.w);
    // PyTorch max_pool2d_with_indices: maxidx starts at the first in-range element, value -inf,
    // update when (val > maxval) || isnan(val)
    int best = hs * s.w + ws;
    float bv = -INFINITY;
    for (int hh = hs; hh < he; ++hh)
      for (int ww = ws; ww < we; ++ww) {
        float v = plane[hh * s.w + ww];
        if (v > bv || isnan(v)) { bv = v; best = hh * s.w + ww; }
      }
    if (y) y[i] = bv;
    if (idx) idx[i] = best;
  }
}

// ---------------------------------------------------------------- max-pool WTA (gather form)
__global__ void maxpool_wta_kernel(const float* __restrict__ x, const float* __restrict__ r_out,
                                   float* __restrict__ r_in, lrpx_pool_shape s, int oh, int ow, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int xw = (int)(i % s.w);
    int xh = (int)((i / s.w) % s.h);
    long long nc = i / ((long long)s.w * s.h);
    const float* plane = x + nc * (long long)s.h * s.w;
    const float* rplane = r_out + nc * (long long)oh * ow;
    int me = xh * s.w + xw;
    // windows (p,q) that contain (xh,xw):  p*stride - pad <= xh < p*stride - pad + kh
    int p_lo = (xh + s.pad_h - s.kh + s.stride_h) / s.stride_h;  // ceil((xh+pad-kh+1)/stride)
    if (xh + s.pad_h - s.kh + 1 <= 0) p_lo = 0;
    int p_hi = min((xh + s.pad_h) / s.stride_h, oh - 1);
    int q_lo = (xw + s.pad_w - s.kw + s.stride_w) / s.stride_w;
    if (xw + s.pad_w - s.kw + 1 <= 0) q_lo = 0;
    int q_hi = min((xw + s.pad_w) / s.stride_w, ow - 1);
    float acc = 0.f;
    for (int p = p_lo; p <= p_hi; ++p)
      for (int q = q_lo; q <= q_hi; ++q) {
        int h0 = p * s.stride_h - s.pad_h, w0 = q * s.stride_w - s.pad_w;
        int hs = max(h0, 0), ws = max(w0, 0);
        int he = min(h0 + s.kh, s.h), we = min(w0 + s.kw, s.w);
        int best = hs * s.w + ws;
        float bv = -INFINITY;
        for (int hh = hs; hh < he; ++hh)
          for (int ww = ws; ww < we; ++ww) {
            float v = plane[hh * s.w + ww];
            if (v > bv || isnan(v)) { bv = v; best = hh * s.w + ww; }
          }
        if (best == me) acc += safe_div(rplane[p * ow + q], bv);
      }
    r_in[i] = plane[me] * acc;
  }
}

// ---------------------------------------------------------------- avg-pool proportional
__global__ void avgpool_prop_kernel(const float* __restrict__ x, const float* __restrict__ r_out,
                                    float* __restrict__ r_in, lrpx_pool_shape s, int oh, int ow, long long total) {
  const float inv = 1.f / (float)(s.kh * s.kw);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int xw = (int)(i % s.w);
    int xh = (int)((i / s.w) % s.h);
    long long nc = i / ((long long)s.w * s.h);
    const float* plane = x + nc * (long long)s.h * s.w;
    const float* rplane = r_out + nc * (long long)oh * ow;
    int p_lo = (xh + s.pad_h - s.kh + 1 <= 0) ? 0 : (xh + s.pad_h - s.kh + s.stride_h) / s.stride_h;
    int p_hi = min((xh + s.pad_h) / s.stride_h, oh - 1);
    int q_lo = (xw + s.pad_w - s.kw + 1 <= 0) ? 0 : (xw + s.pad_w - s.kw + s.stride_w) / s.stride_w;
    int q_hi = min((xw + s.pad_w) / s.stride_w, ow - 1);
    float acc = 0.f;
    for (int p = p_lo; p <= p_hi; ++p)
      for (int q = q_lo; q <= q_hi; ++q) {
        int h0 = p * s.stride_h - s.pad_h, w0 = q * s.stride_w - s.pad_w;
        float sum = 0.f;
        for (int hh = max(h0, 0); hh < min(h0 + s.kh, s.h); ++hh)
          for (int ww = max(w0, 0); ww < min(w0 + s.kw, s.w); ++ww) sum += plane[hh * s.w + ww];
        float z = sum * inv;  // count_include_pad=True
        acc += safe_div(rplane[p * ow + q], z) * inv;
      }
    r_in[i] = plane[xh * s.w + xw] * acc;
  }
}

// ---------------------------------------------------------------- BN / Add / ReLU
__global__ void bn_absratio_kernel(const float* __restrict__ x, const float* __restrict__ r_out, float* __restrict__ r_in,
                                   const float* __restrict__ mean, const float* __restrict__ var,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int c,
                                   int hw, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int ch = (int)((i / hw) % c);
    float sd = sqrtf(var[ch] + eps);
    float w = gamma[ch] / sd;
    float b = beta[ch] - (mean[ch] * gamma[ch]) / sd;
    float xw = fabsf(x[i] * w);
    r_in[i] = safe_div(xw, xw + fabsf(b)) * r_out[i];
  }
}

__global__ void add_split_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                 const float* __restrict__ r, float* __restrict__ r1, float* __restrict__ r2,
                                 size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float a = x1[i], b = x2[i], rr = r[i];
    float out = a + b;
    float half = (out == 0.f) ? 0.5f * rr : 0.f;
    float sg = (out > 0.f) ? 1.f : ((out < 0.f) ? -1.f : 0.f);
    out += LRPX_EPSILON * sg;
    float o1 = rr * a / out, o2 = rr * b / out;
    if (o1 != o1) o1 = 0.f;
    if (o2 != o2) o2 = 0.f;
    r1[i] = o1 + half;
    r2[i] = o2 + half;
  }
}

__global__ void relu_mask_kernel(const float* __restrict__ x, const float* __restrict__ r, float* __restrict__ o,
                                 size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    o[i] = x[i] > 0.f ? r[i] : 0.f;
}

// one warp per row
__global__ void normalize_rel_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int cols,
                                     float temperature) {
  int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * cols;
  float m = 0.f;
  for (int j = lane; j < cols; j += 32) m = fmaxf(m, fabsf(xr[j]));
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (m == 0.f) m = 1.f;
  float add = temperature > 1.f ? temperature : 1.f;
  for (int j = lane; j < cols; j += 32) y[(size_t)row * cols + j] = xr[j] / m * temperature + add;
}

__global__ void sum_f64_kernel(const float* __restrict__ x, size_t n, double* out) {
  double acc = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    acc += (double)x[i];
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double sm[32];
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sm[wid] = acc;
  __syncthreads();
  if (wid == 0) {
    acc = lane < (blockDim.x >> 5) ? sm[lane] : 0.0;
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicAdd(out, acc);
  }
}


// ---------------------------------------------------------------- named vector rules of the explainers
// lrp_linear_eps (gridTDmodel.py:744-765, aoamodel.py:785-810): R_in[j] = x_j * sum_k W[k][j] * r_k / stab(z_k),
// stab(z) = z + 0.01 sign z (0 -> 0.01); z = W x when the caller passes forward_output=False.
// (1) one warp per output row: t_k = r_k / stab(z_k), z_k recomputed when z == nullptr
__global__ void lin_eps_t_kernel(const float* __restrict__ r, const float* __restrict__ x, const float* __restrict__ z,
                                 const float* __restrict__ W, float* __restrict__ t, int n_out, int n_in) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (k >= n_out) return;
  float zk;
  if (z) {
    zk = z[k];
  } else {
    float acc = 0.f;
    for (int j = lane; j < n_in; j += 32) acc = fmaf(W[(size_t)k * n_in + j], x[j], acc);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    zk = acc;
  }
  if (lane == 0) t[k] = r[k] / stab(zk);
}
// (2) column sums over a slice of the rows (coalesced along j), deterministic: part[s][j] = sum_{k in slice s} W[k][j] t_k
__global__ void lin_eps_part_kernel(const float* __restrict__ W, const float* __restrict__ t, float* __restrict__ part,
                                    int n_out, int n_in, int rows_per_slice) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
  if (j >= n_in) return;
  const int k0 = s * rows_per_slice, k1 = min(n_out, k0 + rows_per_slice);
  float acc = 0.f;
  for (int k = k0; k < k1; ++k) acc = fmaf(W[(size_t)k * n_in + j], t[k], acc);
  part[(size_t)s * n_in + j] = acc;
}
__global__ void lin_eps_final_kernel(const float* __restrict__ part, const float* __restrict__ x, float* __restrict__ out,
                                     int n_in, int slices) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_in) return;
  float acc = 0.f;
  for (int s = 0; s < slices; ++s) acc += part[(size_t)s * n_in + j];
  out[j] = x[j] * acc;
}
static inline int lin_eps_slices(int n_out) {
  int s = (n_out + 63) / 64;
  return s < 1 ? 1 : (s > 128 ? 128 : s);
}

// lrp_mha (aoamodel.py:812-862): r_val[p][head*dk + d] = v[p][..] * alpha[head][p] * r_ctx[..] / stab(ctx[..]); other heads 0
__global__ void lrp_mha_kernel(const float* __restrict__ alpha, const float* __restrict__ value,
                               const float* __restrict__ r_ctx, const float* __restrict__ ctx, float* __restrict__ out,
                               int P, int H, int num_head, int head) {
  const int dk = H / num_head;
  const long long total = (long long)P * H;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i / H), c = (int)(i % H);
    float v = 0.f;
    if (c / dk == head) v = value[i] * alpha[(size_t)head * P + p] * r_ctx[c] / stab(ctx[c]);
    out[i] = v;
  }
}

static inline int grid_for(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static int pool_out_dims(const lrpx_pool_shape* s, int* oh, int* ow) {
  if (!s || s->n <= 0 || s->c <= 0 || s->h <= 0 || s->w <= 0 || s->kh <= 0 || s->kw <= 0 || s->stride_h <= 0 ||
      s->stride_w <= 0 || s->pad_h < 0 || s->pad_w < 0 || 2 * s->pad_h > s->kh || 2 * s->pad_w > s->kw)
    return -1;
  *oh = (s->h + 2 * s->pad_h - s->kh) / s->stride_h + 1;  // ceil_mode=False, dilation=1
  *ow = (s->w + 2 * s->pad_w - s->kw) / s->stride_w + 1;
  return (*oh > 0 && *ow > 0) ? 0 : -1;
}

}  // namespace lrpx

using namespace lrpx;

extern "C" {

const char* lrpx_last_error(void) { return lrpx::g_err; }
int lrpx_version(void) { return 100; }
int lrpx_device_cc(void) {
  int dev = 0, maj = 0, mnr = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { set_error("no CUDA device"); return LRPX_E_CUDA; }
  cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&mnr, cudaDevAttrComputeCapabilityMinor, dev);
  return maj * 10 + mnr;
}

int lrpx_maxpool_forward_f32(const float* x, float* y, int64_t* idx, const lrpx_pool_shape* shp, void* stream) {
  int oh, ow;
  LRPX_CHECK_ARG(x && pool_out_dims(shp, &oh, &ow) == 0, "bad pool shape");
  long long total = (long long)shp->n * shp->c * oh * ow;
  maxpool_fwd_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(x, y, idx, *shp, oh, ow, total);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

int lrpx_maxpool_wta_f32(const float* x, const float* r_out, float* r_in, const lrpx_pool_shape* shp, void* stream) {
  int oh, ow;
  LRPX_CHECK_ARG(x && r_out && r_in && pool_out_dims(shp, &oh, &ow) == 0, "bad pool shape");
  long long total = (long long)shp->n * shp->c * shp->h * shp->w;
  maxpool_wta_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(x, r_out, r_in, *shp, oh, ow, total);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

int lrpx_avgpool_prop_f32(const float* x, const float* r_out, float* r_in, const lrpx_pool_shape* shp, void* stream) {
  int oh, ow;
  LRPX_CHECK_ARG(x && r_out && r_in && pool_out_dims(shp, &oh, &ow) == 0, "bad pool shape");
  long long total = (long long)shp->n * shp->c * shp->h * shp->w;
  avgpool_prop_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(x, r_out, r_in, *shp, oh, ow, total);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

int lrpx_bn_absratio_f32(const float* x, const float* r_out, float* r_in, const float* running_mean,
                         const float* running_var, const float* gamma, const float* beta, float eps, int n, int c,
                         int hw, void* stream) {
  LRPX_CHECK_ARG(x && r_out && r_in && running_mean && running_var && gamma && beta && n > 0 && c > 0 && hw > 0,
                 "bad argument");
  long long total = (long long)n * c * hw;
  bn_absratio_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(x, r_out, r_in, running_mean, running_var, gamma,
                                                                     beta, eps, c, hw, total);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

int lrpx_add_split_f32(const float* x1, const float* x2, const float* r_out, float* r1, float* r2, size_t count,
                       void* stream) {
  LRPX_CHECK_ARG(x1 && x2 && r_out && r1 && r2, "null pointer");
  if (count == 0) return LRPX_OK;
  add_split_kernel<<<grid_for((long long)count), 256, 0, as_stream(stream)>>>(x1, x2, r_out, r1, r2, count);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

int lrpx_relu_mask_f32(const float* x, const float* r_out, float* r_in, size_t count, void* stream) {
  LRPX_CHECK_ARG(x && r_out && r_in, "null pointer");
  if (count == 0) return LRPX_OK;
  relu_mask_kernel<<<grid_for((long long)count), 256, 0, as_stream(stream)>>>(x, r_out, r_in, count);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

int lrpx_normalize_relevance_f32(const float* x, float* y, int rows, int cols, float temperature, void* stream) {
  LRPX_CHECK_ARG(x && y && rows >= 0 && cols > 0, "bad argument");
  if (rows == 0) return LRPX_OK;
  normalize_rel_kernel<<<ceil_div(rows, 8), 256, 0, as_stream(stream)>>>(x, y, rows, cols, temperature);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

int lrpx_sum_f64(const float* x, size_t count, double* out, void* stream) {
  LRPX_CHECK_ARG(x && out, "null pointer");
  cudaMemsetAsync(out, 0, sizeof(double), as_stream(stream));
  if (count == 0) return LRPX_OK;
  sum_f64_kernel<<<grid_for((long long)count), 256, 0, as_stream(stream)>>>(x, count, out);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

size_t lrpx_lrp_linear_eps_workspace_bytes(int n_out, int n_in) {
  if (n_out <= 0 || n_in <= 0) return 0;
  return ((size_t)n_out + (size_t)lin_eps_slices(n_out) * n_in) * sizeof(float);
}

int lrpx_lrp_linear_eps_f32(const float* r_out, const float* x, const float* z, const float* W, float* r_in, int n_out,
                            int n_in, void* workspace, size_t workspace_bytes, void* stream) {
  LRPX_CHECK_ARG(r_out && x && W && r_in && n_out > 0 && n_in > 0, "bad argument");
  LRPX_CHECK_ARG(workspace && workspace_bytes >= lrpx_lrp_linear_eps_workspace_bytes(n_out, n_in), "workspace too small");
  cudaStream_t st = as_stream(stream);
  float* t = reinterpret_cast<float*>(workspace);
  float* part = t + n_out;
  const int slices = lin_eps_slices(n_out), rps = (n_out + slices - 1) / slices;
  lin_eps_t_kernel<<<(n_out + 7) / 8, 256, 0, st>>>(r_out, x, z, W, t, n_out, n_in);
  LRPX_CHECK_LAUNCH();
  lin_eps_part_kernel<<<dim3((n_in + 127) / 128, slices), 128, 0, st>>>(W, t, part, n_out, n_in, rps);
  LRPX_CHECK_LAUNCH();
  lin_eps_final_kernel<<<(n_in + 127) / 128, 128, 0, st>>>(part, x, r_in, n_in, slices);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

int lrpx_lrp_mha_f32(const float* alpha, const float* value, const float* r_context, const float* context, float* r_value,
                     int P, int H, int num_head, int head_idx, void* stream) {
  LRPX_CHECK_ARG(alpha && value && r_context && context && r_value && P > 0 && H > 0 && num_head > 0 && H % num_head == 0 &&
                     head_idx >= 0 && head_idx < num_head, "bad argument");
  lrp_mha_kernel<<<grid_for((long long)P * H), 256, 0, as_stream(stream)>>>(alpha, value, r_context, context, r_value, P, H,
                                                                           num_head, head_idx);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

}  // extern "C"
