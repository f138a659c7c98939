// Layout / pooling helpers of the tensor-core path (NHWC bf16), all HBM-bound element-wise kernels.
#include "lrpx_common.cuh"

namespace lrpx {

// fp32 (cout,cin,kh,kw) -> bf16 GEMM operand, see lrpx_weight_prep_bf16 in lrpx.h
__global__ void weight_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wt, int cout, int cin,
                                   int kh, int kw, int mode, int rows_pad, int chan_pad) {
  // output (rows_pad, kh*kw, chan_pad)
  long long total = (long long)rows_pad * kh * kw * chan_pad;
  bool transposed = mode >= 2;
  int rows = transposed ? cin : cout, chans = transposed ? cout : cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % chan_pad);
    int rs = (int)((i / chan_pad) % (kh * kw));
    int row = (int)(i / ((long long)chan_pad * kh * kw));
    float v = 0.f;
    if (row < rows && c < chans) {
      int r = rs / kw, s = rs % kw;
      if (!transposed) {
        v = w[(((size_t)row * cin + c) * kh + r) * kw + s];
        if (mode == 1) v = fmaxf(v, 0.f);
      } else {
        v = w[(((size_t)c * cin + row) * kh + (kh - 1 - r)) * kw + (kw - 1 - s)];
        v = (mode == 2) ? fmaxf(v, 0.f) : fminf(v, 0.f);
      }
    }
    wt[i] = __float2bfloat16(v);
  }
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int n, int c, int hw,
                                    int c_pad) {
  long long total = (long long)n * hw * c_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int ch = (int)(i % c_pad);
    long long pix = i / c_pad;
    int img = (int)(pix / hw), p = (int)(pix % hw);
    float v = ch < c ? src[((size_t)img * c + ch) * hw + p] : 0.f;
    dst[i] = __float2bfloat16(v);
  }
}

__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int n, int c, int hw,
                                    int c_pad) {
  long long total = (long long)n * c * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int p = (int)(i % hw);
    int ch = (int)((i / hw) % c);
    int img = (int)(i / ((long long)hw * c));
    dst[i] = __bfloat162float(src[((size_t)img * hw + p) * c_pad + ch]);
  }
}

// 2x2/2 max-pool over NHWC bf16, 8 channels (16 B) per thread.  Scan order (0,0),(0,1),(1,0),(1,1),
// strict '>' so the first maximum wins, NaN wins (PyTorch max_pool2d_with_indices).
__global__ void maxpool2_nhwc_kernel(const uint4* __restrict__ act, const uint4* __restrict__ gain,
                                     uint4* __restrict__ pooled, uint2* __restrict__ idx, uint4* __restrict__ gpool,
                                     int n, int h, int w, int c8) {
  int oh = h / 2, ow = w / 2;
  long long total = (long long)n * oh * ow * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int cc = (int)(i % c8);
    long long pix = i / c8;
    int q = (int)(pix % ow);
    int p = (int)((pix / ow) % oh);
    int img = (int)(pix / ((long long)ow * oh));
    size_t base = (((size_t)img * h + 2 * p) * w + 2 * q) * c8 + cc;
    size_t off[4] = {base, base + c8, base + (size_t)w * c8, base + (size_t)w * c8 + c8};
    uint4 v[4], g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = act[off[k]];
      if (gain) g[k] = gain[off[k]];
    }
    uint4 outv, outg;
    unsigned char bi[8];
    const __nv_bfloat16* vb[4] = {(const __nv_bfloat16*)&v[0], (const __nv_bfloat16*)&v[1],
                                  (const __nv_bfloat16*)&v[2], (const __nv_bfloat16*)&v[3]};
    const __nv_bfloat16* gb[4] = {(const __nv_bfloat16*)&g[0], (const __nv_bfloat16*)&g[1],
                                  (const __nv_bfloat16*)&g[2], (const __nv_bfloat16*)&g[3]};
    __nv_bfloat16* ov = (__nv_bfloat16*)&outv;
    __nv_bfloat16* og = (__nv_bfloat16*)&outg;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float best = -INFINITY;
      int b = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float f = __bfloat162float(vb[k][e]);
        if (f > best || f != f) { best = f; b = k; }
      }
      bi[e] = (unsigned char)b;
      ov[e] = vb[b][e];
      if (gain) og[e] = gb[b][e];
    }
    pooled[i] = outv;
    if (idx) {
      uint2 pk;
      pk.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
      pk.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
      idx[i] = pk;
    }
    if (gain) gpool[i] = outg;
  }
}

__global__ void scale_rows_kernel(const float* __restrict__ r, const __nv_bfloat16* __restrict__ gain,
                                  const int32_t* __restrict__ row_img, __nv_bfloat16* __restrict__ out, int hw, int c,
                                  long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int ch = (int)(i % c);
    long long pix = i / c;
    int e = (int)(pix / hw), p = (int)(pix % hw);
    int img = row_img ? row_img[e] : e;
    float g = __bfloat162float(gain[((size_t)img * hw + p) * c + ch]);
    out[i] = __float2bfloat16(r[i] * g);
  }
}

static inline int grid_for(long long total) {
  long long g = (total + 255) / 256, cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace lrpx

using namespace lrpx;

extern "C" {

int lrpx_weight_prep_bf16(const float* w, void* wt, int cout, int cin, int kh, int kw, int mode, int rows_pad,
                          int chan_pad, void* stream) {
  LRPX_CHECK_ARG(w && wt && cout > 0 && cin > 0 && kh > 0 && kw > 0 && mode >= 0 && mode <= 3, "bad argument");
  int rows = mode >= 2 ? cin : cout, chans = mode >= 2 ? cout : cin;
  LRPX_CHECK_ARG(rows_pad >= rows && chan_pad >= chans, "padding smaller than the source");
  long long total = (long long)rows_pad * kh * kw * chan_pad;
  weight_prep_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(w, (__nv_bfloat16*)wt, cout, cin, kh, kw, mode,
                                                                    rows_pad, chan_pad);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

int lrpx_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int n, int c, int h, int w, int c_pad, void* stream) {
  LRPX_CHECK_ARG(src && dst && n > 0 && c > 0 && h > 0 && w > 0 && c_pad >= c, "bad argument");
  long long total = (long long)n * h * w * c_pad;
  nchw_to_nhwc_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(src, (__nv_bfloat16*)dst, n, c, h * w, c_pad);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

int lrpx_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int n, int c, int h, int w, int c_pad, void* stream) {
  LRPX_CHECK_ARG(src && dst && n > 0 && c > 0 && h > 0 && w > 0 && c_pad >= c, "bad argument");
  long long total = (long long)n * h * w * c;
  nhwc_to_nchw_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)src, dst, n, c, h * w, c_pad);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

int lrpx_tc_maxpool2_bf16(const void* act, const void* gain_fine, void* pooled, uint8_t* idx, void* gain_pooled, int n,
                          int h, int w, int c, void* stream) {
  LRPX_CHECK_ARG(act && pooled && n > 0 && h > 0 && w > 0 && c > 0 && (h % 2) == 0 && (w % 2) == 0 && (c % 8) == 0,
                 "bad argument (h, w even and c % 8 == 0 required)");
  LRPX_CHECK_ARG((gain_fine == nullptr) == (gain_pooled == nullptr), "gain_fine and gain_pooled go together");
  long long total = (long long)n * (h / 2) * (w / 2) * (c / 8);
  maxpool2_nhwc_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>((const uint4*)act, (const uint4*)gain_fine,
                                                                      (uint4*)pooled, (uint2*)idx, (uint4*)gain_pooled,
                                                                      n, h, w, c / 8);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

int lrpx_tc_scale_rows(const float* r, const void* gain, const int32_t* row_img, void* out, int n_expl, int hw, int c,
                       void* stream) {
  LRPX_CHECK_ARG(r && gain && out && n_expl > 0 && hw > 0 && c > 0, "bad argument");
  long long total = (long long)n_expl * hw * c;
  scale_rows_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(r, (const __nv_bfloat16*)gain, row_img,
                                                                   (__nv_bfloat16*)out, hw, c, total);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

}  // extern "C"
