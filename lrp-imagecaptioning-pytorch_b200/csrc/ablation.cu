// Patch-ablation masks: EvaluationExperiments.block_image, evaluation.py:57-80, batched over requests.
//   spatial relevance = mean over the channels of the heat-map (evaluation.py:128-129; the mean's 1/C does not change
//   the ranking, so the channel SUM is ranked), patch sums over patch x patch pixels, the k patches with the largest
//   sums are blanked.  One block per request: HBM-bound, algorithmic bytes = one read of the heat-map and of the image,
//   one write of the masked image (+ the mask).
#include "lrpx_common.cuh"

namespace lrpx {

constexpr int ABL_THREADS = 256;
constexpr int ABL_MAX_PATCHES = 4096;

__global__ void __launch_bounds__(ABL_THREADS) block_image_kernel(lrpx_block_image_args a) {
  extern __shared__ float s_sum[];                 // [np] patch sums (-inf = blanked) | [C][np] channel partials
  __shared__ float s_rv[ABL_THREADS / 32];
  __shared__ int s_ri[ABL_THREADS / 32];
  const int q = blockIdx.x;
  const int H = a.H, W = a.W, C = a.C, ps = a.patch;
  const int nh = H / ps, nw = W / ps, np = nh * nw;
  const size_t hw = (size_t)H * W;
  const float* heat = a.heat + (size_t)q * C * hw;
  const bool vec4 = (ps % 4 == 0) && (W % 4 == 0) &&
                    ((reinterpret_cast<uintptr_t>(a.heat) | reinterpret_cast<uintptr_t>(a.images) |
                      reinterpret_cast<uintptr_t>(a.mask) | reinterpret_cast<uintptr_t>(a.masked)) & 15) == 0;
  // patch sums: one (patch, channel) item per thread step, fixed summation order -> deterministic; the channel
  // partials are added in channel order afterwards
  float* s_part = s_sum + np;                      // [C][np]
  for (int it = threadIdx.x; it < np * C; it += blockDim.x) {
    const int c = it / np, p = it - c * np;
    const int py = p / nw, px = p - py * nw;
    float acc = 0.f;
    for (int y = 0; y < ps; ++y) {
      const float* row = heat + c * hw + (size_t)(py * ps + y) * W + px * ps;
      if (vec4) {                                    // 16-byte loads: patch rows start on 16-byte boundaries
        for (int x = 0; x < ps; x += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(row + x));
          acc += v.x; acc += v.y; acc += v.z; acc += v.w;
        }
      } else {
        for (int x = 0; x < ps; ++x) acc += row[x];
      }
    }
    s_part[it] = acc;
  }
  __syncthreads();
  for (int p = threadIdx.x; p < np; p += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc += s_part[c * np + p];
    s_sum[p] = acc;
  }
  __syncthreads();
  // the k largest sums, ties to the lower patch index; a selected patch becomes -inf
  for (int j = 0; j < a.k; ++j) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int p = threadIdx.x; p < np; p += blockDim.x) {
      const float v = s_sum[p];
      if (v > bv || (v == bv && p < bi && v != -INFINITY)) { bv = v; bi = p; }
    }
    for (int o = 16; o; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_rv[threadIdx.x >> 5] = bv; s_ri[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < ABL_THREADS / 32; ++w)
        if (s_rv[w] > bv || (s_rv[w] == bv && s_ri[w] < bi)) { bv = s_rv[w]; bi = s_ri[w]; }
      if (bi < np) s_sum[bi] = -INFINITY;
    }
    __syncthreads();
  }
  // mask and masked image (evaluation.py:73-80, :130)
  const int img = a.req_img ? a.req_img[q] : q;
  const float* im = a.images ? a.images + (size_t)img * a.img_c * hw : nullptr;
  float* mk = a.mask ? a.mask + (size_t)q * hw : nullptr;
  float* out = a.masked ? a.masked + (size_t)q * a.img_c * hw : nullptr;
  if (vec4) {                                      // four pixels of one patch row per step
    const int w4 = W / 4;
    for (int i = threadIdx.x; i < H * w4; i += blockDim.x) {
      const int y = i / w4, x = (i - y * w4) * 4;
      const float keep = s_sum[(y / ps) * nw + x / ps] == -INFINITY ? 0.f : 1.f;
      const size_t o = (size_t)y * W + x;
      if (mk) *reinterpret_cast<float4*>(mk + o) = make_float4(keep, keep, keep, keep);
      if (out)
        for (int c = 0; c < a.img_c; ++c) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(im + c * hw + o));
          *reinterpret_cast<float4*>(out + c * hw + o) = make_float4(keep * v.x, keep * v.y, keep * v.z, keep * v.w);
        }
    }
    return;
  }
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    const int y = i / W, x = i - y * W;
    const float keep = s_sum[(y / ps) * nw + x / ps] == -INFINITY ? 0.f : 1.f;
    if (mk) mk[i] = keep;
    if (out)
      for (int c = 0; c < a.img_c; ++c) out[c * hw + i] = keep * im[c * hw + i];
  }
}

}  // namespace lrpx

using namespace lrpx;

extern "C" int lrpx_block_image_f32(const lrpx_block_image_args* a, void* stream) {
  LRPX_CHECK_ARG(a, "null args");
  LRPX_CHECK_ARG(a->Q >= 0 && a->C > 0 && a->H > 0 && a->W > 0 && a->patch > 0 && a->H % a->patch == 0 &&
                     a->W % a->patch == 0,
                 "bad shape (H and W must be multiples of the patch size, evaluation.py:59-60)");
  const int np = (a->H / a->patch) * (a->W / a->patch);
  LRPX_CHECK_ARG(np <= ABL_MAX_PATCHES && (size_t)np * (1 + a->C) * sizeof(float) <= 48 * 1024 && a->k >= 0 && a->k <= np, "k must not exceed the patch count (evaluation.py:65)");
  if (a->Q == 0) return LRPX_OK;
  LRPX_CHECK_ARG(a->heat && (a->mask || a->masked), "null pointer");
  LRPX_CHECK_ARG(!a->masked || (a->images && a->img_c > 0), "masked output needs the images");
  block_image_kernel<<<a->Q, ABL_THREADS, (size_t)np * (1 + a->C) * sizeof(float), as_stream(stream)>>>(*a);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}
