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

// ------------------------------------------------------------------------------------------------
// Bounding-box correctness (evaluation.py:310-335, :403-405, :425-431), batched over requests: the share of the
// (thresholded) positive relevance that falls inside a box.  One block per request; the normalised map lives in shared
// memory, so the heat-map is read from HBM once (algorithmic bytes = the heat-map).
// ------------------------------------------------------------------------------------------------
constexpr int BBOX_MAX_BOXES = 8;

__device__ __forceinline__ float abl_block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  v = 0.f;
  for (int w = 0; w < ABL_THREADS / 32; ++w) v += red[w];
  return v;
}

__global__ void __launch_bounds__(ABL_THREADS) bbox_ratio_kernel(lrpx_bbox_args a) {
  extern __shared__ float s_map[];                 // [H*W] max-abs-normalised positive relevance
  __shared__ float s_red[ABL_THREADS / 32];
  const int q = blockIdx.x, hw = a.H * a.W;
  const float* heat = a.heat + (size_t)q * a.C * hw;
  // np.mean(np.maximum(sign * relevance, 0), axis=(0, 1))   (:398-404)
  float mx = 0.f;
  for (int p = threadIdx.x; p < hw; p += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < a.C; ++c) acc += fmaxf(a.sign * heat[(size_t)c * hw + p], 0.f);
    acc /= (float)a.C;
    s_map[p] = acc;
    mx = fmaxf(mx, acc);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = s_red[0];
  for (int w = 1; w < ABL_THREADS / 32; ++w) mx = fmaxf(mx, s_red[w]);
  __syncthreads();
  // _project_maxabs (:337-342): all zeros when the map is zero
  for (int p = threadIdx.x; p < hw; p += blockDim.x) s_map[p] = mx > 0.f ? s_map[p] / mx : 0.f;
  __syncthreads();
  const int nb = min(a.n_boxes ? a.n_boxes[q] : a.max_boxes, a.max_boxes);
  __shared__ int bx[BBOX_MAX_BOXES][4];
  if (threadIdx.x < BBOX_MAX_BOXES * 4)
    bx[threadIdx.x >> 2][threadIdx.x & 3] =
        (int)(threadIdx.x >> 2) < nb ? a.boxes[((size_t)q * a.max_boxes + (threadIdx.x >> 2)) * 4 + (threadIdx.x & 3)] : 0;
  __syncthreads();
  // Loop order of bbox_experiment (:419-431): boxes outer, thresholds inner.  _calculate_overlaped_pixels zeroes the
  // map IN PLACE (:324-326), so with `inplace_quirk` every later (box, threshold) pair sees the map already cut at the
  // largest threshold applied before it — the effective threshold is the running maximum in loop order.
  float eff = -INFINITY;
  for (int jb = 0; jb < nb; ++jb) {
    const int x0 = bx[jb][0], y0 = bx[jb][1], x1 = bx[jb][2], y1 = bx[jb][3];
    for (int t = 0; t < a.n_thr; ++t) {
      const float thr = a.thresholds[t];
      eff = a.inplace_quirk ? fmaxf(eff, thr) : thr;
      float tot = 0.f, cor = 0.f;
      for (int p = threadIdx.x; p < hw; p += blockDim.x) {
        const float m = s_map[p];
        const float v = m <= eff ? 0.f : m;                        // relevance[relevance <= threshold] = 0 (:324-326)
        const int y = p / a.W, x = p - y * a.W;
        tot += v;
        if (x >= x0 && x < x1 && y >= y0 && y < y1) cor += v;       // bbox_mask[y0:y1, x0:x1] = 1 (:321-322)
      }
      tot = abl_block_sum(tot, s_red);
      cor = abl_block_sum(cor, s_red);
      if (threadIdx.x == 0) {
        const float r = tot == 0.f ? 0.f : cor / tot;               // :327-334
        a.ratio[((size_t)q * a.max_boxes + jb) * a.n_thr + t] = r > 1.f ? 1.f : r;
      }
    }
  }
  // unused box slots (contiguous behind the used ones) report 0
  for (int k = threadIdx.x; k < (a.max_boxes - nb) * a.n_thr; k += blockDim.x)
    a.ratio[((size_t)q * a.max_boxes + nb) * a.n_thr + k] = 0.f;
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

extern "C" int lrpx_bbox_ratio_f32(const lrpx_bbox_args* a, void* stream) {
  LRPX_CHECK_ARG(a, "null args");
  LRPX_CHECK_ARG(a->Q >= 0 && a->C > 0 && a->H > 0 && a->W > 0 && a->n_thr > 0 && a->n_thr <= ABL_THREADS &&
                     a->max_boxes > 0 && a->max_boxes <= BBOX_MAX_BOXES,
                 "bad shape (at most 8 boxes per request)");
  const size_t smem = (size_t)a->H * a->W * sizeof(float);
  LRPX_CHECK_ARG(smem <= 200 * 1024, "H * W too large for the shared-memory map (at most 51200 pixels)");
  if (a->Q == 0) return LRPX_OK;
  LRPX_CHECK_ARG(a->heat && a->thresholds && a->boxes && a->ratio, "null pointer");
  cudaError_t e = cudaFuncSetAttribute(bbox_ratio_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) {
    set_error("lrpx_bbox_ratio_f32: %s", cudaGetErrorString(e));
    return LRPX_E_CUDA;
  }
  bbox_ratio_kernel<<<a->Q, ABL_THREADS, smem, as_stream(stream)>>>(*a);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}
