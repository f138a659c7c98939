// Beam-search bookkeeping on the device: one step of GridTDModel.beam_search
// (models/gridTDmodel.py:400-478; the AoA / adaptive models run the same loop, aoamodel.py:405-485,
// adaptiveattention.py:370-447) without the per-step host round trips (.topk -> python lists -> index tensors).
//
// The reference shrinks its beam whenever a sequence emits <end>.  Here every image keeps k slots; the first
// n_alive slots are the unfinished beams in the reference's order and the others are dead (their rows of the step
// kernels compute values nobody reads).  One step:
//   scores[r][w] = score[r] + log_softmax(logits[r])[w]        over the alive rows (row 0 only at step 0, :436-439)
//   the n_alive best (row, word) pairs in descending order, ties to the lower flat index          (:440 topk)
//   pairs ending in <end> move to the completed list in selection order, the others become the new alive slots
//   in selection order                                                                            (:446-462)
// and `src_row` tells lrpx_beam_gather_f32 which old row each new slot continues (state[beam_idx], :459).
#include "lrpx_common.cuh"

namespace lrpx {

constexpr int BEAM_MAX_K = 8;
constexpr int BEAM_MAX_W = 64;        // L + 1 <= 64 tokens per sequence
constexpr int BEAM_THREADS = 256;

struct Cand {
  float s;
  int i;
};
__device__ __forceinline__ bool better(const Cand& a, const Cand& b) { return a.s > b.s || (a.s == b.s && a.i < b.i); }

__device__ __forceinline__ Cand block_best(Cand c, Cand* red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    Cand t{__shfl_xor_sync(0xffffffffu, c.s, o), __shfl_xor_sync(0xffffffffu, c.i, o)};
    if (better(t, c)) c = t;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = c;
  __syncthreads();
  c = red[0];
  for (int w = 1; w < BEAM_THREADS / 32; ++w)
    if (better(red[w], c)) c = red[w];
  return c;
}
__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  v = red[0];
  for (int w = 1; w < BEAM_THREADS / 32; ++w) v = fmaxf(v, red[w]);
  return v;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  v = 0.f;
  for (int w = 0; w < BEAM_THREADS / 32; ++w) v += red[w];
  return v;
}

__global__ void __launch_bounds__(BEAM_THREADS) beam_step_kernel(lrpx_beam_args a) {
  const int b = blockIdx.x, k = a.k, V = a.V, W = a.L + 1;
  __shared__ float s_m[BEAM_MAX_K], s_lg[BEAM_MAX_K], s_sc[BEAM_MAX_K];
  __shared__ Cand s_sel[BEAM_MAX_K];
  __shared__ Cand s_red[BEAM_THREADS / 32];
  __shared__ float s_redf[BEAM_THREADS / 32];
  __shared__ int s_old[BEAM_MAX_K * BEAM_MAX_W];
  __shared__ int s_dst[BEAM_MAX_K];            // selection j -> new alive slot, or -(completed slot + 1)
  const int na = a.n_alive[b];
  int32_t* src = a.src_row + (size_t)b * k;
  if (na == 0) {                               // every beam of this image has ended: the rows stay where they are
    for (int j = threadIdx.x; j < k; j += blockDim.x) src[j] = b * k + j;
    return;
  }
  const int rows = a.step == 0 ? 1 : na;       // the k start beams are identical: only row 0 competes (:436-437)
  const int len = a.step + 1;                  // tokens per alive sequence so far, <start> included
  const float* lg = a.logits + (size_t)b * k * V;
  for (int r = 0; r < rows; ++r) {             // log-softmax statistics of the alive rows
    const float* x = lg + (size_t)r * V;
    float m = -INFINITY;
    for (int w = threadIdx.x; w < V; w += blockDim.x) m = fmaxf(m, x[w]);
    m = block_max(m, s_redf);
    float s = 0.f;
    for (int w = threadIdx.x; w < V; w += blockDim.x) s += expf(x[w] - m);
    s = block_sum(s, s_redf);
    if (threadIdx.x == 0) { s_m[r] = m; s_lg[r] = logf(s); s_sc[r] = a.scores[(size_t)b * k + r]; }
  }
  for (int t = threadIdx.x; t < na * len; t += blockDim.x)
    s_old[(t / len) * BEAM_MAX_W + t % len] = a.seqs[((size_t)b * k + t / len) * W + t % len];
  __syncthreads();
  // the na best candidates, one block-wide arg-max per rank
  Cand prev{INFINITY, -1};
  for (int j = 0; j < na; ++j) {
    Cand best{-INFINITY, 0x7fffffff};
    for (int r = 0; r < rows; ++r) {
      const float* x = lg + (size_t)r * V;
      const float m = s_m[r], l = s_lg[r], sc = s_sc[r];
      for (int w = threadIdx.x; w < V; w += blockDim.x) {
        Cand c{((x[w] - m) - l) + sc, r * V + w};          // log_softmax = (x - max) - log(sum), then + running score
        const bool after_prev = c.s < prev.s || (c.s == prev.s && c.i > prev.i);
        if (after_prev && better(c, best)) best = c;
      }
    }
    best = block_best(best, s_red);
    // no candidate left (NaN scores compare false everywhere): keep the index inside [0, rows*V) so that the
    // bookkeeping below never leaves its arrays; the score stays -inf
    if (best.i < 0 || best.i >= rows * V) best.i = 0;
    if (threadIdx.x == 0) s_sel[j] = best;
    prev = best;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int alive = 0, nc = a.n_comp[b];
    for (int j = 0; j < na; ++j) {
      const int word = s_sel[j].i % V;
      if (word == a.end_id) {
        s_dst[j] = -(nc + 1);
        a.comp_scores[(size_t)b * k + nc] = s_sel[j].s;
        a.comp_len[(size_t)b * k + nc] = len + 1;
        ++nc;
      } else {
        s_dst[j] = alive;
        a.scores[(size_t)b * k + alive] = s_sel[j].s;
        a.prev_words[(size_t)b * k + alive] = word;
        src[alive] = b * k + s_sel[j].i / V;
        ++alive;
      }
    }
    for (int j = alive; j < k; ++j) src[j] = b * k + j;
    a.n_alive[b] = alive;
    a.n_comp[b] = nc;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < na * (len + 1); t += blockDim.x) {
    const int j = t / (len + 1), p = t % (len + 1);
    const int r = s_sel[j].i / V;
    const int tok = p < len ? s_old[r * BEAM_MAX_W + p] : s_sel[j].i % V;
    const int d = s_dst[j];
    int32_t* dst = d >= 0 ? a.seqs + ((size_t)b * k + d) * W : a.comp_seqs + ((size_t)b * k + (-d - 1)) * W;
    dst[p] = tok;
  }
}

__global__ void beam_gather_kernel(lrpx_beam_gather_args g) {
  const int row = blockIdx.x;
  const int s = g.src_row[row];
  for (int e = 0; e < g.n_pairs; ++e) {
    const float* from = g.src[e] + (size_t)s * g.ld_src[e];
    float* to = g.dst[e] + (size_t)row * g.ld_dst[e];
    for (int c = threadIdx.x; c < g.width[e]; c += blockDim.x) to[c] = from[c];
  }
}

}  // namespace lrpx

using namespace lrpx;

extern "C" {

int lrpx_beam_step(const lrpx_beam_args* a, void* stream) {
  LRPX_CHECK_ARG(a, "null args");
  LRPX_CHECK_ARG(a->B > 0 && a->k > 0 && a->k <= BEAM_MAX_K && a->V >= a->k && a->L > 0 && a->L + 1 <= BEAM_MAX_W &&
                     a->step >= 0 && a->step < a->L,
                 "bad dimensions (k <= 8, L + 1 <= 64, 0 <= step < L)");
  LRPX_CHECK_ARG(a->logits && a->scores && a->n_alive && a->seqs && a->comp_seqs && a->comp_len && a->comp_scores &&
                     a->n_comp && a->prev_words && a->src_row,
                 "null pointer in args");
  beam_step_kernel<<<a->B, BEAM_THREADS, 0, as_stream(stream)>>>(*a);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

int lrpx_beam_gather_f32(const lrpx_beam_gather_args* g, void* stream) {
  LRPX_CHECK_ARG(g, "null args");
  LRPX_CHECK_ARG(g->n_rows >= 0 && g->n_pairs > 0 && g->n_pairs <= LRPX_BEAM_GATHER_MAX && g->src_row, "bad argument");
  for (int e = 0; e < g->n_pairs; ++e)
    LRPX_CHECK_ARG(g->dst[e] && g->src[e] && g->dst[e] != g->src[e] && g->width[e] > 0, "bad pair (dst != src required)");
  if (g->n_rows == 0) return LRPX_OK;
  beam_gather_kernel<<<g->n_rows, 128, 0, as_stream(stream)>>>(*g);
  LRPX_CHECK_LAUNCH();
  return LRPX_OK;
}

}  // extern "C"
