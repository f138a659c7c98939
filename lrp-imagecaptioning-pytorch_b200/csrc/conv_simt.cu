// fp32 implicit-GEMM convolution rules on CUDA cores (parity path, NCHW, any stride/pad/dilation).
//
//   K1  lrpx_conv_rule_s_f32   : z = net(a);  s = r_out / guard(z)      lrp_modules.py:81-84,111-114, utils.py:16-18,26-27
//   K2  lrpx_conv_rule_rin_f32 : r_in (+)= scale * a (.) net^T(s)        utils.py:29-30, lrp_modules.py:134-146
//   fwd lrpx_conv_forward_f32  : conv + bias (+ReLU)                     lrp_wrapper.py:70
//
// GEMM view.  fprop: M = n*P*Q output pixels, N = cout, K = cin*kh*kw.
//             dgrad: M = n*H*W input  pixels, N = cin,  K = cout*kh*kw (gather form, no atomics).
// The sign split of the alpha-beta rule never materialises W+/W-/a+/a- in HBM: a product
// a*w belongs to the positive net iff sign(a) == sign(w) (for the pos-net), so both clamped weight tiles
// are staged in shared memory and the per-(pixel,channel) sign of `a` selects between them.
#include "lrpx_common.cuh"

namespace lrpx {

constexpr int BM = 128, BN = 64, BK = 8, NT = 256, TM = 8, TN = 4;

enum { EPI_S = 0, EPI_FWD = 1, EPI_RIN = 2 };

struct ConvParams {
  const float* a;      // input activations (n,cin,h,w)
  const float* w;      // (cout,cin,kh,kw)
  const float* bias;   // cout or null
  const float* r;      // fprop: r_out (n,cout,P,Q);  dgrad: s (n,cout,P,Q)
  float* out;          // fprop: s or forward output; dgrad: r_in
  float* z_out;        // optional
  lrpx_conv_shape s;
  int P, Q;            // output spatial size
  int net;             // LRPX_NET_*
  float scale;
  int accumulate;
  int relu;
};

template <bool SIGNED, bool DGRAD, int EPI>
__global__ void __launch_bounds__(NT) conv_igemm_kernel(ConvParams p) {
  __shared__ float As[BK][BM];
  __shared__ float Bp[BK][BN];
  __shared__ float Bn[SIGNED ? BK : 1][BN];

  const lrpx_conv_shape& s = p.s;
  const int RS = s.kh * s.kw;
  const int M = DGRAD ? s.n * s.h * s.w : s.n * p.P * p.Q;
  const int N = DGRAD ? s.cin : s.cout;
  const int K = (DGRAD ? s.cout : s.cin) * RS;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int tm = tid & 15, tn = tid >> 4;

  // ---- A-tile loader coordinates: this thread always loads row (tid % BM), k = tid / BM + 2*j
  const int lm = tid % BM, lk0 = tid / BM;
  const int gm = m0 + lm;
  int an = 0, ay = 0, ax = 0;  // image, row, col of the GEMM row (output pixel for fprop, input pixel for dgrad)
  const bool row_ok = gm < M;
  {
    int hw = DGRAD ? s.h * s.w : p.P * p.Q;
    int wd = DGRAD ? s.w : p.Q;
    int g = row_ok ? gm : 0;
    an = g / hw;
    int rem = g % hw;
    ay = rem / wd;
    ax = rem % wd;
  }
  // ---- B-tile loader: 2 elements per thread
  const int bn_l = tid / BK;        // 0..31  (+32)
  const int bk_l = tid % BK;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // dgrad, signed: per-output sign predicate (a >= 0) and value of a for the epilogue
  float aval[DGRAD ? TM : 1][DGRAD ? TN : 1];
  if (DGRAD) {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      int m = m0 + tm + 16 * i;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        int n = n0 + tn * TN + j;
        float v = 0.f;
        if (m < M && n < N) {
          int hw = s.h * s.w;
          int img = m / hw, rem = m % hw;
          v = p.a[((size_t)img * s.cin + n) * hw + rem];
        }
        aval[DGRAD ? i : 0][DGRAD ? j : 0] = v;
      }
    }
  }
  const bool swap_nets = (p.net == LRPX_NET_NEG);

  for (int k0 = 0; k0 < K; k0 += BK) {
    // ---------------- load A tile
#pragma unroll
    for (int j = 0; j < BK / 2; ++j) {
      int kl = lk0 + 2 * j;
      int kk = k0 + kl;
      float v = 0.f;
      if (row_ok && kk < K) {
        int c = kk / RS, rs = kk % RS;
        int r = rs / s.kw, t = rs % s.kw;
        if (!DGRAD) {
          int ih = ay * s.stride_h - s.pad_h + r * s.dil_h;
          int iw = ax * s.stride_w - s.pad_w + t * s.dil_w;
          if (ih >= 0 && ih < s.h && iw >= 0 && iw < s.w) {
            v = p.a[(((size_t)an * s.cin + c) * s.h + ih) * s.w + iw];
            if (!SIGNED && EPI == EPI_S && v == 0.f) v = LRPX_RELEVANCE_RECT;  // Linear rule, lrp_modules.py:14
          }
        } else {
          int ph = ay + s.pad_h - r * s.dil_h, pw = ax + s.pad_w - t * s.dil_w;
          if (ph >= 0 && pw >= 0 && ph % s.stride_h == 0 && pw % s.stride_w == 0) {
            ph /= s.stride_h;
            pw /= s.stride_w;
            if (ph < p.P && pw < p.Q) v = p.r[(((size_t)an * s.cout + c) * p.P + ph) * p.Q + pw];
          }
        }
      }
      As[kl][lm] = v;
    }
    // ---------------- load B tile(s)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      int nl = bn_l + 32 * j;
      int n = n0 + nl, kk = k0 + bk_l;
      float v = 0.f;
      if (n < N && kk < K) {
        if (!DGRAD) {
          v = p.w[(size_t)n * K + kk];
        } else {
          int c = kk / RS, rs = kk % RS;  // c = cout index
          v = p.w[((size_t)c * s.cin + n) * RS + rs];
        }
      }
      if (SIGNED) {
        float vp = fmaxf(v, 0.f), vn = fminf(v, 0.f);
        Bp[bk_l][nl] = swap_nets ? vn : vp;
        Bn[SIGNED ? bk_l : 0][nl] = swap_nets ? vp : vn;
      } else {
        Bp[bk_l][nl] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bp[TN], bn[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) av[i] = As[kk][tm + 16 * i];
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        bp[j] = Bp[kk][tn * TN + j];
        if (SIGNED) bn[j] = Bn[SIGNED ? kk : 0][tn * TN + j];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          float wsel;
          if (SIGNED) {
            bool pos = DGRAD ? (aval[DGRAD ? i : 0][DGRAD ? j : 0] >= 0.f) : (av[i] >= 0.f);
            wsel = pos ? bp[j] : bn[j];
          } else {
            wsel = bp[j];
          }
          acc[i][j] = fmaf(av[i], wsel, acc[i][j]);
        }
    }
    __syncthreads();
  }

  // ---------------- epilogue
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + tm + 16 * i;
    if (m >= M) continue;
    int hw = DGRAD ? s.h * s.w : p.P * p.Q;
    int img = m / hw, rem = m % hw;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tn * TN + j;
      if (n >= N) continue;
      size_t o = ((size_t)img * N + n) * hw + rem;
      float v = acc[i][j];
      if (EPI == EPI_S) {
        if (p.bias) v += p.bias[n];
        if (p.z_out) p.z_out[o] = v;
        float rr = p.r[o];
        p.out[o] = SIGNED ? safe_div(rr, v) : rr / (p.bias ? v : stab(v));  // lrp_modules.py:18-23
      } else if (EPI == EPI_FWD) {
        if (p.bias) v += p.bias[n];
        if (p.relu) v = fmaxf(v, 0.f);
        p.out[o] = v;
      } else {
        float av = aval[DGRAD ? i : 0][DGRAD ? j : 0];
        if (!SIGNED && av == 0.f) av = LRPX_RELEVANCE_RECT;
        float res = p.scale * av * v;
        p.out[o] = p.accumulate ? p.out[o] + res : res;
      }
    }
  }
}

static int check_shape(const lrpx_conv_shape* s, int* P, int* Q) {
  if (!s || s->n <= 0 || s->cin <= 0 || s->h <= 0 || s->w <= 0 || s->cout <= 0 || s->kh <= 0 || s->kw <= 0 ||
      s->stride_h <= 0 || s->stride_w <= 0 || s->pad_h < 0 || s->pad_w < 0 || s->dil_h <= 0 || s->dil_w <= 0)
    return -1;
  *P = (s->h + 2 * s->pad_h - s->dil_h * (s->kh - 1) - 1) / s->stride_h + 1;
  *Q = (s->w + 2 * s->pad_w - s->dil_w * (s->kw - 1) - 1) / s->stride_w + 1;
  if (*P <= 0 || *Q <= 0) return -1;
  if ((long long)s->n * s->h * s->w >= (1LL << 31) || (long long)s->n * (*P) * (*Q) >= (1LL << 31)) return -1;
  return 0;
}

template <bool SIGNED, bool DGRAD, int EPI>
static int launch(const ConvParams& p, cudaStream_t st) {
  const lrpx_conv_shape& s = p.s;
  long long M = DGRAD ? (long long)s.n * s.h * s.w : (long long)s.n * p.P * p.Q;
  int N = DGRAD ? s.cin : s.cout;
  dim3 grid(ceil_div(M, BM), ceil_div(N, BN));
  conv_igemm_kernel<SIGNED, DGRAD, EPI><<<grid, NT, 0, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("conv_igemm launch failed: %s", cudaGetErrorString(e));
    return LRPX_E_CUDA;
  }
  return LRPX_OK;
}

}  // namespace lrpx

using namespace lrpx;

extern "C" {

int lrpx_conv_rule_s_f32(const float* a, const float* w, const float* bias, const float* r_out, float* s_out,
                         float* z_out, const lrpx_conv_shape* shp, int net, void* stream) {
  ConvParams p{};
  LRPX_CHECK_ARG(a && w && r_out && s_out, "null pointer");
  LRPX_CHECK_ARG(check_shape(shp, &p.P, &p.Q) == 0, "bad conv shape");
  LRPX_CHECK_ARG(net == LRPX_NET_POS || net == LRPX_NET_NEG || net == LRPX_NET_PLAIN, "bad net");
  p.a = a; p.w = w; p.bias = bias; p.r = r_out; p.out = s_out; p.z_out = z_out; p.s = *shp; p.net = net;
  if (net == LRPX_NET_PLAIN) return launch<false, false, EPI_S>(p, as_stream(stream));
  return launch<true, false, EPI_S>(p, as_stream(stream));
}

int lrpx_conv_rule_rin_f32(const float* a, const float* w, const float* s_in, float* r_in,
                           const lrpx_conv_shape* shp, int net, float scale, int accumulate, void* stream) {
  ConvParams p{};
  LRPX_CHECK_ARG(a && w && s_in && r_in, "null pointer");
  LRPX_CHECK_ARG(check_shape(shp, &p.P, &p.Q) == 0, "bad conv shape");
  LRPX_CHECK_ARG(net == LRPX_NET_POS || net == LRPX_NET_NEG || net == LRPX_NET_PLAIN, "bad net");
  p.a = a; p.w = w; p.r = s_in; p.out = r_in; p.s = *shp; p.net = net; p.scale = scale; p.accumulate = accumulate;
  if (net == LRPX_NET_PLAIN) return launch<false, true, EPI_RIN>(p, as_stream(stream));
  return launch<true, true, EPI_RIN>(p, as_stream(stream));
}

int lrpx_conv_forward_f32(const float* a, const float* w, const float* bias, float* out, const lrpx_conv_shape* shp,
                          int relu, void* stream) {
  ConvParams p{};
  LRPX_CHECK_ARG(a && w && out, "null pointer");
  LRPX_CHECK_ARG(check_shape(shp, &p.P, &p.Q) == 0, "bad conv shape");
  p.a = a; p.w = w; p.bias = bias; p.out = out; p.s = *shp; p.net = LRPX_NET_PLAIN; p.relu = relu;
  return launch<false, false, EPI_FWD>(p, as_stream(stream));
}

int lrpx_linear_eps_f32(const float* a, const float* w, const float* bias, const float* r_out, float* r_in,
                        float* s_workspace, int n, int in_features, int out_features, int ignore_bias, void* stream) {
  LRPX_CHECK_ARG(a && w && r_out && r_in && s_workspace && n > 0 && in_features > 0 && out_features > 0,
                 "bad argument");
  LRPX_CHECK_ARG(ignore_bias || bias, "bias required when ignore_bias == 0");
  lrpx_conv_shape shp{n, in_features, 1, 1, out_features, 1, 1, 1, 1, 0, 0, 1, 1};
  int rc = lrpx_conv_rule_s_f32(a, w, ignore_bias ? nullptr : bias, r_out, s_workspace, nullptr, &shp, LRPX_NET_PLAIN,
                                stream);
  if (rc) return rc;
  return lrpx_conv_rule_rin_f32(a, w, s_workspace, r_in, &shp, LRPX_NET_PLAIN, 1.f, 0, stream);
}

}  // extern "C"
