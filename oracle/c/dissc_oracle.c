/*
 * Plain-C restatement of the DISSC vocoder forward -- TEST INFRASTRUCTURE ONLY.
 *
 * Independent of PyTorch: direct-form loops with double accumulators, so it is
 * both a check on the index arithmetic (padding, dilation, transposed-conv
 * phases) and a higher-precision ground truth than the fp32 reference.
 * Follows /root/reference/sr/models.py: Generator.forward :98-114,
 * ResBlock1.forward :34-41, CodeGenerator.forward :189,:206-215, with
 * get_padding from sr/utils.py:44-45.
 *
 * Built by oracle/c/Makefile into oracle/_build/libdissc_oracle.so and used only
 * from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LRELU_SLOPE 0.1f /* sr/models.py:13 */

static inline float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

/* y[b,co,t] = bias[co] + sum_{ci,j} w[co,ci,j] * x[b,ci,t*stride + j*dil - pad]   (torch.nn.Conv1d, groups) */
void oracle_conv1d(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int Cout, int Tin,
                   int k, int stride, int dil, int pad, int groups) {
  int Tout = (Tin + 2 * pad - dil * (k - 1) - 1) / stride + 1;
  int cig = Cin / groups, cog = Cout / groups;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int co = 0; co < Cout; ++co) {
      int g = co / cog;
      float* yr = y + ((size_t)b * Cout + co) * Tout;
      double* acc = (double*)malloc(sizeof(double) * (size_t)Tout);
      for (int t = 0; t < Tout; ++t) acc[t] = bias ? (double)bias[co] : 0.0;
      for (int ci = 0; ci < cig; ++ci) {
        const float* xr = x + ((size_t)b * Cin + g * cig + ci) * Tin;
        for (int j = 0; j < k; ++j) {
          double wv = w[((size_t)co * cig + ci) * k + j];
          int off = j * dil - pad;
          for (int t = 0; t < Tout; ++t) {
            int s = t * stride + off;
            if (s >= 0 && s < Tin) acc[t] += wv * (double)xr[s];
          }
        }
      }
      for (int t = 0; t < Tout; ++t) yr[t] = (float)acc[t];
      free(acc);
    }
}

/* torch.nn.ConvTranspose1d: w is (Cin,Cout,k); y[b,co,i*stride + j - pad] += x[b,ci,i]*w[ci,co,j]; Tout=(Tin-1)*stride-2*pad+k */
void oracle_conv_transpose1d(const float* x, const float* w, const float* bias, float* y, int B, int Cin, int Cout,
                             int Tin, int k, int stride, int pad) {
  int Tout = (Tin - 1) * stride - 2 * pad + k;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b)
    for (int co = 0; co < Cout; ++co) {
      double* acc = (double*)malloc(sizeof(double) * (size_t)Tout);
      for (int t = 0; t < Tout; ++t) acc[t] = bias ? (double)bias[co] : 0.0;
      for (int ci = 0; ci < Cin; ++ci) {
        const float* xr = x + ((size_t)b * Cin + ci) * Tin;
        for (int j = 0; j < k; ++j) {
          double wv = w[((size_t)ci * Cout + co) * k + j];
          for (int i = 0; i < Tin; ++i) {
            int t = i * stride + j - pad;
            if (t >= 0 && t < Tout) acc[t] += wv * (double)xr[i];
          }
        }
      }
      float* yr = y + ((size_t)b * Cout + co) * Tout;
      for (int t = 0; t < Tout; ++t) yr[t] = (float)acc[t];
      free(acc);
    }
}

void oracle_leaky_relu(const float* x, float* y, size_t n, float slope) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) y[i] = lrelu(x[i], slope);
}

/* sr/models.py:189,:206-215: x = cat[dict(code)^T, f0, repeat(spkr_emb)^T] -> (B, 2E+1, T) */
void oracle_build_input(const int64_t* code, const float* f0, const int64_t* spkr, const float* dict_w,
                        const float* spkr_w, float* x, int B, int T, int E) {
  int C = 2 * E + 1;
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < T; ++t) {
      const float* e = dict_w + (size_t)code[(size_t)b * T + t] * E;
      const float* s = spkr_w + (size_t)spkr[b] * E;
      for (int c = 0; c < E; ++c) x[((size_t)b * C + c) * T + t] = e[c];
      x[((size_t)b * C + E) * T + t] = f0[(size_t)b * T + t];
      for (int c = 0; c < E; ++c) x[((size_t)b * C + E + 1 + c) * T + t] = s[c];
    }
}

/*
 * Whole Generator.forward for resblock "1".  Weights are FOLDED (weight-norm
 * removed) and passed in module order:
 *   conv_pre.w, conv_pre.b,
 *   for each stage i: ups.i.w, ups.i.b, then for j in kernels, for m in dilations:
 *       convs1.m.w, convs1.m.b, convs2.m.w, convs2.m.b
 *   conv_post.w, conv_post.b
 * x is (B, Cin0, T); y is (B, 1, T*prod(rates)).
 */
int oracle_generator_forward(const float* x_in, float* y, int B, int Cin0, int T, int C0, int n_up, const int* up_rates,
                             const int* up_ks, int n_rk, const int* rks, int n_dil, const int* dils /* n_rk*n_dil */,
                             const float* const* W) {
  int wi = 0;
  size_t cur_T = T;
  int ch = C0;
  float* x = (float*)malloc(sizeof(float) * (size_t)B * C0 * T);
  if (!x) return -1;
  oracle_conv1d(x_in, W[wi], W[wi + 1], x, B, Cin0, C0, T, 7, 1, 1, 3, 1); /* :99 */
  wi += 2;
  for (int i = 0; i < n_up; ++i) {
    int u = up_rates[i], k = up_ks[i], co = ch / 2;
    size_t n_in = (size_t)B * ch * cur_T;
    oracle_leaky_relu(x, x, n_in, LRELU_SLOPE); /* :101 */
    size_t new_T = (cur_T - 1) * u - 2 * ((k - u) / 2) + k;
    float* xu = (float*)malloc(sizeof(float) * (size_t)B * co * new_T);
    if (!xu) return -1;
    oracle_conv_transpose1d(x, W[wi], W[wi + 1], xu, B, ch, co, (int)cur_T, k, u, (k - u) / 2); /* :102 */
    wi += 2;
    free(x);
    ch = co;
    cur_T = new_T;
    size_t n = (size_t)B * ch * cur_T;
    float* xs = (float*)calloc(n, sizeof(float));
    float* r = (float*)malloc(sizeof(float) * n);
    float* xt = (float*)malloc(sizeof(float) * n);
    float* xt2 = (float*)malloc(sizeof(float) * n);
    if (!xs || !r || !xt || !xt2) return -1;
    for (int j = 0; j < n_rk; ++j) { /* :104-108 */
      int rk = rks[j];
      memcpy(r, xu, sizeof(float) * n);
      for (int m = 0; m < n_dil; ++m) { /* ResBlock1.forward :34-41 */
        int d = dils[j * n_dil + m];
        oracle_leaky_relu(r, xt, n, LRELU_SLOPE);
        oracle_conv1d(xt, W[wi], W[wi + 1], xt2, B, ch, ch, (int)cur_T, rk, 1, d, (rk * d - d) / 2, 1);
        oracle_leaky_relu(xt2, xt2, n, LRELU_SLOPE);
        oracle_conv1d(xt2, W[wi + 2], W[wi + 3], xt, B, ch, ch, (int)cur_T, rk, 1, 1, (rk - 1) / 2, 1);
        wi += 4;
        for (size_t e = 0; e < n; ++e) r[e] = xt[e] + r[e];
      }
      for (size_t e = 0; e < n; ++e) xs[e] += r[e];
    }
    for (size_t e = 0; e < n; ++e) xs[e] = xs[e] / (float)n_rk; /* :109 */
    free(r);
    free(xt);
    free(xt2);
    free(xu);
    x = xs;
  }
  size_t n = (size_t)B * ch * cur_T;
  oracle_leaky_relu(x, x, n, 0.01f);                                            /* :110 default slope */
  oracle_conv1d(x, W[wi], W[wi + 1], y, B, ch, 1, (int)cur_T, 7, 1, 1, 3, 1); /* :111 */
  for (size_t e = 0; e < (size_t)B * cur_T; ++e) y[e] = tanhf(y[e]);           /* :112 */
  free(x);
  return 0;
}

/* k-means assignment: argmin_j ||x_n - c_j||^2, lowest index on ties (textless KMeansQuantizer). */
void oracle_kmeans_assign(const float* x, const float* c, int64_t* out, int N, int D, int K) {
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; ++n) {
    double best = INFINITY;
    int bi = 0;
    for (int j = 0; j < K; ++j) {
      double s = 0.0;
      for (int d = 0; d < D; ++d) {
        double diff = (double)x[(size_t)n * D + d] - (double)c[(size_t)j * D + d];
        s += diff * diff;
      }
      if (s < best) {
        best = s;
        bi = j;
      }
    }
    out[n] = bi;
  }
}

/* infer.py:158-172 len_carryover_correction on host floats; round = half-to-even (torch.round). */
void oracle_len_carryover(const float* lens, int32_t* out, int L) {
  float total = 0.f;
  for (int i = 0; i < L; ++i) {
    float c = lens[i] < 1.f ? 1.f : lens[i];
    float r = nearbyintf(c); /* default rounding mode = to nearest even */
    float a = lens[i] - r;
    int v = 0;
    total += a;
    if (total >= 1.f) {
      v = 1;
      total -= 1.f;
    } else if (total <= -1.f) {
      v = -1;
      total += 1.f;
    }
    out[i] = (int32_t)r + v;
  }
}
