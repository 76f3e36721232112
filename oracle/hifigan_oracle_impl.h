/*
 * oracle/hifigan_oracle_impl.h — type-generic body of the CPU oracle.
 * Included twice by hifigan_oracle.c with REAL = float / double.
 * TEST INFRASTRUCTURE ONLY (see hifigan_oracle.c header).
 */

/* y[b,o,t] = bias[o] + sum_c sum_j W[o,c,j] * x[b,c,t + j*d - p], x = 0 outside
 * [0,L).  Same-length (stride 1) Conv1d — hifi/models.py:19-81,152-154,181. */
static void SUF(og_conv1d)(const REAL* x, int B, int Cin, int L, const REAL* W,
                           const REAL* bias, int Cout, int k, int d, int p,
                           REAL* y) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int o = 0; o < Cout; ++o) {
      REAL* yo = y + ((int64_t)b * Cout + o) * L;
      const REAL bo = bias ? bias[o] : (REAL)0;
      for (int t = 0; t < L; ++t) yo[t] = bo;
      for (int c = 0; c < Cin; ++c) {
        const REAL* xc = x + ((int64_t)b * Cin + c) * L;
        for (int j = 0; j < k; ++j) {
          const REAL w = W[((int64_t)o * Cin + c) * k + j];
          const int shift = j * d - p;
          int t0 = shift < 0 ? -shift : 0;
          int t1 = L - shift < L ? L - shift : L;
          for (int t = t0; t < t1; ++t) yo[t] += w * xc[t + shift];
        }
      }
    }
  }
}

/* y[b,o,n] = bias[o] + sum over (i,j) with n = i*s - p + j of W[c,o,j]*x[b,c,i].
 * ConvTranspose1d in scatter form — hifi/models.py:161-171; L_out =
 * (L-1)*s - 2p + k. */
static void SUF(og_conv_transpose1d)(const REAL* x, int B, int Cin, int L,
                                     const REAL* W, const REAL* bias, int Cout,
                                     int k, int s, int p, REAL* y, int Lout) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < B; ++b) {
    for (int o = 0; o < Cout; ++o) {
      REAL* yo = y + ((int64_t)b * Cout + o) * Lout;
      const REAL bo = bias ? bias[o] : (REAL)0;
      for (int n = 0; n < Lout; ++n) yo[n] = bo;
      for (int c = 0; c < Cin; ++c) {
        const REAL* xc = x + ((int64_t)b * Cin + c) * L;
        const REAL* wc = W + ((int64_t)c * Cout + o) * k;
        for (int j = 0; j < k; ++j) {
          const REAL w = wc[j];
          for (int i = 0; i < L; ++i) {
            int n = i * s - p + j;
            if (n >= 0 && n < Lout) yo[n] += w * xc[i];
          }
        }
      }
    }
  }
}

static void SUF(og_leaky_relu)(const REAL* x, int64_t n, REAL slope, REAL* y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) y[i] = x[i] > (REAL)0 ? x[i] : x[i] * slope;
}

/* w = v * g / ||v||, norm over every dim but dim 0 — torch._weight_norm(v,g,0).
 * d0 = C_out for Conv1d, C_in for ConvTranspose1d. */
void SUF(og_weight_norm_fold)(const REAL* v, const REAL* g, int d0, int64_t rest,
                              REAL* w) {
  for (int i = 0; i < d0; ++i) {
    const REAL* vi = v + (int64_t)i * rest;
    REAL ss = 0;
    for (int64_t r = 0; r < rest; ++r) ss += vi[r] * vi[r];
    REAL scale = g[i] / (REAL)sqrt((double)ss);
    for (int64_t r = 0; r < rest; ++r) w[(int64_t)i * rest + r] = vi[r] * scale;
  }
}

/* public op-level entry points (used by tests to pin single layers) */
void SUF(og_op_conv1d)(const REAL* x, int B, int Cin, int L, const REAL* W,
                       const REAL* bias, int Cout, int k, int d, int p, REAL* y) {
  SUF(og_conv1d)(x, B, Cin, L, W, bias, Cout, k, d, p, y);
}
void SUF(og_op_conv_transpose1d)(const REAL* x, int B, int Cin, int L,
                                 const REAL* W, const REAL* bias, int Cout, int k,
                                 int s, int p, REAL* y) {
  SUF(og_conv_transpose1d)(x, B, Cin, L, W, bias, Cout, k, s, p, y,
                           (L - 1) * s - 2 * p + k);
}

/*
 * Generator.forward — hifi/models.py:185-201.
 * weights: flat list, each conv contributes {weight, bias}, in module order:
 *   conv_pre, ups[0..U), resblocks[0..U*K) (ResBlock1: convs1[0..D) then
 *   convs2[0..D); ResBlock2: convs[0..D)), conv_post.
 * mel [B][num_mels][T] -> out [B][1][T*prod(rates)].  Returns 0 on success.
 */
int SUF(og_forward)(const og_config* cfg, const REAL* const* weights,
                    const REAL* mel, int B, int T, REAL* out) {
  const int U = cfg->num_upsamples, K = cfg->num_kernels, D = cfg->num_dilations;
  const REAL slope = (REAL)0.1; /* LRELU_SLOPE, hifi/models.py:9 */
  int wi = 0;
  int C = cfg->upsample_initial_channel;
  int64_t L = T;

  /* largest activation: track max C*L over stages */
  int64_t maxelems = (int64_t)C * L;
  {
    int c = C; int64_t l = L;
    for (int i = 0; i < U; ++i) {
      c = cfg->upsample_initial_channel >> (i + 1);
      l *= cfg->upsample_rates[i];
      if ((int64_t)c * l > maxelems) maxelems = (int64_t)c * l;
    }
  }
  const size_t bytes = sizeof(REAL) * (size_t)maxelems * (size_t)B;
  REAL* x = (REAL*)malloc(bytes);   /* stage input / running tensor */
  REAL* xs = (REAL*)malloc(bytes);  /* MRF sum */
  REAL* r = (REAL*)malloc(bytes);   /* resblock running x */
  REAL* t1 = (REAL*)malloc(bytes);
  REAL* t2 = (REAL*)malloc(bytes);
  if (!x || !xs || !r || !t1 || !t2) {
    free(x); free(xs); free(r); free(t1); free(t2);
    return -1;
  }

  /* x = conv_pre(mel)  :186 */
  SUF(og_conv1d)(mel, B, cfg->num_mels, (int)L, weights[wi], weights[wi + 1], C, 7, 1, 3, x);
  wi += 2;

  const int rb_base = wi + 2 * U; /* first resblock weight index */
  const int per_block = (cfg->resblock_type == 1 ? 2 : 1) * D;

  for (int i = 0; i < U; ++i) {
    const int Cin = cfg->upsample_initial_channel >> i;
    const int Cout = cfg->upsample_initial_channel >> (i + 1);
    const int s = cfg->upsample_rates[i], k = cfg->upsample_kernel_sizes[i];
    const int p = (k - s) / 2;
    const int Lout = (int)((L - 1) * s - 2 * p + k);
    /* x = leaky_relu(x, 0.1); x = ups[i](x)  :188-189 */
    SUF(og_leaky_relu)(x, (int64_t)B * Cin * L, slope, t1);
    SUF(og_conv_transpose1d)(t1, B, Cin, (int)L, weights[2 + 2 * i], weights[3 + 2 * i],
                             Cout, k, s, p, x, Lout);
    L = Lout;
    const int64_t n = (int64_t)B * Cout * L;
    /* xs = sum_j resblocks[i*K + j](x)  :190-195 */
    for (int j = 0; j < K; ++j) {
      const int kk = cfg->resblock_kernel_sizes[j];
      const int base = rb_base + 2 * per_block * (i * K + j);
      memcpy(r, x, sizeof(REAL) * (size_t)n);
      for (int m = 0; m < D; ++m) {
        const int dil = cfg->resblock_dilation_sizes[j][m];
        if (cfg->resblock_type == 1) {
          /* ResBlock1.forward :88-95 */
          const REAL* w1 = weights[base + 2 * m];
          const REAL* b1 = weights[base + 2 * m + 1];
          const REAL* w2 = weights[base + 2 * (D + m)];
          const REAL* b2 = weights[base + 2 * (D + m) + 1];
          SUF(og_leaky_relu)(r, n, slope, t1);
          SUF(og_conv1d)(t1, B, Cout, (int)L, w1, b1, Cout, kk, dil, og_get_padding(kk, dil), t2);
          SUF(og_leaky_relu)(t2, n, slope, t1);
          SUF(og_conv1d)(t1, B, Cout, (int)L, w2, b2, Cout, kk, 1, og_get_padding(kk, 1), t2);
        } else {
          /* ResBlock2.forward :134-139 */
          const REAL* w1 = weights[base + 2 * m];
          const REAL* b1 = weights[base + 2 * m + 1];
          SUF(og_leaky_relu)(r, n, slope, t1);
          SUF(og_conv1d)(t1, B, Cout, (int)L, w1, b1, Cout, kk, dil, og_get_padding(kk, dil), t2);
        }
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < n; ++e) r[e] = t2[e] + r[e]; /* x = xt + x */
      }
      if (j == 0) {
        memcpy(xs, r, sizeof(REAL) * (size_t)n);
      } else {
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < n; ++e) xs[e] += r[e];
      }
    }
    /* x = xs / num_kernels  :196 */
    const REAL nk = (REAL)K;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n; ++e) x[e] = xs[e] / nk;
    C = Cout;
  }
  wi = rb_base + 2 * per_block * U * K;
  /* x = leaky_relu(x) (default slope 0.01); conv_post; tanh  :197-199 */
  SUF(og_leaky_relu)(x, (int64_t)B * C * L, (REAL)0.01, t1);
  SUF(og_conv1d)(t1, B, C, (int)L, weights[wi], weights[wi + 1], 1, 7, 1, 3, out);
  const int64_t nout = (int64_t)B * L;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < nout; ++e) out[e] = (REAL)tanh((double)out[e]);

  free(x); free(xs); free(r); free(t1); free(t2);
  return 0;
}
