/*
 * oracle/hifigan_oracle.c — CPU restatement of tts-king's HiFi-GAN generator.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under tts_king_b200/ may link, import or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg use it, and only as the checker.
 *
 * Parity pin: the reference holds no golden vectors for this path (SURVEY.md
 * §8c), so this restatement is pinned against outputs of the reference itself
 * (hifi/models.py::Generator imported from /root/reference by
 * tools/make_golden.py); the vectors live in tests/golden/.
 *
 * What each function follows (paths relative to the reference tree):
 *   og_conv1d            torch Conv1d as used at hifi/models.py:19-81,152-154,181
 *                        
 *   og_conv_transpose1d  torch ConvTranspose1d as used at hifi/models.py:161-171
 *                        (scatter form — deliberately NOT the
 *                        polyphase form the CUDA path uses)
 *   og_weight_norm_fold  torch._weight_norm(v, g, 0) reached through
 *                        remove_weight_norm, hifi/models.py:97-101,203-210
 *   og_forward           Generator.forward hifi/models.py:185-201 with
 *                        ResBlock1.forward :88-95 / ResBlock2.forward :134-139
 *
 * Layout is the reference's own: activations [B][C][L], Conv1d weights
 * [C_out][C_in][k], ConvTranspose1d weights [C_in][C_out][k].
 * Built twice: REAL=float (og_*_f32) and REAL=double (og_*_f64).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdint.h>

#define OG_MAX_UPS 8
#define OG_MAX_KERNELS 8
#define OG_MAX_DIL 4

typedef struct {
  int num_mels;                 /* 80 */
  int upsample_initial_channel; /* h.upsample_initial_channel */
  int num_upsamples;
  int upsample_rates[OG_MAX_UPS];
  int upsample_kernel_sizes[OG_MAX_UPS];
  int num_kernels;
  int resblock_kernel_sizes[OG_MAX_KERNELS];
  int num_dilations; /* entries used per resblock */
  int resblock_dilation_sizes[OG_MAX_KERNELS][OG_MAX_DIL];
  int resblock_type; /* 1 = ResBlock1, 2 = ResBlock2 */
} og_config;

/* get_padding — hifi/vocoder/utils.py:36-37 */
int og_get_padding(int kernel_size, int dilation) {
  return (kernel_size * dilation - dilation) / 2;
}

/* number of weight arrays og_forward expects (each conv contributes w then b) */
int og_num_weight_arrays(const og_config* c) {
  int per_block = (c->resblock_type == 1 ? 2 : 1) * c->num_dilations;
  int convs = 1 + c->num_upsamples + c->num_upsamples * c->num_kernels * per_block + 1;
  return 2 * convs;
}

#define REAL float
#define SUF(x) x##_f32
#include "hifigan_oracle_impl.h"
#undef REAL
#undef SUF

#define REAL double
#define SUF(x) x##_f64
#include "hifigan_oracle_impl.h"
#undef REAL
#undef SUF

/* float in / float out convenience around the double build: everything is
 * widened, computed in fp64, and rounded once at the end. */
int og_forward_f32_via_f64(const og_config* cfg, const float* const* weights,
                           const int64_t* weight_sizes, const float* mel, int B,
                           int T, float* out) {
  int nw = og_num_weight_arrays(cfg);
  double** w = (double**)calloc((size_t)nw, sizeof(double*));
  if (!w) return -1;
  int rc = 0;
  for (int i = 0; i < nw; ++i) {
    w[i] = (double*)malloc(sizeof(double) * (size_t)weight_sizes[i]);
    if (!w[i]) { rc = -1; goto done; }
    for (int64_t j = 0; j < weight_sizes[i]; ++j) w[i][j] = (double)weights[i][j];
  }
  {
    int64_t nin = (int64_t)B * cfg->num_mels * T;
    int64_t up = 1;
    for (int i = 0; i < cfg->num_upsamples; ++i) up *= cfg->upsample_rates[i];
    int64_t nout = (int64_t)B * T * up;
    double* m = (double*)malloc(sizeof(double) * (size_t)nin);
    double* o = (double*)malloc(sizeof(double) * (size_t)nout);
    if (!m || !o) { free(m); free(o); rc = -1; goto done; }
    for (int64_t j = 0; j < nin; ++j) m[j] = (double)mel[j];
    rc = og_forward_f64(cfg, (const double* const*)w, m, B, T, o);
    for (int64_t j = 0; j < nout; ++j) out[j] = (float)o[j];
    free(m); free(o);
  }
done:
  for (int i = 0; i < nw; ++i) free(w[i]);
  free(w);
  return rc;
}

/* HIFIapi.generate tail — hifiapi.py:50-51: audio * MAX_WAV_VALUE then numpy
 * astype("int16").  numpy's float32->int16 cast on x86-64 goes through a
 * truncating float->int32 conversion followed by a wrap to 16 bits, so +1.0
 * (32768.0) becomes -32768. */
void og_to_int16(const float* wav, int64_t n, float max_wav_value, int16_t* out) {
  for (int64_t i = 0; i < n; ++i) {
    float s = wav[i] * max_wav_value;
    int32_t t = (int32_t)s; /* truncation toward zero */
    out[i] = (int16_t)(uint16_t)(uint32_t)t;
  }
}

/* The fused-pair epilogue's x / num_kernels for num_kernels = 3 (hifi/models.py:196), as
 * tts_king_b200/csrc/common.cuh::div3_rn computes it: q0 = x * RN(1/3), rem = fma(-3, q0, x),
 * q = fma(rem, RN(1/3), q0), sign of x copied onto q; infinities and NaNs take the ordinary division.
 * og_div3_sweep walks the bit patterns start, start + stride, ... (count of them) and returns how many
 * differ from the IEEE quotient x / 3.0f (NaN counts as equal to NaN); *first_bad receives the first one. */
static float og_div3(float v) {
  const float r = 0x1.555556p-2f;
  if (!(fabsf(v) < INFINITY)) return v / 3.0f;
  float q0 = v * r;
  float rem = fmaf(-3.0f, q0, v);
  float q = fmaf(rem, r, q0);
  uint32_t qb, vb;
  memcpy(&qb, &q, 4);
  memcpy(&vb, &v, 4);
  qb = (qb & 0x7fffffffu) | (vb & 0x80000000u);
  memcpy(&q, &qb, 4);
  return q;
}
int64_t og_div3_sweep(uint32_t start, uint32_t stride, int64_t count, uint32_t* first_bad) {
  int64_t bad = 0;
  uint32_t b = start;
  for (int64_t i = 0; i < count; ++i, b += stride) {
    float v, ref, got;
    uint32_t rb, gb;
    memcpy(&v, &b, 4);
    ref = v / 3.0f;
    got = og_div3(v);
    memcpy(&rb, &ref, 4);
    memcpy(&gb, &got, 4);
    if (rb != gb && !(ref != ref && got != got)) {
      if (bad == 0 && first_bad) *first_bad = b;
      ++bad;
    }
  }
  return bad;
}
