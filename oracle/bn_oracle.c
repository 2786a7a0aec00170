/* bn_oracle.c -- CPU ORACLE (test infrastructure, NOT a product path).
 *
 * A plain-C restatement of the reference's chunk-classification hot path:
 *
 *   PCM16 -> host frontend -> int8 TFLite graph -> per-file pooling
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (the CUDA engine in
 * birdnet-stm32_b200/csrc) never links or calls it.
 *
 * PARITY UNPINNED at the third-party boundaries: the arithmetic of this path
 * lives in dependencies that are NOT vendored under /root/reference and are
 * not installable here (no network, not in /opt/wheelhouse):
 *   - TensorFlow Lite builtin int8 kernels, tensorflow==2.19.0
 *     (reference requirements.txt:2; call sites birdnet_stm32/models/runners.py:57,79,93-95)
 *   - librosa==0.11.0 stft (requirements.txt:1; call site audio/spectrogram.py:106-115)
 *   - numpy==1.26.4 promotion rules in normalize() (audio/spectrogram.py:12-21)
 * and the reference's tests hold no golden vectors for them.
 * The restatement follows the published algorithms of those versions:
 *   gemmlowp fixed-point (SaturatingRoundingDoublingHighMul, RoundingDivideByPOT),
 *   tflite::reference_integer_ops::{ConvPerChannel, DepthwiseConvPerChannel,
 *   FullyConnectedPerChannel, Add, Mean}, reference_ops::AffineQuantize,
 *   LUT-based int8 LOGISTIC, and librosa.stft(center=True, pad_mode="constant").
 * What IS pinned: pooling (reference tests/test_pooling.py known answers and the
 * importable reference evaluation/pooling.py, see tests/golden/), chunk geometry
 * (audio/io.py:133-174), the per-op float-shadow check in tests/test_oracle_graph.py, and --
 * at network level -- the reference's own conversion gate against its shipped FLOAT checkpoint:
 * the scores of checkpoints/birdnet_stm32n6_100.keras (read without TensorFlow, oracle/h5min.py +
 * oracle/keras_float_model.py, golden in tests/golden/keras_float_reference.npz) and the int8
 * scores of this oracle have cosine similarity 0.994 (gate: >= 0.95, conversion/validate.py:51-105,
 * cli/convert.py:187-195), tests/test_keras_float_pin.py; the same test file reproduces all 205,185
 * int8 weights of the .tflite exactly from the float checkpoint (BatchNorm folding + per-channel
 * max/127 quantisation), which fixes layouts and layer identity at the bit level.  What stays
 * unpinned is the run-time rounding of the TFLite kernels (BN_OPT_ROUNDING / BN_OPT_MEAN_VARIANT).
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC (see oracle/Makefile).  -ffast-math must
 * NOT be used (rounding behaviour is the point).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/bn_blob.h"

#define BNO_EXPORT __attribute__((visibility("default")))

typedef struct bno_model {
  uint8_t* blob;
  size_t nbytes;
  const bn_blob_header* hdr;
  const bn_blob_tensor* tensors;
  const bn_blob_op* ops;
  int rounding;      /* 0 = gemmlowp double rounding (TFLite default), 1 = single rounding */
  int mean_variant;  /* 0 = auto, 1 = float (i), 2 = folded 1/N (ii), 3 = reference divide (iii) */
  int threads;
} bno_model;

static __thread char g_err[256];
/* per-op-kind seconds (single-threaded runs only; switched on by bno_profile(1)) */
static int g_profile = 0;
static double g_prof_s[32];
static double g_prof_op[256];
BNO_EXPORT void bno_profile(int on, double* out32) {
  if (out32) memcpy(out32, g_prof_s, sizeof g_prof_s);
  if (on >= 0) { g_profile = on; memset(g_prof_s, 0, sizeof g_prof_s); }
}
BNO_EXPORT void bno_profile_ops(double* out256) { memcpy(out256, g_prof_op, sizeof g_prof_op); memset(g_prof_op, 0, sizeof g_prof_op); }
BNO_EXPORT const char* bno_last_error(void) { return g_err; }

/* ------------------------------------------------------------------------- */
/* gemmlowp fixed point (third-party algorithm, restated)                     */
/* ------------------------------------------------------------------------- */
static inline int32_t srdhm(int32_t a, int32_t b) {
  /* SaturatingRoundingDoublingHighMul */
  if (a == INT32_MIN && b == INT32_MIN) return INT32_MAX;
  int64_t ab = (int64_t)a * (int64_t)b;
  int32_t nudge = ab >= 0 ? (1 << 30) : (1 - (1 << 30));
  return (int32_t)((ab + nudge) / ((int64_t)1 << 31)); /* C division truncates toward zero */
}
static inline int32_t rdivpot(int32_t x, int e) {
  /* RoundingDivideByPOT: round half away from zero */
  if (e == 0) return x;
  int32_t mask = (int32_t)(((int64_t)1 << e) - 1);
  int32_t rem = x & mask;
  int32_t thr = (mask >> 1) + (x < 0 ? 1 : 0);
  return (x >> e) + (rem > thr ? 1 : 0);
}
/* tflite::MultiplyByQuantizedMultiplier (common.h), both build flavours */
static inline int32_t mbqm(int32_t x, int32_t qm, int shift, int rounding) {
  if (rounding == 0) {
    int left = shift > 0 ? shift : 0, right = shift > 0 ? 0 : -shift;
    return rdivpot(srdhm((int32_t)((int64_t)x * ((int64_t)1 << left)), qm), right);
  } else {
    /* TFLITE_SINGLE_ROUNDING=1 / ruy::MultiplyByQuantizedMultiplier */
    int total_shift = 31 - shift;
    int64_t round = (int64_t)1 << (total_shift - 1);
    int64_t r = (int64_t)x * (int64_t)qm + round;
    r = r >> total_shift;
    if (r > INT32_MAX) r = INT32_MAX;
    if (r < INT32_MIN) r = INT32_MIN;
    return (int32_t)r;
  }
}
static inline int32_t clampi(int32_t v, int32_t lo, int32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------------------------------- */
/* model loading                                                              */
/* ------------------------------------------------------------------------- */
BNO_EXPORT bno_model* bno_load(const void* blob, size_t nbytes) {
  if (nbytes < sizeof(bn_blob_header)) { snprintf(g_err, sizeof g_err, "blob too small"); return NULL; }
  const bn_blob_header* h = (const bn_blob_header*)blob;
  if (memcmp(h->magic, "BNB200\0\0", 8) != 0 || h->version != BN_BLOB_VERSION) {
    snprintf(g_err, sizeof g_err, "bad blob magic/version"); return NULL;
  }
  if (h->total_bytes != nbytes) { snprintf(g_err, sizeof g_err, "blob size mismatch"); return NULL; }
  bno_model* m = (bno_model*)calloc(1, sizeof *m);
  m->blob = (uint8_t*)malloc(nbytes);
  memcpy(m->blob, blob, nbytes);
  m->nbytes = nbytes;
  m->hdr = (const bn_blob_header*)m->blob;
  m->tensors = (const bn_blob_tensor*)(m->blob + m->hdr->tensors_off);
  m->ops = (const bn_blob_op*)(m->blob + m->hdr->ops_off);
  m->rounding = 0;
  m->mean_variant = 0;
  m->threads = 0;
  return m;
}
BNO_EXPORT void bno_free(bno_model* m) { if (m) { free(m->blob); free(m); } }
BNO_EXPORT void bno_set_option(bno_model* m, int rounding, int mean_variant, int threads) {
  m->rounding = rounding; m->mean_variant = mean_variant; m->threads = threads;
}
BNO_EXPORT int bno_num_classes(const bno_model* m) { return (int)m->hdr->num_classes; }
BNO_EXPORT long bno_input_elems(const bno_model* m) {
  const bn_blob_tensor* t = &m->tensors[m->hdr->input_tensor];
  return (long)t->dims[0] * t->dims[1] * t->dims[2];
}
BNO_EXPORT long bno_tensor_bytes(const bno_model* m, int tfl_id) {
  for (uint32_t i = 0; i < m->hdr->n_tensors; i++)
    if (m->tensors[i].id == tfl_id) return (long)m->tensors[i].nbytes;
  return -1;
}

/* ------------------------------------------------------------------------- */
/* graph execution for ONE chunk                                              */
/* ------------------------------------------------------------------------- */
static int run_one(const bno_model* m, const float* input, float* output, void** buf,
                   int tap_slot, void* tap_out) {
  const bn_blob_header* h = m->hdr;
  const bn_blob_tensor* T = m->tensors;
  const int R = m->rounding;
  memcpy(buf[h->input_tensor], input, T[h->input_tensor].nbytes);
  for (uint32_t oi = 0; oi < h->n_ops; oi++) {
    const bn_blob_op* op = &m->ops[oi];
    struct timespec ts0;
    if (g_profile) clock_gettime(CLOCK_MONOTONIC, &ts0);
    const bn_blob_tensor* to = &T[op->out];
    const bn_blob_tensor* ti = op->n_in > 0 ? &T[op->in[0]] : NULL;
    const int32_t* p = op->p;
    switch (op->kind) {
      case BN_OP_QUANTIZE: {
        /* reference_ops::AffineQuantize: TfLiteRound(val / scale) + zp, float32 division */
        const float* x = (const float*)buf[op->in[0]];
        int8_t* y = (int8_t*)buf[op->out];
        const float scale = op->f[0];
        long n = (long)to->nbytes;
        for (long i = 0; i < n; i++) {
          float q = roundf(x[i] / scale);
          int32_t v = (int32_t)q + p[0];
          y[i] = (int8_t)clampi(v, -128, 127);
        }
      } break;
      case BN_OP_DEQUANTIZE: {
        const int8_t* x = (const int8_t*)buf[op->in[0]];
        float* y = (float*)buf[op->out];
        long n = (long)ti->nbytes;
        for (long i = 0; i < n; i++) y[i] = op->f[0] * (float)((int32_t)x[i] - p[0]);
      } break;
      case BN_OP_REQUANT: {
        const int8_t* x = (const int8_t*)buf[op->in[0]];
        int8_t* y = (int8_t*)buf[op->out];
        for (long i = 0; i < (long)to->nbytes; i++)
          y[i] = (int8_t)clampi(mbqm((int32_t)x[i] - p[0], p[2], p[3], R) + p[1], -128, 127);
      } break;
      case BN_OP_TRANSPOSE: {
        const int8_t* x = (const int8_t*)buf[op->in[0]];
        int8_t* y = (int8_t*)buf[op->out];
        const int32_t* id = ti->dims; const int32_t* od = to->dims;
        long is[3] = { (long)id[1] * id[2], id[2], 1 };
        for (int a = 0; a < od[0]; a++) for (int b = 0; b < od[1]; b++) for (int c = 0; c < od[2]; c++) {
          int oidx[3] = { a, b, c };
          long src = 0;
          for (int d = 0; d < 3; d++) src += oidx[d] * is[p[d]];
          y[((long)a * od[1] + b) * od[2] + c] = x[src];
        }
      } break;
      case BN_OP_SLICE: {
        const int8_t* x = (const int8_t*)buf[op->in[0]];
        int8_t* y = (int8_t*)buf[op->out];
        const int32_t* id = ti->dims; const int32_t* od = to->dims;
        for (int a = 0; a < od[0]; a++) for (int b = 0; b < od[1]; b++)
          memcpy(y + ((long)a * od[1] + b) * od[2],
                 x + ((long)(a + p[0]) * id[1] + (b + p[1])) * id[2] + p[2], (size_t)od[2]);
      } break;
      case BN_OP_RESHAPE:
        memcpy(buf[op->out], buf[op->in[0]], to->nbytes);
        break;
      case BN_OP_FILL:
        memset(buf[op->out], (int8_t)p[0], to->nbytes);
        break;
      case BN_OP_CONCAT: {
        const bn_blob_tensor* t1 = &T[op->in[1]];
        const int8_t* x0 = (const int8_t*)buf[op->in[0]];
        const int8_t* x1 = (const int8_t*)buf[op->in[1]];
        int8_t* y = (int8_t*)buf[op->out];
        int axis = p[0];
        long outer = 1, in0 = 1, in1 = 1;
        for (int d = 0; d < axis; d++) outer *= to->dims[d];
        for (int d = axis; d < 3; d++) { in0 *= ti->dims[d]; in1 *= t1->dims[d]; }
        for (long o = 0; o < outer; o++) {
          memcpy(y + o * (in0 + in1), x0 + o * in0, (size_t)in0);
          memcpy(y + o * (in0 + in1) + in0, x1 + o * in1, (size_t)in1);
        }
      } break;
      case BN_OP_CONV2D: {
        /* reference_integer_ops::ConvPerChannel; weights OHWI */
        const int8_t* x = (const int8_t*)buf[op->in[0]];
        int8_t* y = (int8_t*)buf[op->out];
        const int8_t* w = (const int8_t*)(m->blob + op->off[0]);
        const int32_t* bias = (const int32_t*)(m->blob + op->off[1]);
        const int32_t* mult = (const int32_t*)(m->blob + op->off[2]);
        const int32_t* shift = (const int32_t*)(m->blob + op->off[3]);
        const int kh = p[BN_CONV_KH], kw = p[BN_CONV_KW], sh = p[BN_CONV_SH], sw = p[BN_CONV_SW];
        const int pt = p[BN_CONV_PAD_T], pl = p[BN_CONV_PAD_L];
        const int32_t in_off = -p[BN_CONV_IN_ZP], out_zp = p[BN_CONV_OUT_ZP];
        const int ih = ti->dims[0], iw = ti->dims[1], ic = ti->dims[2];
        const int oh = to->dims[0], ow = to->dims[1], oc = to->dims[2];
        if (kh == 1 && kw == 1 && sh == 1 && sw == 1 && pt == 0 && pl == 0 && ic <= 4096) {
          /* pointwise conv: the same sums as the general loop below, with the zero-point offset applied once per input
           * element and int16 operands so that the compiler can vectorise the channel loop */
          /* weights transposed to [ci][co] int32 so that the inner loop runs over output channels */
          int32_t* ws = (int32_t*)malloc(sizeof(int32_t) * (size_t)oc * ic);
          int32_t* __restrict__ accv = (int32_t*)malloc(sizeof(int32_t) * (size_t)oc);
          for (int co = 0; co < oc; co++) for (int ci = 0; ci < ic; ci++) ws[(long)ci * oc + co] = (int32_t)w[(long)co * ic + ci];
          const int32_t amin = p[BN_CONV_ACT_MIN], amax = p[BN_CONV_ACT_MAX];
          for (long px = 0; px < (long)oh * ow; px++) {
            const int8_t* xp = x + px * ic;
            for (int co = 0; co < oc; co++) accv[co] = bias[co];
            for (int ci = 0; ci < ic; ci++) {
              const int32_t xv = (int32_t)xp[ci] + in_off;
              const int32_t* __restrict__ wr = ws + (long)ci * oc;
              for (int co = 0; co < oc; co++) accv[co] += xv * wr[co];
            }
            for (int co = 0; co < oc; co++) {
              const int32_t acc = mbqm(accv[co], mult[co], shift[co], R) + out_zp;
              y[px * oc + co] = (int8_t)clampi(acc, amin, amax);
            }
          }
          free(accv);
          free(ws);
          break;
        }
        {
          /* general case: ConvPerChannel's sums (out-of-image taps skipped), output channels innermost on weights
           * transposed to [fy][fx][ci][co] */
          int32_t* wt = (int32_t*)malloc(sizeof(int32_t) * (size_t)oc * kh * kw * ic);
          int32_t* __restrict__ accv = (int32_t*)malloc(sizeof(int32_t) * (size_t)oc);
          for (int co = 0; co < oc; co++) for (int t = 0; t < kh * kw; t++) for (int ci = 0; ci < ic; ci++)
            wt[((long)t * ic + ci) * oc + co] = (int32_t)w[((long)co * kh * kw + t) * ic + ci];
          for (int oy = 0; oy < oh; oy++) for (int ox = 0; ox < ow; ox++) {
            for (int co = 0; co < oc; co++) accv[co] = bias[co];
            for (int fy = 0; fy < kh; fy++) {
              int iy = oy * sh - pt + fy;
              if (iy < 0 || iy >= ih) continue;
              for (int fx = 0; fx < kw; fx++) {
                int ix = ox * sw - pl + fx;
                if (ix < 0 || ix >= iw) continue;
                const int8_t* xp = x + ((long)iy * iw + ix) * ic;
                for (int ci = 0; ci < ic; ci++) {
                  const int32_t xv = (int32_t)xp[ci] + in_off;
                  const int32_t* __restrict__ wr = wt + ((long)(fy * kw + fx) * ic + ci) * oc;
                  for (int co = 0; co < oc; co++) accv[co] += xv * wr[co];
                }
              }
            }
            int8_t* yp = y + ((long)oy * ow + ox) * oc;
            for (int co = 0; co < oc; co++) {
              const int32_t acc = mbqm(accv[co], mult[co], shift[co], R) + out_zp;
              yp[co] = (int8_t)clampi(acc, p[BN_CONV_ACT_MIN], p[BN_CONV_ACT_MAX]);
            }
          }
          free(accv); free(wt);
        }
      } break;
      case BN_OP_DWCONV2D: {
        /* reference_integer_ops::DepthwiseConvPerChannel, depth_multiplier 1; weights [kh,kw,C] */
        const int8_t* x = (const int8_t*)buf[op->in[0]];
        int8_t* y = (int8_t*)buf[op->out];
        const int8_t* w = (const int8_t*)(m->blob + op->off[0]);
        const int32_t* bias = (const int32_t*)(m->blob + op->off[1]);
        const int32_t* mult = (const int32_t*)(m->blob + op->off[2]);
        const int32_t* shift = (const int32_t*)(m->blob + op->off[3]);
        const int kh = p[BN_CONV_KH], kw = p[BN_CONV_KW], sh = p[BN_CONV_SH], sw = p[BN_CONV_SW];
        const int pt = p[BN_CONV_PAD_T], pl = p[BN_CONV_PAD_L];
        const int32_t in_off = -p[BN_CONV_IN_ZP], out_zp = p[BN_CONV_OUT_ZP];
        const int ih = ti->dims[0], iw = ti->dims[1], C = ti->dims[2];
        const int oh = to->dims[0], ow = to->dims[1];
        /* same sums as DepthwiseConvPerChannel's loop nest, with the channel loop innermost (taps outside the image are
         * skipped exactly as there) so that it vectorises */
        int32_t* __restrict__ accv = (int32_t*)malloc(sizeof(int32_t) * (size_t)C);
        for (int oy = 0; oy < oh; oy++) for (int ox = 0; ox < ow; ox++) {
          for (int c = 0; c < C; c++) accv[c] = bias[c];
          for (int fy = 0; fy < kh; fy++) {
            int iy = oy * sh - pt + fy;
            if (iy < 0 || iy >= ih) continue;
            for (int fx = 0; fx < kw; fx++) {
              int ix = ox * sw - pl + fx;
              if (ix < 0 || ix >= iw) continue;
              const int8_t* __restrict__ xp = x + ((long)iy * iw + ix) * C;
              const int8_t* __restrict__ wp = w + ((long)fy * kw + fx) * C;
              for (int c = 0; c < C; c++) accv[c] += ((int32_t)xp[c] + in_off) * (int32_t)wp[c];
            }
          }
          int8_t* yp = y + ((long)oy * ow + ox) * C;
          for (int c = 0; c < C; c++) {
            const int32_t acc = mbqm(accv[c], mult[c], shift[c], R) + out_zp;
            yp[c] = (int8_t)clampi(acc, p[BN_CONV_ACT_MIN], p[BN_CONV_ACT_MAX]);
          }
        }
        free(accv);
      } break;
      case BN_OP_FC: {
        /* reference_integer_ops::FullyConnectedPerChannel; weights [N,K]; applied to each
         * leading row of the input (keep_num_dims on [1,1,K] -> rows = dims0*dims1) */
        const int8_t* x = (const int8_t*)buf[op->in[0]];
        int8_t* y = (int8_t*)buf[op->out];
        const int8_t* w = (const int8_t*)(m->blob + op->off[0]);
        const int32_t* bias = (const int32_t*)(m->blob + op->off[1]);
        const int32_t* mult = (const int32_t*)(m->blob + op->off[2]);
        const int32_t* shift = (const int32_t*)(m->blob + op->off[3]);
        const int K = p[BN_CONV_CIN], N = p[BN_CONV_COUT];
        const long rows = (long)ti->nbytes / K;
        const int32_t in_off = -p[BN_CONV_IN_ZP], out_zp = p[BN_CONV_OUT_ZP];
        for (long r = 0; r < rows; r++) for (int n = 0; n < N; n++) {
          int32_t acc = 0;
          for (int k = 0; k < K; k++) acc += ((int32_t)x[r * K + k] + in_off) * (int32_t)w[(long)n * K + k];
          acc += bias[n];
          acc = mbqm(acc, mult[n], shift[n], R) + out_zp;
          y[r * N + n] = (int8_t)clampi(acc, p[BN_CONV_ACT_MIN], p[BN_CONV_ACT_MAX]);
        }
      } break;
      case BN_OP_ADD: {
        /* reference_integer_ops::Add (left_shift 20, three multipliers) */
        const int8_t* a = (const int8_t*)buf[op->in[0]];
        const bn_blob_tensor* tb = &T[op->in[1]];
        const int8_t* b = tb->is_const ? (const int8_t*)(m->blob + tb->data_off) : (const int8_t*)buf[op->in[1]];
        int8_t* y = (int8_t*)buf[op->out];
        const int C = to->dims[2];
        const long n = (long)to->nbytes;
        const int bc = p[BN_ADD_BCAST];
        /* scaled_input = MBQM((x - zp) << left_shift, m, s) is a function of the int8 code alone (per-tensor scales):
         * the same expression evaluated once per code instead of once per element */
        int32_t t1[256], t2[256];
        for (int q = -128; q < 128; q++) {
          t1[q + 128] = mbqm((q - p[BN_ADD_IN1_ZP]) * (1 << p[BN_ADD_LEFT_SHIFT]), p[BN_ADD_M1], p[BN_ADD_S1], R);
          t2[q + 128] = mbqm((q - p[BN_ADD_IN2_ZP]) * (1 << p[BN_ADD_LEFT_SHIFT]), p[BN_ADD_M2], p[BN_ADD_S2], R);
        }
        for (long i = 0; i < n; i++) {
          const int32_t s1 = t1[(int)a[i] + 128];
          const int32_t s2 = t2[(int)b[bc ? (i % C) : i] + 128];
          int32_t o = mbqm(s1 + s2, p[BN_ADD_MO], p[BN_ADD_SO], R) + p[BN_ADD_OUT_ZP];
          y[i] = (int8_t)clampi(o, p[BN_ADD_ACT_MIN], p[BN_ADD_ACT_MAX]);
        }
      } break;
      case BN_OP_MUL: {
        /* reference_integer_ops::Mul; p: in1_zp,in2_zp,out_zp,mult,shift,act_min,act_max,bcast */
        const int8_t* a = (const int8_t*)buf[op->in[0]];
        const bn_blob_tensor* tb = &T[op->in[1]];
        const int8_t* b = tb->is_const ? (const int8_t*)(m->blob + tb->data_off) : (const int8_t*)buf[op->in[1]];
        int8_t* y = (int8_t*)buf[op->out];
        const int C = to->dims[2];
        const int bc = p[7];
        for (long i = 0; i < (long)to->nbytes; i++) {
          int32_t v1 = (int32_t)a[i] - p[0];
          /* bcast: 0 none, 1 const [C], 2 per-chunk [1,1,C] (SE gate), 3 per-position [H,W,1] (attention weights) */
          int32_t v2 = (int32_t)b[bc == 3 ? (i / C) : (bc ? (i % C) : i)] - p[1];
          int32_t o = mbqm(v1 * v2, p[3], p[4], R) + p[2];
          y[i] = (int8_t)clampi(o, p[5], p[6]);
        }
      } break;
      case BN_OP_MEAN: {
        const int8_t* x = (const int8_t*)buf[op->in[0]];
        int8_t* y = (int8_t*)buf[op->out];
        const int C = ti->dims[2];
        const int N = p[BN_MEAN_COUNT];
        int variant = m->mean_variant;
        if (variant == 0) variant = p[BN_MEAN_KEEP_DIMS] ? 3 : 2;
        for (int c = 0; c < C; c++) {
          int32_t sum = 0;
          for (int i = 0; i < N; i++) sum += x[(long)i * C + c];
          int32_t o;
          if (variant == 1) {
            /* (i) reference_ops::QuantizedMeanOrSum, float */
            const float scale = T[op->in[0]].scale / to->scale;
            const float bias = -(float)p[BN_MEAN_IN_ZP] * scale;
            float fm = (float)sum / (float)N;
            o = (int32_t)roundf(fm * scale + bias) + p[BN_MEAN_OUT_ZP];
          } else if (variant == 2) {
            /* (ii) TFLite >= 2.10 reduce.cc: 1/N folded into the multiplier */
            o = mbqm(sum - p[BN_MEAN_IN_ZP] * N, p[BN_MEAN_MULT_N], p[BN_MEAN_SHIFT_N], R) + p[BN_MEAN_OUT_ZP];
          } else {
            /* (iii) reference_integer_ops::Mean: requantise the sum, then rounded divide */
            int32_t acc = mbqm(sum - p[BN_MEAN_IN_ZP] * N, p[BN_MEAN_MULT], p[BN_MEAN_SHIFT], R);
            acc = acc > 0 ? (acc + N / 2) / N : (acc - N / 2) / N;
            o = acc + p[BN_MEAN_OUT_ZP];
          }
          y[c] = (int8_t)clampi(o, -128, 127);
        }
      } break;
      case BN_OP_LOGISTIC: {
        const int8_t* x = (const int8_t*)buf[op->in[0]];
        int8_t* y = (int8_t*)buf[op->out];
        const int8_t* lut = (const int8_t*)(m->blob + op->off[0]);
        for (long i = 0; i < (long)to->nbytes; i++) y[i] = lut[(uint8_t)(x[i] + 128)];
      } break;
      case BN_OP_PAD: {
        /* reference_ops::Pad on int8: the pad value is the tensor's zero point (p[6]) */
        const int8_t* x = (const int8_t*)buf[op->in[0]];
        int8_t* y = (int8_t*)buf[op->out];
        const int d0 = ti->dims[0], d1 = ti->dims[1], d2 = ti->dims[2];
        const int o0 = to->dims[0], o1 = to->dims[1], o2 = to->dims[2];
        for (int h = 0; h < o0; h++)
          for (int w = 0; w < o1; w++)
            for (int c = 0; c < o2; c++) {
              const int sh = h - p[0], sw = w - p[2], sc = c - p[4];
              int v = p[6];
              if (sh >= 0 && sh < d0 && sw >= 0 && sw < d1 && sc >= 0 && sc < d2) v = x[((long)sh * d1 + sw) * d2 + sc];
              y[((long)h * o1 + w) * o2 + c] = (int8_t)v;
            }
      } break;
      case BN_OP_SOFTMAX: {
        /* tflite::optimized_ops::Softmax (int8 in/out, float LUT): the kernel BuiltinOpResolver registers on x86.
         * table[255 - v] = expf(-in_scale * beta * v) is precomputed by the exporter (off[0]); f[1] = output scale. */
        const int8_t* x = (const int8_t*)buf[op->in[0]];
        int8_t* y = (int8_t*)buf[op->out];
        const float* table = (const float*)(m->blob + op->off[0]);
        const int L = to->dims[2];
        const long rows = (long)to->dims[0] * to->dims[1];
        for (long r = 0; r < rows; r++) {
          const int8_t* xp = x + r * L;
          int32_t mx = -128;
          for (int j = 0; j < L; j++) if (xp[j] > mx) mx = xp[j];
          const float* toff = table + (255 - mx);
          volatile float sum = 0.0f;
          for (int j = 0; j < L; j++) sum = sum + toff[xp[j]];
          const float inv = 1.0f / (float)(sum * op->f[1]);
          for (int j = 0; j < L; j++) {
            const float pr = toff[xp[j]] * inv;
            const int32_t q = (int32_t)(float)(pr + 0.5f) + p[0];
            y[r * L + j] = (int8_t)clampi(q, -128, 127);
          }
        }
      } break;
      case BN_OP_SUM: {
        /* reference_ops::QuantizedMeanOrSum(compute_sum = true): p = axis, count, in_zp, out_zp; f = scale, bias */
        const int8_t* x = (const int8_t*)buf[op->in[0]];
        int8_t* y = (int8_t*)buf[op->out];
        const int axis = p[0];
        long outer = 1, inner = 1;
        for (int d = 0; d < axis; d++) outer *= ti->dims[d];
        for (int d = axis + 1; d < 3; d++) inner *= ti->dims[d];
        const int count = ti->dims[axis];
        for (long o = 0; o < outer; o++)
          for (long in_ = 0; in_ < inner; in_++) {
            int32_t sum = 0;
            for (int j = 0; j < count; j++) sum += x[(o * count + j) * inner + in_];
            const float prod = (float)sum * op->f[0];
            const float v = prod + op->f[1];
            y[o * inner + in_] = (int8_t)clampi((int32_t)roundf(v) + p[3], -128, 127);
          }
      } break;
      default:
        snprintf(g_err, sizeof g_err, "oracle: unsupported op kind %d (tflite op %d)", op->kind, op->tfl_index);
        return -1;
    }
    if ((int)op->out == tap_slot && tap_out) memcpy(tap_out, buf[op->out], to->nbytes);
    if (g_profile) {
      struct timespec ts1;
      clock_gettime(CLOCK_MONOTONIC, &ts1);
      const double dt = (double)(ts1.tv_sec - ts0.tv_sec) + 1e-9 * (double)(ts1.tv_nsec - ts0.tv_nsec);
      g_prof_s[op->kind & 31] += dt;
      g_prof_op[oi & 255] += dt;
    }
  }
  memcpy(output, buf[h->output_tensor], T[h->output_tensor].nbytes);
  return 0;
}

/* Run the int8 graph on B chunks.  `input` is the graph's float input [B, in_elems],
 * `output` float [B, C].  If tap_id >= 0 the tensor with that TFLite index is copied to
 * tap_out [B, nbytes].  Replaces tf.lite.Interpreter.invoke (models/runners.py:93-95). */
BNO_EXPORT int bno_run_graph(const bno_model* m, const float* input, int B, float* output,
                             int tap_id, void* tap_out) {
  const bn_blob_header* h = m->hdr;
  const long in_elems = bno_input_elems(m);
  const long out_elems = (long)m->tensors[h->output_tensor].nbytes / 4;
  int tap_slot = -1; long tap_bytes = 0;
  if (tap_id >= 0) {
    for (uint32_t i = 0; i < h->n_tensors; i++) if (m->tensors[i].id == tap_id) { tap_slot = (int)i; tap_bytes = (long)m->tensors[i].nbytes; }
    if (tap_slot < 0) { snprintf(g_err, sizeof g_err, "no tensor with id %d", tap_id); return -1; }
    if (tap_slot == h->input_tensor) { memcpy(tap_out, input, (size_t)B * tap_bytes); tap_slot = -1; }
  }
  int rc = 0;
  int nthreads = m->threads;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
  {
    void** buf = (void**)calloc(h->n_tensors, sizeof(void*));
    for (uint32_t i = 0; i < h->n_tensors; i++)
      if (!m->tensors[i].is_const) buf[i] = malloc((size_t)m->tensors[i].nbytes + 16);
#pragma omp for schedule(dynamic, 1)
    for (int b = 0; b < B; b++) {
      int r = run_one(m, input + (long)b * in_elems, output + (long)b * out_elems, buf, tap_slot,
                      tap_out ? (char*)tap_out + (long)b * tap_bytes : NULL);
      if (r) {
#pragma omp critical
        rc = r;
      }
    }
    for (uint32_t i = 0; i < h->n_tensors; i++) free(buf[i]);
    free(buf);
  }
  return rc;
}

/* ------------------------------------------------------------------------- */
/* host frontend: hybrid (linear |STFT| + min-max)                            */
/* ------------------------------------------------------------------------- */
/* In-place iterative radix-2 complex FFT in double precision. */
static void fft_c2c(double* re, double* im, int n, const double* cs, const double* sn) {
  for (int i = 1, j = 0; i < n; i++) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { double t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
  }
  for (int len = 2; len <= n; len <<= 1) {
    int step = n / len;
    for (int i = 0; i < n; i += len)
      for (int k = 0; k < len / 2; k++) {
        double wr = cs[k * step], wi = -sn[k * step];
        double ur = re[i + k], ui = im[i + k];
        double vr = re[i + k + len / 2] * wr - im[i + k + len / 2] * wi;
        double vi = re[i + k + len / 2] * wi + im[i + k + len / 2] * wr;
        re[i + k] = ur + vr; im[i + k] = ui + vi;
        re[i + k + len / 2] = ur - vr; im[i + k + len / 2] = ui - vi;
      }
  }
}

/* PCM16 -> float32 samples exactly as the reference's decode + peak normalisation:
 * soundfile gives int16/32768 as float32 (audio/io.py:114-116), then y / peak in float32
 * (audio/io.py:124-126).  peak <= 0 means "no normalisation". */
static inline float pcm_to_f32(int16_t s, float peak) {
  float v = (float)s / 32768.0f;
  return peak > 0.0f ? v / peak : v;
}

/* |STFT| of one chunk: librosa.stft(n_fft, hop, win_length=n_fft, window="hann", center=True,
 * pad_mode="constant") -> np.abs -> [:, :spec_width]   (audio/spectrogram.py:106-115,133).
 * x: float32 samples [T].  mag: float32 [n_fft/2+1, spec_width] (bin-major like librosa). */
static void stft_mag(const float* x, int T, int n_fft, int hop, int spec_width, float* mag,
                     const double* win, const double* cs, const double* sn) {
  const int bins = n_fft / 2 + 1, half = n_fft / 2;
  double* re = (double*)malloc(sizeof(double) * n_fft);
  double* im = (double*)malloc(sizeof(double) * n_fft);
  for (int t = 0; t < spec_width; t++) {
    for (int n = 0; n < n_fft; n++) {
      long idx = (long)t * hop + n - half;          /* centred, zero padded */
      double v = (idx >= 0 && idx < T) ? (double)x[idx] : 0.0;
      re[n] = win[n] * v;                            /* float64 window * float32 sample */
      im[n] = 0.0;
    }
    fft_c2c(re, im, n_fft, cs, sn);
    for (int k = 0; k < bins; k++) {
      float r32 = (float)re[k], i32 = (float)im[k];  /* stored as complex64 */
      mag[(long)k * spec_width + t] = hypotf(r32, i32);  /* np.abs(complex64) */
    }
  }
  free(re); free(im);
}

/* normalize(): (S - S.min()) / (S.max() - S.min() + 1e-10)  (audio/spectrogram.py:12-21), numpy 1.26
 * promotion: the scalar denominator is formed in float64 then cast to float32 for the array divide. */
static void minmax_normalize(float* s, long n) {
  float mn = s[0], mx = s[0];
  for (long i = 1; i < n; i++) { if (s[i] < mn) mn = s[i]; if (s[i] > mx) mx = s[i]; }
  float range = mx - mn;
  float den = (float)((double)range + 1e-10);
  for (long i = 0; i < n; i++) s[i] = (s[i] - mn) / den;
}

/* Frontend for `hybrid` models: out float32 [B, n_fft/2+1, spec_width] in [0,1].
 * Restates make_chunks_for_file's hybrid branch (evaluation/metrics.py:55-61). */
BNO_EXPORT int bno_frontend_hybrid(const int16_t* pcm, const float* peak, int B, int T, int n_fft,
                                   int hop, int spec_width, float* out, int threads) {
  if (n_fft <= 0 || (n_fft & (n_fft - 1))) { snprintf(g_err, sizeof g_err, "n_fft must be a power of two"); return -1; }
  const int bins = n_fft / 2 + 1;
  double* win = (double*)malloc(sizeof(double) * n_fft);
  double* cs = (double*)malloc(sizeof(double) * n_fft);
  double* sn = (double*)malloc(sizeof(double) * n_fft);
  for (int n = 0; n < n_fft; n++) {
    win[n] = 0.5 - 0.5 * cos(2.0 * M_PI * (double)n / (double)n_fft);  /* periodic Hann */
    cs[n] = cos(2.0 * M_PI * (double)n / (double)n_fft);
    sn[n] = sin(2.0 * M_PI * (double)n / (double)n_fft);
  }
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int b = 0; b < B; b++) {
    float* x = (float*)malloc(sizeof(float) * T);
    for (int i = 0; i < T; i++) x[i] = pcm_to_f32(pcm[(long)b * T + i], peak ? peak[b] : 0.0f);
    float* s = out + (long)b * bins * spec_width;
    stft_mag(x, T, n_fft, hop, spec_width, s, win, cs, sn);
    minmax_normalize(s, (long)bins * spec_width);
    free(x);
  }
  free(win); free(cs); free(sn);
  return 0;
}

/* Same, from float32 waveforms (chunks that went through host resampling). */
BNO_EXPORT int bno_frontend_hybrid_f32(const float* wav, int B, int T, int n_fft, int hop,
                                       int spec_width, float* out, int threads) {
  const int bins = n_fft / 2 + 1;
  double* win = (double*)malloc(sizeof(double) * n_fft);
  double* cs = (double*)malloc(sizeof(double) * n_fft);
  double* sn = (double*)malloc(sizeof(double) * n_fft);
  for (int n = 0; n < n_fft; n++) {
    win[n] = 0.5 - 0.5 * cos(2.0 * M_PI * (double)n / (double)n_fft);
    cs[n] = cos(2.0 * M_PI * (double)n / (double)n_fft);
    sn[n] = sin(2.0 * M_PI * (double)n / (double)n_fft);
  }
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int b = 0; b < B; b++) {
    float* s = out + (long)b * bins * spec_width;
    stft_mag(wav + (long)b * T, T, n_fft, hop, spec_width, s, win, cs, sn);
    minmax_normalize(s, (long)bins * spec_width);
  }
  free(win); free(cs); free(sn);
  return 0;
}

/* Raw frontend input: x[:T] / (max|x| + 1e-6)  (evaluation/metrics.py:62-69), float32 maths
 * (numpy: float32 array / float64 scalar -> float32 array; the scalar is cast to float32). */
BNO_EXPORT int bno_frontend_raw(const int16_t* pcm, const float* peak, int B, int T, float* out) {
  for (int b = 0; b < B; b++) {
    float mx = 0.0f;
    float* o = out + (long)b * T;
    for (int i = 0; i < T; i++) { o[i] = pcm_to_f32(pcm[(long)b * T + i], peak ? peak[b] : 0.0f); float a = fabsf(o[i]); if (a > mx) mx = a; }
    float den = (float)((double)mx + 1e-6);
    for (int i = 0; i < T; i++) o[i] = o[i] / den;
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* pooling  (evaluation/pooling.py:6-47), float32 like numpy on float32 input */
/* ------------------------------------------------------------------------- */
/* method: 0 avg, 1 max, 2 lme.  scores [N, C] -> out [C].  N == 0 -> zeros. */
BNO_EXPORT int bno_pool(const float* scores, int N, int C, int method, float beta, float* out) {
  if (N == 0) { for (int c = 0; c < C; c++) out[c] = 0.0f; return 0; }
  for (int c = 0; c < C; c++) {
    if (method == 0) {
      /* np.mean float32: pairwise summation; N <= 8 blocks -> plain left-to-right float32 adds */
      float s = 0.0f;
      for (int i = 0; i < N; i++) s += scores[(long)i * C + c];
      out[c] = s / (float)N;
    } else if (method == 1) {
      float mx = scores[c];
      for (int i = 1; i < N; i++) if (scores[(long)i * C + c] > mx) mx = scores[(long)i * C + c];
      out[c] = mx;
    } else if (method == 2) {
      float mx = beta * scores[c];
      for (int i = 1; i < N; i++) { float v = beta * scores[(long)i * C + c]; if (v > mx) mx = v; }
      float s = 0.0f;
      for (int i = 0; i < N; i++) s += expf(beta * scores[(long)i * C + c] - mx);
      float mean = s / (float)N;
      out[c] = (mx + logf(mean + 1e-12f)) / beta;
    } else { snprintf(g_err, sizeof g_err, "Unsupported pooling method: %d", method); return -1; }
  }
  return 0;
}

/* LOGISTIC LUT as TFLite builds it (LUTPopulate<int8_t>, float32, glibc expf) -- used by the
 * tests to cross-check the table the exporter wrote into the blob. */
BNO_EXPORT void bno_logistic_lut(float in_scale, int in_zp, float out_scale, int out_zp, int8_t* lut) {
  const float inv = 1.0f / out_scale;
  for (int q = -128; q < 128; q++) {
    float x = in_scale * (float)(q - in_zp);
    float y = 1.0f / (1.0f + expf(-x));
    int32_t r = (int32_t)roundf(y * inv) + out_zp;
    lut[q + 128] = (int8_t)clampi(r, -128, 127);
  }
}

/* The closed form the CUDA kernels use for a right shift n in [1,31] (bn_common.cuh: rq_fast); exported so a CPU test
 * can prove it equal to the gemmlowp sequence above for all shifts. */
BNO_EXPORT int32_t bno_rq_fast(int32_t acc, int32_t mult, int n) {
  const int64_t p = (int64_t)acc * (int64_t)mult + ((int64_t)1 << 30);
  const int32_t v = (int32_t)(p >> 31);
  return (v + (1 << (n - 1)) + (v >> 31)) >> n;
}

/* expose the fixed-point primitives for unit tests */
BNO_EXPORT int32_t bno_mbqm(int32_t x, int32_t qm, int shift, int rounding) { return mbqm(x, qm, shift, rounding); }
