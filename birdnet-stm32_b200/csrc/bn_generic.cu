// bn_generic.cu -- one CUDA kernel per lowered op (the "generic plan").
//
// Every op of include/bn_blob.h has a kernel here, one thread per output element
// (grid-stride), NHWC int8 activations for a whole wave of chunks.  This plan runs any
// supported graph, keeps every tensor materialised for the debug taps, and is the
// bit-exact reference the fused plan (bn_fast.cu) is checked against on the GPU.
// Semantics: TFLite reference integer kernels, executed by the
// reference inside tf.lite.Interpreter.invoke (birdnet_stm32/models/runners.py:93-95).
#include "bn_common.cuh"
#include "bn_kernels.cuh"

namespace bn {

static inline int grid_for(long n, int block = 256) {
  long g = (n + block - 1) / block;
  const long cap = 148L * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

#define GRID_STRIDE(i, n) for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (long)gridDim.x * blockDim.x)

__global__ void k_quantize(const float* __restrict__ x, int8_t* __restrict__ y, long n, float scale, int zp) {
  GRID_STRIDE(i, n) {
    // AffineQuantize: TfLiteRound(val / scale) + zp, IEEE float32 division, round half away
    float q = roundf(__fdiv_rn(x[i], scale));
    int v = (int)q + zp;
    y[i] = (int8_t)clampi(v, -128, 127);
  }
}

__global__ void k_dequantize(const int8_t* __restrict__ x, float* __restrict__ y, long n, float scale, int zp) {
  GRID_STRIDE(i, n) y[i] = __fmul_rn(scale, (float)((int)x[i] - zp));
}

__global__ void k_requant(const int8_t* __restrict__ x, int8_t* __restrict__ y, long n, int in_zp, int out_zp,
                          int mult, int shift, int R) {
  GRID_STRIDE(i, n) y[i] = (int8_t)clampi(mbqm((int)x[i] - in_zp, mult, shift, R) + out_zp, -128, 127);
}

__global__ void k_transpose(const int8_t* __restrict__ x, int8_t* __restrict__ y, long n, int id0, int id1, int id2,
                            int od0, int od1, int od2, int p0, int p1, int p2) {
  const long is[3] = {(long)id1 * id2, id2, 1};
  const long per = (long)od0 * od1 * od2;
  GRID_STRIDE(i, n) {
    long b = i / per, r = i - b * per;
    int c = (int)(r % od2);
    int bb = (int)((r / od2) % od1);
    int a = (int)(r / ((long)od2 * od1));
    long src = a * is[p0] + bb * is[p1] + c * is[p2];
    y[i] = x[b * per + src];
  }
}

__global__ void k_slice(const int8_t* __restrict__ x, int8_t* __restrict__ y, long n, int id0, int id1, int id2,
                        int od0, int od1, int od2, int b0, int b1, int b2) {
  const long per_o = (long)od0 * od1 * od2, per_i = (long)id0 * id1 * id2;
  GRID_STRIDE(i, n) {
    long b = i / per_o, r = i - b * per_o;
    int c = (int)(r % od2);
    int bb = (int)((r / od2) % od1);
    int a = (int)(r / ((long)od2 * od1));
    y[i] = x[b * per_i + ((long)(a + b0) * id1 + (bb + b1)) * id2 + (c + b2)];
  }
}

__global__ void k_fill(int8_t* __restrict__ y, long n, int val) {
  GRID_STRIDE(i, n) y[i] = (int8_t)val;
}

// y[o, 0:in0] = x0[o, :], y[o, in0:] = x1[o, :] with o over B*outer rows
__global__ void k_concat(const int8_t* __restrict__ x0, const int8_t* __restrict__ x1, int8_t* __restrict__ y,
                         long rows, int in0, int in1) {
  const long w = in0 + in1, n = rows * w;
  GRID_STRIDE(i, n) {
    long o = i / w;
    int j = (int)(i - o * w);
    y[i] = j < in0 ? x0[o * in0 + j] : x1[o * in1 + (j - in0)];
  }
}

// reference_integer_ops::ConvPerChannel, one thread per output element
__global__ void k_conv2d(const int8_t* __restrict__ x, int8_t* __restrict__ y, long n, ConvParams P) {
  const bool vec = (P.ic % 4) == 0;
  GRID_STRIDE(i, n) {
    int co = (int)(i % P.oc);
    long r = i / P.oc;
    int ox = (int)(r % P.ow);
    r /= P.ow;
    int oy = (int)(r % P.oh);
    long b = r / P.oh;
    const int8_t* xb = x + b * ((long)P.ih * P.iw * P.ic);
    int acc = 0;
    for (int fy = 0; fy < P.kh; fy++) {
      int iy = oy * P.sh - P.pt + fy;
      if (iy < 0 || iy >= P.ih) continue;
      for (int fx = 0; fx < P.kw; fx++) {
        int ix = ox * P.sw - P.pl + fx;
        if (ix < 0 || ix >= P.iw) continue;
        const int8_t* xp = xb + ((long)iy * P.iw + ix) * P.ic;
        const int8_t* wp = P.w + (((long)co * P.kh + fy) * P.kw + fx) * P.ic;
        if (vec) {
          const int* x4 = reinterpret_cast<const int*>(xp);
          const int* w4 = reinterpret_cast<const int*>(wp);
          int a = 0, ws = 0;
          for (int k = 0; k < P.ic / 4; k++) {
            int wv = __ldg(w4 + k);
            a = __dp4a(x4[k], wv, a);
            ws = __dp4a(0x01010101, wv, ws);
          }
          acc += a - P.in_zp * ws;
        } else {
          for (int ci = 0; ci < P.ic; ci++) acc += ((int)xp[ci] - P.in_zp) * (int)__ldg(wp + ci);
        }
      }
    }
    acc += __ldg(P.bias + co);
    acc = mbqm(acc, __ldg(P.mult + co), __ldg(P.shift + co), P.rounding) + P.out_zp;
    y[i] = (int8_t)clampi(acc, P.act_min, P.act_max);
  }
}

// reference_integer_ops::DepthwiseConvPerChannel (depth_multiplier 1), weights [kh,kw,C]
__global__ void k_dwconv2d(const int8_t* __restrict__ x, int8_t* __restrict__ y, long n, ConvParams P) {
  GRID_STRIDE(i, n) {
    int c = (int)(i % P.oc);
    long r = i / P.oc;
    int ox = (int)(r % P.ow);
    r /= P.ow;
    int oy = (int)(r % P.oh);
    long b = r / P.oh;
    const int8_t* xb = x + b * ((long)P.ih * P.iw * P.ic);
    int acc = 0;
    for (int fy = 0; fy < P.kh; fy++) {
      int iy = oy * P.sh - P.pt + fy;
      if (iy < 0 || iy >= P.ih) continue;
      for (int fx = 0; fx < P.kw; fx++) {
        int ix = ox * P.sw - P.pl + fx;
        if (ix < 0 || ix >= P.iw) continue;
        acc += ((int)xb[((long)iy * P.iw + ix) * P.ic + c] - P.in_zp) * (int)__ldg(P.w + ((long)fy * P.kw + fx) * P.ic + c);
      }
    }
    acc += __ldg(P.bias + c);
    acc = mbqm(acc, __ldg(P.mult + c), __ldg(P.shift + c), P.rounding) + P.out_zp;
    y[i] = (int8_t)clampi(acc, P.act_min, P.act_max);
  }
}

// reference_integer_ops::FullyConnectedPerChannel: x [rows,K], w [N,K]
__global__ void k_fc(const int8_t* __restrict__ x, int8_t* __restrict__ y, long n, ConvParams P) {
  const int K = P.ic, N = P.oc;
  const bool vec = (K % 4) == 0;
  GRID_STRIDE(i, n) {
    int o = (int)(i % N);
    long row = i / N;
    const int8_t* xp = x + row * K;
    const int8_t* wp = P.w + (long)o * K;
    int acc = 0;
    if (vec) {
      const int* x4 = reinterpret_cast<const int*>(xp);
      const int* w4 = reinterpret_cast<const int*>(wp);
      int a = 0, ws = 0;
      for (int k = 0; k < K / 4; k++) {
        int wv = __ldg(w4 + k);
        a = __dp4a(x4[k], wv, a);
        ws = __dp4a(0x01010101, wv, ws);
      }
      acc = a - P.in_zp * ws;
    } else {
      for (int k = 0; k < K; k++) acc += ((int)xp[k] - P.in_zp) * (int)__ldg(wp + k);
    }
    acc += __ldg(P.bias + o);
    acc = mbqm(acc, __ldg(P.mult + o), __ldg(P.shift + o), P.rounding) + P.out_zp;
    y[i] = (int8_t)clampi(acc, P.act_min, P.act_max);
  }
}

// reference_integer_ops::Add.  bcast: 0 same shape, 1 const [C] over last dim, 2 per-chunk [C]
__global__ void k_add(const int8_t* __restrict__ a, const int8_t* __restrict__ b, int8_t* __restrict__ y, long n,
                      long per_chunk, AddParams P) {
  GRID_STRIDE(i, n) {
    long bi = P.bcast == 0 ? i : (P.bcast == 1 ? (i % P.C) : ((i / per_chunk) * P.C + (i % P.C)));
    int v1 = (int)a[i] - P.in1_zp;
    int v2 = (int)b[bi] - P.in2_zp;
    int s1 = mbqm(v1 << P.left_shift, P.m1, P.s1, P.rounding);
    int s2 = mbqm(v2 << P.left_shift, P.m2, P.s2, P.rounding);
    int o = mbqm(s1 + s2, P.mo, P.so, P.rounding) + P.out_zp;
    y[i] = (int8_t)clampi(o, P.act_min, P.act_max);
  }
}

// reference_integer_ops::Mul
__global__ void k_mul(const int8_t* __restrict__ a, const int8_t* __restrict__ b, int8_t* __restrict__ y, long n,
                      long per_chunk, int in1_zp, int in2_zp, int out_zp, int mult, int shift, int act_min,
                      int act_max, int bcast, int C, int R) {
  GRID_STRIDE(i, n) {
    // bcast: 0 none, 1 const [C], 2 per-chunk [1,1,C] over H,W (SE gate), 3 per-position [H,W,1] over C (attention weights)
    long bi = bcast == 0 ? i : (bcast == 1 ? (i % C) : (bcast == 2 ? ((i / per_chunk) * C + (i % C)) : (i / C)));
    int v = ((int)a[i] - in1_zp) * ((int)b[bi] - in2_zp);
    int o = mbqm(v, mult, shift, R) + out_zp;
    y[i] = (int8_t)clampi(o, act_min, act_max);
  }
}

// Four elements per thread (one 32-bit word of each operand) and 32-bit index arithmetic: same arithmetic per element as
// k_add / k_mul.  Needs n, C and per_chunk to be multiples of 4, word-aligned pointers and n < 2^31.
__global__ void k_add4(const unsigned* __restrict__ a, const int8_t* __restrict__ b, unsigned* __restrict__ y, unsigned n4,
                       unsigned per_chunk, AddParams P) {
  for (unsigned w = blockIdx.x * blockDim.x + threadIdx.x; w < n4; w += gridDim.x * blockDim.x) {
    const unsigned i = 4u * w;
    const unsigned bi = P.bcast == 0 ? i : (P.bcast == 1 ? (i % (unsigned)P.C) : ((i / per_chunk) * (unsigned)P.C + (i % (unsigned)P.C)));
    const unsigned av = a[w], bv = *reinterpret_cast<const unsigned*>(b + bi);
    unsigned o4 = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int v1 = (int)(int8_t)(av >> (8 * j)) - P.in1_zp;
      const int v2 = (int)(int8_t)(bv >> (8 * j)) - P.in2_zp;
      const int s1 = mbqm(v1 << P.left_shift, P.m1, P.s1, P.rounding);
      const int s2 = mbqm(v2 << P.left_shift, P.m2, P.s2, P.rounding);
      const int o = mbqm(s1 + s2, P.mo, P.so, P.rounding) + P.out_zp;
      o4 |= (unsigned)(uint8_t)clampi(o, P.act_min, P.act_max) << (8 * j);
    }
    y[w] = o4;
  }
}

__global__ void k_mul4(const unsigned* __restrict__ a, const int8_t* __restrict__ b, unsigned* __restrict__ y, unsigned n4,
                       unsigned per_chunk, int in1_zp, int in2_zp, int out_zp, int mult, int shift, int act_min,
                       int act_max, int bcast, int C, int R) {
  for (unsigned w = blockIdx.x * blockDim.x + threadIdx.x; w < n4; w += gridDim.x * blockDim.x) {
    const unsigned i = 4u * w;
    const unsigned av = a[w];
    unsigned bv;
    if (bcast == 3) bv = 0x01010101u * (unsigned)(uint8_t)b[i / (unsigned)C];
    else {
      const unsigned bi = bcast == 0 ? i : (bcast == 1 ? (i % (unsigned)C) : ((i / per_chunk) * (unsigned)C + (i % (unsigned)C)));
      bv = *reinterpret_cast<const unsigned*>(b + bi);
    }
    unsigned o4 = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int v = ((int)(int8_t)(av >> (8 * j)) - in1_zp) * ((int)(int8_t)(bv >> (8 * j)) - in2_zp);
      const int o = mbqm(v, mult, shift, R) + out_zp;
      o4 |= (unsigned)(uint8_t)clampi(o, act_min, act_max) << (8 * j);
    }
    y[w] = o4;
  }
}

// Sixteen elements per thread (one 16-byte vector of each operand): MUL with n, C and per_chunk multiples of 16.
__global__ void k_mul16(const uint4* __restrict__ a, const int8_t* __restrict__ b, uint4* __restrict__ y, unsigned n16,
                        unsigned per_chunk, int in1_zp, int in2_zp, int out_zp, int mult, int shift, int act_min,
                        int act_max, int bcast, int C, int R) {
  for (unsigned v = blockIdx.x * blockDim.x + threadIdx.x; v < n16; v += gridDim.x * blockDim.x) {
    const unsigned i = 16u * v;
    const uint4 av = a[v];
    uint4 bv;
    if (bcast == 3) { const unsigned w = 0x01010101u * (unsigned)(uint8_t)b[i / (unsigned)C]; bv = make_uint4(w, w, w, w); }
    else {
      const unsigned bi = bcast == 0 ? i : (bcast == 1 ? (i % (unsigned)C) : ((i / per_chunk) * (unsigned)C + (i % (unsigned)C)));
      bv = *reinterpret_cast<const uint4*>(b + bi);
    }
    const unsigned aw[4] = {av.x, av.y, av.z, av.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
    unsigned ow[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      unsigned o4 = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int p = ((int)(int8_t)(aw[q] >> (8 * j)) - in1_zp) * ((int)(int8_t)(bw[q] >> (8 * j)) - in2_zp);
        const int o = mbqm(p, mult, shift, R) + out_zp;
        o4 |= (unsigned)(uint8_t)clampi(o, act_min, act_max) << (8 * j);
      }
      ow[q] = o4;
    }
    y[v] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

static inline bool words_ok(const void* a, const void* b, const void* y, long n, long per_chunk, int C) {
  return n > 0 && n < (1l << 31) && n % 4 == 0 && per_chunk % 4 == 0 && C % 4 == 0 &&
         (((uintptr_t)a | (uintptr_t)b | (uintptr_t)y) & 3) == 0;
}

// PAD with the zero point: x [B][d0][d1][d2] -> y [B][o0][o1][o2], o = d + before + after
__global__ void k_pad(const int8_t* __restrict__ x, int8_t* __restrict__ y, long n, int d0, int d1, int d2, int o0, int o1, int o2,
                      int b0, int b1, int b2, int val) {
  GRID_STRIDE(i, n) {
    long r = i;
    const int c = (int)(r % o2); r /= o2;
    const int w = (int)(r % o1); r /= o1;
    const int h = (int)(r % o0); r /= o0;
    const int sh = h - b0, sw = w - b1, sc = c - b2;
    int v = val;
    if (sh >= 0 && sh < d0 && sw >= 0 && sw < d1 && sc >= 0 && sc < d2) v = x[((r * d0 + sh) * d1 + sw) * (long)d2 + sc];
    y[i] = (int8_t)v;
  }
}

// SOFTMAX over the last dim, int8 -> int8: the float-LUT kernel of tflite::optimized_ops::Softmax
// (table[255 - v] = expf(-in_scale * beta * v) comes with the blob; sums run left to right in float32).
__global__ void k_softmax(const int8_t* __restrict__ x, int8_t* __restrict__ y, long rows, int L, const float* __restrict__ table,
                          float out_scale, int out_zp) {
  GRID_STRIDE(r, rows) {
    const int8_t* xp = x + r * (long)L;
    int mx = -128;
    for (int j = 0; j < L; j++) mx = max(mx, (int)xp[j]);
    const float* toff = table + (255 - mx);
    float sum = 0.0f;
    for (int j = 0; j < L; j++) sum = __fadd_rn(sum, toff[xp[j]]);
    const float inv = __fdiv_rn(1.0f, __fmul_rn(sum, out_scale));
    for (int j = 0; j < L; j++) {
      const float pr = __fmul_rn(toff[xp[j]], inv);
      const int q = (int)__fadd_rn(pr, 0.5f) + out_zp;
      y[r * (long)L + j] = (int8_t)max(-128, min(127, q));
    }
  }
}

// SUM over one non-batch axis, int8 -> int8: tflite::reference_ops::QuantizedMeanOrSum(compute_sum = true):
// round(float(sum) * scale + bias) + out_zp with scale = in_scale / out_scale, bias = -in_zp * scale * count.
__global__ void k_sum(const int8_t* __restrict__ x, int8_t* __restrict__ y, long n, int outer, int count, int inner, float scale,
                      float bias, int out_zp) {
  GRID_STRIDE(i, n) {
    const long in_ = i % inner;
    const long ob = i / inner;                       // (batch * outer) index
    const int8_t* xp = x + (ob * count) * inner + in_;
    int sum = 0;
    for (int j = 0; j < count; j++) sum += xp[(long)j * inner];
    const float v = __fadd_rn(__fmul_rn((float)sum, scale), bias);
    const int q = (int)roundf(v) + out_zp;
    y[i] = (int8_t)max(-128, min(127, q));
  }
}

// MEAN over H,W: x [B, N, C] -> y [B, C]; variants of SURVEY.md Appendix B.6
__global__ void k_mean(const int8_t* __restrict__ x, int8_t* __restrict__ y, long n, int N, int C, int in_zp,
                       int out_zp, int mult, int shift, int mult_n, int shift_n, float in_scale, float out_scale,
                       int variant, int R) {
  GRID_STRIDE(i, n) {
    long b = i / C;
    int c = (int)(i - b * C);
    const int8_t* xp = x + b * (long)N * C + c;
    int sum = 0;
    for (int j = 0; j < N; j++) sum += xp[(long)j * C];
    int o;
    if (variant == 1) {
      float scale = __fdiv_rn(in_scale, out_scale);
      float bias = __fmul_rn(-(float)in_zp, scale);
      float fm = __fdiv_rn((float)sum, (float)N);
      o = (int)roundf(__fadd_rn(__fmul_rn(fm, scale), bias)) + out_zp;
    } else if (variant == 2) {
      o = mbqm(sum - in_zp * N, mult_n, shift_n, R) + out_zp;
    } else {
      int acc = mbqm(sum - in_zp * N, mult, shift, R);
      acc = acc > 0 ? (acc + N / 2) / N : (acc - N / 2) / N;
      o = acc + out_zp;
    }
    y[i] = (int8_t)clampi(o, -128, 127);
  }
}

__global__ void k_logistic(const int8_t* __restrict__ x, int8_t* __restrict__ y, long n, const int8_t* __restrict__ lut) {
  GRID_STRIDE(i, n) y[i] = __ldg(lut + (uint8_t)(x[i] + 128));
}

// normalize(): (S - min) / (max - min + 1e-10) per chunk (audio/spectrogram.py:12-21)
__global__ void k_minmax_normalize(float* __restrict__ s, long per_chunk, long n, const unsigned* __restrict__ mnmx) {
  GRID_STRIDE(i, n) {
    long b = i / per_chunk;
    float mn = __uint_as_float(mnmx[2 * b]), mx = __uint_as_float(mnmx[2 * b + 1]);
    float den = (float)((double)(mx - mn) + 1e-10);
    s[i] = __fdiv_rn(s[i] - mn, den);
  }
}

// pool_scores (evaluation/pooling.py:25-47): one block per file, one thread per class,
// chunks accumulated row by row in float32 like numpy's axis-0 reduction.
__global__ void k_pool(const float* __restrict__ scores, const int* __restrict__ offs, float* __restrict__ out, int C,
                       int method, float beta) {
  const int f = blockIdx.x;
  const int s = offs[f], e = offs[f + 1];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float r = 0.0f;
    if (e > s) {
      if (method == BN_POOL_AVG) {
        float acc = 0.0f;
        for (int i = s; i < e; i++) acc = __fadd_rn(acc, scores[(long)i * C + c]);
        r = __fdiv_rn(acc, (float)(e - s));
      } else if (method == BN_POOL_MAX) {
        float m = scores[(long)s * C + c];
        for (int i = s + 1; i < e; i++) m = fmaxf(m, scores[(long)i * C + c]);
        r = m;
      } else {
        float m = __fmul_rn(beta, scores[(long)s * C + c]);
        for (int i = s + 1; i < e; i++) m = fmaxf(m, __fmul_rn(beta, scores[(long)i * C + c]));
        float acc = 0.0f;
        for (int i = s; i < e; i++) acc = __fadd_rn(acc, expf(__fsub_rn(__fmul_rn(beta, scores[(long)i * C + c]), m)));
        float mean = __fdiv_rn(acc, (float)(e - s));
        r = __fdiv_rn(__fadd_rn(m, logf(__fadd_rn(mean, 1e-12f))), beta);
      }
    }
    out[(long)f * C + c] = r;
  }
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
#define LAUNCH(kern, n, stream, ...) kern<<<grid_for(n), 256, 0, stream>>>(__VA_ARGS__)

void launch_quantize(const float* x, int8_t* y, long n, float scale, int zp, cudaStream_t st) { LAUNCH(k_quantize, n, st, x, y, n, scale, zp); }
void launch_dequantize(const int8_t* x, float* y, long n, float scale, int zp, cudaStream_t st) { LAUNCH(k_dequantize, n, st, x, y, n, scale, zp); }
void launch_requant(const int8_t* x, int8_t* y, long n, int in_zp, int out_zp, int mult, int shift, int R, cudaStream_t st) { LAUNCH(k_requant, n, st, x, y, n, in_zp, out_zp, mult, shift, R); }
void launch_transpose(const int8_t* x, int8_t* y, long n, const int* id, const int* od, const int* p, cudaStream_t st) {
  LAUNCH(k_transpose, n, st, x, y, n, id[0], id[1], id[2], od[0], od[1], od[2], p[0], p[1], p[2]);
}
void launch_slice(const int8_t* x, int8_t* y, long n, const int* id, const int* od, const int* b, cudaStream_t st) {
  LAUNCH(k_slice, n, st, x, y, n, id[0], id[1], id[2], od[0], od[1], od[2], b[0], b[1], b[2]);
}
void launch_fill(int8_t* y, long n, int val, cudaStream_t st) { LAUNCH(k_fill, n, st, y, n, val); }
void launch_concat(const int8_t* x0, const int8_t* x1, int8_t* y, long rows, int in0, int in1, cudaStream_t st) {
  long n = rows * (in0 + in1);
  LAUNCH(k_concat, n, st, x0, x1, y, rows, in0, in1);
}
void launch_conv2d(const int8_t* x, int8_t* y, long n, const ConvParams& P, cudaStream_t st) { LAUNCH(k_conv2d, n, st, x, y, n, P); }
void launch_dwconv2d(const int8_t* x, int8_t* y, long n, const ConvParams& P, cudaStream_t st) { LAUNCH(k_dwconv2d, n, st, x, y, n, P); }
void launch_fc(const int8_t* x, int8_t* y, long n, const ConvParams& P, cudaStream_t st) { LAUNCH(k_fc, n, st, x, y, n, P); }
void launch_add(const int8_t* a, const int8_t* b, int8_t* y, long n, long per_chunk, const AddParams& P, cudaStream_t st) {
  if (words_ok(a, b, y, n, per_chunk, P.C)) {
    LAUNCH(k_add4, n / 4, st, reinterpret_cast<const unsigned*>(a), b, reinterpret_cast<unsigned*>(y), (unsigned)(n / 4), (unsigned)per_chunk, P);
    return;
  }
  LAUNCH(k_add, n, st, a, b, y, n, per_chunk, P);
}
void launch_mul(const int8_t* a, const int8_t* b, int8_t* y, long n, long per_chunk, const int* p, int C, int R, cudaStream_t st) {
  if (words_ok(a, b, y, n, per_chunk, C) && n % 16 == 0 && per_chunk % 16 == 0 && C % 16 == 0 &&
      (((uintptr_t)a | (uintptr_t)b | (uintptr_t)y) & 15) == 0) {
    LAUNCH(k_mul16, n / 16, st, reinterpret_cast<const uint4*>(a), b, reinterpret_cast<uint4*>(y), (unsigned)(n / 16), (unsigned)per_chunk,
           p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], C, R);
    return;
  }
  if (words_ok(a, b, y, n, per_chunk, C)) {
    LAUNCH(k_mul4, n / 4, st, reinterpret_cast<const unsigned*>(a), b, reinterpret_cast<unsigned*>(y), (unsigned)(n / 4), (unsigned)per_chunk,
           p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], C, R);
    return;
  }
  LAUNCH(k_mul, n, st, a, b, y, n, per_chunk, p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], C, R);
}
void launch_mean(const int8_t* x, int8_t* y, long n, int N, int C, const int* p, float in_scale, float out_scale, int variant, int R, cudaStream_t st) {
  LAUNCH(k_mean, n, st, x, y, n, N, C, p[BN_MEAN_IN_ZP], p[BN_MEAN_OUT_ZP], p[BN_MEAN_MULT], p[BN_MEAN_SHIFT],
         p[BN_MEAN_MULT_N], p[BN_MEAN_SHIFT_N], in_scale, out_scale, variant, R);
}
void launch_logistic(const int8_t* x, int8_t* y, long n, const int8_t* lut, cudaStream_t st) { LAUNCH(k_logistic, n, st, x, y, n, lut); }
void launch_pad(const int8_t* x, int8_t* y, long n, const int* id, const int* od, const int* p, cudaStream_t st) {
  LAUNCH(k_pad, n, st, x, y, n, id[0], id[1], id[2], od[0], od[1], od[2], p[0], p[2], p[4], p[6]);
}
void launch_softmax(const int8_t* x, int8_t* y, long rows, int L, const float* table, float out_scale, int out_zp, cudaStream_t st) {
  LAUNCH(k_softmax, rows, st, x, y, rows, L, table, out_scale, out_zp);
}
void launch_sum(const int8_t* x, int8_t* y, long n, int outer, int count, int inner, float scale, float bias, int out_zp, cudaStream_t st) {
  LAUNCH(k_sum, n, st, x, y, n, outer, count, inner, scale, bias, out_zp);
}
void launch_minmax_normalize(float* s, long per_chunk, long n, const unsigned* mnmx, cudaStream_t st) { LAUNCH(k_minmax_normalize, n, st, s, per_chunk, n, mnmx); }
void launch_pool(const float* scores, const int* offs, float* out, int F, int C, int method, float beta, cudaStream_t st) {
  if (F <= 0) return;
  k_pool<<<F, 128, 0, st>>>(scores, offs, out, C, method, beta);
}

}  // namespace bn
