// bn_common.cuh -- shared device helpers: TFLite fixed-point requantisation on the GPU.
//
// The arithmetic follows gemmlowp's fixedpoint.h / TFLite common.h
// (MultiplyByQuantizedMultiplier) -- third-party code the reference executes inside
// tf.lite.Interpreter.invoke (birdnet_stm32/models/runners.py:93-95).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bn_blob.h"
#include "../../include/bn_engine.h"

namespace bn {

// SaturatingRoundingDoublingHighMul for a multiplier in [0, 2^31): the saturating case
// (both INT32_MIN) cannot occur, and gemmlowp's sign-dependent nudge followed by C's
// truncating division equals (a*b + 2^30) >> 31 with an arithmetic shift (proof in DESIGN.md).
__host__ __device__ __forceinline__ int32_t srdhm(int32_t a, int32_t b) {
  long long ab = (long long)a * (long long)b;
  return (int32_t)((ab + (1ll << 30)) >> 31);
}

// RoundingDivideByPOT: round half away from zero.
__host__ __device__ __forceinline__ int32_t rdivpot(int32_t x, int e) {
  int32_t mask = (int32_t)((1ll << e) - 1);
  int32_t rem = x & mask;
  int32_t thr = (mask >> 1) + (x < 0 ? 1 : 0);
  return (x >> e) + (rem > thr ? 1 : 0);
}

// tflite::MultiplyByQuantizedMultiplier. rounding 0 = double rounding (default TFLite build),
// 1 = single rounding (TFLITE_SINGLE_ROUNDING / ruy).
__host__ __device__ __forceinline__ int32_t mbqm(int32_t x, int32_t qm, int shift, int rounding) {
  if (rounding == 0) {
    int left = shift > 0 ? shift : 0;
    int right = shift > 0 ? 0 : -shift;
    int32_t xs = (int32_t)((uint32_t)x << left);   // int32 wrap like x * (1 << left)
    return rdivpot(srdhm(xs, qm), right);
  } else {
    int total = 31 - shift;
    long long r = (long long)x * (long long)qm + (1ll << (total - 1));
    r >>= total;
    r = r > 2147483647ll ? 2147483647ll : (r < -2147483648ll ? -2147483648ll : r);
    return (int32_t)r;
  }
}

// Fast path for the common case (double rounding, right shift only).
__host__ __device__ __forceinline__ int32_t mbqm_rshift(int32_t x, int32_t qm, int right) {
  return rdivpot(srdhm(x, qm), right);
}

// Double-rounding requantisation specialised for a right shift n = -shift in [1, 31]:
//   SRDHM(acc, mult) = (acc*mult + 2^30) >> 31                          (one wide multiply-add + funnel shift)
//   RoundingDivideByPOT(v, n) = (v + 2^(n-1) + (v >> 31)) >> n          (ties away from zero)
// Bit-identical to mbqm(acc, mult, -n, 0); see tests/test_oracle_graph.py for the equivalence check.
__host__ __device__ __forceinline__ int32_t rq_fast(int32_t acc, int32_t mult, int n) {
  const long long p = (long long)acc * (long long)mult + (1ll << 30);
  const int32_t v = (int32_t)(p >> 31);
  return (v + (1 << (n - 1)) + (v >> 31)) >> n;
}

template <bool FAST>
__host__ __device__ __forceinline__ int32_t requant_t(int32_t acc, int32_t mult, int shift, int rounding) {
  if (FAST) return rq_fast(acc, mult, -shift);
  return mbqm(acc, mult, shift, rounding);
}

__host__ __device__ __forceinline__ int32_t clampi(int32_t v, int32_t lo, int32_t hi) {
  return v < lo ? lo : (v > hi ? hi : v);
}

#ifdef __CUDACC__
// "Saturating form" of the requantisation, for layers whose output zero point and clamp are -128 / [-128, 127]
// (every ReLU / ReLU6 layer of the graph): with C = bias' * mult + 2^30 + (2^(n-1) + zp * 2^n) * 2^31,
//   y = hi32(acc * mult + C) >> (n - 1)
// equals floor((v + 2^(n-1)) / 2^n) + zp with v = SRDHM(acc + bias', mult).  gemmlowp's tie nudge (v >> 31) only
// acts on v < 0, where both forms are <= zp = -128 and saturate to the same byte, so it is dropped.
__device__ __forceinline__ int rq_hi(int acc, int c_lo, int c_hi, int mult) {
  const long long c = ((long long)c_hi << 32) | (unsigned)c_lo;
  return (int)(((long long)acc * (long long)mult + c) >> 32);
}
// four int32 -> four int8 with signed saturation (I2IP.S8.S32.SAT), byte 0 = a
__device__ __forceinline__ unsigned pack4_sat(int a, int b, int c, int d) {
  unsigned r;
  asm("{\n\t.reg .b32 t;\n\t"
      "cvt.pack.sat.s8.s32.b32 t, %4, %3, 0;\n\t"
      "cvt.pack.sat.s8.s32.b32 %0, %2, %1, t;\n\t}"
      : "=r"(r) : "r"(a), "r"(b), "r"(c), "r"(d));
  return r;
}
#endif

// cudaFuncSetAttribute is per device: "once" flags of the launchers are bit masks over device ordinals, so an engine on
// cuda:1 created after one on cuda:0 in the same process still raises its kernels' dynamic shared-memory limit.
inline bool first_use_on_device(unsigned long long& mask) {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) d = 0;
  const unsigned long long bit = 1ull << (d & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

struct ConvParams {
  const int8_t* w;
  const int32_t* bias;
  const int32_t* mult;
  const int32_t* shift;
  int kh, kw, sh, sw, pt, pl;
  int in_zp, out_zp, act_min, act_max;
  int ih, iw, ic, oh, ow, oc;
  int rounding;
};

struct AddParams {
  int in1_zp, in2_zp, out_zp, left_shift;
  int m1, s1, m2, s2, mo, so;
  int act_min, act_max, bcast, C;
  int rounding;
};

}  // namespace bn
