// bn_fft.cuh -- device building blocks of the STFT kernels (bn_frontend.cu, bn_frontend_q.cu): constants, packed FP32 pair
// arithmetic, the register-resident 16-point FFT.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bn {

constexpr int NFFT = 512;
constexpr int NC = 256;          // complex points
constexpr int BINS = 257;
constexpr int FRAMES_PER_CTA = 32;
constexpr int FE_THREADS = 256;  // 8 warps = 16 half-warps = 16 concurrent FFTs, 2 rounds
constexpr int TILE_LD = FRAMES_PER_CTA + 1;

// sqrt.approx.ftz.f32: max relative error 2^-23, far inside the 1e-4 frontend tolerance.  .ftz makes it ONE MUFU.SQRT: without it
// every call carries a subnormal-range fix-up (FSETP + two predicated FMULs); a squared magnitude below 1.2e-38 (an input of
// 1e-19 of full scale -- digital silence is exactly 0 either way) now gives 0 instead of a value below 1.1e-19.
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void cp_async16_fe(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

// Packed FP32 pairs (FADD2 on sm_100a): one issue slot adds or subtracts both halves of a complex number with the same
// IEEE rounding as two scalar FADDs.  K1 is bound by instruction issue, not by the FP32 pipe, so halving the count of the
// butterfly additions is a direct gain.  The mov.b64 packs / unpacks are register naming only (no SASS is emitted).
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  unsigned long long x, y, z;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(z) : "l"(x), "l"(y));
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(z));
  return r;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  unsigned long long x, y, z;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b.x), "f"(b.y));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(z) : "l"(x), "l"(y));
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(z));
  return r;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// In-register 16-point DIF FFT, output in natural order (bit reversal folded into the
// compile-time unrolled index map).  tw16[j] = exp(-2 pi i j / 16).
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
  const float2 tw[8] = {{1.f, 0.f}, {c1, -s1}, {r2, -r2}, {s1, -c1}, {0.f, -1.f}, {-s1, -c1}, {-r2, -r2}, {-c1, -s1}};
#pragma unroll
  for (int half = 8; half >= 1; half >>= 1) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if ((i & half) == 0) {
        float2 a = v[i], b = v[i + half];
        v[i] = add2(a, b);
        float2 d = sub2(a, b);
        const int j = (i & (half - 1)) * (8 / half);   // twiddle index into tw (N=16 base)
        if (j == 0) v[i + half] = d;
        else if (j == 4) v[i + half] = make_float2(d.y, -d.x);
        else v[i + half] = cmul(d, tw[j]);
      }
    }
  }
  // bit-reverse permutation (4 bits)
  float2 o[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int r = ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3);
    o[r] = v[i];
  }
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = o[i];
}

// Real-FFT split partner without shared memory.  After the second radix-16 pass lane l of a half-warp holds
// Z[l + 16 k2] in v[k2].  The split pairs bin k = l + 16 j (j < 8) with bin 256 - k = (16 - l) + 16 (15 - j): lane (16 - l) & 15,
// register 15 - j -- one shuffle per component.  Lane 0 pairs with itself (256 - 16 j = 16 (16 - j), and Z[256] = Z[0]).
// All 32 lanes of the warp must call this (the two half-warps shuffle independently inside one instruction).
template <int J>
__device__ __forceinline__ float2 split_partner(const float2 (&v)[16], int l, int lane_in_warp) {
  const int src = (lane_in_warp & 16) | ((16 - l) & 15);
  float2 zn;
  zn.x = __shfl_sync(0xffffffffu, v[15 - J].x, src);
  zn.y = __shfl_sync(0xffffffffu, v[15 - J].y, src);
  if (l == 0) zn = v[(16 - J) & 15];
  return zn;
}

}  // namespace bn
