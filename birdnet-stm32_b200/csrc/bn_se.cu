// bn_se.cu -- squeeze-and-excitation gate as ONE kernel: MEAN(H, W) -> FULLY_CONNECTED (ReLU) -> FULLY_CONNECTED -> LOGISTIC.
//
// Reference counterpart: se_block (birdnet_stm32/models/blocks.py:27-46: GlobalAveragePooling2D -> Dense(C / r, relu) ->
// Dense(C, sigmoid) -> Multiply) as the converter lowers it into the .tflite.  The generic plan runs the four ops as four
// launches on [C]-sized tensors; here one CTA per chunk reads the feature map once, keeps the channel sums and both dense
// layers in shared memory and writes all four (tiny) op outputs, so every tensor of the graph still exists in memory and the
// integer arithmetic per op is the generic kernels' (k_mean, k_fc, k_logistic in bn_generic.cu), only the launches are fused.
// The broadcast MUL that applies the gate stays a separate (4-wide) launch.
#include "bn_se.cuh"

#include "bn_common.cuh"

namespace bn {

constexpr int SE_THREADS = 256;

__global__ void __launch_bounds__(SE_THREADS)
k_se_gate(const int8_t* __restrict__ x, int8_t* __restrict__ y_mean, int8_t* __restrict__ y_fc1, int8_t* __restrict__ y_fc2,
          int8_t* __restrict__ y_gate, SeParams P, int variant, int R) {
  extern __shared__ __align__(16) int se_smem[];
  int* s_sum = se_smem;                                   // [C]
  int8_t* s_mean = reinterpret_cast<int8_t*>(s_sum + P.C);   // [C]
  int8_t* s_h = s_mean + P.C;                             // [C1]
  const int tid = threadIdx.x;
  const long b = blockIdx.x;
  const int C = P.C, CG = C >> 2;
  for (int i = tid; i < C; i += SE_THREADS) s_sum[i] = 0;
  __syncthreads();
  // ---- channel sums over the N pixels: thread = (pixel stripe, group of 4 channels), one 32-bit word per pixel ----
  {
    const int cg = tid % CG, stripe = tid / CG, nstripes = SE_THREADS / CG;
    if (stripe < nstripes) {
      const unsigned* xp = reinterpret_cast<const unsigned*>(x + b * (long)P.npix * C) + cg;
      int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      for (int j = stripe; j < P.npix; j += nstripes) {
        const int w = (int)__ldg(xp + (long)j * CG);
        a0 = __dp4a(w, 0x00000001, a0); a1 = __dp4a(w, 0x00000100, a1); a2 = __dp4a(w, 0x00010000, a2); a3 = __dp4a(w, 0x01000000, a3);
      }
      atomicAdd(&s_sum[4 * cg + 0], a0); atomicAdd(&s_sum[4 * cg + 1], a1); atomicAdd(&s_sum[4 * cg + 2], a2); atomicAdd(&s_sum[4 * cg + 3], a3);
    }
  }
  __syncthreads();
  // ---- MEAN requantisation (variants of SURVEY Appendix B.6, as in k_mean) ----
  for (int c = tid; c < C; c += SE_THREADS) {
    const int sum = s_sum[c], N = P.npix;
    int o;
    if (variant == 1) {
      const float scale = __fdiv_rn(P.in_scale, P.out_scale);
      const float bias = __fmul_rn(-(float)P.mean_in_zp, scale);
      const float fm = __fdiv_rn((float)sum, (float)N);
      o = (int)roundf(__fadd_rn(__fmul_rn(fm, scale), bias)) + P.mean_out_zp;
    } else if (variant == 2) {
      o = mbqm(sum - P.mean_in_zp * N, P.mean_mult_n, P.mean_shift_n, R) + P.mean_out_zp;
    } else {
      int acc = mbqm(sum - P.mean_in_zp * N, P.mean_mult, P.mean_shift, R);
      acc = acc > 0 ? (acc + N / 2) / N : (acc - N / 2) / N;
      o = acc + P.mean_out_zp;
    }
    const int8_t q = (int8_t)clampi(o, -128, 127);
    s_mean[c] = q;
    y_mean[b * C + c] = q;
  }
  __syncthreads();
  // ---- dense 1: [C] -> [C1], one warp per output (lanes stride over the K / 4 weight words, dp4a; sum(w) for the zero point) ----
  {
    const int warp = tid >> 5, lane = tid & 31;
    const int* xw = reinterpret_cast<const int*>(s_mean);
    for (int o = warp; o < P.C1; o += SE_THREADS / 32) {
      const int* ww = reinterpret_cast<const int*>(P.w1 + (long)o * C);
      int a = 0, ws = 0;
      for (int k = lane; k < CG; k += 32) {
        const int wv = __ldg(ww + k);
        a = __dp4a(xw[k], wv, a);
        ws = __dp4a(0x01010101, wv, ws);
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, d); ws += __shfl_xor_sync(0xffffffffu, ws, d); }
      if (lane == 0) {
        int acc = a - P.fc1_in_zp * ws + __ldg(P.b1 + o);
        acc = mbqm(acc, __ldg(P.m1 + o), __ldg(P.s1 + o), R) + P.fc1_out_zp;
        const int8_t q = (int8_t)clampi(acc, P.fc1_act_min, P.fc1_act_max);
        s_h[o] = q;
        y_fc1[b * P.C1 + o] = q;
      }
    }
  }
  __syncthreads();
  // ---- dense 2: [C1] -> [C], then the LOGISTIC table; one thread per output ----
  for (int o = tid; o < C; o += SE_THREADS) {
    const int8_t* wp = P.w2 + (long)o * P.C1;
    int acc = 0;
    if ((P.C1 & 3) == 0) {
      const int* ww = reinterpret_cast<const int*>(wp);
      const int* hw = reinterpret_cast<const int*>(s_h);
      int a = 0, ws = 0;
      for (int k = 0; k < (P.C1 >> 2); k++) {
        const int wv = __ldg(ww + k);
        a = __dp4a(hw[k], wv, a);
        ws = __dp4a(0x01010101, wv, ws);
      }
      acc = a - P.fc2_in_zp * ws;
    } else {
      for (int k = 0; k < P.C1; k++) acc += ((int)s_h[k] - P.fc2_in_zp) * (int)__ldg(wp + k);
    }
    acc += __ldg(P.b2 + o);
    acc = mbqm(acc, __ldg(P.m2 + o), __ldg(P.s2 + o), R) + P.fc2_out_zp;
    const int8_t q = (int8_t)clampi(acc, P.fc2_act_min, P.fc2_act_max);
    y_fc2[b * C + o] = q;
    y_gate[b * C + o] = __ldg(P.lut + (uint8_t)(q + 128));
  }
}

bool se_supported(const SeParams& P) {
  return P.C % 4 == 0 && P.C >= 4 && P.C / 4 <= SE_THREADS && P.C1 >= 1 && P.C1 <= 1024 && P.npix >= 1 &&
         (size_t)P.C * 4 + P.C + P.C1 + 16 <= 48 * 1024;
}

int launch_se_gate(const int8_t* x, int8_t* y_mean, int8_t* y_fc1, int8_t* y_fc2, int8_t* y_gate, int Bw, const SeParams& P,
                   int variant, int R, cudaStream_t st) {
  if (Bw < 1) return 0;
  if (!se_supported(P)) return BN_ERR_UNSUPPORTED;
  const size_t smem = (size_t)P.C * 4 + P.C + P.C1 + 16;
  k_se_gate<<<Bw, SE_THREADS, smem, st>>>(x, y_mean, y_fc1, y_fc2, y_gate, P, variant, R);
  return 0;
}

}  // namespace bn
