// bn_se.cuh -- fused squeeze-and-excitation gate (see bn_se.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bn {

struct SeParams {
  int npix, C, C1;                  // pixels per map, channels, bottleneck width
  // MEAN (BN_MEAN_* slots of the blob op)
  int mean_in_zp, mean_out_zp, mean_mult, mean_shift, mean_mult_n, mean_shift_n;
  float in_scale, out_scale;
  // FULLY_CONNECTED 1 ([C1][C] weights) and 2 ([C][C1]); device pointers into the blob
  const int8_t* w1; const int32_t* b1; const int32_t* m1; const int32_t* s1;
  int fc1_in_zp, fc1_out_zp, fc1_act_min, fc1_act_max;
  const int8_t* w2; const int32_t* b2; const int32_t* m2; const int32_t* s2;
  int fc2_in_zp, fc2_out_zp, fc2_act_min, fc2_act_max;
  const int8_t* lut;                // LOGISTIC: 256 bytes indexed by code + 128
};

bool se_supported(const SeParams& P);
// x int8 [Bw][npix][C] -> the four op outputs [Bw][C], [Bw][C1], [Bw][C], [Bw][C]; variant = MEAN variant, R = rounding option
int launch_se_gate(const int8_t* x, int8_t* y_mean, int8_t* y_fc1, int8_t* y_fc2, int8_t* y_gate, int Bw, const SeParams& P,
                   int variant, int R, cudaStream_t st);

}  // namespace bn
