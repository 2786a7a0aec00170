// bn_stem_tc.cuh -- stem convolution as an im2col GEMM on tcgen05 (see bn_stem_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bn {

struct StemTcParams {
  const uint8_t* w_img;   // 512 bytes: SWIZZLE_32B K-major image of the weights [16][32] (k = 3 fy + fx, k >= 9 zero)
  int4 rq[16];            // per channel {c_lo, c_hi, mult, n - 1}: saturating-form requantisation (rq_hi, bn_common.cuh)
  int ih, oh, in_zp;      // map height (input = output), input zero point; width is 256 -> 128
};

bool stem_tc_supported(int ih, int iw, int oh, int ow);
// in int8 [Bw][ih][256] -> out int8 [Bw][oh][128][16]
int launch_stem_tc(const int8_t* in, int8_t* out, int Bw, const StemTcParams& P, int num_sms, cudaStream_t st);

}  // namespace bn
