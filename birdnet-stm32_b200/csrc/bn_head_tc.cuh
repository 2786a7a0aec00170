// bn_head_tc.cuh -- K2tc: normalise + QUANTIZE + mel-mixer 1x1 conv on tcgen05 + folded PWL LUT (see bn_head_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace bn {

constexpr int HT_N = 64;          // mel channels
constexpr int HT_KP = 288;        // K padded to 9 MMA k-steps of 32
constexpr int HT_B_BYTES = HT_N * HT_KP;

struct HeadTcParams {
  const uint8_t* w_img;   // HT_B_BYTES: K-major swizzled smem image of the mixer weights (2 x SW128 blocks + 1 x SW32 block)
  const int4* rq;         // [64] {c_lo, c_hi, mult, n - 1}: saturating-form requantisation (rq_hi, bn_common.cuh)
  const uint8_t* lut;     // [64][256] folded element-wise chain, indexed by code + 128
  int K_real;             // 257 real bins; columns K_real .. ldk-1 of the conv input hold `fill`
  int ldk;                // floats per frame row of the magnitude buffer (= conv K, 264)
  int fill;
  int W;                  // frames per chunk (multiple of 128)
  float q_scale;
  int q_zp;
};

void head_tc_weight_image(const int8_t* w, int K, std::vector<uint8_t>& img);
int launch_head_tc(const float* mags, const unsigned* mnmx, int8_t* out, int Bw, const HeadTcParams& P, int num_sms, cudaStream_t st);

}  // namespace bn
