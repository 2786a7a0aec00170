// bn_pw_tc.cuh -- tcgen05 int8 GEMM for the pointwise convolutions (see bn_pw_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace bn {

struct PwTcParams {
  const uint8_t* w_img;   // pre-swizzled smem image of the weights, N * KP bytes
  const int* bias;        // folded bias'
  const int* mult;
  const int* shift;       // all <= -1 (fast requant domain proven at plan build)
  int K, KP, RW, cpr_log; // real K, padded K, smem row width (swizzle span), log2(K/16)
  int N;
  int tmem_cols;          // power of two >= 2N (double-buffered accumulators)
  int out_zp, act_min, act_max;
  int has_add;
  int add_in1_zp, add_in2_zp, add_out_zp;
  int add_m1, add_n1, add_m2, add_n2, add_mo, add_no;   // multipliers and right shifts (>= 0)
  int add_act_min, add_act_max;
};

bool pw_tc_supported(int K, int N);
void pw_tc_weight_image(const int8_t* w, int K, int N, std::vector<uint8_t>& img, int* KP_out, int* RW_out);
size_t pw_tc_smem_bytes(const PwTcParams& P);
int launch_pw_tc(const int8_t* x, const int8_t* res, int8_t* y, long M, const PwTcParams& P, int num_sms, cudaStream_t st);

}  // namespace bn
