// bn_layer.cuh -- parameter blocks and launchers of the per-layer CUDA-core kernels of bn_fast.cu (K4 depthwise 3x3, K5 pointwise
// GEMM), shared with the generic plan's accelerated ops (bn_generic_tc.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/bn_blob.h"
#include "bn_ds.cuh"

namespace bn {

struct PwParams {                 // pointwise conv (+ residual add) -- device pointers
  const int* wt;                  // [K/4][N] words: 4 consecutive k of channel n
  const int* bias;                // folded bias'
  const int* mult;
  const int* shift;
  int K, N;
  int out_zp, act_min, act_max;   // of the conv itself
  int fast;                       // all channels requantise with a right shift >= 1
  int has_add;
  const int* lut_res;             // [256] rescaled residual   (add input 1)
  const int* lut_conv;            // [256] rescaled conv output (add input 2)
  int add_mo, add_so, add_out_zp, add_act_min, add_act_max;
};

struct DwParams {
  const int* wm;                  // [9][C/4][4] masked weight words
  const int* bias; const int* mult; const int* shift;   // [C]
  int C, ih, iw, oh, ow, sh, sw, pt, pl;
  int in_zp, out_zp, act_min, act_max;
  int fast;
};

struct StemParams {
  const int* w;                   // [16][3] words (w0,w1,w2,0) per (co, fy)
  const int* bias; const int* mult; const int* shift;
  int ih, iw, oh, ow, in_zp, out_zp, act_min, act_max;
  int fast;
  const int4* pk;                 // [16][2] {w_fy0, w_fy1, w_fy2, n - 1}, {c_lo, c_hi, mult, 0}: saturating form (rq_hi)
  int sat;                        // pk is valid (zp_out = -128, clamp [-128, 127], int32-safe)
};

// Fused DS-block kernel parameters (k_ds, bn_ds.cu) from the blob ops of a DEPTHWISE_CONV_2D 3x3 -> CONV_2D 1x1 [-> ADD with the
// block input] group; false when the group is outside the kernel's proven domain.  Launch with launch_ds().
bool ds_block_build(const uint8_t* h_blob, const bn_blob_tensor* T, const bn_blob_op* ops, int dw_op, int pw_op, int add_op,
                    std::vector<void*>& owned, DsParams& D, DsLaunch& L);
// Stem parameter block from a CONV_2D blob op (false when the op is not a 3x3 stride-(1,2) 1 -> 16 convolution); device
// allocations are appended to `owned`.  in int8 [Bw][ih][iw] -> out int8 [Bw][oh][ow][16].
bool stem_build(const uint8_t* h_blob, const bn_blob_tensor* T, const bn_blob_op& st, std::vector<void*>& owned, StemParams& S);
int launch_stem(const int8_t* in, int8_t* out, int Bw, const StemParams& S, int R, cudaStream_t st);
// in int8 [Bw][ih][iw][C] -> out int8 [Bw][oh][ow][C]; R = BN_OPT_ROUNDING
int launch_dw3x3(const int8_t* in, int8_t* out, int Bw, const DwParams& D, int R, cudaStream_t st);
// x int8 [M][K] -> y int8 [M][N] (+ residual ADD when P.has_add); whole weight matrix staged per CTA (K * N <= ~128 K)
int launch_pw(const int8_t* x, const int8_t* res, int8_t* y, long M, const PwParams& P, int R, cudaStream_t st);
size_t pw_cuda_core_smem(int K, int N);

}  // namespace bn
