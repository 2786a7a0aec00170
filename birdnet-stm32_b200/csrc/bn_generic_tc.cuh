// bn_generic_tc.cuh -- per-op acceleration records of the generic plan (see bn_generic_tc.cu).
#pragma once
#include <vector>

#include "../../include/bn_blob.h"
#include "bn_layer.cuh"
#include "bn_pw_tc.cuh"
#include "bn_se.cuh"

namespace bn {

struct GenAccelOp {
  bool pw = false;      // 1x1 convolution routed through the tcgen05 GEMM (bn_pw_tc.cu)
  PwTcParams tc{};
  bool pwc = false;     // 1x1 convolution the tensor-core kernel has no build for (K or N of 512, ...): tiled dp4a GEMM (k_pw, bn_fast.cu)
  PwParams pwp{};
  // multi-op fusions (BN_OPT_FUSION bit 0; the ops they cover are skipped and, for the convolution + ADD pair, the convolution's
  // own output tensor is not materialised)
  bool add_fused = false;   // pw: the ADD that follows (op + 1) runs in the GEMM epilogue; tc_add = tc with the ADD constants
  PwTcParams tc_add{};
  int add_res_slot = -1, add_out_slot = -1;
  bool dsb = false;         // dw: DEPTHWISE 3x3 -> CONV 1x1 [-> ADD with the block input] (ops + 0 .. + ds_skip) as ONE fused DS-block
  DsParams ds{};            //     kernel launch (k_ds, bn_ds.cu): the depthwise result never leaves the SM
  DsLaunch dsl{};
  int ds_out_slot = -1, ds_skip = 0;
  bool se = false;          // MEAN -> FC -> FC -> LOGISTIC (ops + 0 .. + 3) as one launch (bn_se.cu)
  SeParams sep{};
  bool stem = false;    // 3x3 stride-(1,2) 1 -> 16 stem convolution: the fused plan's stem kernel (k_stem_sat, bn_fast.cu)
  StemParams stp{};
  bool dw = false;      // depthwise 3x3, stride 1 | 2: register-window kernel of the fused plan's layer path (k_dw3x3, bn_fast.cu)
  DwParams dwp{};
};

struct GenAccel {
  std::vector<GenAccelOp> ops;   // indexed like the blob's op table
  std::vector<void*> owned;
  int n_pw = 0, n_pwc = 0, n_dw = 0, n_stem = 0, n_add_fused = 0, n_se = 0, n_dsb = 0;
};

GenAccel* gen_accel_build(const uint8_t* h_blob, const uint8_t* d_blob, const bn_blob_header* hdr, const bn_blob_tensor* tensors,
                          const bn_blob_op* ops);
void gen_accel_destroy(GenAccel* a);

}  // namespace bn
