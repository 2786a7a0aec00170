// bn_generic_tc.cuh -- per-op acceleration records of the generic plan (see bn_generic_tc.cu).
#pragma once
#include <vector>

#include "../../include/bn_blob.h"
#include "bn_pw_tc.cuh"

namespace bn {

struct GenAccelOp {
  bool pw = false;      // 1x1 convolution routed through the tcgen05 GEMM (bn_pw_tc.cu)
  PwTcParams tc{};
};

struct GenAccel {
  std::vector<GenAccelOp> ops;   // indexed like the blob's op table
  std::vector<void*> owned;
  int n_pw = 0;
};

GenAccel* gen_accel_build(const uint8_t* h_blob, const bn_blob_header* hdr, const bn_blob_tensor* tensors, const bn_blob_op* ops);
void gen_accel_destroy(GenAccel* a);

}  // namespace bn
