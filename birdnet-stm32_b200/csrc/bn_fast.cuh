// bn_fast.cuh -- the fused kernel plan (pattern-matched on the lowered op list).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "bn_common.cuh"

namespace bn {

struct FastPlan {
  bool ok = false;             // the op list matched the fused pattern
  const bn_blob_header* hdr = nullptr;
  const bn_blob_tensor* tensors = nullptr;
  const bn_blob_op* ops = nullptr;
  uint8_t* d_blob = nullptr;
  int wave = 0;
  std::vector<void*> bufs;     // device workspace buffers
  std::vector<int> tap_ids;    // TFLite tensor ids materialised in bufs (same order)
  std::vector<size_t> tap_bytes;
  void* impl = nullptr;        // plan-specific state
};

void fast_plan_build(FastPlan& fp, const bn_blob_header* hdr, const bn_blob_tensor* tensors, const bn_blob_op* ops,
                     uint8_t* d_blob);
int fast_plan_alloc_workspace(FastPlan& fp, int wave, size_t* total_bytes);
void fast_plan_free_workspace(FastPlan& fp);
int fast_run_pcm(FastPlan& fp, const int16_t* d_pcm, const float* d_peak, int Bw, float* d_scores, int rounding,
                 int mean_variant, cudaStream_t st, int64_t* launches);
int fast_run_spec(FastPlan& fp, const float* d_spec, int Bw, float* d_scores, int rounding, int mean_variant,
                  cudaStream_t st, int64_t* launches);
int fast_dump_tensor(FastPlan& fp, int tfl_tensor_id, int Bw, void* out, size_t nbytes);

}  // namespace bn
