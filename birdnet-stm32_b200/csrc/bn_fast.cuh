// bn_fast.cuh -- the fused kernel plan (pattern-matched on the lowered op list).
//
// Plan shape (hybrid-frontend DS-CNN graphs such as the shipped checkpoint):
//   K1  stft_mag_fm   PCM16 -> |STFT| float32, frame-major, + per-chunk min/max       (bn_frontend.cu)
//   K2  head          normalise + QUANTIZE + mel-mixer 1x1 conv + PWL chain (folded into a per-channel
//                     256-entry LUT) + transpose  ->  int8 [mel, frames]
//   K3  stem          3x3 stride-(1,2) conv, Cin = 1
//   K4  dw            depthwise 3x3 (stride 1 or 2)
//   K5  pw            pointwise 1x1 conv as an int8 GEMM with requantisation, residual ADD and
//                     ReLU6 fused into the epilogue
//   K6  tail          MEAN + FULLY_CONNECTED + LOGISTIC + DEQUANTIZE
// Anything that does not match falls back to the generic one-kernel-per-op plan (still CUDA).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "bn_common.cuh"

namespace bn {

// Per-kernel timing with CUDA events on the launching stream (BN_OPT_PROFILE).
struct Profiler {
  bool on = false;
  struct Pending { int slot; cudaEvent_t a, b; };
  std::vector<std::string> names;
  std::vector<double> ms;
  std::vector<int64_t> count;
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> pool;
  int cur = -1;
  cudaEvent_t cur_a = nullptr;
  void begin(const char* name, cudaStream_t st);
  void end(cudaStream_t st);
  void collect();   // synchronises the pending events and accumulates
  void reset();
  ~Profiler();
};

struct FastImpl;

struct FastPlan {
  bool ok = false;             // the op list matched the fused pattern
  const bn_blob_header* hdr = nullptr;
  const bn_blob_tensor* tensors = nullptr;
  const bn_blob_op* ops = nullptr;
  const uint8_t* h_blob = nullptr;
  uint8_t* d_blob = nullptr;
  int wave = 0;
  int use_tc = 1;              // pointwise convs on tcgen05 (BN_OPT_TENSOR_CORE)
  int num_sms = 148;
  int fusion = 139;            // BN_OPT_FUSION -- bit 0: fused DS-block kernels, bit 1: fused frontend, bit 2: depthwise on the tensor core
                               // too, bit 3: whole-stage kernels (bn_stage.cu), bit 4: stage kernels also write their inner block outputs (taps),
                               // bit 5: quantising frontend K1q / K2q (bn_frontend_q.cu) instead of K1 + float32 scratch + K2 (off by default:
                               // measured 0.9 % slower end to end, 27 % less DRAM traffic -- DESIGN.md section 5), bit 6: warp-specialised DS-block
                               // kernel (bn_ds_ws.cu; measured equal, off), bit 7: stem as an im2col GEMM on tcgen05 (bn_stem_tc.cu),
                               // bit 8: stem computed inside the first DS block's kernel (k_ds<..., STEM>; measured slower, off)
  FastImpl* impl = nullptr;
  std::string why;             // why the pattern did not match (diagnostics)
};

void fast_plan_build(FastPlan& fp, const uint8_t* h_blob, const bn_blob_header* hdr, const bn_blob_tensor* tensors,
                     const bn_blob_op* ops, uint8_t* d_blob);
void fast_plan_destroy(FastPlan& fp);
int fast_plan_alloc_workspace(FastPlan& fp, int wave, size_t* total_bytes);
void fast_plan_free_workspace(FastPlan& fp);
// d_pcm: int16 [Bw, T] (f32 = 0) or float32 waveform chunks (f32 = 1)
int fast_run_pcm(FastPlan& fp, const void* d_pcm, int f32, const float* d_peak, int Bw, float* d_scores, int rounding,
                 int mean_variant, cudaStream_t st, int64_t* launches, Profiler* prof);
int fast_run_spec(FastPlan& fp, const float* d_spec, int Bw, float* d_scores, int rounding, int mean_variant,
                  cudaStream_t st, int64_t* launches, Profiler* prof);
int fast_dump_tensor(FastPlan& fp, int tfl_tensor_id, int Bw, void* out, size_t nbytes);

}  // namespace bn
