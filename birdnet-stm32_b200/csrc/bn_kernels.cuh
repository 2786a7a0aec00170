// bn_kernels.cuh -- host-side launchers of the engine's CUDA kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bn_common.cuh"

namespace bn {

int set_error(int code, const char* msg);   // bn_engine.cu: sets bn_last_error() for this thread, returns code

// ---- generic plan (bn_generic.cu) -------------------------------------------------------
void launch_quantize(const float* x, int8_t* y, long n, float scale, int zp, cudaStream_t st);
void launch_dequantize(const int8_t* x, float* y, long n, float scale, int zp, cudaStream_t st);
void launch_requant(const int8_t* x, int8_t* y, long n, int in_zp, int out_zp, int mult, int shift, int R, cudaStream_t st);
void launch_transpose(const int8_t* x, int8_t* y, long n, const int* id, const int* od, const int* perm, cudaStream_t st);
void launch_slice(const int8_t* x, int8_t* y, long n, const int* id, const int* od, const int* begin, cudaStream_t st);
void launch_fill(int8_t* y, long n, int val, cudaStream_t st);
void launch_concat(const int8_t* x0, const int8_t* x1, int8_t* y, long rows, int in0, int in1, cudaStream_t st);
void launch_conv2d(const int8_t* x, int8_t* y, long n, const ConvParams& P, cudaStream_t st);
void launch_dwconv2d(const int8_t* x, int8_t* y, long n, const ConvParams& P, cudaStream_t st);
void launch_fc(const int8_t* x, int8_t* y, long n, const ConvParams& P, cudaStream_t st);
void launch_add(const int8_t* a, const int8_t* b, int8_t* y, long n, long per_chunk, const AddParams& P, cudaStream_t st);
void launch_mul(const int8_t* a, const int8_t* b, int8_t* y, long n, long per_chunk, const int* p, int C, int R, cudaStream_t st);
void launch_mean(const int8_t* x, int8_t* y, long n, int N, int C, const int* p, float in_scale, float out_scale, int variant, int R, cudaStream_t st);
void launch_logistic(const int8_t* x, int8_t* y, long n, const int8_t* lut, cudaStream_t st);
void launch_pad(const int8_t* x, int8_t* y, long n, const int* id, const int* od, const int* p, cudaStream_t st);
void launch_softmax(const int8_t* x, int8_t* y, long rows, int L, const float* table, float out_scale, int out_zp, cudaStream_t st);
void launch_sum(const int8_t* x, int8_t* y, long n, int outer, int count, int inner, float scale, float bias, int out_zp, cudaStream_t st);
void launch_minmax_normalize(float* s, long per_chunk, long n, const unsigned* mnmx, cudaStream_t st);
void launch_pool(const float* scores, const int* offs, float* out, int F, int C, int method, float beta, cudaStream_t st);

// ---- frontend (bn_frontend.cu) -----------------------------------------------------------
// |STFT| of B chunks -> out float32 [B, 257, W] (un-normalised) and per-chunk {min,max} bit patterns.
// Launches 2 kernels.  Returns 0 or a bn_status.
// pcm: int16 [B, T] (f32 = 0) or float32 waveform chunks [B, T] (f32 = 1).
int launch_stft_mag(const void* pcm, int f32, const float* peak, float* out, unsigned* mnmx, int B, int T, int n_fft,
                    int hop, int W, cudaStream_t st);

int launch_stft_mag_fm(const void* pcm, int f32, const float* peak, float* out, unsigned* mnmx, int B, int T, int n_fft,
                       int hop, int W, int ldk, cudaStream_t st);

}  // namespace bn
