// bn_frontend_q.cuh -- K1q / K2q: STFT + chunk-wide min-max + QUANTIZE in one kernel, mel-mixer GEMM on the int8 image (bn_frontend_q.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bn_head_tc.cuh"

namespace bn {

constexpr int HQ_A_BYTES = 128 * HT_KP;     // one 128-frame A operand: 2 x (128 x 128 B, SW128) + 128 x 32 B (SW32) = 36,864 bytes

struct FrontendQParams {
  float q_scale;    // QUANTIZE of the graph input (scale 1/255)
  int q_zp;
  int fill;         // FILL value of the CONCAT padding columns k = 257 .. 263
};

bool frontend_q_supported(int n_fft, int W, int ldk, int K_real, int hop, int f32);
// aimg: uint8 [B][W / 128][HQ_A_BYTES]
// mnmx: uint32 [2 B], arrive: uint32 [B] (scratch, initialised by the launcher).  Launches 2 kernels.
int launch_stft_q(const void* pcm, int f32, const float* peak, uint8_t* aimg, unsigned* mnmx, unsigned* arrive, int B, int T, int n_fft,
                  int hop, int W, const FrontendQParams& Q, int num_sms, cudaStream_t st);
int launch_head_q(const uint8_t* aimg, int8_t* out, int Bw, const HeadTcParams& P, int num_sms, cudaStream_t st);

}  // namespace bn
