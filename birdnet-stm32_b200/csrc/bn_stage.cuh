// bn_stage.cuh -- one stage of the DS-CNN (stride-2 block + the residual blocks after it) per kernel (see bn_stage.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bn_ds.cuh"

namespace bn {

constexpr int STAGE_MAX_BLOCKS = 4;

struct StageParams {
  DsParams L[STAGE_MAX_BLOCKS];   // the per-block constants of bn_ds.cu (block 0: stride 2, no ADD; blocks 1..: stride 1, add_mode 2)
  int8_t* dbg[STAGE_MAX_BLOCKS];  // optional: also write block l's output to global memory (debug taps; nullptr = keep it on the SM)
  int nl;
};

bool stage_supported(int C0, int C, int OH, int OW, int nl);
size_t stage_smem_bytes(int C0, int C, int OH, int OW);
int launch_stage(const int8_t* in, int8_t* out, int Bw, const StageParams& SP, int C0, int C, int OH, int OW, int num_sms, cudaStream_t st);

}  // namespace bn
