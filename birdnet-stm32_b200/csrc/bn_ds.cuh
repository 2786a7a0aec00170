// bn_ds.cuh -- fused depthwise-separable block kernel (see bn_ds.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "bn_stem_tc.cuh"

namespace bn {

// One DS block of models/dscnn.py:28-84 (ds_conv_block): DEPTHWISE_CONV_2D 3x3 (+ReLU6) -> CONV_2D 1x1
// (+ReLU6 | linear) [-> ADD(block input, conv) + ReLU6].  All pointers are device pointers.
struct DsParams {
  // depthwise 3x3
  const int4* dw_wm;      // [9][C/4] masked weight words (byte j of word j = w[tap][4*cg + j], other bytes 0)
  const int4* dw_wt;      // [3][2][C/4] filter-row words for the transposed depthwise (DsLaunch::dwt): word j of entry (ky, 0) =
                          //   (w[ky][0], w[ky][1], w[ky][2], 0) of channel 4 cg + j, entry (ky, 1) = the same taps one byte up
  const int4* dw_rq;      // [C] {c_lo, c_hi, mult, n - 1}, saturating form: c = bias' * mult + 2^30 + (2^(n-1) + zp * 2^n) * 2^31
  const int* dw_rz;       // unused by the saturating form
  // pointwise 1x1
  const uint8_t* w_img;   // N * KP bytes: K-major swizzled shared-memory image of the weights
  const int4* pw_rq;      // [N] saturating form when there is no ADD; else {c_lo, c_hi, mult, n} with c = bias' * mult + 2^30
  const int* pw_rz;       // [N] 2^(n-1) + out_zp * 2^n (ADD blocks: the conv output is signed, tie nudge kept)
  const int4* pw_rq2;     // [N] multiply-high epilogue (DsLaunch::epi): {2c lo, 2c hi, (int)(2 mult) wrapped, 2^(32 - n)}
  const int2* pw_rz2;     // [N] {2^(n-1) + zp2 * 2^n, 128}; constant channels: mult = 0, 2^(32-n) = 0, {0, code + 128}
  long long a_co2;        // a_co with the +128 domain of the clamped conv code folded in
  int two;                // the constant 2, kept opaque to the compiler (IMAD.HI instead of a shift for v >> 31)
  int epi_smem;           // extra shared memory of the multiply-high epilogue variants (0 for the default epilogue)
  int C, N, KP, RW;       // depthwise channels (= GEMM K), output channels, padded K, swizzle row width
  int ih, iw, oh, ow, pt, pl;
  int NB, MT;             // chunks per CTA tile, 128-row MMA tiles per CTA tile
  int nst;                // input-tile buffers in shared memory: 1 = unpipelined, 2 / 3 = software pipeline (bn_ds.cu)
  // tensor-core depthwise variant (bn_ds_tc.cu, stride 1): diag(w[tap]) 32x32 blocks, no-swizzle K-major, [C/32][9][1024 B]
  const uint8_t* dw_img;
  int MTd;                // 128-row M tiles of the depthwise GEMM over padded pixel indices
  int plane_px;           // pixels per channel-chunk plane in shared memory (odd, covers the last tap of the last M tile)
  int TRr;                // output rows per tile (runtime copy of the TR template parameter of k_ds)
  int ow_log, trow_log, cg_log, ppr_log;   // log2 of ow, TR*ow, C/4, iw*C/16
  int sw_sh, sw_mask, rw_log;              // swizzle: chunk ^= (row >> sw_sh) & sw_mask
  int dw_in_zp, dw_lo, dw_hi;
  int pw_lo, pw_hi;
  int tmem_cols;
  // residual ADD, input 1 = block input, input 2 = conv output
  int a_m1, a_n1, a_rz1;  long long a_c1;   // s1 = ((r + 128) * m1 + c1) >> a_n1, a_n1 = 11 + n1, c1 = 2^10 + 2^(n1+10) (zp1 = -128)
  int a_m2, a_n2, a_rz2;  long long a_c2;   // generic conv term
  int a_mo, a_no, a_rzo;  long long a_co;   // y = hi32(t * mo + co) >> a_no, a_no = no - 1, rounding + zp folded into co (saturating form)
  int a_zpo;                                // used when no == 0
  int a_lo, a_hi;
  // Copies of pw_rq / pw_rz for blocks with N <= 64, passed BY VALUE: the k_ds builds with a compile-time channel count (NC)
  // index them with compile-time channel numbers, so the epilogue constants are constant-bank / uniform-register operands
  // instead of two or three shared-memory loads per output.  nc = N when filled, else 0.  (Measured for the 128-channel stage
  // kernel too: 2.04 -> 2.41 ms with a run-time block index into the parameter space, 2.56 ms with the block loop unrolled --
  // four blocks x 128 channels of constants do not fit the uniform registers -- so bn_stage.cu keeps its shared-memory copies.)
  int nc;
  int4 pw_rqc[64];
  int pw_rzc[64];
  // k_ds<..., STEM = 1> (first block only): the kernel's input is the HEAD output int8 [B][ih][256]; the stem convolution of the
  // tile's input rows is computed in the kernel (im2col GEMM of bn_stem_tc.cu) straight into the shared-memory input tile, so the
  // stem output tensor (131 KB per chunk written and read back) never exists in global memory.
  StemTcParams stem;
  int stem_off;           // byte offset of the stem scratch (head rows, im2col operand, weight image) in dynamic shared memory
};

struct DsLaunch {
  int S, TR, add_mode;    // stride, output rows per tile, 0 none / 1 generic / 2 conv term = (o - zp2) << 19
  size_t smem;
  int ctas_per_sm;
  int tcdw;               // 1 = run bn_ds_tc.cu (both convolutions on the tensor core)
  int threads;            // 256, or 512 for single-CTA layers (see k_ds)
  int dwt;                // 1 = depthwise with filter rows along the dp4a axis (PRMT transpose, 3 dp4a per output), 0 = masked words (9)
  int epi;                // residual-ADD epilogue variant of k_ds (add_mode 2): 0 ALU-pipe shifts, 1 multiply-high form, 2 = 1 + residual table
};

int launch_ds(const int8_t* in, int8_t* out, int Bw, const DsParams& P, const DsLaunch& L, int num_sms, cudaStream_t st);
size_t ds_smem_bytes(const DsParams& P, int S, int TR);
// stem + first DS block in one kernel (P.stem / P.stem_off filled, P.nst = 1, TR = 4): in = head output
bool ds_stem_supported(const DsParams& P, int S, int add_mode);
size_t ds_stem_smem_bytes(const DsParams& P, int* stem_off);
int launch_ds_stem(const int8_t* head_out, int8_t* out, int Bw, const DsParams& P, size_t smem, int num_sms, cudaStream_t st);

// warp-specialised pipeline form of the same block (bn_ds_ws.cu): P.nst = 3 | 4 input-tile buffers, two A operands, two accumulators
int launch_dsw(const int8_t* in, int8_t* out, int Bw, const DsParams& P, const DsLaunch& L, int num_sms, cudaStream_t st);
size_t dsw_smem_bytes(const DsParams& P, int S, int TR);
bool dsw_supported(const DsParams& P, int S, int TR, int add_mode);

int launch_dst(const int8_t* in, int8_t* out, int Bw, const DsParams& P, const DsLaunch& L, int num_sms, cudaStream_t st);
size_t dst_smem_bytes(const DsParams& P);
void dst_weight_image(const int8_t* w, int C, std::vector<uint8_t>& img);

}  // namespace bn
