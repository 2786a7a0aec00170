// bn_stem_tc.cu -- K3tc: the stem convolution (3x3, stride (1,2), 1 -> 16 channels) as an im2col GEMM on tcgen05.
//
// Reference counterpart: the first Conv2D of build_dscnn_model (birdnet_stm32/models/dscnn.py:198-262) as lowered into the
// .tflite (CONV_2D 3x3, stride (1,2), SAME, ReLU6).  Same integer results as k_stem_sat (bn_fast.cu): the accumulator is
// sum(x * w) over the nine taps (zero-point padding outside the map), the input zero point is folded into the bias and the
// requantisation is the saturating form of bn_common.cuh (rq_hi).
//
// One output row of the map is 128 pixels = one 128 x 16 x 32 tcgen05.mma kind::i8: the A operand row of pixel ox holds its nine
// input bytes (x[iy-1..iy+1][2 ox .. 2 ox + 2]) in the first 16-byte chunk of a 32-byte K-major row (SWIZZLE_32B), the other
// 23 bytes meet zero weights.  Per tile of 8 output rows a CTA stages the 10 input rows (2.5 KB), every thread assembles the A
// rows of two pixel pairs with byte permutes (6 loads, 9 PRMT, 2 16-byte stores per pair), one thread issues the 8 MMAs, and
// the 8 warps drain TMEM: one tcgen05.ld gives a thread the 16 channels of its pixel, the requantisation constants are kernel
// parameters (constant-bank operands, no loads), and the pixel leaves as ONE 16-byte store.  The nine multiply-adds per
// output, 27 % of the old kernel's instructions, are gone from the CUDA cores.
#include "bn_stem_tc.cuh"

#include "bn_common.cuh"
#include "bn_tc.cuh"

namespace bn {

namespace {

constexpr int ST_ROWS = 8;                 // output rows per tile = MMAs per tile
constexpr int ST_THREADS = 256;
constexpr int ST_PITCH = 256 + 16;         // input row pitch in shared memory: 256 bytes + the right halo (zero point)
constexpr int ST_A_BYTES = ST_ROWS * 128 * 32;

}  // namespace

__global__ void __launch_bounds__(ST_THREADS, 4)
k_stem_tc(const int8_t* __restrict__ in, int8_t* __restrict__ out, int ntiles, StemTcParams P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = smem;                                  // [ST_ROWS][128][32], SWIZZLE_32B
  unsigned char* sB = sA + ST_A_BYTES;                       // [16][32]
  unsigned char* sIn = sB + 1024;                            // [ST_ROWS + 2][ST_PITCH]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(sIn + (ST_ROWS + 2) * ST_PITCH + 16);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), ST_ROWS * 16);
  if (tid == 32) {
    mbar_init(smem_u32(mbar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 512 / 16; i += ST_THREADS) *reinterpret_cast<uint4*>(sB + 16 * i) = __ldg(reinterpret_cast<const uint4*>(P.w_img) + i);
  for (int i = tid; i < ST_A_BYTES / 16; i += ST_THREADS) *reinterpret_cast<uint4*>(sA + 16 * i) = make_uint4(0, 0, 0, 0);
  const unsigned zpw = 0x01010101u * (unsigned)(uint8_t)P.in_zp;
  for (int i = tid; i < (ST_ROWS + 2) * 4; i += ST_THREADS)  // right halo: never overwritten by the staging
    *reinterpret_cast<unsigned*>(sIn + (i >> 2) * ST_PITCH + 256 + 4 * (i & 3)) = zpw;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = make_idesc_i8(128, 16);
  const int bands = P.oh / ST_ROWS;
  const int q = warp & 3, half = warp >> 2;

  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
    const int b = tile / bands, y0 = (tile - b * bands) * ST_ROWS;
    // ---- input rows y0 - 1 .. y0 + ST_ROWS (SAME padding rows = zero point) ----
    if (tid < (ST_ROWS + 2) * 16) {
      const int r = tid >> 4, p = tid & 15;
      const int iy = y0 - 1 + r;
      uint4 v = make_uint4(zpw, zpw, zpw, zpw);
      if (iy >= 0 && iy < P.ih) v = __ldg(reinterpret_cast<const uint4*>(in + ((size_t)b * P.ih + iy) * 256) + p);
      *reinterpret_cast<uint4*>(sIn + r * ST_PITCH + 16 * p) = v;
    }
    __syncthreads();
    // ---- im2col: thread = (output row ry, pixel pair pp) -> pixels 2 pp, 2 pp + 1 ----
#pragma unroll
    for (int k = 0; k < ST_ROWS * 64 / ST_THREADS; k++) {
      const int task = tid + k * ST_THREADS;
      const int ry = task >> 6, pp = task & 63;
      unsigned w0[3], w1[3];
#pragma unroll
      for (int fy = 0; fy < 3; fy++) {
        const unsigned* rp = reinterpret_cast<const unsigned*>(sIn + (ry + fy) * ST_PITCH) + pp;
        w0[fy] = rp[0]; w1[fy] = rp[1];
      }
      // even pixel: bytes 0..2 of w0; odd pixel: bytes 2, 3 of w0 and byte 0 of w1
      unsigned e[3], o[3];
#pragma unroll
      for (int fy = 0; fy < 3; fy++) { e[fy] = w0[fy]; o[fy] = __byte_perm(w0[fy], w1[fy], 0x4432); }
      const uint4 ae = make_uint4(__byte_perm(e[0], e[1], 0x4210), __byte_perm(e[1], e[2], 0x5421), __byte_perm(e[2], 0u, 0x4442), 0u);
      const uint4 ao = make_uint4(__byte_perm(o[0], o[1], 0x4210), __byte_perm(o[1], o[2], 0x5421), __byte_perm(o[2], 0u, 0x4442), 0u);
      const int m = 2 * pp;                                  // rows m and m + 1 share (m >> 2) & 1
      unsigned char* ap = sA + ry * 4096 + m * 32 + (((m >> 2) & 1) << 4);
      *reinterpret_cast<uint4*>(ap) = ae;
      *reinterpret_cast<uint4*>(ap + 32) = ao;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t bd = make_desc(smem_u32(sB), 256, 6u);
#pragma unroll
      for (int ry = 0; ry < ST_ROWS; ry++)
        umma_i8(tmem_base + (uint32_t)(ry * 16), make_desc(smem_u32(sA) + ry * 4096, 256, 6u), bd, idesc, 0u);
      umma_commit(smem_u32(mbar));
    }
    mbar_wait(smem_u32(mbar), (uint32_t)(it & 1));
    tc_fence_after();
    // ---- epilogue: thread = pixel 32 q + lane of rows 4 half .. 4 half + 3; 16 channels -> one 16-byte store ----
#pragma unroll
    for (int rr = 0; rr < ST_ROWS / 2; rr++) {
      const int ry = half * (ST_ROWS / 2) + rr;
      int v[16];
      tmem_ld16(tmem_base + (uint32_t)(ry * 16) + ((uint32_t)(32 * q) << 16), v);
      unsigned ow[4];
#pragma unroll
      for (int g = 0; g < 4; g++) {
        int y[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int4 rq = P.rq[4 * g + j];
          y[j] = rq_hi(v[4 * g + j], rq.x, rq.y, rq.z) >> rq.w;
        }
        ow[g] = pack4_sat(y[0], y[1], y[2], y[3]);
      }
      *reinterpret_cast<uint4*>(out + (((size_t)b * P.oh + y0 + ry) * 128 + 32 * q + lane) * 16) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
    tc_fence_before();
    __syncthreads();                                         // TMEM drained, A operand and input rows free
  }
  if (warp == 0) tmem_dealloc(tmem_base, ST_ROWS * 16);
}

bool stem_tc_supported(int ih, int iw, int oh, int ow) { return iw == 256 && ow == 128 && oh == ih && oh % ST_ROWS == 0; }

int launch_stem_tc(const int8_t* in, int8_t* out, int Bw, const StemTcParams& P, int num_sms, cudaStream_t st) {
  if (Bw < 1) return 0;
  if (!stem_tc_supported(P.ih, 256, P.oh, 128)) return BN_ERR_UNSUPPORTED;
  const size_t smem = ST_A_BYTES + 1024 + (ST_ROWS + 2) * ST_PITCH + 64 + 1024;
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) cudaFuncSetAttribute(k_stem_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const long ntiles = (long)Bw * (P.oh / ST_ROWS);
  long grid = (long)num_sms * 4;
  if (grid > ntiles) grid = ntiles;
  k_stem_tc<<<(int)grid, ST_THREADS, smem, st>>>(in, out, (int)ntiles, P);
  return 0;
}

}  // namespace bn
