// bn_ds.cu -- K45: one depthwise-separable block per kernel, intermediates never leave the SM.
//
//   in  int8 [B][ih][iw][C]  --DW 3x3 (stride 1|2, SAME, ReLU6)-->  A tile in shared memory (K-major, swizzled)
//       --tcgen05.mma kind::i8 with the pointwise weights--> int32 accumulators in TMEM
//       --epilogue: requant (+ residual ADD with the block input still resident in shared memory) + clamp-->
//   out int8 [B][oh][ow][N]
//
// Reference counterpart: ds_conv_block (birdnet_stm32/models/dscnn.py:28-84) as lowered into the .tflite
// (DEPTHWISE_CONV_2D, CONV_2D 1x1, ADD) and executed by tf.lite.Interpreter.invoke
// (birdnet_stm32/models/runners.py:93-95).  Integer semantics: SURVEY Appendix B.3-B.5, bit-exact.
//
// A CTA is persistent over tiles of TR output rows x ow columns (x NB chunks for the late 4x8 layers) =
// MT * 128 output pixels.  Per tile: (1) cp.async the input rows (+halo, SAME padding = zero-point bytes) into
// shared memory, (2) every thread walks one (column, 4-channel group) strip down the rows with a 3x3 register
// window and dp4a against masked weight words, requantises and writes the int8 result straight into the
// swizzled A operand, (3) one thread issues the MMAs, (4) all 8 warps drain TMEM (tcgen05.ld 32x32b.x16),
// requantise, add the residual from the shared-memory input tile and store 16-byte pieces.
//
// Requantisation uses per-channel constants prepared on the host so that bias, rounding nudges and the output
// zero point cost no instructions:   v = (acc * mult + (bias' * mult + 2^30)) >> 31        (SRDHM, 64-bit IMAD)
//                                    y = (v + (2^(n-1) + zp * 2^n) + (v >> 31)) >> n       (RoundingDivideByPOT + zp)
// Equality with gemmlowp's sequence and the int32-safe domain are checked at plan-build time (bn_fast.cu).
#include "bn_ds.cuh"

#include <cstdlib>

#include "bn_common.cuh"
#include "bn_tc.cuh"

namespace bn {

constexpr int DS_THREADS = 256;

__device__ __forceinline__ int rq64(int acc, int c_lo, int c_hi, int mult, int rz, int n) {
  const long long c = ((long long)c_hi << 32) | (unsigned)c_lo;
  const long long p = (long long)acc * (long long)mult + c;
  const int v = (int)(p >> 31);
  return (v + rz + (v >> 31)) >> n;
}
__device__ __forceinline__ int clamp2(int v, int lo, int hi) { return max(lo, min(v, hi)); }
// hi32(a * b) + c as one IMAD.HI (FMA pipe); b is a run-time value so that the compiler cannot turn it back into a shift
__device__ __forceinline__ int mad_hi(int a, int b, int c) {
  int r;
  asm("mad.hi.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
// max(0, min(a, b)): VIMNMX.RELU
__device__ __forceinline__ int min_relu(int a, int b) {
  int r;
  asm("min.relu.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// TMA bulk copy global -> shared (1-D), completion counted in bytes on an mbarrier
__device__ __forceinline__ void ds_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void ds_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// MB = resident CTAs per SM the register allocation is sized for: 2 (128 registers) or 3 (80 registers; only the
// transposed depthwise fits that without meaningful spilling).
// NT = threads per CTA: 256, or 512 for the late layers whose shared-memory footprint (64 KB weight image, four-chunk tiles)
// leaves one CTA per SM -- 16 warps instead of 8 to hide the phase latencies.
// EPI (ADD == 2 only): residual-ADD epilogue variant.  0 = shifts and clamps on the ALU pipe (round 1); 1 = "multiply-high"
// form: every arithmetic right shift is an IMAD.HI by a power of two and SRDHM's >> 31 is folded into a wrapped doubled
// multiplier, so the integer work moves from the saturated ALU pipe (SHF / VIMNMX / LEA) to the FMA pipe (IMAD family),
// and the [-128, 127] clamp is one VIMNMX.RELU in the +128 domain; 2 = variant 1 with the residual rescale
// MBQM((r - zp1) << 20, m1, s1) read from a 256-entry shared-memory table instead of computed.
// NC > 0: the block's output channel count as a compile-time constant (32 | 64).  The epilogue then walks the 16-channel groups of
// a warp with compile-time channel numbers and reads its requantisation constants from the by-value copies in DsParams
// (constant bank -> uniform registers, hoisted out of the per-pixel work): no shared-memory constant loads at all.
// STEM = 1 (S = 2, TR = 4, NC = 32, unpipelined; the first block of the shipped graph): `in` is the head output and the input
// tile is PRODUCED in the kernel by the stem's im2col GEMM instead of being loaded (see DsParams::stem).
template <int S, int TR, int ADD, int DWT, int MB, int NT, int EPI = 0, int NC = 0, int STEM = 0>
__global__ void __launch_bounds__(NT, MB)
k_ds(const int8_t* __restrict__ in, int8_t* __restrict__ out, int Bw, int ntiles, DsParams P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int TRIN = (TR - 1) * S + 3;             // input rows per chunk in a tile
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = P.C, CG = C >> 2, N = P.N, KP = P.KP, RW = P.RW;
  const int TW = P.iw + P.pl + 1;                     // tile columns (left halo only when pad_left = 1)
  const int b_bytes = N * KP, a_bytes = P.MT * 128 * KP;
  const int tile_bytes = (P.NB * TRIN * TW * C + 15) & ~15;
  const int NST = P.nst;                              // input-tile buffers: 1 = unpipelined, 2 / 3 = prefetch distance 1 / 2
  const int NAB = NST > 1 ? 2 : 1;                    // A-operand and TMEM accumulator buffers
  unsigned char* sB = smem;
  unsigned char* sA0 = sB + b_bytes;
  unsigned char* sT0 = sA0 + NAB * a_bytes;
  int4* s_rq = reinterpret_cast<int4*>(sT0 + NST * tile_bytes);
  int* s_rz = reinterpret_cast<int*>(s_rq + N);                 // EPI >= 1: int2 {rz, addc} per channel
  uint64_t* mbar = reinterpret_cast<uint64_t*>(s_rz + (EPI >= 1 ? 2 * N : ((N + 1) & ~1)));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2);
  int* s_lut1 = reinterpret_cast<int*>(tmem_slot + 2);          // EPI == 2: residual term per code, [256]
  // The NC builds fetch the input rows with TMA bulk copies (one cp.async.bulk per row, issued by the lanes of warp 0,
  // completion in bytes on fbar[buffer]) instead of 16-byte cp.async pieces issued by every thread: the staging was 12 % of the
  // first block's instructions (ncu source view), now it is a few instructions of one warp.
  constexpr bool TMA = NC > 0;
  uint64_t* fbar = reinterpret_cast<uint64_t*>(tmem_slot + 2);  // [3] (aliases s_lut1, which only the EPI == 2 builds use)

  // ---- one-time setup ---------------------------------------------------------------------------
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), (uint32_t)P.tmem_cols);
  if (tid == 32) {
    mbar_init(smem_u32(&mbar[0]), 1);
    mbar_init(smem_u32(&mbar[1]), 1);
    if (TMA) for (int s = 0; s < 3; s++) mbar_init(smem_u32(&fbar[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < b_bytes / 16; i += NT) cp_async16(smem_u32(sB + 16 * i), P.w_img + 16 * (size_t)i);
  cp_async_commit();
  if (EPI == 0) {
    for (int i = tid; i < N; i += NT) { s_rq[i] = __ldg(P.pw_rq + i); s_rz[i] = __ldg(P.pw_rz + i); }
  } else {
    // [N] 2c (64-bit) | [N] {(int)(2 m), 2^(32 - n)} | [N] rz << 32
    for (int i = tid; i < N; i += NT) {
      const int4 q = __ldg(P.pw_rq2 + i);
      reinterpret_cast<int2*>(s_rq)[i] = make_int2(q.x, q.y);
      reinterpret_cast<int2*>(s_rq)[N + i] = make_int2(q.z, q.w);
      reinterpret_cast<int2*>(s_rz)[i] = make_int2(0, __ldg(P.pw_rz2 + i).x);
    }
    if (EPI == 2)
      for (int i = tid; i < 256; i += NT) s_lut1[i] = (int)(((unsigned long long)(unsigned)i * (unsigned)P.a_m1 + (unsigned long long)P.a_c1) >> P.a_n1);
  }
  if (P.C < KP) for (int i = tid; i < NAB * a_bytes / 16; i += NT) *reinterpret_cast<uint4*>(sA0 + 16 * i) = make_uint4(0, 0, 0, 0);
  const unsigned zpw = 0x01010101u * (unsigned)(uint8_t)P.dw_in_zp;
  {  // halo columns hold the zero point for the whole kernel (cp.async never touches them)
    const int rows = P.NB * TRIN;
    const int wpc = C >> 2;                           // words per pixel
    for (int sb = 0; sb < NST; sb++) {
      unsigned* tw = reinterpret_cast<unsigned*>(sT0 + sb * tile_bytes);
      for (int i = tid; i < rows * wpc * 2; i += NT) {
        const int side = i & 1, rest = i >> 1;
        const int row = rest / wpc, w = rest - row * wpc;
        if (side == 0 && P.pl == 0) continue;
        const int col = side ? TW - 1 : 0;
        tw[(row * TW + col) * wpc + w] = zpw;
      }
    }
  }
  // per-thread depthwise constants: the channel group of a thread is the same for all of its strips
  const int cg = tid & (CG - 1);
  int4 w[DWT ? 1 : 9];                                // DWT 0: masked words, one tap of 4 channels each
  int4 wl[DWT ? 3 : 1], wr[(DWT && S == 1) ? 3 : 1];   // DWT 1: per filter row, word j = (w[ky][0], w[ky][1], w[ky][2], 0) of channel 4 cg + j
  if (DWT == 0) {
#pragma unroll
    for (int t = 0; t < 9; t++) w[t] = __ldg(P.dw_wm + t * CG + cg);
  } else {
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
      wl[ky] = __ldg(P.dw_wt + (2 * ky) * CG + cg);
      if (S == 1) wr[ky] = __ldg(P.dw_wt + (2 * ky + 1) * CG + cg);   // same taps one byte up: the right output column of a pair
    }
  }
  int4 drq[4];
#pragma unroll
  for (int j = 0; j < 4; j++) drq[j] = __ldg(P.dw_rq + 4 * cg + j);
  // A-operand position of this thread's 4 channels: k-half, 16-byte chunk, byte in chunk
  const int k0 = 4 * cg;
  const int a_kh_off = (k0 >> P.rw_log) * (128 * RW);
  const int a_cc = (k0 & (RW - 1)) >> 4;
  const int a_b = k0 & 15;

  cp_async_wait_all();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t sbo = 8 * RW;
  const uint32_t lt = RW == 128 ? 2u : (RW == 64 ? 4u : 6u);
  const uint32_t idesc = make_idesc_i8(128, N);
  const int ksteps = KP >> 5, ksteps_per_half = RW >> 5;
  const int q = warp & 3, hsel = warp >> 2;
  const int NG = N >> 4;                              // 16-column groups
  const int tiles_per_chunk = P.oh / TR;
  const int ppr = (P.iw * C) >> 4;                    // 16-byte pieces per input row
  const int nstrips = P.NB * P.ow * CG;

  auto tile_origin = [&](int tile, int& b0, int& oy0) {
    if (P.NB == 1) { b0 = tile / tiles_per_chunk; oy0 = (tile - b0 * tiles_per_chunk) * TR; }
    else { b0 = tile * P.NB; oy0 = 0; }
  };

  // ---- (1) stage the input rows of a tile (cp.async; SAME padding rows = zero-point bytes) ---------------------
  auto stage = [&](int tile, unsigned char* sT) {
    int b0, oy0;
    tile_origin(tile, b0, oy0);
    for (int row = warp; row < P.NB * TRIN; row += NT / 32) {
      const int bb = row / TRIN, tr = row - bb * TRIN;
      const int iy = oy0 * S - P.pt + tr;
      const bool ok = (b0 + bb) < Bw && iy >= 0 && iy < P.ih;
      unsigned char* dst = sT + ((size_t)row * TW + P.pl) * C;
      const int8_t* src = in + (((size_t)(b0 + bb) * P.ih + iy) * P.iw) * C;
      if (ok) { for (int p = lane; p < ppr; p += 32) cp_async16(smem_u32(dst + 16 * p), src + 16 * p); }
      else { for (int p = lane; p < ppr; p += 32) *reinterpret_cast<uint4*>(dst + 16 * p) = make_uint4(zpw, zpw, zpw, zpw); }
    }
  };

  // ---- (1') the same with TMA: lane = input row of the tile (warp 0 only); padding rows are filled by the warp -------------
  auto stage_tma = [&](int tile, unsigned char* sT, uint64_t* bar) {
    int b0, oy0;
    tile_origin(tile, b0, oy0);
    const bool mine = lane < P.NB * TRIN;
    const int bb = lane / TRIN, tr = lane - bb * TRIN;
    const int iy = oy0 * S - P.pt + tr;
    const bool ok = mine && (b0 + bb) < Bw && iy >= 0 && iy < P.ih;
    const unsigned okmask = __ballot_sync(0xffffffffu, ok);
    unsigned padmask = __ballot_sync(0xffffffffu, mine && !ok);
    const uint32_t row_bytes = (uint32_t)(P.iw * C);
    if (lane == 0) ds_mbar_expect_tx(smem_u32(bar), (uint32_t)__popc(okmask) * row_bytes);
    __syncwarp();
    if (ok) {
      fence_proxy_async();
      ds_bulk_g2s(smem_u32(sT + ((size_t)lane * TW + P.pl) * C), in + (((size_t)(b0 + bb) * P.ih + iy) * P.iw) * C, row_bytes, smem_u32(bar));
    }
    while (padmask) {
      const int row = __ffs(padmask) - 1;
      padmask &= padmask - 1;
      unsigned char* dst = sT + ((size_t)row * TW + P.pl) * C;
      for (int p = lane; p < ppr; p += 32) *reinterpret_cast<uint4*>(dst + 16 * p) = make_uint4(zpw, zpw, zpw, zpw);
    }
  };

  // ---- (2) depthwise 3x3 -> swizzled A operand --------------------------------------------------------------------
  auto depthwise = [&](const unsigned char* sT, unsigned char* sA) {
    for (int sidx = tid; sidx < nstrips; sidx += NT) {
      const int rest = sidx >> P.cg_log;
      const int ox = rest & (P.ow - 1), bb = rest >> P.ow_log;
      const unsigned* tp = reinterpret_cast<const unsigned*>(sT) + ((size_t)(bb * TRIN) * TW + ox * S) * CG + cg;
      const int mbase = ((bb * TR) << P.ow_log) + ox;
      unsigned x0[3], x1[3], x2[3];
#pragma unroll
      for (int fx = 0; fx < 3; fx++) { x0[fx] = tp[fx * CG]; if (S == 1) x1[fx] = tp[(TW + fx) * CG]; }
#pragma unroll
      for (int r = 0; r < TR; r++) {
        const unsigned* rp = tp + (size_t)(r * S) * TW * CG;
#pragma unroll
        for (int fx = 0; fx < 3; fx++) {
          if (S == 2) x1[fx] = rp[(TW + fx) * CG];
          x2[fx] = rp[(2 * TW + fx) * CG];
        }
        int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
        for (int fx = 0; fx < 3; fx++) {
          const int4 wa = w[fx], wb = w[3 + fx], wc = w[6 + fx];
          a0 = __dp4a((int)x0[fx], wa.x, a0); a1 = __dp4a((int)x0[fx], wa.y, a1); a2 = __dp4a((int)x0[fx], wa.z, a2); a3 = __dp4a((int)x0[fx], wa.w, a3);
          a0 = __dp4a((int)x1[fx], wb.x, a0); a1 = __dp4a((int)x1[fx], wb.y, a1); a2 = __dp4a((int)x1[fx], wb.z, a2); a3 = __dp4a((int)x1[fx], wb.w, a3);
          a0 = __dp4a((int)x2[fx], wc.x, a0); a1 = __dp4a((int)x2[fx], wc.y, a1); a2 = __dp4a((int)x2[fx], wc.z, a2); a3 = __dp4a((int)x2[fx], wc.w, a3);
        }
        const int q0 = rq_hi(a0, drq[0].x, drq[0].y, drq[0].z) >> drq[0].w;
        const int q1 = rq_hi(a1, drq[1].x, drq[1].y, drq[1].z) >> drq[1].w;
        const int q2 = rq_hi(a2, drq[2].x, drq[2].y, drq[2].z) >> drq[2].w;
        const int q3 = rq_hi(a3, drq[3].x, drq[3].y, drq[3].z) >> drq[3].w;
        const int m = mbase + (r << P.ow_log);
        const int j = m >> 7, row = m & 127;
        const int off = j * (128 * KP) + a_kh_off + row * RW + ((a_cc ^ ((row >> P.sw_sh) & P.sw_mask)) << 4) + a_b;
        *reinterpret_cast<unsigned*>(sA + off) = pack4_sat(q0, q1, q2, q3);
#pragma unroll
        for (int fx = 0; fx < 3; fx++) {
          if (S == 1) { x0[fx] = x1[fx]; x1[fx] = x2[fx]; }
          else x0[fx] = x2[fx];
        }
      }
    }
  };

  // ---- (2') depthwise 3x3 with the taps of a filter row along the dp4a axis ------------------------------------
  // The pixel words of an input row (4 channels each) are byte-transposed with PRMT into one word per channel that
  // holds 4 (stride 1) or 3 (stride 2) horizontally adjacent pixels; one dp4a against (w0, w1, w2, 0) then covers a
  // whole filter row of one channel: 3 dp4a per output instead of 9 masked ones.  With stride 1 a thread produces two
  // adjacent output columns from the same transposed words (second weight word = (0, w0, w1, w2)).
  // A-operand byte offset of output pixel m for this thread's 4 channels: [m / 128][k-half][m % 128][RW] with the
  // 16-byte chunk index XOR-swizzled by the row.  All swizzle periods are 8 rows and ow is a multiple of 8, so the
  // swizzle term of a strip does not change from row to row, and the right pixel of an even / odd pair only flips
  // chunk bit 0 when the swizzle uses row bit 0 (RW = 128).
  const int a_jhop = 128 * (KP - RW);
  const unsigned a_pairx = P.sw_sh == 0 ? 16u : 0u;
  auto depthwise_t = [&](const unsigned char* sT, unsigned char* sA) {
    constexpr int NCOL = S == 1 ? 2 : 1;              // output columns per thread
    const int owp_log = P.ow_log - (S == 1 ? 1 : 0);
    const int nst = nstrips >> (S == 1 ? 1 : 0);
    for (int sidx = tid; sidx < nst; sidx += NT) {
      const int rest = sidx >> P.cg_log;
      const int oxp = rest & ((1 << owp_log) - 1), bb = rest >> owp_log;
      const int ox = oxp * NCOL;
      const unsigned* tp = reinterpret_cast<const unsigned*>(sT) + ((size_t)(bb * TRIN) * TW + ox * S) * CG + cg;
      const int mbase = ((bb * TR) << P.ow_log) + ox;
      const int a_thr = a_kh_off + (((a_cc ^ ((mbase >> P.sw_sh) & P.sw_mask)) << 4) | a_b);
      auto load_row = [&](int ir, unsigned (&t)[4]) {   // transposed words of input row ir of the tile
        const unsigned* rp = tp + (size_t)ir * TW * CG;
        const unsigned x0 = rp[0], x1 = rp[CG], x2 = rp[2 * CG];
        const unsigned a = __byte_perm(x0, x1, 0x5140), b = __byte_perm(x0, x1, 0x7362);
        if (S == 1) {
          const unsigned x3 = rp[3 * CG];
          const unsigned c = __byte_perm(x2, x3, 0x5140), d = __byte_perm(x2, x3, 0x7362);
          t[0] = __byte_perm(a, c, 0x5410); t[1] = __byte_perm(a, c, 0x7632);
          t[2] = __byte_perm(b, d, 0x5410); t[3] = __byte_perm(b, d, 0x7632);
        } else {                                        // byte 3 meets a zero weight byte: any value will do
          t[0] = __byte_perm(a, x2, 0x4410); t[1] = __byte_perm(a, x2, 0x5532);
          t[2] = __byte_perm(b, x2, 0x6610); t[3] = __byte_perm(b, x2, 0x7732);
        }
      };
      unsigned t0[4], t1[4], t2[4];
      load_row(0, t0);
      if (S == 1) load_row(1, t1);
#pragma unroll
      for (int r = 0; r < TR; r++) {
        if (S == 2) load_row(2 * r + 1, t1);
        load_row(r * S + 2, t2);
        int aL[4], aR[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const int w0 = c == 0 ? wl[0].x : c == 1 ? wl[0].y : c == 2 ? wl[0].z : wl[0].w;
          const int w1 = c == 0 ? wl[1].x : c == 1 ? wl[1].y : c == 2 ? wl[1].z : wl[1].w;
          const int w2 = c == 0 ? wl[2].x : c == 1 ? wl[2].y : c == 2 ? wl[2].z : wl[2].w;
          aL[c] = __dp4a((int)t2[c], w2, __dp4a((int)t1[c], w1, __dp4a((int)t0[c], w0, 0)));
          if (S == 1) {
            const int v0 = c == 0 ? wr[0].x : c == 1 ? wr[0].y : c == 2 ? wr[0].z : wr[0].w;
            const int v1 = c == 0 ? wr[1].x : c == 1 ? wr[1].y : c == 2 ? wr[1].z : wr[1].w;
            const int v2 = c == 0 ? wr[2].x : c == 1 ? wr[2].y : c == 2 ? wr[2].z : wr[2].w;
            aR[c] = __dp4a((int)t2[c], v2, __dp4a((int)t1[c], v1, __dp4a((int)t0[c], v0, 0)));
          }
        }
        const int m = mbase + (r << P.ow_log);
        const unsigned off = (unsigned)((m << P.rw_log) + (m >> 7) * a_jhop + a_thr);
        *reinterpret_cast<unsigned*>(sA + off) = (pack4_sat(rq_hi(aL[0], drq[0].x, drq[0].y, drq[0].z) >> drq[0].w, rq_hi(aL[1], drq[1].x, drq[1].y, drq[1].z) >> drq[1].w,
                                 rq_hi(aL[2], drq[2].x, drq[2].y, drq[2].z) >> drq[2].w, rq_hi(aL[3], drq[3].x, drq[3].y, drq[3].z) >> drq[3].w));
        if (S == 1)
          *reinterpret_cast<unsigned*>(sA + ((off + RW) ^ a_pairx)) = (pack4_sat(rq_hi(aR[0], drq[0].x, drq[0].y, drq[0].z) >> drq[0].w, rq_hi(aR[1], drq[1].x, drq[1].y, drq[1].z) >> drq[1].w,
                                       rq_hi(aR[2], drq[2].x, drq[2].y, drq[2].z) >> drq[2].w, rq_hi(aR[3], drq[3].x, drq[3].y, drq[3].z) >> drq[3].w));
#pragma unroll
        for (int c = 0; c < 4; c++) {
          if (S == 1) { t0[c] = t1[c]; t1[c] = t2[c]; }
          else t0[c] = t2[c];
        }
      }
    }
  };
  auto run_dw = [&](const unsigned char* sT, unsigned char* sA) {
    if (DWT) depthwise_t(sT, sA); else depthwise(sT, sA);
  };

  // ---- (3) pointwise conv on the tensor core (one thread) -------------------------------------------------------
  auto issue_mma = [&](const unsigned char* sA, uint32_t tmem_d, uint64_t* bar) {
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
    for (int j = 0; j < P.MT; j++) {
      for (int ks = 0; ks < ksteps; ks++) {
        const int h = ks / ksteps_per_half, kk = ks - h * ksteps_per_half;
        const uint64_t ad = make_desc(a_addr + j * (128 * KP) + h * (128 * RW) + kk * 32, sbo, lt);
        const uint64_t bd = make_desc(b_addr + h * (N * RW) + kk * 32, sbo, lt);
        umma_i8(tmem_d + (uint32_t)(j * N), ad, bd, idesc, ks > 0 ? 1u : 0u);
      }
    }
    umma_commit(smem_u32(bar));
  };

  // ---- (4) epilogue: TMEM -> requant (+ residual ADD from the shared-memory input tile) -> global ---------------
  auto epilogue = [&](int tile, const unsigned char* sT, uint32_t tmem_d) {
    int b0, oy0;
    tile_origin(tile, b0, oy0);
    const size_t pix0 = ((size_t)b0 * P.oh + oy0) << P.ow_log;
    auto unit = [&](const int j, const int g) {
      const int m = j * 128 + 32 * q + lane;
      const int bb = m >> P.trow_log;
      const bool ok = (b0 + bb) < Bw;
      int v[16];
      tmem_ld16(tmem_d + (uint32_t)(j * N + 16 * g) + ((uint32_t)(32 * q) << 16), v);
      uint4 rv = make_uint4(0, 0, 0, 0);
      if (ADD) {
        const int rem = m & ((1 << P.trow_log) - 1);
        const int r = rem >> P.ow_log, ox = rem & (P.ow - 1);
        rv = *reinterpret_cast<const uint4*>(sT + ((size_t)(bb * TRIN + r + P.pt) * TW + ox + P.pl) * C + 16 * g);
      }
      // residual bytes as unsigned r - zp1 (zp1 = -128)
      const unsigned rw[4] = {rv.x ^ 0x80808080u, rv.y ^ 0x80808080u, rv.z ^ 0x80808080u, rv.w ^ 0x80808080u};
      unsigned ow4[4];
#pragma unroll
      for (int gg = 0; gg < 4; gg++) {
        int o[4];
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
          const int c = 16 * g + 4 * gg + jj;
          const int4 rq = (ADD == 2 && EPI >= 1) ? make_int4(0, 0, 0, 0) : (NC > 0 ? P.pw_rqc[c] : s_rq[c]);
          if (!ADD) {
            o[jj] = rq_hi(v[4 * gg + jj], rq.x, rq.y, rq.z) >> rq.w;
          } else if (ADD == 2 && EPI >= 1) {
            // Every step is hi32(a * b + c64) -- ONE IMAD.HI with a 64-bit addend -- plus at most one add:
            //   vv = SRDHM(acc + bias', m)   = hi32(acc * (2 m - 2^32) + 2 c) + acc          (2 m wrapped to int32, c = bias' m + 2^30)
            //   t  = vv + rz + (vv >> 31)    = hi32(vv * 2 + rz 2^32) + vv                   (rounding term, zero point, tie nudge)
            //   y  = clamp(t >> n) + 128     = min.relu(hi32(t * 2^(32-n) + 128 2^32), 255)
            const int acc = v[4 * gg + jj];
            const long long c2 = reinterpret_cast<const long long*>(s_rq)[c];
            const int2 mp = reinterpret_cast<const int2*>(s_rq)[N + c];
            const long long rzp = reinterpret_cast<const long long*>(s_rz)[c];
            const int vv = (int)(((long long)acc * (long long)mp.x + c2) >> 32) + acc;
            const int t = (int)(((long long)vv * (long long)P.two + rzp) >> 32) + vv;
            const int y = min_relu((int)(((long long)t * (long long)mp.y + (128ll << 32)) >> 32), 255);
            const unsigned u = __byte_perm(rw[gg], 0u, 0x4440 + jj);
            const int s1 = EPI == 2 ? s_lut1[u] : (int)(((unsigned long long)u * (unsigned)P.a_m1 + (unsigned long long)P.a_c1) >> P.a_n1);
            const int t2 = s1 + (y << 19);                                                    // -(zp2 + 128) << 19 is folded into a_co2
            o[jj] = (int)(((long long)t2 * (long long)P.a_mo + P.a_co2) >> 32) >> P.a_no;
          } else {
            const int rz = NC > 0 ? P.pw_rzc[c] : s_rz[c];
            const int y = clamp2(rq64(v[4 * gg + jj], rq.x, rq.y, rq.z, rz, rq.w), P.pw_lo, P.pw_hi);
            // residual term RoundingDivideByPOT(SRDHM((r - zp1) << 20, m1), n1): the operand is >= 0, so both
            // roundings are plain "add half, shift" and fold into one 64-bit multiply-add + shift
            const unsigned u = __byte_perm(rw[gg], 0u, 0x4440 + jj);
            const int s1 = (int)(((unsigned long long)u * (unsigned)P.a_m1 + (unsigned long long)P.a_c1) >> P.a_n1);
            int t2;
            if (ADD == 2) {
              t2 = s1 + (y << 19);                    // conv term (y - zp2) << 19, -zp2 << 19 folded into a_co
            } else {
              int s2 = (int)(((long long)y * (long long)P.a_m2 + P.a_c2) >> 11);
              if (P.a_n2 > 0) s2 = (s2 + P.a_rz2 + (s2 >> 31)) >> P.a_n2;
              t2 = s1 + s2;
            }
            // output: saturating form (zp_out = -128, clamp [-128, 127]), a_no = right shift - 1
            o[jj] = (int)(((long long)t2 * (long long)P.a_mo + P.a_co) >> 32) >> P.a_no;
          }
        }
        ow4[gg] = pack4_sat(o[0], o[1], o[2], o[3]);
      }
      if (ok) *reinterpret_cast<uint4*>(out + (pix0 + m) * N + 16 * g) = make_uint4(ow4[0], ow4[1], ow4[2], ow4[3]);
    };
    if (NC > 0) {
      // compile-time channel groups: group gi belongs to the warps with hsel == gi % (NT / 128)
#pragma unroll
      for (int gi = 0; gi < NC / 16; gi++) {
        if ((gi % (NT / 128)) != hsel) continue;
        for (int j = 0; j < P.MT; j++) unit(j, gi);
      }
    } else {
      for (int t = hsel; t < P.MT * NG; t += NT / 128) {
        const int j = t / NG;
        unit(j, t - j * NG);
      }
    }
  };

  // ---- (0) STEM builds: the tile's TRIN stem rows from TRIN + 2 head rows ---------------------------------------------------------
  //   scratch at smem + P.stem_off: head rows [TRIN + 2][272] | im2col operand [TRIN][128][32] (SWIZZLE_32B) | weight image [16][32]
  //   tensor memory: pointwise accumulators at columns [0, MT * N), stem accumulators at [MT * N, MT * N + 16 TRIN)
  auto stem_rows = [&](int tile, unsigned char* sT, int it) {
    constexpr int HP = 272;                              // head row pitch: 256 bytes + right halo
    unsigned char* sH = smem + P.stem_off;
    unsigned char* sAs = sH + ((TRIN + 2) * HP + 1023) / 1024 * 1024;
    unsigned char* sBs = sAs + TRIN * 4096;
    const StemTcParams& Q = P.stem;
    int b0, oy0;
    tile_origin(tile, b0, oy0);
    const unsigned hzp = 0x01010101u * (unsigned)(uint8_t)Q.in_zp;
    if (it == 0) {                                       // once per CTA: weight image, zeroed operand, right halo of the head rows
      for (int i = tid; i < 512 / 16; i += NT) *reinterpret_cast<uint4*>(sBs + 16 * i) = __ldg(reinterpret_cast<const uint4*>(Q.w_img) + i);
      for (int i = tid; i < TRIN * 4096 / 16; i += NT) *reinterpret_cast<uint4*>(sAs + 16 * i) = make_uint4(0, 0, 0, 0);
      for (int i = tid; i < (TRIN + 2) * 4; i += NT) *reinterpret_cast<unsigned*>(sH + (i >> 2) * HP + 256 + 4 * (i & 3)) = hzp;
    }
    const int sr0 = oy0 * S - P.pt;                      // first stem row of the tile (S = 2, pt = 0: 2 oy0)
    if (tid < (TRIN + 2) * 16) {
      const int r = tid >> 4, p = tid & 15;
      const int iy = sr0 - 1 + r;
      uint4 v = make_uint4(hzp, hzp, hzp, hzp);
      if (iy >= 0 && iy < Q.ih && b0 < Bw) v = __ldg(reinterpret_cast<const uint4*>(in + ((size_t)b0 * Q.ih + iy) * 256) + p);
      *reinterpret_cast<uint4*>(sH + r * HP + 16 * p) = v;
    }
    __syncthreads();
    for (int task = tid; task < TRIN * 64; task += NT) {  // im2col: (stem row ry, pixel pair pp), see bn_stem_tc.cu
      const int ry = task >> 6, pp = task & 63;
      unsigned e[3], o[3];
#pragma unroll
      for (int fy = 0; fy < 3; fy++) {
        const unsigned* rp = reinterpret_cast<const unsigned*>(sH + (ry + fy) * HP) + pp;
        e[fy] = rp[0]; o[fy] = __byte_perm(rp[0], rp[1], 0x4432);
      }
      const uint4 ae = make_uint4(__byte_perm(e[0], e[1], 0x4210), __byte_perm(e[1], e[2], 0x5421), __byte_perm(e[2], 0u, 0x4442), 0u);
      const uint4 ao = make_uint4(__byte_perm(o[0], o[1], 0x4210), __byte_perm(o[1], o[2], 0x5421), __byte_perm(o[2], 0u, 0x4442), 0u);
      const int m = 2 * pp;
      unsigned char* ap = sAs + ry * 4096 + m * 32 + (((m >> 2) & 1) << 4);
      *reinterpret_cast<uint4*>(ap) = ae;
      *reinterpret_cast<uint4*>(ap + 32) = ao;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    const uint32_t tm_stem = tmem_base + (uint32_t)(P.MT * N);
    if (tid == 0) {
      tc_fence_after();
      const uint64_t bd = make_desc(smem_u32(sBs), 256, 6u);
      const uint32_t idesc_s = make_idesc_i8(128, 16);
#pragma unroll
      for (int ry = 0; ry < TRIN; ry++) umma_i8(tm_stem + (uint32_t)(ry * 16), make_desc(smem_u32(sAs) + ry * 4096, 256, 6u), bd, idesc_s, 0u);
      umma_commit(smem_u32(&mbar[1]));
    }
    mbar_wait(smem_u32(&mbar[1]), (uint32_t)(it & 1));
    tc_fence_after();
    // stem epilogue: thread = pixel 32 q + lane of rows hsel, hsel + 2, ...; 16 channels -> one 16-byte store into the input tile
    for (int ry = hsel; ry < TRIN; ry += NT / 128) {
      unsigned char* dst = sT + ((size_t)ry * TW + P.pl + 32 * q + lane) * C;
      if (sr0 + ry >= P.ih) {                             // SAME padding row below the map
        *reinterpret_cast<uint4*>(dst) = make_uint4(zpw, zpw, zpw, zpw);
        continue;
      }
      int v[16];
      tmem_ld16(tm_stem + (uint32_t)(ry * 16) + ((uint32_t)(32 * q) << 16), v);
      unsigned ow4[4];
#pragma unroll
      for (int g = 0; g < 4; g++) {
        int y[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int4 rq = Q.rq[4 * g + j];
          y[j] = rq_hi(v[4 * g + j], rq.x, rq.y, rq.z) >> rq.w;
        }
        ow4[g] = pack4_sat(y[0], y[1], y[2], y[3]);
      }
      *reinterpret_cast<uint4*>(dst) = make_uint4(ow4[0], ow4[1], ow4[2], ow4[3]);
    }
    tc_fence_before();
  };

  if (NST == 1) {
    // ---- unpipelined: stage -> depthwise -> MMA -> epilogue per tile --------------------------------------------
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
      if (STEM) {
        stem_rows(tile, sT0, it);
      } else if (TMA) {
        if (warp == 0) stage_tma(tile, sT0, &fbar[0]);
        mbar_wait(smem_u32(&fbar[0]), (uint32_t)(it & 1));
      } else {
        stage(tile, sT0);
        cp_async_commit();
        cp_async_wait_all();
      }
      __syncthreads();
      run_dw(sT0, sA0);
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) issue_mma(sA0, tmem_base, &mbar[0]);
      mbar_wait(smem_u32(&mbar[0]), (uint32_t)(it & 1));
      tc_fence_after();
      epilogue(tile, sT0, tmem_base);
      tc_fence_before();
      __syncthreads();                                 // TMEM drained, input tile and A operand free
    }
  } else {
    // ---- software pipeline over this CTA's tiles t_k = blockIdx.x + k * gridDim.x ------------------------------
    //   iteration k:  cp.async(tile k + NST - 1)  |  depthwise(k + 1) -> A[(k+1)&1], MMA(k + 1) -> TMEM[(k+1)&1]  |
    //                 wait MMA(k), epilogue(k)
    // so the global->shared copies have NST - 1 iterations to land and every MMA runs under the next tile's depthwise.
    const int nk = blockIdx.x < ntiles ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int accw = P.MT * N;                         // TMEM columns per accumulator buffer
    auto tile_of = [&](int k) { return blockIdx.x + k * gridDim.x; };
    for (int k = 0; k < NST - 1; k++) {
      if (TMA) { if (k < nk && warp == 0) stage_tma(tile_of(k), sT0 + (k % NST) * tile_bytes, &fbar[k % NST]); continue; }
      if (k < nk) stage(tile_of(k), sT0 + (k % NST) * tile_bytes);
      cp_async_commit();
    }
    if (nk > 0) {
      if (TMA) mbar_wait(smem_u32(&fbar[0]), 0u);
      else if (NST == 2) cp_async_wait_all(); else asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncthreads();
      run_dw(sT0, sA0);
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) issue_mma(sA0, tmem_base, &mbar[0]);
    }
    for (int k = 0; k < nk; k++) {
      const int ab = k & 1;
      if (TMA) {
        if (k + NST - 1 < nk && warp == 0) stage_tma(tile_of(k + NST - 1), sT0 + ((k + NST - 1) % NST) * tile_bytes, &fbar[(k + NST - 1) % NST]);
      } else {
        if (k + NST - 1 < nk) stage(tile_of(k + NST - 1), sT0 + ((k + NST - 1) % NST) * tile_bytes);
        cp_async_commit();
      }
      if (k + 1 < nk) {
        if (TMA) mbar_wait(smem_u32(&fbar[(k + 1) % NST]), (uint32_t)(((k + 1) / NST) & 1));
        else if (NST == 2) cp_async_wait_all(); else asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        run_dw(sT0 + ((k + 1) % NST) * tile_bytes, sA0 + (ab ^ 1) * a_bytes);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) issue_mma(sA0 + (ab ^ 1) * a_bytes, tmem_base + (uint32_t)((ab ^ 1) * accw), &mbar[ab ^ 1]);
      }
      mbar_wait(smem_u32(&mbar[ab]), (uint32_t)((k >> 1) & 1));
      tc_fence_after();
      epilogue(tile_of(k), sT0 + (k % NST) * tile_bytes, tmem_base + (uint32_t)(ab * accw));
      tc_fence_before();
      __syncthreads();                                 // TMEM[ab] drained, sT[k % NST] free for tile k + NST
    }
  }
  cp_async_wait_all();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
size_t ds_smem_bytes(const DsParams& P, int S, int TR) {
  const int trin = (TR - 1) * S + 3, tw = P.iw + P.pl + 1;
  const int nab = P.nst > 1 ? 2 : 1;
  size_t b = (size_t)P.N * P.KP + (size_t)nab * P.MT * 128 * P.KP;
  b += (size_t)P.nst * (((size_t)P.NB * trin * tw * P.C + 15) & ~(size_t)15);
  b += (size_t)P.N * 16 + (size_t)((P.N + 1) & ~1) * 4 + 64;     // constants, 2 + 3 mbarriers, TMEM slot
  b += (size_t)P.epi_smem;                          // multiply-high epilogue variants: second constant word per channel + residual table
  return b + 1024;                                   // alignment slack
}

template <int S, int TR, int ADD, int DWT, int MB, int NT = DS_THREADS, int EPI = 0, int NC = 0>
static int launch_one(const int8_t* in, int8_t* out, int Bw, int ntiles, int grid, size_t smem, const DsParams& P, cudaStream_t st) {
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) cudaFuncSetAttribute(k_ds<S, TR, ADD, DWT, MB, NT, EPI, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  k_ds<S, TR, ADD, DWT, MB, NT, EPI, NC><<<grid, NT, smem, st>>>(in, out, Bw, ntiles, P);
  return 0;
}

// ---- stem + first block --------------------------------------------------------------------------------------------------------------
bool ds_stem_supported(const DsParams& P, int S, int add_mode) {
  return S == 2 && add_mode == 0 && P.C == 16 && P.N == 32 && P.nc == 32 && P.iw == 128 && P.ow == 64 && P.pt == 0 && P.pl == 0 &&
         P.NB == 1 && P.oh % 4 == 0 && P.ih == 2 * P.oh && P.KP == 32 && P.RW == 32;
}

size_t ds_stem_smem_bytes(const DsParams& P, int* stem_off) {
  const int TRIN = 9;
  size_t b = ds_smem_bytes(P, 2, 4);                  // includes 1 KB of alignment slack
  b = (b + 1023) & ~(size_t)1023;
  *stem_off = (int)(b - 1024);                        // relative to the 1024-aligned base the kernel computes
  return b + (((TRIN + 2) * 272 + 1023) / 1024 * 1024) + TRIN * 4096 + 1024;
}

int launch_ds_stem(const int8_t* head_out, int8_t* out, int Bw, const DsParams& P, size_t smem, int num_sms, cudaStream_t st) {
  const int ntiles = Bw * (P.oh / 4);
  int grid = num_sms * 2;
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) return 0;
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) cudaFuncSetAttribute(k_ds<2, 4, 0, 1, 2, 256, 0, 32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  k_ds<2, 4, 0, 1, 2, 256, 0, 32, 1><<<grid, 256, smem, st>>>(head_out, out, Bw, ntiles, P);
  return 0;
}

int launch_ds(const int8_t* in, int8_t* out, int Bw, const DsParams& P, const DsLaunch& L, int num_sms, cudaStream_t st) {
  const int ntiles = P.NB == 1 ? Bw * (P.oh / L.TR) : (Bw + P.NB - 1) / P.NB;
  int grid = num_sms * L.ctas_per_sm;
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) return 0;
  // compile-time channel count builds (constants as kernel-parameter operands): the shapes of the 32- and 64-channel blocks
  if (P.nc == P.N && (P.N == 32 || P.N == 64) && L.dwt && L.epi == 0 && L.threads == 256) {
    static const int nc_on = getenv("BN_DS_NC") ? atoi(getenv("BN_DS_NC")) : 1;
    if (nc_on) {
#define NC_CASE(s, tr, add, mb)                                                                                    \
      if (L.S == s && L.TR == tr && L.add_mode == add && mbs == mb)                                                 \
        return P.N == 32 ? launch_one<s, tr, add, 1, mb, 256, 0, 32>(in, out, Bw, ntiles, grid, L.smem, P, st)      \
                         : launch_one<s, tr, add, 1, mb, 256, 0, 64>(in, out, Bw, ntiles, grid, L.smem, P, st)
      const int mbs = L.ctas_per_sm >= 3 ? 3 : 2;
      NC_CASE(2, 8, 0, 3); NC_CASE(2, 8, 0, 2); NC_CASE(2, 4, 0, 3);
      NC_CASE(1, 16, 2, 2); NC_CASE(1, 8, 2, 2);
#undef NC_CASE
    }
  }
  // 512-thread build for single-CTA layers (transposed depthwise only; the two shapes the late layers use)
  // residual-ADD blocks with the multiply-high epilogue (transposed depthwise builds only)
  if (L.add_mode == 2 && L.epi >= 1 && L.dwt && L.S == 1) {
#define EPI_CASE(tr, mb, nt)                                                                                       \
    if (L.TR == tr && mb_sel == mb && L.threads == nt)                                                              \
      return L.epi == 2 ? launch_one<1, tr, 2, 1, mb, nt, 2>(in, out, Bw, ntiles, grid, L.smem, P, st)              \
                        : launch_one<1, tr, 2, 1, mb, nt, 1>(in, out, Bw, ntiles, grid, L.smem, P, st)
    const int mb_sel = (L.threads == 512 && L.ctas_per_sm == 1) ? 1 : (L.ctas_per_sm >= 3 ? 3 : 2);
    EPI_CASE(4, 1, 512);
    EPI_CASE(4, 2, 256); EPI_CASE(4, 3, 256);
    EPI_CASE(8, 2, 256); EPI_CASE(8, 3, 256);
    EPI_CASE(16, 2, 256); EPI_CASE(16, 3, 256);
#undef EPI_CASE
  }
  if (L.threads == 512 && L.dwt && L.ctas_per_sm == 1) {
    if (L.S == 1 && L.TR == 4 && L.add_mode == 2) return launch_one<1, 4, 2, 1, 1, 512>(in, out, Bw, ntiles, grid, L.smem, P, st);
    if (L.S == 1 && L.TR == 4 && L.add_mode == 1) return launch_one<1, 4, 1, 1, 1, 512>(in, out, Bw, ntiles, grid, L.smem, P, st);
    if (L.S == 2 && L.TR == 4 && L.add_mode == 0) return launch_one<2, 4, 0, 1, 1, 512>(in, out, Bw, ntiles, grid, L.smem, P, st);
  }
#define DS_CASE(s, tr, add)                                                                                         \
  if (L.S == s && L.TR == tr && L.add_mode == add)                                                                  \
    return !L.dwt ? launch_one<s, tr, add, 0, 2>(in, out, Bw, ntiles, grid, L.smem, P, st)                          \
         : L.ctas_per_sm >= 3 ? launch_one<s, tr, add, 1, 3>(in, out, Bw, ntiles, grid, L.smem, P, st)              \
                              : launch_one<s, tr, add, 1, 2>(in, out, Bw, ntiles, grid, L.smem, P, st)
  DS_CASE(1, 4, 0); DS_CASE(1, 4, 1); DS_CASE(1, 4, 2);
  DS_CASE(1, 8, 0); DS_CASE(1, 8, 1); DS_CASE(1, 8, 2);
  DS_CASE(2, 4, 0); DS_CASE(2, 8, 0);
  DS_CASE(1, 16, 0); DS_CASE(1, 16, 1); DS_CASE(1, 16, 2); DS_CASE(2, 16, 0);
#undef DS_CASE
  return BN_ERR_UNSUPPORTED;
}

}  // namespace bn
