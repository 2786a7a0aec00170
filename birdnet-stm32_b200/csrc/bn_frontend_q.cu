// bn_frontend_q.cu -- K1q + K2q: the hybrid frontend without the float32 magnitude scratch.
//
//   K1q  k_stft_q   PCM16 / float32 chunk -> |STFT| (512-point FFT per frame, as K1) -> min / max over the WHOLE chunk
//                   -> normalize() -> QUANTIZE -> int8 codes, written as the ready-made A operand of the mel GEMM
//   K2q  k_head_q   A operand by ONE TMA bulk copy per 128 frames -> tcgen05.mma kind::i8 with the learned mel mixer
//                   -> requant + ReLU -> PWL chain (LUT) -> transposed store int8 [B][64][W]
//
// Reference: get_spectrogram_from_audio (linear branch) + normalize (birdnet_stm32/audio/spectrogram.py:12-21,61,106-115,
// 133,149) followed by the graph's QUANTIZE / CONCAT / mel-mixer CONV_2D / PWL ops (models/frontend.py:299-345,
// models/magnitude.py:179-192).  Same arithmetic as bn_frontend.cu + bn_head_tc.cu, element for element (the quantisation
// code below is the one of bn_head_tc.cu): the int8 codes are identical, only where they live in between changes.
//
// Why: normalize() needs the minimum and maximum of the whole [257, 256] spectrogram before the first code can be formed.
// Round 1 wrote the raw magnitudes to HBM as float32 (248 KB per chunk) and read them back in K2 (270 KB) for a 144 KB
// input.  Here the 8 tiles (32 frames each) of a chunk are worked on by 8 consecutive CTAs of a persistent, co-resident grid
// (cooperative launch).  Every CTA parks its 8,224 magnitudes in TENSOR MEMORY (tcgen05.st: TMEM is otherwise idle in this
// kernel and costs no shared memory, so two CTAs per SM and the cp.async input prefetch stay), publishes its tile's min / max
// with integer atomics and bumps the chunk's arrival counter; one tile LATER -- after the FFTs of its next tile, whose
// magnitudes go to a second TMEM buffer -- it picks the chunk-wide min / max up (by then the other seven tiles have
// long arrived), reads its magnitudes back (tcgen05.ld), quantises them and emits 9 KB of int8: 74 KB per chunk between the
// two kernels instead of 518 KB.
//
// (First version, measured: one thread-block cluster of 8 CTAs per chunk with the min / max exchanged through distributed
// shared memory.  Correct, but `launch__cluster_max_active` = 15 on the B200 -- clusters of 8 are spread one CTA per SM and
// confined to a GPC -- so 120 CTAs were resident instead of 296 and the kernel took 9.9 ms instead of 3.8.  Hence the
// flag-based exchange over the ordinary persistent grid.)
#include <cstdio>
#include <cstdlib>

#include "bn_common.cuh"
#include "bn_fft.cuh"
#include "bn_frontend_q.cuh"
#include "bn_tc.cuh"

namespace bn {

namespace {

constexpr int TILES = 8;                        // tiles (CTAs) per chunk = W / FRAMES_PER_CTA
constexpr int Q_ROWS = FRAMES_PER_CTA;          // A-operand rows produced per CTA
constexpr int Q_KB01 = Q_ROWS * 128;            // bytes of this CTA's rows in each SW128 k-block
constexpr int Q_KB2 = Q_ROWS * 32;              // ... in the SW32 k-block
constexpr int Q_BYTES = 2 * Q_KB01 + Q_KB2;     // 9216
constexpr int TM_BUF = 68;                      // TMEM columns per tile buffer (2 warp halves x 34)
constexpr int TM_COLS = 256;                    // TMEM columns per worker: two tile buffers (136 used)
constexpr int Q_WORKERS = 2;                    // independent 256-thread workers per CTA

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, float v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(v)) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16f(uint32_t taddr, float (&v)[16]) {
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tmem_ld1f(uint32_t taddr) {
  unsigned r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return __uint_as_float(r);
}

// QUANTIZE of the normalised magnitude -- the code of bn_head_tc.cu (see the error bound in its header): one multiply by
// the rounded reciprocal and magic-constant rounding; within 2.5e-4 of a rounding tie the exact two-division chain.
__device__ __forceinline__ int quant_code_q(float f, float mn, float den, float qmul, float scale, int zp_bits, int zp) {
  const float d = f - mn;
  const float t = d * qmul;
  const float tm = t + 12582912.0f;
  const float df = t - (tm - 12582912.0f);
  int q = __float_as_int(tm) - zp_bits;
  if (fabsf(df) > 0.49975f) q = (int)roundf(__fdiv_rn(__fdiv_rn(d, den), scale)) + zp;
  return max(-128, min(127, q));
}

// Two codes at once with packed FP32 pairs (same IEEE results as the scalar form; bn_head_tc.cu: quant_code2).  Fast path only:
// `tie` is raised when either value is within 2.5e-4 of a rounding tie, the caller then redoes the pair with the exact
// two-division chain (quant_code_q).  No clamp: (f - min) / (max - min + 1e-10) <= 1, so the code is in [-128, 127] already.
__device__ __forceinline__ void quant_pair_q(float f0, float f1, float mn, float qmul, int zp_bits, int& q0, int& q1, bool& tie) {
  unsigned long long f, m, k, c, d, t, tm, r, df;
  asm("mov.b64 %0, {%1, %2};" : "=l"(f) : "f"(f0), "f"(f1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(m) : "f"(mn));
  asm("mov.b64 %0, {%1, %1};" : "=l"(k) : "f"(qmul));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(12582912.0f));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f), "l"(m));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(d), "l"(k));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(tm) : "l"(t), "l"(c));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(tm), "l"(c));
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(df) : "l"(t), "l"(r));
  float tm0, tm1, df0, df1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(tm0), "=f"(tm1) : "l"(tm));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(df0), "=f"(df1) : "l"(df));
  q0 = __float_as_int(tm0) - zp_bits;
  q1 = __float_as_int(tm1) - zp_bits;
  tie = tie | (fmaxf(fabsf(df0), fabsf(df1)) > 0.49975f);
}

}  // namespace

__global__ void k_init_minmax_q(unsigned* mnmx, unsigned* arrive, int B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) { mnmx[2 * i] = 0x7f800000u; mnmx[2 * i + 1] = 0u; arrive[i] = 0u; }
}

// dynamic smem: float xs[span + 16] | uint4 sraw[...] | float2 zbuf[16][272] (also the staged A rows, 9216 B) | float red[16] | u32 tmem slot
template <bool F32IN>
__global__ void __launch_bounds__(FE_THREADS * Q_WORKERS, 1)
k_stft_q(const void* __restrict__ pcm_v, const float* __restrict__ peak, uint8_t* __restrict__ aimg, const float4* __restrict__ tables,
         unsigned* __restrict__ mnmx, unsigned* __restrict__ arrive, int T, int hop, int W, int B, FrontendQParams Q, int worker_bytes) {
  using raw_t = typename std::conditional<F32IN, float, int16_t>::type;
  constexpr int G = F32IN ? 4 : 8;
  constexpr int GL = F32IN ? 2 : 3;
  const raw_t* pcm = reinterpret_cast<const raw_t*>(pcm_v);
  extern __shared__ __align__(1024) unsigned char smem_all[];
  // Two independent WORKERS of 256 threads per CTA, each with its own shared-memory slice, named barrier, TMEM columns and tile
  // sequence -- the occupancy of two 256-thread CTAs per SM, but as ONE CTA per SM: the occupancy calculator (and with it the
  // cooperative launch that guarantees the co-residency the inter-tile exchange needs) admits a single CTA per SM for a
  // kernel that allocates tensor memory.
  const int wk = threadIdx.x / FE_THREADS;
  const int span = (FRAMES_PER_CTA - 1) * hop + NFFT;
  unsigned char* smem_raw = smem_all + (size_t)wk * worker_bytes;
  const int vcta = blockIdx.x * Q_WORKERS + wk, vgrid = gridDim.x * Q_WORKERS;     // this worker's place in the persistent tile order
  float* xs = reinterpret_cast<float*>(smem_raw);
  uint4* sraw = reinterpret_cast<uint4*>(xs + ((span + 16 + 3) & ~3));
  float2* zbuf = reinterpret_cast<float2*>(sraw + ((span + 16 + G - 1) >> GL));
  float* red = reinterpret_cast<float*>(zbuf + 16 * (NC + 16));
  // the staged A rows share the FFT exchange buffer: the FFTs of a tile are over (CTA barrier) before the previous tile is
  // finalised, and finalise() ends with a barrier before the next FFTs -- shared memory stays at K1's 90 KB (two CTAs per SM;
  // with a separate 9 KB buffer the occupancy calculator and the cooperative launch only admit ONE CTA per SM)
  unsigned char* sQ = reinterpret_cast<unsigned char*>(zbuf);
  // TMEM base address: an unused padding entry of worker 0's last half-warp buffer
  uint32_t* tmem_slot0 = reinterpret_cast<uint32_t*>(reinterpret_cast<float2*>(smem_all + ((size_t)(reinterpret_cast<unsigned char*>(zbuf) - smem_raw))) + 15 * (NC + 16) + (NC + 15));

  const int tid = threadIdx.x - wk * FE_THREADS, warp = tid >> 5, lane = tid & 31;
  const int hw = tid >> 4;
  const int l = tid & 15;
  float2* zb = zbuf + hw * (NC + 16);
  const unsigned hmask = 0xffffu << (16 * ((tid >> 4) & 1));

  if (threadIdx.x < 32) tmem_alloc(smem_u32(tmem_slot0), TM_COLS * Q_WORKERS);

  const float2* tw512 = reinterpret_cast<const float2*>(tables);
  const float2* win2 = reinterpret_cast<const float2*>(reinterpret_cast<const float*>(tables) + 2 * NFFT);
  float wre[16], wim[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; n1++) {
    const float2 w2 = __ldg(win2 + 16 * n1 + l);
    wre[n1] = 0.5f * w2.x; wim[n1] = 0.5f * w2.y;
  }
  float2 twp[16];
#pragma unroll
  for (int k1 = 1; k1 < 16; k1++) twp[k1] = __ldg(tw512 + ((2 * l * k1) & 511));
  float2 tws[8];
#pragma unroll
  for (int j = 0; j < 8; j++) tws[j] = __ldg(tw512 + l + 16 * j);

  const int a0 = (int)((reinterpret_cast<uintptr_t>(pcm) / sizeof(raw_t)) & (G - 1));
  const raw_t* pcm_al = pcm - a0;
  const long total = (long)B * T;
  const int groups_w = W / FRAMES_PER_CTA;              // == TILES; gridDim.x is a multiple of it: CTAs 8 c .. 8 c + 7 work on the
  const int ntiles = B * groups_w;                      // eight tiles of the same chunk in the same iteration

  auto tile_geom = [&](int tile_id, int& b, int& t0, long& chunk_base, long& g_first, int& shift, int& ngroups) {
    b = tile_id / groups_w;
    t0 = (tile_id - b * groups_w) * FRAMES_PER_CTA;
    chunk_base = (long)b * T + a0;
    const long g_lo = chunk_base + (long)t0 * hop - NFFT / 2;
    g_first = g_lo & ~(long)(G - 1);
    shift = (int)(g_lo - g_first);
    ngroups = (shift + span + G - 1) >> GL;
  };
  auto prefetch = [&](int tile_id) {
    int b, t0, shift, ngroups; long chunk_base, g_first;
    tile_geom(tile_id, b, t0, chunk_base, g_first, shift, ngroups);
    if (g_first >= a0 && g_first + (long)G * ngroups <= a0 + total) {   // whole span inside the buffer: no per-group range checks
      const raw_t* src = pcm_al + g_first;
      for (int grp = tid; grp < ngroups; grp += FE_THREADS) cp_async16_fe(sraw + grp, src + G * grp);
      return;
    }
    for (int grp = tid; grp < ngroups; grp += FE_THREADS) {
      const long g = g_first + (long)G * grp;
      if (g >= a0 && g + G <= a0 + total) {
        cp_async16_fe(sraw + grp, pcm_al + g);
      } else if (F32IN) {
        unsigned fv[4];
#pragma unroll
        for (int j = 0; j < 4; j++) fv[j] = (g + j >= a0 && g + j < a0 + total) ? __float_as_uint((float)pcm_al[g + j]) : 0u;
        sraw[grp] = make_uint4(fv[0], fv[1], fv[2], fv[3]);
      } else {
        unsigned short sv[8];
#pragma unroll
        for (int j = 0; j < 8; j++) sv[j] = (g + j >= a0 && g + j < a0 + total) ? (unsigned short)pcm_al[g + j] : (unsigned short)0;
        sraw[grp] = make_uint4(sv[0] | ((unsigned)sv[1] << 16), sv[2] | ((unsigned)sv[3] << 16), sv[4] | ((unsigned)sv[5] << 16), sv[6] | ((unsigned)sv[7] << 16));
      }
    }
  };

  if (vcta < ntiles) prefetch(vcta);
  asm volatile("cp.async.commit_group;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  auto worker_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(wk + 1), "n"(FE_THREADS) : "memory"); };
  // this thread's TMEM scratch: lane = its lane in the warp's quarter, 34 columns per warp half, two tile buffers
  const uint32_t tm0 = *tmem_slot0 + (uint32_t)(wk * TM_COLS) + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 34);

  // finalise a tile whose magnitudes are parked in TMEM buffer `fbuf`: chunk-wide min / max (all TILES tiles of chunk fb have
  // published theirs once arrive[fb] == TILES), normalize() + QUANTIZE, scatter into the staged A rows, copy them out
  // The counter and the min / max of the previous tile's chunk are loaded BEFORE the FFTs of the current tile (`peek`), so
  // that their latency is hidden; every thread reads them itself (same addresses, no broadcast, no extra barrier).
  auto peek = [&](int fb, unsigned& seen, unsigned& umn, unsigned& umx) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(arrive + fb) : "memory");
    umn = __ldcg(mnmx + 2 * fb);
    umx = __ldcg(mnmx + 2 * fb + 1);
  };
  auto finalise = [&](int fb, int ft0, int fbuf, unsigned seen, unsigned umn, unsigned umx) {
    // lane 0 of every warp holds the peeked values (an acquire load invalidates L1: one lane per warp does it, not all 32)
    float den = 0.0f, qmul = 0.0f, mn = 0.0f;
    if (lane == 0) {
      for (unsigned spin = 0; seen < (unsigned)groups_w; spin++) {   // rare: a tile of the chunk had not been published at peek time
        if (spin > (1u << 22)) __trap();                  // ~1 s: the grid is not co-resident -- fail the launch, never hang the GPU
        __nanosleep(64);
        peek(fb, seen, umn, umx);
      }
      mn = __uint_as_float(umn);
      const float mx = __uint_as_float(umx);
      // normalize(): numpy scalar promotion (float64 add, float32 result)
      den = (float)((double)(mx - mn) + 1e-10);
      qmul = (float)(1.0 / ((double)den * (double)Q.q_scale));
    }
    mn = __shfl_sync(0xffffffffu, mn, 0);
    den = __shfl_sync(0xffffffffu, den, 0);
    qmul = __shfl_sync(0xffffffffu, qmul, 0);
    const uint32_t ftm = tm0 + (uint32_t)(fbuf * TM_BUF);
    const int zp_bits = 0x4B400000 - Q.q_zp;
    const unsigned fill7 = 0x01010101u * (unsigned)(uint8_t)Q.fill;
    // ---- read the magnitudes back, quantise, scatter the codes into the staged A rows (swizzled K-major image) ----
#pragma unroll 1
    for (int round = 0; round < FRAMES_PER_CTA / 16; round++) {
      const int f = round * 16 + hw;                      // row of this CTA's 32-row slice
      const int r = (ft0 + f) & 127;                       // row within the 128-frame MMA tile
      float mg[16];
      tmem_ld16f(ftm + (uint32_t)(round * 17), mg);
      const float m128 = tmem_ld1f(ftm + (uint32_t)(round * 17 + 16));
      unsigned char* row01 = sQ + f * 128;
      const int sx = (r & 7) << 4;
      int qa[8], qb[8];
      bool tie = false;
#pragma unroll
      for (int j = 0; j < 8; j++) quant_pair_q(mg[2 * j], mg[2 * j + 1], mn, qmul, zp_bits, qa[j], qb[j], tie);
      if (tie) {                                          // rare: redo this frame's values with the exact chain
#pragma unroll
        for (int j = 0; j < 8; j++) {
          qa[j] = quant_code_q(mg[2 * j], mn, den, qmul, Q.q_scale, zp_bits, Q.q_zp);
          qb[j] = quant_code_q(mg[2 * j + 1], mn, den, qmul, Q.q_scale, zp_bits, Q.q_zp);
        }
      }
      // bin k = l + 16 j (< 128): k-block 0, 16-byte chunk j (XOR-swizzled by the row), byte l.
      // bin 256 - k: for l > 0 it is 16 (15 - j) + (16 - l) -> k-block 1, chunk 7 - j, byte 16 - l;
      // for l == 0 it is 16 (16 - j): j > 0 -> k-block 1, chunk 8 - j, byte 0; j == 0 -> bin 256 (third k-block)
      unsigned char* pa = row01 + l;
      unsigned char* pb = row01 + Q_KB01 + (l > 0 ? 16 - l : 0);
#pragma unroll
      for (int j = 0; j < 8; j++) {
        pa[(j << 4) ^ sx] = (unsigned char)qa[j];
        if (l > 0) pb[((7 - j) << 4) ^ sx] = (unsigned char)qb[j];
        else if (j > 0) pb[((8 - j) << 4) ^ sx] = (unsigned char)qb[j];
      }
      if (l == 0) {
        // k = 256 (real bin) and 257 .. 263 (the FILL columns of the graph's CONCAT): SW32 block, chunk 0 ^ ((r >> 2) & 1)
        // (bytes 8 .. 15 of that chunk and the row's other chunk pair with zero weights: written as zeros)
        const int cx = ((r >> 2) & 1) << 4;
        unsigned char* p2 = sQ + 2 * Q_KB01 + f * 32;
        *reinterpret_cast<uint4*>(p2 + cx) = make_uint4((unsigned)(unsigned char)qb[0] | (fill7 << 8), fill7, 0u, 0u);
        *reinterpret_cast<uint4*>(p2 + (cx ^ 16)) = make_uint4(0u, 0u, 0u, 0u);
      }
      if (l == 0) row01[Q_KB01 + sx] = (unsigned char)quant_code_q(m128, mn, den, qmul, Q.q_scale, zp_bits, Q.q_zp);   // bin 128: k-block 1, chunk 0, byte 0
    }
    worker_sync();
    // ---- 9 KB of A rows -> global: rows [t0 % 128, +32) of the chunk's tile (t0 / 128), three contiguous pieces ----
    {
      uint8_t* dst = aimg + ((size_t)fb * (W / 128) + (ft0 >> 7)) * (size_t)HQ_A_BYTES;
      const int r0 = ft0 & 127;
      const uint4* s4 = reinterpret_cast<const uint4*>(sQ);
      for (int i = tid; i < Q_BYTES / 16; i += FE_THREADS) {
        const int piece = i < Q_KB01 / 16 ? 0 : (i < 2 * Q_KB01 / 16 ? 1 : 2);
        const int w = i - piece * (Q_KB01 / 16);
        uint8_t* d = piece < 2 ? dst + (size_t)piece * (128 * 128) + (size_t)r0 * 128 + 16 * (size_t)w
                               : dst + (size_t)2 * (128 * 128) + (size_t)r0 * 32 + 16 * (size_t)w;
        *reinterpret_cast<uint4*>(d) = s4[i];
      }
    }
    // no barrier here: sQ (= the FFT exchange buffer) is next written by the FFTs of the following tile, which start
    // behind the worker barrier that follows the sample conversion
  };

  int it = 0, prev_b = 0, prev_t0 = 0;
  for (int tile_id = vcta; tile_id < ntiles; tile_id += vgrid, it++) {
    int b, t0, shift, ngroups; long chunk_base, g_first;
    tile_geom(tile_id, b, t0, chunk_base, g_first, shift, ngroups);
    const float pk = peak ? __ldg(peak + b) : 0.0f;
    const float cs = F32IN ? (pk > 0.0f ? __fdiv_rn(1.0f, pk) : 1.0f) : (pk > 0.0f ? __fdiv_rn(1.0f, 32768.0f * pk) : (1.0f / 32768.0f));
    unsigned p_seen = 0, p_mn = 0, p_mx = 0;
    if (it > 0 && lane == 0) peek(prev_b, p_seen, p_mn, p_mx);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const int rel0 = (int)(g_first - chunk_base);
    const bool inside = rel0 >= 0 && rel0 + G * ngroups <= T;   // no zero padding in this tile (all but the chunk's first / last)
    if (!F32IN && inside) {
      for (int grp = tid; grp < ngroups; grp += FE_THREADS) {
        const uint4 wv = sraw[grp];
        const unsigned ww[4] = {wv.x, wv.y, wv.z, wv.w};
        float fv[8];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int lo = (int)(short)(ww[j] & 0xffffu), hi = (int)ww[j] >> 16;
          fv[2 * j] = (__int_as_float(0x4B400000 + lo) - 12582912.0f) * cs;
          fv[2 * j + 1] = (__int_as_float(0x4B400000 + hi) - 12582912.0f) * cs;
        }
        *reinterpret_cast<float4*>(xs + 8 * grp) = make_float4(fv[0], fv[1], fv[2], fv[3]);
        *reinterpret_cast<float4*>(xs + 8 * grp + 4) = make_float4(fv[4], fv[5], fv[6], fv[7]);
      }
    } else
    for (int grp = tid; grp < ngroups; grp += FE_THREADS) {
      const long rel = g_first + (long)G * grp - chunk_base;
      const uint4 wv = sraw[grp];
      if (F32IN) {
        float f4[4] = {__uint_as_float(wv.x), __uint_as_float(wv.y), __uint_as_float(wv.z), __uint_as_float(wv.w)};
#pragma unroll
        for (int j = 0; j < 4; j++) f4[j] = (rel + j >= 0 && rel + j < T) ? f4[j] * cs : 0.0f;
        *reinterpret_cast<float4*>(xs + 4 * grp) = make_float4(f4[0], f4[1], f4[2], f4[3]);
        continue;
      }
      unsigned ww[4] = {wv.x, wv.y, wv.z, wv.w};
      if (rel < 0 || rel + 8 > T) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const long r0 = rel + 2 * j, r1 = r0 + 1;
          if (!(r0 >= 0 && r0 < T)) ww[j] &= 0xffff0000u;
          if (!(r1 >= 0 && r1 < T)) ww[j] &= 0x0000ffffu;
        }
      }
      float fv[8];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int lo = (int)(short)(ww[j] & 0xffffu), hi = (int)ww[j] >> 16;
        fv[2 * j] = (__int_as_float(0x4B400000 + lo) - 12582912.0f) * cs;
        fv[2 * j + 1] = (__int_as_float(0x4B400000 + hi) - 12582912.0f) * cs;
      }
      *reinterpret_cast<float4*>(xs + 8 * grp) = make_float4(fv[0], fv[1], fv[2], fv[3]);
      *reinterpret_cast<float4*>(xs + 8 * grp + 4) = make_float4(fv[4], fv[5], fv[6], fv[7]);
    }
    if (tile_id + vgrid < ntiles) prefetch(tile_id + vgrid);
    asm volatile("cp.async.commit_group;" ::: "memory");
    worker_sync();

    // ---- FFT of the 32 frames; the magnitudes go to this thread's TMEM columns: per round 16 paired bins + bin 128 ----
    const uint32_t tm = tm0 + (uint32_t)((it & 1) * TM_BUF);
    float lmin = __int_as_float(0x7f800000), lmax = 0.0f;
#pragma unroll 1
    for (int round = 0; round < FRAMES_PER_CTA / 16; round++) {
      const int f = round * 16 + hw;
      const float* xf = xs + shift + f * hop + 2 * l;
      float2 v[16];
#pragma unroll
      for (int n1 = 0; n1 < 16; n1++) v[n1] = make_float2(xf[32 * n1] * wre[n1], xf[32 * n1 + 1] * wim[n1]);
      fft16(v);
      zb[l] = v[0];
#pragma unroll
      for (int k1 = 1; k1 < 16; k1++) zb[k1 * 17 + l] = cmul(v[k1], twp[k1]);
      __syncwarp(hmask);
#pragma unroll
      for (int n2 = 0; n2 < 16; n2++) v[n2] = zb[l * 17 + n2];
      __syncwarp(hmask);
      fft16(v);
      float mg[16];                                       // mg[2 j] = |X[l + 16 j]|, mg[2 j + 1] = |X[256 - (l + 16 j)]|
      auto split_step = [&](const float2 zk, const float2 zn, const int j) {   // partner by register shuffle: see bn_frontend.cu
        const float2 e = make_float2(zk.x + zn.x, zk.y - zn.y);
        const float2 o = make_float2(zk.y + zn.y, zn.x - zk.x);
        const float2 t = cmul(o, tws[j]);
        const float2 xa = add2(e, t), xb = sub2(e, t);
        const float ma = fast_sqrt(xa.x * xa.x + xa.y * xa.y), mb = fast_sqrt(xb.x * xb.x + xb.y * xb.y);
        mg[2 * j] = ma; mg[2 * j + 1] = mb;
        lmin = fminf(lmin, fminf(ma, mb));
        lmax = fmaxf(lmax, fmaxf(ma, mb));
      };
      const int lw = threadIdx.x & 31;
      split_step(v[0], split_partner<0>(v, l, lw), 0); split_step(v[1], split_partner<1>(v, l, lw), 1);
      split_step(v[2], split_partner<2>(v, l, lw), 2); split_step(v[3], split_partner<3>(v, l, lw), 3);
      split_step(v[4], split_partner<4>(v, l, lw), 4); split_step(v[5], split_partner<5>(v, l, lw), 5);
      split_step(v[6], split_partner<6>(v, l, lw), 6); split_step(v[7], split_partner<7>(v, l, lw), 7);
      const float2 z128 = v[8];                           // bin 128 pairs with itself: |X[128]| = 2 |Z[128]| (meaningful in lane l == 0)
      const float m128 = 2.0f * fast_sqrt(z128.x * z128.x + z128.y * z128.y);
      if (l == 0) { lmin = fminf(lmin, m128); lmax = fmaxf(lmax, m128); }
      __syncwarp();
      tmem_st16(tm + (uint32_t)(round * 17), mg);
      tmem_st1(tm + (uint32_t)(round * 17 + 16), m128);
      __syncwarp(hmask);
    }
    tmem_wait_st();

    // ---- min / max of the tile -> global (integer atomics on the bit patterns: magnitudes are >= +0), then arrive ----
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
      lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    }
    if (lane == 0) { red[warp] = lmin; red[8 + warp] = lmax; }
    worker_sync();                                      // also: every warp is done with xs / zbuf of this tile
    if (tid == 0) {
      float mn = red[0], mx = red[8];
      for (int i = 1; i < 8; i++) { mn = fminf(mn, red[i]); mx = fmaxf(mx, red[8 + i]); }
      // fire-and-forget reductions; the release on the counter orders them before it without stalling this thread on a fence
      asm volatile("red.relaxed.gpu.global.min.u32 [%0], %1;" ::"l"(mnmx + 2 * b), "r"(__float_as_uint(mn)) : "memory");
      asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(mnmx + 2 * b + 1), "r"(__float_as_uint(mx)) : "memory");
      asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(arrive + b), "r"(1u) : "memory");
    }
    // ---- the PREVIOUS tile of this CTA: its chunk's other tiles were published an FFT pass ago ----
    if (it > 0) finalise(prev_b, prev_t0, (it - 1) & 1, p_seen, p_mn, p_mx);
    prev_b = b; prev_t0 = t0;
  }
  if (it > 0) finalise(prev_b, prev_t0, (it - 1) & 1, 0u, 0u, 0u);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(*tmem_slot0, TM_COLS * Q_WORKERS);
}

// ---------------------------------------------------------------------------------------------------------------
// K2q: mel-mixer GEMM on the int8 A image
// ---------------------------------------------------------------------------------------------------------------
namespace {
constexpr int HQ_THREADS = 256;
constexpr int HQ_CTAS = 2;
constexpr int HQ_M = 128;
constexpr int HQ_OFF_A = HT_B_BYTES;                 // 18432 (1024-aligned)
constexpr int HQ_OFF_LUT = HQ_OFF_A + HQ_A_BYTES;    // 55296
constexpr int HQ_OFF_OUT = HQ_OFF_LUT + HT_N * 256;  // 71680: transposed output tile [64][128]
constexpr int HQ_OFF_RQ = HQ_OFF_OUT + HT_N * HQ_M;  // 79872
constexpr int HQ_OFF_BAR = HQ_OFF_RQ + HT_N * 16;    // 80896
constexpr int HQ_SMEM = HQ_OFF_BAR + 32 + 1024;

__device__ __forceinline__ void bulk_g2s_q(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_q(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
}  // namespace

__global__ void __launch_bounds__(HQ_THREADS, HQ_CTAS)
k_head_q(const uint8_t* __restrict__ aimg, int8_t* __restrict__ out, int ntiles, HeadTcParams P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* sB = smem;
  unsigned char* sA = smem + HQ_OFF_A;
  unsigned char* sLut = smem + HQ_OFF_LUT;
  unsigned char* sOut = smem + HQ_OFF_OUT;
  int4* s_rq = reinterpret_cast<int4*>(smem + HQ_OFF_RQ);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + HQ_OFF_BAR);    // [0] MMA done, [1] A tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 64);
  if (tid == 32) {
    mbar_init(smem_u32(&mbar[0]), 1);
    mbar_init(smem_u32(&mbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < HT_B_BYTES / 16; i += HQ_THREADS) cp_async16(smem_u32(sB + 16 * i), P.w_img + 16 * (size_t)i);
  for (int i = tid; i < HT_N * 256 / 16; i += HQ_THREADS) cp_async16(smem_u32(sLut + 16 * i), P.lut + 16 * (size_t)i);
  cp_async_commit();
  if (tid < HT_N) s_rq[tid] = __ldg(P.rq + tid);
  cp_async_wait_all();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = make_idesc_i8(HQ_M, HT_N);
  const int halves = P.W / HQ_M;
  const int q = warp & 3, hsel = warp >> 2;

  if (tid == 0 && (int)blockIdx.x < ntiles) {               // first A tile: one TMA bulk copy of 36,864 bytes
    mbar_expect_tx_q(smem_u32(&mbar[1]), HQ_A_BYTES);
    bulk_g2s_q(smem_u32(sA), aimg + (size_t)blockIdx.x * HQ_A_BYTES, HQ_A_BYTES, smem_u32(&mbar[1]));
  }
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
    const int b = tile / halves, t0 = (tile - b * halves) * HQ_M;
    mbar_wait(smem_u32(&mbar[1]), (uint32_t)(it & 1));
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
#pragma unroll
      for (int ks = 0; ks < 8; ks++) {
        const int h = ks >> 2, kk = ks & 3;
        umma_i8(tmem_base, make_desc(a_addr + h * (HQ_M * 128) + kk * 32, 1024, 2u),
                make_desc(b_addr + h * (HT_N * 128) + kk * 32, 1024, 2u), idesc, ks > 0 ? 1u : 0u);
      }
      umma_i8(tmem_base, make_desc(a_addr + 2 * HQ_M * 128, 256, 6u), make_desc(b_addr + 2 * HT_N * 128, 256, 6u), idesc, 1u);
      umma_commit(smem_u32(&mbar[0]));
    }
    mbar_wait(smem_u32(&mbar[0]), (uint32_t)(it & 1));
    tc_fence_after();
    // the MMAs have consumed sA: fetch the next tile under the epilogue
    if (tid == 0 && tile + (int)gridDim.x < ntiles) {
      mbar_expect_tx_q(smem_u32(&mbar[1]), HQ_A_BYTES);
      bulk_g2s_q(smem_u32(sA), aimg + (size_t)(tile + gridDim.x) * HQ_A_BYTES, HQ_A_BYTES, smem_u32(&mbar[1]));
    }
#pragma unroll
    for (int g = 0; g < 2; g++) {
      const int c0 = hsel * 32 + g * 16;
      int v[16];
      tmem_ld16(tmem_base + (uint32_t)c0 + ((uint32_t)(32 * q) << 16), v);
#pragma unroll
      for (int j = 0; j < 16; j++) {
        const int c = c0 + j;
        const int4 rq = s_rq[c];
        int y = rq_hi(v[j], rq.x, rq.y, rq.z) >> rq.w;
        y = max(-128, min(127, y));
        sOut[c * HQ_M + 32 * q + lane] = sLut[c * 256 + y + 128];
      }
    }
    tc_fence_before();
    __syncthreads();
    int8_t* ob = out + (size_t)b * HT_N * P.W + t0;
    for (int i = tid; i < HT_N * HQ_M / 16; i += HQ_THREADS) {
      const int c = i >> 3, piece = i & 7;
      *reinterpret_cast<uint4*>(ob + (size_t)c * P.W + 16 * piece) = *reinterpret_cast<const uint4*>(sOut + c * HQ_M + 16 * piece);
    }
    __syncthreads();                                      // sOut and the TMEM accumulator are free for the next tile
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

// ---------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------
extern const float4* stft_tables_shared();               // bn_frontend.cu

static size_t stft_q_smem_bytes(int hop, bool f32) {
  const int span = (FRAMES_PER_CTA - 1) * hop + NFFT;
  size_t b = sizeof(float) * ((span + 16 + 3) & ~3) + 16 * (size_t)(f32 ? ((span + 16 + 3) >> 2) : ((span + 16 + 7) >> 3));
  b += sizeof(float2) * 16 * (NC + 16) + sizeof(float) * 16;     // the staged A rows (Q_BYTES) alias the exchange buffer
  return b;
}

bool frontend_q_supported(int n_fft, int W, int ldk, int K_real, int hop, int f32) {
  return n_fft == NFFT && W == TILES * FRAMES_PER_CTA && ldk == 264 && K_real == 257 && ((stft_q_smem_bytes(hop, f32 != 0) + 1023) & ~(size_t)1023) * Q_WORKERS + 1024 <= 227 * 1024;
}

// mnmx: uint32 [2 B] min / max bit patterns, arrive: uint32 [B] tiles published per chunk (both initialised here)
int launch_stft_q(const void* pcm, int f32, const float* peak, uint8_t* aimg, unsigned* mnmx, unsigned* arrive, int B, int T, int n_fft,
                  int hop, int W, const FrontendQParams& Q, int num_sms, cudaStream_t st) {
  if (!frontend_q_supported(n_fft, W, 264, 257, hop, f32)) return BN_ERR_UNSUPPORTED;
  if (B < 1) return 0;
  const size_t wbytes = (stft_q_smem_bytes(hop, f32 != 0) + 1023) & ~(size_t)1023;
  const size_t smem = wbytes * Q_WORKERS;
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) {
    cudaFuncSetAttribute(k_stft_q<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(k_stft_q<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  const float4* tab = stft_tables_shared();
  if (!tab) return BN_ERR_CUDA;
  // Persistent grid of CO-RESIDENT CTAs (the tiles of a chunk wait for each other through a global counter): one CTA of two
  // workers per SM, a whole number of chunks (8 workers) per grid, launched cooperatively so that the driver refuses the
  // launch rather than queueing CTAs that others would wait for.
  int grid = (num_sms * Q_WORKERS / TILES) * TILES / Q_WORKERS;
  if (grid * Q_WORKERS > B * TILES) grid = B * TILES / Q_WORKERS;
  if (grid < TILES / Q_WORKERS) return BN_ERR_UNSUPPORTED;
  k_init_minmax_q<<<(B + 255) / 256, 256, 0, st>>>(mnmx, arrive, B);
  int wb = (int)wbytes;
  void* args[] = {(void*)&pcm, (void*)&peak, (void*)&aimg, (void*)&tab, (void*)&mnmx, (void*)&arrive, (void*)&T, (void*)&hop, (void*)&W, (void*)&B, (void*)&Q, (void*)&wb};
  const void* fn = f32 ? (const void*)k_stft_q<true> : (const void*)k_stft_q<false>;
  const cudaError_t le = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(FE_THREADS * Q_WORKERS), args, smem, st);
  if (le != cudaSuccess) {
    if (getenv("BN_DEBUG")) fprintf(stderr, "launch_stft_q: cooperative launch of %d CTAs failed: %s\n", grid, cudaGetErrorString(le));
    cudaGetLastError();
    return BN_ERR_CUDA;
  }
  return 0;
}

int launch_head_q(const uint8_t* aimg, int8_t* out, int Bw, const HeadTcParams& P, int num_sms, cudaStream_t st) {
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) cudaFuncSetAttribute(k_head_q, cudaFuncAttributeMaxDynamicSharedMemorySize, HQ_SMEM);
  if (P.W % HQ_M) return BN_ERR_UNSUPPORTED;
  const int ntiles = Bw * (P.W / HQ_M);
  int grid = num_sms * (getenv("BN_HEAD_CTAS") ? atoi(getenv("BN_HEAD_CTAS")) : 2);
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) return 0;
  k_head_q<<<grid, HQ_THREADS, HQ_SMEM, st>>>(aimg, out, ntiles, P);
  return 0;
}

}  // namespace bn
