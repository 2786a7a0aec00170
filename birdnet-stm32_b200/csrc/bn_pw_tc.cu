// bn_pw_tc.cu -- K5tc: pointwise (1x1) convolution as an int8 tensor-core GEMM on sm_100a.
//
//   y[M, N] = requant( x[M, K] . w[N, K]^T + bias' )  (+ residual ADD, + ReLU6 clamp)
//
// tcgen05.mma.cta_group::1.kind::i8 (SASS UTCIMMA), M = 128 rows per instruction, int32 accumulators in
// TMEM, operands in shared memory in the canonical K-major swizzled layouts (32B / 64B / 128B swizzle
// chosen by K).  One CTA is persistent over 128-row tiles:
//
//   iteration i:   wait cp.async(A_i) -> fence.proxy.async -> bar      (A tile i is in smem)
//                  lane 0 of warp 0 issues K/32 MMAs  A_i x B -> TMEM[i&1], tcgen05.commit -> mbar[i&1]
//                  all 8 warps: wait mbar[(i-1)&1]; start cp.async(A_{i+1}) into the freed smem buffer;
//                               epilogue of tile i-1 from TMEM[(i-1)&1] (tcgen05.ld 32x32b.x16)
//
// so the tensor core, the global->smem copies and the CUDA-core epilogue of three consecutive tiles overlap.
// Epilogue per element (bit-exact TFLite semantics): + folded bias, SRDHM +
// RoundingDivideByPOT (closed form rq_fast), + zero point, clamp; for residual blocks the int8 ADD
// (left shift 20, two input rescales, output rescale) and its ReLU6 clamp.  Each lane owns one output row
// (TMEM lane) and 16 consecutive channels per step: 16-byte residual loads and 16-byte stores.
//
// Reference counterpart: CONV_2D 1x1 (+ ADD) inside tf.lite.Interpreter.invoke
// (birdnet_stm32/models/runners.py:93-95; graph built by models/dscnn.py:28-84).
#include "bn_common.cuh"
#include "bn_pw_tc.cuh"
#include "bn_tc.cuh"

namespace bn {

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
constexpr int TC_THREADS = 256;

__global__ void __launch_bounds__(TC_THREADS)
k_pw_tc(const int8_t* __restrict__ x, const int8_t* __restrict__ res, int8_t* __restrict__ y, int M, PwTcParams P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // swizzled operand tiles need 1024-byte alignment of their shared-memory addresses
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KP = P.KP, N = P.N, RW = P.RW;
  const int KH = KP / RW;                       // 128-byte k-halves (1 or 2)
  const int a_bytes = 128 * KP, b_bytes = N * KP;
  unsigned char* sB = smem;
  unsigned char* sA0 = smem + b_bytes;
  int* prm = reinterpret_cast<int*>(sA0 + 2 * a_bytes);      // bias[N] mult[N] shift[N]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(prm + 3 * N);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2);

  const int ntiles = (M + 127) >> 7;
  const uint32_t ncols = P.tmem_cols;

  // ---- one-time setup ------------------------------------------------------------------------
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), ncols);
  if (tid == 32) {
    mbar_init(smem_u32(&mbar[0]), 1);
    mbar_init(smem_u32(&mbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < b_bytes / 16; i += TC_THREADS) cp_async16(smem_u32(sB + 16 * i), P.w_img + 16 * (size_t)i);
  for (int i = tid; i < N; i += TC_THREADS) { prm[i] = __ldg(P.bias + i); prm[N + i] = __ldg(P.mult + i); prm[2 * N + i] = __ldg(P.shift + i); }
  if (P.K < KP) {   // K = 16: the upper 16 bytes of every 32-byte row stay zero
    for (int i = tid; i < 2 * a_bytes / 16; i += TC_THREADS) *reinterpret_cast<uint4*>(sA0 + 16 * i) = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();

  const int cpr = P.K >> 4;                      // 16-byte chunks per row actually loaded
  const int cpr_log = P.cpr_log;
  const int chunks_per_half = RW >> 4;
  auto load_a = [&](int tile, int buf) {
    unsigned char* dstb = sA0 + buf * a_bytes;
    const long m0 = (long)tile * 128;
    for (int i = tid; i < 128 * cpr; i += TC_THREADS) {
      const int r = i >> cpr_log, cc = i & (cpr - 1);
      const int h = cc / chunks_per_half, c = cc - h * chunks_per_half;
      long m = m0 + r;
      if (m >= M) m = M - 1;                     // tail rows: valid memory, results are not stored
      const uint32_t dst = smem_u32(dstb + h * (128 * RW) + r * RW + ((c ^ swz_xor(r, RW)) << 4));
      cp_async16(dst, x + m * P.K + (cc << 4));
    }
  };

  int tile = blockIdx.x;
  if (tile < ntiles) load_a(tile, 0);
  cp_async_commit();

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t sbo = 8 * RW;                   // bytes between 8-row groups
  const uint32_t lt = RW == 128 ? 2u : (RW == 64 ? 4u : 6u);
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
  const int ksteps = KP >> 5;
  const int ksteps_per_half = RW >> 5;

  // epilogue mapping: warp w -> TMEM lanes 32*(w&3).., column half (w>>2)
  const int q = warp & 3, hsel = warp >> 2;
  const int ncol_half = N >> 1;
  const int col_base = hsel * ncol_half;

  auto epilogue = [&](int tl, int buf) {
    const long m = (long)tl * 128 + 32 * q + lane;
    const bool row_ok = m < M;
    for (int c0 = col_base; c0 < col_base + ncol_half; c0 += 16) {
      int v[16];
      tmem_ld16(tmem_base + (uint32_t)(buf * N + c0) + ((uint32_t)(32 * q) << 16), v);
      uint4 rv = make_uint4(0, 0, 0, 0);
      if (P.has_add && row_ok) rv = __ldg(reinterpret_cast<const uint4*>(res + m * N + c0));
      const unsigned rw[4] = {rv.x, rv.y, rv.z, rv.w};
      unsigned ow[4] = {0, 0, 0, 0};
#pragma unroll
      for (int g = 0; g < 4; g++) {
        const int4 bias = *reinterpret_cast<const int4*>(prm + c0 + 4 * g);
        const int4 mult = *reinterpret_cast<const int4*>(prm + N + c0 + 4 * g);
        const int4 shift = *reinterpret_cast<const int4*>(prm + 2 * N + c0 + 4 * g);
        const int bs[4] = {bias.x, bias.y, bias.z, bias.w}, ms[4] = {mult.x, mult.y, mult.z, mult.w}, ss[4] = {shift.x, shift.y, shift.z, shift.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          int o = clampi(rq_fast(v[4 * g + j] + bs[j], ms[j], -ss[j]) + P.out_zp, P.act_min, P.act_max);
          if (P.has_add) {
            const int r8 = (int)(int8_t)((rw[g] >> (8 * j)) & 0xffu);
            int s1 = srdhm((r8 - P.add_in1_zp) * (1 << 20), P.add_m1);
            if (P.add_n1 > 0) s1 = (s1 + (1 << (P.add_n1 - 1)) + (s1 >> 31)) >> P.add_n1;
            int s2 = srdhm((o - P.add_in2_zp) * (1 << 20), P.add_m2);
            if (P.add_n2 > 0) s2 = (s2 + (1 << (P.add_n2 - 1)) + (s2 >> 31)) >> P.add_n2;
            int so = srdhm(s1 + s2, P.add_mo);
            if (P.add_no > 0) so = (so + (1 << (P.add_no - 1)) + (so >> 31)) >> P.add_no;
            o = clampi(so + P.add_out_zp, P.add_act_min, P.add_act_max);
          }
          ow[g] |= (unsigned)(uint8_t)o << (8 * j);
        }
      }
      if (row_ok) *reinterpret_cast<uint4*>(y + m * N + c0) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
  };

  // ---- persistent main loop ------------------------------------------------------------------------
  int it = 0, prev_tile = -1;
  for (; tile < ntiles; tile += gridDim.x, it++) {
    const int buf = it & 1;
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();                               // A[buf] complete; epilogue reads of TMEM[buf] (tile it-2) are done
    if (warp == 0 && lane == 0) {
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sA0 + buf * a_bytes), b_addr = smem_u32(sB);
      for (int ks = 0; ks < ksteps; ks++) {
        const int h = ks / ksteps_per_half, kk = ks - h * ksteps_per_half;
        const uint64_t ad = make_desc(a_addr + h * (128 * RW) + kk * 32, sbo, lt);
        const uint64_t bd = make_desc(b_addr + h * (N * RW) + kk * 32, sbo, lt);
        umma_i8(tmem_base + (uint32_t)(buf * N), ad, bd, idesc, ks > 0 ? 1u : 0u);
      }
      umma_commit(smem_u32(&mbar[buf]));
    }
    if (it > 0) {
      mbar_wait(smem_u32(&mbar[buf ^ 1]), ((it - 1) >> 1) & 1);   // MMA(it-1) done: TMEM[buf^1] ready, A[buf^1] free
      tc_fence_after();
    }
    const int next = tile + gridDim.x;
    if (next < ntiles) load_a(next, buf ^ 1);
    cp_async_commit();
    if (it > 0) epilogue(prev_tile, buf ^ 1);
    prev_tile = tile;
  }
  if (it > 0) {
    const int buf = (it - 1) & 1;
    mbar_wait(smem_u32(&mbar[buf]), ((it - 1) >> 1) & 1);
    tc_fence_after();
    epilogue(prev_tile, buf);
  }
  cp_async_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
bool pw_tc_supported(int K, int N) {
  const bool k_ok = K == 16 || K == 32 || K == 64 || K == 128 || K == 256;
  const bool n_ok = N == 32 || N == 64 || N == 128 || N == 256;
  return k_ok && n_ok;
}

// Pre-swizzled shared-memory image of the weights [N][K] (K-major), zero padded to KP.
void pw_tc_weight_image(const int8_t* w, int K, int N, std::vector<uint8_t>& img, int* KP_out, int* RW_out) {
  const int KP = K < 32 ? 32 : K;
  const int RW = KP > 128 ? 128 : KP;
  const int KH = KP / RW;
  img.assign((size_t)N * KP, 0);
  for (int n = 0; n < N; n++)
    for (int h = 0; h < KH; h++)
      for (int c = 0; c < RW / 16; c++)
        for (int b = 0; b < 16; b++) {
          const int k = h * RW + c * 16 + b;
          const uint8_t v = k < K ? (uint8_t)w[(size_t)n * K + k] : 0;
          img[(size_t)h * N * RW + (size_t)n * RW + ((c ^ swz_xor(n, RW)) << 4) + b] = v;
        }
  *KP_out = KP;
  *RW_out = RW;
}

size_t pw_tc_smem_bytes(const PwTcParams& P) { return (size_t)P.N * P.KP + 2 * 128 * (size_t)P.KP + 3 * P.N * 4 + 64 + 1024; }

int launch_pw_tc(const int8_t* x, const int8_t* res, int8_t* y, long M, const PwTcParams& P, int num_sms, cudaStream_t st) {
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) cudaFuncSetAttribute(k_pw_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (M <= 0 || M > 0x7fffffffL) return BN_ERR_ARG;
  const int ntiles = (int)((M + 127) / 128);
  const size_t smem = pw_tc_smem_bytes(P);
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  const int tmem_limit = 512 / P.tmem_cols;
  if (per_sm > tmem_limit) per_sm = tmem_limit;
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  int grid = num_sms * per_sm;
  if (grid > ntiles) grid = ntiles;
  k_pw_tc<<<grid, TC_THREADS, smem, st>>>(x, res, y, (int)M, P);
  return 0;
}

}  // namespace bn
