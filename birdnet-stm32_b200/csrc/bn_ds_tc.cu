// bn_ds_tc.cu -- K45t: one stride-1 depthwise-separable block per kernel with BOTH convolutions on the tensor core.
//
//   in  int8 [B][ih][iw][C]
//       --cp.async--> channel-chunk planes in shared memory: plane kc = pixels x 16 bytes (channels 16 kc .. 16 kc + 15)
//       --DW 3x3 as 9 shifted tcgen05.mma kind::i8 per 32-channel group: A = the planes, read through a NO-SWIZZLE
//         K-major descriptor whose start address is moved by (dy * TW + dx) pixels for tap (dy, dx);
//         B = diag(w[tap]) 32 x 32; int32 accumulators in TMEM-->
//       --epilogue 1: TMEM -> requant + ReLU6 (saturating form) -> the swizzled A operand of the pointwise conv-->
//       --tcgen05.mma kind::i8 with the pointwise weights --> TMEM (aliases the depthwise accumulators)
//       --epilogue 2: requant (+ residual ADD from the planes still resident in shared memory) -->
//   out int8 [B][oh][ow][N]
//
// Why it works.  In the canonical no-swizzle K-major layout ((8, m), 2) : ((16 B, SBO), LBO) a matrix row is one
// 16-byte unit per K chunk; with SBO = 128 B consecutive rows are consecutive 16-byte units, i.e. consecutive pixels of
// a plane.  The depthwise output at padded pixel index q = (row * TW + col) is sum_taps in[q + dy * TW + dx] * w[tap],
// so every tap is the same M x 32 operand seen from a start address shifted by whole pixels, and the 3 x 3 window costs
// no CUDA-core instruction at all.  Rows of the accumulator that fall on halo columns / rows are computed and ignored.
// The diagonal B wastes 31/32 of the tensor work, which is free here: the tensor pipe is otherwise > 95 % idle.
//
// Reference counterpart: ds_conv_block (birdnet_stm32/models/dscnn.py:28-84) as lowered into DEPTHWISE_CONV_2D,
// CONV_2D 1x1, ADD and executed by tf.lite.Interpreter.invoke (models/runners.py:93-95).  Bit-exact.
#include "bn_ds.cuh"

#include "bn_common.cuh"
#include "bn_tc.cuh"

namespace bn {

constexpr int DST_THREADS = 256;

// no-swizzle K-major descriptor: rows 16 B apart inside an 8-row group, groups SBO apart, the two K chunks LBO apart
__device__ __forceinline__ uint64_t make_desc_ns(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ int rq64s(int acc, int c_lo, int c_hi, int mult, int rz, int n) {
  const long long c = ((long long)c_hi << 32) | (unsigned)c_lo;
  const long long p = (long long)acc * (long long)mult + c;
  const int v = (int)(p >> 31);
  return (v + rz + (v >> 31)) >> n;
}

template <int ADD>
__global__ void __launch_bounds__(DST_THREADS, 3)
k_dst(const int8_t* __restrict__ in, int8_t* __restrict__ out, int Bw, int ntiles, DsParams P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = P.C, N = P.N, KP = P.KP, RW = P.RW;
  const int CK = C >> 4, G = C >> 5;                  // 16-byte channel chunks, 32-channel groups
  const int TR = P.TRr, TRIN = TR + 2, TW = P.ow + 2;
  const int b_bytes = N * KP, a_bytes = P.MT * 128 * KP, d_bytes = G * 9 * 1024;
  const int plane_bytes = P.plane_px * 16;
  unsigned char* sB = smem;
  unsigned char* sA = sB + b_bytes;
  unsigned char* sD = sA + a_bytes;
  unsigned char* sP = sD + d_bytes;
  int4* s_rq = reinterpret_cast<int4*>(sP + CK * plane_bytes);
  int4* s_drq = s_rq + N;
  int* s_rz = reinterpret_cast<int*>(s_drq + C);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(s_rz + ((N + 1) & ~1));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2);

  // ---- one-time setup -----------------------------------------------------------------------------------------
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), (uint32_t)P.tmem_cols);
  if (tid == 32) {
    mbar_init(smem_u32(&mbar[0]), 1);
    mbar_init(smem_u32(&mbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < b_bytes / 16; i += DST_THREADS) cp_async16(smem_u32(sB + 16 * i), P.w_img + 16 * (size_t)i);
  for (int i = tid; i < d_bytes / 16; i += DST_THREADS) cp_async16(smem_u32(sD + 16 * i), P.dw_img + 16 * (size_t)i);
  cp_async_commit();
  for (int i = tid; i < N; i += DST_THREADS) { s_rq[i] = __ldg(P.pw_rq + i); s_rz[i] = __ldg(P.pw_rz + i); }
  for (int i = tid; i < C; i += DST_THREADS) s_drq[i] = __ldg(P.dw_rq + i);
  // the whole plane area starts as the input zero point: halo columns (and the tail the last M tile reads past the
  // tile) are never written again
  {
    const unsigned zpw = 0x01010101u * (unsigned)(uint8_t)P.dw_in_zp;
    for (int i = tid; i < CK * plane_bytes / 16; i += DST_THREADS) *reinterpret_cast<uint4*>(sP + 16 * i) = make_uint4(zpw, zpw, zpw, zpw);
  }
  cp_async_wait_all();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t sbo = 8 * RW;
  const uint32_t lt = RW == 128 ? 2u : (RW == 64 ? 4u : 6u);
  const uint32_t idesc_pw = make_idesc_i8(128, N);
  const uint32_t idesc_dw = make_idesc_i8(128, 32);
  const int ksteps = KP >> 5, ksteps_per_half = RW >> 5;
  const int q = warp & 3, hsel = warp >> 2;
  const int NG = N >> 4;
  const int tiles_per_chunk = P.oh / TR;
  const unsigned zpw = 0x01010101u * (unsigned)(uint8_t)P.dw_in_zp;
  const int ck_log = P.cg_log - 2;                    // log2(C / 16)

  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
    int b0, oy0;
    if (P.NB == 1) { b0 = tile / tiles_per_chunk; oy0 = (tile - b0 * tiles_per_chunk) * TR; }
    else { b0 = tile * P.NB; oy0 = 0; }
    // ---- (1) stage the input rows into the planes (SAME padding rows = zero point) ----------------------------------
    for (int row = warp; row < P.NB * TRIN; row += DST_THREADS / 32) {
      const int bb = row / TRIN, tr = row - bb * TRIN;
      const int iy = oy0 - 1 + tr;
      const bool ok = (b0 + bb) < Bw && iy >= 0 && iy < P.ih;
      const int px0 = row * TW + 1;                    // plane pixel index of image column 0
      const int8_t* src = in + (((size_t)(b0 + bb) * P.ih + iy) * P.iw) * C;
      const int pieces = P.iw << ck_log;
      if (ok) {
        for (int p = lane; p < pieces; p += 32) {
          const int ix = p >> ck_log, kc = p & (CK - 1);
          cp_async16(smem_u32(sP + kc * plane_bytes + (px0 + ix) * 16), src + 16 * (size_t)p);
        }
      } else {
        for (int p = lane; p < pieces; p += 32) {
          const int ix = p >> ck_log, kc = p & (CK - 1);
          *reinterpret_cast<uint4*>(sP + kc * plane_bytes + (px0 + ix) * 16) = make_uint4(zpw, zpw, zpw, zpw);
        }
      }
    }
    cp_async_commit();
    cp_async_wait_all();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- (2) depthwise 3x3 on the tensor core: 9 shifted MMAs per (M tile, 32-channel group) ----------------------
    if (tid == 0) {
      tc_fence_after();
      const uint32_t p_addr = smem_u32(sP), d_addr = smem_u32(sD);
      for (int j = 0; j < P.MTd; j++) {
        for (int g = 0; g < G; g++) {
          const uint32_t a0 = p_addr + (uint32_t)(2 * g) * plane_bytes + (uint32_t)(j * 128) * 16;
          const uint32_t dcol = tmem_base + (uint32_t)(j * C + g * 32);
#pragma unroll
          for (int t = 0; t < 9; t++) {
            const int dy = t / 3, dx = t - 3 * dy;
            const uint64_t ad = make_desc_ns(a0 + (uint32_t)(dy * TW + dx) * 16, (uint32_t)plane_bytes, 128u);
            const uint64_t bd = make_desc_ns(d_addr + (uint32_t)(g * 9 + t) * 1024, 512u, 128u);
            umma_i8(dcol, ad, bd, idesc_dw, t > 0 ? 1u : 0u);
          }
        }
      }
      umma_commit(smem_u32(&mbar[0]));
    }
    mbar_wait(smem_u32(&mbar[0]), (uint32_t)(it & 1));
    tc_fence_after();
    // ---- (3) epilogue 1: depthwise accumulators -> requant + ReLU6 -> pointwise A operand ----------------------------
    for (int j = 0; j < P.MTd; j++) {
      const int qi = j * 128 + 32 * q + lane;          // padded pixel index of this TMEM lane
      const int rowi = qi / TW, ox = qi - rowi * TW;
      const int bb = rowi / TRIN, r = rowi - bb * TRIN;
      const bool valid = ox < P.ow && r < TR && bb < P.NB;
      if (!__any_sync(0xffffffffu, valid)) continue;
      const int m = (bb * TR + r) * P.ow + ox;         // row of the pointwise GEMM
      const int mrow = m & 127;
      unsigned char* arow = sA + (m >> 7) * (128 * KP) + mrow * RW;
      const int sx = (mrow >> P.sw_sh) & P.sw_mask;
      for (int gi = hsel; gi < CK; gi += 2) {
        int v[16];
        tmem_ld16(tmem_base + (uint32_t)(j * C + 16 * gi) + ((uint32_t)(32 * q) << 16), v);
        unsigned w4[4];
#pragma unroll
        for (int gg = 0; gg < 4; gg++) {
          int o[4];
#pragma unroll
          for (int jj = 0; jj < 4; jj++) {
            const int4 rq = s_drq[16 * gi + 4 * gg + jj];
            o[jj] = rq_hi(v[4 * gg + jj], rq.x, rq.y, rq.z) >> rq.w;
          }
          w4[gg] = pack4_sat(o[0], o[1], o[2], o[3]);
        }
        if (valid) {
          const int k0 = 16 * gi;
          const int off = (k0 >> P.rw_log) * (128 * RW) + ((((k0 & (RW - 1)) >> 4) ^ sx) << 4);
          *reinterpret_cast<uint4*>(arow + off) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- (4) pointwise conv on the tensor core (accumulators alias the depthwise ones) ------------------------------
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
      for (int j = 0; j < P.MT; j++) {
        for (int ks = 0; ks < ksteps; ks++) {
          const int h = ks / ksteps_per_half, kk = ks - h * ksteps_per_half;
          const uint64_t ad = make_desc(a_addr + j * (128 * KP) + h * (128 * RW) + kk * 32, sbo, lt);
          const uint64_t bd = make_desc(b_addr + h * (N * RW) + kk * 32, sbo, lt);
          umma_i8(tmem_base + (uint32_t)(j * N), ad, bd, idesc_pw, ks > 0 ? 1u : 0u);
        }
      }
      umma_commit(smem_u32(&mbar[1]));
    }
    mbar_wait(smem_u32(&mbar[1]), (uint32_t)(it & 1));
    tc_fence_after();
    // ---- (5) epilogue 2: requant (+ residual ADD) -> global -------------------------------------------------------------
    const size_t pix0 = ((size_t)b0 * P.oh + oy0) << P.ow_log;
    for (int t = hsel; t < P.MT * NG; t += 2) {
      const int j = t / NG, g = t - j * NG;
      const int m = j * 128 + 32 * q + lane;
      const int bb = m >> P.trow_log;
      const bool ok = (b0 + bb) < Bw;
      int v[16];
      tmem_ld16(tmem_base + (uint32_t)(j * N + 16 * g) + ((uint32_t)(32 * q) << 16), v);
      uint4 rv = make_uint4(0, 0, 0, 0);
      if (ADD) {
        const int rem = m & ((1 << P.trow_log) - 1);
        const int r = rem >> P.ow_log, ox = rem & (P.ow - 1);
        rv = *reinterpret_cast<const uint4*>(sP + g * plane_bytes + ((bb * TRIN + r + 1) * TW + ox + 1) * 16);
      }
      const unsigned rw[4] = {rv.x ^ 0x80808080u, rv.y ^ 0x80808080u, rv.z ^ 0x80808080u, rv.w ^ 0x80808080u};
      unsigned ow4[4];
#pragma unroll
      for (int gg = 0; gg < 4; gg++) {
        int o[4];
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
          const int c = 16 * g + 4 * gg + jj;
          const int4 rq = s_rq[c];
          if (!ADD) {
            o[jj] = rq_hi(v[4 * gg + jj], rq.x, rq.y, rq.z) >> rq.w;
          } else {
            const int rz = s_rz[c];
            const int y = max(P.pw_lo, min(rq64s(v[4 * gg + jj], rq.x, rq.y, rq.z, rz, rq.w), P.pw_hi));
            const unsigned u = __byte_perm(rw[gg], 0u, 0x4440 + jj);
            const int s1 = (int)(((unsigned long long)u * (unsigned)P.a_m1 + (unsigned long long)P.a_c1) >> P.a_n1);
            int t2;
            if (ADD == 2) {
              t2 = s1 + (y << 19);
            } else {
              int s2 = (int)(((long long)y * (long long)P.a_m2 + P.a_c2) >> 11);
              if (P.a_n2 > 0) s2 = (s2 + P.a_rz2 + (s2 >> 31)) >> P.a_n2;
              t2 = s1 + s2;
            }
            o[jj] = (int)(((long long)t2 * (long long)P.a_mo + P.a_co) >> 32) >> P.a_no;
          }
        }
        ow4[gg] = pack4_sat(o[0], o[1], o[2], o[3]);
      }
      if (ok) *reinterpret_cast<uint4*>(out + (pix0 + m) * N + 16 * g) = make_uint4(ow4[0], ow4[1], ow4[2], ow4[3]);
    }
    tc_fence_before();
    __syncthreads();                                   // TMEM drained, planes and A operand free
  }
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
size_t dst_smem_bytes(const DsParams& P) {
  size_t b = (size_t)P.N * P.KP + (size_t)P.MT * 128 * P.KP + (size_t)(P.C / 32) * 9 * 1024;
  b += (size_t)(P.C / 16) * P.plane_px * 16;
  b += (size_t)P.N * 16 + (size_t)P.C * 16 + (size_t)((P.N + 1) & ~1) * 4 + 32;
  return b + 1024;
}

// diag(w[tap]) blocks in the canonical no-swizzle K-major layout: element (n, k) of block (g, tap) at
// (k / 16) * 512 + (n / 8) * 128 + (n % 8) * 16 + (k % 16); only n == k is non-zero.
void dst_weight_image(const int8_t* w /*[9][C]*/, int C, std::vector<uint8_t>& img) {
  const int G = C / 32;
  img.assign((size_t)G * 9 * 1024, 0);
  for (int g = 0; g < G; g++)
    for (int t = 0; t < 9; t++)
      for (int n = 0; n < 32; n++) {
        const size_t off = (size_t)(g * 9 + t) * 1024 + (n / 16) * 512 + (n / 8) * 128 + (n % 8) * 16 + (n % 16);
        img[off] = (uint8_t)w[(size_t)t * C + 32 * g + n];
      }
}

template <int ADD>
static int launch_one_t(const int8_t* in, int8_t* out, int Bw, int ntiles, int grid, size_t smem, const DsParams& P, cudaStream_t st) {
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) cudaFuncSetAttribute(k_dst<ADD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  k_dst<ADD><<<grid, DST_THREADS, smem, st>>>(in, out, Bw, ntiles, P);
  return 0;
}

int launch_dst(const int8_t* in, int8_t* out, int Bw, const DsParams& P, const DsLaunch& L, int num_sms, cudaStream_t st) {
  const int ntiles = P.NB == 1 ? Bw * (P.oh / P.TRr) : (Bw + P.NB - 1) / P.NB;
  int grid = num_sms * L.ctas_per_sm;
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) return 0;
  if (L.add_mode == 0) return launch_one_t<0>(in, out, Bw, ntiles, grid, L.smem, P, st);
  if (L.add_mode == 1) return launch_one_t<1>(in, out, Bw, ntiles, grid, L.smem, P, st);
  return launch_one_t<2>(in, out, Bw, ntiles, grid, L.smem, P, st);
}

}  // namespace bn
