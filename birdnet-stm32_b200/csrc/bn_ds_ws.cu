// bn_ds_ws.cu -- K45w: the depthwise-separable block of bn_ds.cu as a WARP-SPECIALISED pipeline.
//
// Same block, same integer arithmetic and the same shared-memory operand layouts as k_ds (reference counterpart:
// ds_conv_block, birdnet_stm32/models/dscnn.py:28-84, as lowered into DEPTHWISE_CONV_2D -> CONV_2D 1x1 [-> ADD]), but the
// four phases of a tile no longer run one after the other in every warp with a CTA barrier in between.  A CTA has three
// warpgroups with their own register budgets (setmaxnreg):
//
//   warpgroup 0 (4 warps, 128 registers)  depthwise 3x3 of tile j -> A operand j & 1; its warp 0 also issues the TMA bulk
//                                         copies of the input rows (one cp.async.bulk per row, completion on an mbarrier)
//                                         and the tcgen05.mma of tile j
//   warpgroups 1-2 (8 warps, 56 registers) epilogue of tile j - 1 / j - 2: TMEM -> requantise (+ residual ADD from the input
//                                         tile, still in shared memory) -> global
//
// and the hand-offs are mbarriers, never a CTA-wide barrier:
//
//   full[s]  (tx bytes)     input tile buffer s = j % NST has landed                       TMA        -> depthwise, epilogue
//   mma[a]   (tcgen05.commit) accumulator a = j & 1 is complete, A operand a is free        tensor core -> epilogue, depthwise
//   done[s]  (8 arrivals)   epilogue of tile j finished: tile buffer s and TMEM a are free  epilogue   -> warp 0 of the depthwise group
//
// Two 384-thread CTAs per SM hold 24 warps where two k_ds CTAs hold 16 (the register file is the limit in both cases:
// 2 x (128 x 128 + 256 x 56) = 2 x 30,720 registers), the depthwise warps never wait for a tile load or an MMA round trip of
// their own tile, and the staging costs no instructions in the compute warps.
#include "bn_ds.cuh"

#include "bn_common.cuh"
#include "bn_tc.cuh"

namespace bn {

namespace {

constexpr int WS_THREADS = 384;
constexpr int DW_T = 128;                 // depthwise warpgroup
constexpr int EPI_WARPS = 8;
constexpr int REG_DW = 96, REG_EPI = 64;  // 128 * 96 + 256 * 64 = 28,672 <= 384 * 80 (the depthwise compiles to 87 registers)

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void dw_group_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ int rq64w(int acc, int c_lo, int c_hi, int mult, int rz, int n) {
  const long long c = ((long long)c_hi << 32) | (unsigned)c_lo;
  const long long p = (long long)acc * (long long)mult + c;
  const int v = (int)(p >> 31);
  return (v + rz + (v >> 31)) >> n;
}

}  // namespace

template <int S, int TR, int ADD>
__global__ void __launch_bounds__(WS_THREADS, 2)
k_dsw(const int8_t* __restrict__ in, int8_t* __restrict__ out, int Bw, int ntiles, DsParams P) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int TRIN = (TR - 1) * S + 3;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = P.C, CG = C >> 2, N = P.N, KP = P.KP, RW = P.RW;
  const int TW = P.iw + P.pl + 1;
  const int b_bytes = N * KP, a_bytes = P.MT * 128 * KP;
  const int tile_bytes = (P.NB * TRIN * TW * C + 15) & ~15;
  const int NST = P.nst;                              // 3 or 4 input-tile buffers
  unsigned char* sB = smem;
  unsigned char* sA0 = sB + b_bytes;
  unsigned char* sT0 = sA0 + 2 * a_bytes;
  int4* s_rq = reinterpret_cast<int4*>(sT0 + NST * tile_bytes);
  int* s_rz = reinterpret_cast<int*>(s_rq + N);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(s_rz + ((N + 1) & ~1));
  uint64_t* bar_done = bar_full + 4;
  uint64_t* bar_mma = bar_done + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 2);

  // ---- one-time setup (all warps) ---------------------------------------------------------------------------------
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), (uint32_t)P.tmem_cols);
  if (tid == 32) {
    for (int s = 0; s < 4; s++) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_done[s]), EPI_WARPS); }
    mbar_init(smem_u32(&bar_mma[0]), 1);
    mbar_init(smem_u32(&bar_mma[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < b_bytes / 16; i += WS_THREADS) cp_async16(smem_u32(sB + 16 * i), P.w_img + 16 * (size_t)i);
  cp_async_commit();
  for (int i = tid; i < N; i += WS_THREADS) { s_rq[i] = __ldg(P.pw_rq + i); s_rz[i] = __ldg(P.pw_rz + i); }
  if (P.C < KP) for (int i = tid; i < 2 * a_bytes / 16; i += WS_THREADS) *reinterpret_cast<uint4*>(sA0 + 16 * i) = make_uint4(0, 0, 0, 0);
  const unsigned zpw = 0x01010101u * (unsigned)(uint8_t)P.dw_in_zp;
  {  // halo columns hold the zero point for the whole kernel (the bulk copies never touch them)
    const int rows = P.NB * TRIN;
    const int wpc = C >> 2;
    for (int sb = 0; sb < NST; sb++) {
      unsigned* tw = reinterpret_cast<unsigned*>(sT0 + sb * tile_bytes);
      for (int i = tid; i < rows * wpc * 2; i += WS_THREADS) {
        const int side = i & 1, rest = i >> 1;
        const int row = rest / wpc, w = rest - row * wpc;
        if (side == 0 && P.pl == 0) continue;
        const int col = side ? TW - 1 : 0;
        tw[(row * TW + col) * wpc + w] = zpw;
      }
    }
  }
  cp_async_wait_all();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_chunk = P.oh / TR;
  const int nk = blockIdx.x < ntiles ? (ntiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const int accw = P.MT * N;                          // TMEM columns per accumulator buffer
  auto tile_of = [&](int k) { return blockIdx.x + k * gridDim.x; };
  auto tile_origin = [&](int tile, int& b0, int& oy0) {
    if (P.NB == 1) { b0 = tile / tiles_per_chunk; oy0 = (tile - b0 * tiles_per_chunk) * TR; }
    else { b0 = tile * P.NB; oy0 = 0; }
  };

  if (warp < 4) {
    // =================================================================================================================
    // depthwise warpgroup
    // =================================================================================================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_DW));
    const int cg = tid & (CG - 1);
    int4 wl[3], wr[S == 1 ? 3 : 1];
#pragma unroll
    for (int ky = 0; ky < 3; ky++) {
      wl[ky] = __ldg(P.dw_wt + (2 * ky) * CG + cg);
      if (S == 1) wr[ky] = __ldg(P.dw_wt + (2 * ky + 1) * CG + cg);
    }
    int4 drq[4];
#pragma unroll
    for (int j = 0; j < 4; j++) drq[j] = __ldg(P.dw_rq + 4 * cg + j);
    const int k0 = 4 * cg;
    const int a_kh_off = (k0 >> P.rw_log) * (128 * RW);
    const int a_cc = (k0 & (RW - 1)) >> 4;
    const int a_b = k0 & 15;
    const int a_jhop = 128 * (KP - RW);
    const unsigned a_pairx = P.sw_sh == 0 ? 16u : 0u;
    const int nstrips = P.NB * P.ow * CG;
    const int ppr = (P.iw * C) >> 4;
    const uint32_t row_bytes = (uint32_t)(P.iw * C);

    // input rows of a tile: one TMA bulk copy per row, issued by the lanes of warp 0; SAME-padding rows (and the rows of
    // chunks past the end of the wave) are filled with the zero point by ordinary stores
    auto stage = [&](int tile, unsigned char* sT, uint64_t* bar) {
      int b0, oy0;
      tile_origin(tile, b0, oy0);
      const int rows = P.NB * TRIN;
      const bool mine = lane < rows;
      const int bb = lane / TRIN, tr = lane - bb * TRIN;
      const int iy = oy0 * S - P.pt + tr;
      const bool ok = mine && (b0 + bb) < Bw && iy >= 0 && iy < P.ih;
      const unsigned okmask = __ballot_sync(0xffffffffu, ok);
      unsigned padmask = __ballot_sync(0xffffffffu, mine && !ok);
      if (lane == 0) mbar_expect_tx(smem_u32(bar), (uint32_t)__popc(okmask) * row_bytes);
      __syncwarp();
      if (ok) {
        fence_proxy_async();
        bulk_g2s(smem_u32(sT + ((size_t)lane * TW + P.pl) * C), in + (((size_t)(b0 + bb) * P.ih + iy) * P.iw) * C, row_bytes, smem_u32(bar));
      }
      while (padmask) {
        const int row = __ffs(padmask) - 1;
        padmask &= padmask - 1;
        unsigned char* dst = sT + ((size_t)row * TW + P.pl) * C;
        for (int p = lane; p < ppr; p += 32) *reinterpret_cast<uint4*>(dst + 16 * p) = make_uint4(zpw, zpw, zpw, zpw);
      }
    };

    // depthwise 3x3 with the taps of a filter row along the dp4a axis (see depthwise_t in bn_ds.cu)
    auto depthwise_t = [&](const unsigned char* sT, unsigned char* sA) {
      constexpr int NCOL = S == 1 ? 2 : 1;
      const int owp_log = P.ow_log - (S == 1 ? 1 : 0);
      const int nst = nstrips >> (S == 1 ? 1 : 0);
      for (int sidx = tid; sidx < nst; sidx += DW_T) {
        const int rest = sidx >> P.cg_log;
        const int oxp = rest & ((1 << owp_log) - 1), bb = rest >> owp_log;
        const int ox = oxp * NCOL;
        const unsigned* tp = reinterpret_cast<const unsigned*>(sT) + ((size_t)(bb * TRIN) * TW + ox * S) * CG + cg;
        const int mbase = ((bb * TR) << P.ow_log) + ox;
        const int a_thr = a_kh_off + (((a_cc ^ ((mbase >> P.sw_sh) & P.sw_mask)) << 4) | a_b);
        auto load_row = [&](int ir, unsigned (&t)[4]) {
          const unsigned* rp = tp + (size_t)ir * TW * CG;
          const unsigned x0 = rp[0], x1 = rp[CG], x2 = rp[2 * CG];
          const unsigned a = __byte_perm(x0, x1, 0x5140), b = __byte_perm(x0, x1, 0x7362);
          if (S == 1) {
            const unsigned x3 = rp[3 * CG];
            const unsigned c = __byte_perm(x2, x3, 0x5140), d = __byte_perm(x2, x3, 0x7362);
            t[0] = __byte_perm(a, c, 0x5410); t[1] = __byte_perm(a, c, 0x7632);
            t[2] = __byte_perm(b, d, 0x5410); t[3] = __byte_perm(b, d, 0x7632);
          } else {
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
          *reinterpret_cast<unsigned*>(sA + off) = pack4_sat(rq_hi(aL[0], drq[0].x, drq[0].y, drq[0].z) >> drq[0].w, rq_hi(aL[1], drq[1].x, drq[1].y, drq[1].z) >> drq[1].w,
                                                             rq_hi(aL[2], drq[2].x, drq[2].y, drq[2].z) >> drq[2].w, rq_hi(aL[3], drq[3].x, drq[3].y, drq[3].z) >> drq[3].w);
          if (S == 1)
            *reinterpret_cast<unsigned*>(sA + ((off + RW) ^ a_pairx)) = pack4_sat(rq_hi(aR[0], drq[0].x, drq[0].y, drq[0].z) >> drq[0].w, rq_hi(aR[1], drq[1].x, drq[1].y, drq[1].z) >> drq[1].w,
                                                                                  rq_hi(aR[2], drq[2].x, drq[2].y, drq[2].z) >> drq[2].w, rq_hi(aR[3], drq[3].x, drq[3].y, drq[3].z) >> drq[3].w);
#pragma unroll
          for (int c = 0; c < 4; c++) {
            if (S == 1) { t0[c] = t1[c]; t1[c] = t2[c]; }
            else t0[c] = t2[c];
          }
        }
      }
    };

    const uint32_t sbo = 8 * RW;
    const uint32_t lt = RW == 128 ? 2u : (RW == 64 ? 4u : 6u);
    const uint32_t idesc = make_idesc_i8(128, N);
    const int ksteps = KP >> 5, ksteps_per_half = RW >> 5;
    auto issue_mma = [&](const unsigned char* sA, uint32_t tmem_d, uint64_t* bar) {
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

    if (warp == 0) {
      if (nk > 0) stage(tile_of(0), sT0, &bar_full[0]);
      if (NST == 4 && nk > 1) stage(tile_of(1), sT0 + tile_bytes, &bar_full[1]);
    }
    for (int j = 0; j < nk; j++) {
      const int s = j % NST, a = j & 1;
      if (NST == 3 && warp == 0 && j + 1 < nk) {       // buffer (j + 1) % 3 held tile j - 2
        if (j >= 2) mbar_wait(smem_u32(&bar_done[(j - 2) % NST]), (uint32_t)(((j - 2) / NST) & 1));
        stage(tile_of(j + 1), sT0 + ((j + 1) % NST) * tile_bytes, &bar_full[(j + 1) % NST]);
      }
      mbar_wait(smem_u32(&bar_full[s]), (uint32_t)((j / NST) & 1));
      if (j >= 2) mbar_wait(smem_u32(&bar_mma[a]), (uint32_t)(((j - 2) >> 1) & 1));   // MMA of tile j - 2 has read A operand a
      depthwise_t(sT0 + s * tile_bytes, sA0 + a * a_bytes);
      fence_proxy_async();
      dw_group_sync();
      if (warp == 0) {
        if (j >= 2) mbar_wait(smem_u32(&bar_done[(j - 2) % NST]), (uint32_t)(((j - 2) / NST) & 1));   // accumulator a drained
        tc_fence_after();
        if (lane == 0) issue_mma(sA0 + a * a_bytes, tmem_base + (uint32_t)(a * accw), &bar_mma[a]);
        __syncwarp();
        if (NST == 4 && j + 2 < nk)                      // buffer (j + 2) % 4 held tile j - 2
          stage(tile_of(j + 2), sT0 + ((j + 2) % NST) * tile_bytes, &bar_full[(j + 2) % NST]);
      }
    }
  } else {
    // =================================================================================================================
    // epilogue warpgroups
    // =================================================================================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_EPI));
    const int q = warp & 3, hsel = (warp >> 2) - 1;
    const int NG = N >> 4;
    for (int k = 0; k < nk; k++) {
      const int s = k % NST, a = k & 1;
      const unsigned char* sT = sT0 + s * tile_bytes;
      const uint32_t tmem_d = tmem_base + (uint32_t)(a * accw);
      int b0, oy0;
      tile_origin(tile_of(k), b0, oy0);
      const size_t pix0 = ((size_t)b0 * P.oh + oy0) << P.ow_log;
      if (ADD) mbar_wait(smem_u32(&bar_full[s]), (uint32_t)((k / NST) & 1));
      mbar_wait(smem_u32(&bar_mma[a]), (uint32_t)((k >> 1) & 1));
      tc_fence_after();
      for (int t = hsel; t < P.MT * NG; t += EPI_WARPS / 4) {
        const int j = t / NG, g = t - j * NG;
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
              const int y = max(P.pw_lo, min(rq64w(v[4 * gg + jj], rq.x, rq.y, rq.z, rz, rq.w), P.pw_hi));
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
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_done[s]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
size_t dsw_smem_bytes(const DsParams& P, int S, int TR) {
  const int trin = (TR - 1) * S + 3, tw = P.iw + P.pl + 1;
  size_t b = (size_t)P.N * P.KP + (size_t)2 * P.MT * 128 * P.KP;
  b += (size_t)P.nst * (((size_t)P.NB * trin * tw * P.C + 15) & ~(size_t)15);
  b += (size_t)P.N * 16 + (size_t)((P.N + 1) & ~1) * 4 + 10 * 8 + 16;
  return b + 1024;
}

bool dsw_supported(const DsParams& P, int S, int TR, int add_mode) {
  const int trin = (TR - 1) * S + 3;
  if (P.nst != 3 && P.nst != 4) return false;
  if (P.NB * trin > 32) return false;                 // one lane of warp 0 per input row
  if ((P.iw * P.C) % 16 || P.C % 16) return false;    // bulk copies: 16-byte sizes and addresses
  if (P.ow % 8) return false;
  if (S == 1) return (TR == 4 || TR == 8) && add_mode >= 0 && add_mode <= 2;
  return S == 2 && (TR == 4 || TR == 8) && add_mode == 0;
}

template <int S, int TR, int ADD>
static int launch_w(const int8_t* in, int8_t* out, int Bw, int ntiles, int grid, size_t smem, const DsParams& P, cudaStream_t st) {
  static unsigned long long attr = 0;
  if (first_use_on_device(attr)) cudaFuncSetAttribute(k_dsw<S, TR, ADD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  k_dsw<S, TR, ADD><<<grid, WS_THREADS, smem, st>>>(in, out, Bw, ntiles, P);
  return 0;
}

int launch_dsw(const int8_t* in, int8_t* out, int Bw, const DsParams& P, const DsLaunch& L, int num_sms, cudaStream_t st) {
  const int ntiles = P.NB == 1 ? Bw * (P.oh / L.TR) : (Bw + P.NB - 1) / P.NB;
  int grid = num_sms * L.ctas_per_sm;
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) return 0;
#define W_CASE(s, tr, add) \
  if (L.S == s && L.TR == tr && L.add_mode == add) return launch_w<s, tr, add>(in, out, Bw, ntiles, grid, L.smem, P, st)
  W_CASE(1, 4, 0); W_CASE(1, 4, 1); W_CASE(1, 4, 2);
  W_CASE(1, 8, 0); W_CASE(1, 8, 1); W_CASE(1, 8, 2);
  W_CASE(2, 4, 0); W_CASE(2, 8, 0);
#undef W_CASE
  return BN_ERR_UNSUPPORTED;
}

}  // namespace bn
