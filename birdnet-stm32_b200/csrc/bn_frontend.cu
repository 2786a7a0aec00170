// bn_frontend.cu -- K1: batched STFT magnitude of PCM16 chunks (hybrid frontend).
//
// Replaces the reference's host frontend for `hybrid` models:
//   get_spectrogram_from_audio(mel_bins=-1): librosa.stft(n_fft=512, hop=len//W, hann, center=True,
//   pad_mode="constant") -> np.abs -> [:, :W]      birdnet_stm32/audio/spectrogram.py:61,106-115,133
// preceded by the PCM16 -> float32 decode and file-peak division of audio/io.py:114-126.
//
// Layout: one CTA = 32 consecutive frames of one chunk (grid = W/32 x B).  The CTA stages the
// PCM span it needs into shared memory as float32 (coalesced 32-bit loads of int16 pairs, decode
// and peak division done once per sample), each half-warp computes one 512-point real FFT as a
// 256-point complex FFT (two register-resident radix-16 passes with a shared-memory transpose
// between them, twiddles from a shared table), the real-FFT split and |.| are applied, and the
// 257 x 32 magnitude tile is written bin-major with coalesced 128-byte rows.  Per-chunk min and
// max (needed by normalize()) are reduced per CTA and merged with integer atomics
// (magnitudes are >= +0, so their IEEE bit patterns order like unsigned integers).
#include "bn_common.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "bn_fft.cuh"
#include "bn_kernels.cuh"

namespace bn {

// dynamic smem layout:
//   float  xs[span + 16]                    span = 31*hop + 512 floats (16-byte aligned, indexed like the global buffer)
//   uint4  sraw[(span + 16) / G]            raw samples of the NEXT tile (cp.async prefetch), G = 8 int16 or 4 float32 per 16 bytes
//   float2 zbuf[16][256+16]                 per half-warp exchange buffer (padded)
//   float  tile[257][33]                    magnitude tile (bin-major variant only)
//   float  red[2*8]
//
// The CTA is persistent over (chunk, 32-frame group) tiles.  Everything a thread needs that does not depend on the
// data -- its 32 window samples (pre-scaled by 1/2 for the real-FFT split), the 15 inter-pass twiddles W256^(l k1) and
// the 8 split twiddles W512^(l + 16 j) -- is loaded once into registers and reused for every frame.
//
// F32IN: the chunks are float32 waveforms (what the reference's make_chunks_for_file hands to the frontend after
// load_audio_window resampled / mixed the file, audio/io.py:118-128); the scale is then 1 / peak, or exactly 1 without a peak.
template <bool FRAME_MAJOR, bool F32IN>
__global__ void __launch_bounds__(FE_THREADS, 2)
k_stft_mag(const void* __restrict__ pcm_v, const float* __restrict__ peak, float* __restrict__ out,
           unsigned* __restrict__ mnmx, const float4* __restrict__ tables, int T, int hop, int W, int ldk, int B) {
  using raw_t = typename std::conditional<F32IN, float, int16_t>::type;
  constexpr int G = F32IN ? 4 : 8;                    // samples per 16-byte group
  constexpr int GL = F32IN ? 2 : 3;
  const raw_t* pcm = reinterpret_cast<const raw_t*>(pcm_v);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int span = (FRAMES_PER_CTA - 1) * hop + NFFT;
  float* xs = reinterpret_cast<float*>(smem_raw);
  uint4* sraw = reinterpret_cast<uint4*>(xs + ((span + 16 + 3) & ~3));      // (span + 16) / G groups of G samples
  float2* zbuf = reinterpret_cast<float2*>(sraw + ((span + 16 + G - 1) >> GL));
  float* tile = reinterpret_cast<float*>(zbuf + 16 * (NC + 16));
  float* red = FRAME_MAJOR ? tile : tile + BINS * TILE_LD;

  const int tid = threadIdx.x;
  const int hw = tid >> 4;        // half-warp id 0..15
  const int l = tid & 15;         // lane within the half-warp = n2 (pass 1) / k1 (pass 2)
  float2* zb = zbuf + hw * (NC + 16);
  const unsigned hmask = 0xffffu << (16 * ((tid >> 4) & 1));

  // per-thread constants (tables: tw512[512] float2 followed by win[512] float, computed on the host in double)
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

  // the sample pointer is only known to be element aligned: index samples relative to its 16-byte-aligned floor
  const int a0 = (int)((reinterpret_cast<uintptr_t>(pcm) / sizeof(raw_t)) & (G - 1));
  const raw_t* pcm_al = pcm - a0;
  const long total = (long)B * T;
  const int groups_w = W / FRAMES_PER_CTA;
  const int ntiles = B * groups_w;

  // Raw int16 samples of a tile arrive through cp.async one tile ahead (group grp of 8 samples is always handled by
  // thread grp % FE_THREADS, so a thread only ever touches its own slots of sraw and no barrier is needed for them).
  // groups_w is a power of two for every shipped configuration (W = 256): shift instead of an integer division per tile
  const int gw_log = (groups_w & (groups_w - 1)) == 0 ? 31 - __clz(groups_w) : -1;
  auto tile_geom = [&](int tile_id, int& b, int& t0, long& chunk_base, long& g_first, int& shift, int& ngroups) {
    b = gw_log >= 0 ? (tile_id >> gw_log) : tile_id / groups_w;
    t0 = (tile_id - b * groups_w) * FRAMES_PER_CTA;
    chunk_base = (long)b * T + a0;                        // in pcm_al sample indices
    const long g_lo = chunk_base + (long)t0 * hop - NFFT / 2;
    g_first = g_lo & ~(long)(G - 1);
    shift = (int)(g_lo - g_first);
    ngroups = (shift + span + G - 1) >> GL;
  };
  auto prefetch = [&](int tile_id) {
    int b, t0, shift, ngroups; long chunk_base, g_first;
    tile_geom(tile_id, b, t0, chunk_base, g_first, shift, ngroups);
    if (g_first >= a0 && g_first + (long)G * ngroups <= a0 + total) {
      // the whole span lies inside the buffer (every tile but the first / last of the wave): no per-group range checks
      const raw_t* src = pcm_al + g_first;
      for (int grp = tid; grp < ngroups; grp += FE_THREADS) cp_async16_fe(sraw + grp, src + G * grp);
      return;
    }
    for (int grp = tid; grp < ngroups; grp += FE_THREADS) {
      const long g = g_first + (long)G * grp;
      if (g >= a0 && g + G <= a0 + total) {
        cp_async16_fe(sraw + grp, pcm_al + g);
      } else if (F32IN) {                                 // first / last samples of the whole buffer
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

  if ((int)blockIdx.x < ntiles) prefetch(blockIdx.x);
  asm volatile("cp.async.commit_group;" ::: "memory");

  for (int tile_id = blockIdx.x; tile_id < ntiles; tile_id += gridDim.x) {
    int b, t0, shift, ngroups; long chunk_base, g_first;
    tile_geom(tile_id, b, t0, chunk_base, g_first, shift, ngroups);
    // ---- samples [s0, s0 + span) of chunk b as float32, zero outside [0, T) -------------------------------------
    const float pk = peak ? __ldg(peak + b) : 0.0f;
    // y = (s / 32768) / peak (audio/io.py:114-126) as one multiply by the rounded reciprocal (<= 1.5 ulp from the reference)
    const float cs = F32IN ? (pk > 0.0f ? __fdiv_rn(1.0f, pk) : 1.0f) : (pk > 0.0f ? __fdiv_rn(1.0f, 32768.0f * pk) : (1.0f / 32768.0f));
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const int rel0 = (int)(g_first - chunk_base);           // chunk-relative index of the span's first group (|rel0| < 2^31)
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
      const long rel = g_first + (long)G * grp - chunk_base;   // chunk-relative index of the group's first sample
      const uint4 wv = sraw[grp];
      if (F32IN) {
        float f4[4] = {__uint_as_float(wv.x), __uint_as_float(wv.y), __uint_as_float(wv.z), __uint_as_float(wv.w)};
#pragma unroll
        for (int j = 0; j < 4; j++) f4[j] = (rel + j >= 0 && rel + j < T) ? f4[j] * cs : 0.0f;   // zero padding outside [0, T)
        *reinterpret_cast<float4*>(xs + 4 * grp) = make_float4(f4[0], f4[1], f4[2], f4[3]);
        continue;
      }
      unsigned ww[4] = {wv.x, wv.y, wv.z, wv.w};
      if (rel < 0 || rel + 8 > T) {                       // chunk edge: samples outside [0, T) are zero padding
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
        // int16 -> float32 through the 1.5 * 2^23 magic constant (exact), then the scale
        const int lo = (int)(short)(ww[j] & 0xffffu), hi = (int)ww[j] >> 16;
        fv[2 * j] = (__int_as_float(0x4B400000 + lo) - 12582912.0f) * cs;
        fv[2 * j + 1] = (__int_as_float(0x4B400000 + hi) - 12582912.0f) * cs;
      }
      *reinterpret_cast<float4*>(xs + 8 * grp) = make_float4(fv[0], fv[1], fv[2], fv[3]);
      *reinterpret_cast<float4*>(xs + 8 * grp + 4) = make_float4(fv[4], fv[5], fv[6], fv[7]);
    }
    if (tile_id + (int)gridDim.x < ntiles) prefetch(tile_id + gridDim.x);
    asm volatile("cp.async.commit_group;" ::: "memory");
    __syncthreads();

    float lmin = __int_as_float(0x7f800000), lmax = 0.0f;
#pragma unroll 1
    for (int round = 0; round < FRAMES_PER_CTA / 16; round++) {
      const int f = round * 16 + hw;                      // frame within the tile
      const float* xf = xs + shift + f * hop + 2 * l;     // sample 2 n of this frame for n = l
      // pass 1: thread n2 = l takes z[16*n1 + l], n1 = 0..15, z[n] = x[2n] + i x[2n+1] (windowed, pre-scaled by 1/2)
      float2 v[16];
#pragma unroll
      for (int n1 = 0; n1 < 16; n1++) v[n1] = make_float2(xf[32 * n1] * wre[n1], xf[32 * n1 + 1] * wim[n1]);
      fft16(v);                                           // over n1 -> k1
      // twiddle W256^(l*k1), then transpose through smem
      zb[l] = v[0];
#pragma unroll
      for (int k1 = 1; k1 < 16; k1++) zb[k1 * 17 + l] = cmul(v[k1], twp[k1]);
      __syncwarp(hmask);
      // pass 2: thread k1 = l reads B[n2] = zb[l][n2]
#pragma unroll
      for (int n2 = 0; n2 < 16; n2++) v[n2] = zb[l * 17 + n2];
      __syncwarp(hmask);
      fft16(v);                                           // over n2 -> k2 ; Z[k1 + 16 k2] = v[k2]
      // real-FFT split, two bins per step: with E = Z[k] + conj(Z[N-k]), O = (Z[k] - conj(Z[N-k])) / i (Z carries the
      // factor 1/2) and T = W512^k O:   X[k] = E + T   and   X[N-k] = conj(E - T)   (N = 256), so |X[N-k]| = |E - T|.
      // Z[k] is the thread's own v[j]; its partner Z[N-k] comes from lane 16 - l by register shuffle (split_partner): the
      // second shared-memory exchange (16 stores + 17 loads of 8 bytes per thread) is gone -- K1 runs at 70 % of the
      // shared-memory data pipe, its tightest resource.
      float* orow = FRAME_MAJOR ? out + ((long)b * W + t0 + f) * ldk : tile + f;
      float* const oa = orow + l;                         // bins l + 16 j and 256 - l - 16 j: per-thread bases, compile-time offsets
      float* const ob = orow + NC - l;
      auto split_step = [&](const float2 zk, const float2 zn, const int j) {
        const int k = l + 16 * j;                         // 0..127, partner bin 256 - k
        const float2 e = make_float2(zk.x + zn.x, zk.y - zn.y);
        const float2 o = make_float2(zk.y + zn.y, zn.x - zk.x);
        const float2 t = cmul(o, tws[j]);
        const float2 xa = add2(e, t), xb = sub2(e, t);
        const float ar = xa.x, ai = xa.y, br = xb.x, bi = xb.y;
        const float ma = fast_sqrt(ar * ar + ai * ai), mb = fast_sqrt(br * br + bi * bi);
        if (FRAME_MAJOR) { oa[16 * j] = ma; ob[-16 * j] = mb; }   // 16 lanes -> 64 contiguous bytes each
        else { orow[k * TILE_LD] = ma; orow[(NC - k) * TILE_LD] = mb; }
        lmin = fminf(lmin, fminf(ma, mb));
        lmax = fmaxf(lmax, fmaxf(ma, mb));
      };
      const int lw = tid & 31;
      split_step(v[0], split_partner<0>(v, l, lw), 0); split_step(v[1], split_partner<1>(v, l, lw), 1);
      split_step(v[2], split_partner<2>(v, l, lw), 2); split_step(v[3], split_partner<3>(v, l, lw), 3);
      split_step(v[4], split_partner<4>(v, l, lw), 4); split_step(v[5], split_partner<5>(v, l, lw), 5);
      split_step(v[6], split_partner<6>(v, l, lw), 6); split_step(v[7], split_partner<7>(v, l, lw), 7);
      if (l == 0) {                                       // bin 128 = Z[0 + 16 * 8] pairs with itself: |X[128]| = 2 |Z[128]|
        const float2 z = v[8];
        const float m = 2.0f * fast_sqrt(z.x * z.x + z.y * z.y);
        if (FRAME_MAJOR) orow[128] = m; else orow[128 * TILE_LD] = m;
        lmin = fminf(lmin, m);
        lmax = fmaxf(lmax, m);
      }
      __syncwarp(hmask);                                  // zb is reused by the next round's first exchange
    }

    // CTA reduction of min / max
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
      lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    }
    if ((tid & 31) == 0) { red[tid >> 5] = lmin; red[8 + (tid >> 5)] = lmax; }
    __syncthreads();                                      // also: every warp is done with xs / zbuf / tile writes
    if (tid == 0) {
      float mn = red[0], mx = red[8];
      for (int i = 1; i < 8; i++) { mn = fminf(mn, red[i]); mx = fmaxf(mx, red[8 + i]); }
      atomicMin(mnmx + 2 * b, __float_as_uint(mn));
      atomicMax(mnmx + 2 * b + 1, __float_as_uint(mx));
    }
    if (!FRAME_MAJOR) {
      // write the tile bin-major: out[b][k][t0 + f], 32 consecutive floats per bin row
      float* ob = out + (long)b * BINS * W;
      const int lane = tid & 31, wp = tid >> 5;
      for (int k = wp; k < BINS; k += FE_THREADS / 32) ob[(long)k * W + t0 + lane] = tile[k * TILE_LD + lane];
    }
    __syncthreads();                                      // red / tile are free for the next tile
  }
}

__global__ void k_init_minmax(unsigned* mnmx, int B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) { mnmx[2 * i] = 0x7f800000u; mnmx[2 * i + 1] = 0u; }
}

// tw512[m] = exp(-2 pi i m / 512) (float2) followed by the periodic Hann window win[m] = 0.5 - 0.5 cos(2 pi m / 512)
static const float4* stft_tables() {
  static float4* d_tabs[64] = {nullptr};                // one copy per device ordinal
  int dev = 0;
  cudaGetDevice(&dev);
  float4*& d_tab = d_tabs[dev & 63];
  if (d_tab) return d_tab;
  float h[NFFT * 3];
  const double PI = 3.14159265358979323846;
  for (int m = 0; m < NFFT; m++) {
    h[2 * m] = (float)cos(2.0 * PI * m / NFFT);
    h[2 * m + 1] = (float)(-sin(2.0 * PI * m / NFFT));
    h[2 * NFFT + m] = (float)(0.5 - 0.5 * cos(2.0 * PI * m / NFFT));
  }
  if (cudaMalloc(&d_tab, sizeof h) != cudaSuccess) { d_tab = nullptr; return nullptr; }
  cudaMemcpy(d_tab, h, sizeof h, cudaMemcpyHostToDevice);
  return d_tab;
}

const float4* stft_tables_shared() { return stft_tables(); }

size_t stft_smem_bytes(int hop, bool frame_major, bool f32) {
  const int span = (FRAMES_PER_CTA - 1) * hop + NFFT;
  size_t b = sizeof(float) * ((span + 16 + 3) & ~3) + 16 * (size_t)(f32 ? ((span + 16 + 3) >> 2) : ((span + 16 + 7) >> 3));
  b += sizeof(float2) * 16 * (NC + 16) + sizeof(float) * 16;
  b += frame_major ? 0 : sizeof(float) * BINS * TILE_LD;
  return b;
}

// persistent grid: as many CTAs as fit (2 per SM by registers, fewer if the staged span is large)
static int stft_grid(int B, int W, size_t smem) {
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
  int per = (int)((227 * 1024) / (smem + 1024));
  if (per > 2) per = 2;
  if (per < 1) per = 1;
  const int ntiles = B * (W / FRAMES_PER_CTA);
  const int g = sms * per;
  return g < ntiles ? g : ntiles;
}

static void set_stft_attrs() {
  cudaFuncSetAttribute(k_stft_mag<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(k_stft_mag<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(k_stft_mag<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaFuncSetAttribute(k_stft_mag<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

int launch_stft_mag(const void* pcm, int f32, const float* peak, float* out, unsigned* mnmx, int B, int T, int n_fft,
                    int hop, int W, cudaStream_t st) {
  if (n_fft != NFFT) return BN_ERR_UNSUPPORTED;
  const size_t smem = stft_smem_bytes(hop, false, f32 != 0);
  if (smem > 227 * 1024) return BN_ERR_UNSUPPORTED;
  static unsigned long long attr_done = 0;
  if (first_use_on_device(attr_done)) set_stft_attrs();
  const float4* tab = stft_tables();
  if (!tab) return BN_ERR_CUDA;
  if (W % FRAMES_PER_CTA) return BN_ERR_UNSUPPORTED;
  k_init_minmax<<<(B + 255) / 256, 256, 0, st>>>(mnmx, B);
  if (f32) k_stft_mag<false, true><<<stft_grid(B, W, smem), FE_THREADS, smem, st>>>(pcm, peak, out, mnmx, tab, T, hop, W, 0, B);
  else k_stft_mag<false, false><<<stft_grid(B, W, smem), FE_THREADS, smem, st>>>(pcm, peak, out, mnmx, tab, T, hop, W, 0, B);
  return 0;
}

// Frame-major variant for the fused plan: out float32 [B, W, ldk] (raw magnitudes, bins 0..256 of
// each frame contiguous; columns >= 257 are left untouched).
int launch_stft_mag_fm(const void* pcm, int f32, const float* peak, float* out, unsigned* mnmx, int B, int T, int n_fft,
                       int hop, int W, int ldk, cudaStream_t st) {
  if (n_fft != NFFT || ldk < BINS) return BN_ERR_UNSUPPORTED;
  const size_t smem = stft_smem_bytes(hop, true, f32 != 0);
  if (smem > 227 * 1024) return BN_ERR_UNSUPPORTED;
  static unsigned long long attr_done = 0;
  if (first_use_on_device(attr_done)) set_stft_attrs();
  const float4* tab = stft_tables();
  if (!tab) return BN_ERR_CUDA;
  if (W % FRAMES_PER_CTA) return BN_ERR_UNSUPPORTED;
  if (getenv("BN_DEBUG")) {
    int n = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_stft_mag<true, false>, FE_THREADS, smem);
    fprintf(stderr, "launch_stft_mag_fm: occupancy API says %d CTAs per SM for K1 with %zu bytes of shared memory\n", n, smem);
  }
  k_init_minmax<<<(B + 255) / 256, 256, 0, st>>>(mnmx, B);
  if (f32) k_stft_mag<true, true><<<stft_grid(B, W, smem), FE_THREADS, smem, st>>>(pcm, peak, out, mnmx, tab, T, hop, W, ldk, B);
  else k_stft_mag<true, false><<<stft_grid(B, W, smem), FE_THREADS, smem, st>>>(pcm, peak, out, mnmx, tab, T, hop, W, ldk, B);
  return 0;
}

}  // namespace bn
