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

#include "bn_kernels.cuh"

namespace bn {

constexpr int NFFT = 512;
constexpr int NC = 256;          // complex points
constexpr int BINS = 257;
constexpr int FRAMES_PER_CTA = 32;
constexpr int FE_THREADS = 256;  // 8 warps = 16 half-warps = 16 concurrent FFTs, 2 rounds
constexpr int TILE_LD = FRAMES_PER_CTA + 1;

// sqrt.approx.f32: max relative error 2^-23, far inside the 1e-4 frontend tolerance
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// In-register 16-point DIF FFT, output in natural order (bit reversal folded into the
// compile-time unrolled index map).  tw16[j] = exp(-2 pi i j / 16).
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
  const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
  const float2 tw[8] = {{1.f, 0.f}, {c1, -s1}, {r2, -r2}, {s1, -c1}, {0.f, -1.f}, {-s1, -c1}, {-r2, -r2}, {-c1, -s1}};
#pragma unroll
  for (int half = 8; half >= 1; half >>= 1) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if ((i & half) == 0) {
        float2 a = v[i], b = v[i + half];
        v[i] = make_float2(a.x + b.x, a.y + b.y);
        float2 d = make_float2(a.x - b.x, a.y - b.y);
        const int j = (i & (half - 1)) * (8 / half);   // twiddle index into tw (N=16 base)
        if (j == 0) v[i + half] = d;
        else if (j == 4) v[i + half] = make_float2(d.y, -d.x);
        else v[i + half] = cmul(d, tw[j]);
      }
    }
  }
  // bit-reverse permutation (4 bits)
  float2 o[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int r = ((i & 1) << 3) | ((i & 2) << 1) | ((i & 4) >> 1) | ((i & 8) >> 3);
    o[r] = v[i];
  }
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = o[i];
}

// dynamic smem layout:
//   float2 tw512[512]                       4 KB   exp(-2 pi i m / 512)
//   float  win[512]                         2 KB   periodic Hann
//   float  xs[span]                         span = 31*hop + 512 floats
//   float2 zbuf[16][256+16]                 per half-warp exchange buffer (padded)
//   float  tile[257][33]                    magnitude tile
//   float  red[2*8]
template <bool FRAME_MAJOR>
__global__ void __launch_bounds__(FE_THREADS)
k_stft_mag(const int16_t* __restrict__ pcm, const float* __restrict__ peak, float* __restrict__ out,
           unsigned* __restrict__ mnmx, const float4* __restrict__ tables, int T, int hop, int W, int ldk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int span = (FRAMES_PER_CTA - 1) * hop + NFFT;
  float2* tw512 = reinterpret_cast<float2*>(smem_raw);
  float* win = reinterpret_cast<float*>(tw512 + NFFT);
  float* xs = win + NFFT;
  float2* zbuf = reinterpret_cast<float2*>(xs + ((span + 3) & ~3));
  float* tile = reinterpret_cast<float*>(zbuf + 16 * (NC + 16));
  float* red = FRAME_MAJOR ? tile : tile + BINS * TILE_LD;

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * FRAMES_PER_CTA;

  // tables (computed once on the host in double precision): tw512[512] float2 followed by win[512] float
  for (int m = tid; m < (NFFT * 2 + NFFT) / 4; m += FE_THREADS) reinterpret_cast<float4*>(smem_raw)[m] = __ldg(tables + m);

  // stage samples [s0, s0 + span) of chunk b, zero outside [0, T)
  const float pk = peak ? peak[b] : 0.0f;
  const long chunk_base = (long)b * T;
  const long s0 = (long)t0 * hop - NFFT / 2;           // first sample (chunk-relative), may be < 0
  {
    // 32-bit loads over global sample pairs; g = global sample index (even)
    const long g_lo = chunk_base + s0;
    const long g_first = g_lo & ~1L;                    // floor to even (g_lo may be negative only if b == 0)
    const unsigned* p32 = reinterpret_cast<const unsigned*>(pcm);
    const long total = (long)gridDim.y * T;             // samples in the whole buffer
    const int npairs = (int)(((g_lo + span + 1) - g_first + 1) / 2);
    // batches of 8 independent 32-bit loads per thread (memory-level parallelism), then decode
    for (int base = 0; base < npairs; base += FE_THREADS * 8) {
      unsigned wv[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int w = base + u * FE_THREADS + tid;
        const long g = g_first + 2L * w;
        unsigned word = 0;
        if (w < npairs) {
          if (g >= 0 && g + 1 < total) word = __ldg(p32 + (g >> 1));
          else if (g >= 0 && g < total) word = (unsigned)(unsigned short)pcm[g];
        }
        wv[u] = word;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int w = base + u * FE_THREADS + tid;
        if (w >= npairs) continue;
        const long g = g_first + 2L * w;
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const long gs = g + h;
          const long rel = gs - chunk_base;               // chunk-relative sample index
          const long li = gs - g_lo;                      // index into xs
          if (li >= 0 && li < span) {
            float v = 0.0f;
            if (rel >= 0 && rel < T) {
              const short sv = (short)(h ? (wv[u] >> 16) : (wv[u] & 0xffffu));
              v = (float)sv * (1.0f / 32768.0f);          // exact (power of two)
              if (pk > 0.0f) v = __fdiv_rn(v, pk);        // y / peak, float32 (audio/io.py:124-126)
            }
            xs[li] = v;
          }
        }
      }
    }
  }
  __syncthreads();

  const int hw = tid >> 4;        // half-warp id 0..15
  const int l = tid & 15;         // lane within the half-warp = n2 (pass 1) / k1 (pass 2)
  float2* zb = zbuf + hw * (NC + 16);
  const unsigned hmask = 0xffffu << (16 * ((tid >> 4) & 1));

  float lmin = __int_as_float(0x7f800000), lmax = 0.0f;

#pragma unroll 1
  for (int round = 0; round < FRAMES_PER_CTA / 16; round++) {
    const int f = round * 16 + hw;                      // frame within the tile
    const float* xf = xs + f * hop;                     // 512 samples of this frame
    // pass 1: thread n2 = l takes z[16*n1 + l], n1 = 0..15, z[n] = x[2n] + i x[2n+1] (windowed)
    float2 v[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) {
      const int n = 16 * n1 + l;
      const float2 x2 = make_float2(xf[2 * n], xf[2 * n + 1]);
      const float2 w2 = *reinterpret_cast<const float2*>(win + 2 * n);
      v[n1] = make_float2(x2.x * w2.x, x2.y * w2.y);
    }
    fft16(v);                                           // over n1 -> k1
    // twiddle W256^(l*k1) = tw512[2*l*k1], then transpose through smem: zb[k1*17 + n2]... use [k1][n2]
#pragma unroll
    for (int k1 = 0; k1 < 16; k1++) {
      float2 t = v[k1];
      if (k1 != 0 && l != 0) t = cmul(t, tw512[(2 * l * k1) & 511]);
      zb[k1 * 17 + l] = t;
    }
    __syncwarp(hmask);
    // pass 2: thread k1 = l reads B[n2] = zb[l][n2]
#pragma unroll
    for (int n2 = 0; n2 < 16; n2++) v[n2] = zb[l * 17 + n2];
    __syncwarp(hmask);
    fft16(v);                                           // over n2 -> k2 ; Z[k1 + 16 k2] = v[k2]
#pragma unroll
    for (int k2 = 0; k2 < 16; k2++) zb[l + 16 * k2] = v[k2];   // natural order Z[0..255]
    __syncwarp(hmask);
    // real-FFT split, two bins per step: with E = (Z[k] + conj(Z[N-k]))/2, O = (Z[k] - conj(Z[N-k]))/(2i) and
    // T = W512^k O:   X[k] = E + T   and   X[N-k] = conj(E - T)   (N = 256), so |X[N-k]| = |E - T|.
    const bool fok = t0 + f < W;
    auto emit = [&](int k, float mag) {
      if (fok) {
        if (FRAME_MAJOR) out[((long)b * W + t0 + f) * ldk + k] = mag;   // 16 lanes -> 64 contiguous bytes
        else tile[k * TILE_LD + f] = mag;
        lmin = fminf(lmin, mag);
        lmax = fmaxf(lmax, mag);
      }
    };
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int k = l + 16 * j;                       // 0..127, partner bin 256 - k
      const float2 zk = zb[k];
      const float2 zn = zb[(NC - k) & 255];
      const float2 e = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
      const float2 o = make_float2(0.5f * (zk.y + zn.y), -0.5f * (zk.x - zn.x));
      const float2 t = cmul(o, tw512[k]);
      const float ar = e.x + t.x, ai = e.y + t.y, br = e.x - t.x, bi = e.y - t.y;
      emit(k, fast_sqrt(ar * ar + ai * ai));
      emit(NC - k, fast_sqrt(br * br + bi * bi));
    }
    if (l == 0) {                                     // bin 128 pairs with itself: |X[128]| = |Z[128]|
      const float2 z = zb[128];
      emit(128, fast_sqrt(z.x * z.x + z.y * z.y));
    }
    __syncwarp(hmask);
  }

  // CTA reduction of min / max
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
    lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
  }
  if ((tid & 31) == 0) { red[tid >> 5] = lmin; red[8 + (tid >> 5)] = lmax; }
  __syncthreads();
  if (tid == 0) {
    float mn = red[0], mx = red[8];
    for (int i = 1; i < 8; i++) { mn = fminf(mn, red[i]); mx = fmaxf(mx, red[8 + i]); }
    atomicMin(mnmx + 2 * b, __float_as_uint(mn));
    atomicMax(mnmx + 2 * b + 1, __float_as_uint(mx));
  }

  if (FRAME_MAJOR) return;
  // write the tile bin-major: out[b][k][t0 + f], 32 consecutive floats per bin row
  float* ob = out + (long)b * BINS * W;
  const int lane = tid & 31, wp = tid >> 5;
  for (int k = wp; k < BINS; k += FE_THREADS / 32) {
    if (t0 + lane < W) ob[(long)k * W + t0 + lane] = tile[k * TILE_LD + lane];
  }
}

__global__ void k_init_minmax(unsigned* mnmx, int B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) { mnmx[2 * i] = 0x7f800000u; mnmx[2 * i + 1] = 0u; }
}

// tw512[m] = exp(-2 pi i m / 512) (float2) followed by the periodic Hann window win[m] = 0.5 - 0.5 cos(2 pi m / 512)
static const float4* stft_tables() {
  static float4* d_tab = nullptr;
  static int dev_of = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (d_tab && dev_of == dev) return d_tab;
  float h[NFFT * 3];
  const double PI = 3.14159265358979323846;
  for (int m = 0; m < NFFT; m++) {
    h[2 * m] = (float)cos(2.0 * PI * m / NFFT);
    h[2 * m + 1] = (float)(-sin(2.0 * PI * m / NFFT));
    h[2 * NFFT + m] = (float)(0.5 - 0.5 * cos(2.0 * PI * m / NFFT));
  }
  if (cudaMalloc(&d_tab, sizeof h) != cudaSuccess) return nullptr;
  cudaMemcpy(d_tab, h, sizeof h, cudaMemcpyHostToDevice);
  dev_of = dev;
  return d_tab;
}

size_t stft_smem_bytes(int hop, bool frame_major) {
  const int span = (FRAMES_PER_CTA - 1) * hop + NFFT;
  size_t b = sizeof(float2) * NFFT + sizeof(float) * NFFT + sizeof(float) * ((span + 3) & ~3);
  b += sizeof(float2) * 16 * (NC + 16) + sizeof(float) * 16;
  b += frame_major ? 0 : sizeof(float) * BINS * TILE_LD;
  return b;
}

int launch_stft_mag(const int16_t* pcm, const float* peak, float* out, unsigned* mnmx, int B, int T, int n_fft,
                    int hop, int W, cudaStream_t st) {
  if (n_fft != NFFT) return BN_ERR_UNSUPPORTED;
  const size_t smem = stft_smem_bytes(hop, false);
  if (smem > 227 * 1024) return BN_ERR_UNSUPPORTED;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(k_stft_mag<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(k_stft_mag<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_done = true;
  }
  const float4* tab = stft_tables();
  if (!tab) return BN_ERR_CUDA;
  k_init_minmax<<<(B + 255) / 256, 256, 0, st>>>(mnmx, B);
  dim3 grid((W + FRAMES_PER_CTA - 1) / FRAMES_PER_CTA, B);
  k_stft_mag<false><<<grid, FE_THREADS, smem, st>>>(pcm, peak, out, mnmx, tab, T, hop, W, 0);
  return 0;
}

// Frame-major variant for the fused plan: out float32 [B, W, ldk] (raw magnitudes, bins 0..256 of
// each frame contiguous; columns >= 257 are left untouched).
int launch_stft_mag_fm(const int16_t* pcm, const float* peak, float* out, unsigned* mnmx, int B, int T, int n_fft,
                       int hop, int W, int ldk, cudaStream_t st) {
  if (n_fft != NFFT || ldk < BINS) return BN_ERR_UNSUPPORTED;
  const size_t smem = stft_smem_bytes(hop, true);
  if (smem > 227 * 1024) return BN_ERR_UNSUPPORTED;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(k_stft_mag<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_done = true;
  }
  const float4* tab = stft_tables();
  if (!tab) return BN_ERR_CUDA;
  k_init_minmax<<<(B + 255) / 256, 256, 0, st>>>(mnmx, B);
  dim3 grid((W + FRAMES_PER_CTA - 1) / FRAMES_PER_CTA, B);
  k_stft_mag<true><<<grid, FE_THREADS, smem, st>>>(pcm, peak, out, mnmx, tab, T, hop, W, ldk);
  return 0;
}

}  // namespace bn
