// bn_ingest.cu -- device-side audio ingest: decode -> channel mean -> polyphase resample -> peak normalise -> chunks.
//
// Reference: load_audio_window / fast_resample / split_audio_into_chunks, birdnet_stm32/audio/io.py:14-30,63-174.
// There the work is done per file on the host by libsndfile (int -> float32), numpy (mean over channels, max|y|,
// division) and scipy.signal.resample_poly (firwin Kaiser(5.0) low-pass of 20 * max(up, down) + 1 taps, upfirdn with
// zero extension).  Here the raw interleaved samples of a file window are shipped once (int16: half the bytes of the
// float32 the reference materialises) and every later step runs on the GPU, producing float32 chunks in HBM that
// bn_infer_wave_f32 consumes in place:
//
//   k_decode_mix   interleaved [n, ch] samples -> mono float32 [n]          (HBM bound: read n * ch * bytes, write 4 n)
//   k_resample     y[i] = sum_t H[k0 + t * up] * x[p / up - t], p = (i + pre) * down, k0 = p % up; products and sums
//                  in float32, oldest sample first, no FMA contraction -- the order of scipy's upfirdn inner loop -- plus a
//                  block reduction of max|y| merged with an integer atomicMax            (~ L / up taps per output)
//   k_chunks       chunk c, sample s: y[start_c + s] / peak (IEEE division, as numpy), zero padded short windows
//
// The filter is designed on the host in double precision exactly like scipy.signal.firwin + windows.kaiser and cached per
// (up, down); bn_ingest_filter exposes it so the tests can compare it with scipy's tap by tap.
#include "../../include/bn_ingest.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <utility>
#include <vector>

#include "bn_common.cuh"
#include "bn_kernels.cuh"

namespace bn {

constexpr int IG_THREADS = 256;

// ---------------------------------------------------------------------------------------------
// filter design (host, double precision)
// ---------------------------------------------------------------------------------------------
static long gcd_l(long a, long b) { while (b) { long t = a % b; a = b; b = t; } return a; }

// modified Bessel function I0 by its power series (arguments <= beta = 5 here; converges to double round-off)
static double bessel_i0(double x) {
  const double q = 0.25 * x * x;
  double term = 1.0, sum = 1.0;
  for (int k = 1; k < 200; k++) {
    term *= q / ((double)k * (double)k);
    sum += term;
    if (term < 1e-18 * sum) break;
  }
  return sum;
}

struct PolyFilter {
  int up = 1, down = 1;
  int n_taps = 0;          // padded taps handed to upfirdn
  int n_pre_remove = 0;
  int per_phase = 0;       // ceil(n_taps / up)
  std::vector<float> h;    // [n_taps]
};

static long output_len(long len_h, long in_len, long up, long down) { return (((in_len - 1) * up + len_h) - 1) / down + 1; }

// resample_poly's filter for the reduced ratio up / down and an input of n_in samples (n_in only matters for the rare
// post-padding loop): firwin(2 * half_len + 1, 1 / max_rate, window=("kaiser", 5.0)) cast to float32, times up, padded.
static void design_filter(int up, int down, long n_in, PolyFilter& F) {
  const double PI = 3.14159265358979323846;
  const int max_rate = up > down ? up : down;
  const double f_c = 1.0 / max_rate;
  const int half_len = 10 * max_rate;
  const int numtaps = 2 * half_len + 1;
  const double alpha = 0.5 * (numtaps - 1);
  std::vector<double> h(numtaps);
  const double i0b = bessel_i0(5.0);
  double s = 0.0;
  for (int n = 0; n < numtaps; n++) {
    const double m = n - alpha;
    const double xs = f_c * m;
    const double y = PI * (xs == 0.0 ? 1.0e-20 : xs);
    const double sinc = sin(y) / y;
    const double r = (n - alpha) / alpha;
    double arg = 1.0 - r * r;
    if (arg < 0.0) arg = 0.0;
    const double win = bessel_i0(5.0 * sqrt(arg)) / i0b;
    h[n] = f_c * sinc * win;
    s += h[n];
  }
  const int n_pre_pad = down - half_len % down;
  int n_post_pad = 0;
  const int n_pre_remove = (half_len + n_pre_pad) / down;
  long n_out = n_in * up;
  n_out = n_out / down + (n_out % down ? 1 : 0);
  while (output_len((long)numtaps + n_pre_pad + n_post_pad, n_in, up, down) < n_out + n_pre_remove) n_post_pad++;
  F.up = up; F.down = down;
  F.n_taps = numtaps + n_pre_pad + n_post_pad;
  F.n_pre_remove = n_pre_remove;
  F.per_phase = (F.n_taps + up - 1) / up;
  F.h.assign(F.n_taps, 0.0f);
  const float upf = (float)up;
  for (int n = 0; n < numtaps; n++) F.h[n_pre_pad + n] = (float)(h[n] / s) * upf;   // float32 cast, then the float32 "h *= up"
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
template <int FMT>
__device__ __forceinline__ float decode_sample(const unsigned char* base, long idx) {
  if (FMT == BN_SF_S16) return (float)reinterpret_cast<const int16_t*>(base)[idx] * (1.0f / 32768.0f);
  if (FMT == BN_SF_S32) return (float)reinterpret_cast<const int32_t*>(base)[idx] * (1.0f / 2147483648.0f);
  if (FMT == BN_SF_F32) return reinterpret_cast<const float*>(base)[idx];
  if (FMT == BN_SF_U8) return (float)((int)base[idx] - 128) * (1.0f / 128.0f);
  // packed 24-bit little endian, sign extended
  const unsigned char* p = base + 3 * idx;
  const int v = (int)((unsigned)p[0] << 8 | (unsigned)p[1] << 16 | (unsigned)p[2] << 24) >> 8;
  return (float)v * (1.0f / 8388608.0f);
}

// mono[i] = (((c0 + c1) + c2) + ...) / ch in float32: numpy's mean over a short contiguous axis (sequential add.reduce
// below the pairwise-summation block size of 8, then true_divide by the count)
template <int FMT>
__global__ void __launch_bounds__(IG_THREADS)
k_decode_mix(const unsigned char* __restrict__ frames, float* __restrict__ mono, long n, int ch) {
  const float fch = (float)ch;
  for (long i = blockIdx.x * (long)IG_THREADS + threadIdx.x; i < n; i += (long)gridDim.x * IG_THREADS) {
    float s = decode_sample<FMT>(frames, i * ch);
    for (int c = 1; c < ch; c++) s = __fadd_rn(s, decode_sample<FMT>(frames, i * ch + c));
    mono[i] = ch > 1 ? __fdiv_rn(s, fch) : s;
  }
}

__device__ __forceinline__ void block_absmax(float v, unsigned* peak_bits) {
  __shared__ float red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); i++) m = fmaxf(m, red[i]);
    atomicMax(peak_bits, __float_as_uint(m));        // |y| >= +0: bit patterns order like unsigned integers
  }
  __syncthreads();
}

// hp: taps re-laid-out phase-major [up][per_phase] (hp[k0][t] = H[k0 + t * up], zero beyond the last tap)
__global__ void __launch_bounds__(IG_THREADS)
k_resample(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ hp, unsigned* __restrict__ peak_bits,
           long n_in, long n_out, int up, int down, int per_phase, int n_pre_remove) {
  float amax = 0.0f;
  for (long i0 = blockIdx.x * (long)IG_THREADS; i0 < n_out; i0 += (long)gridDim.x * IG_THREADS) {
    const long i = i0 + threadIdx.x;
    if (i < n_out) {
      const long p = (i + n_pre_remove) * (long)down;
      const long xi = p / up;
      const int k0 = (int)(p - xi * up);
      const float* hrow = hp + (size_t)k0 * per_phase;
      float acc = 0.0f;
      for (int t = per_phase - 1; t >= 0; t--) {        // oldest input sample first (scipy upfirdn's loop order)
        const long idx = xi - t;
        if (idx >= 0 && idx < n_in) acc = __fadd_rn(acc, __fmul_rn(__ldg(x + idx), __ldg(hrow + t)));
      }
      y[i] = acc;
      amax = fmaxf(amax, fabsf(acc));
    }
  }
  block_absmax(amax, peak_bits);
}

// Phase-per-thread resampler.  The filter phase of output i only depends on i mod up, so a block whose output stride S is
// a multiple of up gives every thread ONE phase for its whole life: its per_phase taps sit in shared memory as a column
// hs[t][tid] (conflict-free: consecutive threads, consecutive words) and never have to be gathered again, and the input
// index advances by exactly (S / up) * down per step.  Products and sums are the same float32 operations in the same
// order as k_resample (oldest sample first, no FMA), so both kernels give identical bits.
//   block b owns outputs [b * S * J, (b + 1) * S * J); thread tid handles b * S * J + j * S + tid, j < J.
__global__ void __launch_bounds__(1024)
k_resample_phase(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ hcol, unsigned* __restrict__ peak_bits,
                 long n_in, long n_out, int up, int down, int per_phase, int n_pre_remove, int S, int J) {
  extern __shared__ float hs[];
  const int tid = threadIdx.x;
  const bool active = tid < S;
  // hcol is the tap table already in column order [t][tid] (built on the host for this S): straight coalesced copy
  // (16-byte cp.async pieces: the host pads the table to a multiple of 4 floats and aligns it)
  for (int i = tid; i < (per_phase * S + 3) >> 2; i += blockDim.x)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(hs + 4 * i)), "l"(hcol + 4 * i) : "memory");
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  float amax = 0.0f;
  if (active) {
    const long xi0 = ((long)(tid + n_pre_remove) * down) / up;
    const long xstep = (long)(S / up) * down;
    const long base = (long)blockIdx.x * S * J;
    long xi = xi0 + (base / up) * down;
    const float* hc = hs + tid;
    // four outputs of this thread at a time: they share every tap (same phase), so a tap costs one shared-memory load,
    // four cached global loads and four independent multiply / add chains.  Groups that touch either end of the signal
    // (or the end of the output) take the bounds-checked loop; only the first and last blocks ever do.
    for (int j = 0; j < J; j += 4, xi += 4 * xstep) {
      const long i = base + (long)j * S + tid;
      if (i >= n_out) break;
      if (j + 4 <= J && i + 3L * S < n_out && xi - (per_phase - 1) >= 0 && xi + 3 * xstep < n_in) {
        const float* q0 = x + xi - (per_phase - 1);        // oldest sample of the first output
        const float* q1 = q0 + xstep;
        const float* q2 = q1 + xstep;
        const float* q3 = q2 + xstep;
        const float* hq = hc + (per_phase - 1) * S;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll 4
        for (int tt = 0; tt < per_phase; tt++) {           // oldest sample first (scipy upfirdn's loop order)
          const float h = hq[-tt * S];
          a0 = __fadd_rn(a0, __fmul_rn(__ldg(q0 + tt), h));
          a1 = __fadd_rn(a1, __fmul_rn(__ldg(q1 + tt), h));
          a2 = __fadd_rn(a2, __fmul_rn(__ldg(q2 + tt), h));
          a3 = __fadd_rn(a3, __fmul_rn(__ldg(q3 + tt), h));
        }
        y[i] = a0; y[i + S] = a1; y[i + 2L * S] = a2; y[i + 3L * S] = a3;
        amax = fmaxf(fmaxf(amax, fmaxf(fabsf(a0), fabsf(a1))), fmaxf(fabsf(a2), fabsf(a3)));
      } else {
        for (int jj = 0; jj < 4 && j + jj < J; jj++) {
          const long ii = i + (long)jj * S;
          if (ii >= n_out) break;
          const long xj = xi + jj * xstep;
          float acc = 0.0f;
          for (int t = per_phase - 1; t >= 0; t--) {
            const long idx = xj - t;
            if (idx >= 0 && idx < n_in) acc = __fadd_rn(acc, __fmul_rn(__ldg(x + idx), hc[t * S]));
          }
          y[ii] = acc;
          amax = fmaxf(amax, fabsf(acc));
        }
      }
    }
  }
  block_absmax(amax, peak_bits);
}

__global__ void __launch_bounds__(IG_THREADS)
k_absmax(const float* __restrict__ y, long n, unsigned* __restrict__ peak_bits) {
  float amax = 0.0f;
  for (long i = blockIdx.x * (long)IG_THREADS + threadIdx.x; i < n; i += (long)gridDim.x * IG_THREADS) amax = fmaxf(amax, fabsf(y[i]));
  block_absmax(amax, peak_bits);
}

// in place: y / peak when peak > 0 (audio/io.py:122-124)
__global__ void __launch_bounds__(IG_THREADS)
k_normalize(float* __restrict__ y, long n, const unsigned* __restrict__ peak_bits) {
  const float pk = __uint_as_float(*peak_bits);
  if (!(pk > 0.0f)) return;
  for (long i = blockIdx.x * (long)IG_THREADS + threadIdx.x; i < n; i += (long)gridDim.x * IG_THREADS) y[i] = __fdiv_rn(y[i], pk);
}

// split_audio_into_chunks (audio/io.py:133-174) fused with the peak division: chunk c starts at c * step for c < n_full
// and at n - chunk_len for the tail chunk; a window shorter than a chunk is one right-zero-padded chunk.
__global__ void __launch_bounds__(IG_THREADS)
k_chunks(const float* __restrict__ y, long n, float* __restrict__ out, int n_chunks, int n_full, int chunk_len, int step,
         const unsigned* __restrict__ peak_bits) {
  const float pk = __uint_as_float(*peak_bits);
  const bool norm = pk > 0.0f;
  const long total = (long)n_chunks * chunk_len;
  for (long i = blockIdx.x * (long)IG_THREADS + threadIdx.x; i < total; i += (long)gridDim.x * IG_THREADS) {
    const int c = (int)(i / chunk_len);
    const int s = (int)(i - (long)c * chunk_len);
    const long start = c < n_full ? (long)c * step : n - chunk_len;
    const long src = (n <= chunk_len ? 0 : start) + s;
    float v = 0.0f;
    if (src < n) { v = y[src]; if (norm) v = __fdiv_rn(v, pk); }
    out[i] = v;
  }
}

}  // namespace bn

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
using namespace bn;

struct bn_ingest {
  int device = 0;
  int sms = 148;
  cudaStream_t s_own = nullptr;
  unsigned char* d_frames = nullptr; size_t frames_cap = 0;
  float* d_mono = nullptr; size_t mono_cap = 0;
  float* d_y = nullptr; size_t y_cap = 0;
  float* d_out = nullptr; size_t out_cap = 0;
  unsigned* d_peak = nullptr;
  std::map<std::pair<int, int>, std::pair<PolyFilter, float*>> filters;   // (up, down) -> host taps, device phase-major taps
                                                                          // followed by the column table of k_resample_phase
  int64_t launches = 0;
};

static bool ig_is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

#define IG_CU(call)                                                                                  \
  do {                                                                                               \
    cudaError_t _e = (call);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      char _m[256];                                                                                  \
      snprintf(_m, sizeof _m, "%s: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__);  \
      return set_error(BN_ERR_CUDA, _m);                                                             \
    }                                                                                                \
  } while (0)

static int fmt_bytes(int fmt) {
  switch (fmt) {
    case BN_SF_S16: return 2;
    case BN_SF_S24: return 3;
    case BN_SF_S32: return 4;
    case BN_SF_F32: return 4;
    case BN_SF_U8: return 1;
    default: return 0;
  }
}

template <typename T>
static int ig_reserve(T** p, size_t* cap, size_t need) {
  if (*cap >= need) return 0;
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  size_t n = need + need / 4 + 64;
  if (cudaMalloc((void**)p, n * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return set_error(BN_ERR_CUDA, "cudaMalloc failed in the ingest workspace"); }
  *cap = n;
  return 0;
}

extern "C" int bn_ingest_create(int device, bn_ingest** out) {
  if (!out) return set_error(BN_ERR_ARG, "bn_ingest_create: NULL argument");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    return set_error(BN_ERR_CUDA, "no CUDA device for the ingest kernels (there is no CPU fallback)");
  }
  IG_CU(cudaSetDevice(device));
  bn_ingest* g = new bn_ingest();
  g->device = device;
  cudaDeviceGetAttribute(&g->sms, cudaDevAttrMultiProcessorCount, device);
  if (g->sms <= 0) g->sms = 148;
  if (cudaStreamCreateWithFlags(&g->s_own, cudaStreamNonBlocking) != cudaSuccess || cudaMalloc((void**)&g->d_peak, sizeof(unsigned)) != cudaSuccess) {
    bn_ingest_destroy(g);
    return set_error(BN_ERR_CUDA, "ingest setup failed");
  }
  *out = g;
  return BN_OK;
}

extern "C" void bn_ingest_destroy(bn_ingest* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  for (auto& kv : g->filters) if (kv.second.second) cudaFree(kv.second.second);
  if (g->d_frames) cudaFree(g->d_frames);
  if (g->d_mono) cudaFree(g->d_mono);
  if (g->d_y) cudaFree(g->d_y);
  if (g->d_out) cudaFree(g->d_out);
  if (g->d_peak) cudaFree(g->d_peak);
  if (g->s_own) cudaStreamDestroy(g->s_own);
  delete g;
}

extern "C" int64_t bn_ingest_out_len(int64_t n_frames, int sr_in, int sr_out) {
  if (n_frames <= 0 || sr_in <= 0 || sr_out <= 0) return 0;
  if (sr_in == sr_out) return n_frames;
  const long g = gcd_l(sr_in, sr_out);
  const long up = sr_out / g, down = sr_in / g;
  const long n = n_frames * up;
  return n / down + (n % down ? 1 : 0);
}

extern "C" int bn_ingest_num_chunks(int64_t n_samples, int chunk_len, int step) {
  if (n_samples <= 0 || chunk_len <= 0) return 0;
  if (n_samples <= chunk_len) return 1;
  if (step < 1) step = 1;
  const long n_full = 1 + (n_samples - chunk_len) / step;
  const long tail = (n_samples - chunk_len) % step != 0;
  return (int)(n_full + tail);
}

extern "C" int bn_ingest_filter(int up, int down, float* h_out, int cap, int* n_taps, int* n_pre_remove) {
  if (up < 1 || down < 1 || !n_taps) return set_error(BN_ERR_ARG, "bn_ingest_filter: bad arguments");
  const long g = gcd_l(up, down);
  up /= (int)g; down /= (int)g;
  if (up == 1 && down == 1) { *n_taps = 0; if (n_pre_remove) *n_pre_remove = 0; return BN_OK; }
  PolyFilter F;
  design_filter(up, down, 1L << 20, F);
  *n_taps = F.n_taps;
  if (n_pre_remove) *n_pre_remove = F.n_pre_remove;
  if (h_out) {
    if (cap < F.n_taps) return set_error(BN_ERR_ARG, "bn_ingest_filter: buffer too small");
    memcpy(h_out, F.h.data(), sizeof(float) * F.n_taps);
  }
  return BN_OK;
}

// output stride of a k_resample_phase block: a whole number of phase periods, about 256 threads; 0 = does not fit a block
static int phase_stride(int up) {
  const int S = up <= 256 ? up * ((256 + up - 1) / up) : up;
  return S <= 1024 ? S : 0;
}

static int grid_for(const bn_ingest* g, long n) {
  long b = (n + IG_THREADS - 1) / IG_THREADS;
  const long cap = (long)g->sms * 8;
  if (b > cap) b = cap;
  return b < 1 ? 1 : (int)b;
}

// decode + mix + resample of one window; the result (not yet normalised) is left in *d_res [n_out] and the peak bits in
// g->d_peak.  d_res aliases the caller's buffer when that is device memory and no chunking follows.
static int ingest_core(bn_ingest* g, const void* frames, int fmt, int64_t n_frames, int channels, int sr_in, int sr_out,
                       float* d_dst, float** d_res, int64_t* n_out_p, cudaStream_t st) {
  const int sb = fmt_bytes(fmt);
  if (!sb) return set_error(BN_ERR_ARG, "unknown sample format");
  if (channels < 1 || channels > 64) return set_error(BN_ERR_ARG, "channels must be in 1..64");
  if (sr_in <= 0 || sr_out <= 0) return set_error(BN_ERR_ARG, "sample rates must be positive");
  if (n_frames > (1L << 31)) return set_error(BN_ERR_ARG, "window too long");
  const size_t nbytes = (size_t)n_frames * channels * sb;
  const unsigned char* d_fr = (const unsigned char*)frames;
  if (!ig_is_device_ptr(frames)) {
    int rc = ig_reserve(&g->d_frames, &g->frames_cap, nbytes + 16);
    if (rc) return rc;
    IG_CU(cudaMemcpyAsync(g->d_frames, frames, nbytes, cudaMemcpyHostToDevice, st));
    d_fr = g->d_frames;
  }
  const long gg = gcd_l(sr_in, sr_out);
  const int up = (int)(sr_out / gg), down = (int)(sr_in / gg);
  const bool same = up == 1 && down == 1;
  const int64_t n_out = bn_ingest_out_len(n_frames, sr_in, sr_out);
  *n_out_p = n_out;
  float* d_y = d_dst;
  if (!d_y) {
    int rc = ig_reserve(&g->d_y, &g->y_cap, (size_t)n_out);
    if (rc) return rc;
    d_y = g->d_y;
  }
  float* d_mono = d_y;                                 // same rate: decode straight into the result
  if (!same) {
    int rc = ig_reserve(&g->d_mono, &g->mono_cap, (size_t)n_frames);
    if (rc) return rc;
    d_mono = g->d_mono;
  }
  IG_CU(cudaMemsetAsync(g->d_peak, 0, sizeof(unsigned), st));
  const int gr = grid_for(g, n_frames);
  switch (fmt) {
    case BN_SF_S16: k_decode_mix<BN_SF_S16><<<gr, IG_THREADS, 0, st>>>(d_fr, d_mono, n_frames, channels); break;
    case BN_SF_S24: k_decode_mix<BN_SF_S24><<<gr, IG_THREADS, 0, st>>>(d_fr, d_mono, n_frames, channels); break;
    case BN_SF_S32: k_decode_mix<BN_SF_S32><<<gr, IG_THREADS, 0, st>>>(d_fr, d_mono, n_frames, channels); break;
    case BN_SF_F32: k_decode_mix<BN_SF_F32><<<gr, IG_THREADS, 0, st>>>(d_fr, d_mono, n_frames, channels); break;
    default: k_decode_mix<BN_SF_U8><<<gr, IG_THREADS, 0, st>>>(d_fr, d_mono, n_frames, channels); break;
  }
  g->launches++;
  if (same) {
    k_absmax<<<grid_for(g, n_out), IG_THREADS, 0, st>>>(d_y, n_out, g->d_peak);
    g->launches++;
  } else {
    auto key = std::make_pair(up, down);
    auto it = g->filters.find(key);
    // the post-padding loop of resample_poly depends on n_in only in degenerate cases; redesign when it would differ
    PolyFilter probe;
    if (it == g->filters.end()) {
      design_filter(up, down, n_frames, probe);
    } else {
      const PolyFilter& F0 = it->second.first;
      if (output_len(F0.n_taps, n_frames, up, down) < n_out + F0.n_pre_remove) {
        design_filter(up, down, n_frames, probe);
        cudaFree(it->second.second);
        g->filters.erase(it);
        it = g->filters.end();
      }
    }
    if (it == g->filters.end()) {
      const size_t n_hp = ((size_t)up * probe.per_phase + 3) & ~(size_t)3;   // keeps the column table 16-byte aligned
      const int S = phase_stride(up);
      std::vector<float> hp(n_hp + (S ? (size_t)S * probe.per_phase + 4 : 0), 0.0f);
      for (int k = 0; k < probe.n_taps; k++) hp[(size_t)(k % up) * probe.per_phase + k / up] = probe.h[k];
      for (int tid = 0; tid < S; tid++) {                 // column of thread tid = taps of its phase, [t][tid]
        const long p0 = (long)(tid + probe.n_pre_remove) * down;
        const int k0 = (int)(p0 % up);
        for (int t = 0; t < probe.per_phase; t++) hp[n_hp + (size_t)t * S + tid] = hp[(size_t)k0 * probe.per_phase + t];
      }
      float* d_hp = nullptr;
      IG_CU(cudaMalloc((void**)&d_hp, hp.size() * sizeof(float)));
      IG_CU(cudaMemcpyAsync(d_hp, hp.data(), hp.size() * sizeof(float), cudaMemcpyHostToDevice, st));
      IG_CU(cudaStreamSynchronize(st));                // hp is a local
      it = g->filters.emplace(key, std::make_pair(probe, d_hp)).first;
    }
    const PolyFilter& F = it->second.first;
    // phase-per-thread kernel when a whole number of phase periods fits a block and the tap columns fit shared memory
    const int S = phase_stride(up);
    const size_t hs_bytes = ((size_t)F.per_phase * S + 4) * sizeof(float);
    static const bool force_generic = getenv("BN_INGEST_GENERIC") != nullptr;
    if (S > 0 && hs_bytes <= 200 * 1024 && !force_generic) {
      static unsigned long long attr = 0;
      if (first_use_on_device(attr)) cudaFuncSetAttribute(k_resample_phase, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      const long steps = (n_out + S - 1) / S;                       // block steps in total
      const int per_sm = (int)(200 * 1024 / (hs_bytes + 1024)) < (2048 / ((S + 31) & ~31)) ? (int)(200 * 1024 / (hs_bytes + 1024)) : (2048 / ((S + 31) & ~31));
      long blocks = (long)g->sms * (per_sm < 1 ? 1 : per_sm);
      if (blocks > steps) blocks = steps;
      int J = (int)((steps + blocks - 1) / blocks);
      J = (J + 3) & ~3;                                  // whole groups of four outputs per thread
      blocks = (steps + J - 1) / J;
      k_resample_phase<<<(int)blocks, (S + 31) & ~31, hs_bytes, st>>>(d_mono, d_y, it->second.second + (((size_t)up * F.per_phase + 3) & ~(size_t)3), g->d_peak, n_frames, n_out, up, down,
                                                                   F.per_phase, F.n_pre_remove, S, J);
    } else {
      k_resample<<<grid_for(g, n_out), IG_THREADS, 0, st>>>(d_mono, d_y, it->second.second, g->d_peak, n_frames, n_out, up, down,
                                                           F.per_phase, F.n_pre_remove);
    }
    g->launches++;
  }
  *d_res = d_y;
  return 0;
}

extern "C" int bn_ingest_window(bn_ingest* g, const void* frames, int fmt, int64_t n_frames, int channels, int sr_in, int sr_out,
                                int normalize, float* wave_out, float* peak_out, void* stream) {
  if (!g) return set_error(BN_ERR_ARG, "NULL ingest object");
  if (n_frames < 0 || (n_frames > 0 && (!frames || !wave_out))) return set_error(BN_ERR_ARG, "bn_ingest_window: bad arguments");
  IG_CU(cudaSetDevice(g->device));
  if (n_frames == 0) return BN_OK;
  const bool dev_out = ig_is_device_ptr(wave_out);
  const bool dev_peak = peak_out && ig_is_device_ptr(peak_out);
  const bool all_dev = ig_is_device_ptr(frames) && dev_out && (!peak_out || dev_peak);
  cudaStream_t st = all_dev ? (cudaStream_t)stream : g->s_own;
  float* d_res = nullptr;
  int64_t n_out = 0;
  int rc = ingest_core(g, frames, fmt, n_frames, channels, sr_in, sr_out, dev_out ? wave_out : nullptr, &d_res, &n_out, st);
  if (rc) return rc;
  if (normalize) { k_normalize<<<grid_for(g, n_out), IG_THREADS, 0, st>>>(d_res, n_out, g->d_peak); g->launches++; }
  if (!dev_out) IG_CU(cudaMemcpyAsync(wave_out, d_res, sizeof(float) * (size_t)n_out, cudaMemcpyDeviceToHost, st));
  if (peak_out) IG_CU(cudaMemcpyAsync(peak_out, g->d_peak, sizeof(float), dev_peak ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  IG_CU(cudaGetLastError());
  if (!all_dev) IG_CU(cudaStreamSynchronize(st));
  return BN_OK;
}

extern "C" int bn_ingest_chunks(bn_ingest* g, const void* frames, int fmt, int64_t n_frames, int channels, int sr_in, int sr_out,
                                int chunk_len, int step, float* chunks_out, int max_chunks, int* n_chunks, float* peak_out,
                                void* stream) {
  if (!g) return set_error(BN_ERR_ARG, "NULL ingest object");
  if (!n_chunks || chunk_len <= 0 || n_frames < 0) return set_error(BN_ERR_ARG, "bn_ingest_chunks: bad arguments");
  *n_chunks = 0;
  if (n_frames == 0) return BN_OK;
  if (!frames || !chunks_out) return set_error(BN_ERR_ARG, "bn_ingest_chunks: NULL buffer");
  IG_CU(cudaSetDevice(g->device));
  if (step < 1) step = 1;
  const int64_t n_out = bn_ingest_out_len(n_frames, sr_in, sr_out);
  const int nc = bn_ingest_num_chunks(n_out, chunk_len, step);
  if (nc > max_chunks) return set_error(BN_ERR_ARG, "bn_ingest_chunks: max_chunks too small");
  const int n_full = n_out <= chunk_len ? 1 : (int)(1 + (n_out - chunk_len) / step);
  const bool dev_out = ig_is_device_ptr(chunks_out);
  const bool dev_peak = peak_out && ig_is_device_ptr(peak_out);
  const bool all_dev = ig_is_device_ptr(frames) && dev_out && (!peak_out || dev_peak);
  cudaStream_t st = all_dev ? (cudaStream_t)stream : g->s_own;
  float* d_res = nullptr;
  int64_t n_out2 = 0;
  int rc = ingest_core(g, frames, fmt, n_frames, channels, sr_in, sr_out, nullptr, &d_res, &n_out2, st);
  if (rc) return rc;
  float* d_chunks = chunks_out;
  const size_t total = (size_t)nc * chunk_len;
  if (!dev_out) {
    rc = ig_reserve(&g->d_out, &g->out_cap, total);
    if (rc) return rc;
    d_chunks = g->d_out;
  }
  k_chunks<<<grid_for(g, (long)total), IG_THREADS, 0, st>>>(d_res, n_out, d_chunks, nc, n_full, chunk_len, step, g->d_peak);
  g->launches++;
  if (!dev_out) IG_CU(cudaMemcpyAsync(chunks_out, d_chunks, sizeof(float) * total, cudaMemcpyDeviceToHost, st));
  if (peak_out) IG_CU(cudaMemcpyAsync(peak_out, g->d_peak, sizeof(float), dev_peak ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  IG_CU(cudaGetLastError());
  *n_chunks = nc;
  if (!all_dev) IG_CU(cudaStreamSynchronize(st));
  return BN_OK;
}

extern "C" int64_t bn_ingest_launch_count(const bn_ingest* g) { return g ? g->launches : 0; }
