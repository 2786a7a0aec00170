// bn_features.cu -- precomputed-spectrogram ("librosa") frontends on the GPU: mel / log-mel / MFCC features with
// none | pwl | pcen | db magnitude scaling and per-chunk min-max normalisation.
//
// Reference: get_spectrogram_from_audio, birdnet_stm32/audio/spectrogram.py:24-149 (per chunk, on the host, via
// librosa.feature.melspectrogram / librosa.pcen / librosa.amplitude_to_db / librosa.feature.mfcc).
//
//   K1  k_stft_mag<frame-major>  (bn_frontend.cu)   PCM16 -> |STFT| float32 [B][Wk][264], Wk = frames rounded up to 32
//   KF  k_feat                                      one CTA per chunk: banded mel projection into a shared-memory tile
//                                                   [n_mels][frames], magnitude scaling with CTA-wide min / max
//                                                   reductions, (MFCC: dB + DCT), normalize(), store [rows][W]
//
// Everything a chunk needs after the STFT lives in shared memory (<= 74 KB mel tile), so the features make one trip:
// magnitudes in (scratch written by K1 just before), normalised features out.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/bn_features.h"
#include "bn_common.cuh"
#include "bn_kernels.cuh"

namespace bn {

constexpr int KF_THREADS = 256;
constexpr int KF_LDK = 264;          // floats per frame row of the magnitude scratch
constexpr int KF_FG = 32;            // frames staged per step
constexpr int KF_SUBWAVE = 1184;     // chunks per K1 -> KF round (4 x 296 resident K1 CTAs; measured on the classification path: fewer,
                                     // larger launches beat keeping the 270 KB/chunk scratch inside L2, see bn_fast.cu fe_subwave())

struct FeatDev {
  const float* basis;   // [n_mels][bins]
  const int2* band;     // [n_mels] first / one-past-last non-zero bin
  const float* dct;     // [n_mfcc][n_mels]
  int bins, n_mels, W, Wk, WV, SP;   // WV = frames used by the scaling stage (W, or all frames for MFCC); SP = tile row stride
  int mode, mag_scale, n_mfcc, rows;
  int aux_floats;       // size of the auxiliary shared-memory tile
  float pcen_b;
};

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// CTA-wide min and max of tile[r][t], r < R, t < Wv (row stride SP).  All threads get the result.
__device__ void block_minmax(const float* tile, int R, int Wv, int SP, float* red, float& mn, float& mx) {
  float lo = __int_as_float(0x7f800000), hi = -__int_as_float(0x7f800000);
  for (int i = threadIdx.x; i < R * Wv; i += KF_THREADS) {
    const int r = i / Wv, t = i - r * Wv;
    const float v = tile[r * SP + t];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  lo = warp_min(lo);
  hi = warp_max(hi);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = lo; red[8 + (threadIdx.x >> 5)] = hi; }
  __syncthreads();
  mn = red[0]; mx = red[8];
#pragma unroll
  for (int i = 1; i < KF_THREADS / 32; i++) { mn = fminf(mn, red[i]); mx = fmaxf(mx, red[8 + i]); }
}

// normalize(): (S - min) / (max - min + 1e-10), the denominator formed in double like numpy's scalar promotion
__device__ __forceinline__ float norm_den(float mn, float mx) { return (float)((double)(mx - mn) + 1e-10); }

__global__ void __launch_bounds__(KF_THREADS)
k_feat(const float* __restrict__ mags, float* __restrict__ out, FeatDev P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);                 // [n_mels][SP]
  float* stage = tile + P.n_mels * P.SP;                            // [KF_FG][KF_LDK]
  float* ctile = stage + KF_FG * KF_LDK;                            // MFCC: [n_mfcc][W + 1]; PCEN: smoothed energy [n_mels][WV]
  float* red = ctile + P.aux_floats;
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const float* mb = mags + (size_t)b * P.Wk * KF_LDK;

  // ---- (A) banded mel projection, 32 frames at a time ------------------------------------------------------------
  const bool squared = P.mode == BN_FEAT_MFCC;                      // melspectrogram(power=2.0) for MFCC, 1.0 otherwise
  for (int fg = 0; fg * KF_FG < P.WV; fg++) {
    const float4* src = reinterpret_cast<const float4*>(mb + (size_t)fg * KF_FG * KF_LDK);
    for (int i = tid; i < KF_FG * KF_LDK / 4; i += KF_THREADS) {
      float4 v = __ldg(src + i);
      if (squared) { v.x *= v.x; v.y *= v.y; v.z *= v.z; v.w *= v.w; }
      reinterpret_cast<float4*>(stage)[i] = v;
    }
    __syncthreads();
    for (int task = tid; task < P.n_mels * 4; task += KF_THREADS) {
      const int m = task % P.n_mels, tq = task / P.n_mels;          // 4 groups of 8 frames
      const int2 bd = __ldg(P.band + m);
      const float* wrow = P.basis + (size_t)m * P.bins;
      const float* st = stage + (tq * 8) * KF_LDK;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int f = bd.x; f < bd.y; f++) {
        const float w = __ldg(wrow + f);
#pragma unroll
        for (int j = 0; j < 8; j++) acc[j] = fmaf(w, st[j * KF_LDK + f], acc[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int t = fg * KF_FG + tq * 8 + j;
        if (t < P.WV) tile[m * P.SP + t] = acc[j];
      }
    }
    __syncthreads();
  }

  const int R = P.n_mels, Wv = P.WV, SP = P.SP;
  float mn, mx;
  // ---- (B) magnitude scaling -------------------------------------------------------------------------------------
  if (P.mode == BN_FEAT_LOG_MEL) {
    for (int i = tid; i < R * Wv; i += KF_THREADS) { const int r = i / Wv, t = i - r * Wv; tile[r * SP + t] = log1pf(tile[r * SP + t]); }
  } else if (P.mode == BN_FEAT_MFCC) {
    // power_to_db(S, ref=np.max, amin=1e-10, top_db=80) over ALL frames (the slice to W comes after the DCT)
    block_minmax(tile, R, Wv, SP, red, mn, mx);
    const float refdb = 10.0f * log10f(fmaxf(1e-10f, fabsf(mx)));
    for (int i = tid; i < R * Wv; i += KF_THREADS) {
      const int r = i / Wv, t = i - r * Wv;
      tile[r * SP + t] = 10.0f * log10f(fmaxf(1e-10f, tile[r * SP + t])) - refdb;
    }
    __syncthreads();
    block_minmax(tile, R, Wv, SP, red, mn, mx);
    const float floor_db = mx - 80.0f;
    // orthonormal DCT-II along the mel axis, first n_mfcc rows, frames < W
    for (int i = tid; i < P.n_mfcc * P.W; i += KF_THREADS) {
      const int k = i / P.W, t = i - k * P.W;
      const float* d = P.dct + (size_t)k * R;
      float acc = 0.f;
      for (int m = 0; m < R; m++) acc = fmaf(__ldg(d + m), fmaxf(tile[m * SP + t], floor_db), acc);
      ctile[k * (P.W + 1) + t] = acc;
    }
    __syncthreads();
    block_minmax(ctile, P.n_mfcc, P.W, P.W + 1, red, mn, mx);
    const float den = norm_den(mn, mx);
    float* ob = out + (size_t)b * P.n_mfcc * P.W;
    for (int i = tid; i < P.n_mfcc * P.W; i += KF_THREADS) {
      const int k = i / P.W, t = i - k * P.W;
      ob[i] = __fdiv_rn(ctile[k * (P.W + 1) + t] - mn, den);
    }
    return;
  } else if (P.mag_scale == BN_MAG_PWL) {
    block_minmax(tile, R, Wv, SP, red, mn, mx);
    const float den = norm_den(mn, mx);
    for (int i = tid; i < R * Wv; i += KF_THREADS) {
      const int r = i / Wv, t = i - r * Wv;
      const float sn = __fdiv_rn(tile[r * SP + t] - mn, den);
      float s = 0.40f * sn;
      s += 0.25f * fmaxf(sn - 0.10f, 0.0f);
      s += 0.15f * fmaxf(sn - 0.35f, 0.0f);
      s += 0.08f * fmaxf(sn - 0.65f, 0.0f);
      tile[r * SP + t] = s;
    }
  } else if (P.mag_scale == BN_MAG_PCEN) {
    // librosa.pcen(S * 2^31): first-order IIR along time per band (state = lfilter_zi = 1 - b), then the
    // gain / bias / power compression with gain 0.98, bias 2, power 0.5, eps 1e-6
    const float bb = P.pcen_b;
    for (int r = tid; r < R; r += KF_THREADS) {
      float z = 1.0f - bb;
      for (int t = 0; t < Wv; t++) {
        const float x = tile[r * SP + t] * 2147483648.0f;
        const float y = fmaf(bb, x, z);
        z = (1.0f - bb) * y;
        ctile[r * Wv + t] = y;                                        // smoothed energy
      }
    }
    __syncthreads();
    const float log_eps = logf(1e-6f);
    for (int i = tid; i < R * Wv; i += KF_THREADS) {
      const int r = i / Wv, t = i - r * Wv;
      const float x = tile[r * SP + t] * 2147483648.0f;
      const float smooth = expf(-0.98f * (log_eps + log1pf(ctile[i] / 1e-6f)));
      tile[r * SP + t] = 1.41421356237309515f * expm1f(0.5f * log1pf(x * smooth * 0.5f));
    }
  } else if (P.mag_scale == BN_MAG_DB) {
    // amplitude_to_db(S, ref=np.max): power_to_db(S^2, ref=max^2, amin=1e-10, top_db=80)
    block_minmax(tile, R, Wv, SP, red, mn, mx);
    const float refdb = 10.0f * log10f(fmaxf(1e-10f, mx * mx));
    for (int i = tid; i < R * Wv; i += KF_THREADS) {
      const int r = i / Wv, t = i - r * Wv;
      const float s = tile[r * SP + t];
      tile[r * SP + t] = 10.0f * log10f(fmaxf(1e-10f, s * s)) - refdb;
    }
    __syncthreads();
    block_minmax(tile, R, Wv, SP, red, mn, mx);
    const float floor_db = mx - 80.0f;
    for (int i = tid; i < R * Wv; i += KF_THREADS) { const int r = i / Wv, t = i - r * Wv; tile[r * SP + t] = fmaxf(tile[r * SP + t], floor_db); }
  }
  __syncthreads();
  // ---- (C) normalize() and store ----------------------------------------------------------------------------------
  block_minmax(tile, R, Wv, SP, red, mn, mx);
  const float den = norm_den(mn, mx);
  float* ob = out + (size_t)b * R * P.W;
  for (int i = tid; i < R * P.W; i += KF_THREADS) {
    const int r = i / P.W, t = i - r * P.W;
    ob[i] = __fdiv_rn(tile[r * SP + t] - mn, den);
  }
}

}  // namespace bn

using namespace bn;

struct bn_features {
  int device = 0;
  bn_feat_params p{};
  int hop = 0, bins = 0, frames = 0, Wk = 0, rows = 0;
  float* d_basis = nullptr;
  int2* d_band = nullptr;
  float* d_dct = nullptr;
  float* d_mags = nullptr;
  unsigned* d_mnmx = nullptr;
  int16_t* d_pcm = nullptr;      // staging for host callers (one sub-wave)
  float* d_peak = nullptr;
  float* d_out = nullptr;
  cudaStream_t own = nullptr;
  size_t smem = 0;
  FeatDev dev{};
};

static bool feat_is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

#define FCU(call)                                                                                  \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      char _m[384];                                                                                \
      snprintf(_m, sizeof _m, "%s: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return set_error(BN_ERR_CUDA, _m);                                                           \
    }                                                                                              \
  } while (0)

extern "C" void bn_features_destroy(bn_features* f) {
  if (!f) return;
  cudaSetDevice(f->device);
  cudaFree(f->d_basis); cudaFree(f->d_band); cudaFree(f->d_dct); cudaFree(f->d_mags); cudaFree(f->d_mnmx);
  cudaFree(f->d_pcm); cudaFree(f->d_peak); cudaFree(f->d_out);
  if (f->own) cudaStreamDestroy(f->own);
  delete f;
}

extern "C" int bn_features_create(const bn_feat_params* p, const float* mel_basis, const float* dct, int device, bn_features** out) {
  if (!p || !mel_basis || !out) return set_error(BN_ERR_ARG, "bad arguments");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev)
    return set_error(BN_ERR_CUDA, "no CUDA device for the feature kernels (there is no CPU fallback)");
  if (p->n_fft != 512) return set_error(BN_ERR_UNSUPPORTED, "n_fft must be 512");
  if (p->spec_width <= 0 || p->spec_width % 32 || p->chunk_len < p->spec_width) return set_error(BN_ERR_UNSUPPORTED, "spec_width must be a positive multiple of 32");
  if (p->n_mels <= 0 || p->n_mels > 128) return set_error(BN_ERR_UNSUPPORTED, "n_mels must be in 1..128");
  if (p->mode < BN_FEAT_MEL || p->mode > BN_FEAT_MFCC || p->mag_scale < BN_MAG_NONE || p->mag_scale > BN_MAG_DB) return set_error(BN_ERR_ARG, "unknown mode / mag_scale");
  if (p->mode == BN_FEAT_MFCC && (!dct || p->n_mfcc <= 0 || p->n_mfcc > p->n_mels)) return set_error(BN_ERR_ARG, "MFCC needs a DCT matrix and 0 < n_mfcc <= n_mels");
  FCU(cudaSetDevice(device));
  bn_features* f = new bn_features();
  f->device = device;
  f->p = *p;
  f->hop = p->chunk_len / p->spec_width;
  f->bins = p->n_fft / 2 + 1;
  f->frames = 1 + p->chunk_len / f->hop;                 // librosa.stft(center=True) frame count
  const int wv = p->mode == BN_FEAT_MFCC ? f->frames : p->spec_width;
  f->Wk = (wv + 31) / 32 * 32;
  f->rows = p->mode == BN_FEAT_MFCC ? p->n_mfcc : p->n_mels;
  // band limits of every filter
  std::vector<int2> band(p->n_mels);
  for (int m = 0; m < p->n_mels; m++) {
    int lo = 0, hi = 0;
    bool any = false;
    for (int k = 0; k < f->bins; k++)
      if (mel_basis[(size_t)m * f->bins + k] != 0.0f) { if (!any) lo = k; hi = k + 1; any = true; }
    band[m] = make_int2(lo, hi);
  }
#define FALLOC(ptr, bytes) do { if (cudaMalloc((void**)&(ptr), (bytes)) != cudaSuccess) { bn_features_destroy(f); return set_error(BN_ERR_CUDA, "cudaMalloc failed in bn_features_create"); } } while (0)
  FALLOC(f->d_basis, sizeof(float) * (size_t)p->n_mels * f->bins);
  FALLOC(f->d_band, sizeof(int2) * (size_t)p->n_mels);
  cudaMemcpy(f->d_basis, mel_basis, sizeof(float) * (size_t)p->n_mels * f->bins, cudaMemcpyHostToDevice);
  cudaMemcpy(f->d_band, band.data(), sizeof(int2) * (size_t)p->n_mels, cudaMemcpyHostToDevice);
  if (p->mode == BN_FEAT_MFCC) {
    FALLOC(f->d_dct, sizeof(float) * (size_t)p->n_mfcc * p->n_mels);
    cudaMemcpy(f->d_dct, dct, sizeof(float) * (size_t)p->n_mfcc * p->n_mels, cudaMemcpyHostToDevice);
  }
  FALLOC(f->d_mags, sizeof(float) * (size_t)KF_SUBWAVE * f->Wk * KF_LDK);
  FALLOC(f->d_mnmx, sizeof(unsigned) * 2 * KF_SUBWAVE);
  if (cudaStreamCreateWithFlags(&f->own, cudaStreamNonBlocking) != cudaSuccess) { bn_features_destroy(f); return set_error(BN_ERR_CUDA, "cudaStreamCreate failed"); }
  FeatDev& D = f->dev;
  D.basis = f->d_basis; D.band = f->d_band; D.dct = f->d_dct;
  D.bins = f->bins; D.n_mels = p->n_mels; D.W = p->spec_width; D.Wk = f->Wk; D.WV = wv; D.SP = wv + 1;
  D.mode = p->mode; D.mag_scale = p->mag_scale; D.n_mfcc = p->n_mfcc; D.rows = f->rows;
  D.pcen_b = p->pcen_b;
  D.aux_floats = p->mode == BN_FEAT_MFCC ? p->n_mfcc * (p->spec_width + 1)
                 : (p->mode == BN_FEAT_MEL && p->mag_scale == BN_MAG_PCEN ? p->n_mels * wv : 0);
  f->smem = sizeof(float) * ((size_t)p->n_mels * D.SP + (size_t)KF_FG * KF_LDK + (size_t)D.aux_floats + 32);
  if (f->smem > 227 * 1024) { bn_features_destroy(f); return set_error(BN_ERR_UNSUPPORTED, "feature tile does not fit in shared memory"); }
  if (cudaFuncSetAttribute(k_feat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f->smem) != cudaSuccess) {
    bn_features_destroy(f);
    return set_error(BN_ERR_CUDA, "cudaFuncSetAttribute(k_feat) failed");
  }
  *out = f;
  return BN_OK;
}

extern "C" int bn_features_rows(const bn_features* f) { return f ? f->rows : 0; }

extern "C" int bn_features_pcm16(bn_features* f, const int16_t* pcm, const float* peak, int B, float* out, void* stream) {
  if (!f || !pcm || !out || B < 0) return set_error(BN_ERR_ARG, "bad arguments");
  if (B == 0) return BN_OK;
  FCU(cudaSetDevice(f->device));
  const bool dev_in = feat_is_device_ptr(pcm), dev_out = feat_is_device_ptr(out);
  if (dev_in != dev_out) return set_error(BN_ERR_ARG, "input and output must both be host or both be device pointers");
  if (peak && feat_is_device_ptr(peak) != dev_in) return set_error(BN_ERR_ARG, "peak must live where pcm lives");
  const int T = f->p.chunk_len;
  const size_t out_per = (size_t)f->rows * f->p.spec_width;
  cudaStream_t st = dev_in ? (cudaStream_t)stream : f->own;
  if (!dev_in && !f->d_pcm) {
    FCU(cudaMalloc((void**)&f->d_pcm, sizeof(int16_t) * (size_t)KF_SUBWAVE * T));
    FCU(cudaMalloc((void**)&f->d_peak, sizeof(float) * KF_SUBWAVE));
    FCU(cudaMalloc((void**)&f->d_out, sizeof(float) * (size_t)KF_SUBWAVE * out_per));
  }
  for (int b0 = 0; b0 < B; b0 += KF_SUBWAVE) {
    const int nb = B - b0 < KF_SUBWAVE ? B - b0 : KF_SUBWAVE;
    const int16_t* d_pcm = pcm + (size_t)b0 * T;
    const float* d_peak = peak ? peak + b0 : nullptr;
    float* d_out = out + (size_t)b0 * out_per;
    if (!dev_in) {
      FCU(cudaMemcpyAsync(f->d_pcm, d_pcm, sizeof(int16_t) * (size_t)nb * T, cudaMemcpyHostToDevice, st));
      if (peak) FCU(cudaMemcpyAsync(f->d_peak, d_peak, sizeof(float) * nb, cudaMemcpyHostToDevice, st));
      d_pcm = f->d_pcm;
      d_peak = peak ? f->d_peak : nullptr;
      d_out = f->d_out;
    }
    int rc = launch_stft_mag_fm(d_pcm, 0, d_peak, f->d_mags, f->d_mnmx, nb, T, f->p.n_fft, f->hop, f->Wk, KF_LDK, st);
    if (rc) return set_error(rc, "STFT kernel launch failed");
    k_feat<<<nb, KF_THREADS, f->smem, st>>>(f->d_mags, d_out, f->dev);
    FCU(cudaGetLastError());
    if (!dev_in) FCU(cudaMemcpyAsync(out + (size_t)b0 * out_per, d_out, sizeof(float) * (size_t)nb * out_per, cudaMemcpyDeviceToHost, st));
  }
  if (!dev_in) FCU(cudaStreamSynchronize(st));
  return BN_OK;
}
