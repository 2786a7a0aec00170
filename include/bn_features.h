/* bn_features.h -- C ABI of the precomputed-spectrogram frontends on the GPU (libbn_b200.so).
 *
 * Replaces, for batches of PCM16 chunks, the host feature extraction the reference runs per chunk with librosa:
 *
 *   reference interface                                              replaced by
 *   --------------------------------------------------------------   ----------------------
 *   get_spectrogram_from_audio(audio, sample_rate, n_fft, mel_bins,   bn_features_create +
 *     spec_width, mag_scale, mode, n_mfcc)                             bn_features_pcm16
 *     birdnet_stm32/audio/spectrogram.py:24-149
 *   make_chunks_for_file(frontend="librosa") per-chunk loop           (one call per batch)
 *     birdnet_stm32/evaluation/metrics.py:49-54
 *
 * Same conventions as bn_engine.h: plain C types, 0 or a negative bn_status, message in bn_last_error(),
 * host or device data pointers, no CPU fallback.
 */
#ifndef BN_FEATURES_H
#define BN_FEATURES_H

#include <stddef.h>
#include <stdint.h>

#include "bn_blob.h"   /* BN_MAG_NONE / BN_MAG_PWL / BN_MAG_PCEN / BN_MAG_DB */
#include "bn_engine.h"

#ifdef __cplusplus
extern "C" {
#endif

/* `mode` of get_spectrogram_from_audio (spectrogram.py:31,37-41) */
enum { BN_FEAT_MEL = 0, BN_FEAT_LOG_MEL = 1, BN_FEAT_MFCC = 2 };
/* `mag_scale` (spectrogram.py:30,43-47) uses BN_MAG_* of bn_blob.h; applied in BN_FEAT_MEL mode only, like the reference */

typedef struct bn_feat_params {
  int32_t sample_rate;
  int32_t chunk_len;     /* samples per chunk; hop = chunk_len / spec_width (spectrogram.py:61) */
  int32_t n_fft;         /* 512 */
  int32_t spec_width;    /* frames kept */
  int32_t n_mels;        /* <= 128 */
  int32_t mode;          /* BN_FEAT_* */
  int32_t mag_scale;     /* BN_MAG_*  */
  int32_t n_mfcc;        /* rows kept in BN_FEAT_MFCC mode */
  float pcen_b;          /* librosa.pcen smoothing coefficient for (sample_rate, hop); see audio/mel.py */
  int32_t reserved[7];
} bn_feat_params;

typedef struct bn_features bn_features;

/* mel_basis: float32 [n_mels, n_fft/2+1] (host), e.g. the Slaney filterbank of audio/mel.py;
 * dct: float32 [n_mfcc, n_mels] orthonormal DCT-II rows (host), required for BN_FEAT_MFCC, else NULL. */
BN_API int bn_features_create(const bn_feat_params* p, const float* mel_basis, const float* dct, int device,
                              bn_features** out);
BN_API void bn_features_destroy(bn_features* f);
/* rows of the output: n_mels, or n_mfcc in BN_FEAT_MFCC mode */
BN_API int bn_features_rows(const bn_features* f);
/* pcm: int16 [B, chunk_len]; peak: float32 [B] file-level max|y| (audio/io.py:124-126), NULL or <= 0 = none;
 * out: float32 [B, rows, spec_width], every chunk min-max normalised to [0, 1] (spectrogram.py:12-21,149). */
BN_API int bn_features_pcm16(bn_features* f, const int16_t* pcm, const float* peak, int B, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BN_FEATURES_H */
