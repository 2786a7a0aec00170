/* bn_ingest.h -- C ABI of the device-side audio ingest (libbn_b200.so), SURVEY section 8 row f1.
 *
 *   reference interface                                       replaced by
 *   --------------------------------------------------------  ------------------------------
 *   sf.SoundFile.read(dtype="float32", always_2d=True)         bn_ingest_window (decode step; the container
 *     birdnet_stm32/audio/io.py:112-114                          itself -- RIFF/WAVE -- is parsed on the host)
 *   y.mean(axis=1)                      audio/io.py:118        bn_ingest_window (channel mean, float32)
 *   fast_resample -> scipy.signal.resample_poly                bn_ingest_window (polyphase FIR, Kaiser beta 5,
 *     audio/io.py:14-30,119-120                                  zero extension; bn_ingest_filter = the taps)
 *   peak = max|y| ; y / peak            audio/io.py:122-124    bn_ingest_window (device reduction + IEEE division)
 *   split_audio_into_chunks             audio/io.py:133-174    bn_ingest_chunks (zero-padded short window,
 *                                                                stepped starts, end-anchored tail chunk)
 *
 * The float32 chunks these calls produce feed bn_infer_wave_f32 / bn_infer_pool_wave_f32 (bn_engine.h) without leaving
 * the device.  Conventions are those of bn_engine.h: plain C types, 0 or a negative bn_status, message in
 * bn_last_error(); `frames` and the outputs may each be HOST or DEVICE pointers; with any host pointer the call
 * synchronises before returning, otherwise it only enqueues on `stream`.  There is no CPU fallback.
 */
#ifndef BN_INGEST_H
#define BN_INGEST_H

#include <stddef.h>
#include <stdint.h>

#include "bn_engine.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bn_ingest bn_ingest;

/* sample formats of the interleaved frame buffer, decoded to float32 the way libsndfile does for dtype="float32" */
enum {
  BN_SF_S16 = 0, /* int16   -> s / 2^15                         */
  BN_SF_S24 = 1, /* packed little-endian 24-bit -> s / 2^23     */
  BN_SF_S32 = 2, /* int32   -> (float)s / 2^31                  */
  BN_SF_F32 = 3, /* float32 -> unchanged                        */
  BN_SF_U8 = 4   /* uint8   -> (s - 128) / 2^7                  */
};

BN_API int bn_ingest_create(int device, bn_ingest** out);
BN_API void bn_ingest_destroy(bn_ingest* g);

/* len(load_audio_window(...)) for n_frames read at sr_in: n_frames when the rates agree, else
 * ceil(n_frames * up / down) with up / down = sr_out / sr_in reduced (scipy.signal.resample_poly). */
BN_API int64_t bn_ingest_out_len(int64_t n_frames, int sr_in, int sr_out);

/* Number of chunks split_audio_into_chunks / estimate_num_chunks give for a window of n_samples
 * (chunk_len = int(sr * chunk_duration), step = int(sr * (chunk_duration - overlap)) >= 1). */
BN_API int bn_ingest_num_chunks(int64_t n_samples, int chunk_len, int step);

/* The anti-aliasing filter resample_poly designs for up / down (already reduced or not): float32 taps, scaled by
 * `up` and zero-padded exactly as they are handed to upfirdn.  h_out may be NULL to query *n_taps only.
 * *n_pre_remove = output samples dropped at the front. */
BN_API int bn_ingest_filter(int up, int down, float* h_out, int cap, int* n_taps, int* n_pre_remove);

/* load_audio_window on samples already read from the container: decode -> channel mean -> resample -> (normalize != 0)
 * divide by the peak when it is > 0.  wave_out: float32 [bn_ingest_out_len(n_frames, sr_in, sr_out)];
 * peak_out: optional float32 [1], max|y| after resampling. */
BN_API int bn_ingest_window(bn_ingest* g, const void* frames, int fmt, int64_t n_frames, int channels, int sr_in, int sr_out,
                            int normalize, float* wave_out, float* peak_out, void* stream);

/* bn_ingest_window(normalize = 1) followed by split_audio_into_chunks.  chunks_out: float32 [n_chunks, chunk_len] with
 * n_chunks = bn_ingest_num_chunks(bn_ingest_out_len(...), chunk_len, step) <= max_chunks (else BN_ERR_ARG);
 * *n_chunks receives the count.  n_frames == 0 gives 0 chunks. */
BN_API int bn_ingest_chunks(bn_ingest* g, const void* frames, int fmt, int64_t n_frames, int channels, int sr_in, int sr_out,
                            int chunk_len, int step, float* chunks_out, int max_chunks, int* n_chunks, float* peak_out,
                            void* stream);

/* Kernels launched by this ingest object since creation. */
BN_API int64_t bn_ingest_launch_count(const bn_ingest* g);

#ifdef __cplusplus
}
#endif
#endif /* BN_INGEST_H */
