/* bn_engine.h -- C ABI of the B200 chunk-classification engine (libbn_b200.so).
 *
 * This is the drop-in boundary for the reference's runner plugin point:
 *
 *   reference interface                                   replaced by
 *   ---------------------------------------------------   --------------------------
 *   tf.lite.Interpreter(model_path) + allocate_tensors    bn_create / bn_query
 *     birdnet_stm32/models/runners.py:57-68
 *   TFLiteRunner.predict(x_batch) (set_tensor/invoke/     bn_infer_spec_f32
 *     get_tensor), models/runners.py:82-95
 *   get_spectrogram_from_audio(linear) + normalize        bn_frontend_pcm16
 *     audio/spectrogram.py:12-21,61,106-115,133,149
 *   make_chunks_for_file + predict per <=16-chunk batch   bn_infer_pcm16
 *     evaluation/metrics.py:55-61,129-139
 *   pool_scores / lme_pooling                              bn_pool_scores, bn_infer_pool
 *     evaluation/pooling.py:6-47, metrics.py:143
 *
 * Conventions: plain C types only (no torch / numpy types); every function
 * returns 0 on success or a negative bn_status, never throws; the message of
 * the last failure on the calling thread is bn_last_error().  Data pointers may
 * be HOST or DEVICE pointers (detected with cudaPointerGetAttributes); host
 * buffers are staged through pinned memory inside the call and the call
 * returns after the results are in the caller's buffer.  With device pointers
 * the work is enqueued on `stream` (a cudaStream_t, NULL = default stream) and
 * the call returns without synchronising.  The caller owns all I/O buffers; the
 * engine owns weights and workspace.  One engine per (device, caller thread);
 * calls on one engine are not re-entrant.  There is NO CPU fallback: without a
 * CUDA device bn_create fails with BN_ERR_CUDA.
 */
#ifndef BN_ENGINE_H
#define BN_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BN_API __attribute__((visibility("default")))

typedef struct bn_engine bn_engine;

typedef enum bn_status {
  BN_OK = 0,
  BN_ERR_ARG = -1,      /* bad argument (NULL, shape mismatch, unsupported pooling ...) */
  BN_ERR_BLOB = -2,     /* malformed / unsupported blob                                 */
  BN_ERR_CUDA = -3,     /* CUDA runtime error or no device                              */
  BN_ERR_UNSUPPORTED = -4,
  BN_ERR_STATE = -5     /* e.g. bn_dump_tensor before any inference                      */
} bn_status;

/* pooling methods of evaluation/pooling.py:25-47 */
enum { BN_POOL_AVG = 0, BN_POOL_MAX = 1, BN_POOL_LME = 2 };

/* bn_set_option keys */
enum {
  BN_OPT_ROUNDING = 1,      /* 0 = gemmlowp double rounding (TFLite default build), 1 = single rounding */
  BN_OPT_MEAN_VARIANT = 2,  /* 0 auto, 1 float, 2 folded 1/N, 3 reference rounded divide    */
  BN_OPT_FORCE_GENERIC = 3, /* 1 = run one kernel per op and keep every tensor (debug taps)             */
  BN_OPT_WAVE = 4,          /* chunks processed per internal wave (workspace is sized for it)           */
  BN_OPT_PROFILE = 5,       /* 1 = time every kernel with CUDA events, 0 = off, 2 = on + reset counters  */
  BN_OPT_TENSOR_CORE = 6,   /* 1 (default) = pointwise convs on tcgen05.mma kind::i8, 0 = dp4a CUDA-core GEMM */
  BN_OPT_HOST_WAVE = 8,     /* wave size for calls whose input is in HOST memory (default 592): uploads are double-buffered under
                               compute, so a smaller wave shortens the un-overlapped first upload / last compute of a call */
  BN_OPT_FUSION = 7         /* bit mask, default 139 = 1 | 2 | 8 | 128.
                               bit 0 (1): fused kernels: depthwise + pointwise (+ADD) per DS block (bn_ds.cu); in the generic plan also
                                          the SE gate, 1x1 convolution + ADD and whole DS blocks as single launches
                               bit 1 (2): fused frontend: frame-major K1 + tensor-core head (quantise + mel mixer + PWL LUT)
                               bit 2 (4): depthwise conv of stride-1 blocks on the tensor core too (bit-exact, measured slower, off)
                               bit 3 (8): whole-stage kernel for the 8 x 16 stage (bn_stage.cu)
                               bit 4 (16): the stage kernel also writes its inner block outputs (taps for bn_dump_tensor)
                               bit 5 (32): quantising frontend K1q + K2q (bn_frontend_q.cu; 27 % less DRAM traffic, 1 % slower, off)
                               bit 6 (64): warp-specialised DS-block kernel (bn_ds_ws.cu; measured equal to bit 0's kernel, off)
                               bit 7 (128): stem convolution as an im2col GEMM on tcgen05 (bn_stem_tc.cu)
                               bit 8 (256): stem computed inside the first DS block's kernel (no stem tensor in global memory: 262 KB less DRAM
                                            traffic per chunk; measured 36 % slower for the two layers, off) */
};

typedef struct bn_info {
  int32_t frontend_kind;   /* BN_FE_* of bn_blob.h                     */
  int32_t sample_rate;
  int32_t chunk_len;       /* samples per chunk                         */
  int32_t n_fft;
  int32_t hop;
  int32_t spec_width;
  int32_t fft_bins;        /* n_fft/2 + 1                               */
  int32_t num_classes;
  int64_t input_elems;     /* floats per chunk of the graph input       */
  int32_t n_ops;
  int32_t n_tensors;
  int32_t device;
  int32_t wave;            /* current wave size                         */
  int64_t workspace_bytes; /* device workspace currently allocated      */
  int32_t fast_path;       /* 1 if the fused kernel plan is in use      */
  int32_t reserved[7];
} bn_info;

/* Load a blob (include/bn_blob.h) onto `device`, upload weights, build the kernel plan. */
BN_API int bn_create(const void* blob, size_t nbytes, int device, bn_engine** out);
BN_API void bn_destroy(bn_engine* e);
BN_API int bn_query(const bn_engine* e, bn_info* out);
BN_API int bn_set_option(bn_engine* e, int key, int value);

/* Graph only.  spec: float32 [B, input_elems] exactly as the reference runner receives it
 * (hybrid: [B,257,256,1]); scores: float32 [B, num_classes]. */
BN_API int bn_infer_spec_f32(bn_engine* e, const float* spec, int B, float* scores, void* stream);

/* Frontend only.  pcm: int16 [B, chunk_len]; peak: float32 [B] = the file-level max|y| the
 * reference divides by (audio/io.py:124-126), <= 0 or NULL = none; spec_out: float32
 * [B, input_elems] in the graph-input layout. */
BN_API int bn_frontend_pcm16(bn_engine* e, const int16_t* pcm, const float* peak, int B,
                             float* spec_out, void* stream);

/* Full hot path: PCM16 -> frontend -> int8 graph -> scores [B, num_classes]. */
BN_API int bn_infer_pcm16(bn_engine* e, const int16_t* pcm, const float* peak, int B,
                          float* scores, void* stream);

/* The same with float32 waveform chunks [B, chunk_len] -- what the reference's make_chunks_for_file holds after
 * load_audio_window resampled / mixed / peak-normalised a file (audio/io.py:118-128, evaluation/metrics.py:39-61); produced
 * on the device by bn_ingest_chunks (bn_ingest.h).  peak: optional extra divisor per chunk (NULL = samples used as they are). */
BN_API int bn_infer_wave_f32(bn_engine* e, const float* wave, const float* peak, int B, float* scores, void* stream);
BN_API int bn_frontend_wave_f32(bn_engine* e, const float* wave, const float* peak, int B, float* spec_out, void* stream);
BN_API int bn_infer_pool_wave_f32(bn_engine* e, const float* wave, const float* peak, const int32_t* file_offsets, int F,
                                  int pooling, float beta, float* file_scores, void* stream);

/* Full hot path with per-file pooling on the device.  file_offsets: int32 [F+1], chunk index
 * ranges per file (file f owns chunks [off[f], off[f+1]) ; empty files give zeros);
 * file_scores: float32 [F, num_classes].  B = file_offsets[F]. */
BN_API int bn_infer_pool(bn_engine* e, const int16_t* pcm, const float* peak,
                         const int32_t* file_offsets, int F, int pooling, float beta,
                         float* file_scores, void* stream);

/* Pooling alone: chunk_scores float32 [N, C] -> file_scores [F, C]. */
BN_API int bn_pool_scores(bn_engine* e, const float* chunk_scores, const int32_t* file_offsets, int F,
                          int C, int pooling, float beta, float* file_scores, void* stream);

/* Debug tap: copy tensor `tfl_tensor_id` of the LAST wave of the last inference to host memory
 * `out` (nbytes must equal bytes-per-chunk * chunks in that wave).  Needs BN_OPT_FORCE_GENERIC
 * for tensors the fused plan does not materialise. */
BN_API int bn_dump_tensor(bn_engine* e, int tfl_tensor_id, void* out, size_t nbytes);

/* Number of engine kernels launched by this engine since creation. */
BN_API int64_t bn_launch_count(const bn_engine* e);

/* Per-kernel device times collected while BN_OPT_PROFILE is on (CUDA events on the launching
 * stream).  Iterate index = 0.. until BN_ERR_ARG.  ms = accumulated milliseconds, count = launches. */
BN_API int bn_profile_read(bn_engine* e, int index, char* name, size_t name_cap, double* ms, int64_t* count);

/* Pinned host memory helpers for callers that want overlapped staging. */
BN_API void* bn_host_alloc(size_t nbytes);
BN_API void bn_host_free(void* p);

BN_API const char* bn_last_error(void);
BN_API const char* bn_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BN_ENGINE_H */
