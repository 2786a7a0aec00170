/* bn_reader.h -- C ABI of the native file reader (libbn_b200.so, host code only), SURVEY section 8 row f1.
 *
 *   reference interface                                              replaced by
 *   ------------------------------------------------------------   ---------------------------------------------
 *   sf.info + sf.SoundFile.read per file, serial                      bn_read_pcm16_batch: RIFF/WAVE headers and sample
 *     birdnet_stm32/audio/io.py:90-116                                 data of many files read by a pool of native threads
 *   peak = max|y| (file window)              audio/io.py:122           max|s| / 32768 per file (exact: float32 division by 2^15)
 *   split_audio_into_chunks                  audio/io.py:133-174       chunks written straight into the caller's (pinned)
 *   per-file list building in evaluate()     metrics.py:117-147        batch buffer in file order
 *
 * Scope: RIFF/WAVE and FLAC (RFC 9639, decoded by csrc/bn_flac.h).  Mono files of <= 16 bits at the model rate are turned
 * into int16 chunks ready for bn_infer_pool; every other file (other rate, several channels, 8 / 24 / 32-bit, float) is
 * reported with status BN_RD_NEEDS_INGEST so that the caller sends it through bn_ingest_chunks; anything else (MP3, OGG,
 * M4A: lossy codecs, no decoder here) is BN_RD_UNREADABLE (the reference skips unreadable files, metrics.py:125-126).
 */
#ifndef BN_READER_H
#define BN_READER_H

#include <stddef.h>
#include <stdint.h>

#include "bn_engine.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { BN_RD_OK = 0, BN_RD_NEEDS_INGEST = 1, BN_RD_UNREADABLE = 2 };
enum { BN_CT_WAV = 0, BN_CT_FLAC = 1 };   /* container */

typedef struct bn_reader_file {
  int32_t status;       /* BN_RD_* */
  int32_t n_chunks;     /* chunks written for this file (0 unless BN_RD_OK) */
  int32_t sample_rate;  /* from the header (0 if unreadable) */
  int32_t channels;
  int32_t fmt;          /* BN_SF_* of bn_ingest.h, -1 if not a supported sample format */
  float peak;           /* max|s| / 32768 over the window read (BN_RD_OK) */
  int64_t n_frames;     /* frames in the window: min(frames in file, int(max_seconds * sample_rate)) */
  int64_t data_offset;  /* byte offset of the sample data in the file (WAV) */
  int32_t container;    /* BN_CT_*: RIFF/WAVE, or FLAC (decoded by the reader; fmt is then the container the decoded samples
                           are delivered in: S16 for streams of <= 16 bits, S32 above, left-justified) */
  int32_t reserved;
} bn_reader_file;

/* Header of one file (WAV or FLAC). */
BN_API int bn_wav_probe(const char* path, double max_seconds, bn_reader_file* out);

/* Reads paths[0 ..] in order into `chunks` (int16 [cap_chunks, chunk_len], host memory, ideally from bn_host_alloc) until
 * the next BN_RD_OK file would not fit; chunk geometry as split_audio_into_chunks with step = int(sr * (duration - overlap)).
 * files_out[i] is filled for every consumed file; *chunks_used = chunks written.  Returns the number of files consumed
 * (0 .. n_paths; 0 with n_paths > 0 means the first file alone exceeds cap_chunks) or a negative bn_status. */
BN_API int bn_read_pcm16_batch(const char* const* paths, int n_paths, int sample_rate, int chunk_len, int step, double max_seconds,
                               int16_t* chunks, int cap_chunks, int threads, bn_reader_file* files_out, int* chunks_used);

/* Raw interleaved sample data of paths[0 ..] (any supported format / rate / channel count), file after file, into `dst`
 * (16-byte aligned slots) until the next file would not fit in cap_bytes: the input of bn_ingest_chunks for the files
 * bn_read_pcm16_batch flagged BN_RD_NEEDS_INGEST.  files_out[i] (header fields, n_frames of the window) and
 * byte_offsets[i] (slot start; byte_offsets[consumed] = bytes used) are filled for the consumed files; unreadable files
 * take no space.  Returns the number of files consumed or a negative bn_status. */
BN_API int bn_read_raw_batch(const char* const* paths, int n_paths, double max_seconds, void* dst, int64_t cap_bytes, int threads,
                             bn_reader_file* files_out, int64_t* byte_offsets);

#ifdef __cplusplus
}
#endif
#endif /* BN_READER_H */
