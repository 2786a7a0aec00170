// bn_reader.cu -- native, multi-threaded WAV / FLAC reader feeding the PCM16 batch buffer (host code only; compiled by nvcc with
// the rest of the library, no kernels).
//
// Reference: the per-file host work of evaluate() before any inference -- sf.info / SoundFile.read
// (birdnet_stm32/audio/io.py:90-116), the window peak (io.py:122), split_audio_into_chunks (io.py:133-174) and the
// per-file Python loop (evaluation/metrics.py:117-147).  With inference at about a microsecond per chunk this is what
// bounds an evaluation on real files, so it runs here on a pool of native threads: phase 1 reads the RIFF headers and
// fixes every file's chunk count and position in the batch buffer, phase 2 reads the sample data and writes the chunks.
#include "../../include/bn_reader.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/bn_ingest.h"
#include "bn_flac.h"
#include "bn_kernels.cuh"

namespace {

inline uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
inline uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | p[1] << 8); }

bool pread_all(int fd, void* dst, size_t n, off_t off) {
  unsigned char* d = (unsigned char*)dst;
  while (n) {
    const ssize_t r = pread(fd, d, n, off);
    if (r <= 0) return false;
    d += r; n -= (size_t)r; off += r;
  }
  return true;
}

bool read_whole(int fd, std::vector<unsigned char>& buf) {
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size <= 0) return false;
  buf.resize((size_t)st.st_size);
  return pread_all(fd, buf.data(), buf.size(), 0);
}

// FLAC: STREAMINFO gives rate / channels / sample size / frame count; the samples are decoded in phase 2 (bn_flac.h).
// Samples are delivered left-justified in int16 (<= 16 bits) or int32 containers = BN_SF_S16 / BN_SF_S32.
bool probe_flac(int fd, off_t file_size, double max_seconds, bn_reader_file* o) {
  unsigned char h[10];
  if (!pread_all(fd, h, 10, 0)) return false;
  off_t pos = 0;
  if (memcmp(h, "ID3", 3) == 0) {
    const off_t sz = ((off_t)(h[6] & 0x7f) << 21) | ((off_t)(h[7] & 0x7f) << 14) | ((off_t)(h[8] & 0x7f) << 7) | (off_t)(h[9] & 0x7f);
    pos = 10 + sz + ((h[5] & 0x10) ? 10 : 0);
  }
  unsigned char m[42];
  if (pos + 42 > file_size || !pread_all(fd, m, 42, pos)) return false;
  bnflac::Info info;
  std::string err;
  unsigned char last_flagged[42];
  memcpy(last_flagged, m, 42);
  last_flagged[4] |= 0x80;                              // parse just this block
  if (!bnflac::parse_header(last_flagged, 42, info, err)) return false;
  int64_t frames = (int64_t)info.total;
  if (frames == 0) {                                    // length not recorded: count by decoding
    std::vector<unsigned char> all;
    std::vector<int16_t> o16;
    std::vector<int32_t> o32;
    if (!read_whole(fd, all)) return false;
    frames = bnflac::decode(all.data(), all.size(), 0, info, &o16, &o32, err);
    if (frames < 0) return false;
  }
  o->channels = info.channels;
  o->sample_rate = info.sample_rate;
  o->fmt = info.bps <= 16 ? BN_SF_S16 : BN_SF_S32;
  if (max_seconds > 0) {
    const int64_t lim = (int64_t)(max_seconds * (double)info.sample_rate);
    if (frames > lim) frames = lim;
  }
  o->n_frames = frames;
  o->data_offset = 0;
  o->container = BN_CT_FLAC;
  o->status = BN_RD_NEEDS_INGEST;
  return true;
}

// Decode the first o.n_frames frames of a FLAC file; exactly one of out16 / out32 is filled according to o.fmt.
bool decode_flac_fd(int fd, const bn_reader_file& o, std::vector<int16_t>& out16, std::vector<int32_t>& out32) {
  std::vector<unsigned char> all;
  if (!read_whole(fd, all)) return false;
  bnflac::Info info;
  std::string err;
  const int64_t got = bnflac::decode(all.data(), all.size(), o.n_frames, info, &out16, &out32, err);
  return got == o.n_frames && info.channels == o.channels;
}

// RIFF/WAVE header walk: fmt chunk (PCM, IEEE float, or WAVE_FORMAT_EXTENSIBLE with one of them as sub-format) and data chunk
bool probe_fd(int fd, double max_seconds, bn_reader_file* o) {
  memset(o, 0, sizeof *o);
  o->status = BN_RD_UNREADABLE;
  o->fmt = -1;
  struct stat st;
  if (fstat(fd, &st) != 0) return false;
  unsigned char h[12];
  if (!pread_all(fd, h, 12, 0)) return false;
  if (memcmp(h, "fLaC", 4) == 0 || memcmp(h, "ID3", 3) == 0) return probe_flac(fd, st.st_size, max_seconds, o);
  if (memcmp(h, "RIFF", 4) != 0 || memcmp(h + 8, "WAVE", 4) != 0) return false;
  off_t pos = 12;
  int tag = 0, bits = 0;
  bool have_fmt = false;
  for (;;) {
    unsigned char ch[8];
    if (pos + 8 > st.st_size || !pread_all(fd, ch, 8, pos)) return false;
    const uint32_t size = rd32(ch + 4);
    if (memcmp(ch, "fmt ", 4) == 0) {
      unsigned char f[40] = {0};
      const size_t want = size < sizeof f ? size : sizeof f;
      if (want < 16 || !pread_all(fd, f, want, pos + 8)) return false;
      tag = rd16(f); o->channels = rd16(f + 2); o->sample_rate = (int32_t)rd32(f + 4); bits = rd16(f + 14);
      if (tag == 0xFFFE && want >= 26) tag = rd16(f + 24);
      have_fmt = true;
    } else if (memcmp(ch, "data", 4) == 0) {
      if (!have_fmt || o->channels < 1 || o->sample_rate <= 0) return false;
      if (tag == 1 && bits == 8) o->fmt = BN_SF_U8;
      else if (tag == 1 && bits == 16) o->fmt = BN_SF_S16;
      else if (tag == 1 && bits == 24) o->fmt = BN_SF_S24;
      else if (tag == 1 && bits == 32) o->fmt = BN_SF_S32;
      else if (tag == 3 && bits == 32) o->fmt = BN_SF_F32;
      else return false;
      const int64_t bpf = (int64_t)(bits / 8) * o->channels;
      int64_t avail = (int64_t)st.st_size - (pos + 8);
      if ((int64_t)size < avail) avail = size;
      int64_t frames = avail / bpf;
      if (max_seconds > 0) {
        const int64_t lim = (int64_t)(max_seconds * (double)o->sample_rate);
        if (frames > lim) frames = lim;
      }
      o->n_frames = frames;
      o->data_offset = pos + 8;
      o->status = BN_RD_NEEDS_INGEST;
      return true;
    }
    pos += 8 + (off_t)size + (size & 1);
  }
}

int chunks_for(int64_t n, int chunk_len, int step) {
  if (n <= 0) return 0;
  if (n <= chunk_len) return 1;
  const int64_t n_full = 1 + (n - chunk_len) / step;
  return (int)(n_full + ((n - chunk_len) % step != 0 ? 1 : 0));
}

template <typename F>
void parallel_for(int n, int threads, F fn) {
  if (threads < 1) threads = 1;
  if (threads > n) threads = n;
  if (threads <= 1) { for (int i = 0; i < n; i++) fn(i); return; }
  std::atomic<int> next(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back([&] { for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i); });
  for (auto& th : pool) th.join();
}

}  // namespace

extern "C" int bn_wav_probe(const char* path, double max_seconds, bn_reader_file* out) {
  if (!path || !out) return bn::set_error(BN_ERR_ARG, "bn_wav_probe: NULL argument");
  const int fd = open(path, O_RDONLY | O_CLOEXEC);
  if (fd < 0) { memset(out, 0, sizeof *out); out->status = BN_RD_UNREADABLE; out->fmt = -1; return BN_OK; }
  probe_fd(fd, max_seconds, out);
  close(fd);
  return BN_OK;
}

extern "C" int bn_read_pcm16_batch(const char* const* paths, int n_paths, int sample_rate, int chunk_len, int step, double max_seconds,
                                   int16_t* chunks, int cap_chunks, int threads, bn_reader_file* files_out, int* chunks_used) {
  if (!paths || n_paths < 0 || !files_out || !chunks_used || chunk_len <= 0 || sample_rate <= 0 || (cap_chunks > 0 && !chunks))
    return bn::set_error(BN_ERR_ARG, "bn_read_pcm16_batch: bad arguments");
  if (step < 1) step = 1;
  *chunks_used = 0;
  // ---- phase 1: headers, in parallel ------------------------------------------------------------------------
  parallel_for(n_paths, threads, [&](int i) {
    bn_reader_file* o = files_out + i;
    const int fd = paths[i] ? open(paths[i], O_RDONLY | O_CLOEXEC) : -1;
    if (fd < 0) { memset(o, 0, sizeof *o); o->status = BN_RD_UNREADABLE; o->fmt = -1; return; }
    if (probe_fd(fd, max_seconds, o) && o->fmt == BN_SF_S16 && o->channels == 1 && o->sample_rate == sample_rate) {
      o->n_chunks = chunks_for(o->n_frames, chunk_len, step);
      o->status = o->n_chunks > 0 ? BN_RD_OK : BN_RD_UNREADABLE;          // an empty window is skipped like an unreadable file
    }
    close(fd);
  });
  // ---- positions in the batch buffer: files in order until one does not fit --------------------------------------
  std::vector<int> first(n_paths + 1, 0);
  int consumed = 0, used = 0;
  for (; consumed < n_paths; consumed++) {
    const int nc = files_out[consumed].status == BN_RD_OK ? files_out[consumed].n_chunks : 0;
    if (used + nc > cap_chunks) break;
    first[consumed] = used;
    used += nc;
  }
  // ---- phase 2: sample data -> chunks, in parallel ------------------------------------------------------------------
  std::atomic<int> failed(0);
  parallel_for(consumed, threads, [&](int i) {
    bn_reader_file* o = files_out + i;
    if (o->status != BN_RD_OK) return;
    const int64_t n = o->n_frames;
    int16_t* dst = chunks + (size_t)first[i] * chunk_len;
    const int fd = open(paths[i], O_RDONLY | O_CLOEXEC);
    bool ok = fd >= 0;
    std::vector<int16_t> tmp;
    const int16_t* src = nullptr;
    if (ok && o->container == BN_CT_FLAC) {
      std::vector<int32_t> unused;
      ok = decode_flac_fd(fd, *o, tmp, unused);
      if (ok && (n <= chunk_len || step == chunk_len)) { memcpy(dst, tmp.data(), (size_t)n * 2); src = dst; }
      else src = tmp.data();
    } else if (ok) {
      if (n <= chunk_len || step == chunk_len) {
        // back-to-back chunks: the window is read straight into place (the first n samples ARE the full chunks)
        ok = pread_all(fd, dst, (size_t)n * 2, (off_t)o->data_offset);
        src = dst;
      } else {
        tmp.resize((size_t)n);
        ok = pread_all(fd, tmp.data(), (size_t)n * 2, (off_t)o->data_offset);
        src = tmp.data();
      }
    }
    if (fd >= 0) close(fd);
    if (!ok) { o->status = BN_RD_UNREADABLE; failed.fetch_add(1); memset(dst, 0, (size_t)o->n_chunks * chunk_len * 2); return; }
    int mx = 0;
    for (int64_t k = 0; k < n; k++) { const int v = src[k] < 0 ? -(int)src[k] : (int)src[k]; mx = v > mx ? v : mx; }
    o->peak = (float)mx / 32768.0f;
    if (n <= chunk_len) {
      memset(dst + n, 0, (size_t)(chunk_len - n) * 2);                     // one right-zero-padded chunk
    } else if (step == chunk_len) {
      const int64_t n_full = 1 + (n - chunk_len) / step, rem = n - n_full * chunk_len;
      if (rem > 0) {                                                       // end-anchored tail chunk = samples [n - T, n)
        int16_t* tail = dst + (size_t)n_full * chunk_len;                   // holds samples [n_full T, n) at its start
        memmove(tail + (chunk_len - rem), tail, (size_t)rem * 2);
        memcpy(tail, tail - (chunk_len - rem), (size_t)(chunk_len - rem) * 2);   // samples [n - T, n_full T): the end of the previous chunk
      }
    } else {
      const int64_t n_full = 1 + (n - chunk_len) / step;
      for (int64_t c = 0; c < n_full; c++) memcpy(dst + (size_t)c * chunk_len, src + c * step, (size_t)chunk_len * 2);
      if ((n - chunk_len) % step != 0) memcpy(dst + (size_t)n_full * chunk_len, src + (n - chunk_len), (size_t)chunk_len * 2);
    }
  });
  // a file that failed in phase 2 keeps its (zeroed) slot; the caller drops it through status
  *chunks_used = used;
  return consumed;
}

extern "C" int bn_read_raw_batch(const char* const* paths, int n_paths, double max_seconds, void* dst, int64_t cap_bytes, int threads,
                                 bn_reader_file* files_out, int64_t* byte_offsets) {
  if (!paths || n_paths < 0 || !files_out || !byte_offsets || cap_bytes < 0 || (cap_bytes > 0 && !dst))
    return bn::set_error(BN_ERR_ARG, "bn_read_raw_batch: bad arguments");
  static const int bytes_per_sample[5] = {2, 3, 4, 4, 1};              // BN_SF_S16, S24, S32, F32, U8
  parallel_for(n_paths, threads, [&](int i) {
    bn_reader_file* o = files_out + i;
    const int fd = paths[i] ? open(paths[i], O_RDONLY | O_CLOEXEC) : -1;
    if (fd < 0) { memset(o, 0, sizeof *o); o->status = BN_RD_UNREADABLE; o->fmt = -1; return; }
    probe_fd(fd, max_seconds, o);
    if (o->status != BN_RD_UNREADABLE && o->n_frames <= 0) o->status = BN_RD_UNREADABLE;
    close(fd);
  });
  int consumed = 0;
  int64_t used = 0;
  for (; consumed < n_paths; consumed++) {
    const bn_reader_file& o = files_out[consumed];
    int64_t nb = 0;
    if (o.status != BN_RD_UNREADABLE) nb = o.n_frames * o.channels * bytes_per_sample[o.fmt];
    const int64_t slot = (nb + 15) & ~(int64_t)15;
    if (used + slot > cap_bytes) break;
    byte_offsets[consumed] = used;
    used += slot;
  }
  byte_offsets[consumed] = used;
  parallel_for(consumed, threads, [&](int i) {
    bn_reader_file* o = files_out + i;
    if (o->status == BN_RD_UNREADABLE) return;
    const int64_t nb = o->n_frames * o->channels * bytes_per_sample[o->fmt];
    const int fd = open(paths[i], O_RDONLY | O_CLOEXEC);
    bool ok = fd >= 0;
    if (ok && o->container == BN_CT_FLAC) {
      std::vector<int16_t> o16;
      std::vector<int32_t> o32;
      ok = decode_flac_fd(fd, *o, o16, o32);
      if (ok) memcpy((unsigned char*)dst + byte_offsets[i], o->fmt == BN_SF_S16 ? (const void*)o16.data() : (const void*)o32.data(), (size_t)nb);
    } else if (ok) {
      ok = pread_all(fd, (unsigned char*)dst + byte_offsets[i], (size_t)nb, (off_t)o->data_offset);
    }
    if (fd >= 0) close(fd);
    if (!ok) o->status = BN_RD_UNREADABLE;
  });
  return consumed;
}
