// bn_flac.h -- FLAC stream decoder (host C++, header only) for the native file reader (bn_reader.cu).
//
// The reference reads every container libsndfile reads (birdnet_stm32/audio/io.py:90-116: sf.info + SoundFile.read,
// float32 = integer sample / 2^(bits-1)); FLAC is the lossless one bird-sound archives use.  libsndfile / libFLAC are
// not in this image, so the bit-stream format is decoded here from its public specification (RFC 9639): STREAMINFO,
// frame headers (CRC-8), CONSTANT / VERBATIM / FIXED / LPC subframes, partitioned Rice residuals (4- and 5-bit
// parameters, escape partitions), wasted bits, left-side / side-right / mid-side stereo, frame CRC-16.
//
// Output = interleaved integer samples left-justified in an int16 (bits <= 16) or int32 (bits > 16) container, i.e. the
// sample formats BN_SF_S16 / BN_SF_S32 the rest of the reader and the device ingest already handle:
// (s << (16 - bits)) / 32768 == s / 2^(bits-1), which is the float libsndfile hands to the reference.
#pragma once
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

namespace bnflac {

struct Info {
  int sample_rate = 0, channels = 0, bps = 0;
  int min_block = 0, max_block = 0;
  uint64_t total = 0;       // frames in the stream (0 = unknown)
  size_t audio_off = 0;     // byte offset of the first audio frame
};

inline uint8_t crc8_update(uint8_t c, uint8_t b) {
  c ^= b;
  for (int i = 0; i < 8; i++) c = (uint8_t)((c & 0x80) ? (c << 1) ^ 0x07 : (c << 1));
  return c;
}
struct Crc16Table {
  uint16_t t[256];
  Crc16Table() {
    for (int i = 0; i < 256; i++) {
      uint16_t c = (uint16_t)(i << 8);
      for (int k = 0; k < 8; k++) c = (uint16_t)((c & 0x8000) ? (c << 1) ^ 0x8005 : (c << 1));
      t[i] = c;
    }
  }
};
inline uint16_t crc16(const uint8_t* p, size_t n) {
  static const Crc16Table T;
  uint16_t c = 0;
  for (size_t i = 0; i < n; i++) c = (uint16_t)((c << 8) ^ T.t[(c >> 8) ^ p[i]]);
  return c;
}

// MSB-first bit reader; valid bits are kept left-aligned in `acc`
struct BitReader {
  const uint8_t* p;
  size_t n, pos;
  uint64_t acc = 0;
  int nbits = 0;
  bool overrun = false;
  BitReader(const uint8_t* data, size_t len, size_t start) : p(data), n(len), pos(start) {}
  inline void refill() {
    while (nbits <= 56 && pos < n) { acc |= (uint64_t)p[pos++] << (56 - nbits); nbits += 8; }
  }
  inline uint32_t read(int k) {           // 0 <= k <= 32
    if (k == 0) return 0;
    if (nbits < k) { refill(); if (nbits < k) { overrun = true; return 0; } }
    const uint32_t v = (uint32_t)(acc >> (64 - k));
    acc <<= k;
    nbits -= k;
    return v;
  }
  inline int32_t read_signed(int k) {     // 1 <= k <= 32, two's complement
    const uint32_t v = read(k);
    return k == 32 ? (int32_t)v : (int32_t)(v << (32 - k)) >> (32 - k);
  }
  inline int64_t read_signed_wide(int k) {   // up to 33 bits (side channel of 32-bit streams)
    if (k <= 32) return read_signed(k);
    const uint64_t hi = read(k - 32), lo = read(32);
    const uint64_t v = (hi << 32) | lo;
    return (int64_t)(v << (64 - k)) >> (64 - k);
  }
  inline uint32_t read_unary() {          // number of 0 bits before the next 1 bit
    uint32_t z = 0;
    for (;;) {
      if (nbits == 0) { refill(); if (nbits == 0) { overrun = true; return z; } }
      if (acc == 0) { z += (uint32_t)nbits; nbits = 0; continue; }
      const int lz = __builtin_clzll(acc);
      z += (uint32_t)lz;
      acc = lz == 63 ? 0 : acc << (lz + 1);      // a shift by 64 is undefined
      nbits -= lz + 1;
      return z;
    }
  }
  inline void align() { const int r = nbits & 7; acc <<= r; nbits -= r; }
  inline size_t byte_pos() const { return pos - (size_t)(nbits >> 3); }   // when byte aligned
};

// STREAMINFO (and the position of the first frame).  Skips a leading ID3v2 tag.
inline bool parse_header(const uint8_t* d, size_t n, Info& info, std::string& err) {
  size_t pos = 0;
  if (n >= 10 && memcmp(d, "ID3", 3) == 0) {
    const size_t sz = ((size_t)(d[6] & 0x7f) << 21) | ((size_t)(d[7] & 0x7f) << 14) | ((size_t)(d[8] & 0x7f) << 7) | (size_t)(d[9] & 0x7f);
    pos = 10 + sz + ((d[5] & 0x10) ? 10 : 0);
  }
  if (pos + 4 > n || memcmp(d + pos, "fLaC", 4) != 0) { err = "no fLaC marker"; return false; }
  pos += 4;
  bool have_si = false;
  for (;;) {
    if (pos + 4 > n) { err = "truncated metadata"; return false; }
    const bool last = (d[pos] & 0x80) != 0;
    const int type = d[pos] & 0x7f;
    const size_t len = ((size_t)d[pos + 1] << 16) | ((size_t)d[pos + 2] << 8) | d[pos + 3];
    pos += 4;
    if (pos + len > n) { err = "truncated metadata block"; return false; }
    if (type == 0) {
      if (len < 34) { err = "short STREAMINFO"; return false; }
      const uint8_t* s = d + pos;
      info.min_block = (s[0] << 8) | s[1];
      info.max_block = (s[2] << 8) | s[3];
      info.sample_rate = (s[10] << 12) | (s[11] << 4) | (s[12] >> 4);
      info.channels = ((s[12] >> 1) & 7) + 1;
      info.bps = (((s[12] & 1) << 4) | (s[13] >> 4)) + 1;
      info.total = ((uint64_t)(s[13] & 15) << 32) | ((uint64_t)s[14] << 24) | ((uint64_t)s[15] << 16) | ((uint64_t)s[16] << 8) | s[17];
      have_si = true;
    }
    pos += len;
    if (last) break;
  }
  if (!have_si || info.sample_rate <= 0 || info.bps < 4 || info.bps > 32) { err = "bad STREAMINFO"; return false; }
  info.audio_off = pos;
  return true;
}

namespace detail {

inline bool residual(BitReader& br, int blocksize, int order, int64_t* out, std::string& err) {
  const int method = (int)br.read(2);
  if (method > 1) { err = "reserved residual coding method"; return false; }
  const int pbits = method == 0 ? 4 : 5, esc = method == 0 ? 15 : 31;
  const int porder = (int)br.read(4);
  const int nparts = 1 << porder;
  if ((blocksize >> porder) << porder != blocksize && porder > 0) { err = "partition order does not divide the block"; return false; }
  int i = order;
  for (int p = 0; p < nparts; p++) {
    int count = blocksize >> porder;
    if (p == 0) count -= order;
    if (count < 0) { err = "partition smaller than the predictor order"; return false; }
    const int param = (int)br.read(pbits);
    if (param == esc) {
      const int nb = (int)br.read(5);
      for (int k = 0; k < count; k++) out[i++] = nb ? br.read_signed(nb) : 0;
    } else {
      for (int k = 0; k < count; k++) {
        const uint32_t q = br.read_unary();
        const uint32_t u = (q << param) | br.read(param);
        out[i++] = (int64_t)(u >> 1) ^ -(int64_t)(u & 1);
      }
    }
    if (br.overrun) { err = "frame data ends inside a residual"; return false; }
  }
  return true;
}

inline bool subframe(BitReader& br, int blocksize, int bps, int64_t* s, std::string& err) {
  if (br.read(1) != 0) { err = "subframe padding bit set"; return false; }
  const int type = (int)br.read(6);
  int wasted = 0;
  if (br.read(1)) wasted = (int)br.read_unary() + 1;
  bps -= wasted;
  if (bps < 1) { err = "wasted bits exceed the sample size"; return false; }
  if (type == 0) {
    const int64_t v = br.read_signed_wide(bps);
    for (int i = 0; i < blocksize; i++) s[i] = v;
  } else if (type == 1) {
    for (int i = 0; i < blocksize; i++) s[i] = br.read_signed_wide(bps);
  } else if (type >= 8 && type <= 12) {
    const int order = type - 8;
    if (order > blocksize) { err = "fixed order exceeds the block"; return false; }
    for (int i = 0; i < order; i++) s[i] = br.read_signed_wide(bps);
    if (!residual(br, blocksize, order, s, err)) return false;
    switch (order) {
      case 1: for (int i = 1; i < blocksize; i++) s[i] += s[i - 1]; break;
      case 2: for (int i = 2; i < blocksize; i++) s[i] += 2 * s[i - 1] - s[i - 2]; break;
      case 3: for (int i = 3; i < blocksize; i++) s[i] += 3 * s[i - 1] - 3 * s[i - 2] + s[i - 3]; break;
      case 4: for (int i = 4; i < blocksize; i++) s[i] += 4 * s[i - 1] - 6 * s[i - 2] + 4 * s[i - 3] - s[i - 4]; break;
      default: break;
    }
  } else if (type >= 32) {
    const int order = (type & 31) + 1;
    if (order > blocksize) { err = "LPC order exceeds the block"; return false; }
    for (int i = 0; i < order; i++) s[i] = br.read_signed_wide(bps);
    const int prec = (int)br.read(4) + 1;
    if (prec == 16) { err = "invalid LPC precision"; return false; }
    const int shift = br.read_signed(5);
    if (shift < 0) { err = "negative LPC shift"; return false; }
    int32_t coef[32];
    for (int j = 0; j < order; j++) coef[j] = br.read_signed(prec);
    if (!residual(br, blocksize, order, s, err)) return false;
    for (int i = order; i < blocksize; i++) {
      int64_t acc = 0;
      for (int j = 0; j < order; j++) acc += (int64_t)coef[j] * s[i - 1 - j];
      s[i] += acc >> shift;
    }
  } else {
    err = "reserved subframe type";
    return false;
  }
  if (wasted) for (int i = 0; i < blocksize; i++) s[i] *= ((int64_t)1 << wasted);
  if (br.overrun) { err = "frame data ends inside a subframe"; return false; }
  return true;
}

}  // namespace detail

// Decodes up to `max_frames` sample frames (0 = all).  `out16` receives the samples when info.bps <= 16, `out32` otherwise
// (interleaved, left-justified in the container).  Returns the number of frames decoded or -1 (err is set).
inline int64_t decode(const uint8_t* d, size_t n, int64_t max_frames, Info& info, std::vector<int16_t>* out16, std::vector<int32_t>* out32,
                      std::string& err) {
  if (!parse_header(d, n, info, err)) return -1;
  const int C = info.channels;
  const bool wide = info.bps > 16;
  const int lj = wide ? 32 - info.bps : 16 - info.bps;       // left-justify shift
  size_t pos = info.audio_off;
  int64_t done = 0;
  std::vector<int64_t> buf;
  while (pos + 6 <= n && (max_frames <= 0 || done < max_frames)) {
    if (d[pos] != 0xFF || (d[pos + 1] & 0xFE) != 0xF8) {
      if (done > 0 && info.total && (uint64_t)done >= info.total) break;   // trailing bytes after the last frame
      err = "lost frame sync";
      return -1;
    }
    const size_t fstart = pos;
    const int bs_code = d[pos + 2] >> 4, sr_code = d[pos + 2] & 15;
    const int ch_code = d[pos + 3] >> 4, ss_code = (d[pos + 3] >> 1) & 7;
    if (d[pos + 3] & 1) { err = "reserved bit in the frame header"; return -1; }
    size_t q = pos + 4;
    {  // UTF-8-like coded frame / sample number
      if (q >= n) { err = "truncated frame header"; return -1; }
      const uint8_t b = d[q];
      int extra = 0;
      if (b >= 0xFE) extra = 6; else if (b >= 0xFC) extra = 5; else if (b >= 0xF8) extra = 4; else if (b >= 0xF0) extra = 3;
      else if (b >= 0xE0) extra = 2; else if (b >= 0xC0) extra = 1; else if (b >= 0x80) { err = "bad coded number"; return -1; }
      q += 1 + extra;
    }
    int blocksize = 0;
    if (bs_code == 0) { err = "reserved block size code"; return -1; }
    else if (bs_code == 1) blocksize = 192;
    else if (bs_code <= 5) blocksize = 576 << (bs_code - 2);
    else if (bs_code == 6) { if (q + 1 > n) { err = "truncated frame header"; return -1; } blocksize = d[q] + 1; q += 1; }
    else if (bs_code == 7) { if (q + 2 > n) { err = "truncated frame header"; return -1; } blocksize = ((d[q] << 8) | d[q + 1]) + 1; q += 2; }
    else blocksize = 256 << (bs_code - 8);
    if (sr_code == 12) q += 1; else if (sr_code == 13 || sr_code == 14) q += 2; else if (sr_code == 15) { err = "invalid sample rate code"; return -1; }
    if (q + 1 > n) { err = "truncated frame header"; return -1; }
    uint8_t c8 = 0;
    for (size_t i = fstart; i < q; i++) c8 = crc8_update(c8, d[i]);
    if (c8 != d[q]) { err = "frame header CRC-8 mismatch"; return -1; }
    q += 1;
    int bps = info.bps;
    static const int ss_bits[8] = {0, 8, 12, -1, 16, 20, 24, 32};
    if (ss_code) { if (ss_bits[ss_code] < 0) { err = "reserved sample size code"; return -1; } bps = ss_bits[ss_code]; }
    if (bps != info.bps) { err = "sample size changes inside the stream"; return -1; }
    int nch = 0;
    if (ch_code < 8) nch = ch_code + 1; else if (ch_code <= 10) nch = 2; else { err = "reserved channel assignment"; return -1; }
    if (nch != C) { err = "channel count changes inside the stream"; return -1; }
    buf.resize((size_t)blocksize * C);
    BitReader br(d, n, q);
    for (int c = 0; c < C; c++) {
      int b = bps;
      if ((ch_code == 8 && c == 1) || (ch_code == 9 && c == 0) || (ch_code == 10 && c == 1)) b += 1;   // the side channel
      if (!detail::subframe(br, blocksize, b, buf.data() + (size_t)c * blocksize, err)) return -1;
    }
    br.align();
    const size_t crc_at = br.byte_pos();
    if (crc_at + 2 > n) { err = "truncated frame footer"; return -1; }
    if (crc16(d + fstart, crc_at - fstart) != (uint16_t)((d[crc_at] << 8) | d[crc_at + 1])) { err = "frame CRC-16 mismatch"; return -1; }
    pos = crc_at + 2;
    int64_t* a = buf.data();
    int64_t* bch = buf.data() + blocksize;
    if (ch_code == 8) { for (int i = 0; i < blocksize; i++) bch[i] = a[i] - bch[i]; }
    else if (ch_code == 9) { for (int i = 0; i < blocksize; i++) a[i] = a[i] + bch[i]; }
    else if (ch_code == 10) {
      for (int i = 0; i < blocksize; i++) {
        const int64_t side = bch[i];
        const int64_t mid = (a[i] * 2) | (side & 1);
        a[i] = (mid + side) >> 1;
        bch[i] = (mid - side) >> 1;
      }
    }
    int64_t take = blocksize;
    if (max_frames > 0 && done + take > max_frames) take = max_frames - done;
    if (wide) {
      const size_t base = out32->size();
      out32->resize(base + (size_t)take * C);
      int32_t* o = out32->data() + base;
      for (int64_t i = 0; i < take; i++) for (int c = 0; c < C; c++) o[i * C + c] = (int32_t)((uint32_t)buf[(size_t)c * blocksize + i] << lj);
    } else {
      const size_t base = out16->size();
      out16->resize(base + (size_t)take * C);
      int16_t* o = out16->data() + base;
      for (int64_t i = 0; i < take; i++) for (int c = 0; c < C; c++) o[i * C + c] = (int16_t)((uint16_t)buf[(size_t)c * blocksize + i] << lj);
    }
    done += take;
  }
  return done;
}

}  // namespace bnflac
