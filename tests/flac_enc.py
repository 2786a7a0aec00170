"""A small FLAC ENCODER for the tests (test infrastructure; numpy only).

No FLAC library or encoder binary exists in this image, so the decoder of the native reader (`csrc/bn_flac.h`) is tested
against streams written here from the format specification (RFC 9639), by code that shares nothing with the decoder
(different language, bit WRITER instead of reader, vectorised Rice coding).  The encoder deliberately exercises every
feature the decoder implements, cycling through them block by block:

  subframes  CONSTANT (silent blocks), VERBATIM, FIXED order 0..4, LPC (orders 1..12, quantised least-squares predictors)
  residuals  Rice method 0 (4-bit parameters) and 1 (5-bit), partition orders 0..4, escape partitions (raw bits)
  stereo     independent, left/side, side/right, mid/side
  other      wasted bits, explicit 8- and 16-bit block sizes, the last short block, 8 / 12 / 16 / 20 / 24-bit samples,
             UTF-8 coded frame numbers > 127, CRC-8 / CRC-16, PADDING + VORBIS_COMMENT metadata blocks, an ID3v2 tag

It is not a good compressor; it is a conformance exerciser.
"""

from __future__ import annotations

import numpy as np


def _bits_of(value: int, n: int) -> np.ndarray:
    return np.array([(value >> (n - 1 - i)) & 1 for i in range(n)], dtype=np.uint8)


def _signed_bits(vals: np.ndarray, n: int) -> np.ndarray:
    """Two's complement, MSB first, n bits each -> flat bit array."""
    v = (np.asarray(vals, dtype=np.int64) & ((1 << n) - 1)).astype(np.uint64)
    sh = np.arange(n - 1, -1, -1, dtype=np.uint64)
    return ((v[:, None] >> sh[None, :]) & 1).astype(np.uint8).reshape(-1)


def _rice_bits(res: np.ndarray, k: int) -> np.ndarray:
    """Rice code with parameter k: zigzag, unary quotient (zeros then a one), k remainder bits."""
    r = np.asarray(res, dtype=np.int64)
    u = np.where(r >= 0, 2 * r, -2 * r - 1).astype(np.uint64)
    q = (u >> np.uint64(k)).astype(np.int64)
    lens = q + 1 + k
    ends = np.cumsum(lens)
    starts = ends - lens
    out = np.zeros(int(ends[-1]) if len(ends) else 0, dtype=np.uint8)
    out[starts + q] = 1
    for b in range(k):
        out[starts + q + 1 + b] = ((u >> np.uint64(k - 1 - b)) & np.uint64(1)).astype(np.uint8)
    return out


def _crc8(data: bytes) -> int:
    c = 0
    for b in data:
        c ^= b
        for _ in range(8):
            c = ((c << 1) ^ 0x07) & 0xFF if c & 0x80 else (c << 1) & 0xFF
    return c


def _crc16(data: bytes) -> int:
    c = 0
    for b in data:
        c ^= b << 8
        for _ in range(8):
            c = ((c << 1) ^ 0x8005) & 0xFFFF if c & 0x8000 else (c << 1) & 0xFFFF
    return c


def _utf8_number(v: int) -> bytes:
    if v < 0x80:
        return bytes([v])
    n = 2
    while v >= (1 << (5 * n + 1)):          # n bytes carry 5n + 1 bits
        n += 1
    first = ((0xFF << (8 - n)) & 0xFF) | (v >> (6 * (n - 1)))
    return bytes([first] + [0x80 | ((v >> (6 * (n - 1 - i))) & 0x3F) for i in range(1, n)])


_FIXED = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}


def _predict_residual(x: np.ndarray, coef: list[int], shift: int) -> np.ndarray:
    """x[i] - ((sum_j coef[j] * x[i-1-j]) >> shift) for i >= order (int64, arithmetic shift)."""
    order = len(coef)
    if order == 0:
        return x.copy()
    acc = np.zeros(len(x) - order, dtype=np.int64)
    for j, c in enumerate(coef):
        acc += int(c) * x[order - 1 - j:len(x) - 1 - j]
    return x[order:] - (acc >> shift)


def _lpc_coefficients(x: np.ndarray, order: int, precision: int) -> tuple[list[int], int]:
    """Least-squares predictor of the block, quantised to `precision`-bit coefficients with a non-negative shift."""
    xf = x.astype(np.float64)
    if len(x) <= 2 * order or np.all(xf == xf[0]):
        return [0] * order, 0
    rows = np.stack([xf[order - 1 - j:len(x) - 1 - j] for j in range(order)], axis=1)
    sol, *_ = np.linalg.lstsq(rows, xf[order:], rcond=None)
    cmax = max(float(np.max(np.abs(sol))), 1e-9)
    shift = int(np.clip(precision - 1 - int(np.floor(np.log2(cmax))) - 1, 0, 15))
    lim = (1 << (precision - 1)) - 1
    q = np.clip(np.round(sol * (1 << shift)), -lim - 1, lim).astype(np.int64)
    return [int(v) for v in q], shift


def _residual_section(res: np.ndarray, blocksize: int, order: int, method: int, porder: int, escape_first: bool) -> np.ndarray:
    pbits, esc = (4, 15) if method == 0 else (5, 31)
    parts = [_bits_of(method, 2), _bits_of(porder, 4)]
    n_parts = 1 << porder
    at = 0
    for p in range(n_parts):
        count = (blocksize >> porder) - (order if p == 0 else 0)
        r = res[at:at + count]
        at += count
        if count == 0:
            parts.append(_bits_of(0, pbits))
            continue
        if escape_first and p == 0:
            nb = int(max(1, int(np.max(np.abs(r))).bit_length() + 1))
            parts += [_bits_of(esc, pbits), _bits_of(nb, 5), _signed_bits(r, nb)]
            continue
        mean = float(np.mean(np.abs(r))) if count else 0.0
        k = int(np.clip(np.floor(np.log2(mean + 1.0)), 0, esc - 1))
        parts += [_bits_of(k, pbits), _rice_bits(r, k)]
    assert at == len(res)
    return np.concatenate(parts)


def _subframe(x: np.ndarray, bps: int, kind: str, arg: int, method: int, porder: int, escape_first: bool) -> np.ndarray:
    """One subframe for the int64 samples x at `bps` bits."""
    wasted = 0
    if np.any(x != 0):
        both = np.bitwise_or.reduce(x.astype(np.int64))
        while wasted < bps - 1 and not (both >> wasted) & 1:
            wasted += 1
    xs = x >> wasted
    b = bps - wasted
    n = len(xs)
    if np.all(xs == xs[0]):
        kind = "constant"
    order = arg if kind in ("fixed", "lpc") else 0
    if order >= n:
        kind, order = "verbatim", 0
    while porder > 0 and ((n >> porder) << porder != n or (n >> porder) <= order):
        porder -= 1
    head_type = {"constant": 0, "verbatim": 1}.get(kind, 8 + order if kind == "fixed" else 32 + order - 1)
    parts = [_bits_of(0, 1), _bits_of(head_type, 6)]
    if wasted:
        parts += [_bits_of(1, 1), _bits_of(1, wasted)]       # flag, then k-1 zeros and a one
    else:
        parts.append(_bits_of(0, 1))
    if kind == "constant":
        parts.append(_signed_bits(xs[:1], b))
    elif kind == "verbatim":
        parts.append(_signed_bits(xs, b))
    elif kind == "fixed":
        parts.append(_signed_bits(xs[:order], b) if order else np.zeros(0, np.uint8))
        res = _predict_residual(xs, _FIXED[order], 0)
        parts.append(_residual_section(res, n, order, method, porder, escape_first))
    else:
        precision = 12
        coef, shift = _lpc_coefficients(xs, order, precision)
        parts += [_signed_bits(xs[:order], b), _bits_of(precision - 1, 4), _bits_of(shift & 31, 5), _signed_bits(np.array(coef), precision)]
        res = _predict_residual(xs, coef, shift)
        parts.append(_residual_section(res, n, order, method, porder, escape_first))
    return np.concatenate(parts)


_SS_CODE = {8: 1, 12: 2, 16: 4, 20: 5, 24: 6, 32: 7}
_PLAN = [("verbatim", 0), ("fixed", 0), ("fixed", 1), ("fixed", 2), ("lpc", 8), ("fixed", 3), ("fixed", 4), ("lpc", 2),
         ("lpc", 12), ("lpc", 1), ("lpc", 5)]


def encode(samples: np.ndarray, sample_rate: int, bps: int = 16, blocksize: int = 4096, first_frame_number: int = 0,
           id3: bool = False, explicit_sample_size: bool = True) -> bytes:
    """samples: integer array [frames] or [frames, channels] with values in the `bps`-bit signed range."""
    x = np.asarray(samples)
    if x.ndim == 1:
        x = x[:, None]
    x = x.astype(np.int64)
    n, C = x.shape
    assert 1 <= C <= 8 and n > 0
    frames = []
    fno = first_frame_number
    sizes = []
    for bi, start in enumerate(range(0, n, blocksize)):
        blk = x[start:start + blocksize]
        bs = len(blk)
        stereo_mode = bi % 4 if C == 2 else 0          # 0 independent, 1 left/side, 2 side/right, 3 mid/side
        chans, widths = [], []
        if stereo_mode == 0:
            chans, widths, ch_code = [blk[:, c] for c in range(C)], [bps] * C, C - 1
        else:
            L, R = blk[:, 0], blk[:, 1]
            side = L - R
            if stereo_mode == 1:
                chans, widths, ch_code = [L, side], [bps, bps + 1], 8
            elif stereo_mode == 2:
                chans, widths, ch_code = [side, R], [bps + 1, bps], 9
            else:
                chans, widths, ch_code = [(L + R) >> 1, side], [bps, bps + 1], 10
        if bs == 192:
            bs_code, bs_extra = 1, b""
        elif bs in (576, 1152, 2304, 4608):
            bs_code, bs_extra = 2 + (576, 1152, 2304, 4608).index(bs), b""
        elif bs in (256, 512, 1024, 2048, 4096, 8192, 16384, 32768):
            bs_code, bs_extra = 8 + (bs // 256).bit_length() - 1, b""
        elif bs <= 256:
            bs_code, bs_extra = 6, bytes([bs - 1])
        else:
            bs_code, bs_extra = 7, bytes([(bs - 1) >> 8, (bs - 1) & 0xFF])
        ss = _SS_CODE[bps] if explicit_sample_size and bps in _SS_CODE and bi % 2 == 0 else 0
        head = bytes([0xFF, 0xF8, (bs_code << 4) | 0, (ch_code << 4) | (ss << 1)]) + _utf8_number(fno) + bs_extra
        head += bytes([_crc8(head)])
        bits = []
        for c, (ch, w) in enumerate(zip(chans, widths)):
            kind, arg = _PLAN[(bi + 3 * c) % len(_PLAN)]
            bits.append(_subframe(ch, w, kind, arg, method=(bi + c) % 2, porder=(bi + c) % 5, escape_first=(bi % 7 == 3)))
        body = np.concatenate(bits)
        pad = (-len(body)) % 8
        if pad:
            body = np.concatenate([body, np.zeros(pad, np.uint8)])
        frame = head + np.packbits(body).tobytes()
        frame += _crc16(frame).to_bytes(2, "big")
        frames.append(frame)
        sizes.append(len(frame))
        fno += 1
    si = bytearray()
    si += blocksize.to_bytes(2, "big") * 2
    si += min(sizes).to_bytes(3, "big") + max(sizes).to_bytes(3, "big")
    packed = (sample_rate << 44) | ((C - 1) << 41) | ((bps - 1) << 36) | n
    si += packed.to_bytes(8, "big") + bytes(16)
    out = bytearray()
    if id3:
        tag = b"TIT2" + (5).to_bytes(4, "big") + b"\0\0" + b"\0test"
        sz = len(tag)
        out += b"ID3\x04\x00\x00" + bytes([(sz >> 21) & 0x7F, (sz >> 14) & 0x7F, (sz >> 7) & 0x7F, sz & 0x7F]) + tag
    out += b"fLaC" + bytes([0x00]) + len(si).to_bytes(3, "big") + si
    vendor = b"bn-test-encoder"
    vc = len(vendor).to_bytes(4, "little") + vendor + (0).to_bytes(4, "little")
    out += bytes([0x04]) + len(vc).to_bytes(3, "big") + vc
    out += bytes([0x81]) + (12).to_bytes(3, "big") + bytes(12)          # PADDING, last block
    for f in frames:
        out += f
    return bytes(out)
