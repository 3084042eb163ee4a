"""``decode_audio``: file / file-like -> mono float32 PCM at the model's sampling rate.

The reference hands anything that is not an ndarray to ``faster_whisper.audio.decode_audio`` (PyAV / FFmpeg) before the
hot path (whisper_live/transcriber/transcriber_faster_whisper.py:820-821).  PyAV is not a dependency here; the two
containers the reference itself ships or writes are handled natively -- RIFF/WAVE PCM (what the server dumps) and FLAC
(``assets/jfk.flac``, the input of its WER test) -- and anything else raises with the container named.  The FLAC decoder
is self-checking: the MD5 of the decoded samples must equal the one in STREAMINFO.  Resampling uses a Kaiser-windowed
polyphase filter (scipy's ``resample_poly`` when scipy is importable, else the numpy version below)."""
from __future__ import annotations

import hashlib
import io
import wave
from typing import BinaryIO, Tuple, Union

import numpy as np


class _Bits:
    """MSB-first bit reader over a bytes object."""

    def __init__(self, data: bytes, pos: int = 0):
        self.d, self.p, self.acc, self.n = data, pos, 0, 0

    def read(self, k: int) -> int:
        while self.n < k:
            self.acc = (self.acc << 8) | self.d[self.p]
            self.p += 1
            self.n += 8
        self.n -= k
        v = (self.acc >> self.n) & ((1 << k) - 1)
        self.acc &= (1 << self.n) - 1
        return v

    def signed(self, k: int) -> int:
        v = self.read(k)
        return v - (1 << k) if v >> (k - 1) else v

    def unary(self) -> int:
        """zero bits before the next one bit"""
        q = 0
        while True:
            if self.n == 0:
                self.acc = self.d[self.p]
                self.p += 1
                self.n = 8
            if self.acc == 0:
                q += self.n
                self.n = 0
                continue
            lead = self.n - self.acc.bit_length()
            q += lead
            self.n -= lead + 1
            self.acc &= (1 << self.n) - 1
            return q

    def align(self) -> None:
        self.acc, self.n = 0, 0


_FIXED = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}
_BLOCK = {1: 192, 2: 576, 3: 1152, 4: 2304, 5: 4608}
_DEPTH = {1: 8, 2: 12, 4: 16, 5: 20, 6: 24}


def _residual(br: _Bits, blocksize: int, order: int) -> list:
    method = br.read(2)
    if method > 1:
        raise ValueError("FLAC: reserved residual coding method")
    pbits, esc = (4, 15) if method == 0 else (5, 31)
    porder = br.read(4)
    out = []
    for part in range(1 << porder):
        n = (blocksize >> porder) - (order if part == 0 else 0)
        k = br.read(pbits)
        if k == esc:
            raw = br.read(5)
            out.extend(br.signed(raw) if raw else 0 for _ in range(n))
            continue
        for _ in range(n):
            u = (br.unary() << k) | (br.read(k) if k else 0)
            out.append((u >> 1) ^ -(u & 1))
    return out


def _subframe(br: _Bits, blocksize: int, bps: int) -> list:
    if br.read(1):
        raise ValueError("FLAC: subframe padding bit set")
    typ = br.read(6)
    wasted = 0
    if br.read(1):
        wasted = br.unary() + 1
        bps -= wasted
    if typ == 0:
        s = [br.signed(bps)] * blocksize
    elif typ == 1:
        s = [br.signed(bps) for _ in range(blocksize)]
    elif 8 <= typ <= 12 or typ >= 32:
        if typ >= 32:
            order = (typ & 31) + 1
            s = [br.signed(bps) for _ in range(order)]
            prec = br.read(4) + 1
            shift = br.signed(5)
            if shift < 0:
                raise ValueError("FLAC: negative LPC shift")
            coef = [br.signed(prec) for _ in range(order)]
        else:
            order = typ - 8
            s = [br.signed(bps) for _ in range(order)]
            coef, shift = _FIXED[order], 0
        rc = coef[::-1]
        for r in _residual(br, blocksize, order):
            pred = 0
            for c, v in zip(rc, s[len(s) - order:] if order else ()):
                pred += c * v
            s.append(r + (pred >> shift))
    else:
        raise ValueError(f"FLAC: reserved subframe type {typ}")
    return [v << wasted for v in s] if wasted else s


def decode_flac(data: bytes) -> Tuple[np.ndarray, dict]:
    """FLAC bytes -> (int64 PCM [channels, samples], {rate, ch, bps, total}).  Raises ``ValueError`` when the stream uses a
    reserved feature or the decoded samples do not match the MD5 in STREAMINFO."""
    if data[:4] != b"fLaC":
        raise ValueError("not a FLAC stream")
    pos, info = 4, None
    while True:
        last, btype = data[pos] >> 7, data[pos] & 127
        size = int.from_bytes(data[pos + 1:pos + 4], "big")
        if btype == 0:
            b = _Bits(data, pos + 4)
            b.read(16); b.read(16); b.read(24); b.read(24)
            info = dict(rate=b.read(20), ch=b.read(3) + 1, bps=b.read(5) + 1, total=b.read(36), md5=data[pos + 22:pos + 38])
        pos += 4 + size
        if last:
            break
    if info is None:
        raise ValueError("FLAC: no STREAMINFO block")
    chans = [[] for _ in range(info["ch"])]
    done = 0
    while done < info["total"] or (info["total"] == 0 and pos < len(data)):
        br = _Bits(data, pos)
        if br.read(14) != 0x3FFE:
            raise ValueError(f"FLAC: lost frame sync at byte {pos}")
        br.read(2)
        bs_code, sr_code, ch_code, ss_code = br.read(4), br.read(4), br.read(4), br.read(3)
        br.read(1)
        first = br.read(8)                        # UTF-8 style frame / sample number
        extra = 0
        while first & 0x80:
            first = (first << 1) & 0xFF
            extra += 1
        for _ in range(max(0, extra - 1)):
            br.read(8)
        if bs_code == 6:
            blocksize = br.read(8) + 1
        elif bs_code == 7:
            blocksize = br.read(16) + 1
        elif bs_code >= 8:
            blocksize = 256 << (bs_code - 8)
        elif bs_code in _BLOCK:
            blocksize = _BLOCK[bs_code]
        else:
            raise ValueError("FLAC: reserved block size")
        if sr_code == 12:
            br.read(8)
        elif sr_code in (13, 14):
            br.read(16)
        br.read(8)                                # CRC-8 (the MD5 below covers correctness)
        bps = info["bps"] if ss_code == 0 else _DEPTH[ss_code]
        if ch_code < 8:
            subs = [_subframe(br, blocksize, bps) for _ in range(ch_code + 1)]
        elif ch_code == 8:                        # left / side
            left = _subframe(br, blocksize, bps)
            side = _subframe(br, blocksize, bps + 1)
            subs = [left, [a - b for a, b in zip(left, side)]]
        elif ch_code == 9:                        # side / right
            side = _subframe(br, blocksize, bps + 1)
            right = _subframe(br, blocksize, bps)
            subs = [[a + b for a, b in zip(side, right)], right]
        elif ch_code == 10:                       # mid / side
            mid = _subframe(br, blocksize, bps)
            side = _subframe(br, blocksize, bps + 1)
            left, right = [], []
            for a, b in zip(mid, side):
                a = (a << 1) | (b & 1)
                left.append((a + b) >> 1)
                right.append((a - b) >> 1)
            subs = [left, right]
        else:
            raise ValueError("FLAC: reserved channel assignment")
        br.align()
        br.read(16)                               # CRC-16
        pos = br.p
        for c, sub in zip(chans, subs):
            c.extend(sub)
        done += blocksize
    pcm = np.asarray(chans, dtype=np.int64)
    if info["total"]:
        pcm = pcm[:, :info["total"]]
    if any(info["md5"]):
        nbytes = (info["bps"] + 7) // 8
        inter = pcm.T.reshape(-1)
        if nbytes == 3:
            raw = b"".join(int(v).to_bytes(3, "little", signed=True) for v in inter)
        else:
            raw = inter.astype(f"<i{nbytes}").tobytes()
        if hashlib.md5(raw).digest() != info["md5"]:
            raise ValueError("FLAC: decoded samples do not match the MD5 in STREAMINFO")
    return pcm, info


def resample(x: np.ndarray, src_rate: int, dst_rate: int) -> np.ndarray:
    """Polyphase resampling src_rate -> dst_rate (Kaiser-windowed low-pass at the narrower Nyquist)."""
    if src_rate == dst_rate:
        return np.asarray(x, dtype=np.float64)
    g = int(np.gcd(int(src_rate), int(dst_rate)))
    up, down = int(dst_rate) // g, int(src_rate) // g
    try:
        from scipy.signal import resample_poly
        return resample_poly(np.asarray(x, dtype=np.float64), up, down)
    except ImportError:
        pass
    # numpy version of the same design: firwin(2 * 10 * max(up, down) + 1, 1 / max(up, down), kaiser beta 5), gain `up`
    m = max(up, down)
    half = 10 * m
    t = np.arange(-half, half + 1, dtype=np.float64)
    h = np.sinc(t / m) / m * np.kaiser(2 * half + 1, 5.0)
    h *= up / h.sum()
    x = np.asarray(x, dtype=np.float64)
    n_out = -(-len(x) * up // down)
    z = np.zeros(len(x) * up, dtype=np.float64)
    z[::up] = x
    y = np.convolve(z, h)[half:half + len(z)]
    return y[::down][:n_out]


def decode_audio(source: Union[str, bytes, BinaryIO], sampling_rate: int = 16000) -> np.ndarray:
    """Path, bytes or binary file object holding RIFF/WAVE PCM or FLAC -> mono float32 in [-1, 1) at ``sampling_rate``
    (the reference's ``decode_audio(audio, sampling_rate=...)``, transcriber_faster_whisper.py:820-821)."""
    if isinstance(source, (bytes, bytearray)):
        data = bytes(source)
    elif hasattr(source, "read"):
        data = source.read()
    else:
        with open(source, "rb") as f:
            data = f.read()
    if data[:4] == b"fLaC":
        pcm, info = decode_flac(data)
        mono = pcm.astype(np.float64).mean(axis=0) / float(1 << (info["bps"] - 1))
        rate = info["rate"]
    elif data[:4] == b"RIFF" and data[8:12] == b"WAVE":
        with wave.open(io.BytesIO(data), "rb") as w:
            ch, width, rate, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
            raw = w.readframes(n)
        if width == 1:
            a = (np.frombuffer(raw, dtype=np.uint8).astype(np.float64) - 128.0) / 128.0
        elif width == 2:
            a = np.frombuffer(raw, dtype="<i2").astype(np.float64) / 32768.0
        elif width == 3:
            b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            a = np.where(v >= 1 << 23, v - (1 << 24), v).astype(np.float64) / float(1 << 23)
        elif width == 4:
            a = np.frombuffer(raw, dtype="<i4").astype(np.float64) / float(1 << 31)
        else:
            raise ValueError(f"WAV: unsupported sample width {width}")
        mono = a.reshape(-1, ch).mean(axis=1)
    else:
        raise ValueError(f"decode_audio: unsupported container (magic {data[:4]!r}); pass 16 kHz float32 PCM, WAV or FLAC "
                         "(the reference decodes other formats through PyAV, which is not a dependency here)")
    return resample(mono, rate, sampling_rate).astype(np.float32)
