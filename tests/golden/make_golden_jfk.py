"""Decode the reference's only audio asset (/root/reference/assets/jfk.flac, the input of its WER test,
tests/test_server.py:92-118 and BASELINE config 1) into a committed 16 kHz mono PCM fixture.

No FLAC decoder is installed here (no PyAV / soundfile / ffmpeg), so this script carries a small pure-Python one --
enough of the format for this file (fixed + LPC + verbatim + constant subframes, Rice partitions, the four stereo
modes).  It is self-checking: STREAMINFO holds the MD5 of the decoded PCM and the script refuses to write the fixture
unless it matches.  Resampling 44.1 kHz -> 16 kHz is scipy's polyphase filter (the reference goes through
libswresample inside faster_whisper.decode_audio; the fixture is an input, both sides of every test read the same samples).

    python tests/golden/make_golden_jfk.py        # writes tests/golden/jfk_16k_i16.npy (int16, 176000 samples)
"""
import hashlib
import os
import sys

import numpy as np

SRC = "/root/reference/assets/jfk.flac"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "jfk_16k_i16.npy")


class Bits:
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

    def unary(self) -> int:                      # number of 0 bits before the next 1 bit
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

    def align(self):
        self.acc, self.n = 0, 0


def residual(br: Bits, blocksize: int, order: int):
    method = br.read(2)
    assert method in (0, 1), "reserved residual coding method"
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


FIXED = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}


def subframe(br: Bits, blocksize: int, bps: int):
    assert br.read(1) == 0
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
            assert shift >= 0
            coef = [br.signed(prec) for _ in range(order)]
        else:
            order = typ - 8
            s = [br.signed(bps) for _ in range(order)]
            coef, shift = FIXED[order], 0
        res = residual(br, blocksize, order)
        rc = coef[::-1]
        for r in res:
            pred = 0
            for c, v in zip(rc, s[len(s) - order:] if order else ()):
                pred += c * v
            s.append(r + (pred >> shift))
    else:
        raise ValueError(f"reserved subframe type {typ}")
    if wasted:
        s = [v << wasted for v in s]
    return s


def decode_flac(data: bytes):
    assert data[:4] == b"fLaC"
    pos, info = 4, None
    while True:
        last, btype = data[pos] >> 7, data[pos] & 127
        size = int.from_bytes(data[pos + 1:pos + 4], "big")
        if btype == 0:
            b = Bits(data, pos + 4)
            b.read(16); b.read(16); b.read(24); b.read(24)
            info = dict(rate=b.read(20), ch=b.read(3) + 1, bps=b.read(5) + 1, total=b.read(36), md5=data[pos + 4 + 18:pos + 4 + 34])
        pos += 4 + size
        if last:
            break
    ch_out = [[] for _ in range(info["ch"])]
    BS = {1: 192, 2: 576, 3: 1152, 4: 2304, 5: 4608}
    BPS = {1: 8, 2: 12, 4: 16, 5: 20, 6: 24}
    done = 0
    while done < info["total"]:
        br = Bits(data, pos)
        assert br.read(14) == 0x3FFE, f"lost frame sync at byte {pos}"
        br.read(1); br.read(1)
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
        else:
            blocksize = BS[bs_code]
        if sr_code == 12:
            br.read(8)
        elif sr_code in (13, 14):
            br.read(16)
        br.read(8)                                # CRC-8 (the MD5 at the end covers correctness)
        bps = info["bps"] if ss_code == 0 else BPS[ss_code]
        if ch_code < 8:
            subs = [subframe(br, blocksize, bps) for _ in range(ch_code + 1)]
        elif ch_code == 8:                        # left / side
            l = subframe(br, blocksize, bps); sd = subframe(br, blocksize, bps + 1)
            subs = [l, [a - b for a, b in zip(l, sd)]]
        elif ch_code == 9:                        # side / right
            sd = subframe(br, blocksize, bps + 1); r = subframe(br, blocksize, bps)
            subs = [[a + b for a, b in zip(sd, r)], r]
        elif ch_code == 10:                       # mid / side
            md = subframe(br, blocksize, bps); sd = subframe(br, blocksize, bps + 1)
            subs = []
            l, r = [], []
            for a, b in zip(md, sd):
                a = (a << 1) | (b & 1)
                l.append((a + b) >> 1); r.append((a - b) >> 1)
            subs = [l, r]
        else:
            raise ValueError("reserved channel assignment")
        br.align()
        br.read(16)                               # CRC-16
        pos = br.p
        for c, sub in zip(ch_out, subs):
            c.extend(sub)
        done += blocksize
    pcm = np.asarray(ch_out, dtype=np.int64)[:, :info["total"]]
    nbytes = (info["bps"] + 7) // 8
    inter = pcm.T.reshape(-1)
    raw = b"".join(int(v).to_bytes(nbytes, "little", signed=True) for v in inter) if nbytes == 3 else inter.astype(f"<i{nbytes}").tobytes()
    assert hashlib.md5(raw).digest() == info["md5"], "decoded PCM does not match the MD5 in STREAMINFO"
    return pcm, info


def main():
    from scipy.signal import resample_poly
    data = open(SRC, "rb").read()
    pcm, info = decode_flac(data)
    print("decoded", pcm.shape, info["rate"], "Hz", info["bps"], "bit -- MD5 ok")
    mono = pcm.astype(np.float64).mean(axis=0) / float(1 << (info["bps"] - 1))
    g = np.gcd(16000, info["rate"])
    y = resample_poly(mono, 16000 // g, info["rate"] // g)
    y16 = np.clip(np.round(y * 32768.0), -32768, 32767).astype(np.int16)
    np.save(OUT, y16)
    print("wrote", OUT, y16.shape, "peak", int(np.abs(y16).max()), "rms", float(np.sqrt((y16.astype(np.float64) ** 2).mean())))


if __name__ == "__main__":
    sys.exit(main())
