"""Decode the reference's only audio asset (/root/reference/assets/jfk.flac, the input of its WER test,
tests/test_server.py:92-118 and BASELINE config 1) into a committed 16 kHz mono PCM fixture.

No FLAC decoder is installed here (no PyAV / soundfile / ffmpeg), so the product carries a small pure-Python one
(whisperlive_b200/audio.py: fixed + LPC + verbatim + constant subframes, Rice partitions, the four stereo modes).  It is
self-checking: STREAMINFO holds the MD5 of the decoded PCM and decode_flac raises unless it matches.  Resampling 44.1 kHz -> 16 kHz is scipy's polyphase filter (the reference goes through
libswresample inside faster_whisper.decode_audio; the fixture is an input, both sides of every test read the same samples).

    python tests/golden/make_golden_jfk.py        # writes tests/golden/jfk_16k_i16.npy (int16, 176000 samples)
"""
import hashlib
import os
import sys

import numpy as np

SRC = "/root/reference/assets/jfk.flac"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "jfk_16k_i16.npy")


sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from whisperlive_b200.audio import decode_flac   # noqa: E402  (self-checking: raises unless the STREAMINFO MD5 matches)


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
