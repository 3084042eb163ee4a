"""Encoder time (K2-K7) at S streams, large-v3 shape: device ms of wl_encode (all passes), TFLOP/s against the algorithmic
2.589 TFLOP per window (encoder 2.274 + cross-KV 0.315).
    python tools/enc_time.py --streams 32 --reps 3"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--model", default="large-v3")
ap.add_argument("--streams", type=int, default=32)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()

from whisperlive_b200 import synth
from whisperlive_b200.config import dims_for
from whisperlive_b200.engine import B200Whisper
from whisperlive_b200.weights import random_init

dims = dims_for(a.model)
eng = B200Whisper(dims, random_init(dims, seed=0), max_streams=a.streams, max_beam=1, enc_slots=a.streams + 1,
                  alignment_heads=[(dims.dec_layers - 1, 0)])
waves = [synth.speech_like(30.0, seed=1234 + i) for i in range(a.streams)]
feats = eng.mel(waves)
f3 = np.stack([f[:, :3000] for f in feats])
d, L = dims.d_model, dims.enc_layers
flops = a.streams * (L * (2 * 1500 * d * d * 12 + 4 * 1500 * 1500 * d) + 2 * 1500 * d * d * 2 * dims.dec_layers
                     + 2 * 3000 * dims.n_mels * 3 * d + 2 * 1500 * d * 3 * d)
for rep in range(a.reps + 1):
    enc = eng.encode(f3)
    ms = eng.last_device_ms(1)
    enc.release()
    if rep:
        print(f"encoder {a.streams} streams: {ms:.2f} ms = {flops / ms / 1e9:.0f} TFLOP/s ({flops / 1e12:.2f} TFLOP)", flush=True)
