"""A/B of the two encoder flash-attention kernels inside one process: WLB200_FA_SPLIT=1 (flash_attn_kernel: two softmax groups
on alternating key tiles; opt-in) vs 0 (flash_attn_pair_kernel, the default: all softmax warps on the same tile, two threads per row).
Plain run: device ms of the whole encoder per setting.  Under `ncu --profile-from-start off -k regex:flash_attn`: one
profiled encoder pass per setting.
    python tools/fa_ab.py --streams 16 --reps 3"""
import argparse
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--model", default="large-v3")
ap.add_argument("--streams", type=int, default=16)
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
f3 = np.stack([f[:, :3000] for f in eng.mel(waves)])
rt = ctypes.CDLL("libcudart.so")
outs = {}
for split in ("1", "0"):
    os.environ["WLB200_FA_SPLIT"] = split
    ms = []
    for rep in range(a.reps + 1):
        if rep == a.reps:
            rt.cudaProfilerStart()
        enc = eng.encode(f3)
        if rep == a.reps:
            rt.cudaProfilerStop()
            outs[split] = np.asarray(enc)[0].copy()
        if rep:
            ms.append(eng.last_device_ms(1))
        enc.release()
    print(f"WLB200_FA_SPLIT={split}: encoder {a.streams} streams {np.median(ms):.2f} ms  runs {[round(x, 2) for x in ms]}", flush=True)
print("max |split - pair| over the encoder output of stream 0:", float(np.abs(outs["1"] - outs["0"]).max()))
