"""Sweep of WLB200_XA_PREFETCH (L2 prefetch of the cross-attention K/V from the layer's first LayerNorm) inside one process:
ms per token step of the device-side decode loop at S streams, large-v3 shape, beam 4.
    python tools/sweep_prefetch.py --streams 32 --tokens 60 --values 0,4,8,12,16"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--model", default="large-v3")
ap.add_argument("--streams", default="32")
ap.add_argument("--tokens", type=int, default=60)
ap.add_argument("--beam", type=int, default=4)
ap.add_argument("--values", default="0,4,8,12,16")
ap.add_argument("--cgemm", default="0", help="comma list of WLB200_CGEMM values to cross with the prefetch values")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()

from whisperlive_b200 import synth
from whisperlive_b200.config import dims_for
from whisperlive_b200.engine import B200Whisper
from whisperlive_b200.weights import random_init

dims = dims_for(a.model)
smax = max(int(x) for x in a.streams.split(","))
eng = B200Whisper(dims, random_init(dims, seed=0), max_streams=smax, max_beam=max(a.beam, 1), enc_slots=smax + 1,
                  alignment_heads=[(dims.dec_layers - 1, 0)])
waves = [synth.speech_like(30.0, seed=1234 + i) for i in range(smax)]
sot = [eng.sot] if not dims.multilingual else [eng.sot, eng.sot + 1, eng.sot + 1 + dims.num_languages + 1]
feats = eng.mel(waves)
enc_all = eng.encode(np.stack([f[:, :3000] for f in feats]))
for S in [int(x) for x in a.streams.split(",")]:
    enc = enc_all.select(list(range(S)))
    base = None
    for cg, v in [(cg, int(x)) for cg in a.cgemm.split(",") for x in a.values.split(",")]:
        os.environ["WLB200_XA_PREFETCH"] = str(v)
        os.environ["WLB200_CGEMM"] = cg
        ms = []
        for rep in range(a.reps + 1):
            out = eng.generate(enc, [sot] * S, beam_size=a.beam, suppress_tokens=[eng.eot], suppress_blank=False, max_length=2 * a.tokens)
            if rep:
                ms.append(eng.last_device_ms(2) / max(o.steps for o in out))
        m = float(np.median(ms))
        base = base or m
        print(f"streams {S:3d} cgemm {cg} prefetch {v:3d}: {m:.4f} ms per token step ({m / base:.3f} of the first row)  runs {[round(x, 4) for x in ms]}", flush=True)
