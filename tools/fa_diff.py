"""Where do the two flash-attention kernels disagree?  Encoder outputs of the SAME inputs with WLB200_FA_SPLIT=0 (pair, default)
and =1 (split), compared per stream and per block of 128 encoder positions (= the kernels' query tiles), next to the
oracle's encoder where it is cheap (small models).  Written at the end of round 2: with the split kernel two decode-level
parity tests fail although every encoder-level check passes (profiles/flash_ab_r2.md); this is the first thing to run on
a B200 to see whether the split kernel is wrong on particular (stream, tile)s or merely rounds differently.

    python tools/fa_diff.py --model micro.en --seconds 6,6 --seeds 1,2          # the sampling test's inputs
    python tools/fa_diff.py --model tiny --seconds 6,6,6,14 --seeds 1,2,3,7     # test_generate_matches_oracle[tiny-5]
    python tools/fa_diff.py --model large-v3 --seconds 8,5 --seeds 31,32 --no-oracle"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--model", default="micro.en")
ap.add_argument("--seconds", default="6,6")
ap.add_argument("--seeds", default="1,2")
ap.add_argument("--weights-seed", type=int, default=0)
ap.add_argument("--no-oracle", action="store_true")
ap.add_argument("--repeat", type=int, default=3, help="encodes per setting (a race shows up as run-to-run differences)")
a = ap.parse_args()

from oracle import mel as omel
from whisperlive_b200 import synth
from whisperlive_b200.config import dims_for
from whisperlive_b200.engine import B200Whisper
from whisperlive_b200.weights import random_init

dims = dims_for(a.model)
w = random_init(dims, seed=a.weights_seed)
secs = [float(x) for x in a.seconds.split(",")]
seeds = [int(x) for x in a.seeds.split(",")]
feats = np.stack([omel.pad_or_trim(omel.log_mel(synth.speech_like(s, seed=sd), dims.n_mels)[:, :-1]) for s, sd in zip(secs, seeds)])
eng = B200Whisper(dims, w, max_streams=len(secs), max_beam=1)
outs = {}
for split in ("0", "1"):
    os.environ["WLB200_FA_SPLIT"] = split
    runs = []
    for _ in range(a.repeat):
        enc = eng.encode(feats)
        runs.append(np.asarray(enc).copy())
        enc.release()
    outs[split] = runs
    same = all(np.array_equal(runs[0], r) for r in runs[1:])
    print(f"WLB200_FA_SPLIT={split}: {a.repeat} encodes bit-identical: {same}")
    if not same:
        for i, r in enumerate(runs[1:], 1):
            d = np.abs(r - runs[0])
            print(f"   run {i} vs run 0: max {d.max():.5f} at {np.unravel_index(d.argmax(), d.shape)}")
pair, split = outs["0"][0], outs["1"][0]
ref = None
if not a.no_oracle:
    from oracle.engine import OracleWhisper
    ref = OracleWhisper(w, dims).encode(feats).enc.numpy()
print("stream  tile  max|split-pair|  " + ("max|pair-oracle|  max|split-oracle|" if ref is not None else ""))
for b in range(len(secs)):
    for t in range(12):
        sl = slice(t * 128, min(1500, (t + 1) * 128))
        d = float(np.abs(split[b, sl] - pair[b, sl]).max())
        line = f"{b:6d} {t:5d}  {d:14.5f}"
        if ref is not None:
            line += f"  {float(np.abs(pair[b, sl] - ref[b, sl]).max()):15.5f}  {float(np.abs(split[b, sl] - ref[b, sl]).max()):17.5f}"
        flag = "   <--" if d > 0.03 else ""
        print(line + flag)
