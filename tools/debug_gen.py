"""GPU debug: long-prompt generate vs oracle, all hypotheses."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.test_gpu_parity import engine, feats_for
from whisperlive_b200.tokenizer import build_synthetic_tokenizer, Tokenizer
from whisperlive_b200.transcriber import get_suppressed_tokens

eng, orc = engine("micro.en", seed=0)
dims = eng.dims
sp = orc.spec
tk = Tokenizer(build_synthetic_tokenizer(dims.vocab), False)
sup = list(get_suppressed_tokens(tk, [-1]))
feats = np.stack([feats_for(dims, 6.0, 1), feats_for(dims, 14.0, 7)])
enc, oenc = eng.encode(feats), orc.encode(feats)
rng = np.random.default_rng(0)
for P in (1, 5, 60, 200):
    prev = [sp.timestamp_begin - 3] + rng.integers(256, 50000, P - 2).tolist() if P > 1 else []
    prompts = [prev + [sp.sot]] * 2
    for supp in ([1, 2, 3], sup):
        kw = dict(beam_size=5, num_hypotheses=5, suppress_tokens=supp)
        g = eng.generate(enc, prompts, **kw)
        r = orc.generate(oenc, prompts, **kw)
        for b in range(2):
            same = [x == y for x, y in zip(g[b].sequences_ids, r[b].sequences_ids)]
            print(f"P={len(prompts[b])} nsup={len(supp)} b={b} same={same} lens_g={[len(x) for x in g[b].sequences_ids]} "
                  f"lens_r={[len(x) for x in r[b].sequences_ids]}\n   scores_g={np.round(g[b].scores, 4).tolist()}\n   scores_r={np.round(r[b].scores, 4).tolist()} "
                  f"nsp {g[b].no_speech_prob:.2e} {r[b].no_speech_prob:.2e} steps {g[b].steps} {r[b].steps}")
