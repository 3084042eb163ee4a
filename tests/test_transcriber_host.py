"""Host logic of whisperlive_b200.transcriber (SURVEY.md §8a rows H4-H8, K14 host part)
against tests/golden/transcribe_reference.json, which was produced by running the
reference's own WhisperModel orchestration over the same CPU oracle engine
(tests/golden/make_golden_transcribe.py).  The engine here is the oracle -- injected for
the test only; the product constructs the CUDA engine."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.engine import OracleWhisper
from oracle.mel import OracleFeatureExtractor
from tests.golden.make_golden_transcribe import SCENARIOS, make_audio
from whisperlive_b200.config import dims_for
from whisperlive_b200.tokenizer import build_synthetic_tokenizer
from whisperlive_b200.transcriber import B200WhisperModel
from whisperlive_b200.weights import random_init

GOLD = os.path.join(os.path.dirname(__file__), "golden", "transcribe_reference.json")


def _model(sc):
    dims = dims_for(sc["model"])
    eng = OracleWhisper(random_init(dims, seed=sc["seed"]), dims)
    return B200WhisperModel(sc["model"], engine=eng, hf_tokenizer=build_synthetic_tokenizer(dims.vocab),
                            feature_extractor=OracleFeatureExtractor(dims.n_mels))


def _check(segs, gold):
    assert len(segs) == len(gold)
    for s, g in zip(segs, gold):
        assert s.id == g["id"] and s.seek == g["seek"] and s.tokens == g["tokens"] and s.text == g["text"]
        assert s.temperature == g["temperature"]
        for k in ("start", "end", "avg_logprob", "compression_ratio", "no_speech_prob"):
            assert getattr(s, k) == pytest.approx(g[k], abs=1e-6), k
        if g["words"] is None:
            assert s.words is None
        else:
            assert [(w.word, w.start, w.end) for w in s.words] == [(w["word"], w["start"], w["end"]) for w in g["words"]]
            for w, gw in zip(s.words, g["words"]):
                assert w.probability == pytest.approx(gw["probability"], abs=1e-6)


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_transcribe_matches_reference_orchestration(name):
    torch.set_num_threads(8)
    gold = json.load(open(GOLD))[name]
    sc = SCENARIOS[name]
    segs, info = _model(sc).transcribe(make_audio(sc["audio"]), **sc["kw"])
    _check(segs, gold["segments"])
    assert info.language == gold["language"]
    assert float(info.language_probability) == pytest.approx(gold["language_probability"], abs=1e-6)
    assert info.duration == pytest.approx(gold["duration"])


def test_batched_equals_single():
    """transcribe_batch advances streams in lockstep; results equal per-stream transcribe."""
    torch.set_num_threads(8)
    gold = json.load(open(GOLD))
    names = ["en_two_windows_prompt_hotwords", "en_silence", "en_short_ladder"]
    m = _model(SCENARIOS[names[0]])
    res = m.transcribe_batch([make_audio(SCENARIOS[n]["audio"]) for n in names], [SCENARIOS[n]["kw"] for n in names])
    for n, (segs, info) in zip(names, res):
        _check(segs, gold[n]["segments"])


def test_empty_audio_returns_none():
    m = _model(SCENARIOS["en_silence"])
    assert m.transcribe(np.zeros(0, np.float32)) == (None, None)
