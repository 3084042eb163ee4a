"""Generate tests/golden/transcribe_reference.json by EXECUTING the reference's own
orchestration -- whisper_live/transcriber/transcriber_faster_whisper.py ``WhisperModel``
(transcribe :692, generate_segments :1049, generate_with_fallback :1350, get_prompt :1480,
_split_segments_by_timestamps :970, add_word_timestamps :1515, find_alignment :1646,
detect_language :1716) -- on top of the CPU oracle engine.

The module imports ``ctranslate2`` and ``faster_whisper.*`` at load (:15-30); neither is
installed, so they are stubbed in sys.modules: the engine stub is oracle.engine.OracleWhisper
(same API), FeatureExtractor/pad_or_trim are oracle.mel, Tokenizer is
whisperlive_b200.tokenizer.Tokenizer.  What this pins is therefore the HOST LOGIC of
whisperlive_b200.transcriber (rows H4-H8 of SURVEY.md §8a) against the reference's real
Python, given identical engine outputs.  Runs only in the build container.

    python tests/golden/make_golden_transcribe.py
"""
import dataclasses
import importlib.machinery
import json
import logging
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from oracle import mel as omel  # noqa: E402
from oracle.engine import OracleWhisper  # noqa: E402
from whisperlive_b200 import synth, tokenizer as wtok  # noqa: E402
from whisperlive_b200.config import dims_for  # noqa: E402
from whisperlive_b200.weights import random_init  # noqa: E402

SCENARIOS = {
    "en_short_ladder": dict(model="micro.en", seed=0, audio=("speech", 6.0, 1),
                            kw=dict(temperature=[0.0, 0.4])),
    "en_two_windows_prompt_hotwords": dict(model="micro.en", seed=0, audio=("speech", 33.0, 2),
                                           kw=dict(temperature=[0.0], initial_prompt="hello there",
                                                   hotwords="foo bar", log_prob_threshold=None)),
    "multi_detect_words": dict(model="micro", seed=1, audio=("speech", 5.0, 3),
                               kw=dict(temperature=[0.0], word_timestamps=True, log_prob_threshold=None)),
    "en_notimestamps_prefix": dict(model="micro.en", seed=2, audio=("speech", 4.0, 4),
                                   kw=dict(temperature=[0.0], without_timestamps=True, prefix="hello",
                                           max_new_tokens=20, log_prob_threshold=None)),
    "en_silence": dict(model="micro.en", seed=0, audio=("silence", 2.0, 0), kw=dict(temperature=[0.0])),
    # H8: VAD clipping + restore_speech_timestamps (reference :830-838, :1792-1817) with the deterministic stub detector
    # (tests/stub_vad.py) standing in for Silero on BOTH sides
    "en_vad_two_chunks": dict(model="micro.en", seed=0, audio=("gapped", (4.0, 3.5, 5.0), 6),
                              kw=dict(temperature=[0.0], vad_filter=True, log_prob_threshold=None)),
    "en_vad_words": dict(model="micro.en", seed=3, audio=("gapped", (3.0, 2.6, 2.5, 4.0, 3.0), 7),
                         kw=dict(temperature=[0.0], vad_filter=True, word_timestamps=True, log_prob_threshold=None,
                                 vad_parameters=dict(threshold=0.5, min_silence_duration_ms=1500))),
    "en_vad_all_silence": dict(model="micro.en", seed=0, audio=("silence", 3.0, 0), kw=dict(temperature=[0.0], vad_filter=True)),
    "en_beam2_greedy_words": dict(model="micro.en", seed=3, audio=("speech", 7.5, 5),
                                  kw=dict(temperature=[0.0], beam_size=2, word_timestamps=True,
                                          log_prob_threshold=None, condition_on_previous_text=False)),
}


def make_audio(spec):
    kind, sec, seed = spec
    if kind == "gapped":   # speech, silence, speech, ... (seconds): what VAD gating exists for
        parts = [synth.speech_like(d, seed=seed + i) if i % 2 == 0 else synth.silence(d) for i, d in enumerate(sec)]
        return np.concatenate(parts).astype(np.float32)
    return synth.speech_like(sec, seed=seed) if kind == "speech" else synth.silence(sec)


def _stub(name):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    sys.modules[name] = m
    return m


def install_stubs():
    ct2 = _stub("ctranslate2")
    ct2.models = _stub("ctranslate2.models")
    ct2.models.Whisper = OracleWhisper
    ct2.models.WhisperGenerationResult = object

    class StorageView:
        @staticmethod
        def from_array(a):
            return a
    ct2.StorageView = StorageView

    fw = _stub("faster_whisper")
    audio = _stub("faster_whisper.audio")
    audio.decode_audio = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no decoder"))
    audio.pad_or_trim = omel.pad_or_trim
    fe = _stub("faster_whisper.feature_extractor")
    fe.FeatureExtractor = omel.OracleFeatureExtractor
    tk = _stub("faster_whisper.tokenizer")
    tk._LANGUAGE_CODES = wtok._LANGUAGE_CODES
    tk.Tokenizer = wtok.Tokenizer
    ut = _stub("faster_whisper.utils")
    ut.download_model = lambda *a, **k: None
    ut.format_timestamp = lambda s, *a, **k: f"{s:.3f}"
    ut.get_logger = lambda: logging.getLogger("faster_whisper")

    def get_end(segments):
        return next((w["end"] for s in reversed(segments) for w in reversed(s["words"])),
                    segments[-1]["end"] if segments else None)
    ut.get_end = get_end
    from tests import stub_vad
    vad = _stub("faster_whisper.vad")
    for n in ("SpeechTimestampsMap", "VadOptions", "collect_chunks", "get_speech_timestamps"):
        setattr(vad, n, getattr(stub_vad, n))
    fw.audio, fw.feature_extractor, fw.tokenizer, fw.utils, fw.vad = audio, fe, tk, ut, vad


def reference_model(ref, engine, dims):
    m = ref.WhisperModel.__new__(ref.WhisperModel)
    m.logger = logging.getLogger("faster_whisper")
    m.model = engine
    m.hf_tokenizer = wtok.build_synthetic_tokenizer(dims.vocab)
    m.feat_kwargs = {"feature_size": dims.n_mels}
    m.feature_extractor = omel.OracleFeatureExtractor(dims.n_mels)
    m.input_stride = 2
    m.num_samples_per_token = 320
    m.frames_per_second = 100
    m.tokens_per_second = 50
    m.time_precision = 0.02
    m.max_length = 448
    return m


def seg_to_json(s):
    d = dataclasses.asdict(s)
    for k in ("start", "end", "avg_logprob", "compression_ratio", "no_speech_prob"):
        d[k] = float(d[k])
    if d["words"]:
        for w in d["words"]:
            for k in ("start", "end", "probability"):
                w[k] = float(w[k])
    return d


def main():
    torch.set_num_threads(8)
    install_stubs()
    sys.path.insert(0, "/root/reference")
    from whisper_live.transcriber import transcriber_faster_whisper as ref

    out = {}
    for name, sc in SCENARIOS.items():
        dims = dims_for(sc["model"])
        engine = OracleWhisper(random_init(dims, seed=sc["seed"]), dims)
        model = reference_model(ref, engine, dims)
        segments, info = model.transcribe(make_audio(sc["audio"]), **sc["kw"])
        out[name] = dict(
            segments=None if segments is None else [seg_to_json(s) for s in segments],
            language=None if info is None else info.language,
            language_probability=None if info is None else float(info.language_probability),
            duration=None if info is None else float(info.duration),
            duration_after_vad=None if info is None else float(info.duration_after_vad))
        print(name, None if segments is None else [(s.seek, round(s.start, 2), round(s.end, 2), s.temperature, len(s.tokens)) for s in segments])
    with open(os.path.join(HERE, "transcribe_reference.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
