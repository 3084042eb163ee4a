"""Run the reference's OWN cross-stream batcher (whisper_live/batch_inference.py BatchInferenceWorker._process_multi /
_process_single, :193-438) on top of whisperlive_b200.transcriber.B200WhisperModel -- the drop-in claim of SURVEY.md
section 8(b): everything the batcher reads from the transcriber (feature_extractor, encode, model.generate /
detect_language / is_multilingual, hf_tokenizer, get_prompt, max_length, frames_per_second,
_split_segments_by_timestamps) must exist with the reference's meaning.  The engine underneath is the CPU oracle (no GPU
in the build container); ctranslate2 / faster_whisper are stubbed exactly as in make_golden_transcribe.py.

The same batcher is also run over the reference's own WhisperModel (same oracle engine): both runs must agree segment for
segment.  Prints one JSON object; tests/test_boundary_cpu.py asserts on it.  Build container only (needs /root/reference).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from tests.golden import make_golden_transcribe as G  # noqa: E402
from oracle.engine import OracleWhisper  # noqa: E402
from oracle.mel import OracleFeatureExtractor  # noqa: E402
from whisperlive_b200 import synth  # noqa: E402
from whisperlive_b200 import tokenizer as wtok  # noqa: E402
from whisperlive_b200 import transcriber as ours  # noqa: E402
from whisperlive_b200.config import dims_for  # noqa: E402
from whisperlive_b200.weights import random_init  # noqa: E402


def run_batcher(ref_bi, model, audios, lang):
    worker = ref_bi.BatchInferenceWorker(model, max_batch_size=4, batch_window_ms=1)
    reqs = [ref_bi.BatchRequest(audio=a, language=lang, use_vad=False, initial_prompt="hello" if i == 1 else None)
            for i, a in enumerate(audios)]
    worker._process_multi(reqs)          # the batched path, called synchronously (no thread needed)
    single = ref_bi.BatchRequest(audio=audios[0], language=lang, use_vad=False)
    worker._process_single(single)       # batch of one: delegates to transcriber.transcribe()
    rows = []
    for r in reqs + [single]:
        segs = r.result
        rows.append(dict(
            error=None if r.error is None else repr(r.error), done=r.future.is_set(),
            segments=None if segs is None else [dict(tokens=list(s.tokens), start=round(float(s.start), 6), end=round(float(s.end), 6),
                                                     text=s.text, avg_logprob=round(float(s.avg_logprob), 6),
                                                     no_speech_prob=round(float(s.no_speech_prob), 6), temperature=s.temperature,
                                                     compression_ratio=round(float(s.compression_ratio), 6)) for s in segs],
            language=getattr(r.info, "language", None), duration=round(float(getattr(r.info, "duration", -1.0)), 6),
            segment_type=None if not segs else type(segs[0]).__name__))
    return rows


def main():
    G.install_stubs()
    sys.path.insert(0, "/root/reference")
    from whisper_live import batch_inference as ref_bi
    from whisper_live.transcriber import transcriber_faster_whisper as ref_tr

    out = {}
    for model_name in ("micro.en", "micro"):
        dims = dims_for(model_name)
        engine = OracleWhisper(random_init(dims, seed=0), dims)
        mine = ours.B200WhisperModel(model_name, engine=engine, hf_tokenizer=wtok.build_synthetic_tokenizer(dims.vocab),
                                     feature_extractor=OracleFeatureExtractor(dims.n_mels))
        theirs = G.reference_model(ref_tr, engine, dims)     # the reference's own WhisperModel over the same engine
        audios = [synth.speech_like(6.0, seed=1), synth.speech_like(3.0, seed=2), synth.speech_like(9.5, seed=3)]
        lang = None if dims.multilingual else "en"
        a = run_batcher(ref_bi, mine, audios, lang)
        engine._sampling_calls = 0     # both runs start from the same sampling-noise state (CT2: a fresh generator)
        b = run_batcher(ref_bi, theirs, audios, lang)
        out[model_name] = dict(over_b200_model=a, over_reference_model=b)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
