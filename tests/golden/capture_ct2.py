#!/usr/bin/env python
"""Capture golden vectors from the REAL reference stack (faster-whisper + CTranslate2) -- SURVEY.md §8(c) tier (ii).

The oracle's beam search / timestamp rules / alignment restate CTranslate2 from its published algorithm and are
PARITY-UNPINNED at that boundary (oracle/__init__.py): neither package, nor any checkpoint, exists in the build
container.  This script is the pin: run it ONCE on any machine that has

    pip install faster-whisper==1.2.0        (pulls ctranslate2 4.x, tokenizers)
    a model directory or hub access          (default: tiny.en; --model large-v3 for the metric's architecture)

and commit what it writes.  ``tests/test_ct2_capture.py`` consumes the fixtures when they are present (and is skipped,
loudly, when they are not):

    tests/golden/ct2_capture.npz   features / encoder output samples / alignment matrices (float16-compressed)
    tests/golden/ct2_capture.json  token ids, scores, no_speech_prob, language probabilities, alignments, versions,
                                   the model.bin header (variable names, shapes, dtypes: validates ct2_format.py, N1)

Everything is recorded at the exact call sites the reference uses
(whisper_live/transcriber/transcriber_faster_whisper.py): FeatureExtractor.__call__ :862, Whisper.encode :1348,
Whisper.generate :1394-1407 (beam 5 / patience 1 / length_penalty 1 / suppress_blank / suppress list / timestamps on,
and the sampling variant :1381-1387), Whisper.detect_language :1140, Whisper.align :1657-1663, and the whole
``WhisperModel.transcribe`` on assets/jfk.flac (the reference's only executable pin, tests/test_server.py:92-118).

    python tests/golden/capture_ct2.py [--model tiny.en] [--audio /path/to/jfk.flac] [--device cpu] [--compute-type float32]
"""
from __future__ import annotations

import argparse
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)


def synth_chunks():
    from whisperlive_b200 import synth
    return {"speech_6s": synth.speech_like(6.0, seed=1), "speech_29s": synth.speech_like(29.0, seed=2),
            "noise_5s": synth.white_noise(5.0, seed=3), "silence_2s": synth.silence(2.0)}


def model_bin_header(path: str, max_vars: int = 4000):
    """Walk the CTranslate2 model.bin container without interpreting the payloads: binary version, spec name/revision,
    then per variable (name, shape, dtype id / item size, byte count).  Mirrors what ct2_format.read_ct2_model_bin
    assumes, so a mismatch shows up as a parse error here instead of silently wrong weights there."""
    out = {"variables": []}
    with open(path, "rb") as f:
        def u8():
            return struct.unpack("<B", f.read(1))[0]

        def u16():
            return struct.unpack("<H", f.read(2))[0]

        def u32():
            return struct.unpack("<I", f.read(4))[0]

        def string():
            n = u16()
            return f.read(n).rstrip(b"\x00").decode("utf-8", "replace")
        out["binary_version"] = u32()
        if out["binary_version"] >= 2:
            out["spec"] = string()
            out["spec_revision"] = u32()
        n_vars = u32()
        out["n_variables"] = n_vars
        for _ in range(min(n_vars, max_vars)):
            name = string()
            rank = u8()
            shape = [u32() for _ in range(rank)]
            if out["binary_version"] >= 4:
                dtype_id = u8()
                nbytes = u32()
                item = None
            else:
                item = u8()
                nbytes = u32() * item
                dtype_id = None
            f.seek(nbytes, 1)
            out["variables"].append({"name": name, "shape": shape, "dtype_id": dtype_id, "item_size": item, "nbytes": nbytes})
        if out["binary_version"] >= 3:
            try:
                n_alias = u32()
                out["aliases"] = [[string(), string()] for _ in range(n_alias)]
            except Exception:
                out["aliases"] = None
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="tiny.en")
    ap.add_argument("--audio", default=os.path.join("/root/reference", "assets", "jfk.flac"))
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--compute-type", default="float32")
    a = ap.parse_args()
    try:
        import ctranslate2
        import faster_whisper
        from faster_whisper import WhisperModel
        from faster_whisper.tokenizer import Tokenizer
    except Exception as e:   # the expected outcome in the build container
        print(f"capture_ct2: faster_whisper / ctranslate2 are not importable here ({type(e).__name__}: {e}); nothing captured.")
        return 2

    model = WhisperModel(a.model, device=a.device, compute_type=a.compute_type)
    ct2 = model.model
    fe = model.feature_extractor
    meta = {"model": a.model, "faster_whisper": faster_whisper.__version__, "ctranslate2": ctranslate2.__version__,
            "device": a.device, "compute_type": a.compute_type, "n_mels": int(ct2.n_mels), "multilingual": bool(ct2.is_multilingual)}
    arrays, rec = {}, {"meta": meta, "chunks": {}}
    tok = Tokenizer(model.hf_tokenizer, ct2.is_multilingual, task="transcribe", language="en")
    sot_seq = list(tok.sot_sequence)
    suppress = list(faster_whisper.transcribe.get_suppressed_tokens(tok, [-1]))
    rec["meta"].update(sot_sequence=sot_seq, eot=tok.eot, no_timestamps=tok.no_timestamps, timestamp_begin=tok.timestamp_begin,
                       suppress_tokens=suppress)

    chunks = synth_chunks()
    if os.path.exists(a.audio):
        from faster_whisper.audio import decode_audio
        chunks["jfk"] = decode_audio(a.audio, sampling_rate=16000)
    for name, wav in chunks.items():
        wav = np.asarray(wav, dtype=np.float32)
        feats = fe(wav)                                              # :862
        arrays[f"{name}__features"] = feats.astype(np.float16)
        window = faster_whisper.audio.pad_or_trim(feats[:, :-1])     # :1127
        enc = ct2.encode(ctranslate2.StorageView.from_array(window[None].astype(np.float32)), to_cpu=True)   # :1348
        enc_np = np.array(enc)
        arrays[f"{name}__encoder_sub"] = enc_np[0, ::25].astype(np.float16)
        c = {"n_samples": int(len(wav)), "features_shape": list(feats.shape)}
        # beam search with the reference's live-path defaults (:1394-1407)
        for label, kw in {
            "beam5": dict(beam_size=5, patience=1, length_penalty=1, max_length=448, return_scores=True, return_no_speech_prob=True,
                          suppress_blank=True, suppress_tokens=suppress, max_initial_timestamp_index=50),
            "greedy": dict(beam_size=1, max_length=448, return_scores=True, return_no_speech_prob=True, suppress_blank=True,
                           suppress_tokens=suppress, max_initial_timestamp_index=50),
            "beam5_nots": dict(beam_size=5, max_length=448, return_scores=True, return_no_speech_prob=True, suppress_tokens=suppress),
            "beam2_hyp2": dict(beam_size=2, num_hypotheses=2, max_length=64, return_scores=True, return_no_speech_prob=True,
                               suppress_tokens=suppress),
        }.items():
            prompt = sot_seq + ([tok.no_timestamps] if label == "beam5_nots" else [])
            r = ct2.generate(enc, [prompt], **kw)[0]
            c[label] = {"prompt": prompt, "sequences_ids": [list(map(int, s)) for s in r.sequences_ids],
                        "scores": [float(x) for x in r.scores], "no_speech_prob": float(r.no_speech_prob)}
        # prompt with previous text and with a prefix: pins sample_begin / prefix semantics
        prev = [tok.sot_prev] + tok.encode(" hello there") + sot_seq
        r = ct2.generate(enc, [prev], beam_size=5, max_length=448, return_scores=True, return_no_speech_prob=True,
                         suppress_tokens=suppress)[0]
        c["beam5_prev"] = {"prompt": prev, "sequences_ids": [list(map(int, s)) for s in r.sequences_ids],
                           "scores": [float(x) for x in r.scores], "no_speech_prob": float(r.no_speech_prob)}
        pre = sot_seq + [tok.timestamp_begin] + tok.encode(" And so")
        r = ct2.generate(enc, [pre], beam_size=5, max_length=448, return_scores=True, suppress_tokens=suppress)[0]
        c["beam5_prefix"] = {"prompt": pre, "sequences_ids": [list(map(int, s)) for s in r.sequences_ids],
                             "scores": [float(x) for x in r.scores]}
        # alignment of the beam-5 text tokens (:1657-1663)
        text = [t for t in c["beam5"]["sequences_ids"][0] if t < tok.eot]
        if text:
            al = ct2.align(enc, sot_seq, [text], feats.shape[-1] - 1, median_filter_width=7)[0]
            c["align"] = {"text_tokens": text, "num_frames": int(feats.shape[-1] - 1),
                          "alignments": [[int(x), int(y)] for x, y in al.alignments],
                          "text_token_probs": [float(p) for p in al.text_token_probs]}
        if ct2.is_multilingual:
            c["detect_language"] = [[t, float(p)] for t, p in ct2.detect_language(enc)[0][:10]]   # :1140
        rec["chunks"][name] = c

    # the whole reference path on the fixture the reference's own test pins (tests/test_server.py:92-118)
    if "jfk" in chunks:
        segs, info = model.transcribe(chunks["jfk"], language="en", word_timestamps=True)
        rec["transcribe_jfk"] = {"language": info.language, "duration": float(info.duration), "segments": [
            {"start": float(s.start), "end": float(s.end), "text": s.text, "tokens": list(map(int, s.tokens)),
             "avg_logprob": float(s.avg_logprob), "no_speech_prob": float(s.no_speech_prob), "temperature": s.temperature,
             "words": [[w.word, float(w.start), float(w.end), float(w.probability)] for w in (s.words or [])]} for s in segs]}
        arrays["jfk__pcm16k"] = (np.clip(chunks["jfk"], -1, 1) * 32767).astype(np.int16)

    model_dir = getattr(model, "model_dir", None) or (a.model if os.path.isdir(a.model) else None)
    if model_dir is None:
        try:
            from faster_whisper.utils import download_model
            model_dir = download_model(a.model, local_files_only=True)
        except Exception:
            model_dir = None
    if model_dir and os.path.exists(os.path.join(model_dir, "model.bin")):
        rec["model_bin_header"] = model_bin_header(os.path.join(model_dir, "model.bin"))
        rec["meta"]["model_dir"] = model_dir
    np.savez_compressed(os.path.join(HERE, "ct2_capture.npz"), **arrays)
    with open(os.path.join(HERE, "ct2_capture.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(f"captured {len(rec['chunks'])} chunks from {a.model} -> tests/golden/ct2_capture.{{npz,json}}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
