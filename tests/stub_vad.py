"""Deterministic stand-in for ``faster_whisper.vad`` (TEST INFRASTRUCTURE ONLY).

The reference gates audio with Silero VAD through this module's interface
(whisper_live/transcriber/transcriber_faster_whisper.py:830-838, :1792-1817); the ONNX model and
onnxruntime are absent offline, so the H8 tests replace the *detector* with an energy gate and keep
the interface: ``VadOptions``, ``get_speech_timestamps``, ``collect_chunks`` and ``SpeechTimestampsMap``
(the last one restated from faster-whisper 1.2.0 ``vad.py`` [upstream-recalled]: it is pure
arithmetic on the chunk list and is what ``restore_speech_timestamps`` depends on).  The same module
object is installed as ``faster_whisper.vad`` for the reference's own code in
tests/golden/make_golden_transcribe.py and injected into ``B200WhisperModel(vad=...)``, so both sides
see identical chunks."""
from __future__ import annotations

import bisect
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np


@dataclass
class VadOptions:
    threshold: float = 0.5
    neg_threshold: Optional[float] = None
    min_speech_duration_ms: int = 0
    max_speech_duration_s: float = float("inf")
    min_silence_duration_ms: int = 2000
    speech_pad_ms: int = 400


def get_speech_timestamps(audio: np.ndarray, vad_options: Optional[VadOptions] = None, sampling_rate: int = 16000,
                          **kwargs) -> List[dict]:
    """Energy gate over 512-sample frames (the Silero frame size): a frame is speech when its RMS exceeds
    ``threshold * 0.02``; runs of silence shorter than ``min_silence_duration_ms`` are bridged; chunks are padded by
    ``speech_pad_ms``.  Returns ``[{"start": sample, "end": sample}]`` like the real function."""
    o = vad_options or VadOptions(**kwargs)
    frame = 512
    n = len(audio) // frame
    if n == 0:
        return []
    rms = np.sqrt((audio[:n * frame].reshape(n, frame).astype(np.float64) ** 2).mean(axis=1))
    speech = rms > o.threshold * 0.02
    min_sil = max(1, int(o.min_silence_duration_ms * sampling_rate / 1000 / frame))
    pad = int(o.speech_pad_ms * sampling_rate / 1000)
    chunks: List[dict] = []
    start, silence = None, 0
    for i, s in enumerate(speech):
        if s:
            if start is None:
                start = i
            silence = 0
        elif start is not None:
            silence += 1
            if silence >= min_sil:
                chunks.append({"start": start * frame, "end": (i - silence + 1) * frame})
                start, silence = None, 0
    if start is not None:
        chunks.append({"start": start * frame, "end": (n - silence) * frame})
    out = []
    for c in chunks:
        s, e = max(0, c["start"] - pad), min(len(audio), c["end"] + pad)
        if out and s <= out[-1]["end"]:
            out[-1]["end"] = e
        elif (e - s) * 1000 / sampling_rate >= o.min_speech_duration_ms:
            out.append({"start": s, "end": e})
    return out


def collect_chunks(audio: np.ndarray, chunks: List[dict], sampling_rate: int = 16000,
                   max_duration: float = float("inf")) -> Tuple[List[np.ndarray], List[dict]]:
    if not chunks:
        return [np.array([], dtype=np.float32)], [{"offset": 0, "duration": 0, "segments": []}]
    parts = [audio[c["start"]:c["end"]] for c in chunks]
    meta = [{"offset": c["start"] / sampling_rate, "duration": (c["end"] - c["start"]) / sampling_rate, "segments": [c]}
            for c in chunks]
    return parts, meta


class SpeechTimestampsMap:
    """Maps times on the silence-free axis back to the original audio (faster-whisper ``vad.py``)."""

    def __init__(self, chunks: List[dict], sampling_rate: int, time_precision: int = 2):
        self.sampling_rate = sampling_rate
        self.time_precision = time_precision
        self.chunk_end_sample: List[int] = []
        self.total_silence_before: List[float] = []
        previous_end = 0
        silent_samples = 0
        for chunk in chunks:
            silent_samples += chunk["start"] - previous_end
            previous_end = chunk["end"]
            self.chunk_end_sample.append(chunk["end"] - silent_samples)
            self.total_silence_before.append(silent_samples / sampling_rate)

    def get_original_time(self, time: float, chunk_index: Optional[int] = None, is_end: bool = False) -> float:
        if chunk_index is None:
            chunk_index = self.get_chunk_index(time, is_end)
        return round(self.total_silence_before[chunk_index] + time, self.time_precision)

    def get_chunk_index(self, time: float, is_end: bool = False) -> int:
        sample = int(time * self.sampling_rate)
        if sample in self.chunk_end_sample and is_end:
            return self.chunk_end_sample.index(sample)
        return min(bisect.bisect(self.chunk_end_sample, sample), len(self.chunk_end_sample) - 1)
