"""StreamScheduler: one owner thread drives the GPU engine; client threads submit chunks and wait.

Same contract as the reference's ``BatchInferenceWorker`` (whisper_live/batch_inference.py:87-187:
``submit(request)``, ``request.future`` Event, ``result`` / ``info`` / ``error`` fields, the worker
survives a failing batch and propagates the exception per request) with the gaps of its
``_process_multi`` closed: chunks longer than 30 s, hotwords, word timestamps and
the temperature ladder all go through ``B200WhisperModel.transcribe_batch``.
"""
from __future__ import annotations

import logging
import queue
import threading
import time
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np


@dataclass
class BatchRequest:
    audio: np.ndarray
    language: Optional[str] = None
    task: str = "transcribe"
    initial_prompt: Optional[str] = None
    use_vad: bool = True
    vad_parameters: Optional[Dict] = None
    word_timestamps: bool = False
    client_uid: Optional[str] = None
    hotwords: Optional[str] = None
    future: threading.Event = field(default_factory=threading.Event)
    result: Optional[Any] = None
    info: Optional[Any] = None
    error: Optional[Exception] = None


class StreamScheduler:
    def __init__(self, transcriber, max_batch_size: int = 8, batch_window_ms: int = 50):
        self.transcriber = transcriber
        self.max_batch_size = max_batch_size
        self.batch_window_ms = batch_window_ms
        self._queue: "queue.Queue[BatchRequest]" = queue.Queue()
        self._stop_event = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self.batches_run = 0

    def start(self):
        self._thread = threading.Thread(target=self._worker_loop, daemon=True, name="wlb200-scheduler")
        self._thread.start()
        logging.info("[StreamScheduler] started (max_batch=%d, window=%dms)", self.max_batch_size, self.batch_window_ms)

    def stop(self):
        self._stop_event.set()
        if self._thread:
            self._thread.join(timeout=5)

    def submit(self, request: BatchRequest):
        self._queue.put(request)

    def _collect(self) -> List[BatchRequest]:
        try:
            batch = [self._queue.get(timeout=0.5)]
        except queue.Empty:
            return []
        deadline = time.monotonic() + self.batch_window_ms / 1000.0
        while len(batch) < self.max_batch_size:
            left = deadline - time.monotonic()
            if left <= 0:
                break
            try:
                batch.append(self._queue.get(timeout=left))
            except queue.Empty:
                break
        return batch

    def _worker_loop(self):
        while not self._stop_event.is_set():
            batch = self._collect()
            if not batch:
                continue
            try:
                self._process_batch(batch)
            except Exception as e:  # keep serving: fail the requests of this batch only
                logging.error("[StreamScheduler] batch failed: %s", e)
                for r in batch:
                    if not r.future.is_set():
                        r.error = e
                        r.future.set()

    def _process_batch(self, batch: List[BatchRequest]):
        kws = [dict(language=r.language, task=r.task, initial_prompt=r.initial_prompt, vad_filter=r.use_vad,
                    vad_parameters=r.vad_parameters if r.use_vad else None, hotwords=r.hotwords,
                    word_timestamps=r.word_timestamps) for r in batch]
        out = self.transcriber.transcribe_batch([r.audio for r in batch], kws)
        self.batches_run += 1
        for r, (segments, info) in zip(batch, out):
            r.result = list(segments) if segments is not None else None
            r.info = info
            r.future.set()
