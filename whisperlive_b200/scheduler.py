"""RoundScheduler: the one thread that owns the GPU engine, fed by the per-client threads.

What the reference has in this place is ``BatchInferenceWorker`` (whisper_live/batch_inference.py:87-438): collect a
batch for a time window, run it to completion, answer, repeat -- so a chunk that arrives 10 ms after a batch started
waits for the whole batch, only the first 30 s window is batched, every fallback rung re-encodes, and hotwords / word
timestamps fall back to the unbatched path.  This scheduler shares only its *request record* with it (``BatchRequest``,
the drop-in schema a client thread fills in and waits on); the control flow is different:

* there is no batch.  A ``TranscribeSession`` (transcriber.py) holds every stream that is in flight; the owner thread
  alternates ``admit`` (everything that is in the inbox RIGHT NOW, up to the stream capacity) and ``round`` (one device
  round for everything in flight: encode the next windows, one generate call per option set, align, post-process);
* a stream is answered the moment its last window settles -- it does not wait for the streams it shared rounds with;
* admission happens between rounds, and a round is at most ``step_tokens`` TOKEN STEPS of the device-side decode loop
  (``TranscribeSession.step_round`` over the engine's decode session, ``wl_session_*``): a late chunk is encoded,
  prefilled and joins the loop of the chunks that are already decoding a few token steps after it arrived, and the index
  of a finished stream is refilled immediately;
* an engine error fails the streams it touched, never the scheduler (``TranscribeSession`` isolates them).

``linger_ms`` (default 0) optionally waits for more requests when the engine is idle and a single request arrived --
the latency / batching trade the reference hard-codes as its 50 ms window.
"""
from __future__ import annotations

import collections
import logging
import threading
import time
from dataclasses import dataclass, field
from typing import Any, Deque, Dict, List, Optional

import numpy as np

log = logging.getLogger("whisperlive_b200.scheduler")


@dataclass
class BatchRequest:
    """What a client thread submits and waits on (same fields as the reference's request record,
    whisper_live/batch_inference.py:51-84, so ``ServeClient*`` code can fill either)."""
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
    submitted_at: float = 0.0
    finished_at: float = 0.0

    def kwargs(self) -> dict:
        return dict(language=self.language, task=self.task, initial_prompt=self.initial_prompt, vad_filter=self.use_vad,
                    vad_parameters=self.vad_parameters if self.use_vad else None, hotwords=self.hotwords,
                    word_timestamps=self.word_timestamps)


class RoundScheduler:
    def __init__(self, transcriber, max_batch_size: int = 8, batch_window_ms: int = 0, linger_ms: Optional[int] = None,
                 step_tokens: Optional[int] = 16):
        """``max_batch_size``: streams in flight at once (the engine's ``max_streams``).  ``batch_window_ms`` is accepted
        for signature compatibility with the reference worker and used as ``linger_ms`` when that is not given.
        ``step_tokens``: token steps per device round (``TranscribeSession.step_round``): the inbox is looked at -- and
        a finished stream answered -- at least that often, and new streams join the decode loop already running;
        ``None`` / 0 = window-level rounds (one ``generate`` call run to completion per round)."""
        self.transcriber = transcriber
        self.step_tokens = int(step_tokens or 0)
        self.capacity = max(1, int(max_batch_size))
        self.linger_s = (batch_window_ms if linger_ms is None else linger_ms) / 1000.0
        self._inbox: Deque[BatchRequest] = collections.deque()
        self._cv = threading.Condition()
        self._stop = False
        self._thread: Optional[threading.Thread] = None
        # statistics (tests, metrics)
        self.rounds_run = 0
        self.streams_done = 0
        self.max_in_flight = 0
        self.admitted_mid_flight = 0   # streams that joined while others were already decoding

    # ------------------------------------------------------------------ client side
    def submit(self, request: BatchRequest) -> None:
        request.submitted_at = time.monotonic()
        with self._cv:
            self._inbox.append(request)
            self._cv.notify()

    def start(self) -> None:
        self._thread = threading.Thread(target=self._owner_loop, daemon=True, name="wlb200-rounds")
        self._thread.start()

    def stop(self) -> None:
        with self._cv:
            self._stop = True
            self._cv.notify()
        if self._thread is not None:
            self._thread.join(timeout=10)

    # ------------------------------------------------------------------ owner thread
    def _take(self, room: int, block: bool) -> List[BatchRequest]:
        with self._cv:
            if block:
                while not self._inbox and not self._stop:
                    self._cv.wait(timeout=0.5)
                if self.linger_s > 0 and len(self._inbox) < room and not self._stop:
                    end = time.monotonic() + self.linger_s      # idle engine, first request: optionally wait for company
                    while len(self._inbox) < room and not self._stop:
                        left = end - time.monotonic()
                        if left <= 0:
                            break
                        self._cv.wait(timeout=left)
            out = []
            while self._inbox and len(out) < room:
                out.append(self._inbox.popleft())
            return out

    def _owner_loop(self) -> None:
        session = self.transcriber.open_session() if hasattr(self.transcriber, "open_session") else _OneShotSession(self.transcriber)
        in_flight: Dict[int, BatchRequest] = {}
        while True:
            with self._cv:
                if self._stop and not in_flight and not self._inbox:
                    close = getattr(session, "close", None)
                    if close is not None:
                        close()             # hands the engine's decode session back
                    return
            room = self.capacity - len(in_flight)
            new = self._take(room, block=not in_flight) if room > 0 else []
            if new:
                if in_flight:
                    self.admitted_mid_flight += len(new)
                try:
                    handles = session.add_streams([r.audio for r in new], [r.kwargs() for r in new])
                    for h, r in zip(handles, new):
                        in_flight[h] = r
                except Exception as e:      # admission (VAD / mel / language id) failed: only these requests
                    log.error("admission failed: %s", e)
                    for r in new:
                        self._finish(r, None, None, e)
            self.max_in_flight = max(self.max_in_flight, len(in_flight))
            if not in_flight:
                continue
            try:
                if self.step_tokens > 0 and hasattr(session, "step_round"):
                    session.step_round(self.step_tokens)
                else:
                    session.round()
                self.rounds_run += 1
            except Exception as e:          # the session isolates per-stream errors; anything else fails what is in flight
                log.error("round failed: %s", e)
                for h, r in list(in_flight.items()):
                    self._finish(r, None, None, e)
                in_flight.clear()
                session = self.transcriber.open_session() if hasattr(self.transcriber, "open_session") else _OneShotSession(self.transcriber)
                continue
            for entry in session.pop_finished():
                r = in_flight.pop(entry.handle, None)
                if r is None:
                    continue
                try:
                    segments, info = session.result_of(entry)
                    self._finish(r, segments, info, None)
                except Exception as e:
                    self._finish(r, None, None, e)

    def _finish(self, r: BatchRequest, segments, info, error) -> None:
        r.result = list(segments) if segments is not None else None
        r.info = info
        r.error = error
        r.finished_at = time.monotonic()
        self.streams_done += 1
        r.future.set()


class _OneShotSession:
    """Adapter for transcribers without ``open_session`` (only ``transcribe_batch``): every round is one call."""

    def __init__(self, transcriber):
        self.t = transcriber
        self._pending: List[Any] = []
        self._done: List[Any] = []
        self._n = 0

    def add_streams(self, audios, kws):
        hs = []
        for a, k in zip(audios, kws):
            self._pending.append((self._n, a, k))
            hs.append(self._n)
            self._n += 1
        return hs

    def round(self):
        batch, self._pending = self._pending, []
        try:
            out = self.t.transcribe_batch([a for _, a, _ in batch], [k for _, _, k in batch])
            for (h, _, _), res in zip(batch, out):
                self._done.append(_Done(h, res, None))
        except Exception as e:
            for h, _, _ in batch:
                self._done.append(_Done(h, None, e))

    def pop_finished(self):
        d, self._done = self._done, []
        return d

    def result_of(self, entry):
        if entry.error is not None:
            raise entry.error
        return entry.res


class _Done:
    def __init__(self, handle, res, error):
        self.handle, self.res, self.error = handle, res, error


# the name the backend plugin and round-1 tests import
StreamScheduler = RoundScheduler
