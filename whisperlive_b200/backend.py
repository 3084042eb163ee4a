"""ServeClientB200: the WhisperLive backend plugin (Boundary A).

Subclass of the reference's own ``whisper_live.backend.base.ServeClientBase`` (audio ring buffer,
``speech_to_text`` loop, segment commit logic and the WebSocket JSON are the reference's, unchanged);
mirrors ``ServeClientFasterWhisper`` (whisper_live/backend/faster_whisper_backend.py): same ctor
signature :19-40, ``SINGLE_MODEL`` / ``BATCH_WORKER`` class attributes :15-17, ``set_language``
:180-194, ``transcribe_audio`` :196-250, ``handle_transcription_output`` :252-267, SERVER_READY with
``"backend": "faster_whisper"`` :123-131 (the stock client only collects transcripts for that backend
name: whisper_live/client.py:182).

Differences that are the point of the port: one engine per process shared by all clients, driven by a
single scheduler thread that batches the chunks of all live connections per decode step
(whisperlive_b200.scheduler.StreamScheduler); no CPU device / compute-type probing -- construction
fails loudly without the CUDA engine.
"""
from __future__ import annotations

import json
import logging
import threading

try:  # the reference package must be importable: this module is a plugin for it
    from whisper_live.backend.base import ServeClientBase
except Exception as _e:  # pragma: no cover
    ServeClientBase = None
    _IMPORT_ERROR = _e

from .scheduler import BatchRequest, StreamScheduler

if ServeClientBase is not None:

    class ServeClientB200(ServeClientBase):
        SINGLE_MODEL = None
        SINGLE_MODEL_LOCK = threading.Lock()
        BATCH_WORKER = None
        MAX_STREAMS = 8          # streams batched per decode step on this GPU
        BATCH_WINDOW_MS = 20
        MODEL_FACTORY = None     # tests inject a callable(model_name) -> transcriber

        def __init__(self, websocket, task="transcribe", device=None, language=None, client_uid=None, model="small.en",
                     initial_prompt=None, vad_parameters=None, use_vad=True, single_model=True, send_last_n_segments=10,
                     no_speech_thresh=0.45, clip_audio=False, same_output_threshold=7, cache_path="~/.cache/whisper-live/",
                     translation_queue=None, hotwords=None, diarization=None, word_timestamps=False):
            super().__init__(client_uid, websocket, send_last_n_segments, no_speech_thresh, clip_audio,
                             same_output_threshold, translation_queue, diarization, word_timestamps)
            self.cache_path = cache_path
            self.model_size_or_path = model
            self.language = "en" if (model or "").endswith("en") else language
            self.task = task
            self.initial_prompt = initial_prompt
            self.vad_parameters = vad_parameters or {"threshold": 0.5}
            self.hotwords = hotwords
            self.compute_type = "float16"
            if self.model_size_or_path is None:
                return
            try:
                cls = ServeClientB200
                with cls.SINGLE_MODEL_LOCK:
                    if cls.SINGLE_MODEL is None:
                        cls.SINGLE_MODEL = self.create_model()
                        cls.BATCH_WORKER = StreamScheduler(cls.SINGLE_MODEL, max_batch_size=cls.MAX_STREAMS,
                                                           batch_window_ms=cls.BATCH_WINDOW_MS)
                        cls.BATCH_WORKER.start()
                self.transcriber = cls.SINGLE_MODEL
            except Exception as e:
                logging.error(f"Failed to load model: {e}")
                self.websocket.send(json.dumps({"uid": self.client_uid, "status": "ERROR",
                                                "message": f"Failed to load model: {str(self.model_size_or_path)}"}))
                self.websocket.close()
                return
            self.use_vad = use_vad
            self.trans_thread = threading.Thread(target=self.speech_to_text)
            self.trans_thread.start()
            self.websocket.send(json.dumps({"uid": self.client_uid, "message": self.SERVER_READY, "backend": "faster_whisper"}))

        def create_model(self):
            """Build the shared transcriber (CUDA engine). Raises when no B200 / library is available."""
            if ServeClientB200.MODEL_FACTORY is not None:
                return ServeClientB200.MODEL_FACTORY(self.model_size_or_path)
            from .parallel import MultiDeviceWhisperModel, devices_from_env
            from .transcriber import B200WhisperModel
            devices = devices_from_env()     # WLB200_DEVICES=0,1,...: one engine context per GPU, streams placed i mod G
            if len(devices) > 1:
                return MultiDeviceWhisperModel(self.model_size_or_path, device_index=devices, device="cuda",
                                               compute_type=self.compute_type, max_streams=ServeClientB200.MAX_STREAMS)
            return B200WhisperModel(self.model_size_or_path, device="cuda", device_index=devices[0],
                                    compute_type=self.compute_type, max_streams=ServeClientB200.MAX_STREAMS)

        def set_language(self, info):
            if info.language_probability > 0.5:
                self.language = info.language
                logging.info(f"Detected language {self.language} with probability {info.language_probability}")
                self.websocket.send(json.dumps({"uid": self.client_uid, "language": self.language,
                                                "language_prob": info.language_probability}))

        def transcribe_audio(self, input_sample):
            request = BatchRequest(audio=input_sample, language=self.language, task=self.task,
                                   initial_prompt=self.initial_prompt, use_vad=self.use_vad,
                                   vad_parameters=self.vad_parameters if self.use_vad else None,
                                   word_timestamps=self.word_timestamps, client_uid=self.client_uid, hotwords=self.hotwords)
            ServeClientB200.BATCH_WORKER.submit(request)
            if not request.future.wait(timeout=30):
                raise TimeoutError("transcription request timed out after 30 s")
            if request.error:
                raise request.error
            if self.language is None and request.info is not None:
                self.set_language(request.info)
            return request.result

        def handle_transcription_output(self, result, duration):
            segments = []
            if len(result):
                self.t_start = None
                last_segment = self.update_segments(result, duration)
                segments = self.prepare_segments(last_segment)
            if len(segments):
                self.send_transcription_to_client(segments)

        @classmethod
        def shutdown(cls):
            if cls.BATCH_WORKER is not None:
                cls.BATCH_WORKER.stop()
            cls.BATCH_WORKER = None
            cls.SINGLE_MODEL = None

else:

    class ServeClientB200:  # type: ignore
        def __init__(self, *a, **k):
            raise ImportError("whisper_live (the reference package) is not importable: ServeClientB200 is a plugin "
                              f"for whisper_live.backend.base.ServeClientBase ({_IMPORT_ERROR})")
