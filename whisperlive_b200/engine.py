"""B200Whisper: the CUDA engine with the ``ctranslate2.models.Whisper`` call surface (Boundary C).

Reference call sites it is a drop-in for (whisper_live/transcriber/transcriber_faster_whisper.py):
ctor :634-643, ``encode`` :1348, ``generate`` :1394-1407, ``detect_language`` :1140 / :1771,
``align`` :1657-1663, properties ``is_multilingual`` :652, ``n_mels`` :446, ``device`` /
``device_index`` :1342; and whisper_live/batch_inference.py:271, :283, :355.  All arithmetic runs in
libwlb200.so through the C ABI (include/wlb200.h); numpy arrays are only the host-side containers.
"""
from __future__ import annotations

import ctypes as C
import threading
import weakref
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib
from .config import WhisperDims, dims_for
from .tokenizer import LANGUAGE_CODES

T_MAX = 448


@dataclass
class WhisperGenerationResult:
    sequences_ids: List[List[int]]
    scores: List[float]
    no_speech_prob: float
    steps: int = 0


@dataclass
class WhisperAlignmentResult:
    alignments: List[Tuple[int, int]]
    text_token_probs: List[float]


class _SlotRef:
    """Reference-counted ownership of encoder slots in the library's pool."""

    def __init__(self, engine: "B200Whisper", slots: List[int]):
        self.engine = engine
        self.slots = list(slots)
        self._fin = weakref.finalize(self, B200Whisper._release_slots, weakref.ref(engine), list(slots))

    def release(self) -> None:
        """Give the slots back now (idempotent); every view sharing this owner becomes invalid."""
        self._fin()
        self.slots = []


class _Borrowed:
    """Owner stand-in of a joined handle: holds references to the real owners, releases nothing."""

    def __init__(self, owners, engine):
        self.owners, self.engine, self.slots = owners, engine, []

    def release(self) -> None:
        return


class EncoderOutput:
    """Opaque handle playing the role of the ``ctranslate2.StorageView`` returned by encode():
    a list of pool slots (encoder output + cross-attention K/V resident in HBM)."""

    def __init__(self, owner: _SlotRef, slots: List[int], d_model: int):
        self._owner = owner
        self.slots = list(slots)
        self._d = d_model

    @property
    def shape(self):
        return (len(self.slots), 1500, self._d)

    def select(self, indices: Sequence[int]) -> "EncoderOutput":
        return EncoderOutput(self._owner, [self.slots[i] for i in indices], self._d)

    def join(self, views: Sequence["EncoderOutput"]) -> "EncoderOutput":
        """One batch handle over the slots of several views (of this or other encode calls).  Borrowed: the result
        keeps its parents alive but never releases their slots itself."""
        return EncoderOutput(_Borrowed([v._owner for v in views], self._owner.engine), [s_ for v in views for s_ in v.slots], self._d)

    def release(self) -> None:
        """Explicit end of life of the encoder output and ALL its views (the transcriber calls this at the end of
        every window instead of relying on reference counting)."""
        self._owner.release()
        self.slots = []

    def __len__(self):
        return len(self.slots)

    def __array__(self, dtype=None, copy=None):
        eng = self._owner.engine
        out = np.stack([eng._encoder_output(s) for s in self.slots])
        return out.astype(dtype) if dtype is not None else out


class B200Whisper:
    def __init__(self, dims: WhisperDims, weights: Dict[str, "np.ndarray"], device_index: Union[int, List[int]] = 0,
                 compute_type: str = "float16", max_streams: int = 8, max_beam: int = 5, enc_slots: Optional[int] = None,
                 alignment_heads: Optional[List[Tuple[int, int]]] = None, use_cuda_graph: bool = True):
        if compute_type not in ("float16", "default", "auto"):
            raise ValueError(f"compute_type {compute_type!r}: the B200 engine computes in float16 with fp32 accumulation")
        self.lib = _lib.load()
        self.dims = dims
        self.device = "cuda"
        self.device_index = [device_index] if isinstance(device_index, int) else list(device_index)
        self.compute_type = "float16"
        self.max_streams = max_streams
        self.max_beam = max_beam
        self.enc_slots = enc_slots if enc_slots is not None else 2 * max_streams
        self.use_cuda_graph = use_cuda_graph
        self._lock = threading.RLock()
        from .weights import _special_ids
        eot, ts_begin = _special_ids(dims)
        self.eot, self.sot = eot, eot + 1
        self.timestamp_begin = ts_begin
        self.no_timestamps = ts_begin - 1
        self.no_speech = ts_begin - 2
        n_lang = dims.num_languages if dims.multilingual else 0
        heads = alignment_heads if alignment_heads is not None else dims.default_alignment_heads()
        self.alignment_heads = [(int(l), int(h)) for l, h in heads]
        self._heads_arr = np.asarray(self.alignment_heads, dtype=np.int32).reshape(-1)
        cfg = _lib.WlConfig(
            abi_version=_lib.ABI_VERSION, device=self.device_index[0], d_model=dims.d_model, n_heads=dims.n_heads,
            enc_layers=dims.enc_layers, dec_layers=dims.dec_layers, n_mels=dims.n_mels, vocab=dims.vocab, eot=eot,
            sot=eot + 1, no_speech=self.no_speech, no_timestamps=self.no_timestamps, timestamp_begin=ts_begin, blank=220,
            lang_begin=eot + 2, n_lang=n_lang, max_streams=max_streams, max_beam=max_beam,
            enc_slots=enc_slots if enc_slots is not None else 2 * max_streams,
            n_align_heads=len(self.alignment_heads), align_heads=_lib.ptr(self._heads_arr, C.c_int32))
        ctx = C.c_void_p()
        rc = self.lib.wl_init(C.byref(cfg), C.byref(ctx))
        if rc != 0:
            raise _lib.WlError(f"wl_init failed ({rc}): {self.lib.wl_last_error(None).decode()}")
        self.ctx = ctx
        self._fin = weakref.finalize(self, self.lib.wl_destroy, ctx)
        self._load_weights(weights)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_model(cls, model_size_or_path: str, device_index=0, compute_type="float16", weights=None, seed: int = 0,
                   max_streams: int = 8, max_beam: int = 5, download_root: Optional[str] = None,
                   local_files_only: bool = False, **kw) -> "B200Whisper":
        """Resolve ``model_size_or_path`` the way the reference does (faster_whisper_backend.py:133-178,
        transcriber_faster_whisper.py:620-656): a local directory holding HF ``model.safetensors`` or the
        CTranslate2 ``model.bin``; else a size name / hub id looked up through ``huggingface_hub``
        (``Systran/faster-whisper-<size>``, honouring ``download_root`` / ``local_files_only``).  When no
        checkpoint can be found this RAISES -- a transcriber serving random weights is never built silently.
        ``weights`` may be a tensor dict, or the explicit opt-in ``"random"`` (seeded random initialisation of
        the named architecture: bench.py and the tests, which run without network or checkpoints)."""
        import os
        from . import weights as W
        if isinstance(weights, str):
            if weights != "random":
                raise ValueError(f"weights={weights!r}: pass a tensor dict, None, or the explicit opt-in 'random'")
            dims = dims_for(model_size_or_path)
            return cls(dims, W.random_init(dims, seed=seed), device_index=device_index, compute_type=compute_type,
                       max_streams=max_streams, max_beam=max_beam, **kw)
        model_dir = None
        if weights is None:
            model_dir = W.resolve_model_dir(model_size_or_path, download_root=download_root, local_files_only=local_files_only)
            weights = W.load_model_dir(model_dir)
            if kw.get("alignment_heads") is None:
                from .ct2_format import read_ct2_config
                heads = read_ct2_config(model_dir).get("alignment_heads")
                if heads:
                    kw["alignment_heads"] = heads
        try:
            dims = dims_for(model_size_or_path)
        except KeyError:
            dims = W.infer_dims(weights, str(model_size_or_path))
        eng = cls(dims, weights, device_index=device_index, compute_type=compute_type, max_streams=max_streams,
                  max_beam=max_beam, **kw)
        eng.model_dir = model_dir
        return eng

    def _load_weights(self, weights) -> None:
        from .feature_extractor import mel_filters
        items = dict(weights)
        items["mel_filters"] = mel_filters(self.dims.n_mels)
        for name, t in items.items():
            a = t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
            a = np.ascontiguousarray(a, dtype=np.float32)
            shape = np.asarray(a.shape, dtype=np.int64)
            rc = self.lib.wl_load_tensor(self.ctx, name.encode(), _lib.ptr(a, C.c_float), _lib.ptr(shape, C.c_int64), a.ndim)
            _lib.check(self.lib, self.ctx, rc, f"wl_load_tensor({name})")
        _lib.check(self.lib, self.ctx, self.lib.wl_finalize_weights(self.ctx), "wl_finalize_weights")

    # ------------------------------------------------------------------ properties (ctranslate2 names)
    @property
    def is_multilingual(self) -> bool:
        return self.dims.multilingual

    @property
    def n_mels(self) -> int:
        return self.dims.n_mels

    @property
    def num_languages(self) -> int:
        return self.dims.num_languages

    @property
    def vocab_size(self) -> int:
        return self.dims.vocab

    def kernel_launches(self) -> int:
        return int(self.lib.wl_kernel_launches(self.ctx))

    def last_device_ms(self, which: int) -> float:
        return float(self.lib.wl_last_device_ms(self.ctx, which))

    def profile_cross_attn(self, enable: bool) -> None:
        """Bracket every cross-attention launch of graph-less generate calls with CUDA events (bench.py roofline)."""
        _lib.check(self.lib, self.ctx, self.lib.wl_profile_cross_attn(self.ctx, int(bool(enable))), "wl_profile_cross_attn")

    @staticmethod
    def _release_slots(engine_ref, slots):
        eng = engine_ref()
        if eng is None or not eng._fin.alive:
            return
        arr = np.asarray(slots, dtype=np.int32)
        with eng._lock:
            eng.lib.wl_slots_release(eng.ctx, _lib.ptr(arr, C.c_int32), len(slots))

    def free_slots(self) -> int:
        return int(self.lib.wl_slots_free_count(self.ctx))

    def _encoder_output(self, slot: int) -> np.ndarray:
        out = np.empty((1500, self.dims.d_model), dtype=np.float32)
        with self._lock:
            _lib.check(self.lib, self.ctx, self.lib.wl_encoder_output(self.ctx, slot, _lib.ptr(out, C.c_float)), "wl_encoder_output")
        return out

    # ------------------------------------------------------------------ K1 (used by FeatureExtractor)
    def mel(self, waveforms: Sequence[np.ndarray]) -> List[np.ndarray]:
        """log-mel of each waveform: float32 [n_mels, len//160 + 1] (the last frame is the one callers drop)."""
        outs: List[Optional[np.ndarray]] = [None] * len(waveforms)
        for i0 in range(0, len(waveforms), self.max_streams):
            chunk = [np.ascontiguousarray(w, dtype=np.float32) for w in waveforms[i0:i0 + self.max_streams]]
            lens = [len(w) for w in chunk]
            if min(lens) <= 0:
                raise ValueError("mel: empty waveform")
            off = np.zeros(len(chunk) + 1, dtype=np.int64)
            off[1:] = np.cumsum(lens)
            frames = [n // 160 + 1 for n in lens]
            ooff = np.zeros(len(chunk) + 1, dtype=np.int64)
            ooff[1:] = np.cumsum([f * self.n_mels for f in frames])
            pcm = np.concatenate(chunk)
            out = np.empty(int(ooff[-1]), dtype=np.float32)
            with self._lock:
                rc = self.lib.wl_mel(self.ctx, _lib.ptr(pcm, C.c_float), _lib.ptr(off, C.c_int64), len(chunk),
                                     _lib.ptr(out, C.c_float), _lib.ptr(ooff, C.c_int64))
                _lib.check(self.lib, self.ctx, rc, "wl_mel")
            for j, f in enumerate(frames):
                outs[i0 + j] = out[ooff[j]:ooff[j + 1]].reshape(self.n_mels, f)
        return outs  # type: ignore

    # ------------------------------------------------------------------ resident features (K1 -> K2 without leaving HBM)
    def mel_device(self, waveforms: Sequence[np.ndarray]) -> List[int]:
        """log-mel of up to ``max_streams`` waveforms, kept on the device until the next call; returns the frame
        count of each (len // 160 + 1, the last frame being the one callers drop).  Pair with ``encode_windows``."""
        chunk = [np.ascontiguousarray(w, dtype=np.float32) for w in waveforms]
        if not chunk or len(chunk) > self.max_streams:
            raise ValueError(f"mel_device takes 1..{self.max_streams} waveforms, got {len(chunk)}")
        if min(len(w) for w in chunk) <= 0:
            raise ValueError("mel: empty waveform")
        off = np.zeros(len(chunk) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(w) for w in chunk])
        pcm = np.concatenate(chunk)
        frames = np.zeros(len(chunk), dtype=np.int32)
        with self._lock:
            rc = self.lib.wl_mel_device(self.ctx, _lib.ptr(pcm, C.c_float), _lib.ptr(off, C.c_int64), len(chunk), _lib.ptr(frames, C.c_int32))
            _lib.check(self.lib, self.ctx, rc, "wl_mel_device")
            self._resident_epoch = getattr(self, "_resident_epoch", 0) + 1
        return [int(f) for f in frames]

    def encode_windows(self, windows: Sequence[Tuple[int, int, int]]) -> EncoderOutput:
        """Encode windows ``(stream, seek, length)`` cut from the resident log-mel (zero-padded to 3000 frames on the
        device, what ``pad_or_trim`` does on the host in the reference, transcriber_faster_whisper.py:1127)."""
        slots_all: List[int] = []
        with self._lock:
            free = self.free_slots()
            if len(windows) > free:
                raise RuntimeError(f"encode: {len(windows)} windows requested but only {free} encoder slots are free")
            for b0 in range(0, len(windows), self.max_streams):
                part = windows[b0:b0 + self.max_streams]
                arr = np.asarray(part, dtype=np.int32).reshape(-1, 3)
                st, sk, ln = (np.ascontiguousarray(arr[:, i]) for i in range(3))
                slots = np.zeros(len(part), dtype=np.int32)
                rc = self.lib.wl_encode_windows(self.ctx, len(part), _lib.ptr(st, C.c_int32), _lib.ptr(sk, C.c_int32),
                                                _lib.ptr(ln, C.c_int32), _lib.ptr(slots, C.c_int32))
                if rc != 0 and slots_all:
                    a = np.asarray(slots_all, dtype=np.int32)
                    self.lib.wl_slots_release(self.ctx, _lib.ptr(a, C.c_int32), len(slots_all))
                _lib.check(self.lib, self.ctx, rc, "wl_encode_windows")
                slots_all.extend(int(x) for x in slots)
        return EncoderOutput(_SlotRef(self, slots_all), slots_all, self.dims.d_model)

    # ------------------------------------------------------------------ ctranslate2.models.Whisper.encode
    def encode(self, features, to_cpu: bool = False) -> EncoderOutput:
        f = np.ascontiguousarray(np.asarray(features), dtype=np.float32)
        if f.ndim == 2:
            f = f[None]
        if f.ndim != 3 or f.shape[1] != self.n_mels or f.shape[2] != 3000:
            raise ValueError(f"encode expects [batch, {self.n_mels}, 3000] features, got {f.shape}")
        slots_all: List[int] = []
        with self._lock:
            free = self.free_slots()
            if f.shape[0] > free:
                raise RuntimeError(f"encode: {f.shape[0]} windows requested but only {free} encoder slots are free "
                                   f"(pool of {self.enc_slots}; release earlier EncoderOutput handles or encode in groups "
                                   f"of at most max_streams={self.max_streams})")
            for b0 in range(0, f.shape[0], self.max_streams):
                part = f[b0:b0 + self.max_streams]
                slots = np.zeros(part.shape[0], dtype=np.int32)
                rc = self.lib.wl_encode(self.ctx, _lib.ptr(part, C.c_float), part.shape[0], _lib.ptr(slots, C.c_int32))
                if rc != 0 and slots_all:
                    arr = np.asarray(slots_all, dtype=np.int32)
                    self.lib.wl_slots_release(self.ctx, _lib.ptr(arr, C.c_int32), len(slots_all))
                _lib.check(self.lib, self.ctx, rc, "wl_encode")
                slots_all.extend(int(s) for s in slots)
        return EncoderOutput(_SlotRef(self, slots_all), slots_all, self.dims.d_model)

    def _as_encoded(self, features) -> EncoderOutput:
        return features if isinstance(features, EncoderOutput) else self.encode(features)

    # ------------------------------------------------------------------ ctranslate2.models.Whisper.generate
    def generate(self, features, prompts: Sequence[Sequence[int]], *, beam_size: int = 5, patience: float = 1,
                 num_hypotheses: int = 1, length_penalty: float = 1, repetition_penalty: float = 1,
                 no_repeat_ngram_size: int = 0, max_length: int = 448, return_scores: bool = False,
                 return_no_speech_prob: bool = False, max_initial_timestamp_index: int = 50, suppress_blank: bool = True,
                 suppress_tokens: Optional[Sequence[int]] = (-1,), sampling_topk: int = 1, sampling_temperature: float = 1,
                 seed: Optional[int] = None, max_length_per_stream: Optional[Sequence[int]] = None,
                 prefill: Optional[bool] = None) -> List[WhisperGenerationResult]:
        if repetition_penalty != 1 or no_repeat_ngram_size != 0:
            raise NotImplementedError("repetition_penalty / no_repeat_ngram_size other than the reference's 1 / 0")
        if seed is None:
            # like CT2's generator state, the noise advances from one sampling call to the next: the rungs of the
            # temperature ladder and successive windows never replay each other's draws
            if int(beam_size) == 1 and sampling_topk != 1 and sampling_temperature > 0:
                self._sampling_calls = getattr(self, "_sampling_calls", 0) + 1
            seed = getattr(self, "_sampling_calls", 0)
        enc = self._as_encoded(features)
        if len(prompts) != len(enc):
            raise ValueError(f"{len(prompts)} prompts for {len(enc)} encoded streams")
        sup = np.asarray(sorted({int(t) for t in (suppress_tokens or ()) if t >= 0}), dtype=np.int32)
        results: List[WhisperGenerationResult] = []
        NH = int(num_hypotheses)
        for b0 in range(0, len(prompts), self.max_streams):
            ps = [list(map(int, p)) for p in prompts[b0:b0 + self.max_streams]]
            B = len(ps)
            off = np.zeros(B + 1, dtype=np.int32)
            off[1:] = np.cumsum([len(p) for p in ps])
            flat = np.asarray([t for p in ps for t in p], dtype=np.int32)
            slots = np.asarray(enc.slots[b0:b0 + B], dtype=np.int32)
            opts = _lib.WlGenOpts(
                beam_size=int(beam_size), patience=float(patience), num_hypotheses=NH, length_penalty=float(length_penalty),
                max_length=int(max_length), suppress_blank=int(bool(suppress_blank)),
                max_initial_timestamp_index=int(max_initial_timestamp_index), sampling_topk=int(sampling_topk),
                sampling_temperature=float(sampling_temperature), seed=int(seed) & 0xFFFFFFFF,
                suppress_tokens=_lib.ptr(sup, C.c_int32) if len(sup) else None, n_suppress=len(sup),
                use_cuda_graph=int(self.use_cuda_graph), max_length_per_stream=None,
                prefill=0 if prefill is None else (1 if prefill else 2))
            mlps = None
            if max_length_per_stream is not None:
                mlps = np.asarray(list(max_length_per_stream)[b0:b0 + B], dtype=np.int32)
                opts.max_length_per_stream = _lib.ptr(mlps, C.c_int32)
            ids = np.zeros((B, NH, T_MAX), dtype=np.int32)
            lens = np.zeros((B, NH), dtype=np.int32)
            score = np.zeros((B, NH), dtype=np.float32)
            nsp = np.zeros(B, dtype=np.float32)
            steps = np.zeros(B, dtype=np.int32)
            with self._lock:
                rc = self.lib.wl_generate(self.ctx, _lib.ptr(slots, C.c_int32), B, _lib.ptr(flat, C.c_int32),
                                          _lib.ptr(off, C.c_int32), C.byref(opts), _lib.ptr(ids, C.c_int32),
                                          _lib.ptr(lens, C.c_int32), _lib.ptr(score, C.c_float), _lib.ptr(nsp, C.c_float),
                                          _lib.ptr(steps, C.c_int32))
                _lib.check(self.lib, self.ctx, rc, "wl_generate")
            for b in range(B):
                seqs, scs = [], []
                for h in range(NH):
                    if lens[b, h] >= 0:
                        seqs.append(ids[b, h, :lens[b, h]].tolist())
                        scs.append(float(score[b, h]))
                results.append(WhisperGenerationResult(seqs, scs, float(nsp[b]), int(steps[b])))
        self.last_steps = max((r.steps for r in results), default=0)
        return results

    # ------------------------------------------------------------------ N2: decode session (step-level admission)
    def open_decode_session(self, capacity: Optional[int] = None, **generate_kwargs) -> "DecodeSession":
        """A decode loop whose streams come and go independently (``wl_session_*``): same keyword arguments as
        ``generate`` (``max_length`` is given per stream at admission; sampling options are not accepted)."""
        # one session per engine context: whatever an earlier owner left behind (a scheduler stopped mid-decode) is dropped
        with self._lock:
            rc = self.lib.wl_session_close(self.ctx)
            _lib.check(self.lib, self.ctx, rc, "wl_session_close")
        return DecodeSession(self, capacity or self.max_streams, **generate_kwargs)

    # ------------------------------------------------------------------ ctranslate2.models.Whisper.detect_language
    def detect_language(self, features) -> List[List[Tuple[str, float]]]:
        if not self.is_multilingual:
            raise RuntimeError("detect_language can only be called on multilingual models")
        enc = self._as_encoded(features)
        n_lang = self.num_languages
        out: List[List[Tuple[str, float]]] = []
        for b0 in range(0, len(enc), self.max_streams):
            slots = np.asarray(enc.slots[b0:b0 + self.max_streams], dtype=np.int32)
            probs = np.zeros((len(slots), n_lang), dtype=np.float32)
            with self._lock:
                rc = self.lib.wl_detect_language(self.ctx, _lib.ptr(slots, C.c_int32), len(slots), _lib.ptr(probs, C.c_float))
                _lib.check(self.lib, self.ctx, rc, "wl_detect_language")
            for p in probs:
                order = np.lexsort((np.arange(n_lang), -p))
                out.append([(f"<|{LANGUAGE_CODES[i]}|>", float(p[i])) for i in order])
        return out

    # ------------------------------------------------------------------ ctranslate2.models.Whisper.align
    def align(self, features, start_sequence: Sequence[int], text_tokens: Sequence[Sequence[int]],
              num_frames: Union[int, Sequence[int]], *, median_filter_width: int = 7) -> List[WhisperAlignmentResult]:
        enc = self._as_encoded(features)
        if len(text_tokens) != len(enc):
            raise ValueError(f"{len(text_tokens)} token lists for {len(enc)} encoded streams")
        out: List[WhisperAlignmentResult] = []
        start = np.asarray(list(start_sequence), dtype=np.int32)
        for b0 in range(0, len(enc), self.max_streams):
            tt = [list(map(int, t)) for t in text_tokens[b0:b0 + self.max_streams]]
            B = len(tt)
            slots = np.asarray(enc.slots[b0:b0 + B], dtype=np.int32)
            toff = np.zeros(B + 1, dtype=np.int32)
            toff[1:] = np.cumsum([len(t) for t in tt])
            flat = np.asarray([x for t in tt for x in t] or [0], dtype=np.int32)
            nf = np.asarray([num_frames] * B if isinstance(num_frames, (int, np.integer)) else list(num_frames)[b0:b0 + B],
                            dtype=np.int32)
            cap = int(sum(len(t) + 1 + max(1, int(f) // 2) for t, f in zip(tt, nf)) + 8)
            pairs = np.zeros((cap, 2), dtype=np.int32)
            poff = np.zeros(B + 1, dtype=np.int32)
            probs = np.zeros(max(1, int(toff[-1])), dtype=np.float32)
            with self._lock:
                rc = self.lib.wl_align(self.ctx, _lib.ptr(slots, C.c_int32), B, _lib.ptr(start, C.c_int32), len(start),
                                       _lib.ptr(flat, C.c_int32), _lib.ptr(toff, C.c_int32), _lib.ptr(nf, C.c_int32),
                                       int(median_filter_width), _lib.ptr(pairs, C.c_int32), cap, _lib.ptr(poff, C.c_int32),
                                       _lib.ptr(probs, C.c_float))
                _lib.check(self.lib, self.ctx, rc, "wl_align")
            for b in range(B):
                al = [(int(a), int(t)) for a, t in pairs[poff[b]:poff[b + 1]]]
                out.append(WhisperAlignmentResult(al, probs[toff[b]:toff[b + 1]].tolist()))
        return out

    # ------------------------------------------------------------------ parity hooks
    def decode_logits(self, features, token_lists: Sequence[Sequence[int]]) -> List[np.ndarray]:
        """Teacher-forced logits [T, vocab] per stream (test hook: wl_decode_logits)."""
        enc = self._as_encoded(features)
        tl = [list(map(int, t)) for t in token_lists]
        B = len(tl)
        off = np.zeros(B + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(t) for t in tl])
        flat = np.asarray([x for t in tl for x in t], dtype=np.int32)
        slots = np.asarray(enc.slots, dtype=np.int32)
        out = np.zeros((int(off[-1]), self.dims.vocab), dtype=np.float32)
        with self._lock:
            rc = self.lib.wl_decode_logits(self.ctx, _lib.ptr(slots, C.c_int32), B, _lib.ptr(flat, C.c_int32),
                                           _lib.ptr(off, C.c_int32), _lib.ptr(out, C.c_float))
            _lib.check(self.lib, self.ctx, rc, "wl_decode_logits")
        return [out[off[b]:off[b + 1]] for b in range(B)]

    def test_wgemm(self, w: np.ndarray, x: np.ndarray, bias: Optional[np.ndarray] = None, mode: int = 0,
                   resid: Optional[np.ndarray] = None) -> np.ndarray:
        """Y[R, n_out] = X[R, K] W[n_out, K]^T through the small-batch decode GEMM (wl_test_wgemm); ``mode | 8`` runs the
        cluster split-K GEMM (any row count) with epilogue ``mode``."""
        w16 = np.ascontiguousarray(w, dtype=np.float16)
        x16 = np.ascontiguousarray(x, dtype=np.float16)
        n_out, K = w16.shape
        R = x16.shape[0]
        out = np.zeros((R, n_out), dtype=np.float32) if resid is None else np.ascontiguousarray(resid, dtype=np.float32).copy()
        bp = None
        if bias is not None:
            bias = np.ascontiguousarray(bias, dtype=np.float32)
            bp = _lib.ptr(bias, C.c_float)
        with self._lock:
            rc = self.lib.wl_test_wgemm(self.ctx, _lib.ptr(w16.view(np.uint16), C.c_uint16), _lib.ptr(x16.view(np.uint16), C.c_uint16),
                                        bp, _lib.ptr(out, C.c_float), R, n_out, K, int(mode))
            _lib.check(self.lib, self.ctx, rc, "wl_test_wgemm")
        return out

    def test_gemm(self, a: np.ndarray, b: np.ndarray, bias: Optional[np.ndarray] = None, transposed_store: bool = False,
                  gelu: bool = False, use_simt: bool = False) -> np.ndarray:
        """C[z] = A[z] @ B[z]^T through the tcgen05 kernel (or the CUDA-core checker)."""
        a16 = np.ascontiguousarray(a, dtype=np.float16)
        b16 = np.ascontiguousarray(b, dtype=np.float16)
        if a16.ndim == 2:
            a16, b16 = a16[None], b16[None]
        Z, M, K = a16.shape
        N = b16.shape[1]
        c = np.zeros((Z, N, M) if transposed_store else (Z, M, N), dtype=np.float32)
        bp = None
        if bias is not None:
            bias = np.ascontiguousarray(bias, dtype=np.float32)
            bp = _lib.ptr(bias, C.c_float)
        with self._lock:
            rc = self.lib.wl_test_gemm(self.ctx, _lib.ptr(a16.view(np.uint16), C.c_uint16), _lib.ptr(b16.view(np.uint16), C.c_uint16),
                                       bp, _lib.ptr(c, C.c_float), M, N, K, Z, int(transposed_store), int(gelu), int(use_simt))
            _lib.check(self.lib, self.ctx, rc, "wl_test_gemm")
        return c


class DecodeSession:
    """Step-level continuous batching on one engine context (``include/wlb200.h``: ``wl_session_*``).

    The reference's batcher runs a batch to completion before it looks at the queue again
    (whisper_live/batch_inference.py:155-187); here a stream is admitted at any token-step boundary into a free index
    of the running decode loop, and a finished stream is collected while the others keep decoding:

        sess = engine.open_decode_session(beam_size=5, suppress_tokens=...)
        idx = sess.admit([enc_a, enc_b], [prompt_a, prompt_b], [448, 448])
        while sess.live:
            for i in sess.run(max_steps=16):        # returns early when a stream finishes
                result = sess.collect(i)            # WhisperGenerationResult, the index is free again
            ... admit whoever arrived meanwhile ...

    One-shot engine calls (``encode``, ``generate`` for a temperature-fallback retry, ``align``, ``detect_language``)
    may be interleaved between two ``run`` calls: the session owns its decode state and self-attention cache."""

    def __init__(self, engine: B200Whisper, capacity: int, *, beam_size: int = 5, patience: float = 1, num_hypotheses: int = 1,
                 length_penalty: float = 1, repetition_penalty: float = 1, no_repeat_ngram_size: int = 0,
                 max_initial_timestamp_index: int = 50, suppress_blank: bool = True,
                 suppress_tokens: Optional[Sequence[int]] = (-1,), sampling_topk: int = 1, sampling_temperature: float = 1,
                 return_scores: bool = False, return_no_speech_prob: bool = False, max_length: int = T_MAX, **_ignored):
        if repetition_penalty != 1 or no_repeat_ngram_size != 0:
            raise NotImplementedError("repetition_penalty / no_repeat_ngram_size other than the reference's 1 / 0")
        if int(beam_size) == 1 and sampling_topk != 1 and sampling_temperature > 0:
            raise ValueError("a decode session does not sample: run temperature-fallback retries through generate()")
        self.engine = engine
        self.capacity = int(capacity)
        self.num_hypotheses = int(num_hypotheses)
        self._sup = np.asarray(sorted({int(t) for t in (suppress_tokens or ()) if t >= 0}), dtype=np.int32)
        self._opts = _lib.WlGenOpts(
            beam_size=int(beam_size), patience=float(patience), num_hypotheses=self.num_hypotheses,
            length_penalty=float(length_penalty), max_length=int(max_length), suppress_blank=int(bool(suppress_blank)),
            max_initial_timestamp_index=int(max_initial_timestamp_index), sampling_topk=1, sampling_temperature=1.0, seed=0,
            suppress_tokens=_lib.ptr(self._sup, C.c_int32) if len(self._sup) else None, n_suppress=len(self._sup),
            use_cuda_graph=int(engine.use_cuda_graph), max_length_per_stream=None, prefill=1)
        self._held: Dict[int, EncoderOutput] = {}     # index -> the encoder view its stream decodes against
        self._finished: List[int] = []
        self.steps = 0
        self.runs = 0
        self.closed = False
        with engine._lock:
            rc = engine.lib.wl_session_open(engine.ctx, C.byref(self._opts), self.capacity)
            _lib.check(engine.lib, engine.ctx, rc, "wl_session_open")

    # -- bookkeeping -------------------------------------------------------------------------------
    @property
    def live(self) -> int:
        """streams admitted and not yet collected"""
        return len(self._held)

    def free_indices(self) -> List[int]:
        return [i for i in range(self.capacity) if i not in self._held]

    # -- admission ---------------------------------------------------------------------------------
    def admit(self, features: Sequence[EncoderOutput], prompts: Sequence[Sequence[int]], max_lengths: Sequence[int],
              indices: Optional[Sequence[int]] = None) -> List[int]:
        """Admit one stream per (single-stream encoder view, prompt, max_length); returns the indices they decode in."""
        n = len(prompts)
        if n == 0:
            return []
        if len(features) != n or len(max_lengths) != n:
            raise ValueError("admit: features / prompts / max_lengths differ in length")
        free = self.free_indices()
        if indices is None:
            if n > len(free):
                raise RuntimeError(f"admit: {n} streams for {len(free)} free indices")
            indices = free[:n]
        slots = []
        for f in features:
            if not isinstance(f, EncoderOutput) or len(f) != 1:
                raise TypeError("admit: every stream needs its own single-stream EncoderOutput view")
            slots.append(int(f.slots[0]))
        ps = [list(map(int, p)) for p in prompts]
        off = np.zeros(n + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(p) for p in ps])
        flat = np.asarray([t for p in ps for t in p], dtype=np.int32)
        idx = np.asarray(list(indices), dtype=np.int32)
        sl = np.asarray(slots, dtype=np.int32)
        ml = np.asarray(list(max_lengths), dtype=np.int32)
        eng = self.engine
        with eng._lock:
            rc = eng.lib.wl_session_admit(eng.ctx, n, _lib.ptr(idx, C.c_int32), _lib.ptr(sl, C.c_int32), _lib.ptr(flat, C.c_int32),
                                          _lib.ptr(off, C.c_int32), _lib.ptr(ml, C.c_int32))
            _lib.check(eng.lib, eng.ctx, rc, "wl_session_admit")
        for i, f in zip(idx.tolist(), features):
            self._held[i] = f
        return idx.tolist()

    # -- the token loop ----------------------------------------------------------------------------
    def run(self, max_steps: int = 16, break_on_finish: bool = True) -> List[int]:
        """Up to ``max_steps`` token steps over every admitted stream; returns the indices that are finished and waiting
        to be collected."""
        eng = self.engine
        done = np.zeros(self.capacity, dtype=np.int32)
        ran = C.c_int32(0)
        with eng._lock:
            rc = eng.lib.wl_session_run(eng.ctx, int(max_steps), int(bool(break_on_finish)), _lib.ptr(done, C.c_int32), C.byref(ran))
            _lib.check(eng.lib, eng.ctx, rc, "wl_session_run")
        self.steps += int(ran.value)
        self.runs += 1
        self.last_steps = int(ran.value)
        self._finished = [i for i in range(self.capacity) if done[i]]
        return list(self._finished)

    def collect(self, index: int) -> WhisperGenerationResult:
        eng = self.engine
        NH = self.num_hypotheses
        ids = np.zeros((NH, T_MAX), dtype=np.int32)
        lens = np.zeros(NH, dtype=np.int32)
        score = np.zeros(NH, dtype=np.float32)
        nsp = C.c_float(0.0)
        steps = C.c_int32(0)
        with eng._lock:
            rc = eng.lib.wl_session_collect(eng.ctx, int(index), _lib.ptr(ids, C.c_int32), _lib.ptr(lens, C.c_int32),
                                            _lib.ptr(score, C.c_float), C.byref(nsp), C.byref(steps))
            _lib.check(eng.lib, eng.ctx, rc, "wl_session_collect")
        self._held.pop(int(index), None)
        seqs, scs = [], []
        for h in range(NH):
            if lens[h] >= 0:
                seqs.append(ids[h, :lens[h]].tolist())
                scs.append(float(score[h]))
        return WhisperGenerationResult(seqs, scs, float(nsp.value), int(steps.value))

    def close(self) -> None:
        if self.closed:
            return
        self.closed = True
        eng = self.engine
        with eng._lock:
            rc = eng.lib.wl_session_close(eng.ctx)
            _lib.check(eng.lib, eng.ctx, rc, "wl_session_close")
        self._held.clear()
