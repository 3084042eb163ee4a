"""CPU stand-in with the ``ctranslate2.models.Whisper`` API surface (TEST INFRASTRUCTURE ONLY).

Same methods / argument names / result fields as the engine the reference drives at
whisper_live/transcriber/transcriber_faster_whisper.py:634-643 (ctor), :1348 (encode),
:1394-1407 (generate), :1140/:1771 (detect_language), :1657-1663 (align), and as the
mock at tests/test_batch_inference.py:58-68.  Used (a) to check the CUDA engine
(whisperlive_b200.engine.B200Whisper) call by call, and (b) as the
``cpu_baseline`` / ``--impl reference`` arm of bench.py.  The product never imports it.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import align as oalign
from . import model as om
from .search import GenOptions, StreamResult, VocabSpec, search_stream

LANG_CODES = (
    "en zh de es ru ko fr ja pt tr pl ca nl ar sv it id hi fi vi he uk el ms cs ro da hu ta no th ur hr bg lt la mi "
    "ml cy sk te fa lv bn sr az sl kn et mk br eu is hy ne mn bs kk sq sw gl mr pa si km sn yo so af oc ka be tg sd "
    "gu am yi lo uz fo ht ps tk nn mt sa lb my bo tl mg as tt haw ln ha ba jw su yue").split()


@dataclass
class OracleEncoderOutput:
    enc: torch.Tensor                      # [B,1500,d] f32
    xkv: list                              # per decoder layer (K,V) [B,H,1500,64]

    @property
    def shape(self):
        return tuple(self.enc.shape)

    def select(self, indices):
        idx = torch.as_tensor(list(indices), dtype=torch.long)
        return OracleEncoderOutput(self.enc.index_select(0, idx),
                                   [(k.index_select(0, idx), v.index_select(0, idx)) for k, v in self.xkv])

    def join(self, views):
        """Batch of several (single- or multi-stream) views: the oracle's handles are tensors, so they concatenate."""
        return OracleEncoderOutput(torch.cat([v.enc for v in views], 0),
                                   [(torch.cat([v.xkv[l][0] for v in views], 0), torch.cat([v.xkv[l][1] for v in views], 0))
                                    for l in range(len(self.xkv))])

    def __array__(self, dtype=None, copy=None):
        a = self.enc.numpy()
        return a.astype(dtype) if dtype is not None else a


@dataclass
class GenerationResult:
    sequences_ids: List[List[int]]
    scores: List[float]
    no_speech_prob: float
    steps: int = 0
    margins: List[float] = field(default_factory=list)


@dataclass
class AlignmentResult:
    alignments: List[Tuple[int, int]]
    text_token_probs: List[float]


class OracleWhisper:
    def __init__(self, weights: Dict[str, torch.Tensor], dims, num_threads: Optional[int] = None):
        self.w = weights
        self.dims = dims
        self.spec = VocabSpec.from_vocab_size(dims.vocab)
        self.device = "cpu"
        self.device_index = [0]
        if num_threads:
            torch.set_num_threads(num_threads)

    @property
    def is_multilingual(self) -> bool:
        return self.dims.multilingual

    @property
    def n_mels(self) -> int:
        return self.dims.n_mels

    @property
    def vocab_size(self) -> int:
        return self.dims.vocab

    @property
    def num_languages(self) -> int:
        return self.dims.num_languages

    # -- ctranslate2.models.Whisper.encode -------------------------------------------------
    def encode(self, features, to_cpu: bool = False) -> OracleEncoderOutput:
        f = torch.as_tensor(np.asarray(features), dtype=torch.float32)
        if f.ndim == 2:
            f = f[None]
        with torch.no_grad():
            enc = om.encoder_forward(self.w, f, self.dims.n_heads, self.dims.enc_layers)
            xkv = om.cross_kv(self.w, enc, self.dims.n_heads, self.dims.dec_layers)
        return OracleEncoderOutput(enc, xkv)

    def _as_encoded(self, features) -> OracleEncoderOutput:
        return features if isinstance(features, OracleEncoderOutput) else self.encode(features)

    def _stream_step_fn(self, eo: OracleEncoderOutput, b: int):
        xkv_b = [(k[b:b + 1], v[b:b + 1]) for k, v in eo.xkv]
        state = om.DecoderState(self.dims.dec_layers)

        def step(tokens: torch.Tensor, parents):
            if parents is not None:
                state.reorder(parents)
            r = tokens.shape[0]
            with torch.no_grad():
                return om.decoder_forward(self.w, tokens, xkv_b, state, self.dims.n_heads, self.dims.dec_layers,
                                          row_to_stream=torch.zeros(r, dtype=torch.long))
        return step

    # -- ctranslate2.models.Whisper.generate -----------------------------------------------
    def generate(self, features, prompts: Sequence[Sequence[int]], *, beam_size=5, patience=1, num_hypotheses=1,
                 length_penalty=1, repetition_penalty=1, no_repeat_ngram_size=0, max_length=448,
                 return_scores=False, return_no_speech_prob=False, max_initial_timestamp_index=50,
                 suppress_blank=True, suppress_tokens=(-1,), sampling_topk=1, sampling_temperature=1,
                 seed=None, max_length_per_stream=None, trace: bool = False) -> List[GenerationResult]:
        if repetition_penalty != 1 or no_repeat_ngram_size != 0:
            raise NotImplementedError("repetition_penalty / no_repeat_ngram_size: the reference passes 1 / 0")
        if seed is None:   # same rule as whisperlive_b200.engine.B200Whisper.generate: a per-engine sampling-call counter
            if int(beam_size) == 1 and sampling_topk != 1 and sampling_temperature > 0:
                self._sampling_calls = getattr(self, "_sampling_calls", 0) + 1
            seed = getattr(self, "_sampling_calls", 0)
        eo = self._as_encoded(features)
        sup = [t for t in (suppress_tokens or ()) if t >= 0]
        out = []
        for b, prompt in enumerate(prompts):
            opts = GenOptions(beam_size=beam_size, patience=patience, num_hypotheses=num_hypotheses,
                              length_penalty=length_penalty,
                              max_length=max_length if max_length_per_stream is None else max_length_per_stream[b],
                              suppress_blank=suppress_blank,
                              suppress_tokens=sup, max_initial_timestamp_index=max_initial_timestamp_index,
                              sampling_topk=sampling_topk, sampling_temperature=sampling_temperature, seed=seed, trace=trace)
            r: StreamResult = search_stream(self._stream_step_fn(eo, b), list(prompt), self.spec, opts, stream_index=b)
            g = GenerationResult(r.sequences_ids, r.scores, r.no_speech_prob, r.steps, r.margins)
            g.trace, g.row_margins, g.row_tokens = r.trace, r.row_margins, r.row_tokens
            out.append(g)
        return out

    # -- whisperlive_b200.engine.B200Whisper.open_decode_session (CPU model of wl_session_*) -----
    def open_decode_session(self, capacity: Optional[int] = None, **generate_kwargs) -> "OracleDecodeSession":
        return OracleDecodeSession(self, capacity or getattr(self, "max_streams", 8), **generate_kwargs)

    # -- ctranslate2.models.Whisper.detect_language ----------------------------------------
    def detect_language(self, features) -> List[List[Tuple[str, float]]]:
        if not self.is_multilingual:
            raise RuntimeError("detect_language can only be called on multilingual models")
        eo = self._as_encoded(features)
        ids = self.spec.language_ids()
        res = []
        for b in range(eo.enc.shape[0]):
            logits = self._stream_step_fn(eo, b)(torch.tensor([[self.spec.sot]]), None)[0, -1]
            p = torch.softmax(logits[ids], -1).numpy()
            order = np.lexsort((np.arange(len(ids)), -p))
            res.append([(f"<|{LANG_CODES[i]}|>", float(p[i])) for i in order])
        return res

    # -- ctranslate2.models.Whisper.align ----------------------------------------------------
    def align(self, features, start_sequence: Sequence[int], text_tokens: Sequence[Sequence[int]],
              num_frames: Union[int, Sequence[int]], *, median_filter_width: int = 7) -> List[AlignmentResult]:
        eo = self._as_encoded(features)
        heads = self.dims.default_alignment_heads()
        out = []
        for b, text in enumerate(text_tokens):
            nf = num_frames if isinstance(num_frames, int) else num_frames[b]
            seq = list(start_sequence) + [self.spec.no_timestamps] + list(text) + [self.spec.eot]
            xkv_b = [(k[b:b + 1], v[b:b + 1]) for k, v in eo.xkv]
            cross: list = []
            with torch.no_grad():
                logits = om.decoder_forward(self.w, torch.tensor([seq]), xkv_b, om.DecoderState(self.dims.dec_layers),
                                            self.dims.n_heads, self.dims.dec_layers, collect_cross=cross)[0]
            n0 = len(start_sequence)
            probs = torch.softmax(logits[n0:n0 + len(text)], -1)
            tok_probs = probs[torch.arange(len(text)), torch.tensor(list(text), dtype=torch.long)].tolist() if len(text) else []
            attn = np.stack([cross[l][0, h].numpy() for (l, h) in heads])
            out.append(AlignmentResult(oalign.alignment_from_attention(attn, n0, nf, median_filter_width), tok_probs))
        return out


class OracleDecodeSession:
    """CPU model of the engine's decode session (include/wlb200.h, wl_session_*) for the host-side tests: a stream's
    hypotheses are what ``generate`` returns for it alone (streams are independent), and it occupies its index for as
    many token steps as that decode took after the (prefilled) prompt.  TEST INFRASTRUCTURE like the rest of oracle/."""

    def __init__(self, engine: OracleWhisper, capacity: int, **kw):
        if int(kw.get("beam_size", 5)) == 1 and kw.get("sampling_topk", 1) != 1 and kw.get("sampling_temperature", 0) > 0:
            raise ValueError("a decode session does not sample")
        self.engine, self.capacity = engine, int(capacity)
        self.kw = {k: v for k, v in kw.items() if k not in ("max_length", "max_length_per_stream")}
        self._res: Dict[int, GenerationResult] = {}
        self._left: Dict[int, int] = {}
        self.steps = 0
        self.runs = 0
        self.closed = False

    @property
    def live(self) -> int:
        return len(self._res)

    def free_indices(self) -> List[int]:
        return [i for i in range(self.capacity) if i not in self._res]

    def admit(self, features, prompts, max_lengths, indices=None) -> List[int]:
        free = self.free_indices()
        if indices is None:
            if len(prompts) > len(free):
                raise RuntimeError(f"admit: {len(prompts)} streams for {len(free)} free indices")
            indices = free[:len(prompts)]
        for i, f, p, ml in zip(indices, features, prompts, max_lengths):
            r = self.engine.generate(f, [list(p)], max_length=int(ml), **self.kw)[0]
            self._res[i] = r
            self._left[i] = max(1, int(r.steps) - (len(p) - 1))
        return list(indices)

    def run(self, max_steps: int = 16, break_on_finish: bool = True) -> List[int]:
        self.runs += 1
        ran = 0
        while ran < max_steps and any(v > 0 for v in self._left.values()):
            ran += 1
            newly = False
            for i in self._left:
                if self._left[i] > 0:
                    self._left[i] -= 1
                    newly = newly or self._left[i] == 0
            if newly and break_on_finish:
                break
        self.steps += ran
        self.last_steps = ran
        return sorted(i for i, v in self._left.items() if v == 0)

    def collect(self, index: int) -> GenerationResult:
        if self._left.get(index, 1) != 0:
            raise RuntimeError(f"stream index {index} has not finished")
        del self._left[index]
        return self._res.pop(index)

    def close(self) -> None:
        self.closed = True
        self._res.clear()
        self._left.clear()
