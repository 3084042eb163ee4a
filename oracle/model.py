"""Oracle K2-K11, K13: Whisper encoder / decoder arithmetic in plain torch fp32
(TEST INFRASTRUCTURE ONLY).

What the reference runs for these rows is ``ctranslate2.models.Whisper`` (call sites
whisper_live/transcriber/transcriber_faster_whisper.py:1348 encode, :1394 generate,
:1140/:1771 detect_language, :1657 align) -- a pip dependency that is not in
/root/reference and not installed.  The network itself is the published Whisper
architecture; this restatement is pinned against HF transformers 5.5.0
``models/whisper/modeling_whisper.py`` (WhisperEncoder :541, WhisperDecoder :650,
WhisperAttention :241) by tests/golden/make_golden_model.py.

Weights use HF tensor names (whisperlive_b200/weights.py).  GELU is the exact erf form.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

W = Dict[str, torch.Tensor]


def _ln(x, w: W, prefix: str):
    return F.layer_norm(x, (x.shape[-1],), w[prefix + ".weight"], w[prefix + ".bias"], 1e-5)


def _lin(x, w: W, prefix: str):
    return F.linear(x, w[prefix + ".weight"], w.get(prefix + ".bias"))


def _split_heads(x, n_heads):
    b, t, d = x.shape
    return x.view(b, t, n_heads, d // n_heads).transpose(1, 2)  # [B,H,T,64]


def _merge_heads(x):
    b, h, t, hd = x.shape
    return x.transpose(1, 2).reshape(b, t, h * hd)


def encoder_forward(w: W, features: torch.Tensor, n_heads: int, n_layers: int,
                    collect: Optional[dict] = None) -> torch.Tensor:
    """features [B, n_mels, 3000] f32 -> [B, 1500, d]  (K2-K6)."""
    p = "model.encoder"
    x = F.gelu(F.conv1d(features, w[p + ".conv1.weight"], w[p + ".conv1.bias"], padding=1))
    if collect is not None:
        collect["conv1"] = x.transpose(1, 2).clone()
    x = F.gelu(F.conv1d(x, w[p + ".conv2.weight"], w[p + ".conv2.bias"], stride=2, padding=1))
    x = x.transpose(1, 2) + w[p + ".embed_positions.weight"][None]
    if collect is not None:
        collect["stem"] = x.clone()
    scale = 64 ** -0.5
    for i in range(n_layers):
        lp = f"{p}.layers.{i}"
        h = _ln(x, w, lp + ".self_attn_layer_norm")
        q = _split_heads(_lin(h, w, lp + ".self_attn.q_proj"), n_heads)
        k = _split_heads(_lin(h, w, lp + ".self_attn.k_proj"), n_heads)
        v = _split_heads(_lin(h, w, lp + ".self_attn.v_proj"), n_heads)
        a = torch.softmax((q @ k.transpose(-1, -2)) * scale, dim=-1) @ v
        x = x + _lin(_merge_heads(a), w, lp + ".self_attn.out_proj")
        h = _ln(x, w, lp + ".final_layer_norm")
        x = x + _lin(F.gelu(_lin(h, w, lp + ".fc1")), w, lp + ".fc2")
        if collect is not None:
            collect[f"layer{i}"] = x.clone()
    return _ln(x, w, p + ".layer_norm")


def cross_kv(w: W, enc_out: torch.Tensor, n_heads: int, n_layers: int) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """K7: per decoder layer (K,V) [B,H,1500,64] from the encoder output."""
    out = []
    for i in range(n_layers):
        lp = f"model.decoder.layers.{i}.encoder_attn"
        out.append((_split_heads(_lin(enc_out, w, lp + ".k_proj"), n_heads),
                    _split_heads(_lin(enc_out, w, lp + ".v_proj"), n_heads)))
    return out


class DecoderState:
    """Self-attention KV cache for R rows: per layer (k, v) [R,H,T,64]."""

    def __init__(self, n_layers: int):
        self.k: List[Optional[torch.Tensor]] = [None] * n_layers
        self.v: List[Optional[torch.Tensor]] = [None] * n_layers

    @property
    def length(self) -> int:
        return 0 if self.k[0] is None else self.k[0].shape[2]

    def reorder(self, index: torch.Tensor) -> None:
        """Gather cache rows by parent-beam index (CT2 does this every beam step)."""
        for i in range(len(self.k)):
            if self.k[i] is not None:
                self.k[i] = self.k[i].index_select(0, index)
                self.v[i] = self.v[i].index_select(0, index)

    def clone(self) -> "DecoderState":
        s = DecoderState(len(self.k))
        s.k = [None if t is None else t.clone() for t in self.k]
        s.v = [None if t is None else t.clone() for t in self.v]
        return s


def decoder_forward(w: W, tokens: torch.Tensor, xkv: List[Tuple[torch.Tensor, torch.Tensor]],
                    state: DecoderState, n_heads: int, n_layers: int,
                    row_to_stream: Optional[torch.Tensor] = None,
                    collect_cross: Optional[list] = None,
                    return_hidden: bool = False) -> torch.Tensor:
    """tokens [R,T_new] int64 appended after ``state.length`` cached positions (K8-K12 logits).

    Returns logits [R,T_new,vocab] f32.  ``row_to_stream`` maps decoder rows (beams) to
    encoder batch entries.  ``collect_cross`` receives per-layer cross-attention
    probabilities [R,H,T_new,1500] (K14)."""
    p = "model.decoder"
    r, t_new = tokens.shape
    past = state.length
    pos = w[p + ".embed_positions.weight"][past:past + t_new]
    x = w[p + ".embed_tokens.weight"][tokens] + pos[None]
    scale = 64 ** -0.5
    causal = None
    if t_new > 1:
        total = past + t_new
        causal = torch.full((t_new, total), float("-inf"))
        causal = torch.triu(causal, diagonal=past + 1)
    for i in range(n_layers):
        lp = f"{p}.layers.{i}"
        h = _ln(x, w, lp + ".self_attn_layer_norm")
        q = _split_heads(_lin(h, w, lp + ".self_attn.q_proj"), n_heads)
        k = _split_heads(_lin(h, w, lp + ".self_attn.k_proj"), n_heads)
        v = _split_heads(_lin(h, w, lp + ".self_attn.v_proj"), n_heads)
        if state.k[i] is not None:
            k = torch.cat([state.k[i], k], dim=2)
            v = torch.cat([state.v[i], v], dim=2)
        state.k[i], state.v[i] = k, v
        s = (q @ k.transpose(-1, -2)) * scale
        if causal is not None:
            s = s + causal
        x = x + _lin(_merge_heads(torch.softmax(s, dim=-1) @ v), w, lp + ".self_attn.out_proj")

        h = _ln(x, w, lp + ".encoder_attn_layer_norm")
        q = _split_heads(_lin(h, w, lp + ".encoder_attn.q_proj"), n_heads)
        ck, cv = xkv[i]
        if row_to_stream is not None:
            ck, cv = ck.index_select(0, row_to_stream), cv.index_select(0, row_to_stream)
        pr = torch.softmax((q @ ck.transpose(-1, -2)) * scale, dim=-1)
        if collect_cross is not None:
            collect_cross.append(pr)
        x = x + _lin(_merge_heads(pr @ cv), w, lp + ".encoder_attn.out_proj")

        h = _ln(x, w, lp + ".final_layer_norm")
        x = x + _lin(F.gelu(_lin(h, w, lp + ".fc1")), w, lp + ".fc2")
    x = _ln(x, w, p + ".layer_norm")
    if return_hidden:
        return x
    return x @ w[p + ".embed_tokens.weight"].t()
