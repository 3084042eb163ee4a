"""CTranslate2 ``model.bin`` reader / writer for Whisper checkpoints.

The reference loads ``<model_dir>/model.bin`` + ``config.json`` through ``ctranslate2.models.Whisper(model_path, ...)``
(/root/reference/whisper_live/transcriber/transcriber_faster_whisper.py:634-643; the directory comes from
``download_model`` / a local path, backend/faster_whisper_backend.py:133-178).  CTranslate2 is not vendored in the
reference tree and not installed here, so the container format below is RESTATED FROM THE PUBLISHED SOURCE FROM MEMORY
(ctranslate2 4.x ``python/ctranslate2/specs/model_spec.py::ModelSpec._serialize``, binary version 6, and the variable
names produced by ``specs/whisper_spec.py`` + ``transformer_spec.py`` + ``attention_spec.py``).  It is exercised by a
write -> read round trip only (tests/test_transcriber_host.py); it has NOT been checked against a real converted model.
Anything unexpected in a file makes the reader fail loudly instead of guessing.

Layout (little endian):
    u32 binary_version (6) | str spec_name ("WhisperSpec") | u32 spec_revision | u32 n_variables
    n_variables x { str name | u8 rank | rank x u32 dim | u8 dtype_id | u32 n_bytes | bytes }
    u32 n_aliases | n_aliases x { str alias | str variable_name }
    str = u16 (len + 1) | utf-8 bytes | NUL
dtype ids (ctranslate2 ``DataType``): 0 float32, 1 int8, 2 int16, 3 int32, 4 float16, 5 bfloat16.
"""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

BINARY_VERSION = 6
_DTYPES = {0: np.float32, 1: np.int8, 2: np.int16, 3: np.int32, 4: np.float16}
_DTYPE_IDS = {np.dtype(np.float32): 0, np.dtype(np.int8): 1, np.dtype(np.int16): 2, np.dtype(np.int32): 3,
              np.dtype(np.float16): 4}
_BF16 = 5


def _read_str(f) -> str:
    (n,) = struct.unpack("<H", f.read(2))
    raw = f.read(n)
    if len(raw) != n or n == 0 or raw[-1] != 0:
        raise ValueError("model.bin: malformed string field")
    return raw[:-1].decode("utf-8")


def _write_str(f, s: str) -> None:
    b = s.encode("utf-8")
    f.write(struct.pack("<H", len(b) + 1))
    f.write(b)
    f.write(b"\0")


def read_variables(path: str) -> Tuple[Dict[str, np.ndarray], Dict[str, str], Dict[str, object]]:
    """Return (variables, aliases, header) of a CTranslate2 model.bin.  bfloat16 payloads come back as float32."""
    variables: Dict[str, np.ndarray] = {}
    with open(path, "rb") as f:
        (version,) = struct.unpack("<I", f.read(4))
        if version != BINARY_VERSION:
            raise ValueError(f"model.bin: binary version {version}, this reader understands {BINARY_VERSION} only")
        spec = _read_str(f)
        revision, n_var = struct.unpack("<II", f.read(8))
        for _ in range(n_var):
            name = _read_str(f)
            (rank,) = struct.unpack("<B", f.read(1))
            shape = struct.unpack(f"<{rank}I", f.read(4 * rank)) if rank else ()
            dtype_id, n_bytes = struct.unpack("<BI", f.read(5))
            raw = f.read(n_bytes)
            if len(raw) != n_bytes:
                raise ValueError(f"model.bin: variable {name!r} truncated")
            count = int(np.prod(shape)) if rank else 1
            if dtype_id == _BF16:
                if n_bytes != 2 * count:
                    raise ValueError(f"model.bin: variable {name!r}: {n_bytes} bytes for {count} bfloat16 values")
                u = np.frombuffer(raw, dtype=np.uint16).astype(np.uint32) << 16
                arr = u.view(np.float32).reshape(shape)
            elif dtype_id in _DTYPES:
                dt = np.dtype(_DTYPES[dtype_id])
                if n_bytes != dt.itemsize * count:
                    raise ValueError(f"model.bin: variable {name!r}: {n_bytes} bytes for {count} x {dt}")
                arr = np.frombuffer(raw, dtype=dt).reshape(shape).copy()
            else:
                raise ValueError(f"model.bin: variable {name!r} has unknown dtype id {dtype_id}")
            variables[name] = arr
        aliases: Dict[str, str] = {}
        tail = f.read(4)
        if tail:
            (n_alias,) = struct.unpack("<I", tail)
            for _ in range(n_alias):
                alias = _read_str(f)
                aliases[alias] = _read_str(f)
        if f.read(1):
            raise ValueError("model.bin: trailing bytes after the alias table")
    for alias, target in aliases.items():
        if target not in variables:
            raise ValueError(f"model.bin: alias {alias!r} points at missing variable {target!r}")
    return variables, aliases, {"spec": spec, "revision": revision, "version": version}


def write_variables(path: str, variables: Dict[str, np.ndarray], aliases: Optional[Dict[str, str]] = None,
                    spec: str = "WhisperSpec", revision: int = 3) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<I", BINARY_VERSION))
        _write_str(f, spec)
        f.write(struct.pack("<II", revision, len(variables)))
        for name in sorted(variables):
            arr = np.ascontiguousarray(variables[name])
            if arr.dtype not in _DTYPE_IDS:
                raise ValueError(f"cannot serialise {name!r} of dtype {arr.dtype}")
            _write_str(f, name)
            f.write(struct.pack("<B", arr.ndim))
            for d in arr.shape:
                f.write(struct.pack("<I", d))
            f.write(struct.pack("<BI", _DTYPE_IDS[arr.dtype], arr.nbytes))
            f.write(arr.tobytes())
        aliases = aliases or {}
        f.write(struct.pack("<I", len(aliases)))
        for alias in sorted(aliases):
            _write_str(f, alias)
            _write_str(f, aliases[alias])


# ------------------------------------------------------------------------------------------ name mapping
def _f32(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def _attention_to_hf(get, src: str, dst: str, out: Dict[str, torch.Tensor], cross: bool) -> None:
    """CT2 fuses the projections: self-attention linear_0 = [q; k; v], linear_1 = out; cross-attention linear_0 = q,
    linear_1 = [k; v], linear_2 = out.  Whisper's k_proj has no bias (CT2 stores zeros in its slice)."""
    if cross:
        wq, bq = get(f"{src}/linear_0/weight"), get(f"{src}/linear_0/bias")
        wkv, bkv = get(f"{src}/linear_1/weight"), get(f"{src}/linear_1/bias")
        d = wq.shape[0]
        wk, wv, bv = wkv[:d], wkv[d:], bkv[d:]
        wo, bo = get(f"{src}/linear_2/weight"), get(f"{src}/linear_2/bias")
    else:
        w, b = get(f"{src}/linear_0/weight"), get(f"{src}/linear_0/bias")
        d = w.shape[0] // 3
        wq, wk, wv = w[:d], w[d:2 * d], w[2 * d:]
        bq, bv = b[:d], b[2 * d:]
        wo, bo = get(f"{src}/linear_1/weight"), get(f"{src}/linear_1/bias")
    out[f"{dst}.q_proj.weight"], out[f"{dst}.q_proj.bias"] = _f32(wq), _f32(bq)
    out[f"{dst}.k_proj.weight"] = _f32(wk)
    out[f"{dst}.v_proj.weight"], out[f"{dst}.v_proj.bias"] = _f32(wv), _f32(bv)
    out[f"{dst}.out_proj.weight"], out[f"{dst}.out_proj.bias"] = _f32(wo), _f32(bo)


def _norm_to_hf(get, src: str, dst: str, out: Dict[str, torch.Tensor]) -> None:
    out[f"{dst}.weight"], out[f"{dst}.bias"] = _f32(get(f"{src}/gamma")), _f32(get(f"{src}/beta"))


def load_ct2_model_bin(path: str) -> Dict[str, torch.Tensor]:
    """``model.bin`` (or its directory) -> the canonical HF-named fp32 dict the engine uploads (weights.load_safetensors
    produces the same key space).  int8 / int16 quantised checkpoints are rejected: the engine computes in fp16."""
    if os.path.isdir(path):
        path = os.path.join(path, "model.bin")
    variables, aliases, header = read_variables(path)
    if header["spec"] != "WhisperSpec":
        raise ValueError(f"model.bin holds a {header['spec']!r}, not a WhisperSpec")

    def get(name: str) -> np.ndarray:
        key = aliases.get(name, name)
        if key not in variables:
            raise KeyError(f"model.bin: variable {name!r} is missing")
        arr = variables[key]
        if arr.dtype in (np.int8, np.int16) and arr.ndim >= 1:
            raise ValueError(f"model.bin: {name!r} is quantised ({arr.dtype}); convert with --quantization float16")
        return arr

    out: Dict[str, torch.Tensor] = {}
    for conv in ("conv1", "conv2"):
        out[f"model.encoder.{conv}.weight"] = _f32(get(f"encoder/{conv}/weight"))
        out[f"model.encoder.{conv}.bias"] = _f32(get(f"encoder/{conv}/bias"))
    out["model.encoder.embed_positions.weight"] = _f32(get("encoder/position_encodings/encodings"))
    _norm_to_hf(get, "encoder/layer_norm", "model.encoder.layer_norm", out)
    n_enc = 0
    while f"encoder/layer_{n_enc}/self_attention/linear_0/weight" in variables:
        s, h = f"encoder/layer_{n_enc}", f"model.encoder.layers.{n_enc}"
        _norm_to_hf(get, f"{s}/self_attention/layer_norm", f"{h}.self_attn_layer_norm", out)
        _attention_to_hf(get, f"{s}/self_attention", f"{h}.self_attn", out, cross=False)
        _norm_to_hf(get, f"{s}/ffn/layer_norm", f"{h}.final_layer_norm", out)
        for i, fc in enumerate(("fc1", "fc2")):
            out[f"{h}.{fc}.weight"], out[f"{h}.{fc}.bias"] = _f32(get(f"{s}/ffn/linear_{i}/weight")), _f32(get(f"{s}/ffn/linear_{i}/bias"))
        n_enc += 1
    out["model.decoder.embed_tokens.weight"] = _f32(get("decoder/embeddings/weight"))
    out["model.decoder.embed_positions.weight"] = _f32(get("decoder/position_encodings/encodings"))
    _norm_to_hf(get, "decoder/layer_norm", "model.decoder.layer_norm", out)
    n_dec = 0
    while f"decoder/layer_{n_dec}/self_attention/linear_0/weight" in variables:
        s, h = f"decoder/layer_{n_dec}", f"model.decoder.layers.{n_dec}"
        _norm_to_hf(get, f"{s}/self_attention/layer_norm", f"{h}.self_attn_layer_norm", out)
        _attention_to_hf(get, f"{s}/self_attention", f"{h}.self_attn", out, cross=False)
        _norm_to_hf(get, f"{s}/attention/layer_norm", f"{h}.encoder_attn_layer_norm", out)
        _attention_to_hf(get, f"{s}/attention", f"{h}.encoder_attn", out, cross=True)
        _norm_to_hf(get, f"{s}/ffn/layer_norm", f"{h}.final_layer_norm", out)
        for i, fc in enumerate(("fc1", "fc2")):
            out[f"{h}.{fc}.weight"], out[f"{h}.{fc}.bias"] = _f32(get(f"{s}/ffn/linear_{i}/weight")), _f32(get(f"{s}/ffn/linear_{i}/bias"))
        n_dec += 1
    if n_enc == 0 or n_dec == 0:
        raise ValueError("model.bin: no encoder / decoder layers found under the expected variable names")
    return out


def expected_ct2_names(n_enc: int, n_dec: int) -> List[str]:
    """Every variable name ``load_ct2_model_bin`` reads for a Whisper with n_enc / n_dec layers -- compared against the
    variable table of a REAL converted model by tests/test_ct2_capture.py (the container layout and these names are
    restated from memory; that test is what validates them)."""
    names = ["encoder/conv1/weight", "encoder/conv1/bias", "encoder/conv2/weight", "encoder/conv2/bias",
             "encoder/position_encodings/encodings", "encoder/layer_norm/gamma", "encoder/layer_norm/beta",
             "decoder/embeddings/weight", "decoder/position_encodings/encodings", "decoder/layer_norm/gamma", "decoder/layer_norm/beta"]

    def norm(p):
        return [f"{p}/gamma", f"{p}/beta"]

    def lin(p, i):
        return [f"{p}/linear_{i}/weight", f"{p}/linear_{i}/bias"]
    for l in range(n_enc):
        s_ = f"encoder/layer_{l}"
        names += norm(f"{s_}/self_attention/layer_norm") + lin(f"{s_}/self_attention", 0) + lin(f"{s_}/self_attention", 1)
        names += norm(f"{s_}/ffn/layer_norm") + lin(f"{s_}/ffn", 0) + lin(f"{s_}/ffn", 1)
    for l in range(n_dec):
        s_ = f"decoder/layer_{l}"
        names += norm(f"{s_}/self_attention/layer_norm") + lin(f"{s_}/self_attention", 0) + lin(f"{s_}/self_attention", 1)
        names += norm(f"{s_}/attention/layer_norm") + lin(f"{s_}/attention", 0) + lin(f"{s_}/attention", 1) + lin(f"{s_}/attention", 2)
        names += norm(f"{s_}/ffn/layer_norm") + lin(f"{s_}/ffn", 0) + lin(f"{s_}/ffn", 1)
    return names


def save_ct2_model_bin(weights: Dict[str, torch.Tensor], path: str, dtype=np.float16) -> None:
    """Inverse of load_ct2_model_bin (tests, and to hand a checkpoint to a CTranslate2 install for cross-checks)."""
    def npy(name):
        return weights[name].detach().cpu().numpy().astype(dtype)

    def zeros_like_bias(w):
        return np.zeros((w.shape[0],), dtype=dtype)

    v: Dict[str, np.ndarray] = {}
    for conv in ("conv1", "conv2"):
        v[f"encoder/{conv}/weight"], v[f"encoder/{conv}/bias"] = npy(f"model.encoder.{conv}.weight"), npy(f"model.encoder.{conv}.bias")
    v["encoder/position_encodings/encodings"] = npy("model.encoder.embed_positions.weight")
    v["encoder/layer_norm/gamma"], v["encoder/layer_norm/beta"] = npy("model.encoder.layer_norm.weight"), npy("model.encoder.layer_norm.bias")

    def norm(src, dst):
        v[f"{dst}/gamma"], v[f"{dst}/beta"] = npy(f"{src}.weight"), npy(f"{src}.bias")

    def ffn(h, s):
        norm(f"{h}.final_layer_norm", f"{s}/ffn/layer_norm")
        for i, fc in enumerate(("fc1", "fc2")):
            v[f"{s}/ffn/linear_{i}/weight"], v[f"{s}/ffn/linear_{i}/bias"] = npy(f"{h}.{fc}.weight"), npy(f"{h}.{fc}.bias")

    def self_attn(h, s):
        wq, wk, wv = npy(f"{h}.q_proj.weight"), npy(f"{h}.k_proj.weight"), npy(f"{h}.v_proj.weight")
        v[f"{s}/linear_0/weight"] = np.concatenate([wq, wk, wv], 0)
        v[f"{s}/linear_0/bias"] = np.concatenate([npy(f"{h}.q_proj.bias"), zeros_like_bias(wk), npy(f"{h}.v_proj.bias")], 0)
        v[f"{s}/linear_1/weight"], v[f"{s}/linear_1/bias"] = npy(f"{h}.out_proj.weight"), npy(f"{h}.out_proj.bias")

    i = 0
    while f"model.encoder.layers.{i}.fc1.weight" in weights:
        h, s = f"model.encoder.layers.{i}", f"encoder/layer_{i}"
        norm(f"{h}.self_attn_layer_norm", f"{s}/self_attention/layer_norm")
        self_attn(f"{h}.self_attn", f"{s}/self_attention")
        ffn(h, s)
        i += 1
    v["decoder/embeddings/weight"] = npy("model.decoder.embed_tokens.weight")
    v["decoder/position_encodings/encodings"] = npy("model.decoder.embed_positions.weight")
    v["decoder/layer_norm/gamma"], v["decoder/layer_norm/beta"] = npy("model.decoder.layer_norm.weight"), npy("model.decoder.layer_norm.bias")
    i = 0
    while f"model.decoder.layers.{i}.fc1.weight" in weights:
        h, s = f"model.decoder.layers.{i}", f"decoder/layer_{i}"
        norm(f"{h}.self_attn_layer_norm", f"{s}/self_attention/layer_norm")
        self_attn(f"{h}.self_attn", f"{s}/self_attention")
        norm(f"{h}.encoder_attn_layer_norm", f"{s}/attention/layer_norm")
        a = f"{h}.encoder_attn"
        wk, wv = npy(f"{a}.k_proj.weight"), npy(f"{a}.v_proj.weight")
        v[f"{s}/attention/linear_0/weight"], v[f"{s}/attention/linear_0/bias"] = npy(f"{a}.q_proj.weight"), npy(f"{a}.q_proj.bias")
        v[f"{s}/attention/linear_1/weight"] = np.concatenate([wk, wv], 0)
        v[f"{s}/attention/linear_1/bias"] = np.concatenate([zeros_like_bias(wk), npy(f"{a}.v_proj.bias")], 0)
        v[f"{s}/attention/linear_2/weight"], v[f"{s}/attention/linear_2/bias"] = npy(f"{a}.out_proj.weight"), npy(f"{a}.out_proj.bias")
        ffn(h, s)
        i += 1
    # the output projection is tied to the embedding: CT2 stores it once and lists the second name as an alias
    write_variables(path, v, aliases={"decoder/projection/weight": "decoder/embeddings/weight"})


def read_ct2_config(model_dir: str) -> Dict[str, object]:
    """``config.json`` next to model.bin: alignment heads and special-token id lists written by the CT2 converter
    (keys ``alignment_heads``, ``lang_ids``, ``suppress_ids``, ``suppress_ids_begin``); absent keys are omitted."""
    p = os.path.join(model_dir, "config.json")
    if not os.path.exists(p):
        return {}
    cfg = json.load(open(p))
    out: Dict[str, object] = {}
    if "alignment_heads" in cfg:
        out["alignment_heads"] = [(int(a), int(b)) for a, b in cfg["alignment_heads"]]
    for k in ("lang_ids", "suppress_ids", "suppress_ids_begin"):
        if k in cfg:
            out[k] = [int(x) for x in cfg[k]]
    return out
