"""Generate tests/golden/model_hf.npz: HF transformers Whisper (random weights from
whisperlive_b200.weights.random_init) encoder output + decoder logits, sub-sampled.
Pins oracle/model.py's network arithmetic to modeling_whisper.py (transformers 5.5.0).

    python tests/golden/make_golden_model.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from transformers import WhisperConfig, WhisperForConditionalGeneration  # noqa: E402

from oracle import mel as omel  # noqa: E402
from whisperlive_b200 import synth  # noqa: E402
from whisperlive_b200.config import dims_for  # noqa: E402
from whisperlive_b200.weights import random_init  # noqa: E402

torch.manual_seed(0)
out = {}
for name, seed in (("micro.en", 3), ("tiny", 4)):
    dims = dims_for(name)
    cfg = WhisperConfig(
        vocab_size=dims.vocab, num_mel_bins=dims.n_mels, d_model=dims.d_model,
        encoder_layers=dims.enc_layers, decoder_layers=dims.dec_layers,
        encoder_attention_heads=dims.n_heads, decoder_attention_heads=dims.n_heads,
        encoder_ffn_dim=dims.d_ff, decoder_ffn_dim=dims.d_ff,
        max_source_positions=1500, max_target_positions=448, activation_function="gelu",
        dropout=0.0, attention_dropout=0.0, activation_dropout=0.0)
    cfg._attn_implementation = "eager"
    model = WhisperForConditionalGeneration(cfg).eval()
    w = random_init(dims, seed=seed)
    sd = dict(w)
    sd["proj_out.weight"] = w["model.decoder.embed_tokens.weight"]
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("k_proj.bias" in m or m == "proj_out.weight" for m in missing), missing
    wav = synth.speech_like(7.3, seed=100 + seed)
    feats = omel.pad_or_trim(omel.log_mel(wav, dims.n_mels)[:, :-1])
    tokens = torch.tensor([[50257, 50362, 400, 1234, 50256 - 7, 11]], dtype=torch.long)
    with torch.no_grad():
        enc = model.model.encoder(torch.from_numpy(feats)[None]).last_hidden_state
        logits = model(input_features=torch.from_numpy(feats)[None], decoder_input_ids=tokens).logits
    out[f"{name}__wav_seed"] = np.array([100 + seed])
    out[f"{name}__init_seed"] = np.array([seed])
    out[f"{name}__tokens"] = tokens.numpy()
    out[f"{name}__enc_sub"] = enc[0, ::25].numpy()              # [60, d]
    out[f"{name}__logits_head"] = logits[0, :, :512].numpy()    # [T, 512]
    out[f"{name}__logits_tail"] = logits[0, :, -1700:].numpy()  # specials + timestamps
    out[f"{name}__logits_lse"] = torch.logsumexp(logits[0], -1).numpy()
np.savez_compressed(os.path.join(HERE, "model_hf.npz"), **out)
print("wrote model_hf.npz", {k: v.shape for k, v in out.items()})
