"""The oracle against the committed golden vectors (CPU).

mel_reference.npz   <- reference's in-repo mel code (tests/golden/make_golden_mel.py)
model_hf.npz        <- HF transformers Whisper (tests/golden/make_golden_model.py)
"""
import os

import numpy as np
import pytest
import torch

from oracle import mel as omel
from oracle import model as om
from whisperlive_b200 import synth
from whisperlive_b200.config import dims_for
from whisperlive_b200.weights import random_init

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_mel_oracle_matches_reference_formula():
    g = np.load(os.path.join(GOLD, "mel_reference.npz"))
    keys = [k for k in g.files if not k.startswith("wav__")]
    assert len(keys) == 10
    for k in keys:
        name, n_mels = k.split("__")
        out = omel.log_mel(g["wav__" + name], int(n_mels))
        assert out.shape == g[k].shape and out.dtype == np.float32
        # fp32 FFT implementations differ in the last bits; log-mel units
        np.testing.assert_allclose(out, g[k], atol=5e-5, rtol=0)


def test_mel_filterbank_matches_hf():
    from transformers.audio_utils import mel_filter_bank
    for n in (80, 128):
        hf = mel_filter_bank(num_frequency_bins=201, num_mel_filters=n, min_frequency=0.0, max_frequency=8000.0,
                             sampling_rate=16000, norm="slaney", mel_scale="slaney").T
        np.testing.assert_allclose(omel.slaney_mel_filters(n), hf.astype(np.float32), atol=1e-7)


def test_mel_shapes_and_padding():
    for n in (1600, 16000, 17001, 480000):
        f = omel.log_mel(np.zeros(n, np.float32), 80)
        assert f.shape == (80, n // 160 + 1)
    f = omel.pad_or_trim(np.ones((80, 10), np.float32))
    assert f.shape == (80, 3000) and f[:, 10:].sum() == 0  # zero padding in feature space
    assert omel.pad_or_trim(np.ones((80, 3100), np.float32)).shape == (80, 3000)


@pytest.mark.parametrize("name", ["micro.en", "tiny"])
def test_network_oracle_matches_hf(name):
    g = np.load(os.path.join(GOLD, "model_hf.npz"))
    dims = dims_for(name)
    w = random_init(dims, seed=int(g[name + "__init_seed"][0]))
    wav = synth.speech_like(7.3, seed=int(g[name + "__wav_seed"][0]))
    feats = omel.pad_or_trim(omel.log_mel(wav, dims.n_mels)[:, :-1])
    with torch.no_grad():
        enc = om.encoder_forward(w, torch.from_numpy(feats)[None], dims.n_heads, dims.enc_layers)
        xkv = om.cross_kv(w, enc, dims.n_heads, dims.dec_layers)
        toks = torch.from_numpy(g[name + "__tokens"])
        logits = om.decoder_forward(w, toks, xkv, om.DecoderState(dims.dec_layers), dims.n_heads, dims.dec_layers)
        st = om.DecoderState(dims.dec_layers)
        inc = torch.cat([om.decoder_forward(w, toks[:, i:i + 1], xkv, st, dims.n_heads, dims.dec_layers)
                         for i in range(toks.shape[1])], 1)
    np.testing.assert_allclose(enc[0, ::25].numpy(), g[name + "__enc_sub"], atol=2e-4)
    np.testing.assert_allclose(logits[0, :, :512].numpy(), g[name + "__logits_head"], atol=5e-4)
    np.testing.assert_allclose(logits[0, :, -1700:].numpy(), g[name + "__logits_tail"], atol=5e-4)
    np.testing.assert_allclose(torch.logsumexp(logits[0], -1).numpy(), g[name + "__logits_lse"], atol=5e-4)
    # KV-cached incremental decoding == full teacher-forced pass
    np.testing.assert_allclose(inc.numpy(), logits.numpy(), atol=2e-4)
