"""Generate tests/golden/mel_*.npz by EXECUTING the reference's in-repo mel code.

Runs only in the build container (needs /root/reference).  The reference module
whisper_live/transcriber/tensorrt_utils.py imports audio I/O packages that are
not installed (kaldialign, soundfile, av); none of them is touched by
``log_mel_spectrogram`` when it is handed a tensor, so they are stubbed in
sys.modules.  Its ``mel_filters.npz`` asset is not shipped in the repository
(tensorrt_utils.py:107-127 documents it as librosa.filters.mel output) so the
file is produced here from HF transformers' independent Slaney filterbank.

Passing a tensor skips the 30 s audio padding (tensorrt_utils.py:160-170), and
padding=160 gives F.pad(audio,(0,160)) -- exactly the faster-whisper front end
the reference's live path uses.

    python tests/golden/make_golden_mel.py
"""
import importlib.machinery
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from whisperlive_b200 import synth  # noqa: E402

from transformers.audio_utils import mel_filter_bank  # noqa: E402  (before the stubs: it probes soundfile)

for name in ("kaldialign", "soundfile", "av"):
    if name not in sys.modules:
        stub = types.ModuleType(name)
        stub.__spec__ = importlib.machinery.ModuleSpec(name, None)
        sys.modules[name] = stub
sys.path.insert(0, "/root/reference")
from whisper_live.transcriber import tensorrt_utils as ref  # noqa: E402

tmp = tempfile.mkdtemp()
banks = {}
for n in (80, 128):
    banks[f"mel_{n}"] = mel_filter_bank(
        num_frequency_bins=201, num_mel_filters=n, min_frequency=0.0, max_frequency=8000.0,
        sampling_rate=16000, norm="slaney", mel_scale="slaney").T.astype(np.float32)
np.savez_compressed(os.path.join(tmp, "mel_filters.npz"), **banks)

cases = {
    "speech_1p0s": synth.speech_like(1.0, seed=11),
    "speech_2p37s": synth.speech_like(2.37, seed=12),
    "noise_1p5s": synth.white_noise(1.5, seed=13),
    "silence_1p0s": synth.silence(1.0),
    "speech_odd_17001": synth.speech_like(17001 / 16000, seed=14),
}
out = {}
for n_mels in (80, 128):
    for name, wav in cases.items():
        feats = ref.log_mel_spectrogram(torch.from_numpy(wav), n_mels, padding=160, mel_filters_dir=tmp)
        # the reference function already drops the extra STFT frame (stft[..., :-1])
        out[f"{name}__{n_mels}"] = feats.numpy().astype(np.float32)
for name, wav in cases.items():
    out[f"wav__{name}"] = wav
np.savez_compressed(os.path.join(HERE, "mel_reference.npz"), **out)
print("wrote", os.path.join(HERE, "mel_reference.npz"), {k: v.shape for k, v in out.items() if "__" in k and not k.startswith("wav")})
