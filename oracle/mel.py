"""Oracle K1: log-mel features, numpy restatement (TEST INFRASTRUCTURE ONLY).

Follows the feature extractor the reference calls at
  whisper_live/transcriber/transcriber_faster_whisper.py:655,:862 (FeatureExtractor(audio))
  whisper_live/batch_inference.py:258-259 (features + pad_or_trim)
whose arithmetic lives in faster-whisper==1.2.0 (requirements/server.txt:1),
``faster_whisper/feature_extractor.py`` -- NOT vendored in /root/reference.  The
published algorithm is restated here; the only in-repo statement of the same
formula is whisper_live/transcriber/tensorrt_utils.py:177-190 (torch.stft, hann,
drop last frame, filters @ |stft|^2, clamp 1e-10, log10, max-8, (x+4)/4), which
``tests/golden/make_golden_mel.py`` executes to pin this file.

Differences from the OpenAI/HF/TRT front end that this oracle must keep
(SURVEY.md A.1): the waveform is padded by 160 zero samples only (no 30 s audio
padding), the global max is over the real chunk, and callers zero-pad *features*
to 3000 frames (pad_or_trim).
"""
from __future__ import annotations

import numpy as np

SAMPLING_RATE = 16000
N_FFT = 400
HOP = 160
N_FREQ = N_FFT // 2 + 1  # 201
NB_MAX_FRAMES = 3000


def slaney_mel_filters(n_mels: int, sr: int = SAMPLING_RATE, n_fft: int = N_FFT) -> np.ndarray:
    """[n_mels, 201] float32 Slaney-scale, area-normalised triangular filterbank
    (== librosa.filters.mel(sr=16000, n_fft=400, n_mels=n_mels))."""
    bin_hz = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    # mel axis: linear below 1 kHz (200/3 Hz per mel), logarithmic above
    mel_pts = np.linspace(0.0, 45.245640471924965, n_mels + 2)
    hz_per_mel = 200.0 / 3.0
    knee_hz = 1000.0
    knee_mel = knee_hz / hz_per_mel
    log_step = np.log(6.4) / 27.0
    edges = hz_per_mel * mel_pts
    above = mel_pts >= knee_mel
    edges[above] = knee_hz * np.exp(log_step * (mel_pts[above] - knee_mel))

    width = np.diff(edges)
    dist = edges[:, None] - bin_hz[None, :]
    rising = -dist[:-2] / width[:-1, None]
    falling = dist[2:] / width[1:, None]
    tri = np.maximum(0.0, np.minimum(rising, falling))
    tri *= (2.0 / (edges[2:n_mels + 2] - edges[:n_mels]))[:, None]
    return tri.astype(np.float32)


def hann_window(n_fft: int = N_FFT) -> np.ndarray:
    """Periodic Hann window: np.hanning(n_fft + 1)[:-1] as float32."""
    return np.hanning(n_fft + 1)[:-1].astype(np.float32)


def stft_power(wave: np.ndarray, n_fft: int = N_FFT, hop: int = HOP) -> np.ndarray:
    """|STFT|^2 with center=True / reflect padding, one-sided. Returns [201, n_frames]
    float32, n_frames = 1 + len(wave)//hop (caller drops the last one)."""
    window = hann_window(n_fft)
    padded = np.pad(wave, (n_fft // 2, n_fft // 2), mode="reflect")
    n_frames = 1 + (padded.shape[0] - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(n_frames)[:, None]
    frames = padded[idx] * window[None, :]
    spec = np.fft.rfft(frames, n=n_fft, axis=-1).astype(np.complex64)  # [n_frames, 201]
    return (np.abs(spec) ** 2).T.astype(np.float32)


def log_mel(wave: np.ndarray, n_mels: int = 80, padding: int = 160) -> np.ndarray:
    """FeatureExtractor.__call__: float32 [n_mels, len(wave)//160 + 1].

    The final column is the extra frame callers drop (content_frames = shape[-1]-1,
    transcriber_faster_whisper.py:1057)."""
    wave = np.asarray(wave, dtype=np.float32)
    if padding:
        wave = np.pad(wave, (0, padding))
    power = stft_power(wave)[:, :-1]
    mel = slaney_mel_filters(n_mels) @ power
    log_spec = np.log10(np.clip(mel, 1e-10, None))
    log_spec = np.maximum(log_spec, log_spec.max() - 8.0)
    return ((log_spec + 4.0) / 4.0).astype(np.float32)


def pad_or_trim(features: np.ndarray, length: int = NB_MAX_FRAMES) -> np.ndarray:
    """faster_whisper.audio.pad_or_trim on the last axis (zero padding in feature space;
    call sites transcriber_faster_whisper.py:1127, batch_inference.py:259)."""
    n = features.shape[-1]
    if n > length:
        return features[..., :length]
    if n < length:
        pad = [(0, 0)] * features.ndim
        pad[-1] = (0, length - n)
        return np.pad(features, pad)
    return features


class OracleFeatureExtractor:
    """Same attribute surface as faster_whisper.feature_extractor.FeatureExtractor."""

    def __init__(self, feature_size=80, sampling_rate=16000, hop_length=160, chunk_length=30, n_fft=400):
        assert sampling_rate == 16000 and hop_length == 160 and n_fft == 400
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.chunk_length = chunk_length
        self.n_samples = chunk_length * sampling_rate
        self.nb_max_frames = self.n_samples // hop_length
        self.time_per_frame = hop_length / sampling_rate
        self.sampling_rate = sampling_rate
        self.feature_size = feature_size
        self.mel_filters = slaney_mel_filters(feature_size)

    def __call__(self, waveform, padding=160, chunk_length=None):
        if chunk_length is not None:
            self.n_samples = chunk_length * self.sampling_rate
            self.nb_max_frames = self.n_samples // self.hop_length
        return log_mel(waveform, self.feature_size, padding)

    def batch(self, waveforms, padding=160, chunk_length=None):
        return [self(w, padding, chunk_length) for w in waveforms]
