"""Oracle K14: cross-attention word alignment post-processing (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED: the reference calls ``ctranslate2.models.Whisper.align``
(whisper_live/transcriber/transcriber_faster_whisper.py:1657-1663); CT2 is not
vendored/installed.  Restated from OpenAI whisper/timing.py (find_alignment,
median_filter, dtw_cpu) with CT2's ordering: attention probabilities are the
decoder layer's softmax over all 1500 encoder positions, THEN sliced to
num_frames // 2 (OpenAI slices before the softmax; SURVEY.md A.3 flags this as
the first thing to verify next to a real CT2).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def median_filter_time(x: np.ndarray, width: int) -> np.ndarray:
    """Median over a sliding window along the last axis, reflect padding."""
    pad = width // 2
    if pad == 0 or x.shape[-1] <= pad:
        return x
    xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(pad, pad)], mode="reflect")
    win = np.lib.stride_tricks.sliding_window_view(xp, width, axis=-1)
    return np.sort(win, axis=-1)[..., pad]


def dtw_path(cost_matrix: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Monotone alignment minimising the summed cost; moves (1,1),(1,0),(0,1), diagonal preferred
    only when strictly cheaper than both, then 'up', then 'left' (OpenAI dtw_cpu order)."""
    n, m = cost_matrix.shape
    acc = np.full((n + 1, m + 1), np.inf, dtype=np.float32)
    move = -np.ones((n + 1, m + 1), dtype=np.int8)
    acc[0, 0] = 0.0
    for j in range(1, m + 1):
        for i in range(1, n + 1):
            c0, c1, c2 = acc[i - 1, j - 1], acc[i - 1, j], acc[i, j - 1]
            if c0 < c1 and c0 < c2:
                c, t = c0, 0
            elif c1 < c0 and c1 < c2:
                c, t = c1, 1
            else:
                c, t = c2, 2
            acc[i, j] = cost_matrix[i - 1, j - 1] + c
            move[i, j] = t
    move[0, :] = 2
    move[:, 0] = 1
    i, j = n, m
    path = []
    while i > 0 or j > 0:
        path.append((i - 1, j - 1))
        t = move[i, j]
        if t == 0:
            i, j = i - 1, j - 1
        elif t == 1:
            i -= 1
        else:
            j -= 1
    path = np.array(path[::-1], dtype=np.int64)
    return path[:, 0], path[:, 1]


def alignment_from_attention(attn: np.ndarray, n_start: int, num_frames: int, median_width: int = 7) -> List[Tuple[int, int]]:
    """attn [n_align_heads, n_tok, 1500] softmax probabilities for the token sequence
    start_sequence + [no_timestamps] + text + [eot]; returns [(text_idx, time_idx)]."""
    w = attn[:, :, : num_frames // 2].astype(np.float32)
    mean = w.mean(axis=-2, keepdims=True)
    std = w.std(axis=-2, keepdims=True)
    w = (w - mean) / std
    w = median_filter_time(w, median_width)
    mat = w.mean(axis=0)[n_start:-1]
    ti, fi = dtw_path(-mat)
    return list(zip(ti.tolist(), fi.tolist()))
