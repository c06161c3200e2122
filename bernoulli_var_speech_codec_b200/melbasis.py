"""Slaney-normalised mel filterbank (host-side constant table for the log-mel kernel).

The reference obtains this matrix from ``librosa.filters.mel`` (un-vendored,
pinned ``librosa==0.8.1`` in reference third_party/BigVGAN/requirements.txt:3;
call site third_party/BigVGAN/meldataset.py:68 with sr=22050, n_fft=1024,
n_mels=80, fmin=0, fmax=8000, defaults htk=False, norm='slaney', float32).
librosa is not part of this image, so the published construction is restated
here: triangular filters on the Slaney auditory scale (linear below 1 kHz,
logarithmic above), each normalised to unit area.
"""
from __future__ import annotations

import numpy as np

_F_SP = 200.0 / 3.0
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f / _F_SP
    with np.errstate(divide="ignore", invalid="ignore"):
        logpart = _MIN_LOG_MEL + np.log(np.maximum(f, 1e-30) / _MIN_LOG_HZ) / _LOGSTEP
    return np.where(f >= _MIN_LOG_HZ, logpart, lin)


def mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    lin = _F_SP * m
    logpart = _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL))
    return np.where(m >= _MIN_LOG_MEL, logpart, lin)


def slaney_mel_basis(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """Returns float32 ``[n_mels, n_fft//2 + 1]``."""
    n_bins = 1 + n_fft // 2
    fft_hz = np.linspace(0.0, sr / 2.0, n_bins)
    edges = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    width = np.diff(edges)
    ramps = edges[:, None] - fft_hz[None, :]
    basis = np.zeros((n_mels, n_bins), dtype=np.float32)
    for m in range(n_mels):
        rising = -ramps[m] / width[m]
        falling = ramps[m + 2] / width[m + 1]
        basis[m] = np.maximum(0.0, np.minimum(rising, falling))
    area_norm = 2.0 / (edges[2:n_mels + 2] - edges[:n_mels])
    basis *= area_norm[:, None]
    return basis


def sparse_rows(basis: np.ndarray):
    """Per mel row: first non-zero FFT bin, tap count, and a dense tap table.

    Returns (start[int32 n_mels], count[int32 n_mels], taps[float32 n_mels, max_count]).
    The filters are contiguous triangles, so [start, start+count) covers every
    non-zero weight of the row.
    """
    n_mels = basis.shape[0]
    start = np.zeros(n_mels, dtype=np.int32)
    count = np.zeros(n_mels, dtype=np.int32)
    for m in range(n_mels):
        nz = np.nonzero(basis[m])[0]
        if nz.size:
            start[m] = nz[0]
            count[m] = nz[-1] - nz[0] + 1
    width = int(count.max()) if n_mels else 0
    taps = np.zeros((n_mels, max(width, 1)), dtype=np.float32)
    for m in range(n_mels):
        taps[m, :count[m]] = basis[m, start[m]:start[m] + count[m]]
    return start, count, taps
