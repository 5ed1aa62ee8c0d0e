"""CPU restatement of the reference's mel front-end (TEST INFRASTRUCTURE: only tests/, smoke() and
bench.py's cpu_baseline leg may import this; the product path is smart-vocoder_b200/csrc/mel_frontend.cu).

Follows /root/reference/mel_processing.py:
  * spectrogram_torch        :51-69   reflect pad (n_fft-hop)/2, Hann(win) STFT hop `hop`, center=False,
                                       onesided, sqrt(re^2 + im^2 + 1e-6)
  * spec_to_mel_torch        :72-81   mel_basis @ spec, log(clamp(., 1e-5))   (:16-22, :39-41)
  * mel_spectrogram_torch    :84-112  the two fused

The mel basis comes from librosa.filters.mel (third-party, absent here and not vendored by the reference;
requirements.txt pins librosa==0.8.0).  Its published algorithm (Slaney auditory-toolbox mel scale,
htk=False, area ("slaney") normalisation) is restated in `slaney_mel_basis`.  PINNING: the STFT/magnitude/log
part is pinned against the reference's own functions run in the build container
(tests/golden/make_golden_mel.py -> tests/golden/mel_frontend.npz); the basis is pinned against two
independent implementations of the same published filterbank (torchaudio.functional.melscale_fbanks and
transformers.audio_utils.mel_filter_bank, both norm="slaney", mel_scale="slaney"), stored in the same fixture.
"""
from __future__ import annotations

import numpy as np


def hann_window(win: int, dtype=np.float64) -> np.ndarray:
    """The window the reference uses: `torch.hann_window(win_size).to(dtype=y.dtype)` (mel_processing.py:59-60) --
    i.e. ALWAYS the fp32 periodic Hann window, whatever the signal dtype.  torch builds it in fp32 as
    arange(win) * float(2 pi / win) -> cos -> * -0.5 + 0.5; restated with the same roundings (torch's vectorised
    cosf may differ by 1 ulp = 6e-8 on some taps)."""
    a = (np.arange(win, dtype=np.float32) * np.float32(2.0 * np.pi / win)).astype(np.float32)
    w = (np.cos(a.astype(np.float64)).astype(np.float32) * np.float32(-0.5) + np.float32(0.5)).astype(np.float32)
    return w.astype(dtype)


def _hz_to_mel(f):
    f = np.asarray(f, np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore", invalid="ignore"):
        log_part = min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep
    return np.where(f >= min_log_hz, log_part, mels)


def _mel_to_hz(m):
    m = np.asarray(m, np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def slaney_mel_basis(sr: int, n_fft: int, n_mels: int, fmin: float = 0.0, fmax=None) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults htk=False, norm='slaney':
    triangular filters between n_mels+2 points equally spaced on the Slaney mel scale, each scaled by
    2 / (f_hi - f_lo).  Returns float32 [n_mels, n_fft//2 + 1] (librosa's default dtype)."""
    if fmax is None:
        fmax = sr / 2.0
    n_bins = n_fft // 2 + 1
    fftfreqs = np.linspace(0.0, sr / 2.0, n_bins)
    mel_pts = np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2)
    mel_f = _mel_to_hz(mel_pts)
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    weights = np.zeros((n_mels, n_bins), np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(np.float32)


def frame_count(n_samples: int, n_fft: int, hop: int) -> int:
    """Frames of torch.stft(center=False) over the reflect-padded signal (mel_processing.py:62-66)."""
    pad = (n_fft - hop) // 2
    padded = n_samples + 2 * pad
    return 0 if padded < n_fft else 1 + (padded - n_fft) // hop


def spectrogram(y: np.ndarray, n_fft: int, hop: int, win: int, dtype=np.float64) -> np.ndarray:
    """mel_processing.py:51-69.  y [B, n] -> [B, n_fft//2+1, T].  `dtype` is the arithmetic type."""
    y = np.asarray(y, dtype)
    pad = (n_fft - hop) // 2
    if y.shape[1] <= pad:
        raise ValueError("reflect padding needs n_samples > (n_fft - hop) / 2")
    yp = np.pad(y, ((0, 0), (pad, pad)), mode="reflect")
    T = frame_count(y.shape[1], n_fft, hop)
    w = np.zeros(n_fft, dtype)
    off = (n_fft - win) // 2  # torch.stft centres a short window inside n_fft
    w[off:off + win] = hann_window(win, dtype)
    idx = np.arange(n_fft)[None, :] + hop * np.arange(T)[:, None]
    frames = yp[:, idx] * w  # [B, T, n_fft]
    spec = np.fft.rfft(frames.astype(np.float64), axis=-1)  # numpy computes in double either way
    re, im = spec.real.astype(dtype), spec.imag.astype(dtype)
    mag = np.sqrt(re * re + im * im + dtype(1e-6))
    return np.ascontiguousarray(np.transpose(mag, (0, 2, 1))).astype(dtype)


def spec_to_mel(spec: np.ndarray, basis: np.ndarray, dtype=np.float64) -> np.ndarray:
    """mel_processing.py:72-81: log(clamp(basis @ spec, 1e-5))."""
    m = np.einsum("mf,bft->bmt", np.asarray(basis, dtype), np.asarray(spec, dtype))
    return np.log(np.maximum(m, dtype(1e-5))).astype(dtype)


def mel_spectrogram(y, n_fft, n_mels, sr, hop, win, fmin, fmax, dtype=np.float64) -> np.ndarray:
    """mel_processing.py:84-112."""
    return spec_to_mel(spectrogram(y, n_fft, hop, win, dtype), slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax), dtype)
