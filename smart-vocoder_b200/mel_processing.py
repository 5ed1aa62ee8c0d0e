"""Drop-in for the reference's ``mel_processing`` on the inference path (B200, libsvk).

Mirrors (reference file:line) the three functions the callers of ``SynthesizerTrn.infer`` use to make its
input (inference.ipynb:100-111, train.py:266-272), with the same names, argument order and module-level caches:

  * ``spectrogram_torch(y, n_fft, sampling_rate, hop_size, win_size, center=False)``          mel_processing.py:51-69
  * ``spec_to_mel_torch(spec, n_fft, num_mels, sampling_rate, fmin, fmax)``                    mel_processing.py:72-81
  * ``mel_spectrogram_torch(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False)``  :84-112
  * ``spectral_normalize_torch`` / ``dynamic_range_compression_torch`` and inverses            mel_processing.py:19-47

The STFT, magnitude, mel projection and log run in ONE hand-written CUDA kernel behind the C ABI
(csrc/mel_frontend.cu, include/svk.h ``svk_*spectrogram`` / ``svk_spec_to_mel``); torch only owns the buffers and
the stream.  There is no CPU compute path: a CPU input tensor (what inference.ipynb cells 100-111 and
data_utils.py:64 pass) is copied to the current CUDA device, processed there and returned on the CPU, like the
reference returns its result on the input's device.  ``librosa`` is not needed -- the mel basis
is built by ``svk_mel_basis`` (librosa.filters.mel's published algorithm).
"""
from __future__ import annotations

import ctypes

import torch

import svk_runtime as rt

MAX_WAV_VALUE = 32768.0

# reference-style module caches (mel_processing.py:47-48), keyed the same way; values are front-end handles
mel_basis = {}
hann_window = {}
_frontends = {}


class _Frontend:
    def __init__(self, n_fft, hop_size, win_size, sampling_rate, num_mels, fmin, fmax, device_index):
        self._h = ctypes.c_void_p()
        rt.check(rt.lib().svk_frontend_create(int(n_fft), int(hop_size), int(win_size), int(sampling_rate), int(num_mels),
                                              float(fmin), 0.0 if fmax is None else float(fmax), int(device_index),
                                              ctypes.byref(self._h)))

    @property
    def ptr(self):
        return self._h

    def frames(self, n):
        return int(rt.lib().svk_frontend_frames(self._h, int(n)))

    def __del__(self):  # pragma: no cover
        try:
            if self._h:
                rt.lib().svk_frontend_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass


def _frontend(t: torch.Tensor, n_fft, hop_size, win_size, sampling_rate, num_mels=80, fmin=0.0, fmax=None) -> _Frontend:
    if not t.is_cuda:  # callers go through _to_device() first
        raise rt.SvkError(rt.SVK_ERR_CUDA, "mel_processing: input must be a CUDA tensor (libsvk has no CPU fallback)")
    if t.dtype != torch.float32:
        raise rt.SvkError(rt.SVK_ERR_INVALID, "mel_processing: fp32 tensors only")
    key = (int(n_fft), int(hop_size), int(win_size), int(sampling_rate), int(num_mels), float(fmin),
           None if fmax is None else float(fmax), t.device.index if t.device.index is not None else torch.cuda.current_device())
    fe = _frontends.get(key)
    if fe is None:
        fe = _frontends[key] = _Frontend(*key)
        dtype_device = str(t.dtype) + '_' + str(t.device)
        mel_basis[str(fmax) + '_' + dtype_device] = fe      # same cache keys as the reference
        hann_window[str(win_size) + '_' + dtype_device] = fe
    return fe


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _to_device(t: torch.Tensor):
    """The reference's callers feed CPU tensors (inference.ipynb cells 100-111 call spectrogram_torch /
    spec_to_mel_torch before `.cuda()`; data_utils.py:64 too).  There is no CPU compute path here: a CPU input
    is copied to the current CUDA device, the kernel runs there, and the result goes back to the input's
    device, so the caller sees the reference's behaviour (same device in, same device out)."""
    if t.is_cuda:
        return t, (lambda r: r)
    if not torch.cuda.is_available():
        raise rt.SvkError(rt.SVK_ERR_CUDA, "mel_processing: no CUDA device (libsvk has no CPU fallback)")
    src = t.device
    return t.to("cuda", non_blocking=False), (lambda r: r.to(src))


def dynamic_range_compression_torch(x, C=1, clip_val=1e-5):
    return torch.log(torch.clamp(x, min=clip_val) * C)


def dynamic_range_decompression_torch(x, C=1):
    return torch.exp(x) / C


def spectral_normalize_torch(magnitudes):
    return dynamic_range_compression_torch(magnitudes)


def spectral_de_normalize_torch(magnitudes):
    return dynamic_range_decompression_torch(magnitudes)


def _check_range(y):
    # the reference prints (not raises) when the waveform leaves [-1, 1] (mel_processing.py:52-55)
    if torch.min(y) < -1.:
        print('min value is ', torch.min(y))
    if torch.max(y) > 1.:
        print('max value is ', torch.max(y))


def spectrogram_torch(y, n_fft, sampling_rate, hop_size, win_size, center=False):
    if center:
        raise NotImplementedError("center=True is never used by the reference's callers; only center=False is built")
    _check_range(y)
    y, back = _to_device(y)
    fe = _frontend(y, n_fft, hop_size, win_size, sampling_rate)
    y = y.contiguous()
    B, n = y.shape
    spec = torch.empty(B, n_fft // 2 + 1, fe.frames(n), device=y.device, dtype=torch.float32)
    with torch.cuda.device(y.device):
        rt.check(rt.lib().svk_spectrogram(fe.ptr, y.data_ptr(), B, n, spec.data_ptr(), _stream()))
    return back(spec)


def spec_to_mel_torch(spec, n_fft, num_mels, sampling_rate, fmin, fmax):
    # hop/win do not enter the projection; any valid pair selects the same mel basis
    spec, back = _to_device(spec)
    fe = _frontend(spec, n_fft, n_fft // 4, n_fft, sampling_rate, num_mels, fmin, fmax)
    spec = spec.contiguous()
    B, nb, T = spec.shape
    if nb != n_fft // 2 + 1:
        raise rt.SvkError(rt.SVK_ERR_INVALID, f"spec has {nb} bins, n_fft={n_fft} needs {n_fft // 2 + 1}")
    mel = torch.empty(B, num_mels, T, device=spec.device, dtype=torch.float32)
    with torch.cuda.device(spec.device):
        rt.check(rt.lib().svk_spec_to_mel(fe.ptr, spec.data_ptr(), B, T, mel.data_ptr(), _stream()))
    return back(mel)


def mel_spectrogram_torch(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False):
    if center:
        raise NotImplementedError("center=True is never used by the reference's callers; only center=False is built")
    _check_range(y)
    y, back = _to_device(y)
    fe = _frontend(y, n_fft, hop_size, win_size, sampling_rate, num_mels, fmin, fmax)
    y = y.contiguous()
    B, n = y.shape
    mel = torch.empty(B, num_mels, fe.frames(n), device=y.device, dtype=torch.float32)
    with torch.cuda.device(y.device):
        rt.check(rt.lib().svk_mel_spectrogram(fe.ptr, y.data_ptr(), B, n, mel.data_ptr(), None, _stream()))
    return back(mel)
