"""Golden vectors for the mel front-end, from the UNMODIFIED reference mel_processing.py run in the build container.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_mel.py     ->  tests/golden/mel_frontend.npz

Two things the reference module needs are absent here, so the script supplies them before importing it:

  * `librosa` (requirements.txt pins 0.8.0; not installed, no network).  mel_processing.py only uses
    `librosa.filters.mel` on this path (:14, :76, :95).  A stub package is registered whose `filters.mel` is
    torchaudio.functional.melscale_fbanks(norm="slaney", mel_scale="slaney") -- an independent implementation
    of the same published filterbank -- so the reference functions run end to end.  The basis itself is stored
    together with transformers.audio_utils.mel_filter_bank's version of it; oracle/mel_frontend.slaney_mel_basis
    is checked against both (tests/test_mel_frontend.py).
  * `torch.stft` without `return_complex` (the reference targets torch 1.6, mel_processing.py:63-64): wrapped so that
    the old real-valued [..., 2] layout is returned by today's torch.  Nothing else of the reference is touched.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SVK_REFERENCE", "/root/reference")

N_FFT, HOP, WIN, SR, N_MELS, FMIN, FMAX = 1024, 256, 1024, 22050, 80, 0.0, None  # configs/iitp_base.json "data"


def torchaudio_basis(sr, n_fft, n_mels, fmin, fmax):
    import torchaudio
    fb = torchaudio.functional.melscale_fbanks(n_fft // 2 + 1, float(fmin), float(sr / 2 if fmax is None else fmax), n_mels,
                                               sr, norm="slaney", mel_scale="slaney")
    return fb.T.contiguous().numpy().astype(np.float32)


def transformers_basis(sr, n_fft, n_mels, fmin, fmax):
    from transformers.audio_utils import mel_filter_bank
    fb = mel_filter_bank(n_fft // 2 + 1, n_mels, float(fmin), float(sr / 2 if fmax is None else fmax), sr, norm="slaney",
                         mel_scale="slaney")
    return np.ascontiguousarray(fb.T).astype(np.float32)


def install_stubs():
    lib = types.ModuleType("librosa")
    util = types.ModuleType("librosa.util")
    filters = types.ModuleType("librosa.filters")
    util.normalize = util.pad_center = util.tiny = None  # imported by name, never called on this path
    filters.mel = lambda sr, n_fft, n_mels, fmin, fmax: torchaudio_basis(sr, n_fft, n_mels, fmin, fmax)
    lib.util, lib.filters = util, filters
    sys.modules.update({"librosa": lib, "librosa.util": util, "librosa.filters": filters})
    real_stft = torch.stft

    def stft_compat(*a, **k):
        if "return_complex" in k:
            return real_stft(*a, **k)
        return torch.view_as_real(real_stft(*a, return_complex=True, **k))

    torch.stft = stft_compat


def signal(seed, B, n):
    """Speech-like test signal in [-1, 1]: a few chirping harmonics + noise, with a silent stretch."""
    rng = np.random.Generator(np.random.Philox(key=[seed, 77]))
    t = np.arange(n) / SR
    y = np.zeros((B, n))
    for b in range(B):
        f0 = 110.0 * (b + 1) + 30.0 * np.sin(2 * np.pi * 1.5 * t)
        ph = 2 * np.pi * np.cumsum(f0) / SR
        for h in range(1, 12):
            y[b] += np.sin(h * ph + rng.uniform(0, 6.28)) / h
        y[b] = 0.25 * y[b] + 0.02 * rng.standard_normal(n)
        y[b, n // 3: n // 3 + 700] = 0.0
    return np.clip(y, -1, 1).astype(np.float32)


def main():
    bases = {"basis_torchaudio": torchaudio_basis(SR, N_FFT, N_MELS, FMIN, FMAX),
             "basis_transformers": transformers_basis(SR, N_FFT, N_MELS, FMIN, FMAX)}  # before the librosa stub exists
    install_stubs()
    sys.path.insert(0, REF)
    import mel_processing as ref  # noqa: E402  (reference, imported in place)

    out = {}
    cases = {"a": (2, 8192), "b": (1, 5000), "c": (3, 1024)}  # 5000: not a multiple of hop; 1024: T = 4
    for tag, (B, n) in cases.items():
        y = signal(11 + len(tag) + B, B, n)
        out[f"{tag}_y"] = y
        for dt, name in ((torch.float32, "f32"), (torch.float64, "f64")):
            ref.mel_basis.clear(), ref.hann_window.clear()
            yt = torch.from_numpy(y).to(dt)
            spec = ref.spectrogram_torch(yt, N_FFT, SR, HOP, WIN, center=False)
            mel = ref.spec_to_mel_torch(spec, N_FFT, N_MELS, SR, FMIN, FMAX)
            mel2 = ref.mel_spectrogram_torch(yt, N_FFT, N_MELS, SR, HOP, WIN, FMIN, FMAX, center=False)
            assert torch.equal(mel, mel2)
            out[f"{tag}_spec_{name}"] = spec.numpy()
            out[f"{tag}_mel_{name}"] = mel.numpy()
        print(tag, y.shape, "->", out[f"{tag}_spec_f32"].shape, out[f"{tag}_mel_f32"].shape,
              "fp32 vs fp64: spec", float(np.abs(out[f"{tag}_spec_f32"] - out[f"{tag}_spec_f64"]).max()),
              "mel", float(np.abs(out[f"{tag}_mel_f32"] - out[f"{tag}_mel_f64"]).max()))
    out.update(bases)
    print("basis: torchaudio vs transformers max-abs", float(np.abs(out["basis_torchaudio"] - out["basis_transformers"]).max()))
    # the drop-in boundary: names, parameter order and defaults of the functions the callers use
    import inspect
    import json
    sigs = {n: str(inspect.signature(getattr(ref, n))) for n in
            ("spectrogram_torch", "spec_to_mel_torch", "mel_spectrogram_torch", "dynamic_range_compression_torch",
             "dynamic_range_decompression_torch", "spectral_normalize_torch", "spectral_de_normalize_torch")}
    sigs["MAX_WAV_VALUE"] = ref.MAX_WAV_VALUE
    json.dump(sigs, open(os.path.join(HERE, "mel_processing_signatures.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "mel_frontend.npz"), **out)
    print("wrote", os.path.join(HERE, "mel_frontend.npz"), os.path.getsize(os.path.join(HERE, "mel_frontend.npz")), "bytes")


if __name__ == "__main__":
    main()
