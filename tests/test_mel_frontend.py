"""Mel front-end: oracle pinned against the reference's mel_processing.py outputs
(tests/golden/make_golden_mel.py), host-side tables of the C ABI, and -- on the GPU -- the CUDA kernel through the
reference-named shim against the oracle and the reference goldens.

Tolerances (floating point, stated here):
  * reference fp32 vs reference fp64 on the golden signals: spec 7.3e-6 max-abs (|spec| <= 66), log-mel 2.1e-6.
  * oracle fp64 vs reference fp64: <= 5e-7 (the reference's fp32 Hann window restated to 1 ulp).
  * CUDA fp32 kernel vs fp64 oracle / reference: spec <= 1e-4 max-abs, log-mel <= 1e-4 max-abs.
"""
import numpy as np
import pytest

from conftest import load_golden

CFG = dict(n_fft=1024, hop=256, win=1024, sr=22050, n_mels=80, fmin=0.0, fmax=None)  # configs/iitp_base.json "data"
SPEC_TOL = 1e-4
MEL_TOL = 1e-4


@pytest.fixture(scope="module")
def golden():
    return load_golden("mel_frontend")


# ------------------------------------------------------------------------------------------ CPU: oracle + host tables
def test_oracle_basis_matches_two_independent_filterbanks(golden):
    from oracle import mel_frontend as M
    b = M.slaney_mel_basis(CFG["sr"], CFG["n_fft"], CFG["n_mels"], CFG["fmin"], CFG["fmax"])
    assert b.shape == (80, 513) and b.dtype == np.float32
    assert np.abs(b - golden["basis_transformers"]).max() <= 1e-9
    assert np.abs(b - golden["basis_torchaudio"]).max() <= 1e-7   # torchaudio builds it in fp32
    # slaney normalisation: every filter has (approximately) unit area in Hz / bin spacing terms
    assert (b >= 0).all() and (b.sum(axis=1) > 0).all()


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_oracle_against_reference_goldens(golden, tag):
    from oracle import mel_frontend as M
    y = golden[f"{tag}_y"]
    spec = M.spectrogram(y, CFG["n_fft"], CFG["hop"], CFG["win"])
    assert spec.shape == golden[f"{tag}_spec_f64"].shape
    assert spec.shape[2] == M.frame_count(y.shape[1], CFG["n_fft"], CFG["hop"])
    assert np.abs(spec - golden[f"{tag}_spec_f64"]).max() <= 5e-7
    mel = M.spec_to_mel(spec, golden["basis_torchaudio"])  # the basis the reference run used (librosa stub)
    assert np.abs(mel - golden[f"{tag}_mel_f64"]).max() <= 1e-6
    # and the reference's own fp32 run sits at its rounding floor around the oracle
    assert np.abs(spec - golden[f"{tag}_spec_f32"]).max() <= 2e-5
    assert np.abs(mel - golden[f"{tag}_mel_f32"]).max() <= 1e-5


def test_abi_host_tables_match_oracle(golden):
    import svk_runtime as rt
    from oracle import mel_frontend as M
    L = rt.lib()
    b = np.zeros((80, 513), np.float32)
    rt.check(L.svk_mel_basis(CFG["sr"], CFG["n_fft"], CFG["n_mels"], 0.0, 0.0, b.ctypes.data))
    assert np.abs(b - M.slaney_mel_basis(CFG["sr"], CFG["n_fft"], CFG["n_mels"], 0.0, None)).max() <= 1e-9
    b2 = np.zeros((40, 257), np.float32)
    rt.check(L.svk_mel_basis(16000, 512, 40, 50.0, 7600.0, b2.ctypes.data))
    assert np.abs(b2 - M.slaney_mel_basis(16000, 512, 40, 50.0, 7600.0)).max() <= 1e-9
    w = np.zeros(1024, np.float32)
    rt.check(L.svk_hann_window(1024, 1024, w.ctypes.data))
    assert np.array_equal(w, M.hann_window(1024, np.float32))
    import torch
    assert np.abs(w - torch.hann_window(1024).numpy()).max() <= 6e-8  # 1 ulp of torch's vectorised cosf
    w2 = np.zeros(1024, np.float32)
    rt.check(L.svk_hann_window(800, 1024, w2.ctypes.data))  # short window centred like torch.stft does
    assert (w2[:112] == 0).all() and (w2[912:] == 0).all() and np.array_equal(w2[112:912], M.hann_window(800, np.float32))
    with pytest.raises(rt.SvkError):
        rt.check(L.svk_hann_window(2048, 1024, w.ctypes.data))
    with pytest.raises(rt.SvkError):
        rt.check(L.svk_mel_basis(22050, 1024, 80, 9000.0, 8000.0, b.ctypes.data))


def test_frontend_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import ctypes
    import mel_processing as mp
    import svk_runtime as rt
    h = ctypes.c_void_p()
    assert rt.lib().svk_frontend_create(1024, 256, 1024, 22050, 80, 0.0, 0.0, 0, ctypes.byref(h)) == rt.SVK_ERR_CUDA
    with pytest.raises(rt.SvkError):
        mp.mel_spectrogram_torch(torch.zeros(1, 4096), 1024, 80, 22050, 256, 1024, 0.0, None)


# ------------------------------------------------------------------------------------------------ GPU: the kernel
def _gpu_mods():
    import torch
    import mel_processing as mp
    return torch, mp


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_gpu_frontend_against_reference_goldens(golden, tag):
    torch, mp = _gpu_mods()
    from oracle import mel_frontend as M
    y = torch.from_numpy(golden[f"{tag}_y"]).cuda()
    spec = mp.spectrogram_torch(y, CFG["n_fft"], CFG["sr"], CFG["hop"], CFG["win"], center=False)
    mel = mp.spec_to_mel_torch(spec, CFG["n_fft"], CFG["n_mels"], CFG["sr"], CFG["fmin"], CFG["fmax"])
    mel_fused = mp.mel_spectrogram_torch(y, CFG["n_fft"], CFG["n_mels"], CFG["sr"], CFG["hop"], CFG["win"], CFG["fmin"],
                                         CFG["fmax"], center=False)
    torch.cuda.synchronize()
    assert tuple(spec.shape) == golden[f"{tag}_spec_f64"].shape and tuple(mel.shape) == golden[f"{tag}_mel_f64"].shape
    assert torch.equal(mel, mel_fused), "fused and two-step paths must agree bit for bit (same arithmetic)"
    e_spec = np.abs(spec.cpu().numpy() - golden[f"{tag}_spec_f64"]).max()
    e_mel = np.abs(mel.cpu().numpy() - golden[f"{tag}_mel_f64"]).max()
    o_spec = M.spectrogram(golden[f"{tag}_y"], CFG["n_fft"], CFG["hop"], CFG["win"])
    o_mel = M.spec_to_mel(o_spec, M.slaney_mel_basis(CFG["sr"], CFG["n_fft"], CFG["n_mels"], CFG["fmin"], CFG["fmax"]))
    e_ospec = np.abs(spec.cpu().numpy() - o_spec).max()
    e_omel = np.abs(mel.cpu().numpy() - o_mel).max()
    print(f"mel front-end {tag}: vs reference fp64 spec {e_spec:.2e} mel {e_mel:.2e}; vs oracle spec {e_ospec:.2e} mel {e_omel:.2e}")
    assert e_spec <= SPEC_TOL and e_ospec <= SPEC_TOL
    assert e_mel <= MEL_TOL and e_omel <= MEL_TOL


@pytest.mark.gpu
@pytest.mark.parametrize("B,n,n_fft,hop,win,n_mels,sr,fmin,fmax", [
    (1, 385, 1024, 256, 1024, 80, 22050, 0.0, None),        # shortest legal signal: one frame, both edges reflected
    (3, 4111, 1024, 256, 1024, 80, 22050, 0.0, None),       # odd frame count (pairing tail) and ragged tile
    (2, 3000, 512, 128, 400, 40, 16000, 50.0, 7600.0),      # short window centred in n_fft, fmin/fmax set
    (1, 9000, 2048, 512, 2048, 128, 44100, 0.0, None),      # larger transform
    (5, 16 * 256 * 3, 1024, 256, 1024, 80, 22050, 0.0, 8000.0),  # exactly 3 tiles per utterance
])
def test_gpu_frontend_against_oracle_shapes(B, n, n_fft, hop, win, n_mels, sr, fmin, fmax):
    torch, mp = _gpu_mods()
    from oracle import mel_frontend as M
    rng = np.random.Generator(np.random.Philox(key=[n, n_fft]))
    y = np.clip(0.3 * rng.standard_normal((B, n)), -1, 1).astype(np.float32)
    yt = torch.from_numpy(y).cuda()
    spec = mp.spectrogram_torch(yt, n_fft, sr, hop, win)
    mel = mp.mel_spectrogram_torch(yt, n_fft, n_mels, sr, hop, win, fmin, fmax)
    torch.cuda.synchronize()
    o_spec = M.spectrogram(y, n_fft, hop, win)
    o_mel = M.spec_to_mel(o_spec, M.slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax))
    assert tuple(spec.shape) == o_spec.shape and tuple(mel.shape) == o_mel.shape
    assert np.abs(spec.cpu().numpy() - o_spec).max() <= SPEC_TOL
    assert np.abs(mel.cpu().numpy() - o_mel).max() <= MEL_TOL


@pytest.mark.gpu
def test_gpu_frontend_properties_at_full_size():
    """BASELINE configs[2] shape (16 utterances x 1024 frames): linearity in the signal before the magnitude is not
    observable, so use shift-invariance (a hop-aligned shift moves frames) and silence -> log floor."""
    torch, mp = _gpu_mods()
    B, T = 16, 1024
    n = T * 256
    g = torch.Generator(device="cuda").manual_seed(3)
    y = (0.2 * torch.randn(B, n + 256, device="cuda", generator=g)).clamp_(-1, 1)
    a = mp.mel_spectrogram_torch(y[:, :n].contiguous(), 1024, 80, 22050, 256, 1024, 0.0, None)
    b = mp.mel_spectrogram_torch(y[:, 256:].contiguous(), 1024, 80, 22050, 256, 1024, 0.0, None)
    torch.cuda.synchronize()
    assert tuple(a.shape) == (B, 80, T)
    # interior frames (away from the reflected edges: 2 frames per side) coincide after a one-hop shift; not bit for
    # bit, because a frame changes role (real <-> imaginary part) in the two-frames-per-FFT pairing
    assert float((a[:, :, 3:-2] - b[:, :, 2:-3]).abs().max()) <= 2e-5
    z = mp.mel_spectrogram_torch(torch.zeros(2, 8192, device="cuda"), 1024, 80, 22050, 256, 1024, 0.0, None)
    # silence: spec = sqrt(1e-6) = 1e-3 in every bin, mel = log(max(1e-3 * sum(basis_m), 1e-5))
    import svk_runtime as rt
    basis = np.zeros((80, 513), np.float32)
    rt.check(rt.lib().svk_mel_basis(22050, 1024, 80, 0.0, 0.0, basis.ctypes.data))
    want = np.log(np.maximum(1e-3 * basis.astype(np.float64).sum(axis=1), 1e-5))
    assert np.abs(z.cpu().numpy() - want[None, :, None]).max() <= 1e-5


@pytest.mark.gpu
def test_gpu_frontend_errors():
    torch, mp = _gpu_mods()
    import svk_runtime as rt
    with pytest.raises(rt.SvkError):
        mp.spectrogram_torch(torch.zeros(1, 384, device="cuda"), 1024, 22050, 256, 1024)   # n <= pad: reflect impossible
    with pytest.raises(rt.SvkError):
        mp.spectrogram_torch(torch.zeros(1, 4096, device="cuda"), 1000, 22050, 250, 1000)   # n_fft not a power of two
    with pytest.raises(NotImplementedError):
        mp.spectrogram_torch(torch.zeros(1, 4096, device="cuda"), 1024, 22050, 256, 1024, center=True)


@pytest.mark.gpu
def test_gpu_wave_to_wave_through_frontend(base_cfg, base_sd):
    """inference.ipynb:100-118 end to end on the device: wav -> mel_spectrogram_torch -> SynthesizerTrn.infer."""
    torch, mp = _gpu_mods()
    from models import SynthesizerTrn
    d = base_cfg["data"]
    net = SynthesizerTrn(d["filter_length"] // 2 + 1, base_cfg["train"]["segment_size"] // d["hop_length"],
                         **base_cfg["model"])
    net.load_state_dict({k: torch.from_numpy(v) for k, v in base_sd.items()})
    net = net.cuda().eval()
    n = 40 * d["hop_length"]
    t = torch.arange(n, device="cuda") / d["sampling_rate"]
    y = (0.4 * torch.sin(2 * np.pi * 220.0 * t))[None, :].contiguous()
    mel = mp.mel_spectrogram_torch(y, d["filter_length"], d["n_mel_channels"], d["sampling_rate"], d["hop_length"],
                                   d["win_length"], d["mel_fmin"], d["mel_fmax"])
    assert tuple(mel.shape) == (1, 80, 40)
    with torch.no_grad():
        o, mask, _ = net.infer(mel, torch.tensor([40], device="cuda"), noise_scale=0.667)
    torch.cuda.synchronize()
    assert tuple(o.shape) == (1, 1, n) and torch.isfinite(o).all() and float(o.abs().max()) <= 1.0


@pytest.mark.gpu
def test_gpu_notebook_cell4_replay_with_cpu_waveform(base_cfg, base_sd):
    """inference.ipynb cell 4, line for line: the waveform and the spectrogram are CPU tensors
    (`spectrogram_torch(audio_norm, ...)`, `spec_to_mel_torch(spec, ...)`), `mel.cuda()` comes afterwards.  The shim
    must take them and give back CPU tensors equal to what the CUDA-tensor path produces."""
    torch, mp = _gpu_mods()
    from models import SynthesizerTrn
    d = base_cfg["data"]
    net_g = SynthesizerTrn(d["filter_length"] // 2 + 1, base_cfg["train"]["segment_size"] // d["hop_length"],
                           n_speakers=d["n_speakers"], **base_cfg["model"]).cuda()
    _ = net_g.eval()
    net_g.load_state_dict({k: torch.from_numpy(v) for k, v in base_sd.items()})
    n = 30 * d["hop_length"]
    audio = (0.3 * 32768.0 * torch.sin(2 * np.pi * 330.0 * torch.arange(n) / d["sampling_rate"]))  # load_wav_to_torch: CPU float
    audio_norm = audio / 32768.0
    audio_norm = audio_norm.unsqueeze(0)
    spec = mp.spectrogram_torch(audio_norm, 1024, 22050, 256, 1024, center=False)
    mel = mp.spec_to_mel_torch(spec, d["filter_length"], d["n_mel_channels"], d["sampling_rate"], d["mel_fmin"], d["mel_fmax"])
    assert spec.device.type == "cpu" and mel.device.type == "cpu" and tuple(mel.shape) == (1, 80, 30)
    spec_d = mp.spectrogram_torch(audio_norm.cuda(), 1024, 22050, 256, 1024, center=False)
    assert torch.equal(spec, spec_d.cpu())
    with torch.no_grad():
        mel = mel.cuda()
        spec_lengths = torch.LongTensor([mel.size(2)]).cuda()
        audio_ = net_g.infer(mel, spec_lengths, sid=None, noise_scale=.667, noise_scale_w=0.8, length_scale=1)[0][0, 0].data.cpu().float().numpy()
    assert audio_.shape == (n,) and np.isfinite(audio_).all()
    # `.float()` is a no-op in the reference and must not unbind the uploaded weights
    net_g = net_g.float()
    net_g.to(torch.float32)
    with torch.no_grad():
        net_g.infer(mel, spec_lengths, noise_scale=.667)
    assert net_g.cpu()._handle is None


def test_shim_signatures_match_the_reference():
    """Drop-in boundary: same names, parameter order and defaults as the reference's mel_processing.py
    (signatures recorded by tests/golden/make_golden_mel.py from the reference itself)."""
    import inspect
    import json
    import os
    import mel_processing as mp
    from conftest import GOLDEN
    ref = json.load(open(os.path.join(GOLDEN, "mel_processing_signatures.json")))
    assert mp.MAX_WAV_VALUE == ref.pop("MAX_WAV_VALUE")
    for name, sig in ref.items():
        assert str(inspect.signature(getattr(mp, name))) == sig, name
