"""Windowed / chunked / streaming synthesis through the C ABI (svk_infer_window,
svk_infer_chunked).  The reference only synthesises whole utterances (models.py:331-339), so parity here is
equality with the whole-utterance svk_infer on the same inputs (which the other GPU tests pin against the
reference goldens) -- plus one direct check against the CPU oracle.

Bar: a window recomputed with svk_halo_frames() frames of context evaluates every kept output with the same
operands in the same order as the whole-utterance run, so the results must be IDENTICAL (torch.equal).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["tc", "fp32"])
def net(request, base_cfg, base_sd):
    from gpu_util import build_net
    return build_net(base_cfg["model"], base_sd, engine=request.param)


def _inputs(B, T, lengths, seed):
    rng = np.random.Generator(np.random.Philox(key=[seed, 9]))
    mel = torch.from_numpy((rng.standard_normal((B, 80, T)) * 2 - 5).astype(np.float32)).cuda()
    eps = torch.from_numpy(rng.standard_normal((B, 192, T)).astype(np.float32)).cuda()
    return mel, torch.tensor(lengths, dtype=torch.int64, device="cuda"), eps


def _full(net, mel, lengths, eps, noise_scale, max_len=None):
    from gpu_util import inject_eps
    with inject_eps(eps.cpu().numpy()), torch.no_grad():
        out = net.infer(mel, lengths, noise_scale=noise_scale, max_len=max_len)
    torch.cuda.synchronize()
    return out


def test_halo_is_the_receptive_field(net):
    # enc 16 x 2 + flow 4 x 8 x 2 + decoder 14 frames per side
    assert net.halo_frames() == 32 + 64 + 14


@pytest.mark.parametrize("chunk", [64, 100, 333])
def test_chunked_equals_whole_utterance(net, chunk):
    B, T = 2, 333  # > chunk + 2 * halo for the small chunks: interior windows with two artificial edges
    mel, lengths, eps = _inputs(B, T, [333, 201], 5)
    o, mask, lat = _full(net, mel, lengths, eps, 0.667)
    oc, maskc, latc = net.infer_chunked(mel, lengths, chunk_frames=chunk, noise_scale=0.667, eps=eps)
    torch.cuda.synchronize()
    assert oc.shape == o.shape and torch.equal(maskc, mask)
    d = float((oc - o).abs().max())
    print(f"chunk {chunk}: max |chunked - whole| = {d:.3e}")
    assert torch.equal(oc, o)
    for a, b, nm in zip(latc, lat, ("z", "z_p", "m_p", "logs_p")):
        assert torch.equal(a, b), nm


def test_chunked_with_max_len(net):
    B, T, ml = 2, 300, 257
    mel, lengths, eps = _inputs(B, T, [300, 280], 6)
    o, _, _ = _full(net, mel, lengths, eps, 0.5, max_len=ml)
    oc, _, latc = net.infer_chunked(mel, lengths, chunk_frames=96, noise_scale=0.5, max_len=ml, eps=eps)
    torch.cuda.synchronize()
    assert tuple(oc.shape) == (B, 1, 256 * ml) and all(v is None for v in latc)
    assert torch.equal(oc, o)


def test_stream_chunks_concatenate_to_whole(net):
    B, T = 1, 290
    mel, lengths, eps = _inputs(B, T, [290], 7)
    o, _, _ = _full(net, mel, lengths, eps, 0.667)
    chunks = list(net.infer_stream(mel, lengths, chunk_frames=48, noise_scale=0.667, eps=eps))
    torch.cuda.synchronize()
    assert len(chunks) == (T + 47) // 48 and all(c.shape[2] == 256 * 48 for c in chunks[:-1])
    assert torch.equal(torch.cat(chunks, dim=2), o)


def test_chunked_against_oracle(base_sd, base_dims, base_cfg):
    """Direct check of the windowed path against the CPU oracle (fp64), 1e-4 max-abs like every waveform test."""
    from gpu_util import build_net
    from oracle.oracle import Oracle
    net = build_net(base_cfg["model"], base_sd, engine="tc")
    B, T = 1, 150
    mel, lengths, eps = _inputs(B, T, [150], 8)
    oc, _, _ = net.infer_chunked(mel, lengths, chunk_frames=20, noise_scale=0.667, eps=eps)
    torch.cuda.synchronize()
    ro, _, _ = Oracle(np.float64).infer(base_sd, base_dims, mel.cpu().numpy(), lengths.cpu().numpy(), eps.cpu().numpy(), 0.667, None)
    err = float(np.abs(oc.cpu().numpy() - ro).max())
    print(f"chunked (20-frame windows) vs fp64 oracle: {err:.2e}")
    assert err <= 1e-4


def test_window_argument_errors(net):
    import svk_runtime as rt
    mel, lengths, eps = _inputs(1, 40, [40], 9)
    with pytest.raises(rt.SvkError):
        net.infer_chunked(mel, lengths, chunk_frames=0, eps=eps)
    ws = torch.empty(16, dtype=torch.uint8, device="cuda")
    o = torch.empty(1, 1, 256 * 8, device="cuda")
    with pytest.raises(rt.SvkError):  # workspace too small
        rt.check(rt.lib().svk_infer_window(net._handle.ptr, mel.data_ptr(), lengths.data_ptr(), eps.data_ptr(), 1.0, 1, 40, 0,
                                           0, 8, o.data_ptr(), 256 * 8, None, None, None, None, 0, ws.data_ptr(), 16, None))
    n = int(rt.lib().svk_window_workspace_bytes(net._handle.ptr, 1, 8))
    ws = torch.empty(n, dtype=torch.uint8, device="cuda")
    with pytest.raises(rt.SvkError):  # window outside [0, T)
        rt.check(rt.lib().svk_infer_window(net._handle.ptr, mel.data_ptr(), lengths.data_ptr(), eps.data_ptr(), 1.0, 1, 40, 0,
                                           36, 44, o.data_ptr(), 256 * 8, None, None, None, None, 0, ws.data_ptr(), n, None))
