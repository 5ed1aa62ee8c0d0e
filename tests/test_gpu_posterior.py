"""Analysis direction through the C ABI: PosteriorEncoder.forward (models.py:103-110) and
ResidualCouplingBlock.forward(reverse=False) (models.py:73-76), against the reference goldens
(tests/golden/make_golden_posterior.py) and the CPU oracle.  1e-4 max-abs like every fp32 tensor of the path;
the mask and the untouched coupling halves bit-exact."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module", params=["tc", "fp32"])
def net(request, base_cfg, base_sd):
    from gpu_util import build_net
    return build_net(base_cfg["model"], base_sd, engine=request.param)


def _np(t):
    return t.detach().cpu().numpy()


def test_posterior_encoder_matches_reference_golden(net):
    from gpu_util import dev, inject_eps
    g = load_golden("posterior_base_b2_t24")
    with inject_eps(g["eps"]), torch.no_grad():
        z, m, logs, mask = net.enc_q(dev(g["spec"]), dev(g["lengths"], torch.int64))
    torch.cuda.synchronize()
    assert np.array_equal(_np(mask), g["ref32_y_mask"])
    for nm, v in (("z", z), ("m_q", m), ("logs_q", logs)):
        err = np.abs(_np(v) - g["ref64_" + nm]).max()
        print(nm, "GPU vs ref64", err, "ref32 vs ref64", np.abs(g["ref32_" + nm] - g["ref64_" + nm]).max())
        assert err <= TOL, nm
    assert net.last_launch_count() > 15  # mask, pre, image, 16 fused WN layers, proj, sample


def test_flow_forward_matches_reference_golden(net):
    from gpu_util import dev
    g = load_golden("posterior_base_b2_t24")
    z = dev(g["ref64_z"].astype(np.float32))
    mask = dev(g["ref32_y_mask"])
    z_p = net.flow(z, mask)
    torch.cuda.synchronize()
    assert np.abs(_np(z_p) - g["ref64_z_p"]).max() <= TOL
    # bijection: the reverse pass (the infer path's flow) undoes it
    back = net.flow(z_p, mask, reverse=True)
    assert float((back - z).abs().max()) <= 1e-5


def test_flow_forward_vs_oracle_ragged_and_round_trip_at_full_size(net, base_sd, base_dims):
    from gpu_util import dev
    rng = np.random.Generator(np.random.Philox(key=[31, 2]))
    B, T = 3, 150
    z = rng.standard_normal((B, 192, T)).astype(np.float32)
    lengths = np.array([150, 0, 77], np.int64)
    mask = Oracle.sequence_mask(lengths, T).astype(np.float32)
    want = Oracle(np.float64).flow_forward(base_sd, base_dims, z * mask, mask.astype(np.float64))
    got = net.flow(dev(z * mask), dev(mask))
    assert np.abs(_np(got) - want).max() <= TOL
    # size-independent property at BASELINE configs[2]'s latent shape: reverse(forward(z)) == z
    g = torch.Generator(device="cuda").manual_seed(5)
    zz = torch.randn(16, 192, 1024, device="cuda", generator=g)
    mm = torch.ones(16, 1, 1024, device="cuda")
    back = net.flow(net.flow(zz, mm), mm, reverse=True)
    torch.cuda.synchronize()
    assert float((back - zz).abs().max()) <= 2e-5


def test_analysis_resynthesis_chain(net, base_sd, base_dims):
    """wav -> spectrogram_torch -> enc_q -> flow -> flow^-1 -> dec: the reconstruction path of models.py:317-329
    (without the random slice), every stage on the device through the C ABI; compared with the CPU oracle."""
    import mel_processing as mp
    from gpu_util import inject_eps
    n = 30 * 256
    t = torch.arange(n, device="cuda") / 22050.0
    y = (0.3 * torch.sin(2 * np.pi * 330.0 * t) + 0.1 * torch.sin(2 * np.pi * 1234.0 * t))[None].contiguous()
    spec = mp.spectrogram_torch(y, 1024, 22050, 256, 1024)
    lengths = torch.tensor([30], device="cuda")
    eps = np.random.Generator(np.random.Philox(key=[3, 3])).standard_normal((1, 192, 30)).astype(np.float32)
    with inject_eps(eps), torch.no_grad():
        z, m, logs, mask = net.enc_q(spec, lengths)
    z_p = net.flow(z, mask)
    z2 = net.flow(z_p, mask, reverse=True)
    o = net.dec(z2 * mask)
    torch.cuda.synchronize()
    orc = Oracle(np.float64)
    oz, om, ol, omask = orc.posterior_encoder(base_sd, base_dims, _np(spec), _np(lengths), eps)
    ozp = orc.flow_forward(base_sd, base_dims, oz, omask)
    oo = orc.generator(base_sd, base_dims, oz * omask)
    assert np.abs(_np(z) - oz).max() <= TOL and np.abs(_np(z_p) - ozp).max() <= TOL
    assert np.abs(_np(o) - oo).max() <= TOL


def test_posterior_needs_enc_q_weights(base_cfg, base_sd, base_dims):
    """svk_infer works without enc_q.*; svk_posterior_encoder then fails loudly."""
    import svk_runtime as rt
    import svk_weights as W
    h = rt.Handle(base_dims, 0)
    for k, v in base_sd.items():
        if not W.is_dead_key(k):
            h.load_tensor(k, v)
    h.finalize()
    live, loaded = h.weight_status()
    assert live == loaded
    x = torch.zeros(1, 513, 8, device="cuda")
    lengths = torch.tensor([8], device="cuda")
    eps = torch.zeros(1, 192, 8, device="cuda")
    outs = [torch.empty(1, 192, 8, device="cuda") for _ in range(3)]
    mask = torch.empty(1, 1, 8, device="cuda")
    ws = torch.empty(1 << 22, dtype=torch.uint8, device="cuda")
    st = rt.lib().svk_posterior_encoder(h.ptr, x.data_ptr(), lengths.data_ptr(), eps.data_ptr(), 1, 8, outs[0].data_ptr(),
                                        outs[1].data_ptr(), outs[2].data_ptr(), mask.data_ptr(), ws.data_ptr(), ws.numel(), None)
    assert st == rt.SVK_ERR_STATE
    h.close()
