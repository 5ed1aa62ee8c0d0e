"""GPU parity tests (run with -m gpu on the B200 box): every check goes through the C ABI of
libsvk.so and compares with the CPU oracle / the reference-generated golden vectors.

Tolerances: fp32 path, 1e-4 max-abs on the waveform and latents (BASELINE north_star); integer /
indexing work (mask, Flip, split/cat pass-through, spline bins) bit-exact.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(scope="module", params=["tc", "fp32"])
def net(request, base_cfg, base_sd):
    """Both engines of libsvk must meet the same bar: "tc" = tcgen05 fp16x3 split (default), "fp32" = FFMA."""
    from gpu_util import build_net
    return build_net(base_cfg["model"], base_sd, engine=request.param)


def _np(t):
    return t.detach().cpu().numpy()


def _run_infer(net, g):
    from gpu_util import dev, inject_eps
    ml = int(g["max_len"])
    with inject_eps(g["eps"]), torch.no_grad():
        o, mask, lat = net.infer(dev(g["mel"]), dev(g["lengths"], torch.int64), noise_scale=float(g["noise_scale"]),
                                 max_len=None if ml < 0 else ml)
    torch.cuda.synchronize()
    return o, mask, lat


@pytest.mark.parametrize("name", ["infer_base_b2_t40", "infer_base_b1_t12_maxlen9", "infer_base_b3_t3"])
def test_infer_matches_reference_golden(net, name):
    """Padded ragged batch, max_len < T, T shorter than every halo."""
    g = load_golden(name)
    o, mask, (z, z_p, m_p, logs_p) = _run_infer(net, g)
    assert o.shape == g["ref64_o"].shape
    assert np.array_equal(_np(mask), g["ref32_x_mask"])  # integer compare: bit-exact
    assert np.abs(_np(o) - g["ref64_o"]).max() <= TOL
    for nm, v in (("z", z), ("z_p", z_p), ("m_p", m_p), ("logs_p", logs_p)):
        assert np.abs(_np(v) - g["ref64_" + nm]).max() <= TOL, nm
    # report how close we are to the reference's own fp32 noise floor
    print(name, "GPU vs ref64:", np.abs(_np(o) - g["ref64_o"]).max(), " ref32 vs ref64:",
          np.abs(g["ref32_o"] - g["ref64_o"]).max())
    assert net.last_launch_count() > 100  # our kernels actually ran


def test_infer_trace_stages(net):
    """Module-level intermediates of the golden trace: flow after each coupling, decoder stages."""
    from gpu_util import dev
    g = load_golden("infer_base_b1_t12_maxlen9")
    mask = dev(g["ref32_x_mask"])
    z = net.flow(dev(g["ref64_z_p"]), mask, reverse=True)
    assert np.abs(_np(z) - g["trace_flow3"]).max() <= TOL
    o = net.dec(dev((g["ref64_z"] * g["ref32_x_mask"])[:, :, :9]))
    assert np.abs(_np(o) - g["ref64_o"]).max() <= TOL
    xo, m, logs, mk = net.enc_p(dev(g["mel"]), dev(g["lengths"], torch.int64))
    assert np.abs(_np(m) - g["ref64_m_p"]).max() <= TOL and np.abs(_np(logs) - g["ref64_logs_p"]).max() <= TOL
    assert np.array_equal(_np(mk), g["ref32_x_mask"])


def test_infer_vs_oracle_ragged(net, base_sd, base_dims):
    """Fresh shapes not in the fixtures: tile-unaligned T, zero-length item, max_len cut."""
    from gpu_util import dev, inject_eps
    rng = np.random.Generator(np.random.Philox(key=[77, 1]))
    B, T, ml = 3, 70, 61
    mel = (rng.standard_normal((B, 80, T)) * 2 - 5).astype(np.float32)
    eps = rng.standard_normal((B, 192, T)).astype(np.float32)
    lengths = np.array([70, 0, 33], np.int64)
    with inject_eps(eps), torch.no_grad():
        o, mask, (z, z_p, m_p, logs_p) = net.infer(dev(mel), dev(lengths, torch.int64), noise_scale=0.5, max_len=ml)
    ro, rmask, (rz, rzp, rm, rl) = Oracle(np.float64).infer(base_sd, base_dims, mel, lengths, eps, 0.5, ml)
    assert o.shape == ro.shape == (B, 1, 256 * ml)
    assert np.array_equal(_np(mask), rmask.astype(np.float32))
    assert np.abs(_np(o) - ro).max() <= TOL
    assert np.abs(_np(z) - rz).max() <= TOL and np.abs(_np(z_p) - rzp).max() <= TOL


def _sd_with_live_couplings(base_dims, live):
    """The seeded recipe with `post` (weight and bias) of every coupling NOT in `live` zero, as the reference initialises
    it (modules.py:321-322): such a coupling is the identity on x1 apart from the mask."""
    import svk_weights as W
    sd = W.make_state_dict(base_dims, seed=1234)
    for f in range(base_dims.n_flows):
        if f not in live:
            for leaf in ("weight", "bias"):
                sd[f"flow.flows.{2 * f}.post.{leaf}"] = np.zeros_like(sd[f"flow.flows.{2 * f}.post.{leaf}"])
    return sd


def _flow_inputs(B=2, T=150, lengths=(150, 97), seed=31):
    rng = np.random.Generator(np.random.Philox(key=[seed, T]))
    z_p = rng.standard_normal((B, 192, T)).astype(np.float32)
    mask = Oracle.sequence_mask(np.asarray(lengths, np.int64), T).astype(np.float32)
    return z_p, mask


@pytest.mark.parametrize("engine", ["tc", "fp32"])
def test_identity_flow_is_bit_exact_on_gpu(engine, base_cfg, base_dims):
    """north_star: 'bit-identical on the flow's permutation/indexing ops'.  In the product path the four Flips are folded
    into weight-index permutations and a reversed-channel store, and split/cat into channel offsets -- so the standalone
    svk_flip test says nothing about them.  With zero-initialised `post` (the reference's own init) every coupling is
    x1 -> x1 * mask, the whole reverse flow is z_p * mask, and the GPU must reproduce that BIT FOR BIT through all four
    folded Flips, on two time tiles with a ragged row."""
    from gpu_util import build_net, dev, inject_eps
    net0 = build_net(base_cfg["model"], _sd_with_live_couplings(base_dims, live=()), engine=engine)
    z_p, mask = _flow_inputs()
    want = z_p * mask
    live_rows = np.broadcast_to(mask != 0, z_p.shape)
    for reverse in (True, False):
        z = _np(net0.flow(dev(z_p), dev(mask), reverse=reverse))
        assert np.array_equal(z, want)  # torch.equal semantics (a masked zero may differ in sign: -0 + 0 = +0)
        assert z[live_rows].tobytes() == z_p[live_rows].tobytes()  # every unmasked element bit for bit
    # and through infer: the returned z equals the returned z_p * x_mask
    rng = np.random.Generator(np.random.Philox(key=[32, 1]))
    mel = (rng.standard_normal((2, 80, 150)) * 2 - 5).astype(np.float32)
    eps = rng.standard_normal((2, 192, 150)).astype(np.float32)
    with inject_eps(eps), torch.no_grad():
        _, xm, (z2, zp2, _, _) = net0.infer(dev(mel), dev(np.array([150, 97]), torch.int64), noise_scale=0.667)
    assert torch.equal(z2, zp2 * xm)


@pytest.mark.parametrize("live", [0, 1, 2, 3])
def test_single_live_coupling_pass_through_half_bit_exact(net, live, base_cfg, base_dims):
    """One coupling alive, the other three identities: its x0 half must come out as z_p * mask bit for bit, in the
    channel positions the Flips put it (coupling f sees the input after n_flows - f Flips, models.py:77-79: odd ->
    x0 is the reversed upper half), while its x1 half carries x1 - m within 1e-4 of the fp64 oracle."""
    from gpu_util import build_net, dev
    sd = _sd_with_live_couplings(base_dims, live=(live,))
    net1 = build_net(base_cfg["model"], sd, engine=net.engine)
    z_p, mask = _flow_inputs(seed=40 + live)
    z = _np(net1.flow(dev(z_p), dev(mask), reverse=True))
    ref = Oracle(np.float64).flow_reverse(sd, base_dims, z_p.astype(np.float64), mask.astype(np.float64))
    x0 = slice(0, 96) if (base_dims.n_flows - live) % 2 == 0 else slice(96, 192)
    x1 = slice(96, 192) if x0.start == 0 else slice(0, 96)
    # value equality like torch.equal: under the mask a zero may be +0 where z_p * 0 is -0 (the identity couplings add
    # post(h) = +0 to x1 = -0); every unmasked element is compared bit for bit
    assert np.array_equal(z[:, x0], (z_p * mask)[:, x0])
    live_rows = np.broadcast_to(mask != 0, z.shape)[:, x0]
    assert z[:, x0][live_rows].tobytes() == z_p[:, x0][live_rows].tobytes()
    assert np.abs(z[:, x1] - ref[:, x1]).max() <= TOL
    assert np.abs(ref[:, x1] - (z_p * mask)[:, x1]).max() > 1e-2  # the live half really moved


def test_flow_tail_against_reference_trace(net, base_sd, base_dims):
    """All couplings alive (reference trace of infer_base_b1_t12_maxlen9): the GPU's reverse flow from the reference's
    z_p agrees with the reference's state after the last coupling on BOTH halves within 1e-4.  (Bit-exactness of the
    pass-through half is pinned by the two tests above, where the GPU's own input to that half is known.)"""
    from gpu_util import dev
    g = load_golden("infer_base_b1_t12_maxlen9")
    z = _np(net.flow(dev(g["ref64_z_p"]), dev(g["ref32_x_mask"]), reverse=True))
    assert np.abs(z - g["trace_flow3"]).max() <= TOL
    assert np.abs(z - g["ref64_z"]).max() <= TOL


def test_infer_matches_reference_golden_multitile(net):
    """B=3, T=300, lengths (300, 257, 129): tile borders inside every utterance, ragged rows (F10), generated by the
    unmodified reference (tests/golden/make_golden_multitile.py)."""
    g = load_golden("infer_base_b3_t300_ragged")
    o, mask, (z, z_p, m_p, logs_p) = _run_infer(net, g)
    assert np.array_equal(_np(mask), g["ref32_x_mask"])
    e_o = np.abs(_np(o) - g["ref64_o"]).max()
    print("multi-tile golden: GPU vs ref64 on o %.2e (reference fp32 vs fp64 %.2e); z %.2e" %
          (e_o, float(g["ref32_vs_ref64_o"]), np.abs(_np(z) - g["ref64_z"]).max()))
    assert e_o <= TOL
    assert np.abs(_np(z) - g["ref64_z"]).max() <= TOL and np.abs(_np(m_p) - g["ref64_m_p"]).max() <= TOL


# ------------------------------------------------------------------------ operand range of the fp16 hi/lo engine
def _rel_err(y, ref):
    return float(np.abs(y - ref).max() / np.sqrt((ref ** 2).mean()))


def test_cold_recipe_relative_error(net, base_cfg, base_dims):
    """Reference random init (zero `post`, unit decoder gains): |o| ~ 0.02, activations down to 1e-3.  The
    absolute 1e-4 bar is vacuous there, so the error is stated RELATIVE to the signal's rms: the lo halves of the
    fp16 pairs are stored scaled by 2^11 (tc_common.cuh), which keeps the pairs exact to 22 bits down to |x| = 6e-5."""
    import svk_weights as W
    from gpu_util import build_net, dev, inject_eps
    sd = W.make_state_dict(base_dims, seed=1234, alive=False)
    netc = build_net(base_cfg["model"], sd, engine=net.engine)
    rng = np.random.Generator(np.random.Philox(key=[51, 2]))
    B, T = 2, 40
    mel = (rng.standard_normal((B, 80, T)) * 2 - 5).astype(np.float32)
    eps = rng.standard_normal((B, 192, T)).astype(np.float32)
    lengths = np.array([40, 31], np.int64)
    with inject_eps(eps), torch.no_grad():
        o = _np(netc.infer(dev(mel), dev(lengths, torch.int64), noise_scale=0.667)[0])
    ro = Oracle(np.float64).infer(sd, base_dims, mel, lengths, eps, 0.667, None)[0]
    ro32 = Oracle(np.float32).infer(sd, base_dims, mel, lengths, eps, 0.667, None)[0]
    print("cold recipe (%s): |o| max %.4f rms %.4f; GPU rel err %.2e (abs %.2e); fp32 CPU oracle rel err %.2e" %
          (net.engine, np.abs(ro).max(), np.sqrt((ro ** 2).mean()), _rel_err(o, ro), np.abs(o - ro).max(), _rel_err(ro32, ro)))
    assert np.abs(o - ro).max() <= TOL
    assert _rel_err(o, ro) <= 2e-4  # of the signal's rms


@pytest.mark.parametrize("scale", [1e-3, 1.0, 1e3])
@pytest.mark.parametrize("idx", [4, 11])
def test_resblock_relative_error_over_operand_range(net, base_sd, base_dims, idx, scale):
    """A ResBlock1 is positively homogeneous up to its biases, so feeding x * scale sweeps the magnitude of every operand
    image: small (lo halves would be fp16 subnormals unscaled), nominal, large (|activations| ~ 1e4, near the top of the
    fp16 range).  Error is reported relative to the output's rms and must stay at fp32 class on all three."""
    import svk_runtime as rt
    from gpu_util import dev
    C = {4: 128, 11: 32}[idx]
    L = 700
    rng = np.random.Generator(np.random.Philox(key=[61, idx]))
    x = (rng.standard_normal((1, C, L)) * scale).astype(np.float32)
    k = base_dims.resblock_kernel_sizes[idx % 3]
    ref = Oracle(np.float64).resblock1(base_sd, f"dec.resblocks.{idx}", x, k, (1, 3, 5))
    xd, y = dev(x), torch.empty(1, C, L, device="cuda")
    ws = torch.empty(rt.lib().svk_resblock1_workspace_bytes(net._handle.ptr, idx, 1, L), dtype=torch.uint8, device="cuda")
    rt.check(rt.lib().svk_resblock1(net._handle.ptr, idx, xd.data_ptr(), 1, L, y.data_ptr(), ws.data_ptr(), ws.numel(),
                                    torch.cuda.current_stream().cuda_stream))
    rel = _rel_err(_np(y), ref)
    print("resblock %d (%s) at input scale %g: |y| max %.3g, rel err %.2e" % (idx, net.engine, scale, np.abs(ref).max(), rel))
    assert rel <= 2e-5


def test_hot_recipe_fails_loudly(base_cfg, base_dims):
    """Decoder gains x2 on top of the alive recipe: the reference's own fp32 activations reach 3e7 (measured with forward
    hooks), far above fp16's 65504.  The tensor-core engine must not hand back NaN audio silently: infer raises
    SVK_ERR_RANGE (svk_check_range); with range_check off the NaNs are there and check_range() reports them; the fp32
    FFMA engine computes the (saturated) waveform."""
    import svk_runtime as rt
    import svk_weights as W
    from gpu_util import build_net, dev, inject_eps
    sd = W.make_state_dict(base_dims, seed=1234)
    for k in sd:
        if k.startswith("dec.") and k.endswith("weight_g"):
            sd[k] = sd[k] * np.float32(2.0)
    rng = np.random.Generator(np.random.Philox(key=[5, 5]))
    mel = (rng.standard_normal((1, 80, 24)) * 2 - 5).astype(np.float32)
    eps = rng.standard_normal((1, 192, 24)).astype(np.float32)
    ln = np.array([24], np.int64)
    hot = build_net(base_cfg["model"], sd, engine="tc")
    with inject_eps(eps), torch.no_grad():
        with pytest.raises(rt.SvkError) as ei:
            hot.infer(dev(mel), dev(ln, torch.int64), noise_scale=0.667)
    assert ei.value.code == rt.SVK_ERR_RANGE
    hot.range_check = False
    with inject_eps(eps), torch.no_grad():
        o = hot.infer(dev(mel), dev(ln, torch.int64), noise_scale=0.667)[0]
    assert not torch.isfinite(o).all()
    with pytest.raises(rt.SvkError):
        hot.check_range()
    hot.check_range()  # the flag is cleared by the failing check
    with pytest.raises(rt.SvkError):  # host entry point checks before it returns
        hot.infer_host(mel, ln, eps, 0.667)
    ok = build_net(base_cfg["model"], sd, engine="fp32")
    with inject_eps(eps), torch.no_grad():
        o32 = ok.infer(dev(mel), dev(ln, torch.int64), noise_scale=0.667)[0]
    ro = Oracle(np.float32).infer(sd, base_dims, mel, ln, eps, 0.667, None)[0]
    # a saturated square wave: only samples whose pre-tanh value happens to be O(1) among values of 1e6 can differ
    assert torch.isfinite(o32).all() and (np.abs(_np(o32) - ro) > 1e-3).mean() <= 1e-3


def test_wn_fused_layer_equals_two_launch_form(base_cfg, base_sd, monkeypatch):
    """wn_layer.cu (in_layer + gate + res_skip + residual/skip in ONE launch, acts in shared memory) performs the same
    arithmetic in the same order as the two-launch form ($SVK_FUSE_WN=0: conv_tc MODE_GATE + split STORE epilogue), so
    the encoder and the flow must agree BIT FOR BIT -- on a ragged batch spanning three time tiles."""
    from gpu_util import build_net, dev
    g = load_golden("infer_base_b3_t300_ragged")
    outs = []
    # 2 = fused whatever the batch size (the default fuses when the batch fills the GPU); SVK_WN_STACK: the whole stack in
    # one launch (tiles synchronise with their neighbours inside the kernel) / one launch per layer / the two-launch form
    for fuse, stack in (("2", "1"), ("2", "0"), ("0", "0")):
        monkeypatch.setenv("SVK_FUSE_WN", fuse)
        monkeypatch.setenv("SVK_WN_STACK", stack)
        n = build_net(base_cfg["model"], base_sd, engine="tc")
        xo, m, logs, mask = n.enc_p(dev(g["mel"]), dev(g["lengths"], torch.int64))
        z = n.flow(dev(g["ref64_z"]), mask, reverse=False)
        torch.cuda.synchronize()
        outs.append((_np(xo), _np(m), _np(logs), _np(z), n.last_launch_count()))
    for other in outs[1:]:
        for a, b in zip(outs[0][:4], other[:4]):
            assert np.array_equal(a, b)
    assert outs[0][4] < outs[1][4] < outs[2][4]  # fewer launches
    assert np.abs(outs[0][1] - g["ref64_m_p"]).max() <= TOL


def test_determinism(net):
    g = load_golden("infer_base_b2_t40")
    o1, _, _ = _run_infer(net, g)
    o2, _, _ = _run_infer(net, g)
    assert torch.equal(o1, o2)


def test_host_entry_matches_device_entry(net):
    """svk_infer_host (H2D + infer + D2H inside the library) == svk_infer on device tensors."""
    g = load_golden("infer_base_b2_t40")
    o_dev, mask_dev, lat_dev = _run_infer(net, g)
    o, mask, lat = net.infer_host(g["mel"], g["lengths"], g["eps"], float(g["noise_scale"]), None, want_latents=True)
    assert np.array_equal(o, _np(o_dev)) and np.array_equal(mask, _np(mask_dev))
    for a, b in zip(lat, lat_dev):
        assert np.array_equal(a, _np(b))


def test_seeded_rng_path(net):
    """Without injection the draw comes from torch's CUDA generator, like torch.randn_like in the reference."""
    from gpu_util import dev
    mel = dev(np.full((1, 80, 8), -5.0, np.float32))
    ln = dev(np.array([8]), torch.int64)
    torch.manual_seed(3)
    o1 = net.infer(mel, ln, noise_scale=0.667)[0]
    torch.manual_seed(3)
    o2 = net.infer(mel, ln, noise_scale=0.667)[0]
    o3 = net.infer(mel, ln, noise_scale=0.667)[0]
    assert torch.equal(o1, o2) and not torch.equal(o1, o3)
    o4 = net.infer(mel, ln, noise_scale=0.0)[0]
    o5 = net.infer(mel, ln, noise_scale=0.0)[0]
    assert torch.equal(o4, o5)


# ------------------------------------------------------------------------------- operators
@pytest.mark.parametrize("k,dil,cin,cout,L", [(1, 1, 80, 192, 50), (5, 1, 192, 384, 37), (7, 1, 192, 512, 12),
                                              (3, 3, 256, 256, 300), (7, 5, 64, 64, 1000), (11, 5, 32, 32, 700),
                                              (11, 1, 128, 128, 257), (7, 1, 32, 1, 3000), (3, 1, 8, 96, 5)])
def test_conv1d_vs_oracle(k, dil, cin, cout, L):
    _conv1d_case(k, dil, cin, cout, L, "fp32")


@pytest.mark.parametrize("k,dil,cin,cout,L", [(1, 1, 96, 192, 50), (5, 1, 192, 384, 137), (7, 1, 192, 512, 12),
                                              (3, 3, 256, 256, 300), (7, 5, 64, 64, 1000), (11, 5, 32, 32, 700),
                                              (11, 1, 128, 128, 257), (7, 1, 32, 16, 3000), (1, 1, 192, 96, 129),
                                              (2, 1, 512, 2048, 40), (3, 1, 32, 40, 5)])
def test_conv1d_tc_vs_oracle(k, dil, cin, cout, L):
    """tcgen05 engine: taps as row-shifted descriptors, every tile width the model uses (N = 16..128,
    multi-tile Cout, ragged last tile), time tiles with ragged ends."""
    _conv1d_case(k, dil, cin, cout, L, "tc")


def _conv1d_case(k, dil, cin, cout, L, engine):
    from gpu_util import conv1d, dev
    rng = np.random.Generator(np.random.Philox(key=[k, dil * 1000 + cin]))
    x = rng.standard_normal((2, cin, L)).astype(np.float32)
    w = (rng.standard_normal((cout, cin, k)) / np.sqrt(cin * k)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    pad = (k * dil - dil) // 2
    orc = Oracle(np.float64)
    ref = orc.conv1d(orc.leaky_relu(x, 0.1), w, b, dil, pad)
    y = _np(conv1d(dev(x), dev(w), dev(b), dil, pad, pre_slope=0.1, engine=engine))
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 2e-5
    y2 = _np(conv1d(dev(x), dev(w), None, dil, pad, engine=engine))
    assert np.abs(y2 - orc.conv1d(x, w, None, dil, pad)).max() <= 2e-5


@pytest.mark.parametrize("cin,cout,k,s,L", [(512, 256, 16, 8, 12), (256, 128, 16, 8, 96), (128, 64, 4, 2, 700),
                                            (64, 32, 4, 2, 1025), (16, 8, 6, 2, 33), (16, 8, 3, 1, 9)])
def test_conv_transpose1d_vs_oracle(cin, cout, k, s, L):
    _convt_case(cin, cout, k, s, L, "fp32")


@pytest.mark.parametrize("cin,cout,k,s,L", [(512, 256, 16, 8, 12), (256, 128, 16, 8, 196), (128, 64, 4, 2, 700),
                                            (64, 32, 4, 2, 1025), (32, 8, 6, 2, 33), (32, 16, 3, 1, 9)])
def test_conv_transpose1d_tc_vs_oracle(cin, cout, k, s, L):
    _convt_case(cin, cout, k, s, L, "tc")


def _convt_case(cin, cout, k, s, L, engine):
    from gpu_util import conv_transpose1d, dev
    rng = np.random.Generator(np.random.Philox(key=[cin, k * 100 + s]))
    x = rng.standard_normal((2, cin, L)).astype(np.float32)
    w = (rng.standard_normal((cin, cout, k)) / np.sqrt(cin)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    p = (k - s) // 2
    orc = Oracle(np.float64)
    ref = orc.conv_transpose1d(orc.leaky_relu(x, 0.1), w, b, s, p)
    y = _np(conv_transpose1d(dev(x), dev(w), dev(b), s, p, pre_slope=0.1, engine=engine))
    assert y.shape == ref.shape
    assert np.abs(y - ref).max() <= 2e-5


def test_sequence_mask_and_flip_bit_exact():
    from gpu_util import dev, flip, sequence_mask
    lengths = np.array([0, 1, 17, 64, 65, 1000], np.int64)
    m = _np(sequence_mask(dev(lengths, torch.int64), 65))
    assert np.array_equal(m[:, None, :], Oracle.sequence_mask(lengths, 65))
    x = np.random.Generator(np.random.Philox(key=[1, 2])).standard_normal((3, 192, 33)).astype(np.float32)
    assert np.array_equal(_np(flip(dev(x))), Oracle.flip(x))


def test_weight_norm_vs_oracle():
    from gpu_util import dev, weight_norm
    rng = np.random.Generator(np.random.Philox(key=[3, 4]))
    v = rng.standard_normal((512, 256, 16)).astype(np.float32)
    g = rng.uniform(0.5, 2, (512, 1, 1)).astype(np.float32)
    ref = Oracle(np.float64).weight_norm(v, g)
    assert np.abs(_np(weight_norm(dev(v), dev(g))) - ref).max() <= 1e-6


@pytest.mark.parametrize("inverse", [False, True])
def test_rq_spline_matches_reference(inverse):
    from gpu_util import dev, rq_spline
    g = load_golden("rq_spline")
    tag = f"inv{int(inverse)}"
    y, lad, bins = rq_spline(dev(g["x"]), dev(g["uw"]), dev(g["uh"]), dev(g["ud"]), inverse)
    y, lad, bins = _np(y), _np(lad), _np(bins)
    assert np.abs(y - g["y64_" + tag]).max() <= 2e-5
    assert np.abs(lad - g["lad64_" + tag]).max() <= 2e-4
    out = np.abs(g["x"]) > 5.0
    assert np.array_equal(y[out], g["x"][out]) and np.all(lad[out] == 0)  # identity tails, bit-exact
    _, _, obins = Oracle(np.float32).rq_spline(g["x"], g["uw"], g["uh"], g["ud"], inverse)
    assert np.array_equal(bins, obins)  # searchsorted index: integer work


def test_convflow_matches_reference_golden():
    """modules.ConvFlow(4, 32, 3, 3) (pre -> 3 x DDSConv layer -> proj -> spline -> cat) against the reference's fp64
    outputs: forward, its log-determinant, and the reverse pass; the bin indices (integer work) against the fp32 oracle."""
    from gpu_util import convflow, dev
    g = load_golden("convflow")
    w = {k[2:]: v for k, v in g.items() if k.startswith("w_")}
    mask_np = Oracle.sequence_mask(g["lengths"], g["x"].shape[2]).astype(np.float32)
    x, mask = dev(g["x"]), dev(mask_np[:, 0])
    y, logdet, bins = convflow(x, mask, w, 32, 3, 3, reverse=False)
    assert np.abs(_np(y) - g["fwd64"]).max() <= 2e-5
    assert np.abs(_np(logdet) - g["logdet64"]).max() <= 2e-4
    yr, _, bins_r = convflow(x, mask, w, 32, 3, 3, reverse=True)
    # the inverse solves a quadratic per element and amplifies parameter noise where a bin is nearly flat: fp32 bar (1e-4)
    assert np.abs(_np(yr) - g["rev64"]).max() <= TOL
    print("ConvFlow GPU vs ref64: fwd %.2e logdet %.2e rev %.2e (ref32 vs ref64: %.2e / %.2e / %.2e)" % (
        np.abs(_np(y) - g["fwd64"]).max(), np.abs(_np(logdet) - g["logdet64"]).max(), np.abs(_np(yr) - g["rev64"]).max(),
        np.abs(g["fwd32"] - g["fwd64"]).max(), np.abs(g["logdet32"] - g["logdet64"]).max(), np.abs(g["rev32"] - g["rev64"]).max()))
    orc = Oracle(np.float32)
    _, _, ob = orc.convflow(w, g["x"], mask_np, 32, 3, 3, reverse=False)
    _, _, obr = orc.convflow(w, g["x"], mask_np, 32, 3, 3, reverse=True)
    assert np.array_equal(_np(bins), ob) and np.array_equal(_np(bins_r), obr)
    # x0 passes through untouched apart from the mask: bit for bit
    assert np.array_equal(_np(y)[:, :2], g["x"][:, :2] * mask_np)
    # the flow is a bijection where the mask is on: reverse(forward(x)) == x * mask
    back, _, _ = convflow(y, mask, w, 32, 3, 3, reverse=True)
    assert np.abs(_np(back) - g["x"] * mask_np).max() <= 1e-4


def test_convflow_module_shim_matches_reference_surface():
    """svk_modules.ConvFlow: the reference's constructor, state_dict key names and forward signature / return values."""
    from svk_modules import ConvFlow
    g = load_golden("convflow")
    w = {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w_")}
    cf = ConvFlow(4, 32, 3, 3)
    assert list(cf.state_dict().keys()) == list(w.keys())  # same names, same order as modules.ConvFlow(...).state_dict()
    assert float(cf.state_dict()["proj.weight"].abs().max()) == 0.0  # zero-initialised like modules.py:358-359
    cf.load_state_dict(w)
    cf = cf.cuda().eval()
    mask = torch.from_numpy(Oracle.sequence_mask(g["lengths"], 37).astype(np.float32)).cuda()
    y, logdet = cf(torch.from_numpy(g["x"]).cuda(), mask)
    assert np.abs(_np(y) - g["fwd64"]).max() <= 2e-5 and np.abs(_np(logdet) - g["logdet64"]).max() <= 2e-4
    back = cf(y, mask, reverse=True)
    assert np.abs(_np(back) - g["x"] * _np(mask)).max() <= TOL


def test_convflow_vs_oracle_wide():
    """VITS-sized instance (in_channels 2, filter_channels 192, 3 layers: the widest tile the kernels stage) on a ragged
    batch spanning several 64-step tiles, against the fp64 oracle; and with the reference's zero-initialised proj (modules.py:358-359)."""
    from gpu_util import convflow, dev
    rng = np.random.Generator(np.random.Philox(key=[192, 3]))
    B, C, T, F, k, n = 2, 2, 150, 192, 3, 3
    w = {"pre.weight": rng.standard_normal((F, C // 2, 1)) * 0.5, "pre.bias": rng.standard_normal(F) * 0.1,
         "proj.weight": rng.standard_normal((C // 2 * 29, F, 1)) * 0.05, "proj.bias": rng.standard_normal(C // 2 * 29) * 0.1}
    for i in range(n):
        w[f"convs.convs_sep.{i}.weight"] = rng.standard_normal((F, 1, k)) * 0.5
        w[f"convs.convs_sep.{i}.bias"] = rng.standard_normal(F) * 0.1
        w[f"convs.convs_1x1.{i}.weight"] = rng.standard_normal((F, F, 1)) / np.sqrt(F)
        w[f"convs.convs_1x1.{i}.bias"] = rng.standard_normal(F) * 0.1
        for nm in ("norms_1", "norms_2"):
            w[f"convs.{nm}.{i}.gamma"] = 1 + 0.1 * rng.standard_normal(F)
            w[f"convs.{nm}.{i}.beta"] = 0.1 * rng.standard_normal(F)
    w = {kk: v.astype(np.float32) for kk, v in w.items()}
    xx = (rng.standard_normal((B, C, T)) * 2.5).astype(np.float32)
    mask_np = Oracle.sequence_mask(np.array([150, 70]), T).astype(np.float32)
    y, logdet, _ = convflow(dev(xx), dev(mask_np[:, 0]), w, F, k, n)
    ry, rl, _ = Oracle(np.float64).convflow(w, xx, mask_np, F, k, n)
    assert np.abs(_np(y) - ry).max() <= 1e-4 and np.abs(_np(logdet) - rl).max() <= 2e-3
    # zero-initialised proj (modules.py:358-359): every element sees the same parameter-free spline (uniform bins, interior
    # derivative 1e-3 + softplus(0)); the tails beyond +-5 are the identity bit for bit
    w0 = dict(w)
    w0["proj.weight"], w0["proj.bias"] = np.zeros_like(w["proj.weight"]), np.zeros_like(w["proj.bias"])
    y0, l0, _ = convflow(dev(xx), dev(mask_np[:, 0]), w0, F, k, n)
    ry0, rl0, _ = Oracle(np.float64).convflow(w0, xx, mask_np, F, k, n)
    assert np.abs(_np(y0) - ry0).max() <= 1e-5 and np.abs(_np(l0) - rl0).max() <= 1e-3
    outside = np.broadcast_to(np.abs(xx) > 5.0, xx.shape).copy()
    outside[:, 0] = True  # x0 is never transformed
    assert np.array_equal(_np(y0)[outside], (xx * mask_np)[outside])


def test_rq_spline_round_trip_large():
    """Size-independent property at scale: inverse(forward(x)) == x, logdets cancel."""
    from gpu_util import rq_spline
    gen = torch.Generator(device="cuda").manual_seed(5)
    n = 1 << 20
    x = (torch.rand(n, device="cuda", generator=gen) * 12 - 6)
    uw = torch.randn(n, 10, device="cuda", generator=gen)
    uh = torch.randn(n, 10, device="cuda", generator=gen)
    ud = torch.randn(n, 9, device="cuda", generator=gen)
    y, lad, _ = rq_spline(x, uw, uh, ud, False)
    xr, lad2, _ = rq_spline(y, uw, uh, ud, True)
    # fp32 inversion is ill-conditioned where a bin is nearly flat (the reference's own fp32 round trip on
    # ConvFlow is 2e-5 over 148 points, tests/golden/make_golden.py); bound the bulk tightly, the tail loosely
    err = (xr - x).abs()
    assert torch.quantile(err[:1 << 16], 0.999).item() <= 5e-4
    assert err.max().item() <= 5e-2
    assert torch.quantile((lad + lad2).abs()[:1 << 16], 0.999).item() <= 2e-3
    assert torch.equal(y[x.abs() > 5], x[x.abs() > 5])
    # and a 20k-element slice against the fp32 CPU oracle, element for element
    n2 = 20000
    oy, olad, obins = Oracle(np.float32).rq_spline(_np(x[:n2]), _np(uw[:n2]), _np(uh[:n2]), _np(ud[:n2]), False)
    assert np.abs(_np(y[:n2]) - oy).max() <= 1e-4


def test_resblock_and_generator_vs_oracle(net, base_sd, base_dims):
    import svk_runtime as rt
    from gpu_util import dev
    rng = np.random.Generator(np.random.Philox(key=[9, 9]))
    orc = Oracle(np.float64)
    for idx, (C, L) in ((0, (256, 40)), (4, (128, 333)), (8, (64, 600)), (11, (32, 1100))):
        x = rng.standard_normal((2, C, L)).astype(np.float32)
        k = base_dims.resblock_kernel_sizes[idx % 3]
        ref = orc.resblock1(base_sd, f"dec.resblocks.{idx}", x, k, (1, 3, 5))
        xd, y = dev(x), torch.empty(2, C, L, device="cuda")
        ws = torch.empty(rt.lib().svk_resblock1_workspace_bytes(net._handle.ptr, idx, 2, L), dtype=torch.uint8, device="cuda")
        rt.check(rt.lib().svk_resblock1(net._handle.ptr, idx, xd.data_ptr(), 2, L, y.data_ptr(), ws.data_ptr(),
                                        ws.numel(), torch.cuda.current_stream().cuda_stream))
        assert np.abs(_np(y) - ref).max() <= TOL, idx
    z = rng.standard_normal((2, 192, 21)).astype(np.float32)
    assert np.abs(_np(net.dec(dev(z))) - orc.generator(base_sd, base_dims, z)).max() <= TOL


# ------------------------------------------------------------------------ bf16 engine (BASELINE configs[3])
def test_bf16_engine_against_oracle(base_cfg, base_sd):
    """bf16 operands (activation images and weights), one tcgen05 pass, fp32 accumulate, fp32 residual stream.
    The reference has no reduced-precision inference path, so the tolerance is ours to state:
    waveform SNR >= 30 dB against the reference's fp64 output and max-abs <= 0.05 on a signal of amplitude ~0.9;
    everything that is integer / indexing work stays bit-exact."""
    from gpu_util import build_net
    net16 = build_net(base_cfg["model"], base_sd, engine="bf16")
    g = load_golden("infer_base_b2_t40")
    o, mask, (z, z_p, m_p, logs_p) = _run_infer(net16, g)
    assert np.array_equal(_np(mask), g["ref32_x_mask"])
    ref = g["ref64_o"]
    err = _np(o).astype(np.float64) - ref
    snr = 10 * np.log10((ref ** 2).sum() / (err ** 2).sum())
    print("bf16 engine: waveform SNR %.1f dB, max-abs %.3e (|ref| max %.3f); z max-abs %.3e" %
          (snr, np.abs(err).max(), np.abs(ref).max(), np.abs(_np(z) - g["ref64_z"]).max()))
    assert snr >= 30.0 and np.abs(err).max() <= 0.05
    assert np.abs(_np(z) - g["ref64_z"]).max() <= 0.05
    assert net16.last_launch_count() > 100


# ------------------------------------------------------------------------ full-size properties
def test_full_size_window_vs_oracle_and_batch_independence(net, base_sd, base_dims):
    """BASELINE config 3 shape (B=16, T=1024).  The path is convolutional with a receptive field of
    32+64+14 = 110 frames per side, so the oracle run on a cropped window must reproduce the
    interior of the full-size GPU result; and utterances never interact."""
    from gpu_util import dev, inject_eps
    B, T = 16, 1024
    rng = np.random.Generator(np.random.Philox(key=[1024, 16]))
    mel = (rng.standard_normal((B, 80, T)) * 2 - 5).astype(np.float32)
    eps = rng.standard_normal((B, 192, T)).astype(np.float32)
    lengths = np.full(B, T, np.int64)
    with inject_eps(eps), torch.no_grad():
        o = _np(net.infer(dev(mel), dev(lengths, torch.int64), noise_scale=0.667)[0])
    assert o.shape == (B, 1, 256 * T) and np.isfinite(o).all() and np.abs(o).max() <= 1.0
    # (1) oracle on a window of utterance 5: frames [400, 640), compare interior [512, 528)
    b, lo, hi, a0, a1 = 5, 400, 640, 512, 528
    ro, _, _ = Oracle(np.float32).infer(base_sd, base_dims, mel[b:b + 1, :, lo:hi], np.array([hi - lo]),
                                        eps[b:b + 1, :, lo:hi], 0.667, None)
    ref = ro[0, 0, (a0 - lo) * 256:(a1 - lo) * 256]
    assert np.abs(o[b, 0, a0 * 256:a1 * 256] - ref).max() <= TOL
    # (2) batch independence: utterance 11 alone == utterance 11 inside the batch, bit for bit
    with inject_eps(eps[11:12]), torch.no_grad():
        o1 = _np(net.infer(dev(mel[11:12]), dev(lengths[11:12], torch.int64), noise_scale=0.667)[0])
    assert np.array_equal(o1[0], o[11])


def test_config2_resblocks_at_b1_t1024(net, base_sd, base_dims):
    """BASELINE configs[1]: batch 1, 80x1024 mel, decoder ResBlock convs only.  Stage-1 blocks (C=128 at
    64 samples per frame -> L = 65536) through svk_resblock1; a ResBlock1 is local (halo 12 / 36 / 60 samples for
    k = 3 / 7 / 11), so the fp64 oracle on a cropped window must reproduce the interior."""
    import svk_runtime as rt
    from gpu_util import dev
    rng = np.random.Generator(np.random.Philox(key=[2, 1024]))
    C, L = 128, 64 * 1024
    x = rng.standard_normal((1, C, L)).astype(np.float32)
    orc = Oracle(np.float64)
    for idx in (3, 4, 5):
        k = base_dims.resblock_kernel_sizes[idx % 3]
        xd, y = dev(x), torch.empty(1, C, L, device="cuda")
        ws = torch.empty(rt.lib().svk_resblock1_workspace_bytes(net._handle.ptr, idx, 1, L), dtype=torch.uint8, device="cuda")
        rt.check(rt.lib().svk_resblock1(net._handle.ptr, idx, xd.data_ptr(), 1, L, y.data_ptr(), ws.data_ptr(), ws.numel(),
                                        torch.cuda.current_stream().cuda_stream))
        yh = _np(y)
        assert np.isfinite(yh).all()
        for w0 in (0, 30000, L - 400):  # left edge (zero padding), interior across tile borders, right edge
            lo, hi = max(0, w0 - 64), min(L, w0 + 400 + 64)
            ref = orc.resblock1(base_sd, f"dec.resblocks.{idx}", x[:, :, lo:hi], k, (1, 3, 5))
            a0 = w0 if lo == 0 else w0  # interior [w0, w0+400) is >= 60 samples from any artificial cut
            a1 = min(L, w0 + 400)
            assert np.abs(yh[0, :, a0:a1] - ref[0, :, a0 - lo:a1 - lo]).max() <= TOL, (idx, w0)


def test_config4_bf16_at_b64_t512(base_cfg, base_sd):
    """BASELINE configs[3] shape: batch 64, 80x512 mel, bf16 operands with fp32 accumulate.  No reference
    counterpart exists; the fp32-class engine (itself within 1e-4 of the reference) is the yardstick: per-utterance
    waveform SNR >= 30 dB, and utterances stay independent (batch item alone == inside the batch, bit for bit)."""
    from gpu_util import build_net, dev, inject_eps
    B, T = 64, 512
    rng = np.random.Generator(np.random.Philox(key=[64, 512]))
    mel = (rng.standard_normal((B, 80, T)) * 2 - 5).astype(np.float32)
    eps = rng.standard_normal((B, 192, T)).astype(np.float32)
    lengths = np.full(B, T, np.int64)
    outs = {}
    for eng in ("bf16", "tc"):
        net_e = build_net(base_cfg["model"], base_sd, engine=eng)
        with inject_eps(eps), torch.no_grad():
            outs[eng] = _np(net_e.infer(dev(mel), dev(lengths, torch.int64), noise_scale=0.667)[0]).astype(np.float64)
        if eng == "bf16":
            with inject_eps(eps[40:41]), torch.no_grad():
                alone = _np(net_e.infer(dev(mel[40:41]), dev(lengths[40:41], torch.int64), noise_scale=0.667)[0])
            assert np.array_equal(alone[0].astype(np.float64), outs[eng][40])
        del net_e
        torch.cuda.empty_cache()
    assert outs["bf16"].shape == (B, 1, 256 * T) and np.isfinite(outs["bf16"]).all()
    err = outs["bf16"] - outs["tc"]
    snr = 10 * np.log10((outs["tc"] ** 2).sum(axis=(1, 2)) / (err ** 2).sum(axis=(1, 2)))
    print("bf16 vs tc at 64x512: SNR min %.1f dB, median %.1f dB" % (snr.min(), np.median(snr)))
    assert snr.min() >= 30.0


def test_config5_shard_at_b64_t1024(net, base_sd, base_dims):
    """BASELINE configs[4] per-GPU shape: 64 utterances x 1024 frames (16 GB workspace; 2^31-element tensors in the
    narrow stages).  Oracle on cropped windows: the interior of a full-length utterance late in the batch, and a window
    straddling the end of a shorter (padded) one, where the unmasked decoder runs on into the padding."""
    from gpu_util import dev, inject_eps
    if net.engine != "tc":
        pytest.skip("full-size shape: default engine only (the FFMA engine takes 9x longer)")
    B, T = 64, 1024
    rng = np.random.Generator(np.random.Philox(key=[1024, 64]))
    mel = (rng.standard_normal((B, 80, T)) * 2 - 5).astype(np.float32)
    eps = rng.standard_normal((B, 192, T)).astype(np.float32)
    lengths = np.full(B, T, np.int64)
    lengths[9] = 700
    with inject_eps(eps), torch.no_grad():
        o_d, mask_d, _ = net.infer(dev(mel), dev(lengths, torch.int64), noise_scale=0.667)
    assert tuple(o_d.shape) == (B, 1, 256 * T) and bool(torch.isfinite(o_d).all())
    assert np.array_equal(_np(mask_d), Oracle.sequence_mask(lengths, T).astype(np.float32))
    for b, lo, hi, a0, a1 in ((61, 400, 640, 512, 528), (9, 580, 830, 690, 706)):
        ln = np.array([min(max(int(lengths[b]) - lo, 0), hi - lo)], np.int64)
        ro, _, _ = Oracle(np.float32).infer(base_sd, base_dims, mel[b:b + 1, :, lo:hi], ln, eps[b:b + 1, :, lo:hi], 0.667, None)
        ref = ro[0, 0, (a0 - lo) * 256:(a1 - lo) * 256]
        got = _np(o_d[b, 0, a0 * 256:a1 * 256])
        assert np.abs(got - ref).max() <= TOL, b
    # last utterance alone == inside the batch, bit for bit (no cross-utterance term, no 32-bit offset wrap)
    with inject_eps(eps[63:64]), torch.no_grad():
        o1 = net.infer(dev(mel[63:64]), dev(lengths[63:64], torch.int64), noise_scale=0.667)[0]
    assert torch.equal(o1[0], o_d[63])


# ------------------------------------------------------------------------ resblock: "2" (modules.ResBlock2, models.py:121)
@pytest.mark.parametrize("engine", ["tc", "fp32", "bf16"])
def test_resblock2_model_matches_reference_golden(engine):
    """Generator built from ResBlock2 (two convs per block, dilations (1,3)/(2,5)/(3,8)): whole infer against the reference's
    fp64 golden, and one block per kernel size through svk_resblock1 (which runs whichever block type the handle holds)."""
    import svk_runtime as rt
    from gpu_util import build_net, dev
    from test_oracle import resblock2_case
    model, dims, sd, g = resblock2_case()
    net = build_net(model, sd, engine=engine)
    tol = TOL if engine != "bf16" else 0.05 * float(np.abs(g["ref64_o"]).max())
    o, mask, (z, *_rest) = _run_infer(net, g)
    assert np.array_equal(_np(mask), g["ref32_x_mask"])
    assert np.abs(_np(o) - g["ref64_o"]).max() <= tol
    if engine != "bf16":
        assert np.abs(_np(z) - g["ref64_z"]).max() <= TOL
    x = g["rb_in"]
    B, C, L = x.shape
    for j in range(3):
        xd, y = dev(x), torch.empty(B, C, L, device="cuda")
        ws = torch.empty(rt.lib().svk_resblock1_workspace_bytes(net._handle.ptr, j, B, L), dtype=torch.uint8, device="cuda")
        rt.check(rt.lib().svk_resblock1(net._handle.ptr, j, xd.data_ptr(), B, L, y.data_ptr(), ws.data_ptr(), ws.numel(),
                                        torch.cuda.current_stream().cuda_stream))
        err = np.abs(_np(y) - g[f"rb{j}"]).max()
        assert err <= (TOL if engine != "bf16" else 0.05 * float(np.abs(g[f"rb{j}"]).max())), (j, err)
