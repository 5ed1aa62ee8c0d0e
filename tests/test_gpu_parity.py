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


def test_flow_passthrough_is_bit_exact(net, base_sd, base_dims):
    """The last coupling (RCL0) leaves its x0 half untouched and Flip/split/cat are pure indexing, so
    z[:, :96] must equal -- bit for bit -- what the oracle carries there, given the same input to RCL0."""
    from gpu_util import dev
    g = load_golden("infer_base_b1_t12_maxlen9")
    mask = g["ref32_x_mask"]
    z_in = g["trace_flow2"]  # fp32 copy of the state after RCL1
    # run only the tail [Flip, RCL0] through the oracle in fp32
    orc = Oracle(np.float32)
    ref = orc.coupling_reverse(base_sd, "flow.flows.0", base_dims, orc.flip(z_in), mask)
    # and the full reverse flow on the GPU from z_p; compare the untouched half of the final result
    z = _np(net.flow(dev(g["ref64_z_p"]), dev(mask), reverse=True))
    assert ref[:, :96].tobytes() == orc.flip(z_in)[:, :96].tobytes()  # oracle: x0 passes through
    assert np.abs(z[:, 96:] - ref[:, 96:]).max() <= TOL


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
