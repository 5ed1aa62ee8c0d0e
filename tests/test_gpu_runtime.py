"""Runtime around svk_infer (C ABI: svk_graph_*, svk_pipeline_*, svk_randn): CUDA-graph replay and the pipelined
host-buffer entry must return exactly what svk_infer / svk_infer_host return; the device N(0,1) generator is checked
for its statistics and its counter semantics."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def net(base_cfg, base_sd):
    from gpu_util import build_net
    return build_net(base_cfg["model"], base_sd, engine="tc")


def _np(t):
    return t.detach().cpu().numpy()


def test_infer_graph_replays_infer_bit_for_bit(net):
    from gpu_util import dev, inject_eps
    g = load_golden("infer_base_b2_t40")
    mel, ln = dev(g["mel"]), dev(g["lengths"], torch.int64)
    with inject_eps(g["eps"]), torch.no_grad():
        o, mask, lat = net.infer(mel, ln, noise_scale=float(g["noise_scale"]))
        og, maskg, latg = net.infer_graph(mel, ln, noise_scale=float(g["noise_scale"]))
        assert torch.equal(o, og) and torch.equal(mask, maskg)
        for a, b in zip(lat, latg):
            assert torch.equal(a, b)
        # replay with other inputs (same shape -> same graph), then the first ones again
        keep = og.clone()
        mel2 = mel.flip(0).contiguous()
        o2 = net.infer(mel2, ln, noise_scale=float(g["noise_scale"]))[0]
        o2g = net.infer_graph(mel2, ln, noise_scale=float(g["noise_scale"]))[0]
        assert torch.equal(o2, o2g) and not torch.equal(o2g, keep)
        assert torch.equal(net.infer_graph(mel, ln, noise_scale=float(g["noise_scale"]))[0], keep)
    assert len(net._graphs) == 1
    gr = next(iter(net._graphs.values()))
    assert gr.kernel_nodes > 100
    print("graph: %d kernel nodes, programmatic edges %s" % (gr.kernel_nodes, gr.programmatic_edges))
    assert np.abs(_np(og) - g["ref64_o"]).max() <= 1e-4
    # max_len < T and another noise_scale are other graphs
    gm = load_golden("infer_base_b1_t12_maxlen9")
    with inject_eps(gm["eps"]), torch.no_grad():
        om = net.infer_graph(dev(gm["mel"]), dev(gm["lengths"], torch.int64), noise_scale=float(gm["noise_scale"]), max_len=9)[0]
    assert np.abs(_np(om) - gm["ref64_o"]).max() <= 1e-4
    assert len(net._graphs) == 2


def test_infer_graph_with_multi_layer_wn_launches(base_cfg, base_sd, monkeypatch):
    """The whole-stack WN launch (tiles synchronise through flags that a memset node zeroes before every launch) captured in
    a CUDA graph: replays must keep returning what eager infer returns -- a flag buffer left over from the previous replay
    would let a tile read its neighbour's halo too early."""
    from gpu_util import build_net, dev, inject_eps
    monkeypatch.setenv("SVK_FUSE_WN", "2")   # fused layers whatever the batch size ...
    monkeypatch.setenv("SVK_WN_STACK", "1")  # ... and the whole stack in one launch
    n = build_net(base_cfg["model"], base_sd, engine="tc")
    g = load_golden("infer_base_b3_t300_ragged")
    mel, ln = dev(g["mel"]), dev(g["lengths"], torch.int64)
    with inject_eps(g["eps"]), torch.no_grad():
        o = n.infer(mel, ln, noise_scale=float(g["noise_scale"]))[0]
        launches = n.last_launch_count()
        for _ in range(4):
            assert torch.equal(n.infer_graph(mel, ln, noise_scale=float(g["noise_scale"]))[0], o)
        mel2 = mel.flip(0).contiguous()
        assert torch.equal(n.infer_graph(mel2, ln, noise_scale=float(g["noise_scale"]))[0],
                           n.infer(mel2, ln, noise_scale=float(g["noise_scale"]))[0])
    assert launches < 100  # 5 WN launches instead of 48 (or 96)
    assert np.abs(_np(o) - g["ref64_o"]).max() <= 1e-4


def test_pipeline_matches_infer_host(net, base_dims):
    B, T, n_calls = 2, 70, 5
    rng = np.random.Generator(np.random.Philox(key=[70, 2]))
    mels = [(rng.standard_normal((B, 80, T)) * 2 - 5).astype(np.float32) for _ in range(n_calls)]
    epss = [rng.standard_normal((B, 192, T)).astype(np.float32) for _ in range(n_calls)]
    lens = [np.array([T, 10 + 9 * i], np.int64) for i in range(n_calls)]
    want = [net.infer_host(m, l, e, 0.667)[:2] for m, l, e in zip(mels, lens, epss)]
    pipe = net.pipeline(B, T, depth=2)
    outs = [np.zeros((B, 1, 256 * T), np.float32) for _ in range(n_calls)]
    masks = [np.zeros((B, 1, T), np.float32) for _ in range(n_calls)]
    tickets = []
    for i in range(n_calls):  # submissions run ahead of the waits: slots are reused after an implicit wait
        tickets.append(pipe.submit(mels[i], lens[i], outs[i], eps=epss[i], noise_scale=0.667, x_mask=masks[i]))
        if i >= 1:
            pipe.wait(tickets[i - 1])
    pipe.drain()
    for i in range(n_calls):
        assert np.array_equal(outs[i], want[i][0]), i
        assert np.array_equal(masks[i], want[i][1]), i
    # device-drawn eps: reproducible from (seed, submission order), different across calls and seeds
    def run(seed):
        p = net.pipeline(B, T, depth=3)
        res = [np.zeros((B, 1, 256 * T), np.float32) for _ in range(3)]
        for i in range(3):
            p.submit(mels[0], lens[0], res[i], seed=seed, noise_scale=0.667)
        p.drain()
        p.close()
        return res
    a, b, c = run(11), run(11), run(12)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert not np.array_equal(a[0], a[1]) and not np.array_equal(a[0], c[0])
    assert all(np.isfinite(x).all() and np.abs(x).max() <= 1.0 for x in a)
    import svk_runtime as rt
    with pytest.raises(rt.SvkError):
        pipe.wait(99)
    pipe.close()


def test_device_randn_statistics_and_counter(net):
    import svk_runtime as rt
    n = 1 << 22
    x = torch.empty(n, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    rt.check(rt.lib().svk_randn(net._handle.ptr, 1234, 0, n, x.data_ptr(), s))
    y = torch.empty(n, device="cuda")
    rt.check(rt.lib().svk_randn(net._handle.ptr, 1234, 0, n, y.data_ptr(), s))
    assert torch.equal(x, y)
    xd = x.double()
    m, v = xd.mean().item(), xd.var().item()
    k = ((xd - m) ** 4).mean().item() / v ** 2
    print("svk_randn: mean %.2e var %.5f kurtosis %.4f max|x| %.2f" % (m, v, k, x.abs().max().item()))
    assert abs(m) < 3e-3 and abs(v - 1) < 5e-3 and abs(k - 3) < 3e-2 and torch.isfinite(x).all()
    # P(|x| > 3) = 0.0027 for a normal variable
    assert abs((x.abs() > 3).double().mean().item() - 0.0027) < 3e-4
    # counter semantics: element i = f(seed, offset + i / 4), independent of the call's length
    z = torch.empty(8, device="cuda")
    rt.check(rt.lib().svk_randn(net._handle.ptr, 1234, 5, 8, z.data_ptr(), s))
    assert torch.equal(z, x[20:28])
    w = torch.empty(n, device="cuda")
    rt.check(rt.lib().svk_randn(net._handle.ptr, 1235, 0, n, w.data_ptr(), s))
    assert abs((x * w).double().mean().item()) < 3e-3  # different seeds: uncorrelated streams


def test_device_int16_egress_matches_host_conversion():
    import svk_wav
    g = torch.Generator(device="cuda").manual_seed(4)
    x = (torch.randn(3, 1, 256 * 37 + 5, device="cuda", generator=g) * 0.6)
    x[0, 0, :4] = torch.tensor([1.0, -1.0, 0.99999, 3.0517578125e-05 * 0.5])
    q = svk_wav.to_int16(x)
    assert q.dtype == torch.int16 and q.is_cuda and q.shape == x.shape
    assert np.array_equal(q.cpu().numpy(), svk_wav.to_int16(x.cpu().numpy()))
