"""Thin torch-tensor wrappers over the stateless C-ABI operators (test helpers only)."""
import contextlib

import numpy as np
import torch

import svk_runtime as rt


def _s():
    return torch.cuda.current_stream().cuda_stream


def dev(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype)


def conv1d(x, w, b, dilation=1, padding=0, pre_slope=1.0, engine="fp32"):
    B, Cin, L = x.shape
    Cout, _, k = w.shape
    y = torch.empty(B, Cout, L + 2 * padding - dilation * (k - 1), device="cuda")
    fn = rt.lib().svk_conv1d_tc if engine == "tc" else rt.lib().svk_conv1d
    rt.check(fn(x.data_ptr(), B, Cin, L, w.data_ptr(), None if b is None else b.data_ptr(), Cout, k,
                                 dilation, padding, pre_slope, y.data_ptr(), _s()))
    return y


def conv_transpose1d(x, w, b, stride, padding, pre_slope=1.0, engine="fp32"):
    B, Cin, L = x.shape
    _, Cout, k = w.shape
    y = torch.empty(B, Cout, (L - 1) * stride - 2 * padding + k, device="cuda")
    fn = rt.lib().svk_conv_transpose1d_tc if engine == "tc" else rt.lib().svk_conv_transpose1d
    rt.check(fn(x.data_ptr(), B, Cin, L, w.data_ptr(), None if b is None else b.data_ptr(),
                                           Cout, k, stride, padding, pre_slope, y.data_ptr(), _s()))
    return y


def sequence_mask(lengths, T):
    B = lengths.shape[0]
    m = torch.empty(B, T, device="cuda")
    rt.check(rt.lib().svk_sequence_mask(lengths.data_ptr(), B, T, m.data_ptr(), _s()))
    return m


def flip(x):
    B, C, T = x.shape
    y = torch.empty_like(x)
    rt.check(rt.lib().svk_flip(x.data_ptr(), B, C, T, y.data_ptr(), _s()))
    return y


def weight_norm(v, g):
    w = torch.empty_like(v)
    rt.check(rt.lib().svk_weight_norm(v.data_ptr(), g.data_ptr(), v.shape[0], v.numel() // v.shape[0], w.data_ptr(), _s()))
    return w


def rq_spline(x, uw, uh, ud, inverse, tail_bound=5.0):
    y, lad = torch.empty_like(x), torch.empty_like(x)
    bins = torch.empty(x.shape, device="cuda", dtype=torch.int32)
    rt.check(rt.lib().svk_rq_spline(x.data_ptr(), uw.data_ptr(), uh.data_ptr(), ud.data_ptr(), x.numel(), uw.shape[-1],
                                    int(inverse), tail_bound, 1e-3, 1e-3, 1e-3, y.data_ptr(), lad.data_ptr(),
                                    bins.data_ptr(), _s()))
    return y, lad, bins


def convflow(x, mask, w, filter_channels, kernel, n_layers, num_bins=10, tail_bound=5.0, reverse=False):
    """svk_convflow with the state_dict of a reference modules.ConvFlow (numpy arrays keyed like the module)."""
    import ctypes
    B, C, T = x.shape
    st = lambda name, leaf: dev(np.stack([w[f"convs.{name}.{i}.{leaf}"] for i in range(n_layers)]))  # noqa: E731
    t = dict(pre_w=dev(w["pre.weight"]), pre_b=dev(w["pre.bias"]), sep_w=st("convs_sep", "weight"), sep_b=st("convs_sep", "bias"),
             pw_w=st("convs_1x1", "weight"), pw_b=st("convs_1x1", "bias"), norm1_g=st("norms_1", "gamma"), norm1_b=st("norms_1", "beta"),
             norm2_g=st("norms_2", "gamma"), norm2_b=st("norms_2", "beta"), proj_w=dev(w["proj.weight"]), proj_b=dev(w["proj.bias"]))
    ws_ = rt.SvkConvFlowWeights(**{k: v.data_ptr() for k, v in t.items()})
    y = torch.empty_like(x)
    logdet = torch.empty(B, device="cuda")
    bins = torch.empty(B, C // 2, T, device="cuda", dtype=torch.int32)
    nbytes = rt.lib().svk_convflow_workspace_bytes(B, C, T, filter_channels, num_bins)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    rt.check(rt.lib().svk_convflow(x.data_ptr(), mask.data_ptr(), B, C, T, filter_channels, kernel, n_layers, num_bins, tail_bound,
                                   ctypes.byref(ws_), int(reverse), y.data_ptr(), logdet.data_ptr(), bins.data_ptr(), ws.data_ptr(),
                                   nbytes, _s()))
    torch.cuda.synchronize()
    return y, logdet, bins


@contextlib.contextmanager
def inject_eps(eps_np):
    """Feed the path's only RNG draw (models.py:336) with a fixed tensor, as make_golden.py does."""
    orig = torch.randn_like
    torch.randn_like = lambda t, *a, **k: torch.from_numpy(eps_np).to(t.device, t.dtype)
    try:
        yield
    finally:
        torch.randn_like = orig


def build_net(cfg_model, sd_np, **extra):
    from models import SynthesizerTrn
    net = SynthesizerTrn(513, 32, n_speakers=109, **cfg_model, **extra)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd_np.items()})
    return net.cuda().eval()
