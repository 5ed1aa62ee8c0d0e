"""Generate golden vectors by running the UNMODIFIED reference in the build container.

Usage (build container only; /root/reference does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so parity is pinned on outputs of
the reference's own modules run here on seeded inputs:

  * weights   : svk_weights.make_state_dict(dims, seed) loaded with strict=True into the reference
                ``SynthesizerTrn`` (proves the key/shape surface); only the checksum is stored.
  * inputs    : numpy Philox streams (mel ~ N(-5, 2), eps ~ N(0,1)); ``torch.randn_like`` is patched
                during the reference call so the reference consumes our ``eps``.
  * outputs   : reference fp32 AND fp64 (``.double()``) results of infer / sub-modules.

Nothing from the reference is copied; it is imported from where it lies.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "smart-vocoder_b200"))
sys.path.insert(0, ROOT)
REF = os.environ.get("SVK_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

import torch  # noqa: E402

import svk_weights as W  # noqa: E402

import models as ref_models  # noqa: E402  (reference)
import modules as ref_modules  # noqa: E402  (reference)
import transforms as ref_transforms  # noqa: E402  (reference)
import commons as ref_commons  # noqa: E402  (reference)

torch.set_grad_enabled(False)


def philox(seed, tag):
    import zlib
    return np.random.Generator(np.random.Philox(key=[seed, zlib.crc32(tag.encode())]))


def make_inputs(seed, B, T, lengths):
    mel = (philox(seed, "mel").standard_normal((B, 80, T)) * 2.0 - 5.0).astype(np.float32)
    eps = philox(seed, "eps").standard_normal((B, 192, T)).astype(np.float32)
    return mel, np.asarray(lengths, np.int64), eps


def build_ref(model_kwargs, sd_np, spec_channels=513, dtype=torch.float32):
    net = ref_models.SynthesizerTrn(spec_channels, 32, n_speakers=109, **model_kwargs)
    net.eval()
    missing = net.load_state_dict({k: torch.from_numpy(v) for k, v in sd_np.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return net.to(dtype)


class inject_eps:
    def __init__(self, eps):
        self.eps = eps

    def __enter__(self):
        self.orig = torch.randn_like
        torch.randn_like = lambda t, *a, **k: self.eps.to(t.dtype)
        return self

    def __exit__(self, *a):
        torch.randn_like = self.orig


def run_infer(net, mel, lengths, eps, noise_scale, max_len, dtype):
    with inject_eps(torch.from_numpy(eps)):
        o, mask, (z, z_p, m_p, logs_p) = net.infer(torch.from_numpy(mel).to(dtype), torch.from_numpy(lengths),
                                                   noise_scale=noise_scale, max_len=max_len)
    return dict(o=o, x_mask=mask, z=z, z_p=z_p, m_p=m_p, logs_p=logs_p)


def infer_case(name, model_kwargs, sd, seed, B, T, lengths, noise_scale, max_len, with_trace=False):
    mel, lengths, eps = make_inputs(seed, B, T, lengths)
    eps = eps[:, :model_kwargs["inter_channels"]]
    out = {"mel": mel, "lengths": lengths, "eps": eps, "noise_scale": np.float64(noise_scale),
           "max_len": np.int64(-1 if max_len is None else max_len)}
    net32 = build_ref(model_kwargs, sd, dtype=torch.float32)
    r32 = run_infer(net32, mel, lengths, eps, noise_scale, max_len, torch.float32)
    net64 = build_ref(model_kwargs, sd, dtype=torch.float64)
    r64 = run_infer(net64, mel, lengths, eps, noise_scale, max_len, torch.float64)
    for k, v in r32.items():
        out[f"ref32_{k}"] = v.numpy()
    for k, v in r64.items():
        out[f"ref64_{k}"] = v.numpy().astype(np.float64 if k == "o" else np.float32)
    err = (r32["o"].double() - r64["o"]).abs().max().item()
    print(f"[{name}] |o|max={r64['o'].abs().max():.3f} std={r64['o'].std():.3f} z std={r64['z'].std():.3f} "
          f"ref fp32-vs-fp64 max-abs on o: {err:.2e}")
    if with_trace:
        # module-level intermediates (fp64 truth stored as fp32 is enough to localise a bug)
        mask = r64["x_mask"]
        z = r64["z_p"]
        flows = list(net64.flow.flows)
        n = 0
        for fl in reversed(flows):
            z = fl(z, mask, g=None, reverse=True)
            if isinstance(fl, ref_modules.ResidualCouplingLayer):
                out[f"trace_flow{n}"] = z.numpy().astype(np.float32)
                n += 1
        dec = net64.dec
        x = dec.conv_pre((r64["z"] * mask)[:, :, :max_len])
        out["trace_conv_pre"] = x.numpy().astype(np.float32)
        for i in range(dec.num_upsamples):
            x = torch.nn.functional.leaky_relu(x, 0.1)
            x = dec.ups[i](x)
            out[f"trace_ups{i}"] = x.numpy().astype(np.float32)
            xs = None
            for j in range(dec.num_kernels):
                r = dec.resblocks[i * dec.num_kernels + j](x)
                if i == 0:
                    out[f"trace_rb{j}"] = r.numpy().astype(np.float32)
                xs = r if xs is None else xs + r
            x = xs / dec.num_kernels
            out[f"trace_stage{i}"] = x.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    return out


def spline_case(name, seed):
    """transforms.piecewise_rational_quadratic_transform fwd+inverse, incl. tails and exact knots."""
    rng = philox(seed, "spline")
    B, C, T, nb = 2, 3, 40, 10
    x = rng.uniform(-7.0, 7.0, size=(B, C, T)).astype(np.float32)
    x[0, 0, :6] = [-5.0, 5.0, 0.0, -4.9999995, 4.9999995, 5.0000005]
    uw = (rng.standard_normal((B, C, T, nb)) * 1.5).astype(np.float32)
    uh = (rng.standard_normal((B, C, T, nb)) * 1.5).astype(np.float32)
    ud = (rng.standard_normal((B, C, T, nb - 1)) * 2.0).astype(np.float32)
    ud[1, 2, :3, :] = 25.0  # softplus threshold branch
    out = dict(x=x, uw=uw, uh=uh, ud=ud)
    for inv in (False, True):
        for dt, tag in ((torch.float32, "32"), (torch.float64, "64")):
            y, lad = ref_transforms.piecewise_rational_quadratic_transform(
                torch.from_numpy(x).to(dt), torch.from_numpy(uw).to(dt), torch.from_numpy(uh).to(dt),
                torch.from_numpy(ud).to(dt), inverse=inv, tails="linear", tail_bound=5.0)
            out[f"y{tag}_inv{int(inv)}"] = y.numpy()
            out[f"lad{tag}_inv{int(inv)}"] = lad.numpy()
    # round trip in fp64 through the reference
    y = torch.from_numpy(out["y64_inv0"])
    xr, _ = ref_transforms.piecewise_rational_quadratic_transform(
        y, torch.from_numpy(uw).double(), torch.from_numpy(uh).double(), torch.from_numpy(ud).double(),
        inverse=True, tails="linear", tail_bound=5.0)
    print(f"[{name}] reference fp64 round-trip error {np.abs(xr.numpy() - x).max():.2e}")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def convflow_case(name, seed):
    """modules.ConvFlow (never instantiated by the reference model) fwd + reverse."""
    torch.manual_seed(seed)
    cf = ref_modules.ConvFlow(4, 32, 3, 3).eval()
    rng = philox(seed, "convflow")
    sd = {}
    for k, v in cf.state_dict().items():
        a = (rng.standard_normal(tuple(v.shape)) * (0.3 if "proj" in k else 0.5)).astype(np.float32)
        if "gamma" in k:
            a = 1.0 + 0.1 * a
        sd[k] = a
    cf.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    B, T = 2, 37
    x = (rng.standard_normal((B, 4, T)) * 2.5).astype(np.float32)
    lengths = np.array([37, 20], np.int64)
    mask = ref_commons.sequence_mask(torch.from_numpy(lengths), T)[:, None, :].float()
    out = {"x": x, "lengths": lengths}
    out.update({"w_" + k: v for k, v in sd.items()})
    for dt, tag in ((torch.float32, "32"), (torch.float64, "64")):
        m = cf.to(dt)
        y, logdet = m(torch.from_numpy(x).to(dt), mask.to(dt), reverse=False)
        xr = m(y, mask.to(dt), reverse=True)
        yr = m(torch.from_numpy(x).to(dt), mask.to(dt), reverse=True)
        out[f"fwd{tag}"] = y.numpy()
        out[f"logdet{tag}"] = logdet.numpy()
        out[f"rev{tag}"] = yr.numpy()
        print(f"[{name}] {tag}: reverse(forward(x)) max err {(xr - torch.from_numpy(x).to(dt) * mask.to(dt)).abs().max():.2e}")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def main():
    cfg = json.load(open(os.path.join(ROOT, "configs", "iitp_base.json")))
    model = cfg["model"]
    dims = W.dims_from_model_kwargs(513, **model)
    sd = W.make_state_dict(dims, seed=1234)
    meta = {"iitp_base_seed1234_checksum": W.state_dict_checksum(sd)}

    # case A: padded batch with unequal lengths, T not a multiple of any tile
    infer_case("infer_base_b2_t40", model, sd, seed=7, B=2, T=40, lengths=[40, 29], noise_scale=0.667, max_len=None)
    # case B: max_len < T, single utterance, with per-module trace
    infer_case("infer_base_b1_t12_maxlen9", model, sd, seed=8, B=1, T=12, lengths=[12], noise_scale=0.667,
               max_len=9, with_trace=True)
    # case C: T shorter than every halo, noise_scale=1
    infer_case("infer_base_b3_t3", model, sd, seed=9, B=3, T=3, lengths=[3, 1, 2], noise_scale=1.0, max_len=None)

    # tiny topology: weights are small enough to be regenerated instantly; full trace
    tdims = W.dims_from_model_kwargs(513, **W.TINY_MODEL)
    tsd = W.make_state_dict(tdims, seed=4321)
    meta["tiny_seed4321_checksum"] = W.state_dict_checksum(tsd)
    infer_case("infer_tiny_b2_t33", W.TINY_MODEL, tsd, seed=11, B=2, T=33, lengths=[33, 17], noise_scale=0.667,
               max_len=None, with_trace=True)

    spline_case("rq_spline", seed=21)
    convflow_case("convflow", seed=22)
    json.dump(meta, open(os.path.join(HERE, "meta.json"), "w"), indent=1)
    print(meta)


if __name__ == "__main__":
    main()
