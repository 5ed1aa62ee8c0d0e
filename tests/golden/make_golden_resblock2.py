"""Golden for the ``resblock: "2"`` decoder variant (modules.ResBlock2, modules.py:232-256), from the UNMODIFIED reference.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_resblock2.py      (build container only)

The base model with Generator built from ResBlock2 (models.py:121): two convs per block, dilations
(1,3) / (2,5) / (3,8) for the kernel sizes 3 / 7 / 11, B=2, T=150 with lengths (150, 97) -- two encoder tiles,
decoder stages of 10..300 tiles.  Stored: inputs, the reference's fp64 waveform, fp64 latent z as fp32, one
ResBlock2 output per kernel size of the first stage (fp64 truth as fp32) with its input, fp32-vs-fp64 distance.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (sets sys.path for the reference and the package)

import torch  # noqa: E402

import svk_weights as W  # noqa: E402

DILATIONS = [[1, 3], [2, 5], [3, 8]]


def model_kwargs():
    cfg = json.load(open(os.path.join(G.ROOT, "configs", "iitp_base.json")))
    model = dict(cfg["model"])
    model["resblock"] = "2"
    model["resblock_dilation_sizes"] = DILATIONS
    return model


def main():
    model = model_kwargs()
    dims = W.dims_from_model_kwargs(513, **model)
    sd = W.make_state_dict(dims, seed=2222)
    B, T, lengths, ns = 2, 150, [150, 97], 0.667
    mel, lengths, eps = G.make_inputs(17, B, T, lengths)
    net32 = G.build_ref(model, sd, dtype=torch.float32)
    r32 = G.run_infer(net32, mel, lengths, eps, ns, None, torch.float32)
    net64 = G.build_ref(model, sd, dtype=torch.float64)
    r64 = G.run_infer(net64, mel, lengths, eps, ns, None, torch.float64)
    err32 = (r32["o"].double() - r64["o"]).abs().max().item()
    out = {"mel": mel, "lengths": lengths, "eps": eps, "noise_scale": np.float64(ns), "max_len": np.int64(-1),
           "ref64_o": r64["o"].numpy(), "ref64_z": r64["z"].numpy().astype(np.float32),
           "ref32_x_mask": r32["x_mask"].numpy(), "ref32_vs_ref64_o": np.float64(err32),
           "checksum": np.array(W.state_dict_checksum(sd))}
    dec = net64.dec
    x = dec.ups[0](torch.nn.functional.leaky_relu(dec.conv_pre(r64["z"] * r64["x_mask"]), 0.1))
    x = x[:1, :, 100:356].contiguous()  # a 256-step crop keeps the fixture small; the block sees it as a whole signal
    out["rb_in"] = x.numpy().astype(np.float32)
    for j in range(dec.num_kernels):
        out[f"rb{j}"] = dec.resblocks[j](x.float().double()).numpy().astype(np.float32)
    print(f"|o|max={r64['o'].abs().max():.3f} std={r64['o'].std():.3f}  reference fp32-vs-fp64 max-abs on o: {err32:.2e}")
    np.savez_compressed(os.path.join(HERE, "infer_resblock2_b2_t150.npz"), **out)


if __name__ == "__main__":
    main()
