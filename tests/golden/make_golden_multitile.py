"""One more golden from the UNMODIFIED reference: a ragged batch that spans several 128-frame tiles.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_multitile.py      (build container only)

The cases of make_golden.py have T <= 40 -- one time tile of the tensor-core kernels.  This one is B=3,
T=300 with lengths (300, 257, 129): three tiles per utterance in the encoder / flow (tile borders at 128 and
256, one length one past a border, one exactly one past the first tile), 38..600 tiles in the decoder stages,
padded rows.  Stored: inputs, the reference's fp64 waveform (float64), fp64 latents z / m_p as
float32, the fp32 mask, and the reference's own fp32-vs-fp64 distance for the report.
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


def main():
    cfg = json.load(open(os.path.join(G.ROOT, "configs", "iitp_base.json")))
    model = cfg["model"]
    dims = W.dims_from_model_kwargs(513, **model)
    sd = W.make_state_dict(dims, seed=1234)
    B, T, lengths, ns = 3, 300, [300, 257, 129], 0.667
    mel, lengths, eps = G.make_inputs(13, B, T, lengths)
    net32 = G.build_ref(model, sd, dtype=torch.float32)
    r32 = G.run_infer(net32, mel, lengths, eps, ns, None, torch.float32)
    net64 = G.build_ref(model, sd, dtype=torch.float64)
    r64 = G.run_infer(net64, mel, lengths, eps, ns, None, torch.float64)
    err32 = (r32["o"].double() - r64["o"]).abs().max().item()
    out = {"mel": mel, "lengths": lengths, "eps": eps, "noise_scale": np.float64(ns), "max_len": np.int64(-1),
           "ref64_o": r64["o"].numpy(), "ref64_z": r64["z"].numpy().astype(np.float32),
           "ref64_m_p": r64["m_p"].numpy().astype(np.float32), "ref32_x_mask": r32["x_mask"].numpy(),
           "ref32_vs_ref64_o": np.float64(err32)}
    print(f"|o|max={r64['o'].abs().max():.3f}  reference fp32-vs-fp64 max-abs on o: {err32:.2e}")
    np.savez_compressed(os.path.join(HERE, "infer_base_b3_t300_ragged.npz"), **out)


if __name__ == "__main__":
    main()
