"""Golden vectors for the analysis direction: PosteriorEncoder + flow forward of the UNMODIFIED
reference, run in the build container on the seeded weights of make_golden.py.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_posterior.py   ->  tests/golden/posterior_base_b2_t24.npz

Reference call sites: models.py:103-110 (PosteriorEncoder.forward), :73-76 (ResidualCouplingBlock.forward, reverse=False),
:321-323 (how SynthesizerTrn.forward chains them).  eps (the randn_like draw of models.py:109) is injected.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (sets up sys.path, imports the reference modules)

import torch  # noqa: E402


def main():
    cfg = json.load(open(os.path.join(G.ROOT, "configs", "iitp_base.json")))
    model = cfg["model"]
    dims = G.W.dims_from_model_kwargs(513, **model)
    sd = G.W.make_state_dict(dims, seed=1234)
    B, T = 2, 24
    rng = G.philox(21, "posterior")
    spec = np.abs(rng.standard_normal((B, 513, T)) * 3.0).astype(np.float32) + 1e-3  # linear-spectrogram-like, >= 1e-3
    eps = rng.standard_normal((B, 192, T)).astype(np.float32)
    lengths = np.array([24, 17], np.int64)
    out = {"spec": spec, "eps": eps, "lengths": lengths}
    for dt, tag in ((torch.float32, "ref32"), (torch.float64, "ref64")):
        net = G.build_ref(model, sd, dtype=dt)
        with G.inject_eps(torch.from_numpy(eps)):
            z, m_q, logs_q, y_mask = net.enc_q(torch.from_numpy(spec).to(dt), torch.from_numpy(lengths), g=None)
        z_p = net.flow(z, y_mask, g=None)
        back = net.flow(z_p, y_mask, g=None, reverse=True)
        for k, v in (("z", z), ("m_q", m_q), ("logs_q", logs_q), ("y_mask", y_mask), ("z_p", z_p)):
            out[f"{tag}_{k}"] = v.numpy()
        print(tag, "z std", float(z.std()), "z_p std", float(z_p.std()), "logs range", float(logs_q.min()), float(logs_q.max()),
              "reverse(forward(z)) err", float((back - z).abs().max()))
    for k in ("z", "m_q", "logs_q", "z_p"):
        print(k, "ref32 vs ref64", float(np.abs(out[f"ref32_{k}"] - out[f"ref64_{k}"]).max()))
    np.savez_compressed(os.path.join(HERE, "posterior_base_b2_t24.npz"), **out)
    print("wrote", os.path.getsize(os.path.join(HERE, "posterior_base_b2_t24.npz")), "bytes")


if __name__ == "__main__":
    main()
