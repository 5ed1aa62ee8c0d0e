"""Minimal target for ncu: N passes of SynthesizerTrn.infer at a bench shape (no timing, no CPU work)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "smart-vocoder_b200"))
import torch  # noqa: E402

import svk_weights as W  # noqa: E402
from models import SynthesizerTrn  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--frames", type=int, default=1024)
ap.add_argument("--iters", type=int, default=1)
a = ap.parse_args()
cfg = json.load(open(os.path.join(ROOT, "configs", "iitp_base.json")))
dims = W.dims_from_model_kwargs(513, **cfg["model"])
net = SynthesizerTrn(513, 32, n_speakers=109, **cfg["model"])
net.load_state_dict({k: torch.from_numpy(v) for k, v in W.make_state_dict(dims, seed=1234).items()})
net = net.cuda().eval()
g = torch.Generator().manual_seed(0)
mel = (torch.randn(a.batch, 80, a.frames, generator=g) * 2 - 5).cuda()
lengths = torch.full((a.batch,), a.frames, dtype=torch.int64).cuda()
torch.manual_seed(1)
for _ in range(a.iters):
    o = net.infer(mel, lengths, noise_scale=0.667)[0]
torch.cuda.synchronize()
print("ok", tuple(o.shape), net.last_launch_count(), "launches/iter")
