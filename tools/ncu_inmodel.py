"""Ordinal of an in-model launch among the launches of one kernel, for `ncu -k regex:<kernel> -s <ordinal> -c 1`.

    python tools/ncu_inmodel.py --layer resblock_conv2 --cin 128 --k 11 [--nth 0] [--kernel conv_tc_kernel]

Runs one profiled SynthesizerTrn.infer (svk_profile_begin/end records are in launch order) and prints how many
launches of the same kernel precede the wanted one in an infer call.  conv_tc_kernel launches = tensor-engine records
that are not fused pairs / fused WN layers; the FFMA, pair and WN-layer kernels are separate kernels.
"""
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

KERNEL_OF = {"resblock_pair": "conv_tc_pair_kernel", "wn_layer": "wn_layer_kernel", "split_image": "split_image_kernel"}

ap = argparse.ArgumentParser()
ap.add_argument("--layer", required=True)
ap.add_argument("--cin", type=int, required=True)
ap.add_argument("--k", type=int, required=True)
ap.add_argument("--nth", type=int, default=0)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--frames", type=int, default=1024)
a = ap.parse_args()
cfg = json.load(open(os.path.join(ROOT, "configs", "iitp_base.json")))
dims = W.dims_from_model_kwargs(513, **cfg["model"])
net = SynthesizerTrn(513, 32, n_speakers=109, range_check=False, **cfg["model"])
net.load_state_dict({k: torch.from_numpy(v) for k, v in W.make_state_dict(dims, seed=1234).items()})
net = net.cuda().eval()
mel = (torch.randn(a.batch, 80, a.frames) * 2 - 5).cuda()
lengths = torch.full((a.batch,), a.frames, dtype=torch.int64).cuda()
net._handle.profile_begin(4096)
net.infer(mel, lengths, noise_scale=0.667)
torch.cuda.synchronize()
recs = net._handle.profile_end()


def kernel_of(r):
    if r["layer"] in KERNEL_OF:
        return KERNEL_OF[r["layer"]]
    return "conv_tc_kernel" if r["engine"] == "tc" else "ffma"


want_kernel, seen, hits = None, {}, 0
for r in recs:
    kn = kernel_of(r)
    if r["layer"] == a.layer and r["cin"] == a.cin and r["k"] == a.k:
        if hits == a.nth:
            print(seen.get(kn, 0))
            sys.exit(0)
        hits += 1
    seen[kn] = seen.get(kn, 0) + 1
print("not found", file=sys.stderr)
sys.exit(1)
