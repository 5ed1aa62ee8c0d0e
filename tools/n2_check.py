"""2-GPU check of svk_parallel.time_sharded_infer with the real model: rank 0 holds one batch of long utterances, the
ranks synthesise halves of the time axis (halo-widened windows), rank 0 compares with its own whole-utterance infer.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 tools/n2_check.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "smart-vocoder_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import svk_parallel as P  # noqa: E402
import svk_weights as W  # noqa: E402
from models import SynthesizerTrn  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = json.load(open(os.path.join(ROOT, "configs", "iitp_base.json")))
dims = W.dims_from_model_kwargs(513, **cfg["model"])
net = SynthesizerTrn(513, 32, n_speakers=109, **cfg["model"])
net.load_state_dict({k: torch.from_numpy(v) for k, v in W.make_state_dict(dims, seed=1234).items()})
net = net.cuda().eval()
B, T = 2, 1500
g = torch.Generator().manual_seed(3)
mel = (torch.randn(B, 80, T, generator=g) * 2 - 5).to(dev) if rank == 0 else None
eps = torch.randn(B, 192, T, generator=g).to(dev) if rank == 0 else None
lengths = torch.tensor([T, T - 333], dtype=torch.int64, device=dev) if rank == 0 else None
fn = lambda m, l, e, lo, hi: net.infer_window(m, l, e, lo, hi, noise_scale=0.667)  # noqa: E731
out = P.time_sharded_infer(fn, mel, lengths, eps, B, 80, 192, T, dims.hop, net.halo_frames(), dev)
torch.cuda.synchronize()
if rank == 0:
    full = net.infer_chunked(mel, lengths, chunk_frames=T, noise_scale=0.667, eps=eps)[0]
    torch.cuda.synchronize()
    print("time_sharded_infer over 2 GPUs: identical to whole-utterance infer:", bool(torch.equal(out, full)),
          "max |diff|", float((out - full).abs().max()), "shape", tuple(out.shape))
dist.barrier()
dist.destroy_process_group()
