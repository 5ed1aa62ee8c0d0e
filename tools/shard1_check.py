"""Developer check: svk_parallel.sharded_infer on ONE GPU (NCCL world of 1: no NCCL kernels), per-step wall time with 1 and 2 micro-batches."""
import json, os, sys, time
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "smart-vocoder_b200"))
import torch, torch.distributed as dist
import svk_weights as W, svk_parallel as P
from models import SynthesizerTrn
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("nccl", rank=0, world_size=1)
cfg = json.load(open(os.path.join(ROOT, "configs", "iitp_base.json")))
dims = W.dims_from_model_kwargs(513, **cfg["model"])
net = SynthesizerTrn(513, 32, n_speakers=109, range_check=False, **cfg["model"])
net.load_state_dict({k: torch.from_numpy(v) for k, v in W.make_state_dict(dims, seed=1234).items()})
net = net.cuda().eval()
dev = torch.device("cuda:0")
mel = (torch.randn(16, 80, 1024) * 2 - 5).pin_memory(); lengths = torch.full((16,), 1024, dtype=torch.int64).pin_memory()
out = torch.empty(16, 1, 256 * 1024).pin_memory()
fn = lambda m, l: net.infer(m, l, noise_scale=0.667)[0]
for rep in range(2):
    for M in (1, 2):
        for _ in range(2):
            P.sharded_infer(fn, mel, lengths, 16, 80, 1024, 256, dev, micro_batches=M, out_host=out)
        ts = []
        for _ in range(8):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            P.sharded_infer(fn, mel, lengths, 16, 80, 1024, 256, dev, micro_batches=M, out_host=out)
            torch.cuda.synchronize(); ts.append(round(1e3 * (time.perf_counter() - t0), 1))
        print("stack", os.environ.get("SVK_WN_STACK", "1"), "micro", M, "ms/step", ts)
