"""Per-step device time of N back-to-back SynthesizerTrn.infer calls (CUDA event per step): looks for outlier steps."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "smart-vocoder_b200"))
import numpy as np, torch
import svk_weights as W
from models import SynthesizerTrn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
cfg = json.load(open(os.path.join(ROOT, "configs", "iitp_base.json")))
dims = W.dims_from_model_kwargs(513, **cfg["model"])
net = SynthesizerTrn(513, 32, n_speakers=109, range_check=False, **cfg["model"])
net.load_state_dict({k: torch.from_numpy(v) for k, v in W.make_state_dict(dims, seed=1234).items()})
net = net.cuda().eval()
mel = (torch.randn(16, 80, 1024) * 2 - 5).cuda(); lengths = torch.full((16,), 1024, dtype=torch.int64).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(3):
    flush.fill_(i); net.infer(mel, lengths, noise_scale=0.667)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
ev[0].record()
for i in range(n):
    flush.fill_(i & 0xFF); net.infer(mel, lengths, noise_scale=0.667); ev[i + 1].record()
torch.cuda.synchronize()
t = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(n)])
print(f"steps {n}: mean {t.mean():.3f} median {np.median(t):.3f} min {t.min():.3f} max {t.max():.3f} p99 {np.percentile(t, 99):.3f}")
print("slowest:", sorted(((round(float(x), 2), i) for i, x in enumerate(t)), reverse=True)[:6])
print("first 12:", np.round(t[:12], 2).tolist())
