"""B = 1 latency distribution of SynthesizerTrn.infer (one 80x1024 utterance, CUDA events around each call)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "smart-vocoder_b200"))
import numpy as np, torch
import svk_weights as W
from models import SynthesizerTrn
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
cfg = json.load(open(os.path.join(ROOT, "configs", "iitp_base.json")))
dims = W.dims_from_model_kwargs(513, **cfg["model"])
net = SynthesizerTrn(513, 32, n_speakers=109, range_check=False, **cfg["model"])
net.load_state_dict({k: torch.from_numpy(v) for k, v in W.make_state_dict(dims, seed=1234).items()})
net = net.cuda().eval()
mel = (torch.randn(1, 80, 1024) * 2 - 5).cuda(); lengths = torch.full((1,), 1024, dtype=torch.int64).cuda()
for i in range(5):
    net.infer(mel, lengths, noise_scale=0.667)
torch.cuda.synchronize()
lat = []
for i in range(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); net.infer(mel, lengths, noise_scale=0.667); e1.record(); torch.cuda.synchronize()
    lat.append(e0.elapsed_time(e1))
lat = np.array(lat)
print(f"{os.environ.get('TAG','')}: median {np.median(lat):.3f} min {lat.min():.3f} max {lat.max():.3f} p90 {np.percentile(lat,90):.3f}  >4ms: {(lat>4).sum()}/{n}")
