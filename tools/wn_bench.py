import os, sys, json, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/smart-vocoder_b200')
import numpy as np, torch
import svk_weights as W
from models import SynthesizerTrn
cfg = json.load(open('/root/repo/configs/iitp_base.json'))
dims = W.dims_from_model_kwargs(513, **cfg['model'])
net = SynthesizerTrn(513, 32, n_speakers=109, range_check=False, **cfg['model'])
net.load_state_dict({k: torch.from_numpy(v) for k, v in W.make_state_dict(dims, seed=1234).items()})
net = net.cuda().eval()
B, T = int(sys.argv[1]) if len(sys.argv) > 1 else 16, 1024
mel = (torch.randn(B, 80, T) * 2 - 5).cuda(); ln = torch.full((B,), T, dtype=torch.int64).cuda()
for _ in range(3): net.enc_p(mel, ln)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): net.enc_p(mel, ln)
e1.record(); torch.cuda.synchronize()
print('enc_p (16 WN layers + pre/proj) B=%d: %.3f ms per call, launches %d' % (B, e0.elapsed_time(e1) / 20, net.last_launch_count()))
