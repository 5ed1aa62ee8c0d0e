"""Small end-to-end run for compute-sanitizer: front-end -> infer (both tcgen05 paths incl. lean epilogues, pairs, running sum,
conv_post stream kernel) -> chunked infer -> posterior encoder / flow forward, at shapes with ragged tiles."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "smart-vocoder_b200"))
import torch  # noqa: E402

import mel_processing as mp  # noqa: E402
import svk_weights as W  # noqa: E402
from models import SynthesizerTrn  # noqa: E402

cfg = json.load(open(os.path.join(ROOT, "configs", "iitp_base.json")))
dims = W.dims_from_model_kwargs(513, **cfg["model"])
net = SynthesizerTrn(513, 32, n_speakers=109, **cfg["model"])
net.load_state_dict({k: torch.from_numpy(v) for k, v in W.make_state_dict(dims, seed=1234).items()})
net = net.cuda().eval()
g = torch.Generator(device="cuda").manual_seed(0)
y = (0.3 * torch.randn(2, 37 * 256 + 77, device="cuda", generator=g)).clamp_(-1, 1)
mel = mp.mel_spectrogram_torch(y, 1024, 80, 22050, 256, 1024, 0.0, None)
spec = mp.spectrogram_torch(y, 1024, 22050, 256, 1024)
T = mel.shape[2]
lengths = torch.tensor([T, T - 9], device="cuda")
o = net.infer(mel, lengths, noise_scale=0.667)[0]
oc = net.infer_chunked(mel, lengths, chunk_frames=16, noise_scale=0.667)[0]
z, m, logs, mask = net.enc_q(spec, lengths)
zp = net.flow(z, mask)
torch.cuda.synchronize()
print("ok", tuple(o.shape), tuple(oc.shape), tuple(zp.shape), bool(torch.isfinite(o).all()))
