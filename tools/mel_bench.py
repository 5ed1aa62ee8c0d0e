"""Throughput of the mel front-end kernel (svk_mel_spectrogram) at BASELINE configs[2]'s shape: 16 utterances x 1024 frames
of 22.05 kHz audio.  CUDA events on the launching stream, L2 flushed between iterations.  --ncu: one launch only."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "smart-vocoder_b200"))
import torch  # noqa: E402

import mel_processing as mp  # noqa: E402

B, T = 16, 1024
n = T * 256
g = torch.Generator(device="cuda").manual_seed(0)
y = (0.3 * torch.randn(B, n, device="cuda", generator=g)).clamp_(-1, 1)
args = (1024, 80, 22050, 256, 1024, 0.0, None)
mel = mp.mel_spectrogram_torch(y, *args)
torch.cuda.synchronize()
if "--ncu" in sys.argv:
    sys.exit(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
times = []
for i in range(20):
    flush.fill_(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mel = mp.mel_spectrogram_torch(y, *args)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
times.sort()
ms = times[len(times) // 2]
alg_bytes = 4 * B * n + 4 * B * 80 * T  # waveform in, log-mel out
print(json.dumps({"kernel": "mel_frontend_kernel (fused STFT + magnitude + mel + log)", "shape": f"{B} x {n} samples -> {B} x 80 x {T}",
                  "ms_median": ms, "ms_best": times[0], "samples_per_s": B * n / (ms * 1e-3),
                  "algorithmic_bytes": alg_bytes, "algorithmic_gbs": alg_bytes / (ms * 1e-3) / 1e9,
                  "share_of_infer_step": f"{ms:.3f} ms vs ~37 ms"}))
