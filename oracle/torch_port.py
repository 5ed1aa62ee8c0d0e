"""TEST / BASELINE INFRASTRUCTURE ONLY -- the reference path restated on torch CPU ops.

The reference's arithmetic lives in PyTorch (third party: pinned torch==1.6.0 in the reference's
requirements.txt:6; this image has 2.11.0 with MKL-DNN): ``nn.Conv1d`` / ``nn.ConvTranspose1d`` /
``weight_norm`` / ``tanh`` / ``sigmoid`` / ``leaky_relu``.  This module restates
``SynthesizerTrn.infer`` (reference models.py:331-339) as one flat function over a ``state_dict``
using the same torch operators in the same order, INCLUDING the per-call weight_norm recompute
(172 per infer) that the reference pays -- so timing it on the host cores is the
reference's own CPU cost.  It cannot import /root/reference (absent on the GPU box).

Used by bench.py (``cpu_baseline`` leg and ``--impl reference``) and by tests as a second checker.
Never imported by the product path.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F


def _w(sd: Dict[str, torch.Tensor], p: str) -> torch.Tensor:
    """Effective weight; legacy weight_norm hook = g * v / ||v|| over dims != 0, recomputed per call."""
    if p + ".weight_v" in sd:
        return torch._weight_norm(sd[p + ".weight_v"], sd[p + ".weight_g"], 0)
    return sd[p + ".weight"]


def _wn(sd, p: str, x, mask, n_layers: int, hidden: int, k: int):
    """modules.WN.forward, g=None (modules.py:148-176)."""
    out = torch.zeros_like(x)
    for i in range(n_layers):
        a = F.conv1d(x, _w(sd, f"{p}.in_layers.{i}"), sd[f"{p}.in_layers.{i}.bias"], padding=(k - 1) // 2)
        a = a + torch.zeros_like(a)  # g_l = zeros (modules.py:161), kept: the reference pays for it
        acts = torch.tanh(a[:, :hidden]) * torch.sigmoid(a[:, hidden:])
        rs = F.conv1d(acts, _w(sd, f"{p}.res_skip_layers.{i}"), sd[f"{p}.res_skip_layers.{i}.bias"])
        if i < n_layers - 1:
            x = (x + rs[:, :hidden]) * mask
            out = out + rs[:, hidden:]
        else:
            out = out + rs
    return out * mask


@torch.no_grad()
def infer(sd: Dict[str, torch.Tensor], d, mel: torch.Tensor, lengths: torch.Tensor, eps: Optional[torch.Tensor],
          noise_scale: float = 1.0, max_len: Optional[int] = None):
    """SynthesizerTrn.infer (models.py:331-339).  ``eps=None`` draws with torch.randn_like like the reference."""
    H, C, k = d.hidden_channels, d.inter_channels, d.wn_kernel
    # MelEncoder.forward (models.py:35-47)
    x = F.conv1d(mel, sd["enc_p.pre_enc.weight"], sd["enc_p.pre_enc.bias"])
    T = mel.size(2)
    mask = (torch.arange(T, dtype=lengths.dtype)[None, :] < lengths[:, None])[:, None, :].to(mel.dtype)
    x = _wn(sd, "enc_p.encoder", x * mask, mask, d.enc_layers, H, k)
    stats = F.conv1d(x, sd["enc_p.proj.weight"], sd["enc_p.proj.bias"]) * mask
    m_p, logs_p = torch.split(stats, C, dim=1)
    e = torch.randn_like(m_p) if eps is None else eps
    z_p = m_p + e * torch.exp(logs_p) * noise_scale
    # ResidualCouplingBlock reverse (models.py:77-79; modules.py:324-343)
    z = z_p
    for f in reversed(range(d.n_flows)):
        z = torch.flip(z, [1])
        p = f"flow.flows.{2 * f}"
        x0, x1 = torch.split(z, [C // 2] * 2, 1)
        h = F.conv1d(x0, sd[p + ".pre.weight"], sd[p + ".pre.bias"]) * mask
        h = _wn(sd, p + ".enc", h, mask, d.flow_layers, H, k)
        m = F.conv1d(h, sd[p + ".post.weight"], sd[p + ".post.bias"]) * mask
        logs = torch.zeros_like(m)
        x1 = (x1 - m) * torch.exp(-logs) * mask
        z = torch.cat([x0, x1], 1)
    # Generator.forward (models.py:141-160)
    y = F.conv1d((z * mask)[:, :, :max_len], sd["dec.conv_pre.weight"], sd["dec.conv_pre.bias"], padding=3)
    nk = len(d.resblock_kernel_sizes)
    for i, (u, ku) in enumerate(zip(d.upsample_rates, d.upsample_kernel_sizes)):
        y = F.leaky_relu(y, 0.1)
        y = F.conv_transpose1d(y, _w(sd, f"dec.ups.{i}"), sd[f"dec.ups.{i}.bias"], stride=u, padding=(ku - u) // 2)
        xs = None
        for j, (rk, rd) in enumerate(zip(d.resblock_kernel_sizes, d.resblock_dilation_sizes)):
            p = f"dec.resblocks.{i * nk + j}"
            r = y
            for l, dil in enumerate(rd):  # ResBlock1.forward (modules.py:210-223)
                xt = F.leaky_relu(r, 0.1)
                xt = F.conv1d(xt, _w(sd, f"{p}.convs1.{l}"), sd[f"{p}.convs1.{l}.bias"], dilation=dil,
                              padding=(rk * dil - dil) // 2)
                xt = F.leaky_relu(xt, 0.1)
                xt = F.conv1d(xt, _w(sd, f"{p}.convs2.{l}"), sd[f"{p}.convs2.{l}.bias"], padding=(rk - 1) // 2)
                r = xt + r
            xs = r if xs is None else xs + r
        y = xs / nk
    y = F.leaky_relu(y)
    y = F.conv1d(y, sd["dec.conv_post.weight"], None, padding=3)
    return torch.tanh(y), mask, (z, z_p, m_p, logs_p)
