"""Module-level operators of the reference's ``modules.py`` that exist as standalone B200 kernels.

``ConvFlow`` mirrors ``modules.ConvFlow`` (reference modules.py:346-390: constructor arguments, parameter names of its
``state_dict()``, ``forward(x, x_mask, g=None, reverse=False)`` and return values) on top of ``svk_convflow``
(csrc/convflow.cu).  The reference defines the class but never instantiates it, so it is not wired into
``SynthesizerTrn``; the file is not called ``modules.py`` so that it does not shadow the reference's module of that name
for callers that keep importing ``modules.WN`` etc. from the reference checkout.
"""
from __future__ import annotations

import ctypes
import math
from collections import OrderedDict

import torch
from torch import nn

import svk_runtime as rt


class ConvFlow(nn.Module):
    def __init__(self, in_channels, filter_channels, kernel_size, n_layers, num_bins=10, tail_bound=5.0):
        super().__init__()
        self.in_channels = in_channels
        self.filter_channels = filter_channels
        self.kernel_size = kernel_size
        self.n_layers = n_layers
        self.num_bins = num_bins
        self.tail_bound = tail_bound
        self.half_channels = in_channels // 2
        F, h, k = filter_channels, self.half_channels, kernel_size
        P = nn.Parameter
        bound = lambda fan_in: 1.0 / math.sqrt(fan_in)  # noqa: E731  (torch's Conv1d default init scale)
        u = lambda *shape, fan_in: P(torch.empty(*shape).uniform_(-bound(fan_in), bound(fan_in)), requires_grad=False)  # noqa: E731
        self._p = OrderedDict()
        self._p["pre.weight"], self._p["pre.bias"] = u(F, h, 1, fan_in=h), u(F, fan_in=h)
        for i in range(n_layers):
            self._p[f"convs.convs_sep.{i}.weight"], self._p[f"convs.convs_sep.{i}.bias"] = u(F, 1, k, fan_in=k), u(F, fan_in=k)
        for i in range(n_layers):
            self._p[f"convs.convs_1x1.{i}.weight"], self._p[f"convs.convs_1x1.{i}.bias"] = u(F, F, 1, fan_in=F), u(F, fan_in=F)
        for nm in ("norms_1", "norms_2"):
            for i in range(n_layers):
                self._p[f"convs.{nm}.{i}.gamma"] = P(torch.ones(F), requires_grad=False)
                self._p[f"convs.{nm}.{i}.beta"] = P(torch.zeros(F), requires_grad=False)
        # proj is zero-initialised (modules.py:358-359)
        self._p["proj.weight"] = P(torch.zeros(h * (num_bins * 3 - 1), F, 1), requires_grad=False)
        self._p["proj.bias"] = P(torch.zeros(h * (num_bins * 3 - 1)), requires_grad=False)
        for k_, v in self._p.items():
            self.register_parameter(k_.replace(".", "__"), v)
        self._packed = None

    # state_dict surface with the reference's dotted names
    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        out = OrderedDict() if destination is None else destination
        for k_, v in self._p.items():
            out[prefix + k_] = v if keep_vars else v.detach()
        return out

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        missing = [k_ for k_ in self._p if k_ not in state_dict]
        unexpected = [k_ for k_ in state_dict if k_ not in self._p]
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for ConvFlow: missing {missing}, unexpected {unexpected}")
        with torch.no_grad():
            for k_, v in state_dict.items():
                if k_ in self._p:
                    if tuple(v.shape) != tuple(self._p[k_].shape):
                        raise RuntimeError(f"size mismatch for {k_}: {tuple(v.shape)} vs {tuple(self._p[k_].shape)}")
                    self._p[k_].copy_(v.to(self._p[k_].dtype))
        self._packed = None

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        self._packed = None
        return self

    def _pack(self, device):
        if self._packed is None or self._packed[0] != device:
            st = lambda name, leaf: torch.stack([self._p[f"convs.{name}.{i}.{leaf}"] for i in range(self.n_layers)]).to(  # noqa: E731
                device, torch.float32).contiguous() if self.n_layers else torch.empty(0, device=device)
            t = dict(pre_w=self._p["pre.weight"], pre_b=self._p["pre.bias"], sep_w=st("convs_sep", "weight"), sep_b=st("convs_sep", "bias"),
                     pw_w=st("convs_1x1", "weight"), pw_b=st("convs_1x1", "bias"), norm1_g=st("norms_1", "gamma"),
                     norm1_b=st("norms_1", "beta"), norm2_g=st("norms_2", "gamma"), norm2_b=st("norms_2", "beta"),
                     proj_w=self._p["proj.weight"], proj_b=self._p["proj.bias"])
            t = {k_: v.detach().to(device, torch.float32).contiguous() for k_, v in t.items()}
            self._packed = (device, t, rt.SvkConvFlowWeights(**{k_: v.data_ptr() for k_, v in t.items()}))
        return self._packed[2]

    def forward(self, x, x_mask, g=None, reverse=False):
        if g is not None:
            raise NotImplementedError("ConvFlow conditioning (g) is not built: no caller of the reference passes it")
        if not x.is_cuda:
            raise rt.SvkError(rt.SVK_ERR_CUDA, "ConvFlow (B200): input must be a CUDA tensor (no CPU path)")
        x = x.to(torch.float32).contiguous()
        B, C, T = x.shape
        if C != self.in_channels:
            raise RuntimeError(f"expected input[B, {self.in_channels}, T], got {list(x.shape)}")
        mask = x_mask.to(torch.float32).reshape(B, T).contiguous()
        w = self._pack(x.device)
        y = torch.empty_like(x)
        logdet = torch.empty(B, device=x.device, dtype=torch.float32)
        nbytes = rt.lib().svk_convflow_workspace_bytes(B, C, T, self.filter_channels, self.num_bins)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            rt.check(rt.lib().svk_convflow(x.data_ptr(), mask.data_ptr(), B, C, T, self.filter_channels, self.kernel_size, self.n_layers,
                                           self.num_bins, float(self.tail_bound), ctypes.byref(w), int(bool(reverse)), y.data_ptr(),
                                           None if reverse else logdet.data_ptr(), None, ws.data_ptr(), nbytes,
                                           torch.cuda.current_stream(x.device).cuda_stream))
        if not reverse:
            return y, logdet
        return y
