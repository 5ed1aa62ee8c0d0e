"""Drop-in for the reference's ``models.SynthesizerTrn`` on the inference path (B200, libsvk).

Put this directory ahead of the reference checkout on ``sys.path`` and ``inference.ipynb`` runs
unchanged: ``from models import SynthesizerTrn`` and ``from mel_processing import ...`` resolve here
(the mel front-end shim accepts the notebook's CPU tensors), while ``commons``, ``utils``,
``data_utils`` ... keep coming from the reference.

What is mirrored (reference file:line):
  * constructor signature and ignored keys            models.py:266-314   
  * ``state_dict()`` / ``load_state_dict()`` surface   utils.py:18-43      (659 keys)
  * ``infer(x, x_lengths, sid, noise_scale, length_scale, noise_scale_w, max_len)``
        -> ``(o, x_mask, (z, z_p, m_p, logs_p))``      models.py:331-339
  * sub-module calls ``dec(z)``, ``enc_p(x, x_lengths)``, ``flow(z, mask, reverse=True|False)``,
    ``enc_q(spec, lengths)``                           models.py:141-160, 35-47, 73-80, 103-110

All arithmetic runs in hand-written CUDA behind the C ABI of include/svk.h; PyTorch only owns the
device buffers, the RNG draw and the stream.  Training-time methods are out of scope and raise.
"""
from __future__ import annotations

import os

from collections import OrderedDict, namedtuple
from typing import Optional

import numpy as np
import torch
from torch import nn

import svk_runtime as rt
import svk_weights as W

_IncompatibleKeys = namedtuple("_IncompatibleKeys", ["missing_keys", "unexpected_keys"])


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class _Workspace:
    """Grow-only device scratch, one per module."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = None
            self.buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        return self.buf


class _SubModule:
    """Callable view of one stage, for callers that use ``net_g.dec(...)`` style access."""

    def __init__(self, parent: "SynthesizerTrn", fn):
        self._parent, self._fn = parent, fn

    def __call__(self, *a, **k):
        return self._fn(*a, **k)

    forward = __call__


class _DevView:
    """Zero-copy torch view of a device buffer owned by libsvk (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(int(s) for s in shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def _view(ptr, shape, typestr, device):
    return torch.as_tensor(_DevView(ptr, shape, typestr), device=device)


class InferGraph:
    """One captured ``infer`` (svk_graph_*): fixed (B, T, max_len, noise_scale); I/O lives in buffers the graph owns."""

    def __init__(self, handle: "rt.Handle", dims, B: int, T: int, Tp: int, noise_scale: float, device):
        import ctypes
        self._g = ctypes.c_void_p()
        rt.check(rt.lib().svk_graph_create(handle.ptr, B, T, Tp, float(noise_scale), ctypes.byref(self._g)))
        io = rt.SvkGraphIO()
        rt.check(rt.lib().svk_graph_buffers(self._g, ctypes.byref(io)))
        C = dims.inter_channels
        self.mel = _view(io.mel, (B, dims.n_mel, T), "<f4", device)
        self.lengths = _view(io.lengths, (B,), "<i8", device)
        self.eps = _view(io.eps, (B, C, T), "<f4", device)
        self.o = _view(io.o, (B, 1, dims.hop * Tp), "<f4", device)
        self.x_mask = _view(io.x_mask, (B, 1, T), "<f4", device)
        self.lat = tuple(_view(p, (B, C, T), "<f4", device) for p in (io.z, io.z_p, io.m_p, io.logs_p))
        self.kernel_nodes, self.programmatic_edges = int(io.kernel_nodes), bool(io.programmatic_edges)

    def launch(self, stream):
        rt.check(rt.lib().svk_graph_launch(self._g, stream))

    def close(self):
        if self._g:
            rt.lib().svk_graph_destroy(self._g)
            self._g = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class InferPipeline:
    """Pipelined host-buffer ``infer`` (svk_pipeline_*): submit() enqueues H2D + kernels + D2H on three streams and
    returns a ticket; wait() blocks until that call's PCM is in host memory.  The H2D of call i+1 and the D2H of call
    i-1 overlap the kernels of call i.  Host arrays should be pinned and must outlive the wait."""

    def __init__(self, handle: "rt.Handle", dims, B: int, T: int, Tp: int, depth: int):
        import ctypes
        self._p = ctypes.c_void_p()
        rt.check(rt.lib().svk_pipeline_create(handle.ptr, B, T, Tp, int(depth), ctypes.byref(self._p)))
        self.B, self.T, self.Tp, self.dims, self.depth = B, T, Tp, dims, int(depth)
        self._keep = {}

    def submit(self, mel: np.ndarray, lengths: np.ndarray, out: np.ndarray, eps: Optional[np.ndarray] = None, seed: int = 0,
               noise_scale: float = 1.0, x_mask: Optional[np.ndarray] = None) -> int:
        import ctypes
        d = self.dims
        for a, shape, dt, nm in ((mel, (self.B, d.n_mel, self.T), np.float32, "mel"), (lengths, (self.B,), np.int64, "lengths"),
                                 (out, (self.B, 1, d.hop * self.Tp), np.float32, "out")):
            if a.dtype != dt or tuple(a.shape) != shape or not a.flags["C_CONTIGUOUS"]:
                raise ValueError(f"{nm}: expected a C-contiguous {np.dtype(dt).name} array of shape {shape}")
        if eps is not None and (eps.dtype != np.float32 or tuple(eps.shape) != (self.B, d.inter_channels, self.T) or
                                not eps.flags["C_CONTIGUOUS"]):
            raise ValueError("eps: expected a C-contiguous float32 array [B, inter_channels, T]")
        if x_mask is not None and (x_mask.dtype != np.float32 or x_mask.size != self.B * self.T or not x_mask.flags["C_CONTIGUOUS"]):
            raise ValueError("x_mask: expected a C-contiguous float32 array [B, 1, T]")
        t = ctypes.c_int64()
        vp = lambda a: None if a is None else a.ctypes.data  # noqa: E731
        rt.check(rt.lib().svk_pipeline_submit(self._p, vp(mel), vp(lengths), vp(eps), int(seed), float(noise_scale), vp(out),
                                              vp(x_mask), ctypes.byref(t)))
        self._keep[t.value] = (mel, lengths, eps, out, x_mask)  # the copies are asynchronous: keep the arrays alive
        self._keep.pop(t.value - 2 * self.depth, None)
        return t.value

    def wait(self, ticket: int):
        rt.check(rt.lib().svk_pipeline_wait(self._p, int(ticket)))
        self._keep.pop(int(ticket), None)

    def drain(self):
        rt.check(rt.lib().svk_pipeline_drain(self._p))
        self._keep.clear()

    def close(self):
        if self._p:
            rt.lib().svk_pipeline_destroy(self._p)
            self._p = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


class SynthesizerTrn(nn.Module):
    """
    Synthesizer for Training  (inference path only; see module docstring)
    """

    def __init__(self,
                 spec_channels,
                 segment_size,
                 inter_channels,
                 hidden_channels,
                 filter_channels,
                 n_heads,
                 n_layers,
                 kernel_size,
                 p_dropout,
                 resblock,
                 resblock_kernel_sizes,
                 resblock_dilation_sizes,
                 upsample_rates,
                 upsample_initial_channel,
                 upsample_kernel_sizes,
                 n_speakers=0,
                 gin_channels=0,
                 **kwargs):
        super().__init__()
        # stored exactly like the reference does (models.py:287-303), even the ignored ones
        self.spec_channels = spec_channels
        self.inter_channels = inter_channels
        self.hidden_channels = hidden_channels
        self.filter_channels = filter_channels
        self.n_heads = n_heads
        self.n_layers = n_layers
        self.kernel_size = kernel_size
        self.p_dropout = p_dropout
        self.resblock = resblock
        self.resblock_kernel_sizes = resblock_kernel_sizes
        self.resblock_dilation_sizes = resblock_dilation_sizes
        self.upsample_rates = upsample_rates
        self.upsample_initial_channel = upsample_initial_channel
        self.upsample_kernel_sizes = upsample_kernel_sizes
        self.segment_size = segment_size
        self.n_speakers = n_speakers
        self.gin_channels = gin_channels

        self.dims = W.dims_from_model_kwargs(
            spec_channels, inter_channels=inter_channels, hidden_channels=hidden_channels, resblock=resblock,
            resblock_kernel_sizes=resblock_kernel_sizes, resblock_dilation_sizes=resblock_dilation_sizes,
            upsample_rates=upsample_rates, upsample_initial_channel=upsample_initial_channel,
            upsample_kernel_sizes=upsample_kernel_sizes, gin_channels=gin_channels)
        self.dims.validate()
        self._spec = W.state_dict_spec(self.dims)
        # Same key surface as the reference; values are a seeded stand-in for torch's random init
        # (flow.*.post zero like modules.py:321-322).  Real use loads a checkpoint over them.
        init = W.make_state_dict(self.dims, seed=int(kwargs.get("init_seed", 0)), alive=False)
        self._sd = OrderedDict((k, torch.from_numpy(v)) for k, v in init.items())
        # engine: "tc" (tcgen05 fp16x3 split, default) or "fp32" (FFMA); kwarg `engine=` or $SVK_ENGINE.
        # Not a reference hyper-parameter: the reference swallows unknown kwargs (models.py:284).
        eng = str(kwargs.get("engine", os.environ.get("SVK_ENGINE", "tc"))).lower()
        if eng not in rt.PRECISIONS:
            raise ValueError(f"engine must be one of {sorted(rt.PRECISIONS)}, got {eng!r}")
        self.engine = eng
        self._handle: Optional[rt.Handle] = None
        self._device: Optional[torch.device] = None
        self._ws = _Workspace()
        self._graphs = {}
        # Range guard (svk_check_range): the fp16 hi/lo engine turns an activation above 65504 into NaN audio.  With
        # `range_check` True (default) every infer()/dec() call ends with a check (one 4-byte D2H + stream sync, the
        # sync the reference's callers do anyway with `.cpu()`) and raises SvkError instead of returning NaNs;
        # set it False to keep calls asynchronous and call check_range() yourself.
        self.range_check = bool(kwargs.get("range_check", True))

        self.dec = _SubModule(self, self._dec_forward)
        self.enc_p = _SubModule(self, self._enc_p_forward)
        self.flow = _SubModule(self, self._flow_forward)
        self.enc_q = _SubModule(self, self._enc_q_forward)

    # ------------------------------------------------------------------ nn.Module surface
    def _apply(self, fn, *args, **kwargs):
        # probe on the device the module currently lives on: dtype-only calls (`.float()`, `.to(torch.float32)`,
        # no-ops in the reference) then keep the binding, and only a real move to the CPU releases the handle
        probe = fn(torch.empty(0, dtype=torch.float32, device=self._device if self._device is not None else "cpu"))
        if probe.dtype != torch.float32:
            raise TypeError("SynthesizerTrn (B200) computes in fp32; .half()/.double() are not supported")
        if probe.device.type == "cuda":
            self._bind(probe.device)
        elif probe.device.type == "cpu":
            self._release()
        return self

    def _release(self):
        for g in self._graphs.values():
            g.close()
        self._graphs = {}
        if self._handle is not None:
            self._handle.close()
        self._handle, self._device = None, None

    def _bind(self, device: torch.device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        device = torch.device("cuda", idx)
        if self._handle is not None and self._device == device:
            return
        self._release()
        self._handle = rt.Handle(self.dims, idx, rt.PRECISIONS[self.engine])
        self._device = device
        self._upload()

    def _upload(self):
        h = self._handle
        for k, v in self._sd.items():
            if not W.is_dead_key(k) or W.is_posterior_key(k):  # enc_q.* is optional: only enc_q(...) reads it
                h.load_tensor(k, v.detach().cpu().numpy())
        h.finalize()

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):
        out = OrderedDict() if destination is None else destination
        for k, v in self._sd.items():
            out[prefix + k] = v if keep_vars else v.detach()
        return out

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        missing = [k for k in self._sd if k not in state_dict]
        unexpected = [k for k in state_dict if k not in self._sd]
        errors = []
        for k, v in state_dict.items():
            if k in self._sd and tuple(v.shape) != tuple(self._sd[k].shape):
                errors.append(f"size mismatch for {k}: copying a param with shape {tuple(v.shape)} from checkpoint, "
                              f"the shape in current model is {tuple(self._sd[k].shape)}.")
        if strict and (missing or unexpected):
            if missing:
                errors.append("Missing key(s) in state_dict: " + ", ".join(f'"{k}"' for k in missing) + ".")
            if unexpected:
                errors.append("Unexpected key(s) in state_dict: " + ", ".join(f'"{k}"' for k in unexpected) + ".")
        if errors:
            raise RuntimeError("Error(s) in loading state_dict for SynthesizerTrn:\n\t" + "\n\t".join(errors))
        for k, v in state_dict.items():
            if k in self._sd:
                self._sd[k] = v.detach().to("cpu", torch.float32).contiguous().clone()
        if self._handle is not None:
            self._upload()
        return _IncompatibleKeys(missing, unexpected)

    def parameters(self, recurse: bool = True):
        return iter(self._sd.values())

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("training is out of scope of the B200 inference path")
        return super().train(False)

    # ------------------------------------------------------------------ helpers
    def _need_cuda(self, *tensors):
        if self._handle is None:
            raise RuntimeError("SynthesizerTrn (B200) has no CPU path: call .cuda() first "
                               "(libsvk runs on sm_100a only)")
        for t in tensors:
            if t is not None and t.device != self._device:
                raise RuntimeError(f"Expected all tensors to be on the same device, but found {t.device} and {self._device}")

    @staticmethod
    def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
        if t.dtype != torch.float32:
            raise RuntimeError(f"{name}: expected a float32 tensor but found {t.dtype}")
        return t.contiguous()

    def _stream(self):
        return torch.cuda.current_stream(self._device).cuda_stream

    @staticmethod
    def _clip_len(T: int, max_len) -> int:
        """Python slice semantics of ``[:, :, :max_len]`` (models.py:338)."""
        if max_len is None:
            return T
        max_len = int(max_len)
        return max(T + max_len, 0) if max_len < 0 else min(max_len, T)

    # ------------------------------------------------------------------ the hot path
    def infer(self, x, x_lengths, sid=None, noise_scale=1, length_scale=1, noise_scale_w=1., max_len=None):
        # sid, length_scale, noise_scale_w are accepted and ignored, exactly like models.py:331-339
        self._need_cuda(x, x_lengths)
        x = self._f32(x, "x")
        if x.dim() != 3 or x.shape[1] != self.dims.n_mel:
            raise RuntimeError(f"expected input[B, {self.dims.n_mel}, T], got {list(x.shape)}")
        B, _, T = x.shape
        lengths = x_lengths.to(torch.int64).contiguous()
        C, dev = self.dims.inter_channels, self._device
        Tp = self._clip_len(T, max_len)
        if B == 0 or T == 0 or Tp == 0:
            raise ValueError("infer: empty batch / zero frames")
        m_p = torch.empty(B, C, T, device=dev, dtype=torch.float32)
        # the only RNG consumer of the path (models.py:336); drawn by torch so that a seeded
        # generator -- or a patched randn_like -- feeds this path exactly like the reference
        eps = torch.randn_like(m_p).contiguous()
        logs_p = torch.empty_like(m_p)
        z_p = torch.empty_like(m_p)
        z = torch.empty_like(m_p)
        x_mask = torch.empty(B, 1, T, device=dev, dtype=torch.float32)
        o = torch.empty(B, 1, self.dims.hop * Tp, device=dev, dtype=torch.float32)
        nbytes = self._handle.workspace_bytes(B, T, Tp)
        ws = self._ws.get(nbytes, dev)
        with torch.cuda.device(dev):
            rt.check(rt.lib().svk_infer(self._handle.ptr, _ptr(x), _ptr(lengths), _ptr(eps), float(noise_scale),
                                        B, T, Tp, _ptr(o), _ptr(x_mask), _ptr(z), _ptr(z_p), _ptr(m_p), _ptr(logs_p),
                                        _ptr(ws), nbytes, self._stream()))
            if self.range_check:
                self._handle.check_range(self._stream())
        return o, x_mask, (z, z_p, m_p, logs_p)

    def infer_graph(self, x, x_lengths, sid=None, noise_scale=1, length_scale=1, noise_scale_w=1., max_len=None):
        """``infer`` replayed from a CUDA graph (svk_graph_*): one launch call instead of ~135, for small-batch latency.
        Same arguments and return value as ``infer`` -- but the returned tensors ALIAS buffers owned by the graph of
        this (B, T, max_len, noise_scale) and are overwritten by its next replay: clone what you keep.  The first
        call per shape captures (one eager run + instantiation)."""
        self._need_cuda(x, x_lengths)
        x = self._f32(x, "x")
        if x.dim() != 3 or x.shape[1] != self.dims.n_mel:
            raise RuntimeError(f"expected input[B, {self.dims.n_mel}, T], got {list(x.shape)}")
        B, _, T = x.shape
        Tp = self._clip_len(T, max_len)
        if B == 0 or T == 0 or Tp == 0:
            raise ValueError("infer_graph: empty batch / zero frames")
        key = (B, T, Tp, float(noise_scale))
        with torch.cuda.device(self._device):
            g = self._graphs.get(key)
            if g is None:
                if len(self._graphs) >= 8:  # bounded cache: a graph owns a full workspace
                    self._graphs.pop(next(iter(self._graphs))).close()
                g = self._graphs[key] = InferGraph(self._handle, self.dims, B, T, Tp, float(noise_scale), self._device)
            g.mel.copy_(x)
            g.lengths.copy_(x_lengths.to(torch.int64))
            g.eps.copy_(torch.randn_like(g.lat[2]))  # the draw of models.py:336, through torch like infer()
            g.launch(self._stream())
            if self.range_check:
                self._handle.check_range(self._stream())
        return g.o, g.x_mask, g.lat

    def pipeline(self, B, T, depth=2, max_len=None) -> InferPipeline:
        """Pipelined host-buffer entry for a fixed batch shape (svk_pipeline_*); see InferPipeline."""
        self._need_cuda()
        return InferPipeline(self._handle, self.dims, int(B), int(T), self._clip_len(int(T), max_len), depth)

    def check_range(self):
        """Raise SvkError(SVK_ERR_RANGE) if any call since the last check produced a non-finite sample (synchronises)."""
        self._need_cuda()
        with torch.cuda.device(self._device):
            self._handle.check_range(self._stream())

    def infer_host(self, mel: np.ndarray, lengths: np.ndarray, eps: np.ndarray, noise_scale=1.0, max_len=None,
                   out: Optional[np.ndarray] = None, want_latents: bool = False):
        """Host-buffer entry (svk_infer_host): what inference.ipynb:114-118 does around ``infer``
        (``.cuda()`` ... ``.cpu()``), copies included.  Arrays should live in pinned memory."""
        self._need_cuda()
        # the C side copies raw bytes: coerce dtype / layout here (no-ops for the arrays a careful caller passes)
        mel = np.ascontiguousarray(mel, dtype=np.float32)
        if mel.ndim != 3 or mel.shape[1] != self.dims.n_mel:
            raise ValueError(f"mel: expected [B, {self.dims.n_mel}, T], got {list(mel.shape)}")
        B, _, T = mel.shape
        lengths = np.ascontiguousarray(lengths, dtype=np.int64).reshape(-1)
        eps = np.ascontiguousarray(eps, dtype=np.float32)
        if lengths.shape[0] != B:
            raise ValueError(f"lengths: expected [{B}], got {list(lengths.shape)}")
        if tuple(eps.shape) != (B, self.dims.inter_channels, T):
            raise ValueError(f"eps: expected [{B}, {self.dims.inter_channels}, {T}], got {list(eps.shape)}")
        Tp = self._clip_len(T, max_len)
        if B == 0 or T == 0 or Tp == 0:
            raise ValueError("infer_host: empty batch / zero frames")
        if out is not None and (out.dtype != np.float32 or not out.flags["C_CONTIGUOUS"] or
                                out.size != B * self.dims.hop * Tp):
            raise ValueError(f"out: expected a C-contiguous float32 array of {B * self.dims.hop * Tp} elements")
        if out is None:
            out = np.empty((B, 1, self.dims.hop * Tp), np.float32)
        mask = np.empty((B, 1, T), np.float32)
        lat = [np.empty((B, self.dims.inter_channels, T), np.float32) if want_latents else None for _ in range(4)]
        vp = lambda a: None if a is None else a.ctypes.data  # noqa: E731
        rt.check(rt.lib().svk_infer_host(self._handle.ptr, vp(mel), vp(lengths), vp(eps), float(noise_scale), B, T, Tp,
                                         vp(out), vp(mask), vp(lat[0]), vp(lat[1]), vp(lat[2]), vp(lat[3])))
        return out, mask, tuple(lat)

    def last_launch_count(self) -> int:
        return self._handle.last_launch_count() if self._handle else 0

    # ------------------------------------------------------------------ chunked / streaming synthesis
    def halo_frames(self) -> int:
        """Frames of context per side that make a window's PCM equal to the whole-utterance result."""
        self._need_cuda()
        return int(rt.lib().svk_halo_frames(self._handle.ptr))

    def _chunk_inputs(self, x, x_lengths, eps):
        self._need_cuda(x, x_lengths)
        x = self._f32(x, "x")
        if x.dim() != 3 or x.shape[1] != self.dims.n_mel:
            raise RuntimeError(f"expected input[B, {self.dims.n_mel}, T], got {list(x.shape)}")
        B, _, T = x.shape
        lengths = x_lengths.to(torch.int64).contiguous()
        if eps is None:  # same single draw over the whole utterance as models.py:336
            eps = torch.randn_like(torch.empty(B, self.dims.inter_channels, T, device=self._device, dtype=torch.float32))
        eps = self._f32(eps, "eps")
        if tuple(eps.shape) != (B, self.dims.inter_channels, T):
            raise RuntimeError(f"eps: expected [{B}, {self.dims.inter_channels}, {T}], got {list(eps.shape)}")
        return x, lengths, eps, B, T

    def infer_chunked(self, x, x_lengths, chunk_frames=256, noise_scale=1, max_len=None, eps=None):
        """``infer`` with bounded workspace: the time axis is walked in windows of ``chunk_frames`` frames, each
        recomputed with ``halo_frames()`` frames of context, so the result equals ``infer`` on the same inputs
        (svk_infer_chunked).  Same return value as ``infer`` (models.py:331-339)."""
        x, lengths, eps, B, T = self._chunk_inputs(x, x_lengths, eps)
        Tp = self._clip_len(T, max_len)
        if B == 0 or T == 0 or Tp == 0:
            raise ValueError("infer_chunked: empty batch / zero frames")
        C, dev = self.dims.inter_channels, self._device
        want_lat = Tp == T
        lat = [torch.empty(B, C, T, device=dev, dtype=torch.float32) if want_lat else None for _ in range(4)]
        x_mask = torch.empty(B, 1, T, device=dev, dtype=torch.float32)
        o = torch.empty(B, 1, self.dims.hop * Tp, device=dev, dtype=torch.float32)
        chunk = min(int(chunk_frames), Tp)
        nbytes = int(rt.lib().svk_window_workspace_bytes(self._handle.ptr, B, chunk))
        ws = self._ws.get(nbytes, dev)
        with torch.cuda.device(dev):
            rt.check(rt.lib().svk_infer_chunked(self._handle.ptr, _ptr(x), _ptr(lengths), _ptr(eps), float(noise_scale), B, T,
                                                Tp, chunk, _ptr(o), _ptr(x_mask), _ptr(lat[0]), _ptr(lat[1]), _ptr(lat[2]),
                                                _ptr(lat[3]), _ptr(ws), nbytes, self._stream()))
        return o, x_mask, tuple(lat)

    def infer_stream(self, x, x_lengths, chunk_frames=64, noise_scale=1, max_len=None, eps=None):
        """Generator over PCM chunks ``[B, 1, hop * frames]`` in time order (svk_infer_window): the first audio is
        available after one window instead of after the whole utterance.  Concatenated along time the chunks equal
        ``infer``'s ``o`` on the same inputs."""
        x, lengths, eps, B, T = self._chunk_inputs(x, x_lengths, eps)
        Tp = self._clip_len(T, max_len)
        chunk = max(1, min(int(chunk_frames), Tp))
        nbytes = int(rt.lib().svk_window_workspace_bytes(self._handle.ptr, B, chunk))
        ws = self._ws.get(nbytes, self._device)
        hop = self.dims.hop
        for t0 in range(0, Tp, chunk):
            t1 = min(t0 + chunk, Tp)
            o = torch.empty(B, 1, hop * (t1 - t0), device=self._device, dtype=torch.float32)
            with torch.cuda.device(self._device):
                rt.check(rt.lib().svk_infer_window(self._handle.ptr, _ptr(x), _ptr(lengths), _ptr(eps), float(noise_scale), B, T,
                                                   Tp, t0, t1, _ptr(o), hop * (t1 - t0), None, None, None, None, 0, _ptr(ws),
                                                   nbytes, self._stream()))
            yield o

    def infer_window(self, x, x_lengths, eps, lo, hi, noise_scale=1):
        """PCM ``[B, 1, hop * (hi - lo)]`` of frames [lo, hi) of ``x`` (svk_infer_window): equal to the same slice of
        ``infer(x, ...)``'s output when ``x`` itself carries ``halo_frames()`` frames of context around [lo, hi) or
        reaches the true sequence ends.  Building block of ``svk_parallel.time_sharded_infer``."""
        x, lengths, eps, B, T = self._chunk_inputs(x, x_lengths, eps)
        lo, hi = int(lo), int(hi)
        nbytes = int(rt.lib().svk_window_workspace_bytes(self._handle.ptr, B, max(hi - lo, 1)))
        ws = self._ws.get(nbytes, self._device)
        hop = self.dims.hop
        o = torch.empty(B, 1, hop * max(hi - lo, 0), device=self._device, dtype=torch.float32)
        with torch.cuda.device(self._device):
            rt.check(rt.lib().svk_infer_window(self._handle.ptr, _ptr(x), _ptr(lengths), _ptr(eps), float(noise_scale), B, T, 0,
                                               lo, hi, _ptr(o), hop * (hi - lo), None, None, None, None, 0, _ptr(ws), nbytes,
                                               self._stream()))
        return o

    # ------------------------------------------------------------------ sub-module views
    def _dec_forward(self, x, g=None):
        """Generator.forward(x, g=None) (models.py:141-160)."""
        self._need_cuda(x)
        x = self._f32(x, "x")
        B, C, L = x.shape
        if C != self.dims.inter_channels:
            raise RuntimeError(f"expected input[B, {self.dims.inter_channels}, L], got {list(x.shape)}")
        o = torch.empty(B, 1, self.dims.hop * L, device=self._device, dtype=torch.float32)
        nbytes = self._handle.workspace_bytes(B, L, L)
        ws = self._ws.get(nbytes, self._device)
        with torch.cuda.device(self._device):
            rt.check(rt.lib().svk_generator(self._handle.ptr, _ptr(x), B, L, _ptr(o), _ptr(ws), nbytes, self._stream()))
            if self.range_check:
                self._handle.check_range(self._stream())
        return o

    def _enc_p_forward(self, x, x_lengths, g=None):
        """MelEncoder.forward (models.py:35-47) -> (x, m, logs, x_mask)."""
        self._need_cuda(x, x_lengths)
        x = self._f32(x, "x")
        B, _, T = x.shape
        dev, H, C = self._device, self.dims.hidden_channels, self.dims.inter_channels
        lengths = x_lengths.to(torch.int64).contiguous()
        xo = torch.empty(B, H, T, device=dev, dtype=torch.float32)
        m = torch.empty(B, C, T, device=dev, dtype=torch.float32)
        logs = torch.empty_like(m)
        mask = torch.empty(B, 1, T, device=dev, dtype=torch.float32)
        nbytes = self._handle.workspace_bytes(B, T, T)
        ws = self._ws.get(nbytes, dev)
        with torch.cuda.device(dev):
            rt.check(rt.lib().svk_mel_encoder(self._handle.ptr, _ptr(x), _ptr(lengths), B, T, _ptr(xo), _ptr(m),
                                              _ptr(logs), _ptr(mask), _ptr(ws), nbytes, self._stream()))
        return xo, m, logs, mask

    def _flow_forward(self, x, x_mask, g=None, reverse=False):
        """ResidualCouplingBlock.forward (models.py:73-80): reverse=True is on the infer path, reverse=False is the
        analysis direction (z -> z_p, models.py:323)."""
        self._need_cuda(x, x_mask)
        z = self._f32(x, "x").clone()
        mask = self._f32(x_mask, "x_mask")
        B, C, T = z.shape
        nbytes = self._handle.workspace_bytes(B, T, T)
        ws = self._ws.get(nbytes, self._device)
        fn = rt.lib().svk_flow_reverse if reverse else rt.lib().svk_flow_forward
        with torch.cuda.device(self._device):
            rt.check(fn(self._handle.ptr, _ptr(z), _ptr(mask), B, T, _ptr(ws), nbytes, self._stream()))
        return z

    def _enc_q_forward(self, x, x_lengths, g=None):
        """PosteriorEncoder.forward (models.py:103-110): linear spectrogram -> (z, m, logs, x_mask)."""
        self._need_cuda(x, x_lengths)
        x = self._f32(x, "x")
        if x.dim() != 3 or x.shape[1] != self.dims.spec_channels:
            raise RuntimeError(f"expected input[B, {self.dims.spec_channels}, T], got {list(x.shape)}")
        B, _, T = x.shape
        dev, C = self._device, self.dims.inter_channels
        lengths = x_lengths.to(torch.int64).contiguous()
        m = torch.empty(B, C, T, device=dev, dtype=torch.float32)
        eps = torch.randn_like(m).contiguous()  # the draw of models.py:109
        logs, z = torch.empty_like(m), torch.empty_like(m)
        mask = torch.empty(B, 1, T, device=dev, dtype=torch.float32)
        nbytes = self._handle.workspace_bytes(B, T, T)
        ws = self._ws.get(nbytes, dev)
        with torch.cuda.device(dev):
            rt.check(rt.lib().svk_posterior_encoder(self._handle.ptr, _ptr(x), _ptr(lengths), _ptr(eps), B, T, _ptr(z), _ptr(m),
                                                    _ptr(logs), _ptr(mask), _ptr(ws), nbytes, self._stream()))
        return z, m, logs, mask

    # ------------------------------------------------------------------ out of scope (training)
    def forward(self, x, x_lengths, y, y_lengths, sid=None):
        raise NotImplementedError("SynthesizerTrn.forward is the training graph (models.py:317-329); "
                                  "out of scope of the B200 inference path")

    def voice_conversion(self, y, y_lengths, sid_src, sid_tgt):
        # the reference asserts, then dies on the never-created emb_g (models.py:342-343)
        assert self.n_speakers > 0, "n_speakers have to be larger than 0."
        raise AttributeError("'SynthesizerTrn' object has no attribute 'emb_g'")
