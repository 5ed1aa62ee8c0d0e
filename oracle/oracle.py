"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the SMART-Vocoder mel->waveform path.

Orchestration (numpy) over the plain-C operators of svk_oracle.c.  Follows the reference
call stack of ``SynthesizerTrn.infer`` (reference models.py:331-339):

  MelEncoder.forward          models.py:35-47
  WN.forward                  modules.py:148-176
  ResidualCouplingBlock(rev)  models.py:73-80
  ResidualCouplingLayer(rev)  modules.py:324-343
  Flip                        modules.py:270-277
  Generator.forward           models.py:141-160
  ResBlock1.forward           modules.py:210-223
  spline operator             transforms.py:12-193
  ConvFlow / DDSConv / LayerNorm   modules.py:363-390, 70-108, 20-32  (standalone operator)

Pinning: the reference ships no golden vectors; this oracle is pinned against outputs of the
reference itself (tests/golden/*.npz, produced by tests/golden/make_golden.py in the build
container where /root/reference is importable).  See tests/test_oracle.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
The product path (smart-vocoder_b200/) never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsvk_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile libsvk_oracle.so with gcc (building the checker is not using it)."""
    src = [os.path.join(_HERE, "svk_oracle.c"), os.path.join(_HERE, "svk_oracle_impl.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libsvk_oracle.so"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        try:
            _lib = ctypes.CDLL(_LIB_PATH)
        except OSError:
            build(force=True)
            _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class Oracle:
    """Operators in fp32 (``np.float32``) or fp64 (``np.float64``, ground truth)."""

    def __init__(self, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        assert self.dtype in (np.dtype(np.float32), np.dtype(np.float64))
        self.sfx = "f32" if self.dtype == np.float32 else "f64"
        self.creal = ctypes.c_float if self.dtype == np.float32 else ctypes.c_double

    def _fn(self, name):
        return getattr(lib(), f"svko_{self.sfx}_{name}")

    def arr(self, a) -> np.ndarray:
        return np.ascontiguousarray(np.asarray(a), dtype=self.dtype)

    # ---- operators ------------------------------------------------------------------
    def weight_norm(self, v, g) -> np.ndarray:
        v, g = self.arr(v), self.arr(g)
        w = np.empty_like(v)
        d0 = v.shape[0]
        self._fn("weight_norm")(_ptr(v), _ptr(g), ctypes.c_int64(d0), ctypes.c_int64(v.size // d0), _ptr(w))
        return w

    def conv1d(self, x, w, b=None, dilation: int = 1, padding: int = 0) -> np.ndarray:
        x, w = self.arr(x), self.arr(w)
        b = None if b is None else self.arr(b)
        B, Cin, L = x.shape
        Cout, Cin2, k = w.shape
        assert Cin == Cin2
        Lout = L + 2 * padding - dilation * (k - 1)
        y = np.empty((B, Cout, Lout), self.dtype)
        i64 = ctypes.c_int64
        self._fn("conv1d")(_ptr(x), i64(B), i64(Cin), i64(L), _ptr(w), _ptr(b), i64(Cout), i64(k),
                           i64(dilation), i64(padding), _ptr(y))
        return y

    def conv_transpose1d(self, x, w, b, stride: int, padding: int) -> np.ndarray:
        x, w = self.arr(x), self.arr(w)
        b = None if b is None else self.arr(b)
        B, Cin, L = x.shape
        Cin2, Cout, k = w.shape
        assert Cin == Cin2
        Lout = (L - 1) * stride - 2 * padding + k
        y = np.empty((B, Cout, Lout), self.dtype)
        i64 = ctypes.c_int64
        self._fn("conv_transpose1d")(_ptr(x), i64(B), i64(Cin), i64(L), _ptr(w), _ptr(b), i64(Cout),
                                     i64(k), i64(stride), i64(padding), _ptr(y))
        return y

    def gate(self, a) -> np.ndarray:
        a = self.arr(a)
        B, C2, L = a.shape
        out = np.empty((B, C2 // 2, L), self.dtype)
        i64 = ctypes.c_int64
        self._fn("gate")(_ptr(a), i64(B), i64(C2 // 2), i64(L), _ptr(out))
        return out

    def leaky_relu(self, x, slope: float) -> np.ndarray:
        x = self.arr(x)
        y = np.empty_like(x)
        self._fn("leaky_relu")(_ptr(x), ctypes.c_int64(x.size), self.creal(slope), _ptr(y))
        return y

    @staticmethod
    def sequence_mask(lengths, T: int) -> np.ndarray:
        """commons.sequence_mask (commons.py:121-125) -> float [B,1,T] as models.py:40."""
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        B = lengths.shape[0]
        m = np.empty((B, T), np.float32)
        lib().svko_sequence_mask(_ptr(lengths), ctypes.c_int64(B), ctypes.c_int64(T), _ptr(m))
        return m[:, None, :]

    @staticmethod
    def flip(x) -> np.ndarray:
        """modules.Flip (modules.py:270-277): reverse the channel axis."""
        return np.ascontiguousarray(x[:, ::-1, :])

    def rq_spline(self, inputs, uw, uh, ud, inverse: bool, tail_bound: float = 5.0,
                  min_bin_width=1e-3, min_bin_height=1e-3, min_derivative=1e-3):
        """piecewise_rational_quadratic_transform(tails='linear').  Returns (out, logabsdet, bin)."""
        x, uw, uh, ud = self.arr(inputs), self.arr(uw), self.arr(uh), self.arr(ud)
        nb = uw.shape[-1]
        assert uh.shape[-1] == nb and ud.shape[-1] == nb - 1
        assert uw.shape[:-1] == x.shape
        # reference transforms.py:108-111
        if min_bin_width * nb > 1.0:
            raise ValueError("Minimal bin width too large for the number of bins")
        if min_bin_height * nb > 1.0:
            raise ValueError("Minimal bin height too large for the number of bins")
        out = np.empty_like(x)
        lad = np.empty_like(x)
        bins = np.empty(x.shape, np.int32)
        r = self.creal
        self._fn("rq_spline")(_ptr(x), _ptr(uw), _ptr(uh), _ptr(ud), ctypes.c_int64(x.size),
                              ctypes.c_int(nb), ctypes.c_int(1 if inverse else 0), r(tail_bound),
                              r(min_bin_width), r(min_bin_height), r(min_derivative), _ptr(out),
                              _ptr(lad), _ptr(bins))
        return out, lad, bins

    # ---- ConvFlow (standalone operator; never instantiated by the reference model) ----
    def layer_norm(self, x, gamma, beta, eps: float = 1e-5) -> np.ndarray:
        """modules.LayerNorm.forward (modules.py:20-32): F.layer_norm over the channel axis of [B, C, T]."""
        x = self.arr(x)
        mean = x.mean(axis=1, keepdims=True)
        var = ((x - mean) ** 2).mean(axis=1, keepdims=True)
        y = (x - mean) / np.sqrt(var + self.dtype.type(eps))
        return (y * self.arr(gamma)[None, :, None] + self.arr(beta)[None, :, None]).astype(self.dtype)

    def gelu(self, x) -> np.ndarray:
        """F.gelu (erf form, torch's default)."""
        from math import erf, sqrt
        x = self.arr(x)
        e = np.vectorize(erf, otypes=[np.float64])(x.astype(np.float64) / sqrt(2.0))
        return (0.5 * x * (1.0 + e)).astype(self.dtype)

    def dds_conv(self, w: Dict[str, np.ndarray], prefix: str, x, mask, channels: int, kernel: int, n_layers: int) -> np.ndarray:
        """modules.DDSConv.forward, g=None (modules.py:96-108): depthwise dilated conv (dilation kernel**i) of x * mask,
        LayerNorm, GELU, 1x1 conv, LayerNorm, GELU, residual; result * mask."""
        x, mask = self.arr(x), self.arr(mask)
        for i in range(n_layers):
            dil = kernel ** i
            pad = (kernel * dil - dil) // 2
            sw, sb = self.arr(w[f"{prefix}.convs_sep.{i}.weight"]), self.arr(w[f"{prefix}.convs_sep.{i}.bias"])
            xm = x * mask
            y = np.empty_like(x)
            for c in range(channels):  # groups = channels: one single-channel conv per channel
                y[:, c:c + 1] = self.conv1d(xm[:, c:c + 1], sw[c:c + 1], sb[c:c + 1], dilation=dil, padding=pad)
            y = self.gelu(self.layer_norm(y, w[f"{prefix}.norms_1.{i}.gamma"], w[f"{prefix}.norms_1.{i}.beta"]))
            y = self.conv1d(y, self.arr(w[f"{prefix}.convs_1x1.{i}.weight"]), self.arr(w[f"{prefix}.convs_1x1.{i}.bias"]))
            y = self.gelu(self.layer_norm(y, w[f"{prefix}.norms_2.{i}.gamma"], w[f"{prefix}.norms_2.{i}.beta"]))
            x = x + y
        return x * mask

    def convflow(self, w: Dict[str, np.ndarray], x, mask, filter_channels: int, kernel: int, n_layers: int, num_bins: int = 10,
                 tail_bound: float = 5.0, reverse: bool = False):
        """modules.ConvFlow.forward (modules.py:363-390).  `w` holds the module's state_dict.  Returns (y, logdet, bins);
        logdet is None for reverse=True, like the reference."""
        x, mask = self.arr(x), self.arr(mask)
        B, C, T = x.shape
        half = C // 2
        x0, x1 = x[:, :half], x[:, half:]
        h = self.conv1d(x0, self.arr(w["pre.weight"]), self.arr(w["pre.bias"]))
        h = self.dds_conv(w, "convs", h, mask, filter_channels, kernel, n_layers)
        h = self.conv1d(h, self.arr(w["proj.weight"]), self.arr(w["proj.bias"])) * mask
        h = h.reshape(B, half, -1, T).transpose(0, 1, 3, 2)  # [b, c, t, 3nb-1]
        sf = self.dtype.type(np.sqrt(float(filter_channels)))
        uw = h[..., :num_bins] / sf
        uh = h[..., num_bins:2 * num_bins] / sf
        ud = h[..., 2 * num_bins:]
        y1, lad, bins = self.rq_spline(np.ascontiguousarray(x1), np.ascontiguousarray(uw), np.ascontiguousarray(uh),
                                       np.ascontiguousarray(ud), reverse, tail_bound)
        y = np.concatenate([x0, y1], axis=1) * mask
        logdet = None if reverse else (lad * mask).sum(axis=(1, 2))
        return y, logdet, bins

    # ---- weights --------------------------------------------------------------------
    def conv_w(self, sd: Dict[str, np.ndarray], prefix: str) -> Tuple[np.ndarray, Optional[np.ndarray]]:
        """Effective (weight, bias) of a layer: weight-normed (``weight_g``/``weight_v``) or plain."""
        if prefix + ".weight_v" in sd:
            w = self.weight_norm(sd[prefix + ".weight_v"], sd[prefix + ".weight_g"])
        else:
            w = self.arr(sd[prefix + ".weight"])
        b = sd.get(prefix + ".bias")
        return w, (None if b is None else self.arr(b))

    # ---- modules --------------------------------------------------------------------
    def wn(self, sd, prefix: str, x, mask, n_layers: int, hidden: int, kernel: int = 5) -> np.ndarray:
        """modules.WN.forward with g=None, dilation_rate=1 (modules.py:148-176)."""
        x = self.arr(x)
        mask = self.arr(mask)
        out = np.zeros_like(x)
        for i in range(n_layers):
            w_in, b_in = self.conv_w(sd, f"{prefix}.in_layers.{i}")
            a = self.conv1d(x, w_in, b_in, dilation=1, padding=(kernel - 1) // 2)
            acts = self.gate(a)
            w_rs, b_rs = self.conv_w(sd, f"{prefix}.res_skip_layers.{i}")
            rs = self.conv1d(acts, w_rs, b_rs)
            if i < n_layers - 1:
                x = (x + rs[:, :hidden]) * mask
                out = out + rs[:, hidden:]
            else:
                out = out + rs
        return out * mask

    def mel_encoder(self, sd, d, mel, lengths):
        """MelEncoder.forward (models.py:35-47) -> (x, m, logs, mask)."""
        mel = self.arr(mel)
        w, b = self.conv_w(sd, "enc_p.pre_enc")
        x = self.conv1d(mel, w, b)
        mask = self.arr(self.sequence_mask(lengths, mel.shape[2]))
        x = self.wn(sd, "enc_p.encoder", x * mask, mask, d.enc_layers, d.hidden_channels, d.wn_kernel)
        w, b = self.conv_w(sd, "enc_p.proj")
        stats = self.conv1d(x, w, b) * mask
        C = d.inter_channels
        return x, np.ascontiguousarray(stats[:, :C]), np.ascontiguousarray(stats[:, C:]), mask

    def coupling_reverse(self, sd, prefix: str, d, x, mask) -> np.ndarray:
        """ResidualCouplingLayer.forward(reverse=True), mean_only (modules.py:324-343)."""
        half = d.half
        x0, x1 = x[:, :half], x[:, half:]
        w, b = self.conv_w(sd, prefix + ".pre")
        h = self.conv1d(x0, w, b) * mask
        h = self.wn(sd, prefix + ".enc", h, mask, d.flow_layers, d.hidden_channels, d.wn_kernel)
        w, b = self.conv_w(sd, prefix + ".post")
        m = self.conv1d(h, w, b) * mask
        x1 = (x1 - m) * self.dtype.type(1.0) * mask  # exp(-logs) == exp(-0) == 1
        return np.ascontiguousarray(np.concatenate([x0, x1], axis=1))

    def flow_reverse(self, sd, d, z, mask, trace: Optional[List[np.ndarray]] = None) -> np.ndarray:
        """ResidualCouplingBlock.forward(reverse=True) (models.py:77-79): Flip, RCL3, Flip, RCL2, ..."""
        for f in reversed(range(d.n_flows)):
            z = self.flip(z)
            z = self.coupling_reverse(sd, f"flow.flows.{2 * f}", d, z, mask)
            if trace is not None:
                trace.append(z)
        return z

    def posterior_encoder(self, sd, d, spec, lengths, eps):
        """PosteriorEncoder.forward with g=None (models.py:103-110) -> (z, m, logs, mask); ``eps`` is the randn_like draw."""
        spec = self.arr(spec)
        mask = self.arr(self.sequence_mask(lengths, spec.shape[2]))
        w, b = self.conv_w(sd, "enc_q.pre")
        x = self.conv1d(spec, w, b) * mask
        x = self.wn(sd, "enc_q.enc", x, mask, 16, d.hidden_channels, 5)  # hard-coded 5, 1, 16 (models.py:312)
        w, b = self.conv_w(sd, "enc_q.proj")
        stats = self.conv1d(x, w, b) * mask
        C = d.inter_channels
        m, logs = np.ascontiguousarray(stats[:, :C]), np.ascontiguousarray(stats[:, C:])
        z = (m + self.arr(eps) * np.exp(logs)) * mask
        return z, m, logs, mask

    def coupling_forward(self, sd, prefix: str, d, x, mask) -> np.ndarray:
        """ResidualCouplingLayer.forward(reverse=False), mean_only (modules.py:324-339): x1 = m + x1 * exp(0) * mask."""
        half = d.half
        x0, x1 = x[:, :half], x[:, half:]
        w, b = self.conv_w(sd, prefix + ".pre")
        h = self.conv1d(x0, w, b) * mask
        h = self.wn(sd, prefix + ".enc", h, mask, d.flow_layers, d.hidden_channels, d.wn_kernel)
        w, b = self.conv_w(sd, prefix + ".post")
        m = self.conv1d(h, w, b) * mask
        x1 = m + x1 * self.dtype.type(1.0) * mask
        return np.ascontiguousarray(np.concatenate([x0, x1], axis=1))

    def flow_forward(self, sd, d, z, mask) -> np.ndarray:
        """ResidualCouplingBlock.forward(reverse=False) (models.py:73-76): RCL0, Flip, RCL1, Flip, ..."""
        z = self.arr(z)
        for f in range(d.n_flows):
            z = self.coupling_forward(sd, f"flow.flows.{2 * f}", d, z, mask)
            z = self.flip(z)
        return z

    def resblock1(self, sd, prefix: str, x, kernel: int, dilations: Sequence[int]) -> np.ndarray:
        """ResBlock1.forward with x_mask=None (modules.py:210-223)."""
        for l, dil in enumerate(dilations):
            xt = self.leaky_relu(x, 0.1)
            w, b = self.conv_w(sd, f"{prefix}.convs1.{l}")
            xt = self.conv1d(xt, w, b, dilation=dil, padding=(kernel * dil - dil) // 2)
            xt = self.leaky_relu(xt, 0.1)
            w, b = self.conv_w(sd, f"{prefix}.convs2.{l}")
            xt = self.conv1d(xt, w, b, dilation=1, padding=(kernel - 1) // 2)
            x = xt + x
        return x

    def resblock2(self, sd, prefix: str, x, kernel: int, dilations: Sequence[int]) -> np.ndarray:
        """modules.ResBlock2.forward, x_mask=None (modules.py:243-252): x = conv_l(leaky_relu(x, 0.1)) + x for the two convs."""
        x = self.arr(x)
        for l in range(2):
            dil = int(dilations[l])
            w, b = self.conv_w(sd, f"{prefix}.convs.{l}")
            x = self.conv1d(self.leaky_relu(x, 0.1), w, b, dilation=dil, padding=(kernel * dil - dil) // 2) + x
        return x

    def generator(self, sd, d, z, trace: Optional[Dict[str, np.ndarray]] = None) -> np.ndarray:
        """Generator.forward with g=None (models.py:141-160)."""
        w, b = self.conv_w(sd, "dec.conv_pre")
        x = self.conv1d(self.arr(z), w, b, padding=3)
        if trace is not None:
            trace["conv_pre"] = x
        nk = len(d.resblock_kernel_sizes)
        for i, (u, k) in enumerate(zip(d.upsample_rates, d.upsample_kernel_sizes)):
            x = self.leaky_relu(x, 0.1)
            w, b = self.conv_w(sd, f"dec.ups.{i}")
            x = self.conv_transpose1d(x, w, b, stride=u, padding=(k - u) // 2)
            if trace is not None:
                trace[f"ups{i}"] = x
            xs = None
            for j, (rk, rd) in enumerate(zip(d.resblock_kernel_sizes, d.resblock_dilation_sizes)):
                block = self.resblock1 if str(getattr(d, "resblock", "1")) == "1" else self.resblock2  # models.py:121
                r = block(sd, f"dec.resblocks.{i * nk + j}", x, rk, rd)
                xs = r if xs is None else xs + r
            x = xs / self.dtype.type(nk)
            if trace is not None:
                trace[f"stage{i}"] = x
        x = self.leaky_relu(x, 0.01)  # F.leaky_relu default slope
        w, _ = self.conv_w(sd, "dec.conv_post")
        x = self.conv1d(x, w, None, padding=3)
        return np.tanh(x)

    def infer(self, sd, d, mel, lengths, eps, noise_scale: float = 1.0, max_len: Optional[int] = None):
        """SynthesizerTrn.infer (models.py:331-339) with the N(0,1) draw injected as ``eps``.

        Returns (o, x_mask, (z, z_p, m_p, logs_p)).
        """
        _, m_p, logs_p, mask = self.mel_encoder(sd, d, mel, lengths)
        z_p = m_p + self.arr(eps) * np.exp(logs_p) * self.dtype.type(noise_scale)
        z = self.flow_reverse(sd, d, z_p, mask)
        o = self.generator(sd, d, np.ascontiguousarray((z * mask)[:, :, :max_len]))
        return o, mask, (z, z_p, m_p, logs_p)
