"""ctypes binding of libsvk.so (include/svk.h).  No compute happens in Python.

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  Importing this module
never falls back to another implementation: if libsvk.so is missing or a CUDA device is absent,
calls raise ``SvkError``.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsvk.so")

SVK_OK = 0
SVK_IGNORED = 1
SVK_ERR_INVALID = -1
SVK_ERR_CUDA = -2
SVK_ERR_UNKNOWN_KEY = -3
SVK_ERR_STATE = -4
SVK_ERR_WORKSPACE = -5
SVK_ERR_RANGE = -6

SVK_MAX_UPSAMPLES = 8
SVK_MAX_RESBLOCK_KERNELS = 8
SVK_RESBLOCK_PAIRS = 3

PRECISION_FP32 = 0   # fp32 FFMA everywhere
PRECISION_TC = 1     # tcgen05 3-product fp16 split (fp32-class results), the default engine
PRECISION_BF16 = 2   # bf16 operand images + weights, one tcgen05 pass, fp32 accumulate (BASELINE configs[3])
PRECISIONS = {"fp32": PRECISION_FP32, "ffma": PRECISION_FP32, "tc": PRECISION_TC, "bf16": PRECISION_BF16}


class SvkError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libsvk error {code}: {message}")
        self.code = code


class SvkConfig(ctypes.Structure):
    _fields_ = [
        ("n_mel", ctypes.c_int32),
        ("spec_channels", ctypes.c_int32),
        ("inter_channels", ctypes.c_int32),
        ("hidden_channels", ctypes.c_int32),
        ("enc_layers", ctypes.c_int32),
        ("flow_layers", ctypes.c_int32),
        ("n_flows", ctypes.c_int32),
        ("wn_kernel", ctypes.c_int32),
        ("gin_channels", ctypes.c_int32),
        ("upsample_initial_channel", ctypes.c_int32),
        ("n_upsamples", ctypes.c_int32),
        ("upsample_rates", ctypes.c_int32 * SVK_MAX_UPSAMPLES),
        ("upsample_kernel_sizes", ctypes.c_int32 * SVK_MAX_UPSAMPLES),
        ("n_resblock_kernels", ctypes.c_int32),
        ("resblock_kernel_sizes", ctypes.c_int32 * SVK_MAX_RESBLOCK_KERNELS),
        ("resblock_dilations", (ctypes.c_int32 * SVK_RESBLOCK_PAIRS) * SVK_MAX_RESBLOCK_KERNELS),
        ("precision", ctypes.c_int32),
        ("resblock_type", ctypes.c_int32),
    ]


class SvkLaunchRecord(ctypes.Structure):
    _fields_ = [("layer", ctypes.c_int32), ("cin", ctypes.c_int32), ("cout", ctypes.c_int32), ("k", ctypes.c_int32),
                ("dilation", ctypes.c_int32), ("batch", ctypes.c_int32), ("length", ctypes.c_int64),
                ("flops", ctypes.c_double), ("bytes", ctypes.c_double), ("ms", ctypes.c_float),
                ("engine", ctypes.c_int32), ("gap_ms", ctypes.c_float), ("reserved", ctypes.c_int32),
                ("dup_bytes", ctypes.c_double)]


class SvkConvFlowWeights(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("pre_w", "pre_b", "sep_w", "sep_b", "pw_w", "pw_b", "norm1_g", "norm1_b",
                                               "norm2_g", "norm2_b", "proj_w", "proj_b")]


class SvkGraphIO(ctypes.Structure):
    _fields_ = [("mel", ctypes.c_void_p), ("lengths", ctypes.c_void_p), ("eps", ctypes.c_void_p), ("o", ctypes.c_void_p),
                ("x_mask", ctypes.c_void_p), ("z", ctypes.c_void_p), ("z_p", ctypes.c_void_p), ("m_p", ctypes.c_void_p),
                ("logs_p", ctypes.c_void_p), ("B", ctypes.c_int32), ("T", ctypes.c_int32), ("T_out", ctypes.c_int32),
                ("programmatic_edges", ctypes.c_int32), ("kernel_nodes", ctypes.c_int64)]


LAYER_NAMES = {0: "other", 1: "pre_enc", 2: "wn_in", 3: "wn_res_skip", 4: "proj", 5: "flow_pre", 6: "flow_post",
               7: "conv_pre", 8: "upsample", 9: "resblock_conv1", 10: "resblock_conv2", 11: "conv_post", 12: "resblock_pair",
               13: "split_image", 14: "wn_layer"}

_lib: Optional[ctypes.CDLL] = None

_vp, _i, _i64, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); mirrors include/svk.h one to one (tests check every symbol loads)
SIGNATURES = {
    "svk_abi_version": (_i, []),
    "svk_last_error": (ctypes.c_char_p, []),
    "svk_create": (_i, [ctypes.POINTER(SvkConfig), _i, ctypes.POINTER(_vp)]),
    "svk_destroy": (None, [_vp]),
    "svk_load_tensor": (_i, [_vp, ctypes.c_char_p, _vp, ctypes.POINTER(_i64), _i]),
    "svk_finalize_weights": (_i, [_vp]),
    "svk_weight_status": (_i, [_vp, ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "svk_workspace_bytes": (_sz, [_vp, _i, _i, _i]),
    "svk_infer": (_i, [_vp, _vp, _vp, _vp, _f, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "svk_infer_host": (_i, [_vp, _vp, _vp, _vp, _f, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "svk_check_range": (_i, [_vp, _vp]),
    "svk_get_config": (_i, [_vp, ctypes.POINTER(SvkConfig)]),
    "svk_graph_create": (_i, [_vp, _i, _i, _i, _f, ctypes.POINTER(_vp)]),
    "svk_graph_buffers": (_i, [_vp, _vp]),
    "svk_graph_launch": (_i, [_vp, _vp]),
    "svk_graph_destroy": (None, [_vp]),
    "svk_pipeline_create": (_i, [_vp, _i, _i, _i, _i, ctypes.POINTER(_vp)]),
    "svk_pipeline_submit": (_i, [_vp, _vp, _vp, _vp, ctypes.c_uint64, _f, _vp, _vp, ctypes.POINTER(_i64)]),
    "svk_pipeline_wait": (_i, [_vp, _i64]),
    "svk_pipeline_drain": (_i, [_vp]),
    "svk_pipeline_destroy": (None, [_vp]),
    "svk_randn": (_i, [_vp, ctypes.c_uint64, ctypes.c_uint64, _i64, _vp, _vp]),
    "svk_halo_frames": (_i, [_vp]),
    "svk_window_workspace_bytes": (_sz, [_vp, _i, _i]),
    "svk_infer_window": (_i, [_vp, _vp, _vp, _vp, _f, _i, _i, _i, _i, _i, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "svk_infer_chunked": (_i, [_vp, _vp, _vp, _vp, _f, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "svk_last_launch_count": (_i64, [_vp]),
    "svk_profile_begin": (_i, [_vp, _i]),
    "svk_profile_end": (_i, [_vp, _vp, _i, ctypes.POINTER(_i)]),
    "svk_mel_encoder": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "svk_flow_reverse": (_i, [_vp, _vp, _vp, _i, _i, _vp, _sz, _vp]),
    "svk_flow_forward": (_i, [_vp, _vp, _vp, _i, _i, _vp, _sz, _vp]),
    "svk_posterior_encoder": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "svk_generator": (_i, [_vp, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "svk_resblock1_workspace_bytes": (_sz, [_vp, _i, _i, _i]),
    "svk_resblock1": (_i, [_vp, _i, _vp, _i, _i, _vp, _vp, _sz, _vp]),
    "svk_conv1d": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp]),
    "svk_conv_transpose1d": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp]),
    "svk_conv1d_tc": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp]),
    "svk_conv_transpose1d_tc": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp]),
    "svk_sequence_mask": (_i, [_vp, _i, _i, _vp, _vp]),
    "svk_pcm_to_int16": (_i, [_vp, _i64, _f, _vp, _vp]),
    "svk_flip": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "svk_weight_norm": (_i, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "svk_rq_spline": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _f, _f, _f, _f, _vp, _vp, _vp, _vp]),
    "svk_convflow_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "svk_convflow": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, ctypes.POINTER(SvkConvFlowWeights), _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "svk_frontend_create": (_i, [_i, _i, _i, _i, _i, _f, _f, _i, ctypes.POINTER(_vp)]),
    "svk_frontend_destroy": (None, [_vp]),
    "svk_frontend_frames": (_i64, [_vp, _i64]),
    "svk_spectrogram": (_i, [_vp, _vp, _i, _i64, _vp, _vp]),
    "svk_spec_to_mel": (_i, [_vp, _vp, _i, _i64, _vp, _vp]),
    "svk_mel_spectrogram": (_i, [_vp, _vp, _i, _i64, _vp, _vp, _vp]),
    "svk_mel_basis": (_i, [_i, _i, _i, _f, _f, _vp]),
    "svk_hann_window": (_i, [_i, _i, _vp]),
}


def lib() -> ctypes.CDLL:
    """Load libsvk.so; raises if the CUDA extension has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SvkError(SVK_ERR_STATE, f"{LIB_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.svk_abi_version() != 2:
            raise SvkError(SVK_ERR_STATE, "libsvk ABI version mismatch")
        _lib = l
    return _lib


def check(status: int) -> int:
    if status < 0:
        raise SvkError(status, lib().svk_last_error().decode("utf-8", "replace"))
    return status


def make_config(dims, precision: int = PRECISION_TC) -> SvkConfig:
    """ModelDims (svk_weights.py) -> svk_config."""
    c = SvkConfig()
    c.n_mel = dims.n_mel
    c.spec_channels = dims.spec_channels
    c.inter_channels = dims.inter_channels
    c.hidden_channels = dims.hidden_channels
    c.enc_layers = dims.enc_layers
    c.flow_layers = dims.flow_layers
    c.n_flows = dims.n_flows
    c.wn_kernel = dims.wn_kernel
    c.gin_channels = dims.gin_channels
    c.upsample_initial_channel = dims.upsample_initial_channel
    if len(dims.upsample_rates) > SVK_MAX_UPSAMPLES or len(dims.resblock_kernel_sizes) > SVK_MAX_RESBLOCK_KERNELS:
        raise SvkError(SVK_ERR_INVALID, "too many upsample stages / resblock kernels")
    c.n_upsamples = len(dims.upsample_rates)
    for i, (u, k) in enumerate(zip(dims.upsample_rates, dims.upsample_kernel_sizes)):
        c.upsample_rates[i] = int(u)
        c.upsample_kernel_sizes[i] = int(k)
    c.n_resblock_kernels = len(dims.resblock_kernel_sizes)
    for j, (k, ds) in enumerate(zip(dims.resblock_kernel_sizes, dims.resblock_dilation_sizes)):
        c.resblock_kernel_sizes[j] = int(k)
        need = 2 if str(getattr(dims, "resblock", "1")) == "2" else SVK_RESBLOCK_PAIRS
        if (need == SVK_RESBLOCK_PAIRS and len(ds) != need) or len(ds) < need or len(ds) > SVK_RESBLOCK_PAIRS:
            raise SvkError(SVK_ERR_INVALID, "ResBlock1 needs exactly 3 dilations per kernel size, ResBlock2 uses the first 2")
        for l, dil in enumerate(ds):
            c.resblock_dilations[j][l] = int(dil)
        for l in range(len(ds), SVK_RESBLOCK_PAIRS):
            c.resblock_dilations[j][l] = 1
    c.precision = precision
    c.resblock_type = 2 if str(getattr(dims, "resblock", "1")) == "2" else 1
    return c


class Handle:
    """Owns one svk_handle (one device)."""

    def __init__(self, dims, device: int, precision: int = PRECISION_TC):
        self._h = _vp()
        self.cfg = make_config(dims, precision)
        check(lib().svk_create(ctypes.byref(self.cfg), int(device), ctypes.byref(self._h)))
        self.device = int(device)

    @property
    def ptr(self):
        return self._h

    def close(self):
        if self._h:
            lib().svk_destroy(self._h)
            self._h = _vp()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def load_tensor(self, key: str, array) -> int:
        """array: C-contiguous float32 numpy array (host)."""
        import numpy as np
        a = np.ascontiguousarray(array, dtype=np.float32)
        shape = (_i64 * a.ndim)(*a.shape)
        return check(lib().svk_load_tensor(self._h, key.encode(), a.ctypes.data_as(_vp), shape, a.ndim))

    def finalize(self):
        check(lib().svk_finalize_weights(self._h))

    def weight_status(self):
        a, b = _i(), _i()
        check(lib().svk_weight_status(self._h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def workspace_bytes(self, B: int, T: int, max_len: int) -> int:
        return int(lib().svk_workspace_bytes(self._h, B, T, max_len))

    def last_launch_count(self) -> int:
        return int(lib().svk_last_launch_count(self._h))

    def check_range(self, stream=None):
        """Synchronises `stream`; raises SvkError(SVK_ERR_RANGE) if a non-finite sample was produced since the last check."""
        check(lib().svk_check_range(self._h, stream))

    def profile_begin(self, max_records: int = 8192):
        self._prof_cap = int(max_records)
        check(lib().svk_profile_begin(self._h, self._prof_cap))

    def profile_end(self):
        """-> list of dicts (layer, cin, cout, k, dilation, batch, length, flops, bytes, ms, engine)."""
        buf = (SvkLaunchRecord * self._prof_cap)()
        n = _i()
        check(lib().svk_profile_end(self._h, buf, self._prof_cap, ctypes.byref(n)))
        out = []
        for r in buf[:min(n.value, self._prof_cap)]:
            out.append(dict(layer=LAYER_NAMES.get(r.layer, str(r.layer)), cin=r.cin, cout=r.cout, k=r.k,
                            dilation=r.dilation, batch=r.batch, length=r.length, flops=r.flops, bytes=r.bytes, ms=r.ms,
                            engine="tc" if r.engine else "ffma", gap_ms=r.gap_ms, dup_bytes=r.dup_bytes))
        return out
