"""Checkpoint surface of the mel->waveform path: key names, shapes, synthetic weights.

The reference has no weight file format of its own; weights enter through
``utils.load_checkpoint`` (reference utils.py:18-43) which walks the model's
``state_dict()`` keys.  This module restates that key surface (SURVEY App. C,
659 keys for iitp_base.json) from the hyper-parameters alone, so the B200 shim
can answer ``state_dict()`` / ``load_state_dict()`` without building any
``nn.Conv1d``.

It also holds the seeded "alive" weight recipe used by tests and bench.py
(reference random init makes the flow an exact identity and the
waveform ~0.01 in amplitude, which makes parity vacuous).  The recipe is keyed
per tensor name, so it does not depend on construction order, and uses numpy's
Philox bit generator, which is stable across numpy versions and hosts.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np

# Hard-coded by the reference constructor, not read from the JSON
# (reference models.py:305-314).
ENC_LAYERS = 16
FLOW_LAYERS = 8
N_FLOWS = 4
WN_KERNEL = 5
N_MEL = 80
CONV_PRE_KERNEL = 7
CONV_POST_KERNEL = 7


@dataclass
class ModelDims:
    """Effective hyper-parameters of SynthesizerTrn (reference models.py:266-314)."""

    spec_channels: int = 513
    inter_channels: int = 192
    hidden_channels: int = 192
    resblock: str = "1"
    resblock_kernel_sizes: Sequence[int] = (3, 7, 11)
    resblock_dilation_sizes: Sequence[Sequence[int]] = ((1, 3, 5), (1, 3, 5), (1, 3, 5))
    upsample_rates: Sequence[int] = (8, 8, 2, 2)
    upsample_initial_channel: int = 512
    upsample_kernel_sizes: Sequence[int] = (16, 16, 4, 4)
    gin_channels: int = 0
    n_mel: int = N_MEL
    enc_layers: int = ENC_LAYERS
    flow_layers: int = FLOW_LAYERS
    n_flows: int = N_FLOWS
    wn_kernel: int = WN_KERNEL

    @property
    def hop(self) -> int:
        h = 1
        for u in self.upsample_rates:
            h *= int(u)
        return h

    @property
    def half(self) -> int:
        return self.inter_channels // 2

    def stage_channels(self, i: int) -> int:
        return self.upsample_initial_channel // (2 ** (i + 1))

    def validate(self) -> None:
        # reference modules.py:308 / modules.py:114 / models.py:122
        assert self.inter_channels % 2 == 0, "channels should be divisible by 2"
        assert self.wn_kernel % 2 == 1
        # models.py:121: '1' -> ResBlock1, anything else -> ResBlock2 (two convs: dilation[0], dilation[1])
        for ds in self.resblock_dilation_sizes:
            assert len(ds) == 3 if str(self.resblock) == "1" else 2 <= len(ds) <= 3, "ResBlock1 takes 3 dilations, ResBlock2 2"
        assert len(self.upsample_rates) == len(self.upsample_kernel_sizes)
        assert len(self.resblock_kernel_sizes) == len(self.resblock_dilation_sizes)
        for k in self.resblock_kernel_sizes:
            assert k % 2 == 1
        for u, k in zip(self.upsample_rates, self.upsample_kernel_sizes):
            assert (k - u) % 2 == 0, "ConvTranspose1d padding (k-u)//2 must be exact"


def dims_from_model_kwargs(spec_channels: int, **model) -> ModelDims:
    """Map the ``hps.model`` block of iitp_base.json to ModelDims (ignored keys: SURVEY F6)."""
    return ModelDims(
        spec_channels=int(spec_channels),
        inter_channels=int(model["inter_channels"]),
        hidden_channels=int(model["hidden_channels"]),
        resblock=str(model["resblock"]),
        resblock_kernel_sizes=tuple(int(k) for k in model["resblock_kernel_sizes"]),
        resblock_dilation_sizes=tuple(tuple(int(d) for d in ds) for ds in model["resblock_dilation_sizes"]),
        upsample_rates=tuple(int(u) for u in model["upsample_rates"]),
        upsample_initial_channel=int(model["upsample_initial_channel"]),
        upsample_kernel_sizes=tuple(int(k) for k in model["upsample_kernel_sizes"]),
        gin_channels=int(model.get("gin_channels", 0)),
    )


def _wn_keys(prefix: str, hidden: int, kernel: int, n_layers: int, gin: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """modules.WN parameter surface (reference modules.py:111-146), legacy weight_norm names."""
    out: List[Tuple[str, Tuple[int, ...]]] = []
    for i in range(n_layers):
        p = f"{prefix}.in_layers.{i}"
        out += [(p + ".bias", (2 * hidden,)), (p + ".weight_g", (2 * hidden, 1, 1)),
                (p + ".weight_v", (2 * hidden, hidden, kernel))]
    for i in range(n_layers):
        rs = 2 * hidden if i < n_layers - 1 else hidden
        p = f"{prefix}.res_skip_layers.{i}"
        out += [(p + ".bias", (rs,)), (p + ".weight_g", (rs, 1, 1)), (p + ".weight_v", (rs, hidden, 1))]
    if gin != 0:
        p = f"{prefix}.cond_layer"
        c = 2 * hidden * n_layers
        out += [(p + ".bias", (c,)), (p + ".weight_g", (c, 1, 1)), (p + ".weight_v", (c, gin, 1))]
    return out


def state_dict_spec(d: ModelDims) -> List[Tuple[str, Tuple[int, ...]]]:
    """Ordered (key, shape) list identical to ``SynthesizerTrn(...).state_dict()`` of the reference."""
    H, C, gin = d.hidden_channels, d.inter_channels, d.gin_channels
    spec: List[Tuple[str, Tuple[int, ...]]] = []
    # enc_p : MelEncoder (models.py:15-33)
    spec += _wn_keys("enc_p.encoder", H, d.wn_kernel, d.enc_layers, gin)
    spec += [("enc_p.pre_enc.weight", (H, d.n_mel, 1)), ("enc_p.pre_enc.bias", (H,)),
             ("enc_p.proj.weight", (2 * C, H, 1)), ("enc_p.proj.bias", (2 * C,))]
    # dec : Generator (models.py:116-139)
    U = d.upsample_initial_channel
    spec += [("dec.conv_pre.weight", (U, C, CONV_PRE_KERNEL)), ("dec.conv_pre.bias", (U,))]
    for i, (u, k) in enumerate(zip(d.upsample_rates, d.upsample_kernel_sizes)):
        cin, cout = U // (2 ** i), U // (2 ** (i + 1))
        p = f"dec.ups.{i}"
        spec += [(p + ".bias", (cout,)), (p + ".weight_g", (cin, 1, 1)), (p + ".weight_v", (cin, cout, k))]
    nk = len(d.resblock_kernel_sizes)
    for i in range(len(d.upsample_rates)):
        ch = d.stage_channels(i)
        for j, k in enumerate(d.resblock_kernel_sizes):
            n = i * nk + j
            if str(d.resblock) != "1":  # ResBlock2 (modules.py:232-241): one ModuleList `convs` of two layers
                for l in range(2):
                    p = f"dec.resblocks.{n}.convs.{l}"
                    spec += [(p + ".bias", (ch,)), (p + ".weight_g", (ch, 1, 1)), (p + ".weight_v", (ch, ch, k))]
                continue
            for grp in ("convs1", "convs2"):
                for l in range(3):
                    p = f"dec.resblocks.{n}.{grp}.{l}"
                    spec += [(p + ".bias", (ch,)), (p + ".weight_g", (ch, 1, 1)), (p + ".weight_v", (ch, ch, k))]
    ch = d.stage_channels(len(d.upsample_rates) - 1)
    spec += [("dec.conv_post.weight", (1, ch, CONV_POST_KERNEL))]
    if gin != 0:
        spec += [("dec.cond.weight", (U, gin, 1)), ("dec.cond.bias", (U,))]
    # enc_q : PosteriorEncoder (models.py:83-103) -- dead at inference, must be accepted
    spec += [("enc_q.pre.weight", (H, d.spec_channels, 1)), ("enc_q.pre.bias", (H,))]
    spec += _wn_keys("enc_q.enc", H, 5, 16, gin)
    spec += [("enc_q.proj.weight", (2 * C, H, 1)), ("enc_q.proj.bias", (2 * C,))]
    # flow : ResidualCouplingBlock (models.py:50-71); odd indices are Flip (no params)
    for f in range(d.n_flows):
        p = f"flow.flows.{2 * f}"
        spec += [(p + ".pre.weight", (H, d.half, 1)), (p + ".pre.bias", (H,))]
        spec += _wn_keys(p + ".enc", H, d.wn_kernel, d.flow_layers, gin)
        spec += [(p + ".post.weight", (d.half, H, 1)), (p + ".post.bias", (d.half,))]
    return spec


def is_dead_key(key: str) -> bool:
    """Keys the inference path never reads."""
    return key.startswith("enc_q.") or ".cond_layer." in key or key.startswith("dec.cond.")


def is_posterior_key(key: str) -> bool:
    """enc_q.* without its (gin-only) cond_layer: read only by the analysis direction (PosteriorEncoder, models.py:83-110)."""
    return key.startswith("enc_q.") and ".cond_layer." not in key


def _rng_for(seed: int, key: str) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[int(seed), zlib.crc32(key.encode())]))


DEC_GAIN = 2.4      # SURVEY 8(c) "alive" recipe: every dec.*weight_g scaled so |o| max ~0.9
POST_STD = 0.05     # flow.flows.*.post is zero-init in the reference -> randomise


def make_state_dict(d: ModelDims, seed: int = 1234, alive: bool = True,
                    include_dead: bool = True) -> "OrderedDict[str, np.ndarray]":
    """Seeded fp32 weights with the reference key surface.

    weight_v / plain weights ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (torch's conv default scale),
    weight_g = ||v|| over dims != 0 (what weight_norm starts from), times DEC_GAIN for ``dec.``
    when ``alive``; ``flow.*.post`` ~ N(0, POST_STD) when ``alive`` else zero.
    """
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    spec = state_dict_spec(d)
    shapes = dict(spec)
    for key, shape in spec:
        if is_dead_key(key) and not include_dead:
            continue
        rng = _rng_for(seed, key)
        leaf = key.rsplit(".", 1)[1]
        if leaf == "weight_g":
            continue  # filled after its weight_v
        if leaf in ("weight_v", "weight"):
            fan_in = int(np.prod(shape[1:]))
            bound = 1.0 / np.sqrt(fan_in)
            if ".post." in key and key.startswith("flow."):
                w = (rng.standard_normal(shape) * POST_STD if alive else np.zeros(shape)).astype(np.float32)
            else:
                w = rng.uniform(-bound, bound, size=shape).astype(np.float32)
            sd[key] = w
            if leaf == "weight_v":
                gkey = key[:-1] + "g"
                norm = np.sqrt((w.astype(np.float64) ** 2).reshape(shape[0], -1).sum(1)).astype(np.float32)
                gain = DEC_GAIN if (alive and key.startswith("dec.")) else 1.0
                sd[gkey] = (norm * np.float32(gain)).reshape(shapes[gkey]).astype(np.float32)
        elif leaf == "bias":
            wkey = key[:-4] + ("weight_v" if (key[:-4] + "weight_v") in shapes else "weight")
            fan_in = int(np.prod(shapes[wkey][1:]))
            # ConvTranspose1d: torch computes fan_in from dim 1 * k as well
            bound = 1.0 / np.sqrt(fan_in)
            if ".post." in key and key.startswith("flow."):
                b = (rng.standard_normal(shape) * POST_STD if alive else np.zeros(shape)).astype(np.float32)
            else:
                b = rng.uniform(-bound, bound, size=shape).astype(np.float32)
            sd[key] = b
        else:  # pragma: no cover
            raise KeyError(key)
    # restore reference ordering
    ordered: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for key, _ in spec:
        if key in sd:
            ordered[key] = sd[key]
    return ordered


def state_dict_checksum(sd: Dict[str, np.ndarray]) -> str:
    """Order-independent checksum used to prove the GPU box regenerated identical weights."""
    acc = 0
    for k in sorted(sd):
        acc = zlib.crc32(np.ascontiguousarray(sd[k]).tobytes(), zlib.crc32(k.encode(), acc))
    return f"{acc:08x}"


# Tiny config for fast CPU tests of the oracle / host logic (same topology, small widths).
TINY_MODEL = dict(
    inter_channels=16, hidden_channels=16, filter_channels=32, n_heads=2, n_layers=6, kernel_size=3,
    p_dropout=0.1, resblock="1", resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], upsample_rates=[8, 8, 2, 2],
    upsample_initial_channel=64, upsample_kernel_sizes=[16, 16, 4, 4], n_layers_q=3,
    use_spectral_norm=False, gin_channels=8,
)
