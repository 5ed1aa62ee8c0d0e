"""CPU-side checks: the C-ABI library loads and exports every symbol include/svk.h declares, the
checkpoint key surface matches the reference, and the shim refuses to compute without CUDA."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from conftest import GOLDEN, ROOT

import svk_runtime as rt
import svk_weights as W


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "svk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svk_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build_libsvk()
    return rt.lib()


def test_header_symbols_exported(lib):
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/svk.h but not exported by libsvk.so"
    assert set(names) == set(rt.SIGNATURES), "ctypes binding and header disagree"
    assert lib.svk_abi_version() == 2


def test_config_struct_layout():
    n_i32 = 11 + 8 + 8 + 1 + 8 + 8 * 3 + 1 + 1  # + resblock_type (ABI 2)
    assert ctypes.sizeof(rt.SvkConfig) == 4 * n_i32


def test_key_surface_matches_reference(base_dims):
    ref = json.load(open(os.path.join(GOLDEN, "state_dict_keys_iitp_base.json")))
    spec = W.state_dict_spec(base_dims)
    assert [[k, list(s)] for k, s in spec] == ref
    assert len(spec) == 659
    live = [k for k, _ in spec if not W.is_dead_key(k)]
    assert sum(int(np.prod(s)) for k, s in spec if not W.is_dead_key(k)) == 35_696_448
    assert len(live) == 659 - 103 - 3 * 5 - 2


def test_ms_config_is_same_model():
    a = json.load(open(os.path.join(ROOT, "configs", "iitp_base.json")))
    b = json.load(open(os.path.join(ROOT, "configs", "iitp_base_ms.json")))
    assert a["model"] == b["model"] and a["train"] == b["train"]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu(lib, base_dims):
    with pytest.raises(rt.SvkError) as ei:
        rt.Handle(base_dims, 0)
    assert ei.value.code == rt.SVK_ERR_CUDA
    assert "no CPU path" in str(ei.value)


def test_create_rejects_bad_config(lib, base_dims):
    import copy
    d = copy.copy(base_dims)
    d.inter_channels = 191
    with pytest.raises(rt.SvkError) as ei:
        rt.Handle(d, 0)
    assert ei.value.code == rt.SVK_ERR_INVALID and "divisible by 2" in str(ei.value)


def test_stateless_ops_validate_arguments(lib):
    assert lib.svk_conv1d(None, 1, 8, 4, None, None, 8, 3, 1, 1, 1.0, None, None) == rt.SVK_ERR_INVALID
    assert b"svk_conv1d" in lib.svk_last_error()
    assert lib.svk_rq_spline(1, 1, 1, 1, 1, 10, 0, 5.0, 0.2, 1e-3, 1e-3, 1, 1, None, None) == rt.SVK_ERR_INVALID
    assert b"Minimal bin width too large" in lib.svk_last_error()


def test_shim_surface_on_cpu(base_cfg, base_sd):
    from models import SynthesizerTrn
    net = SynthesizerTrn(513, 32, n_speakers=109, **base_cfg["model"])
    sd = net.state_dict()
    ref = json.load(open(os.path.join(GOLDEN, "state_dict_keys_iitp_base.json")))
    assert [[k, list(v.shape)] for k, v in sd.items()] == ref
    # zero-init post like the reference (modules.py:321-322)
    assert float(sd["flow.flows.0.post.weight"].abs().max()) == 0.0
    # utils.load_checkpoint protocol (utils.py:31-39): take from file, fall back to model's own
    saved = {k: torch.from_numpy(v) for k, v in base_sd.items() if not k.startswith("dec.conv_post")}
    new = {k: saved.get(k, v) for k, v in sd.items()}
    res = net.load_state_dict(new)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(net.state_dict()["dec.ups.0.weight_v"], torch.from_numpy(base_sd["dec.ups.0.weight_v"]))
    with pytest.raises(RuntimeError, match="Missing key"):
        net.load_state_dict({k: v for k, v in new.items() if k != "enc_p.proj.bias"})
    with pytest.raises(RuntimeError, match="size mismatch"):
        bad = dict(new)
        bad["enc_p.proj.bias"] = torch.zeros(3)
        net.load_state_dict(bad)
    res = net.load_state_dict({"enc_p.proj.bias": new["enc_p.proj.bias"], "bogus": torch.zeros(1)}, strict=False)
    assert "bogus" in res.unexpected_keys and len(res.missing_keys) == 658
    assert net.eval() is net
    with pytest.raises(RuntimeError, match="no CPU path"):
        net.infer(torch.zeros(1, 80, 4), torch.tensor([4]))
    with pytest.raises(NotImplementedError):
        net(None, None, None, None)
    with pytest.raises(AttributeError, match="emb_g"):
        net.voice_conversion(None, None, 0, 1)


def test_shim_ctor_asserts(base_cfg):
    from models import SynthesizerTrn
    m = dict(base_cfg["model"])
    m["inter_channels"] = 191
    with pytest.raises(AssertionError, match="divisible by 2"):
        SynthesizerTrn(513, 32, **m)
    m = dict(base_cfg["model"])
    m["resblock"] = "2"  # models.py:121: ResBlock2 takes the first two dilations of each triple (modules.py:232-241)
    keys = SynthesizerTrn(513, 32, **m).state_dict().keys()
    assert "dec.resblocks.0.convs.1.weight_v" in keys and "dec.resblocks.0.convs1.0.weight_v" not in keys
    assert not any(".convs.2." in k for k in keys if k.startswith("dec.resblocks"))


def test_clip_len_matches_python_slicing():
    from models import SynthesizerTrn
    for T in (1, 5, 12):
        for ml in (None, 0, 1, 4, 12, 100, -1, -3, -50):
            assert SynthesizerTrn._clip_len(T, ml) == len(list(range(T))[:ml])


def test_shim_signatures_match_the_reference_modules():
    """the nn.Module surface the callers use -- constructor, infer, and the sub-module calls -- has the
    reference's parameter names, order and defaults (recorded from the reference by make_golden_signatures.py)."""
    import inspect
    from models import SynthesizerTrn
    ref = json.load(open(os.path.join(GOLDEN, "models_signatures.json")))
    for name in ("__init__", "infer", "forward", "voice_conversion"):
        assert str(inspect.signature(getattr(SynthesizerTrn, name))) == ref["SynthesizerTrn." + name], name
    sub = {"Generator.forward": "_dec_forward", "MelEncoder.forward": "_enc_p_forward",
           "PosteriorEncoder.forward": "_enc_q_forward", "ResidualCouplingBlock.forward": "_flow_forward"}
    for ref_name, ours in sub.items():
        assert str(inspect.signature(getattr(SynthesizerTrn, ours))) == ref[ref_name], ref_name
