"""TEST / BASELINE INFRASTRUCTURE ONLY -- import the unmodified reference vendored by oracle/vendor_ref.py.

The reference's module is called ``models`` like the drop-in shim, so it is loaded under the alias
``svk_ref_models``; its own bare imports (``commons``, ``modules``, ``transforms``) resolve through a sys.path
entry appended at the END, so nothing of the product package is shadowed.
"""
import importlib.util
import json
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "models.py")) and os.path.isfile(os.path.join(REF_DIR, "MANIFEST.json"))


def load_reference_models():
    """-> the reference's ``models`` module (alias svk_ref_models), or None when oracle/_ref/ was not vendored."""
    if not available():
        return None
    if "svk_ref_models" in sys.modules:
        return sys.modules["svk_ref_models"]
    if REF_DIR not in sys.path:
        sys.path.append(REF_DIR)
    spec = importlib.util.spec_from_file_location("svk_ref_models", os.path.join(REF_DIR, "models.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["svk_ref_models"] = mod
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    return mod


def manifest() -> dict:
    return json.load(open(os.path.join(REF_DIR, "MANIFEST.json"))) if available() else {}


def build_reference_net(cfg_model: dict, state_dict_np: dict, spec_channels: int = 513, segment_frames: int = 32,
                        n_speakers: int = 109):
    """The reference's SynthesizerTrn with the seeded weights loaded strictly, in eval mode (inference.ipynb cell 3)."""
    import torch
    ref = load_reference_models()
    if ref is None:
        return None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # legacy weight_norm FutureWarning
        net = ref.SynthesizerTrn(spec_channels, segment_frames, n_speakers=n_speakers, **cfg_model)
    net.eval()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in state_dict_np.items()}, strict=True)
    return net
