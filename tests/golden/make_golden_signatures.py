"""Record the call signatures of the reference's module surface (the drop-in boundary of SURVEY 8b) from the reference itself.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_signatures.py   ->  tests/golden/models_signatures.json
"""
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get("SVK_REFERENCE", "/root/reference"))
import models as ref  # noqa: E402  (reference, imported in place)

out = {
    "SynthesizerTrn.__init__": str(inspect.signature(ref.SynthesizerTrn.__init__)),
    "SynthesizerTrn.infer": str(inspect.signature(ref.SynthesizerTrn.infer)),
    "SynthesizerTrn.forward": str(inspect.signature(ref.SynthesizerTrn.forward)),
    "SynthesizerTrn.voice_conversion": str(inspect.signature(ref.SynthesizerTrn.voice_conversion)),
    "Generator.forward": str(inspect.signature(ref.Generator.forward)),
    "MelEncoder.forward": str(inspect.signature(ref.MelEncoder.forward)),
    "PosteriorEncoder.forward": str(inspect.signature(ref.PosteriorEncoder.forward)),
    "ResidualCouplingBlock.forward": str(inspect.signature(ref.ResidualCouplingBlock.forward)),
}
json.dump(out, open(os.path.join(HERE, "models_signatures.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
