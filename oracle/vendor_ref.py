"""TEST / BASELINE INFRASTRUCTURE ONLY -- recipe that makes the UNMODIFIED reference available on the GPU box.

    python oracle/vendor_ref.py            (build container only; also run by __graft_entry__.build())

/root/reference does not exist on the GPU box, and the reference is a flat directory of scripts that pip cannot
install (no setup.py / pyproject.toml).  This recipe copies, byte for byte, the four files the hot path lives in
(models.py, modules.py, commons.py, transforms.py -- SURVEY 8(a)) from where they lie under /root/reference into
oracle/_ref/, which is git-ignored (nothing of the reference enters the history) but not gpurun-ignored, so it
travels with the snapshot like a built .so.  A manifest with the files' sha256 is written beside them.

Users: bench.py's reference arm (`--impl reference`, `cpu_baseline.kind == "reference"`) and its `--impl torch-gpu`
arm, through oracle/ref_loader.py.  Never imported by the product path.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ("models.py", "modules.py", "commons.py", "transforms.py")


def vendor(src: str = None) -> bool:
    src = src or os.environ.get("SVK_REFERENCE", "/root/reference")
    if not all(os.path.isfile(os.path.join(src, f)) for f in FILES):
        return False
    os.makedirs(DEST, exist_ok=True)
    manifest = {"source": src, "files": {}}
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(DEST, f))
        manifest["files"][f] = hashlib.sha256(open(os.path.join(DEST, f), "rb").read()).hexdigest()
    json.dump(manifest, open(os.path.join(DEST, "MANIFEST.json"), "w"), indent=1)
    return True


if __name__ == "__main__":
    ok = vendor(sys.argv[1] if len(sys.argv) > 1 else None)
    print("vendored into", DEST if ok else "(nothing: reference checkout not found)")
