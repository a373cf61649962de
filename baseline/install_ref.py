"""Stage the UNMODIFIED reference model files for the GPU box (TEST / BENCH INFRASTRUCTURE, not product).

    python baseline/install_ref.py          # also run by __graft_entry__.build() when /root/reference is present

The reference is pure Python (no FFI, nothing to compile), and /root/reference does not exist on the GPU box.  This
copies the files of the masked-LM path -- model/{deberta,adapter,bert}.py, util/*.py and the main.py / mc.py loops --
byte for byte into baseline/_ref/ (git-ignored, so never part of the repo's history; NOT gpurun-ignored, so it travels with the
snapshot like a built .so).  oracle/ref_loader.py loads them by path under the version shims of SURVEY.md 8(c), which
lets ``bench.py --impl reference`` and ``cpu_baseline`` time the REAL reference implementation (kind "reference") on
the box's host cores; without baseline/_ref they fall back to the oracle port (kind "port").
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")   # baseline/_ref: the sanctioned, git-ignored home of the reference install
SRC = os.environ.get("FROZENBILM_REFERENCE_SRC", "/root/reference")
FILES = ["model/deberta.py", "model/adapter.py", "model/bert.py", "util/misc.py",
         # the training / evaluation loops that tests/test_dropin_gpu.py drives UNMODIFIED against this build
         "main.py", "mc.py", "args.py", "util/__init__.py", "util/dist.py", "util/metrics.py"]


def install(src=SRC, dest=DEST, quiet=False):
    if not os.path.isfile(os.path.join(src, FILES[0])):
        if not quiet:
            print(f"install_ref: {src} not present, nothing staged")
        return False
    manifest = {}
    for rel in FILES:
        d = os.path.join(dest, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), d)
        manifest[rel] = hashlib.sha256(open(d, "rb").read()).hexdigest()
    json.dump({"source": src, "sha256": manifest}, open(os.path.join(dest, "MANIFEST.json"), "w"), indent=1)
    if not quiet:
        print(f"install_ref: staged {len(FILES)} reference files under {dest}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
