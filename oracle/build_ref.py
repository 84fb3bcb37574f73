"""TEST / BASELINE INFRASTRUCTURE ONLY — recipe for ``oracle/_ref``: the UNMODIFIED reference sources of the hot path.

The reference's path is two pure-Python files (``models/resunet.py``, ``models/base.py``) plus the un-vendored torchlibrosa
(restated in ``oracle/torchlibrosa``).  ``/root/reference`` exists only in the build container, so this recipe copies those
two files, byte for byte, from where they lie into the git-ignored ``oracle/_ref/models/`` — the Python analogue of compiling a
C reference into ``oracle/_ref``: the copy travels to the GPU box with the snapshot (it is not gpurun-ignored) and lets
``bench.py`` time the reference itself there (``cpu_baseline.kind = "reference"``) and the GPU tests compare against it.
Nothing under ``oracle/_ref`` is ever committed, and the product (``lass_b200/``) never imports it.

    python oracle/build_ref.py            (also run by __graft_entry__.build() when /root/reference is present)
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
FILES = ("models/resunet.py", "models/base.py", "losses.py", "optimizers/lr_schedulers.py", "data/waveform_mixers.py")


def build(reference_root="/root/reference"):
    if not os.path.isfile(os.path.join(reference_root, FILES[0])):
        return False
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(reference_root, rel), os.path.join(REF_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(REF_DIR, "MANIFEST.json"), "w") as f:
        json.dump({"source": reference_root, "sha256": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    ok = build(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("oracle/_ref built" if ok else "reference tree not present: oracle/_ref not built")
