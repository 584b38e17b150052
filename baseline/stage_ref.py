"""Stage the UNMODIFIED reference into git-ignored baseline/_ref/ so that it can travel to the GPU box.

The reference (Lac-quan-yeu-doi/Face-Recognition-Models) has no packaging metadata, so
`pip install --target baseline/_ref /root/reference` has nothing to install (recorded in DESIGN.md); its Python modules
are importable as they lie.  This script copies the `main_code` package byte for byte (no edits, .py files only) from
/root/reference, which exists in the build container only.  baseline/_ref/ is listed in .gitignore (never committed:
no reference sources enter this repo's history) and NOT in .gpurunignore, so `gpurun` and the round-end driver ship it.

Consumers (all of them timing / checking arms, never the product path):
  * bench.py --impl reference and bench.py's cpu_baseline leg: `main_code.utils.criterion.<Head>` +
    nn.CrossEntropyLoss + `main_code.utils.metrics.accuracy` (criterion.py:232-301, model_utils.py:179-182);
  * examples/ref_train_model.py: the reference's own `train_model` (model_utils.py:147-216) behind the INTEGRATION.md shim.

    python baseline/stage_ref.py            # copy; prints the file count and a content hash
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("MH_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")


def stage(verbose: bool = True) -> bool:
    """Returns True when baseline/_ref holds the reference's main_code package afterwards."""
    src_pkg = os.path.join(SRC, "main_code")
    dst_pkg = os.path.join(DST, "main_code")
    if not os.path.isdir(src_pkg):
        if verbose:
            print(f"stage_ref: {src_pkg} not found (only the build container has the reference); "
                  f"baseline/_ref {'present' if os.path.isdir(dst_pkg) else 'absent'}")
        return os.path.isdir(dst_pkg)
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    h = hashlib.sha256()
    n = 0
    for root, _dirs, files in os.walk(src_pkg):
        rel = os.path.relpath(root, SRC)
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            os.makedirs(os.path.join(DST, rel), exist_ok=True)
            data = open(os.path.join(root, f), "rb").read()
            open(os.path.join(DST, rel, f), "wb").write(data)
            h.update(os.path.join(rel, f).encode() + b"\0" + data)
            n += 1
    with open(os.path.join(DST, "STAGED_FROM.txt"), "w") as fh:
        fh.write(f"{n} unmodified .py files copied from {SRC}/main_code by baseline/stage_ref.py\nsha256 {h.hexdigest()}\n")
    if verbose:
        print(f"stage_ref: {n} files -> {dst_pkg} (sha256 {h.hexdigest()[:16]})")
    return True


def reference_path() -> str | None:
    """Directory to put on sys.path to `import main_code...`, or None when the reference was never staged."""
    return DST if os.path.isfile(os.path.join(DST, "main_code", "utils", "criterion.py")) else None


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
