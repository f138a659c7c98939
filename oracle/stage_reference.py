"""Stages the UNMODIFIED Python reference for the GPU box: copies the files the hot path needs from /root/reference
(read-only, build container only) into baseline/_ref/ — git-ignored (the history stays source-only) but NOT
gpurun-ignored, so it travels with the snapshot.  `bench.py --impl reference` then times the reference's own
`ExplainGridTDAttention.explain_caption_wordt` + `explain_cnn` on the box's host cores through oracle/ref_shim.py
(`cpu_baseline.kind: "reference"`); without the staged copy it falls back to the oracle port (`kind: "port"`).

The reference has no setup.py / pyproject.toml, so the `pip install --target baseline/_ref` recipe does not apply
(DESIGN.md section 2); this script is the committed recipe instead.  TEST / BENCH INFRASTRUCTURE ONLY."""
import os
import shutil
import sys

SRC = os.environ.get("LRPX_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
WANT = ["LRPtools", "models", "evaluation.py", "config.py"]


def stage(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "LRPtools")):
        if verbose:
            print(f"reference not mounted at {SRC}: nothing staged")
        return False
    os.makedirs(DST, exist_ok=True)
    for name in WANT:
        s, d = os.path.join(SRC, name), os.path.join(DST, name)
        if os.path.isdir(s):
            shutil.copytree(s, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.pth", "*.pt"))
        elif os.path.exists(s):
            shutil.copy2(s, d)
    if verbose:
        print(f"staged {WANT} -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
