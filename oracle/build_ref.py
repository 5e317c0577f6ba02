#!/usr/bin/env python
"""Recipe for oracle/_ref/: the UNMODIFIED reference hot path as a CPU arm that travels to the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY.  Nothing from the reference is committed: `oracle/_ref/` is git-ignored (it is NOT
gpurun-ignored, so it ships to the GPU box next to the built .so files).  `__graft_entry__.build()` runs this in the
build container, where `/root/reference` exists; on the GPU box the prebuilt directory is used as it is.

The reference's matching path is four pure-Python files that import with torch + numpy + scipy only
(SURVEY.md section 8c): they are copied byte for byte, together with the three package `__init__.py` files of the
reference, and a manifest with their sha256 is written so that `bench.py` can state what it timed.

    python oracle/build_ref.py            # -> oracle/_ref/dmm/...
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("DMM_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

# the hot path of SURVEY.md section 8(a) + the package files python needs to import it
FILES = [
    "dmm/__init__.py",
    "dmm/modules/__init__.py",
    "dmm/utils/__init__.py",
    "dmm/modules/match_model.py",                 # MatchModel (match_model.py:13-152)
    "dmm/modules/submodules/relax_match.py",      # relax_matching / project_row / project_col / hungarian_matching
    "dmm/utils/match_helper.py",                  # compute_iou_binary_mask_2D / get_cosine_score / compute_matching_loss
    "dmm/utils/checker.py",                       # CHECK* asserts
]


def available() -> bool:
    """True when a usable oracle/_ref exists (built here earlier, or shipped to the GPU box)."""
    return os.path.exists(os.path.join(OUT, "MANIFEST.json")) and \
        os.path.exists(os.path.join(OUT, "dmm", "modules", "match_model.py"))


def build(force: bool = False) -> str:
    """Copy the files (only when the reference tree is present); returns OUT, or '' when it cannot be built."""
    if not os.path.isdir(os.path.join(REF_ROOT, "dmm")):
        return OUT if available() else ""
    manifest = {"source": REF_ROOT, "files": {}}
    for rel in FILES:
        src, dst = os.path.join(REF_ROOT, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        data = open(src, "rb").read()
        manifest["files"][rel] = hashlib.sha256(data).hexdigest()
        if force or not os.path.exists(dst) or open(dst, "rb").read() != data:
            shutil.copyfile(src, dst)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return OUT


def verify() -> bool:
    """The shipped copy still is what the manifest says (nobody edited the reference arm)."""
    if not available():
        return False
    man = json.load(open(os.path.join(OUT, "MANIFEST.json")))
    for rel, digest in man["files"].items():
        p = os.path.join(OUT, rel)
        if not os.path.exists(p) or hashlib.sha256(open(p, "rb").read()).hexdigest() != digest:
            return False
    return True


def import_reference():
    """-> the reference's (MatchModel, relax_matching, match_helper module), imported from oracle/_ref."""
    if not verify():
        raise RuntimeError("oracle/_ref is missing or does not match its manifest: run `python oracle/build_ref.py` "
                           "in a container that has /root/reference")
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    from dmm.modules.match_model import MatchModel
    from dmm.modules.submodules.relax_match import relax_matching
    from dmm.utils import match_helper
    return MatchModel, relax_matching, match_helper


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    print(out or "reference tree not found and no prebuilt oracle/_ref")
