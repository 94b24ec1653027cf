#!/usr/bin/env python
"""Stages the reference's OWN implementation of the hot path under oracle/_ref/ so that bench.py's CPU legs
can time the reference itself on the GPU box's host cores (`cpu_baseline.kind = "reference"`), where
/root/reference does not exist.  TEST / BENCH INFRASTRUCTURE ONLY: oracle/_ref/ is git-ignored (nothing from
the reference enters this repository's history) but travels with the working tree like a built .so.

    python oracle/make_ref.py            # needs /root/reference (the build container)

Only the files SURVEY.md section 8(c) lists for the DE-GAP-FFDnet path are staged, unmodified; they are
imported through the same import-time shims as in the container (oracle/ref_import.py, with
DEQSCI_REFERENCE_ROOT pointing at the staged tree).  Weights come from tests/golden/weights_*.npz."""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("DEQSCI_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref", "reference")
FILES = [
    "solvers/equilibrium_solvers_yaping.py",       # EquilibriumProxGradSCI :382-436
    "solvers/new_equilibrium_utils_yaping.py",     # andersonexp :153-189, DEQFixedPoint :241-281
    "solvers/cg_utils.py",                         # imported by the file above (:7), unused on the path
    "utils/__init__.py",
    "utils/cg_utils.py",                           # A_torch_ / At_torch_ :85-129
    "networks/__init__.py",
    "networks/ffdnet/models.py",                   # FFDNet :70-108
    "networks/ffdnet/functions.py",                # noise-map concat / pixel shuffle
    "networks/provable/model/SimpleCNN_models.py", # DnCNN (DE-GAP-CNN / RSN-CNN side legs)
    "networks/provable/model/conv_sn_chen.py",
    "networks/provable/model/bn_sn_chen.py",
]


def main():
    if not os.path.isdir(os.path.join(SRC, "solvers")):
        print("oracle/make_ref.py: no reference tree at %s; nothing staged" % SRC)
        return 1
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = []
    for rel in FILES:
        src = os.path.join(SRC, rel)
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest.append("%s  %s" % (hashlib.sha256(open(src, "rb").read()).hexdigest(), rel))
    with open(os.path.join(HERE, "_ref", "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    print("staged %d reference files under %s" % (len(FILES), DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
