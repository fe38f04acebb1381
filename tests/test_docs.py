"""Evidence hygiene: every file the documents cite exists in the tree (profiles/, tests/, tools/, csrc/, oracle/ ...)."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# names of REFERENCE files (not in this tree), scratch output the text mentions as such, and patterns
NOT_OURS = {"CooLBM_MRT_combustion.cpp", "PF/apps/laplace3D.h", "SC/apps/laplace2D.h", "_b.txt", "_c.txt", "_d.txt"}
SEARCH = ["", "profiles/", "multiphase-lbm_b200/", "multiphase-lbm_b200/csrc/", "multiphase-lbm_b200/apps/", "tools/", "tools/probe/",
          "tools/gpu_runs/", "tests/", "tests/golden/", "tests/host_check/", "tests/host_check/dropin/", "oracle/", "oracle/ref_harness/",
          "include/"]


def test_cited_files_exist():
    missing = []
    for doc in ("DESIGN.md", "README.md", "INTEGRATION.md", "profiles/README.md", "tools/README.md"):
        text = open(os.path.join(ROOT, doc)).read()
        for m in set(re.findall(r"`([A-Za-z0-9_\-./*]+\.(?:json|txt|csv|py|cu|cuh|h|c|cpp|sh|md))`", text)):
            if m in NOT_OURS or m.startswith(("gpurun_out/", "/", "SC/", "PF/", "AB/")):      # SC/ PF/ AB/ = the reference's trees
                continue
            if not any(glob.glob(os.path.join(ROOT, d + m)) for d in SEARCH):
                missing.append((doc, m))
    assert not missing, missing
