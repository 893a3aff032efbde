"""Regenerates tests/golden/ from the reference tree (run in the development container, where /root/reference is
mounted): the fixtures are the reference's OWN shipped inputs and answer files (matrix/*), copied verbatim -- they are
data, not source.  Prints the sha256 of every file so tests/golden/PROVENANCE.md can be checked."""
import hashlib
import os
import shutil
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/matrix"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(DST, exist_ok=True)
for name in sorted(os.listdir(SRC)):
    if not name.startswith("ELSES_MATRIX_"):
        continue
    shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
    print(hashlib.sha256(open(os.path.join(DST, name), "rb").read()).hexdigest(), name)
