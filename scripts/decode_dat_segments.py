"""Decode the segment tables of the reference's shipped setup files (test/*.dat) into tests/golden/dat_segments.json.
Run in the build container (needs /root/reference); the GPU box only sees the committed JSON."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aces4_b200.setup_reader import occ_virt_segments, read_setup  # noqa: E402

out = {}
for f in sorted(glob.glob("/root/reference/test/*.dat")):
    s = read_setup(open(f, "rb").read())
    assert s["trailing_bytes"] == 0, f
    occ, virt = occ_virt_segments(s)
    out[os.path.basename(f)] = {"segments": s["segments"], "occ": occ, "virt": virt, "programs": s["programs"],
                                "ints": {k: s["ints"][k] for k in ("baocc", "eaocc", "bavirt", "eavirt", "norb") if k in s["ints"]}}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "dat_segments.json"), "w"), indent=1, sort_keys=True)
print(len(out), "files decoded")
