"""Extract the distinct block-contraction label patterns `D[..] = L[..] * R[..]` from the reference's SIAL programs.

Run HERE (the reference checkout is not on the GPU box); the output tests/golden/sial_contraction_patterns.txt is
committed and is what tests and the sweep read.  One line per distinct pattern:

    <dlabels> <llabels> <rlabels> <kinds> <count> <first occurrence file:line>

Labels are canonicalised by first appearance (a, b, c, ...); <kinds> gives, per canonical label, the index kind from
the SIAL declaration: o = occupied moaindex (b?occ:e?occ), v = virtual, p = other moaindex range, n = aoindex,
x = simple `index`, s = subindex.  Index kinds decide segment extents in the tests.
"""
import os
import re
import sys
from collections import OrderedDict

REF = "/root/reference/src/sialx/qm"
FILES = ["cc/rccsd_rhf.sialx", "cc/rlccd_rhf.sialx", "cc/rccsdpt_aab.sialx", "cc/rccsdpt_aaa.sialx",
         "cc/rlambda_rhf.sialx", "cc/mp2_rhf_disc.sialx", "eom/eom_ccsd_rhf_right.sialx", "eom/eom_ccsd_rhf_left.sialx"]
DECL = re.compile(r"^\s*(aoindex|moaindex|mobindex|moindex|index|laindex|subindex)\s+(\w+)\s*(?:=\s*([^#]*?)|of\s+(\w+))\s*(?:#.*)?$", re.I)
STMT = re.compile(r"^\s*(\w+)\s*\[([^\]]*)\]\s*(=|\+=|-=)\s*(\w+)\s*\[([^\]]*)\]\s*\*\s*(\w+)\s*\[([^\]]*)\]\s*(?:#.*)?$")


def kind_of(decl, rng, parent, kinds):
    decl = decl.lower()
    if decl == "subindex":
        return kinds.get(parent, "s")
    if decl in ("aoindex", "laindex"):
        return "n"
    if decl == "index":
        return "x"
    r = (rng or "").lower()
    lo, _, hi = r.partition(":")
    if "occ" in lo and "occ" in hi:
        return "o"
    if "virt" in lo and "virt" in hi:
        return "v"
    return "p"


def main(out_path):
    pats = OrderedDict()
    for f in FILES:
        path = os.path.join(REF, f)
        if not os.path.exists(path):
            continue
        kinds = {}
        lines = open(path, errors="replace").read().splitlines()
        for ln in lines:
            m = DECL.match(ln)
            if m:
                kinds[m.group(2)] = kind_of(m.group(1), m.group(3), m.group(4), kinds)
        for no, ln in enumerate(lines, 1):
            m = STMT.match(ln)
            if not m:
                continue
            dl = [x.strip() for x in m.group(2).split(",")]
            ll = [x.strip() for x in m.group(5).split(",")]
            rl = [x.strip() for x in m.group(7).split(",")]
            if not all(x in kinds for x in dl + ll + rl):
                continue
            canon = {}
            for x in dl + ll + rl:
                if x not in canon:
                    canon[x] = "abcdefghijklmnopqrstuvwxyz"[len(canon)]
            cnt = {}
            for x in dl + ll + rl:
                cnt[x] = cnt.get(x, 0) + 1
            if any(c != 2 for c in cnt.values()):
                continue  # not a plain contraction (repeated label / trace): the reference rejects these too
            key = ("".join(canon[x] for x in dl), "".join(canon[x] for x in ll), "".join(canon[x] for x in rl),
                   "".join(kinds[x] for x in canon))
            if key not in pats:
                pats[key] = [0, f"{f}:{no}"]
            pats[key][0] += 1
    with open(out_path, "w") as fh:
        fh.write("# distinct contraction label patterns of the reference's SIAL programs (scripts/extract_sial_patterns.py)\n")
        fh.write("# dlabels llabels rlabels kinds count first_occurrence\n")
        for (d, l, r, k), (c, where) in pats.items():
            fh.write(f"{d} {l} {r} {k} {c} {where}\n")
    print(len(pats), "distinct patterns ->", out_path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                            "tests", "golden", "sial_contraction_patterns.txt"))
