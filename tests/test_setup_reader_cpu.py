"""CPU tests of the `.dat` setup reader (aces4_b200/setup_reader.py): a file assembled here in the reference's stream
format parses back; the decoded segment tables of the shipped test inputs (tests/golden/dat_segments.json, produced by
scripts/decode_dat_segments.py) match SURVEY.md 8(d); and -- when the reference checkout is present -- the golden file is
regenerated and compared."""
import glob
import json
import os
import struct

import pytest

from aces4_b200.setup_reader import MAGIC, SetupFormatError, occ_virt_segments, read_setup

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "dat_segments.json")))


def s_(x):
    return struct.pack("<i", len(x) + 1) + x.encode() + b"\0"


def ints_(v):
    return struct.pack(f"<i{len(v)}i", len(v), *v)


def build_dat():
    b = struct.pack("<ii", MAGIC, 1)
    b += struct.pack("<i", 2) + s_("scf_rhf.siox") + s_("rlccd_rhf.siox")
    b += struct.pack("<i", 4) + b"".join(s_(k) + struct.pack("<i", v) for k, v in
                                         [("baocc", 1), ("eaocc", 2), ("bavirt", 3), ("eavirt", 4)])
    b += struct.pack("<i", 1) + s_("damp") + struct.pack("<d", 0.25)
    b += struct.pack("<i", 2) + struct.pack("<i", 1001) + ints_([9, 7]) + struct.pack("<i", 1003) + ints_([3, 2, 8, 8])
    b += struct.pack("<i", 1) + s_("charge") + struct.pack("<i", 1) + ints_([3]) + struct.pack("<i3d", 3, 8.0, 1.0, 1.0)
    b += struct.pack("<i", 1) + s_("flags") + struct.pack("<i", 2) + ints_([2, 1]) + ints_([4, 5])
    b += struct.pack("<i", 1) + s_("rlccd_rhf.siox") + struct.pack("<i", 1) + s_("cc_iter") + s_("12")
    return b


def test_round_trip_of_a_file_in_the_reference_format():
    s = read_setup(build_dat())
    assert s["programs"] == ["scf_rhf.siox", "rlccd_rhf.siox"]
    assert s["segments"] == {"ao": [9, 7], "moa": [3, 2, 8, 8]}
    assert occ_virt_segments(s) == ([3, 2], [8, 8])
    assert s["scalars"] == {"damp": 0.25} and s["arrays"]["charge"] == ([3], [8.0, 1.0, 1.0])
    assert s["int_arrays"]["flags"] == ([2, 1], [4, 5]) and s["configs"]["rlccd_rhf.siox"] == {"cc_iter": "12"}
    assert s["trailing_bytes"] == 0


def test_bad_files_are_rejected():
    good = build_dat()
    with pytest.raises(SetupFormatError):
        read_setup(b"\0\0\0\0" + good[4:])
    with pytest.raises(SetupFormatError):
        read_setup(good[:40])


def test_golden_segment_tables_match_the_survey():
    assert GOLD["lccd_test.dat"]["segments"]["ao"] == [13] and GOLD["lccd_test.dat"]["occ"] == [5]
    assert GOLD["lccd_test.dat"]["virt"] == [8]
    assert GOLD["ccsdpt_test.dat"]["segments"]["ao"] == [14] and GOLD["ccsdpt_test.dat"]["virt"] == [9]
    assert GOLD["second_ccsdpt_test.dat"]["segments"]["moa"] == [5, 6]
    assert GOLD["lccd_frozencore_test.dat"]["segments"]["moa"] == [1, 4, 8]
    assert GOLD["lccd_frozencore_test.dat"]["occ"] == [4] and GOLD["lccd_frozencore_test.dat"]["virt"] == [8]
    assert GOLD["eom_ccsd_water_test.dat"]["segments"]["moa"] == [5, 8]


@pytest.mark.skipif(not os.path.isdir("/root/reference/test"), reason="reference checkout not present (GPU box)")
def test_golden_file_is_what_the_shipped_inputs_decode_to():
    files = sorted(glob.glob("/root/reference/test/*.dat"))
    assert sorted(os.path.basename(f) for f in files) == sorted(GOLD)
    for f in files:
        s = read_setup(open(f, "rb").read())
        assert s["trailing_bytes"] == 0
        g = GOLD[os.path.basename(f)]
        assert s["segments"] == g["segments"] and list(occ_virt_segments(s)) == [g["occ"], g["virt"]]
