"""Decode the geometry / basis / segment tables of the reference's LCCD setup files (test/lccd_frozencore_test.dat,
test/lccd_test.dat: water, 3-21G) and the energy goldens of the reference's tests for them (test/test_qm.cpp:371-468)
into tests/golden/water_321g_setup.json.  Run in the build container (needs /root/reference); the GPU box only sees the
committed JSON.  floats are written with repr() precision (json round-trips doubles exactly)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aces4_b200.setup_reader import read_setup  # noqa: E402

GOLDEN = {  # test/test_qm.cpp:396-397, 412-415 (DISABLED_lccd_dropcoreinsial_test) and :447-448, 459-462 (lccd_frozencore_test)
    "scf_energy": -75.58432674274046, "lccd_correlation": -0.12610179886435, "lccd_energy": -75.71042854160481,
    "tolerance": 1e-10,
    # all-electron runs of the same molecule: test/test_qm.cpp:651-652, 677-678 (DISABLED_eom_lccd_test: rlccd_rhf.siox
    # without drop_mo) and :732-733, 758-759 (DISABLED_eom_mp2_test: mp2_rhf_disc.siox)
    "all_electron": {"scf_energy": -75.58432674274034, "lccd_energy": -75.71210049055006, "mp2_energy": -75.70540831822183,
                     # test/test_qm.cpp:526-529 (DISABLED_lccsd_test: rlccsd_rhf.siox, all electron)
                     "lccsd_correlation": -0.12865706498547, "lccsd_energy": -75.71298380772593,
                     # test/test_qm.cpp:252-253 (DISABLED_eom_test: rccsd_rhf.siox, cc_conv 1e-12) and :907-908 / :990-991
                     # (eom_ccsd_water_right_test / eom_ccsd_water_test = BASELINE config 3: the same run stopped at cc_conv 1e-10)
                     "ccsd_energy": -75.71251002928709, "ccsd_energy_cc_conv_1e-10": -75.71251002936883}}
# hydrogen fluoride / 3-21G: test/test_qm.cpp:86-102 (second_ccsdpt_test, the CCSD stage of the reference's enabled CCSD(T) test;
# cc_conv 1e-10) and :810-824 (lamccsdpt_test: drop_mo=1-1, cc_conv 1e-12)
GOLDEN["hf"] = {"scf_energy": -99.45975176375698, "ccsd_correlation": -0.12588695910754, "ccsd_energy": -99.58563872286452,
                "frozen_core_ccsd_energy": -99.583972376431,
                # :110-126 (second_ccsdpt_test): ccsdpt_energy = ccsd_energy + eaaa + esaaa + eaab + esaab
                "eaaa": -0.00001091437340, "esaaa": 0.00000240120432, "eaab": -0.00058787722879, "esaab": 0.00003533079603,
                "ccsdpt_energy": -99.58619978246637}
# neon / (9s4p1d) -> [3s2p1d], spherical d: test/test_qm.cpp:22-52 (DISABLED_ccsdpt_test = BASELINE config 2, test/ccsdpt_test.dat:
# scf_rhf_coreh, tran_rhf_no4v, rccsd_rhf, rccsdpt_aaa, rccsdpt_aab).  The setup stops the SCF at 1e-8 and the CCSD at 1e-7, so
# these two numbers carry that run's convergence error: they pin a tightly converged run only to about the .dat's own cc_conv.
GOLDEN["ne_ccsdpt_test"] = {"eaab": -0.0010909774775509193, "esaab": 8.5547845910409156e-05, "cc_conv": 1e-07, "scf_conv": 1e-08}
# water / 3-21G excited states: test/test_qm.cpp:1005-1031 (eom_ccsd_water_test = BASELINE config 3) and :921-933
# (eom_ccsd_water_right_test): the four EOM-CCSD roots `sek0` of eom_ccsd_rhf_right.siox, asserted at 1e-8 (the setup's
# eom_tol is 1e-6, cc_conv 1e-10); :265-272 (DISABLED_eom_test, eom_test.dat: cc_conv 1e-12): the two CIS roots of rcis_rhf.siox
# at 1e-10 -- they pin the CIS starting vectors the harness hands to the EOM program -- and its two EOM roots at 1e-8
GOLDEN["eom_ccsd_water_test"] = {"sek0": [0.32850657002707, 0.41193399006592, 0.42288344162832, 0.51159731180444],
                                 "tolerance": 1e-8, "oscnorm": [0.00680956, 0.0, 0.09037060, 0.11312310]}
# eom_ccsd_water_right_test (:927-932): right-hand transition moments, the first two asserted at 1e-4
GOLDEN["eom_ccsd_water_right_test"] = {"rdipmom": [0.17558771, 0.0, 0.56188382, 0.56905669]}
GOLDEN["eom_test"] = {"cis_sek0": [0.36275490375537, 0.43493738840536], "eom_sek0": [0.32850656893104, 0.41193399028059]}
# hydrogen fluoride / 3-21G, the reference's enabled rlambda_test (test/test_qm.cpp:307-341: scf, tran, rccsd_rhf, rlambda_rhf;
# scf_conv / cc_conv 1e-12): the lambda pseudo-energy at 1e-10
GOLDEN["rlambda_test"] = {"lambda_pseudo": -0.12592115116563,
                          # :317, :337 the z components the test carries in its `expected` arrays (it asserts x and y only)
                          "scf_dipole_z": 0.84792717246707, "ccsd_dipole_z": 0.80028992302928}
# hydrogen fluoride / 3-21G, frozen core: the reference's enabled lamccsdpt_test (test/test_qm.cpp:798-869: scf, tran, rccsd_rhf,
# rlambda_rhf, rlamccsdpt_aaa, rlamccsdpt_aab; cc_conv 1e-12), Lambda-CCSD(T): every number asserted at 1e-10
GOLDEN["lamccsdpt_test"] = {"ccsd_energy": -99.583972376431, "eaaa": -0.00001109673867, "esaaa": 0.00000190144932,
                            "eaab": -0.00057100058054, "esaab": 0.00002785417018, "ccsdpt_energy": -99.584524718131}
# hydrogen fluoride / 3-21G: the reference's enabled cis_test (test/test_qm.cpp:153-202: scf, tran, rcis_rhf, rcis_d_rhf; two roots,
# the degenerate 1-Pi pair): CIS roots `sek0` and the CIS(D) corrections `ekd`, all at 1e-10
GOLDEN["cis_test"] = {"sek0": [0.43427913493064, 0.43427913493252], "ekd": [-0.03032115569246, -0.03032115569276]}
out = {"golden": GOLDEN, "source": "UFParLab/aces4 test/*.dat decoded by aces4_b200/setup_reader.py", "setups": {}}
for name in ("lccd_frozencore_test.dat", "lccd_test.dat", "eom_lccd_test.dat", "lccsd_test.dat", "second_ccsdpt_test.dat",
             "lamccsdpt_test.dat", "ccsdpt_test.dat", "eom_ccsd_water_test.dat", "eom_test.dat", "rlambda_test.dat", "cis_test.dat"):
    s = read_setup(open(os.path.join("/root/reference/test", name), "rb").read())
    assert s["trailing_bytes"] == 0
    keep_f = ("alphas", "charge", "coords", "pcoeffs")
    keep_i = ("atom", "ivangmom", "ncfps", "npfps", "ixalphas", "ixpcoeffs", "ccbeg", "ccend", "end_nfps",
              "moa_seg_ranges", "ao_seg_ranges")
    out["setups"][name] = {
        "programs": s["programs"], "ints": s["ints"], "scalars": s["scalars"], "segments": s["segments"],
        "arrays": {k: s["arrays"][k] for k in keep_f}, # (ccsdpt_test.dat is an older setup without the shell -> atom table: one centre)
        "int_arrays": {k: s["int_arrays"][k] for k in keep_i if k in s["int_arrays"]}}
path = os.path.join(ROOT, "tests", "golden", "water_321g_setup.json")
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
print("wrote", path, os.path.getsize(path), "bytes")
