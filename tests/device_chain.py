"""Device-side harness shared by the GPU tests that run the reference's programs VERBATIM on libsipgpu
(tests/test_gpu_z_cc_reference_programs.py, tests/test_gpu_z_eom_ccsd.py): uploads what the SCF / transformation programs hand
over into resident distributed arrays, registers them under the reference's persistence labels, and walks a program with
the SIAL front-end on the DeviceBackend.  Test infrastructure; `sip` is aces4_b200.api (or the numpy stand-in of
tests/fake_device_api.py in the CPU twin)."""
import numpy as np

import lccd_water as lw
from oracle import qm_inputs as qm


def upload(sip, A, blocks):
    for idx, b in blocks.items():
        view = A.block_view(idx)
        sip._check(sip.lib().sipgpu_h2d(view.ptr, sip._hp(np.asfortranarray(b)), view.size), "h2d")


def resident(sip, seg_lists, blocks):
    A = sip.DistArray(seg_lists)
    A.fill_local(0.0)
    upload(sip, A, blocks)
    return A


def hand_over_scf_and_transformation(sip, case, inp, transformed=True):
    """persist the MO integral classes (transformed: from the dense numpy transformation; else the reference's transformation
    program has to run first), ca / fock_a (over ALL orbital segments) and scf_energy under the labels the CC programs restore
    (rccsd_rhf.sialx:225-244, rlccd_rhf.sialx:243-251, :917-919).  -> (seg_ext per index kind, resident aoint, device Fock block)"""
    sip.set_predefined_int_array("moa_seg_ranges", inp["moa_seg_ranges"])
    seg_ext = lw.segs_with_all_orbitals(inp)
    seg_ext["p"] = list(inp["segs"]["o"]) + list(inp["segs"]["v"])
    given = {lab: resident(sip, [seg_ext[k] for k in lw.KINDS[lab.lower()]], inp["arrays"][lab.lower()])
             for lab in (lw.PERSISTED if transformed else ())}
    statics = lw.all_orbital_statics(case, inp)
    given["ca"] = resident(sip, [seg_ext["ao"], seg_ext["pa"]], statics["ca"])
    given["fock_a"] = resident(sip, [seg_ext["pa"], seg_ext["pa"]], statics["fock_a"])
    aoint = resident(sip, [seg_ext[k] for k in lw.KINDS["aoint"]], inp["arrays"]["aoint"])
    sip.sync()
    for label, A in given.items():
        A.persist(label)
    sip.persist_scalar("scf_energy", inp["e_scf"])
    return seg_ext, aoint, sip.DeviceBlock.from_numpy(inp["fock"])


def static_arrays(sip, seg_ext):
    """empty arrays of the defs files' statics, for `restore_persistent ca "ca"` / `Fock_a "fock_a"` to adopt into"""
    return {"ca": sip.DistArray([seg_ext["ao"], seg_ext["pa"]]), "fock_a": sip.DistArray([seg_ext["pa"], seg_ext["pa"]])}


class TracingWalker:
    """mixin: records, per excited state, whether the EOM program's own Davidson solver reported convergence (`converged = 1`)"""

    def _x_call(self, name):
        out = super()._x_call(name)
        if name == "collapse_davidson":
            self.__dict__.setdefault("state_converged", {})[self.idx["kstate"]] = self.be.value(self.scalars["converged"]) == 1.0
        return out


def run_program_on_device(sip, text, case, inp, seg_ext, aoint, fock, record, constants, extra_arrays=None, host_data=None, trace=False):
    """-> (walker, backend, scalars as floats)"""
    from aces4_b200.sial_frontend import DeviceBackend, Program, Walker

    prog = Program(text)
    arr = lw.device_program_arrays(sip, prog, text, constants, inp["segs"], skip=("aoint",))
    arr["aoint"] = aoint
    arr.update(extra_arrays or {})
    be = DeviceBackend(sip, arr, record=record)
    be.fock, be.seg_ranges = fock, inp["moa_seg_ranges"]
    if host_data is None:      # what the dipole integral engine and the SCF program would supply (lambda / EOM property parts)
        host_data, scf_dipole = lw.dipole_data(lw.CASES[case][0])
        Walker.host_registry.setdefault("scf_dipole", scf_dipole)
    cls = type("TracedWalker", (TracingWalker, Walker), {}) if trace else Walker
    w = cls(prog, be, lw.segs_with_all_orbitals(inp), index_base=inp["index_base"], constants=constants, host_data=host_data)
    sc = w.run()
    return w, be, {k: be.value(v) for k, v in sc.items()}
