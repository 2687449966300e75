"""The full LCCD, LCCSD and CCSD programs (tests/golden/lccd_program.sialx, lccsd_program.sialx, ccsd_program.sialx) walked on the DEVICE backend without a GPU: the deferred op
stream in DRY mode (fake device addresses, nothing executes) takes every C-ABI call the GPU energy test
(tests/test_gpu_z_lccd_water_energy.py) will make -- label validation, pattern analysis, the recorder and the
scheduler all run on the host.  Checks that the program is accepted end to end, that the schedule honours the hazards
of the recorded order and that the hot loops fuse (chains / fused accumulates); the arithmetic is the GPU test's job."""
import pytest

import lccd_water as lw
from aces4_b200.sial_frontend import DeviceBackend, Program, Walker


@pytest.fixture(scope="module")
def sip():
    import aces4_b200 as s

    s.build()
    return s.api


class DryArray:
    """stands in for api.DistArray (whose slab needs a GPU): one dry block per SIAL block, single owner"""

    def __init__(self, sip, seg_ext):
        self.sip, self.seg_ext, self.blocks = sip, seg_ext, {}

    def owner(self, idx):
        return 0

    def block_view(self, idx):
        idx = tuple(idx)
        if idx not in self.blocks:
            self.blocks[idx] = self.sip.DeviceBlock(tuple(s[i - 1] for s, i in zip(self.seg_ext, idx)))
        return self.blocks[idx]

    def put(self, idx, blk):
        self.block_view(idx).scale_and_copy(blk, 1.0)

    def put_accumulate(self, idx, blk):
        self.block_view(idx).accumulate(blk)

    def put_initialize(self, idx, v):
        self.block_view(idx).fill(v)


class DryBackend(DeviceBackend):
    """DeviceBackend with its two device-touching calls (scalar read-back, device sync) stubbed"""

    def _val(self, s):
        return 0.0

    def barrier(self):            # api.sync() needs a device; a barrier drains the recording, which dry mode can do
        self.api.wl_flush()
        self.cache.clear()


@pytest.mark.parametrize("segmentation,program", [("dat", "lccd"), ("fine", "lccd"), ("all_dat", "lccd"),
                                                  ("all_fine", "lccd"), ("all_dat", "lccsd"), ("all_fine", "lccsd"),
                                                  ("all_dat", "ccsd"), ("all_fine", "ccsd"), ("hf_fine", "ccsd+t"),
                                                  ("ne_fine", "ccsd+t"), ("ne_dat", "ccsd+t")])
def test_lccd_program_records_and_schedules_on_the_device_backend(sip, segmentation, program):
    inp = lw.inputs(segmentation)
    sip.set_predefined_int_array("moa_seg_ranges", inp["moa_seg_ranges"])
    prog = Program({"lccd": lw.PROGRAM, "lccsd": lw.PROGRAM_LCCSD, "ccsd": lw.PROGRAM_CCSD, "ccsd+t": lw.PROGRAM_CCSD}[program])
    with sip.recording(dry=True):
        arrays = {name: DryArray(sip, [inp["segs"][k] for k in kinds]) for name, kinds in lw.KINDS.items()}
        be = DryBackend(sip, arrays, record=False)       # one recording around everything (ended by the with block)
        be.fock = sip.DeviceBlock(inp["fock"].shape)
        if program == "ccsd+t":      # ... with the integral transformation program in front
            Walker(Program(lw.PROGRAM_TRAN), be, inp["segs"], index_base=inp["index_base"]).run()
            sip.wl_flush()
        w = Walker(prog, be, inp["segs"], index_base=inp["index_base"])
        w.run()
        sip.wl_flush()
        w.run_proc("iteration")
        sip.wl_flush()
        if program == "ccsd+t":      # the rank-6 stream of tests/golden/ccsd_t_restated.sialx
            Walker(Program(lw.PROGRAM_PT), be, inp["segs"], index_base=inp["index_base"]).run()
            sip.wl_flush()
        st = sip.wl_stats()
        level, unit = sip.wl_last_plan()
    assert st["recorded"] > (2000 if segmentation == "fine" else 60)
    assert st["scheduled"] < st["recorded"]
    assert st["fused_accumulates"] > 0 and st["temps_elided"] > 0
    assert len(level) == len(unit) > 0 and min(level) >= 1
    assert not w.locals


def test_static_array_slices_in_place_record_through_the_c_abi(sip):
    """the in-place path of DeviceBackend (`static_in_place`): every contraction with a block of `ca` goes through
    sipgpu_block_contract_sliced -- dry mode checks that the library accepts the whole transformation program and an
    LCCD iteration that way (pattern, slice bounds against the parent extents)"""
    inp = lw.inputs("fine")
    ca, ca_segs = lw.dense_ca(inp)
    sip.set_predefined_int_array("moa_seg_ranges", inp["moa_seg_ranges"])
    with sip.recording(dry=True):
        arrays = {name: DryArray(sip, [inp["segs"][k] for k in kinds]) for name, kinds in lw.KINDS.items()}
        be = DryBackend(sip, arrays, record=False, static_in_place={"ca": (sip.DeviceBlock(ca.shape), ca_segs)})
        be.fock = sip.DeviceBlock(inp["fock"].shape)
        Walker(Program(lw.PROGRAM_TRAN), be, inp["segs"], index_base=inp["index_base"]).run()
        sip.wl_flush()
        w = Walker(Program(lw.PROGRAM), be, inp["segs"], index_base=inp["index_base"])
        w.run()
        w.run_proc("iteration")
        sip.wl_flush()
        st = sip.wl_stats()
    assert st["recorded"] > 5000 and not arrays["ca"].blocks       # no block of ca was ever materialised
