"""The bodies of the device energy tests (tests/test_gpu_z_lccd_water_energy.py), executed on the CPU against a numpy
stand-in for the Python API of libsipgpu (tests/fake_device_api.py).  Those GPU tests were written after the round's
GPU minutes were spent; this runs everything in them that is not the CUDA library -- the front-end's DeviceBackend
(device-resident scalars across iterations, block views, put / get traffic, local arrays) and the assertions of the
tests themselves -- so that what remains unexercised until the first B200 run is libsipgpu's arithmetic alone, which the
rest of the GPU suite has already covered op by op."""
import os
import pytest

import test_gpu_z_cross_product as xp
import test_gpu_z_lccd_water_energy as dev
from fake_device_api import FakeApi


@pytest.mark.parametrize("case,record", [("fine", True), ("all_dat", True)])
def test_lccd_device_test_body_on_the_fake_api(oracle, case, record):
    dev.test_lccd_energy_on_the_device_matches_the_reference_golden(FakeApi(oracle), case, record)


@pytest.mark.parametrize("case,record", [("all_dat", False)])
def test_lccsd_device_test_body_on_the_fake_api(oracle, case, record):
    dev.test_lccsd_energy_on_the_device_matches_the_reference_golden(FakeApi(oracle), case, record)


def test_ccsd_device_test_body_on_the_fake_api(oracle):
    dev.test_ccsd_energy_on_the_device_matches_the_reference_golden(FakeApi(oracle), "all_dat", True)


@pytest.mark.skipif(not os.environ.get("SIPGPU_SLOW_TESTS"), reason="a twin of a test that runs anyway; SIPGPU_SLOW_TESTS=1 (keeps the CPU suite at a few minutes)")
def test_ccsd_t_device_test_body_on_the_fake_api(oracle):
    dev.test_ccsd_t_energy_of_hydrogen_fluoride_on_the_device(FakeApi(oracle), "hf_dat", True)
    dev.test_ccsd_energy_of_hydrogen_fluoride_on_the_device(FakeApi(oracle))


@pytest.mark.parametrize("case,record", [("ne_dat", True), ("ne_dat", False)])
def test_neon_ccsd_t_device_test_body_on_the_fake_api(oracle, case, record):
    dev.test_ccsd_t_of_neon_on_the_device_matches_the_goldens_of_ccsdpt_test(FakeApi(oracle), case, record)


def test_transformation_pipeline_device_test_body_on_the_fake_api(oracle):
    dev.test_transformation_then_cc_program_on_the_device(FakeApi(oracle), "fine", "lccd")


def test_persistence_chain_device_test_body_on_the_fake_api(oracle):
    dev.test_programs_chained_through_persistent_arrays_on_the_device(FakeApi(oracle))


def test_static_in_place_device_test_body_on_the_fake_api(oracle):
    dev.test_static_array_blocks_read_in_place_through_the_programs(FakeApi(oracle))


def test_reference_triples_programs_device_test_body_on_the_fake_api(oracle):
    import test_gpu_z_ccsdpt_reference as pt
    pt.test_reference_triples_programs_on_the_device(FakeApi(oracle), "hf_dat", True)


@pytest.mark.skipif(not os.environ.get("SIPGPU_SLOW_TESTS"), reason="a twin of a test that runs anyway; SIPGPU_SLOW_TESTS=1 (keeps the CPU suite at a few minutes)")
def test_reference_eom_program_device_test_body_on_the_fake_api(oracle):
    import test_gpu_z_eom_ccsd as eom
    # (without the left-hand program, 25 s more: its CPU twin is tests/test_eom_ccsd_cpu.py::test_eom_ccsd_water_test_in_full)
    eom.test_reference_eom_program_on_the_device(FakeApi(oracle), "eom_dat", True, with_left=bool(os.environ.get("SIPGPU_SLOW_TESTS")))


def test_reference_cc_programs_device_test_bodies_on_the_fake_api(oracle):
    import test_gpu_z_cc_reference_programs as cc
    cc.test_reference_lccd_program_on_the_device(FakeApi(oracle), "dat", True)
    cc.test_reference_lccsd_and_ccsd_programs_on_the_device(FakeApi(oracle))
    cc.test_config_1_from_ao_integrals_on_the_device(FakeApi(oracle))
    cc.test_reference_cis_program_on_the_device(FakeApi(oracle))
    cc.test_reference_lambda_program_on_the_device(FakeApi(oracle))
    cc.test_reference_lambda_ccsdpt_programs_on_the_device(FakeApi(oracle), "hf_fc_dat")
    cc.test_reference_cis_and_cis_d_programs_on_the_device(FakeApi(oracle))


def test_reference_unit_program_device_test_bodies_on_the_fake_api(oracle):
    import ref_unit_programs as rp
    import test_gpu_z_reference_unit_programs as up
    for case in rp.ALL:
        up.test_reference_unit_program_on_the_device(FakeApi(oracle), case, True)


def test_cross_product_test_bodies_on_the_fake_api(oracle):
    """tests/test_gpu_z_cross_product.py at a block size the CPU finishes in seconds"""
    xp.test_full_cross_product_s16_against_the_oracle(FakeApi(oracle), oracle, s=3)
    xp.test_full_cross_product_s32_equivariance(FakeApi(oracle), oracle, s=3)
