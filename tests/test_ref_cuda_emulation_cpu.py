"""CPU pre-check of tests/test_gpu_vs_ref_cuda.py: the reference's legacy CUDA algorithm, emulated statement by statement
in numpy (oracle/legacy_cuda_emulation.py, after src/sip/cuda/gpu_super_instructions.cu:307-684), agrees with the oracle on
every case the GPU test runs against the real thing."""
import numpy as np

import test_gpu_vs_ref_cuda as gpu_cases
from oracle import legacy_cuda_emulation as emu
from oracle import ref_gpu


def test_legacy_contraction_algorithm_equals_the_oracle_on_all_gpu_cases(oracle):
    cases = gpu_cases.contraction_cases() + gpu_cases.config4_cases()
    assert len(cases) >= 190
    for c in cases:
        if int(np.prod(c["x1"][0])) > 3_000_000:
            continue                                  # the 8 MB bench-shape case ran on the GPU; too slow for this emulation
        x1, x2 = ref_gpu.case_inputs(c)
        got = emu.gpu_contract(c["y"][0], c["y"][1], x1, c["x1"][0], c["x1"][1], x2, c["x2"][0], c["x2"][1])
        want, ierr = oracle.contract_labels(c["y"][1], c["y"][0], c["x1"][1], x1, c["x2"][1], x2)
        assert ierr == 0
        want = want.reshape(got.shape)
        assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want)), (c["where"], c["y"], c["x1"], c["x2"])


def test_legacy_permutation_algorithm_equals_the_oracle_on_all_gpu_cases(oracle):
    for c in gpu_cases.permute_cases():
        x1, _ = ref_gpu.case_inputs(c)
        got = emu.gpu_permute(c["y"][0], c["y"][1], x1, c["x1"][0], c["x1"][1])
        assert np.array_equal(got, oracle.permute_labels(c["y"][1], c["x1"][1], x1)), (c["y"], c["x1"])
