"""Copies the SIAL programs of the reference's block-operation unit tests (src/sialx/test/*.sialx, driven by
test/test_basic_sial.cpp and test/test_sial.cpp: the known-answer tests of SURVEY.md section 8c) into
tests/golden/ref_unit_programs/ VERBATIM, under a two-line header -- TEST INPUTS of the SIAL front-end
(tests/test_reference_unit_programs_cpu.py, tests/test_gpu_z_reference_unit_programs.py), generated in the build container where
/root/reference exists:   python scripts/make_unit_program_goldens.py"""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/src/sialx/test/"
NAMES = ("contraction_small_test", "contraction_small_test2", "transpose_tmp", "transpose4d_tmp", "transpose4d_square_tmp",
         "contract_to_scalar", "sum_op_test", "self_multiply_test", "put_test", "get_mpi", "put_accumulate_mpi",
         "put_accumulate_stress", "tmp_arrays", "tmp_arrays_2", "block_scale_assign",
         # the programs of the reference's dormant CUDA path: the same operations between gpu_on / gpu_put / gpu_allocate / gpu_get /
         # gpu_free / gpu_off statements
         "gpu_contraction_small_test", "gpu_sum_op_test", "gpu_self_multiply_test", "gpu_transpose_tmp", "gpu_contract_to_scalar",
         "gpu_ops", "put_initialize", "put_increment",
         # persistence between consecutive programs
         "persistent_distributed_array_mpi1", "persistent_distributed_array_mpi2", "persistent_scalars_1", "persistent_scalars_2",
         "persistent_static_array_test1", "persistent_static_array_test2",
         "persistent_distributed_array_one_of_three", "persistent_distributed_array_two_of_three", "persistent_distributed_array_three_of_three",
         # the pardo work distribution and the interpreter's scalar / int / if-else arithmetic (no block operations)
         "pardo_loop", "pardo_loop_1d", "pardo_loop_2d", "pardo_loop_3d", "pardo_loop_4d", "pardo_loop_5d", "pardo_loop_6d",
         "pardo_loop_with_pragma", "pardo_with_where", "scalar_ops", "int_ops", "int_self_ops", "ifelse", "index_scalar_cast",
         "exit_statement_test", "return_sval_test",
         # programs whose printed blocks the reference compares with fixture files (test/expected_output/*.txt)
         "static_array_test", "scalar_valued_blocks", "simple_indices_assignments", "local_arrays", "local_arrays_wild", "cast_indices_to_simple",
         # rank-5 served arrays with a leading simple index through a contiguous local array (the EOM programs' idiom)
         "contig_local3")
os.makedirs(os.path.join(ROOT, "tests", "golden", "ref_unit_programs"), exist_ok=True)
for name in NAMES:
    text = open(SRC + name + ".sialx", errors="replace").read()
    head = (f"# tests/golden/ref_unit_programs/{name}.sialx -- the reference's src/sialx/test/{name}.sialx, verbatim, as a TEST INPUT of the\n"
            "# SIAL front-end (copied by scripts/make_unit_program_goldens.py)\n")
    open(os.path.join(ROOT, "tests", "golden", "ref_unit_programs", name + ".sialx"), "w").write(head + text)
print(len(NAMES), "programs")

# the reference's expected-output fixtures (printed blocks) for those programs, verbatim
FIX = "/root/reference/test/expected_output/"
os.makedirs(os.path.join(ROOT, "tests", "golden", "ref_expected_output"), exist_ok=True)
for name in ("static_array_test", "tmp_arrays", "tmp_arrays_2", "scalar_valued_blocks", "local_arrays", "local_arrays_wild"):
    open(os.path.join(ROOT, "tests", "golden", "ref_expected_output", name + ".txt"), "w").write(open(FIX + name + ".txt").read())
