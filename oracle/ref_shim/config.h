/* oracle/ref_shim/config.h -- TEST INFRASTRUCTURE (never linked into the product).
 *
 * Stand-in for the header the reference's build system generates from config.h.cmake (four optional
 * defines: HAVE_BLAS, HAVE_CUDA, HAVE_MPI, SIP_DEVEL).  None of them is set: the reference sources that
 * oracle/Makefile compiles in place (under /root/reference) are built in their single-process, host-only
 * configuration, which is the one the reference's own unit tests (test_basic_sial.cpp) run in.
 */
#ifndef ACES4_B200_ORACLE_REF_SHIM_CONFIG_H
#define ACES4_B200_ORACLE_REF_SHIM_CONFIG_H
#endif
