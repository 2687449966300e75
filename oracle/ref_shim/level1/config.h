/* oracle/ref_shim/level1/config.h -- TEST INFRASTRUCTURE.  Stand-in for the reference's generated config.h with the one
 * optional feature INTEGRATION.md level 1 re-enables: HAVE_CUDA (the device half of sip::Block, block.cpp:377-429). */
#ifndef ACES4_B200_ORACLE_REF_SHIM_LEVEL1_CONFIG_H
#define ACES4_B200_ORACLE_REF_SHIM_LEVEL1_CONFIG_H
#define HAVE_CUDA 1
#endif
