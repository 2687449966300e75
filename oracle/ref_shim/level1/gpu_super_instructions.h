/* oracle/ref_shim/level1/gpu_super_instructions.h -- TEST INFRASTRUCTURE.
 *
 * The replacement INTEGRATION.md (level 1) tells a maintainer to install for src/sip/cuda/gpu_super_instructions.h: the
 * same twelve `_gpu_*` names, now `extern "C"` with a status return, declared by the product's public header.  This
 * directory precedes the reference's src/sip/cuda on the include path of `make -C oracle ref_l1`, so the reference's
 * block.cpp (`#include "gpu_super_instructions.h"`, compiled unmodified with HAVE_CUDA) picks it up. */
#ifndef GPU_SUPER_INSTRUCTIONS_H_
#define GPU_SUPER_INSTRUCTIONS_H_
#define SIPGPU_NO_TENSORDIL_PROTOTYPES /* aces4 declares those ten itself (tensor_ops_c_prototypes.h, `int&` parameters) */
#include "sipgpu.h"
#endif
