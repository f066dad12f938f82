// Interface between tree.cu and the hierarchy-specialised straight-line kernels compiled at run time
// (tree_jit.cu): one thread per row, the hierarchy as compile-time structure, NVRTC -> cubin -> on-disk cache.
#pragma once
#include "../../include/nbdt_b200.h"
#include "tree_spec.h"

struct TreeJit;  // generator state + loaded modules of one hierarchy on one device

// *out stays NULL when the hierarchy is outside the generator's limits or the JIT is disabled
// (NBDT_B200_TREE_JIT=0) -- not an error.
int tree_jit_create(const nbdt_tree_desc* d, int device, int sm_count, int max_smem, TreeJit** out);
void tree_jit_destroy(TreeJit* t);
// mode: kSpecSoftLoss / kSpecHardLoss / kSpecSoftFwd.  Returns NBDT_OK, an error code, or -1 (not handled: no
// NVRTC, compile failure, unsupported alignment -- the caller falls back to the table-driven kernel).
int tree_jit_launch(TreeJit* t, int mode, int dtype, const SpecParams& p, void* stream);
// Name of the last kernel variant launched / compiled for this tree ("" if none): for bench lines and tests.
const char* tree_jit_last_variant(const TreeJit* t);
