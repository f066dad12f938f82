// Interface between tree.cu and the second-generation generic tree kernel (tree_v2.cu).
#pragma once
#include "../../include/nbdt_b200.h"
#include "tree_spec.h"

struct TreeV2;  // DFS/BFS-renumbered tables of one hierarchy on one device

// *out stays NULL when the hierarchy is outside the kernel's limits (more than 1024 leaves) -- not an error.
int tree_v2_create(const nbdt_tree_desc* d, int device, int sm_count, int max_smem, TreeV2** out);
void tree_v2_destroy(TreeV2* t);
// mode: kSpecSoftLoss / kSpecHardLoss / kSpecSoftFwd.  Returns NBDT_OK, an error code, or -1 (not handled: use tree.cu's kernel).
int tree_v2_launch(const TreeV2* t, int mode, int dtype, const SpecParams& p, void* stream);
