// Interface between tree.cu and the generated hierarchy-specialised kernels (tree_spec_gen.cu).
#pragma once

enum { kSpecSoftLoss = 0, kSpecHardLoss = 1, kSpecSoftFwd = 2 };

struct SpecParams {
  const void* x;
  long long B, ld;
  const long long* targets;
  float* class_probs;  // kSpecSoftFwd
  float* loss;         // loss modes (pre-zeroed device scalar)
  void* grad;          // loss modes; may be NULL
  float w_xent, w_tree, inv_rows, grad_scale;
  long long hw;        // 0: rows are contiguous (stride ld).  > 0: class-major (N, C, H, W) logits with hw = H*W pixels per image:
                       // row r = (image r / hw, pixel r % hw), class c at image * C * hw + c * hw + pixel (segmentation variants).
                       // Only the run-time specialised kernels read it.
};

struct SpecEntry {
  unsigned long long hash;  // FNV-1a of (C, N, root, child_ptr, child_node) -- see tree_hash() in tree.cu
  int C;
  const void* fn[3][3];     // [mode][dtype: f32, bf16, f16] __global__ function pointers
};

extern const SpecEntry kSpecTable[];
extern const int kSpecCount;
