/*
 * nbdt_b200.h -- C ABI of libnbdt_b200.so: the B200 (sm_100a) implementation of the NBDT hot path.
 *
 * The reference (alvinwan/neural-backed-decision-trees) is pure Python; it has no FFI.  The entry
 * points below are what a binding for the hot path replaces, one per reference call site
 * (file:line into the reference tree):
 *
 *   nbdt_tree_create        nbdt/tree.py:146-174 (Tree.__init__/load_hierarchy) + tree.py:105-125
 *                           (Node.build_class_mappings): the per-node class<->child tables,
 *                           flattened once and copied to the device.
 *   nbdt_soft_forward       nbdt/model.py:268-273 SoftEmbeddedDecisionRules.forward
 *                           (= forward_nodes model.py:101-123 + traverse_tree model.py:207-242)
 *   nbdt_soft_backward      the autograd backward of the above (the reference relies on torch autograd
 *                           through model.py:94-117,241)
 *   nbdt_hard_forward       nbdt/model.py:194-203 HardEmbeddedDecisionRules.forward_with_decisions
 *                           (= forward_nodes + traverse_tree model.py:145-186)
 *   nbdt_node_outputs       nbdt/model.py:101-120 EmbeddedDecisionRules.get_all_node_outputs
 *   nbdt_soft_loss          nbdt/loss.py:191-203 TreeSupLoss.forward with
 *                           nbdt/loss.py:260-266 SoftTreeSupLoss.forward_tree, and its autograd backward
 *   nbdt_hard_loss          nbdt/loss.py:191-203 with nbdt/loss.py:212-257 HardTreeSupLoss.forward_tree
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer owned by the caller.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*); no host sync.
 *   - return value 0 = ok, otherwise an NBDT_E* code; nbdt_last_error() gives a thread-local message.
 *   - dtype of logits / grad_logits: NBDT_F32, NBDT_BF16 or NBDT_F16.  All arithmetic is fp32.
 *   - node ids: leaf with class index c -> c ; inner node i (index into Tree.inodes, i.e. sorted by
 *     wnid, reference nbdt/tree.py:172-173) -> num_classes + i.
 *   - There is no CPU implementation behind this ABI: without a CUDA device every compute entry
 *     point fails with NBDT_ECUDA.
 */
#ifndef NBDT_B200_H
#define NBDT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBDT_B200_VERSION 105 /* 105: nbdt_bn_stats_apply (BatchNorm finalize folded into the consumer kernels); 104: class-major tree entry points (nbdt_*_nchw); 103: BatchNorm statistics from the convolution epilogue (nbdt_conv2d_fwd_stats, nbdt_bn_finalize_partials); 102: run-time specialised tree kernels (nbdt_tree_jit_*, nbdt_tree_kernel_variant); out-of-range targets are ignored rows */

enum { NBDT_OK = 0, NBDT_EINVAL = 1, NBDT_ECUDA = 2, NBDT_ETREE = 3, NBDT_ETOOLARGE = 4 };
enum { NBDT_F32 = 0, NBDT_BF16 = 1, NBDT_F16 = 2 };

typedef struct nbdt_tree nbdt_tree; /* opaque; owns the device copies of the tree tables */

typedef struct {
  int32_t num_classes;       /* C: leaves, in wnids-file order (class index)                  */
  int32_t num_inodes;        /* N: inner nodes, in Tree.inodes order                           */
  int32_t root;              /* index into inodes of the root                                  */
  const int32_t* child_ptr;  /* [N+1] CSR offsets into child_node (HOST pointer)               */
  const int32_t* child_node; /* [S] node ids of the children, in Node.children order (HOST)    */
} nbdt_tree_desc;

int nbdt_version(void);
const char* nbdt_last_error(void);
int nbdt_device_count(void);

/* Validates that desc is a rooted tree covering all leaves; derives leaf counts, depths and the
 * level schedules; copies everything to `device`.  */
int nbdt_tree_create(const nbdt_tree_desc* desc, int device, nbdt_tree** out);
void nbdt_tree_destroy(nbdt_tree* tree);
int nbdt_tree_max_depth(const nbdt_tree* tree);    /* max #decisions root->leaf                 */
/* Name of the hierarchy-specialised kernel variant the last soft-forward / loss launch on this tree used ("" when the
 * table-driven kernel ran): "ka<generator>_<hierarchy hash>_m<mode>_t<dtype>_v<vector bytes>_o<...>".  For bench lines. */
const char* nbdt_tree_kernel_variant(const nbdt_tree* tree);
int nbdt_tree_num_slots(const nbdt_tree* tree);    /* S = sum of children over inner nodes      */

/* class_probs[b,c] = prod over inner nodes n on the path to c of softmax_k(node_logits_n[b,:])[k(n,c)],
 * node_logits_n[b,k] = mean of logits[b, leaves under child k].  class_probs: fp32 [B, C] contiguous. */
int nbdt_soft_forward(const nbdt_tree* tree, const void* logits, int dtype, int64_t B, int64_t ld,
                      float* class_probs, void* stream);

/* Backward of nbdt_soft_forward for an arbitrary criterion on the probabilities:
 * grad_logits = grad_scale * (d class_probs / d logits)^T grad_probs, recomputing the forward from
 * `logits`.  grad_probs: fp32 [B, C] contiguous; grad_logits: dtype/ld of logits. */
int nbdt_soft_backward(const nbdt_tree* tree, const void* logits, int dtype, int64_t B, int64_t ld,
                       const float* grad_probs, float grad_scale, void* grad_logits, void* stream);

/* Greedy root->leaf walk by per-node argmax (first maximum).  pred: int64 [B] class index.
 * The path_* outputs may each be NULL; otherwise [B, max_depth]: child index taken (-1 past the
 * leaf), inode index the decision was taken at (-1 pad), probability of the taken child and
 * entropy of that node's child distribution.  */
int nbdt_hard_forward(const nbdt_tree* tree, const void* logits, int dtype, int64_t B, int64_t ld,
                      int64_t* pred, int32_t* path_child, int32_t* path_inode, float* path_prob,
                      float* path_entropy, void* stream);

/* Per-node outputs, packed in child-slot order (slot s of inode i is child_ptr[i] + k):
 * node_logits, node_probs: fp32 [B, S]; preds: int64 [B, N]; entropy: fp32 [B, N]. Any may be NULL. */
int nbdt_node_outputs(const nbdt_tree* tree, const void* logits, int dtype, int64_t B, int64_t ld,
                      float* node_logits, float* node_probs, int64_t* preds, float* entropy,
                      void* stream);

/* loss = (1/batch_rows) * sum_b [ w_xent * CE(logits_b, y_b) + w_tree * CE(class_probs_b, y_b) ]
 * (the second CE takes the probabilities as logits -- reference loss.py:264-266).
 * loss: device fp32 scalar (overwritten).  grad_logits: same dtype/shape/ld as logits, receives
 * grad_scale * dloss/dlogits; may be NULL (forward only).  batch_rows: the batch size the mean is
 * taken over (pass B on one GPU, the global batch when rows are sharded); <=0 means B. */
int nbdt_soft_loss(const nbdt_tree* tree, const void* logits, int dtype, const int64_t* targets,
                   int64_t B, int64_t ld, int64_t batch_rows, float w_xent, float w_tree,
                   float grad_scale, float* loss, void* grad_logits, void* stream);

/* loss = (1/batch_rows) * sum_b [ w_xent * CE(logits_b, y_b)
 *          + w_tree * tsw * (2/N) * sum_{n on path(y_b)} -log softmax(node_logits_n[b,:])[child toward y_b] ]
 * (reference loss.py:212-257: the group-wise CE means weighted by rows/(B*N/2)*tsw collapse to this). */
int nbdt_hard_loss(const nbdt_tree* tree, const void* logits, int dtype, const int64_t* targets,
                   int64_t B, int64_t ld, int64_t batch_rows, float w_xent, float w_tree, float tsw,
                   float grad_scale, float* loss, void* grad_logits, void* stream);
/* Both losses: a row whose target is outside [0, C) is ignored -- it adds nothing to the loss and its gradient row is zero
 * (nn.CrossEntropyLoss's ignore_index rows, reference loss.py:266 / model.py:127-143); the mean is still taken over
 * batch_rows, so pass the number of non-ignored rows there to reproduce CrossEntropyLoss's normalisation. */

/* Class-major variants for the segmentation wrappers: logits, gradient and class probabilities are (N, C, H, W) contiguous,
 * targets / predictions (N, H, W); row r of the hierarchy = pixel r % hw of image r / hw, hw = H*W.  They replace the
 * coerce_tensor / uncoerce_tensor permute + reshape copies (reference nbdt/utils.py:338-347) around the rules in
 * SegNBDT.forward (nbdt/model.py:376-399) and SoftSegTreeSupLoss.forward (nbdt/loss.py:318-327).  Run-time specialised kernels
 * when hw % 32 == 0, the table-driven kernel otherwise; arguments as in the row-major calls. */
int nbdt_soft_forward_nchw(const nbdt_tree* tree, const void* logits, int dtype, int64_t N, int64_t hw,
                           float* class_probs, void* stream);
int nbdt_hard_forward_nchw(const nbdt_tree* tree, const void* logits, int dtype, int64_t N, int64_t hw, int64_t* pred,
                           void* stream);
int nbdt_soft_loss_nchw(const nbdt_tree* tree, const void* logits, int dtype, const int64_t* targets, int64_t N,
                        int64_t hw, int64_t batch_rows, float w_xent, float w_tree, float grad_scale, float* loss,
                        void* grad_logits, void* stream);

/* Hierarchy-specialised kernels (csrc/tree_jit.cu).  nbdt_soft_forward / nbdt_soft_loss / nbdt_hard_loss run, for any
 * hierarchy of up to 1536 leaves, a straight-line kernel generated from the nbdt_tree_desc and compiled with NVRTC on
 * first use (cached on disk next to the library; NBDT_B200_TREE_JIT=0 disables, =cache uses only cached kernels).  The
 * two calls below need no device: the generated CUDA source (tests compile it for the host and check it against the
 * oracle; mode 0 soft loss, 1 hard loss, 2 soft forward; vec = bytes per logits vector access, 0 = the class-major variant), and ahead-of-time
 * compilation into the cache (the build does this for the shipped hierarchies). */
long long nbdt_tree_jit_source(const nbdt_tree_desc* desc, int mode, int dtype, int vec, char* buf, long long cap);
int nbdt_tree_jit_precompile(const nbdt_tree_desc* desc, int mode, int dtype);


/* ---------------------------------------------------------------------------------------------
 * Backbone: implicit-GEMM convolutions on tcgen05 tensor cores (csrc/conv_gemm.cu).  These replace the
 * cuDNN calls behind nn.Conv2d in the reference's backbones (nbdt/models/resnet.py:47-65,120 and the
 * pytorchcv WRN units, nbdt/models/wideresnet.py:1-5).  Activations: NHWC bf16, channel counts
 * multiples of 64 (pad with zero channels).  ksize in {1,3} with padding ksize/2, stride in {1,2}.
 * ------------------------------------------------------------------------------------------- */

/* y[N,H/s,W/s,Cout] = conv(x[N,H,W,Cin]; wpack) (+ residual, same shape as y; may be NULL).
 * wpack: bf16 [ksize*ksize][Cout][Cin] from nbdt_conv_pack_weights. */
int nbdt_conv2d_fwd(const void* x, const void* wpack, void* y, const void* residual, int N, int H, int W,
                    int Cin, int Cout, int ksize, int stride, void* stream);

/* The same convolution, with the BatchNorm batch statistics of y accumulated in its epilogue: every CTA writes one row of
 * per-channel partial sums, partials[cta][0][c] = sum y[...,c], partials[cta][1][c] = sum y[...,c]^2 (of the bf16 values it
 * stored).  partials: nbdt_conv2d_stats_floats(Cout) floats (0 = more channels than the fused path covers: use nbdt_bn_stats);
 * *parts (host) = rows written.  nbdt_bn_finalize_partials turns them into the consumer BatchNorm's coefficients; the separate
 * statistics pass over y is gone.  Replaces the batch-statistics half of F.batch_norm behind the nn.BatchNorm2d modules that
 * follow a convolution in the reference's backbones (nbdt/models/resnet.py:47-54,69-74; pytorchcv pre-activation WRN units). */
int nbdt_conv2d_fwd_stats(const void* x, const void* wpack, void* y, const void* residual, int N, int H, int W,
                          int Cin, int Cout, int ksize, int stride, float* partials, int* parts, void* stream);
long long nbdt_conv2d_stats_floats(int Cout);

/* dx[N,H,W,Cin] = data gradient of the above for dy[N,H/s,W/s,Cout] (+ accumulate, same shape as dx; may
 * alias dx).  wpack_t: bf16 [ksize*ksize][Cin][Cout] from nbdt_conv_pack_weights. */
int nbdt_conv2d_dgrad(const void* dy, const void* wpack_t, void* dx, const void* accumulate, int N, int H,
                      int W, int Cin, int Cout, int ksize, int stride, void* stream);

/* dwp[ksize*ksize][Cout][Cin] (fp32, overwritten) = weight gradient given x and dy. */
int nbdt_conv2d_wgrad(const void* x, const void* dy, float* dwp, int N, int H, int W, int Cin, int Cout,
                      int ksize, int stride, void* stream);

/* fp32 OIHW master weights -> bf16 tap-major matrices, zero padded to (cout_pad, cin_pad); either output
 * may be NULL.  wpack [taps][cout_pad][cin_pad], wpack_t [taps][cin_pad][cout_pad]. */
int nbdt_conv_pack_weights(const float* w_oihw, void* wpack, void* wpack_t, int Cout, int Cin, int cout_pad,
                           int cin_pad, int ksize, void* stream);

/* fp32 dwp [taps][cout_pad][cin_pad] -> OIHW gradient [Cout][Cin][k][k] (accumulate != 0: add into it). */
int nbdt_conv_unpack_wgrad(const float* dwp, float* dw_oihw, int Cout, int Cin, int cout_pad, int cin_pad,
                           int ksize, int accumulate, void* stream);

/* The same two conversions for EVERY convolution layer of a network in one launch each (one pack and one unpack launch per train step
 * instead of one per layer).  jobs_device: device array of njobs x 8 int64:
 *   pack:   {w_oihw, wpack, wpack_t, Cout, Cin, cout_pad, cin_pad, taps | first_tile << 32}, tiles = ceil(cin_pad/32) * ceil(cout_pad/32)
 *   unpack: {dw_oihw (accumulated into), dwp, 0, Cout, Cin, cout_pad, cin_pad, taps | first_tile << 32}, tiles = ceil(Cin/32) * ceil(Cout/32)
 * first_tile = running sum of the previous jobs' tile counts, total_tiles = the sum over all jobs. */
int nbdt_conv_pack_weights_batch(const void* jobs_device, int njobs, int total_tiles, void* stream);
int nbdt_conv_unpack_wgrad_batch(const void* jobs_device, int njobs, int total_tiles, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backbone: memory-bound kernels around the convolutions (csrc/bn_ops.cu).  Activations bf16 [M][C]
 * (M = N*H*W rows of an NHWC tensor), parameters / statistics fp32.  They replace nn.BatchNorm2d,
 * F.relu, the residual add, F.avg_pool2d and optim.SGD of the reference train step
 * (nbdt/models/resnet.py:50-54,70-73,142; main.py:207,239).  c_real <= C marks zero-padded channels.
 * ------------------------------------------------------------------------------------------- */

/* Floats the `sums` / `dsums` scratch buffers below must hold (reduced [2][C] + coefficients [3][C] + per-CTA partial sums). */
long long nbdt_bn_scratch_floats(int C);
/* Training-mode batch statistics: mean_invstd[2][C], scale_shift[2][C] (y = x*scale + shift), running-stat
 * update with `momentum` (running_* may be NULL).  sums: fp32 scratch of nbdt_bn_scratch_floats(C). */
int nbdt_bn_stats(const void* x, long long M, int C, int c_real, const float* gamma, const float* beta,
                  float* running_mean, float* running_var, float momentum, float eps, float* sums,
                  float* mean_invstd, float* scale_shift, void* stream);
/* nbdt_bn_stats + nbdt_bn_apply in two launches instead of three: the statistics kernel accumulates into one fp64 [2][C]
 * buffer (the head of `sums`, 8-byte aligned) and the apply kernel finalizes mean / invstd / scale / shift in its prologue,
 * writes mean_invstd / scale_shift (for the backward pass) and updates the running statistics.  The training-mode
 * nn.BatchNorm2d (+ F.relu, + residual) of the reference's blocks (nbdt/models/resnet.py:50-54,70-73). */
int nbdt_bn_stats_apply(const void* x, const void* residual, void* y, long long M, int C, int c_real, const float* gamma,
                        const float* beta, float* running_mean, float* running_var, float momentum, float eps, float* sums,
                        float* mean_invstd, float* scale_shift, int relu, void* stream);
/* The same from per-CTA partial sums [nparts][2][C] over M rows in total (written by nbdt_conv2d_fwd_stats). */
int nbdt_bn_finalize_partials(const float* partials, int nparts, long long M, int C, int c_real, const float* gamma,
                              const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                              float* mean_invstd, float* scale_shift, void* stream);
/* Eval mode: scale_shift from the running statistics. */
int nbdt_bn_eval_scale(int C, int c_real, const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, float eps, float* scale_shift, void* stream);
/* y = act(x*scale + shift (+ residual)); relu: activation code 0 none, 1 ReLU, 2 SiLU (SiLU: inference only). */
int nbdt_bn_apply(const void* x, const void* residual, void* y, long long M, int C, const float* scale_shift,
                  int relu, void* stream);
/* Backward of nbdt_bn_stats + nbdt_bn_apply.  ReLU mask of the forward pass: `act` (the forward output) when
 * given; else recomputed as x*scale+shift > 0 from `scale_shift` when given (valid when no residual was added before
 * the ReLU; saves one tensor read per kernel); else no activation.  dx = gradient w.r.t. x (+ extra, may be NULL);
 * dres (may be NULL) = gradient w.r.t. the residual; dsums: fp32 buffer of nbdt_bn_scratch_floats(C); its first
 * [2][C] receive (dbeta, dgamma). */
int nbdt_bn_backward(const void* x, const void* act, const float* scale_shift, const void* dy, const void* extra,
                     void* dx, void* dres, long long M, int C, int c_real, const float* gamma,
                     const float* mean_invstd, float* dsums, void* stream);
/* Same for y = SiLU(x*scale + shift) (no residual): the activation gradient is recomputed from x and scale_shift. */
int nbdt_bn_backward_silu(const void* x, const float* scale_shift, const void* dy, const void* extra, void* dx,
                          long long M, int C, int c_real, const float* gamma, const float* mean_invstd, float* dsums,
                          void* stream);
/* feat[N][C] (fp32) = mean over the HW rows of each image;  da = dfeat / HW broadcast. */
int nbdt_avgpool_fwd(const void* a, float* feat, int N, int HW, int C, void* stream);
int nbdt_avgpool_bwd(const float* dfeat, void* da, int N, int HW, int C, void* stream);
/* 3-channel 3x3 stem conv: x NCHW fp32 [N,3,H,W], w fp32 [cout][3][3][3] -> y NHWC bf16 [N,H,W,c_phys]
 * (channels >= cout are written as zero); and its weight gradient. */
int nbdt_stem_fwd(const float* x, const float* w, void* y, int N, int H, int W, int cout, int c_phys, void* stream);
/* same with stride 1 or 2 (y [N,ceil(H/s),ceil(W/s),c_phys]): the EfficientNet stem (torchvision efficientnet.py: Conv2d(3,32,3,stride=2)). */
int nbdt_stem_fwd_strided(const float* x, const float* w, void* y, int N, int H, int W, int cout, int c_phys, int stride,
                          void* stream);
int nbdt_stem_wgrad(const float* x, const void* dy, float* dw, int N, int H, int W, int cout, int c_phys,
                    void* stream);
int nbdt_stem_wgrad_strided(const float* x, const void* dy, float* dw, int N, int H, int W, int cout, int c_phys,
                            int stride, void* stream);
/* Fused SGD over a flat fp32 buffer: m = momentum*m + (g*grad_scale + wd*p); p -= lr*m (first_step: m = g'). */
int nbdt_sgd_step(float* params, const float* grads, float* momentum_buf, long long n, float lr, float momentum,
                  float weight_decay, float grad_scale, int first_step, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backbone: MBConv kernels (csrc/mbconv_ops.cu) for the EfficientNet backbones the reference pulls in with
 * `from torchvision.models import *` / `from pytorchcv.models.efficientnet import *` (nbdt/models/__init__.py:3-4).
 * Inference path (folded BatchNorm).
 * ------------------------------------------------------------------------------------------- */

/* Depthwise conv (groups = C) + folded BatchNorm + activation: y = act(dwconv(x; w)*scale + shift).
 * x bf16 [N,H,W,C]; w fp32 [ksize*ksize][C] (tap-major); scale_shift fp32 [2][C] or NULL; act 0 none / 1 ReLU / 2 SiLU;
 * y bf16 [N,ceil(H/stride),ceil(W/stride),C]; ksize in {3,5}, padding ksize/2, stride in {1,2}. */
int nbdt_dwconv_fwd(const void* x, const float* w, const float* scale_shift, void* y, int N, int H, int W, int C,
                    int ksize, int stride, int act, void* stream);
/* Squeeze-excitation scaling y[n,p,c] = x[n,p,c]*s[n][c]; x, y bf16 [N][HW][C] (may alias), s fp32 [N][C]. */
int nbdt_se_scale(const void* x, const float* s, void* y, int N, int HW, int C, void* stream);
/* Training path of the MBConv blocks (depthwise conv without the fused affine, i.e. nbdt_dwconv_fwd with scale_shift = NULL,
 * act = 0): data gradient dx [N,H,W,C] from dy [N,ceil(H/s),ceil(W/s),C]; weight gradient dw fp32 [ksize*ksize][C]
 * (overwritten); gradient of the squeeze-excitation scale ds[n][c] = sum_p x[n,p,c]*dy[n,p,c]. */
int nbdt_dwconv_dgrad(const void* dy, const float* w, void* dx, int N, int H, int W, int C, int ksize, int stride,
                      void* stream);
int nbdt_dwconv_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int C, int ksize, int stride,
                      void* stream);
int nbdt_se_dscale(const void* x, const void* dy, float* ds, int N, int HW, int C, void* stream);

/* Classifier head (csrc/linear.cu): C[M,N] = A B (+ bias[n]) (+ C) in fp32 with element strides A(m,k) = A[m*a_row_stride +
 * k*a_col_stride], B(k,n) = B[k*b_row_stride + n*b_col_stride] -- logits = feat W^T + b and its gradients
 * (reference nbdt/models/resnet.py:142-148) without copies for the transposes; nbdt_colsum: out[n] = sum_m X[m,n] (bias gradient). */
int nbdt_sgemm(int M, int N, int K, const float* A, long long a_row_stride, long long a_col_stride, const float* B, long long b_row_stride,
               long long b_col_stride, float* C, long long ldc, const float* bias_n, int accumulate, void* stream);
int nbdt_colsum(const float* X, int M, int N, long long ld, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Either side of the hot path (csrc/aux_ops.cu): evaluation counters, hierarchy induction, batch augmentation.
 * ------------------------------------------------------------------------------------------- */

/* Top-k accuracy without a host sync per batch (reference nbdt/metrics.py:16-20, fed by nbdt/analysis.py:224-229 with the
 * decision rules' output): counters[0] += #rows whose target is among the k largest scores (ties: lower index first, as
 * torch.topk lists them), counters[1] += #rows with a target in [0, C).  counters: two device uint64, zeroed by the caller. */
int nbdt_topk_count(const void* scores, int dtype, const int64_t* targets, int64_t B, int64_t C, int64_t ld, int k,
                    unsigned long long* counters, void* stream);

/* Ward agglomerative clustering of C points (the classifier rows, fp32 [C][D]) on the device: what
 * sklearn.cluster.AgglomerativeClustering(linkage="ward").fit(X).children_ returns in the reference's build_induced_graph
 * (nbdt/graph.py:434-438).  children: int32 [C-1][2], ids < C are points, C + s the cluster made by merge s (smaller id
 * first); heights: fp64 [C-1] merge distances (nondecreasing).  workspace: nbdt_ward_workspace_bytes(C) device bytes.
 * C <= 1024. */
long long nbdt_ward_workspace_bytes(int C);
int nbdt_ward_linkage(const float* centers, int C, int D, int32_t* children, double* heights, void* workspace, void* stream);

/* RandomCrop(H, padding=pad) + RandomHorizontalFlip + ToTensor + Normalize (reference nbdt/data/cifar.py:10-21) for a batch of
 * uint8 images resident on the device.  src: uint8 [N][H][W][3]; params: device int32 [N][3] = (dx, dy in [0, 2*pad], flip 0/1)
 * drawn by the caller; dst: fp32 [N][3][H][W]; mean3 / std3: HOST pointers to three floats. */
int nbdt_augment_crop_flip(const void* src_u8_nhwc, const int32_t* params, float* dst_nchw, int N, int H, int W, int pad,
                           const float* mean3, const float* std3, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NBDT_B200_H */
