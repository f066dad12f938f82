"""ctypes loader for libnbdt_b200.so.

The library is the only implementation of the hot path: if it is missing, or there is no CUDA
device, calls fail loudly -- there is no CPU fallback.
"""
import ctypes
import os
import threading

_PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB_PATH = os.environ.get("NBDT_B200_LIB", os.path.join(_PKG, "lib", "libnbdt_b200.so"))

F32, BF16, F16 = 0, 1, 2

c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32
c_f32 = ctypes.c_float
c_ptr = ctypes.c_void_p


class TreeDesc(ctypes.Structure):
    _fields_ = [
        ("num_classes", c_i32),
        ("num_inodes", c_i32),
        ("root", c_i32),
        ("child_ptr", ctypes.POINTER(c_i32)),
        ("child_node", ctypes.POINTER(c_i32)),
    ]


# name -> (restype, argtypes); every symbol include/nbdt_b200.h declares
SIGNATURES = {
    "nbdt_version": (ctypes.c_int, []),
    "nbdt_last_error": (ctypes.c_char_p, []),
    "nbdt_device_count": (ctypes.c_int, []),
    "nbdt_tree_create": (ctypes.c_int, [ctypes.POINTER(TreeDesc), ctypes.c_int, ctypes.POINTER(c_ptr)]),
    "nbdt_tree_destroy": (None, [c_ptr]),
    "nbdt_tree_max_depth": (ctypes.c_int, [c_ptr]),
    "nbdt_tree_num_slots": (ctypes.c_int, [c_ptr]),
    "nbdt_tree_kernel_variant": (ctypes.c_char_p, [c_ptr]),
    "nbdt_tree_jit_source": (ctypes.c_longlong, [ctypes.POINTER(TreeDesc), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_longlong]),
    "nbdt_tree_jit_precompile": (ctypes.c_int, [ctypes.POINTER(TreeDesc), ctypes.c_int, ctypes.c_int]),
    "nbdt_soft_forward": (ctypes.c_int, [c_ptr, c_ptr, ctypes.c_int, c_i64, c_i64, c_ptr, c_ptr]),
    "nbdt_soft_backward": (ctypes.c_int, [c_ptr, c_ptr, ctypes.c_int, c_i64, c_i64, c_ptr, c_f32, c_ptr, c_ptr]),
    "nbdt_hard_forward": (ctypes.c_int, [c_ptr, c_ptr, ctypes.c_int, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "nbdt_node_outputs": (ctypes.c_int, [c_ptr, c_ptr, ctypes.c_int, c_i64, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr]),
    "nbdt_soft_forward_nchw": (ctypes.c_int, [c_ptr, c_ptr, ctypes.c_int, c_i64, c_i64, c_ptr, c_ptr]),
    "nbdt_hard_forward_nchw": (ctypes.c_int, [c_ptr, c_ptr, ctypes.c_int, c_i64, c_i64, c_ptr, c_ptr]),
    "nbdt_soft_loss_nchw": (ctypes.c_int, [c_ptr, c_ptr, ctypes.c_int, c_ptr, c_i64, c_i64, c_i64, c_f32, c_f32, c_f32, c_ptr, c_ptr, c_ptr]),
    "nbdt_soft_loss": (ctypes.c_int, [c_ptr, c_ptr, ctypes.c_int, c_ptr, c_i64, c_i64, c_i64, c_f32, c_f32, c_f32, c_ptr, c_ptr, c_ptr]),
    "nbdt_hard_loss": (ctypes.c_int, [c_ptr, c_ptr, ctypes.c_int, c_ptr, c_i64, c_i64, c_i64, c_f32, c_f32, c_f32, c_f32, c_ptr, c_ptr, c_ptr]),
    "nbdt_conv2d_fwd": (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr] + [ctypes.c_int] * 7 + [c_ptr]),
    "nbdt_conv2d_fwd_stats": (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr] + [ctypes.c_int] * 7 + [c_ptr, ctypes.POINTER(ctypes.c_int), c_ptr]),
    "nbdt_conv2d_stats_floats": (ctypes.c_longlong, [ctypes.c_int]),
    "nbdt_conv2d_dgrad": (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr] + [ctypes.c_int] * 7 + [c_ptr]),
    "nbdt_conv2d_wgrad": (ctypes.c_int, [c_ptr, c_ptr, c_ptr] + [ctypes.c_int] * 7 + [c_ptr]),
    "nbdt_conv_pack_weights": (ctypes.c_int, [c_ptr, c_ptr, c_ptr] + [ctypes.c_int] * 5 + [c_ptr]),
    "nbdt_conv_pack_weights_batch": (ctypes.c_int, [c_ptr, ctypes.c_int, ctypes.c_int, c_ptr]),
    "nbdt_conv_unpack_wgrad_batch": (ctypes.c_int, [c_ptr, ctypes.c_int, ctypes.c_int, c_ptr]),
    "nbdt_conv_unpack_wgrad": (ctypes.c_int, [c_ptr, c_ptr] + [ctypes.c_int] * 6 + [c_ptr]),
    "nbdt_bn_scratch_floats": (ctypes.c_longlong, [ctypes.c_int]),
    "nbdt_bn_stats_apply": (ctypes.c_int, [c_ptr, c_ptr, c_ptr, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_f32, c_f32, c_ptr, c_ptr, c_ptr, ctypes.c_int, c_ptr]),
    "nbdt_bn_finalize_partials": (ctypes.c_int, [c_ptr, ctypes.c_int, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_f32, c_f32, c_ptr, c_ptr, c_ptr]),
    "nbdt_bn_stats": (ctypes.c_int, [c_ptr, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_f32, c_f32, c_ptr, c_ptr, c_ptr, c_ptr]),
    "nbdt_bn_eval_scale": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_ptr, c_ptr, c_ptr, c_ptr, c_f32, c_ptr, c_ptr]),
    "nbdt_bn_apply": (ctypes.c_int, [c_ptr, c_ptr, c_ptr, ctypes.c_longlong, ctypes.c_int, c_ptr, ctypes.c_int, c_ptr]),
    "nbdt_bn_backward": (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "nbdt_avgpool_fwd": (ctypes.c_int, [c_ptr, c_ptr, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_ptr]),
    "nbdt_avgpool_bwd": (ctypes.c_int, [c_ptr, c_ptr, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_ptr]),
    "nbdt_stem_fwd": (ctypes.c_int, [c_ptr, c_ptr, c_ptr] + [ctypes.c_int] * 5 + [c_ptr]),
    "nbdt_stem_fwd_strided": (ctypes.c_int, [c_ptr, c_ptr, c_ptr] + [ctypes.c_int] * 6 + [c_ptr]),
    "nbdt_dwconv_fwd": (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr] + [ctypes.c_int] * 7 + [c_ptr]),
    "nbdt_dwconv_dgrad": (ctypes.c_int, [c_ptr, c_ptr, c_ptr] + [ctypes.c_int] * 6 + [c_ptr]),
    "nbdt_dwconv_wgrad": (ctypes.c_int, [c_ptr, c_ptr, c_ptr] + [ctypes.c_int] * 6 + [c_ptr]),
    "nbdt_se_dscale": (ctypes.c_int, [c_ptr, c_ptr, c_ptr] + [ctypes.c_int] * 3 + [c_ptr]),
    "nbdt_stem_wgrad_strided": (ctypes.c_int, [c_ptr, c_ptr, c_ptr] + [ctypes.c_int] * 6 + [c_ptr]),
    "nbdt_bn_backward_silu": (ctypes.c_int, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, ctypes.c_longlong, ctypes.c_int, ctypes.c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "nbdt_se_scale": (ctypes.c_int, [c_ptr, c_ptr, c_ptr] + [ctypes.c_int] * 3 + [c_ptr]),
    "nbdt_stem_wgrad": (ctypes.c_int, [c_ptr, c_ptr, c_ptr] + [ctypes.c_int] * 5 + [c_ptr]),
    "nbdt_sgemm": (ctypes.c_int, [ctypes.c_int] * 3 + [c_ptr, ctypes.c_longlong, ctypes.c_longlong, c_ptr, ctypes.c_longlong, ctypes.c_longlong,
                                  c_ptr, ctypes.c_longlong, c_ptr, ctypes.c_int, c_ptr]),
    "nbdt_colsum": (ctypes.c_int, [c_ptr, ctypes.c_int, ctypes.c_int, ctypes.c_longlong, c_ptr, c_ptr]),
    "nbdt_topk_count": (ctypes.c_int, [c_ptr, ctypes.c_int, c_ptr, c_i64, c_i64, c_i64, ctypes.c_int, c_ptr, c_ptr]),
    "nbdt_ward_workspace_bytes": (ctypes.c_longlong, [ctypes.c_int]),
    "nbdt_ward_linkage": (ctypes.c_int, [c_ptr, ctypes.c_int, ctypes.c_int, c_ptr, c_ptr, c_ptr, c_ptr]),
    "nbdt_augment_crop_flip": (ctypes.c_int, [c_ptr, c_ptr, c_ptr, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.POINTER(c_f32), ctypes.POINTER(c_f32), c_ptr]),
    "nbdt_sgd_step": (ctypes.c_int, [c_ptr, c_ptr, c_ptr, ctypes.c_longlong, c_f32, c_f32, c_f32, c_f32, ctypes.c_int, c_ptr]),
}

_lib = None
_lock = threading.Lock()


class NBDTB200Error(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NBDTB200Error(
                f"libnbdt_b200.so not found at {LIB_PATH}. Build it with "
                "`python neural-backed-decision-trees_b200/build.py` (or __graft_entry__.build()). "
                "nbdt-b200 has no CPU or PyTorch fallback for the tree head, the losses or the backbones."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is missing
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = load().nbdt_last_error()
        raise NBDTB200Error(f"libnbdt_b200 error {rc}: {msg.decode() if msg else '?'}")


def dtype_code(dtype):
    import torch

    try:
        return {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}[dtype]
    except KeyError:
        raise TypeError(f"nbdt-b200 kernels take float32, bfloat16 or float16 logits, got {dtype}")


def require_cuda(t, what):
    if not t.is_cuda:
        raise NBDTB200Error(
            f"{what}: expected a CUDA tensor, got device '{t.device}'. nbdt-b200 runs the tree head and "
            "the tree-supervision losses as sm_100a CUDA kernels only; there is no CPU path."
        )


def stream_ptr(device):
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
