"""ctypes binding of libtd3d.so (the C ABI declared in include/td3d.h).

PyTorch is used for device memory, streams and torch.distributed only; every compute call goes
through this file into hand-written sm_100a kernels.  There is no CPU fallback: if the shared
library is missing, or the device is not a B200, calls raise.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtd3d.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_HSWISH, ACT_SILU = 0, 1, 2, 3
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05 = 0, 1, 2
OPT_SGD, OPT_ADAMW, OPT_RMSPROP, OPT_ADADELTA = 0, 1, 2, 3


class BlockDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("kernel", "stride", "in_ch", "exp_ch", "out_ch", "use_se", "se_hidden", "use_hs", "name_stage", "name_index")]


class NetDesc(C.Structure):
    _fields_ = [("arch", C.c_int), ("stem_ch", C.c_int), ("n_blocks", C.c_int), ("blocks", C.POINTER(BlockDesc)),
                ("last_ch", C.c_int), ("head_ch", C.c_int), ("num_classes", C.c_int),
                ("max_classes", C.c_int), ("num_points", C.c_int)]


class Sizes(C.Structure):
    _fields_ = [(n, C.c_int64) for n in
                ("n_param_tensors", "param_floats", "n_bn", "bn_floats", "packed_bytes",
                 "workspace_bytes", "head_param_offset", "head_param_stride")]


class ParamInfo(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("offset", C.c_int64), ("numel", C.c_int64),
                ("ndim", C.c_int32), ("shape", C.c_int64 * 4), ("bn_index", C.c_int32)]


class BnInfo(C.Structure):
    _fields_ = [("name", C.c_char * 96), ("offset", C.c_int64), ("channels", C.c_int32)]


class LossDesc(C.Structure):
    _fields_ = [(n, C.c_float) for n in
                ("w_l1", "w_smoothl1", "w_mse", "w_add", "w_diag", "w_wing", "w_ce",
                 "smoothl1_beta", "wing_w", "wing_eps")]


class OptimDesc(C.Structure):
    _fields_ = [("kind", C.c_int), ("lr", C.c_float), ("weight_decay", C.c_float),
                ("momentum", C.c_float), ("nesterov", C.c_int), ("beta1", C.c_float),
                ("beta2", C.c_float), ("eps", C.c_float), ("alpha", C.c_float), ("rho", C.c_float),
                ("grad_scale", C.c_float)]


# every symbol include/td3d.h declares (tests/test_cabi.py checks the .so exports all of them)
SYMBOLS = [
    "td3d_abi_version", "td3d_last_error", "td3d_device_check",
    "td3d_plan_create", "td3d_plan_destroy", "td3d_plan_sizes", "td3d_plan_param_info",
    "td3d_plan_bn_info", "td3d_plan_bind", "td3d_plan_set_dropout_counter", "td3d_plan_profile", "td3d_plan_profile_read", "td3d_plan_profile_launch",
    "td3d_pack_weights",
    "td3d_forward", "td3d_forward_export", "td3d_backward_stages", "td3d_backward",
    "td3d_backward_ready_range", "td3d_loss_fwd_bwd", "td3d_metrics_accum", "td3d_optim_step", "td3d_roi_crop_resize",
    "td3d_lift_2d", "td3d_iou_2d_based", "td3d_k_stem_fwd", "td3d_k_stem_wgrad", "td3d_k_dw_fwd", "td3d_k_dw_bwd", "td3d_k_dw_fwd_ex", "td3d_k_gemm_nt",
    "td3d_k_gemm_tn", "td3d_k_apply_xform", "td3d_k_affine2", "td3d_k_act_bwd_stats", "td3d_debug_tc_timeline",
]

_lib = None


class Td3dError(RuntimeError):
    pass


def lib():
    """Load libtd3d.so (built in-tree by __graft_entry__.build / csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Td3dError(f"{LIB_PATH} not found: build it with `make -C csrc` "
                            "(there is no CPU or PyTorch fallback for the td3d kernels)")
        L = C.CDLL(LIB_PATH)
        L.td3d_last_error.restype = C.c_char_p
        L.td3d_plan_destroy.restype = None
        for s in SYMBOLS:
            getattr(L, s)
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise Td3dError(f"td3d error {rc}: {lib().td3d_last_error().decode()}")


def require_b200():
    if not torch.cuda.is_available():
        raise Td3dError("td3d needs a CUDA device (sm_100a / B200); there is no CPU fallback")
    check(lib().td3d_device_check())


def ptr(t):
    """Device pointer of a tensor (or NULL)."""
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "td3d expects contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(dtype):
    if dtype is torch.float32 or (isinstance(dtype, str) and dtype in ("fp32", "f32", "float32")):
        return F32
    if dtype is torch.bfloat16 or (isinstance(dtype, str) and dtype in ("bf16", "bfloat16")):
        return BF16
    if isinstance(dtype, int) and not isinstance(dtype, bool) and dtype in (F32, BF16):
        return dtype
    raise ValueError(f"unsupported compute dtype {dtype!r} (fp32 or bf16)")


def torch_dtype(code):
    return torch.bfloat16 if code == BF16 else torch.float32
