"""ctypes binding of include/densereg.h (libdensereg_sm100.so).  Fails loudly if the library is
missing -- there is no fallback path."""
import ctypes as C
import os

_LIB = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libdensereg_sm100.so")

DR_PREC_FP32, DR_PREC_TF32, DR_PREC_TF32X3 = 0, 1, 2
PRECISIONS = {"fp32": DR_PREC_FP32, "tf32": DR_PREC_TF32, "tf32x3": DR_PREC_TF32X3}


class DrConfig(C.Structure):
    _fields_ = [("num_stack", C.c_int32), ("num_fea", C.c_int32), ("kernel_size", C.c_int32),
                ("num_jnt", C.c_int32), ("in_hw", C.c_int32), ("out_hw", C.c_int32),
                ("max_batch", C.c_int32), ("precision", C.c_int32), ("device", C.c_int32),
                ("reserved", C.c_int32 * 7)]


class DrLayerInfo(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("k", C.c_int32), ("stride", C.c_int32), ("cin", C.c_int32),
                ("cout", C.c_int32), ("brn", C.c_int32), ("relu", C.c_int32), ("wd", C.c_float),
                ("w_off", C.c_int64), ("p_off", C.c_int64), ("s_off", C.c_int64),
                ("in_hw", C.c_int32), ("out_hw", C.c_int32)]


class DrOpInfo(C.Structure):
    _fields_ = [("kind", C.c_int32), ("lane", C.c_int32), ("layer", C.c_int32), ("need_dgrad", C.c_int32), ("raw_buf", C.c_int32),
                ("in_buf", C.c_int32), ("in_c0", C.c_int32), ("in_c", C.c_int32),
                ("out_buf", C.c_int32), ("out_c0", C.c_int32), ("out_c", C.c_int32),
                ("res_buf", C.c_int32), ("res_c0", C.c_int32), ("res_c", C.c_int32),
                ("nwait", C.c_int32 * 2), ("wait_op", (C.c_int32 * 3) * 2), ("record", C.c_int32 * 2),
                ("gsrc_buf", C.c_int32), ("gsrc_c0", C.c_int32), ("gsrc_c", C.c_int32),
                ("dres_buf", C.c_int32), ("dres_c0", C.c_int32), ("dres_c", C.c_int32),
                ("res_grad_fused", C.c_int32), ("in_grad_fused", C.c_int32)]


class DrTraceRec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("B", C.c_int32), ("hw", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32), ("k", C.c_int32),
                ("kernel", C.c_int32), ("ms", C.c_float)]


class DrPipeOp(C.Structure):         # include/densereg.h: dr_pipe_op
    _fields_ = [("kind", C.c_int32), ("stream", C.c_int32), ("event", C.c_int32)]


# every symbol include/densereg.h declares: name -> (restype, argtypes)
_P, _F, _I32P = C.c_void_p, C.c_void_p, C.c_void_p
SIGNATURES = {
    "dr_version": (C.c_int, []),
    "dr_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(DrConfig)]),
    "dr_destroy": (C.c_int, [_P]),
    "dr_last_error": (C.c_char_p, [_P]),
    "dr_param_count": (C.c_size_t, [_P]),
    "dr_state_count": (C.c_size_t, [_P]),
    "dr_num_layers": (C.c_int, [_P]),
    "dr_get_layer": (C.c_int, [_P, C.c_int, C.POINTER(DrLayerInfo)]),
    "dr_bind": (C.c_int, [_P, _F, _F, _F, _F, _F]),
    "dr_params_changed": (C.c_int, [_P]),
    "dr_init_params": (C.c_int, [_P, C.c_uint64, C.c_float, _P]),
    "dr_norm_dm": (C.c_int, [_P, C.c_int, C.c_int, _F, _F, _F, _P]),
    "dr_forward": (C.c_int, [_P, C.c_int, _F, _F, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                             C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_uint64, _P]),
    "dr_vote": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _F, _F, _F, _F, _F, _F, _F, _I32P, _I32P, _P]),
    "dr_infer": (C.c_int, [_P, C.c_int, _F, _F, _F, _F, _I32P, _P]),
    "dr_loss_backward": (C.c_int, [_P, C.c_int, _F, _F, _F, _F, _F, C.c_uint64, C.c_int, _P]),
    "dr_pipeline_join": (C.c_int, [_P, _P]),
    "dr_pipeline_depth": (C.c_int, [_P]),
    "dr_debug_pipeline_plan": (C.c_int, [_P, C.c_int, C.POINTER(DrPipeOp), C.c_int]),
    "dr_comm_unique_id": (C.c_int, [_P]),
    "dr_comm_init": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "dr_comm_overlap_next_backward": (C.c_int, [_P]),
    "dr_comm_allreduce_count": (C.c_int64, [_P]),
    "dr_zero_grads": (C.c_int, [_P, _P]),
    "dr_optimizer_step": (C.c_int, [_P, C.c_int, C.c_int, C.c_float, C.c_int64, _P]),
    "dr_crop_from_xyz_pose": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _F, _F, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_float,
                                        C.c_int, _F, _F, _F, _P]),
    "dr_crop_from_bbx": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _F, _F, C.POINTER(C.c_float), C.c_int, _F, _F, _F, _P]),
    "dr_data_aug": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _F, _F, _F, _F, _F, _F, _F, _F, _P]),
    "dr_debug_conv": (C.c_int, [_P, C.c_int, C.c_int, _F, _F, C.c_int, _P]),
    "dr_debug_conv_bwd": (C.c_int, [_P, C.c_int, C.c_int, _F, _F, _F, _F, C.c_int, _P]),
    "dr_debug_get_output": (C.c_int, [_P, C.c_int, C.c_int, _F, C.c_int, _P]),
    "dr_num_ops": (C.c_int, [_P]),
    "dr_debug_op": (C.c_int, [_P, C.c_int, C.POINTER(DrOpInfo)]),
    "dr_trace": (C.c_int, [_P, C.c_int]),
    "dr_trace_count": (C.c_int, [_P]),
    "dr_trace_get": (C.c_int, [_P, C.c_int, C.POINTER(DrTraceRec)]),
    "dr_launch_count": (C.c_int64, [_P]),
    "dr_tc_launch_count": (C.c_int64, [_P]),
    "dr_workspace_bytes": (C.c_size_t, [_P]),
}


def load():
    """dlopen the in-tree library and bind every declared symbol."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "densereg_b200: %s is missing -- run `python __graft_entry__.py` (nvcc, sm_100a) first. "
            "There is no CPU / PyTorch fallback for the hot path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib
