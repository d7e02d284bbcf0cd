"""ctypes binding of the C ABI declared in include/crnn_b200.h (libcrnn_b200.so, built in-tree by build.py).

There is deliberately NO fallback: if the CUDA library is missing or a call fails, this raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcrnn_b200.so")

CRNN_CELL_GRU, CRNN_CELL_LSTM = 0, 1


class CrnnConfig(ctypes.Structure):
    _fields_ = [("imgh", ctypes.c_int32), ("imgw", ctypes.c_int32), ("num_classes", ctypes.c_int32), ("cell", ctypes.c_int32),
                ("n_units", ctypes.c_int32), ("time_dense", ctypes.c_int32), ("max_len", ctypes.c_int32), ("max_batch", ctypes.c_int32)]


class TensorInfo(ctypes.Structure):
    _fields_ = [("offset", ctypes.c_int64), ("numel", ctypes.c_int64), ("is_int", ctypes.c_int32)]


class Optimizer(ctypes.Structure):
    """struct crnn_optimizer (include/crnn_b200.h)."""
    _fields_ = [("kind", ctypes.c_int), ("lr", ctypes.c_float), ("beta1", ctypes.c_float), ("beta2", ctypes.c_float), ("eps", ctypes.c_float),
                ("decay", ctypes.c_float), ("momentum", ctypes.c_float), ("clipnorm", ctypes.c_float)]


# every symbol include/crnn_b200.h declares
SYMBOLS = ["crnn_last_error", "crnn_version", "crnn_workspace_bytes", "crnn_create", "crnn_destroy", "crnn_num_tensors",
           "crnn_tensor_name", "crnn_tensor_lookup", "crnn_forward", "crnn_forward_host", "crnn_train_fwd_bwd", "crnn_train_on_batch_host", "crnn_adam_step",
           "crnn_sgd_step", "crnn_get_iterations", "crnn_set_iterations", "crnn_ctc_status", "crnn_ctc_loss_grad", "crnn_ctc_greedy",
           "crnn_ctc_beam", "crnn_ctc_beam_topk", "crnn_ctc_beam_host", "crnn_ctc_greedy_host", "crnn_edit_distance", "crnn_edit_distance_host", "crnn_normalize_u8", "crnn_gemm", "crnn_debug_block_backward", "crnn_debug_materialize_blocks", "crnn_gemm_tc", "crnn_gemm_tc_dw", "crnn_gemm_tc_scratch_floats", "crnn_launch_count",
           "crnn_profile_enable", "crnn_profile_num_stages", "crnn_profile_stage_name", "crnn_profile_report",
           "crnn_profile_num_families", "crnn_profile_family_name", "crnn_profile_report2",
           "crnn_bilinear_sample", "crnn_nccl_unique_id", "crnn_comm_init_rank", "crnn_set_comm", "crnn_set_dp_fused", "crnn_comm_ranks", "crnn_allreduce_grads"]

_lib = None


class CrnnError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CrnnError(f"{LIB_PATH} not found: build it with `python crnn-ocr-lite_b200/build.py` "
                        "(or __graft_entry__.build()); there is no CPU fallback for this path")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, f32, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_uint64
    lib.crnn_last_error.restype = ctypes.c_char_p
    lib.crnn_version.restype = ctypes.c_char_p
    lib.crnn_tensor_name.restype = ctypes.c_char_p
    lib.crnn_tensor_name.argtypes = [vp, i32]
    lib.crnn_workspace_bytes.argtypes = [ctypes.POINTER(CrnnConfig), ctypes.POINTER(ctypes.c_size_t)]
    lib.crnn_create.argtypes = [ctypes.POINTER(CrnnConfig), vp, ctypes.c_size_t, ctypes.POINTER(vp)]
    lib.crnn_destroy.argtypes = [vp]
    lib.crnn_num_tensors.argtypes = [vp]
    lib.crnn_tensor_lookup.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(TensorInfo)]
    lib.crnn_forward.argtypes = [vp, vp, i32, vp, vp]
    lib.crnn_forward_host.argtypes = [vp, vp, i32, vp, vp]
    lib.crnn_train_fwd_bwd.argtypes = [vp, vp, vp, vp, vp, i32, vp, u64, vp]
    lib.crnn_train_on_batch_host.argtypes = [vp, vp, i32, f32, f32, vp, vp, vp, i32, u64, ctypes.POINTER(Optimizer), f32, vp,
                                             ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32), vp]
    lib.crnn_adam_step.argtypes = [vp, f32, f32, f32, f32, f32, f32, vp]
    lib.crnn_sgd_step.argtypes = [vp, f32, f32, f32, f32, f32, vp]
    lib.crnn_get_iterations.argtypes = [vp, ctypes.POINTER(ctypes.c_int64)]
    lib.crnn_set_iterations.argtypes = [vp, ctypes.c_int64]
    lib.crnn_ctc_status.argtypes = [vp, ctypes.POINTER(ctypes.c_int32), vp]
    lib.crnn_ctc_loss_grad.argtypes = [vp, i32, i32, i32, i32, vp, i32, vp, vp, f32, vp, vp, vp, f32, vp, vp]
    lib.crnn_ctc_greedy.argtypes = [vp, vp, i32, i32, i32, f32, vp, vp, vp, vp]
    lib.crnn_ctc_beam.argtypes = [vp, vp, i32, i32, i32, f32, i32, i32, vp, vp, vp, vp]
    lib.crnn_ctc_beam_topk.argtypes = [vp, vp, i32, i32, i32, f32, i32, i32, i32, vp, vp, vp, vp]
    lib.crnn_ctc_beam_host.argtypes = [vp, i32, i32, i32, f32, i32, i32, vp, vp, vp, vp]
    lib.crnn_ctc_greedy_host.argtypes = [vp, i32, i32, i32, f32, vp, vp, vp, vp]
    lib.crnn_normalize_u8.argtypes = [vp, vp, ctypes.c_longlong, f32, f32, vp]
    lib.crnn_edit_distance.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp]
    lib.crnn_edit_distance_host.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp]
    lib.crnn_gemm.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, i32, i32, vp]
    lib.crnn_launch_count.restype = ctypes.c_longlong
    lib.crnn_gemm_tc.argtypes = [vp, i32, vp, i32, i32, vp, i32, i32, i32, i32, vp, vp, vp, vp, vp]
    lib.crnn_debug_block_backward.argtypes = [vp, i32, vp, vp, i32, u64, vp]
    lib.crnn_debug_materialize_blocks.argtypes = [vp, vp]
    lib.crnn_gemm_tc_dw.argtypes = [vp, i32, i32, vp, i32, i32, vp, i32, i32, vp, vp, vp]
    lib.crnn_gemm_tc_scratch_floats.restype = ctypes.c_longlong
    lib.crnn_gemm_tc_scratch_floats.argtypes = [i32, i32]
    lib.crnn_profile_enable.argtypes = [vp, i32]
    lib.crnn_profile_stage_name.restype = ctypes.c_char_p
    lib.crnn_profile_stage_name.argtypes = [i32]
    lib.crnn_profile_report.argtypes = [vp, vp, vp, vp]
    lib.crnn_profile_family_name.restype = ctypes.c_char_p
    lib.crnn_profile_family_name.argtypes = [i32]
    lib.crnn_profile_report2.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    lib.crnn_bilinear_sample.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    lib.crnn_nccl_unique_id.argtypes = [vp]
    lib.crnn_comm_init_rank.argtypes = [vp, vp, i32, i32]
    lib.crnn_set_comm.argtypes = [vp, vp, i32]
    lib.crnn_set_dp_fused.argtypes = [vp, i32]
    lib.crnn_comm_ranks.argtypes = [vp]
    lib.crnn_allreduce_grads.argtypes = [vp, vp, vp]
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().crnn_last_error().decode(errors="replace")
        if status == -5:
            raise ValueError("Not enough time for target transition sequence: " + msg)
        raise CrnnError(f"libcrnn_b200 status {status}: {msg}")
