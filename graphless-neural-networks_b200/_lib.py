"""ctypes binding of libglnn_b200.so (include/glnn_b200.h).  There is no CPU fallback: every call
raises if the library is missing or the device is not sm_100."""
import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GLNN_B200_LIB") or os.path.join(HERE, "lib", "libglnn_b200.so")

_lock = threading.Lock()
_lib = None

c_i32, c_i64, c_f32, c_vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class MlpDesc(C.Structure):
    _fields_ = [("num_layers", c_i32), ("feat_dim", c_i32), ("hidden_dim", c_i32),
                ("label_dim", c_i32), ("norm", c_i32), ("dropout", c_f32), ("bn_eps", c_f32),
                ("bn_momentum", c_f32)]


class AdamHParams(C.Structure):
    _fields_ = [("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
                ("weight_decay", C.c_double)]


class GnnLayer(C.Structure):
    _fields_ = [("weight", c_vp), ("bias", c_vp), ("bn_scale", c_vp), ("bn_shift", c_vp),
                ("d_in", c_i32), ("d_out", c_i32)]


class SpmmDesc(C.Structure):
    _fields_ = [("indptr", c_vp), ("indices", c_vp), ("indptr64", c_i32), ("d", c_i32),
                ("n_dst", c_i64), ("n_src", c_i64), ("X", c_vp), ("ldx", c_i64), ("X_q24", c_vp),
                ("ldq", c_i64), ("Y", c_vp), ("ldy", c_i64), ("Y_hi", c_vp), ("Y_lo", c_vp),
                ("ldyp", c_i64), ("self_add", c_i32), ("mean_plus_one", c_i32), ("src_scale", c_vp),
                ("dst_scale", c_vp), ("bias", c_vp), ("col_scale", c_vp), ("col_shift", c_vp),
                ("relu", c_i32), ("log_softmax", c_i32), ("hot_below", c_i32), ("reserved", c_i32),
                ("Y_init", c_vp), ("ldyi", c_i64)]


class DpGroup(C.Structure):
    _fields_ = [("world", c_i32), ("rank", c_i32), ("base", c_vp * 8), ("bytes", c_i64)]


class ActDesc(C.Structure):
    _fields_ = [("n", c_i64), ("d", c_i32), ("relu_post", c_i32), ("X", c_vp), ("ldx", c_i64),
                ("Y", c_vp), ("ldy", c_i64), ("gamma", c_vp), ("beta", c_vp), ("running_mean", c_vp),
                ("running_var", c_vp), ("save_mean", c_vp), ("save_invstd", c_vp), ("eps", c_f32),
                ("momentum", c_f32), ("p_drop", c_f32), ("relu_input", c_i32), ("seed", C.c_uint64),
                ("keep_mask", c_vp)]


class SageLayerHost(C.Structure):
    _fields_ = [("weight", c_vp), ("bias", c_vp), ("bn_gamma", c_vp), ("bn_beta", c_vp),
                ("bn_mean", c_vp), ("bn_var", c_vp), ("d_in", c_i32), ("d_out", c_i32)]


# name -> (restype, argtypes); kept in the order of include/glnn_b200.h
SIGNATURES = {
    "glnn_version": (C.c_int, []),
    "glnn_last_error": (C.c_char_p, []),
    "glnn_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "glnn_spmm_csr_f32": (C.c_int, [c_vp, C.c_int, c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_i64,
                                    C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int,
                                    c_vp]),
    "glnn_spmm_csr_planes": (C.c_int, [c_vp, C.c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_i64,
                                       C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                       C.c_int, c_vp]),
    "glnn_gemm_f32": (C.c_int, [c_vp, c_i64, C.c_int, c_vp, c_i64, C.c_int, c_vp, c_i64, c_i64, c_i64,
                                c_i64, c_vp, c_vp, c_vp, c_vp, C.c_int, C.c_int, c_vp]),
    "glnn_split_planes_f32": (C.c_int, [c_vp, c_i64, c_i64, C.c_int, c_vp, c_vp, c_i64, c_vp]),
    "glnn_gemm_bf16x3_planes": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, c_vp, c_vp, c_i64, C.c_int, c_vp,
                                          c_i64, c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp,
                                          c_vp, c_vp, C.c_int, c_vp]),
    "glnn_q24_row_bytes": (c_i64, [C.c_int]),
    "glnn_quantize_q24_f32": (C.c_int, [c_vp, c_i64, c_i64, C.c_int, c_vp, c_i64, c_vp]),
    "glnn_gemm_bf16x3_planes_q24": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, C.c_int, c_vp, c_i64,
                                              c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, C.c_int,
                                              c_vp]),
    "glnn_spmm_csr_q24_planes": (C.c_int, [c_vp, C.c_int, c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, c_i64,
                                           C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp]),
    "glnn_spmm_csr": (C.c_int, [C.POINTER(SpmmDesc), c_vp]),
    "glnn_s24_row_words": (c_i64, [C.c_int]),
    "glnn_compact_s24": (C.c_int, [c_vp, c_i64, c_i64, C.c_int, c_vp, c_i64, c_vp, c_vp]),
    "glnn_spmm_csr_s24": (C.c_int, [C.POINTER(SpmmDesc), c_vp, c_i64, c_vp, c_vp]),
    "glnn_exp_spmm_tma_q24": (C.c_int, [C.POINTER(SpmmDesc), C.c_int, c_vp]),
    "glnn_bn_fold_f32": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, C.c_int, c_vp]),
    "glnn_log_softmax_f32": (C.c_int, [c_vp, c_i64, c_vp, c_i64, c_i64, C.c_int, c_vp]),
    "glnn_nll_acc_f32": (C.c_int, [c_vp, c_i64, C.c_int, c_vp, c_vp, c_i64, c_vp, c_vp]),
    "glnn_mlp_param_count": (c_i64, [C.POINTER(MlpDesc)]),
    "glnn_mlp_bn_stat_count": (c_i64, [C.POINTER(MlpDesc)]),
    "glnn_mlp_workspace_bytes": (c_i64, [C.POINTER(MlpDesc), c_i64]),
    "glnn_mlp_train_pass": (C.c_int, [C.POINTER(MlpDesc), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_i64,
                                      C.POINTER(AdamHParams), c_vp, c_i64, c_vp, C.c_int, c_vp, c_i64,
                                      c_i64, c_vp, C.c_uint64, c_f32, c_vp, c_vp, c_i64, c_vp]),
    "glnn_adam_step_f32": (C.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_i64, C.POINTER(AdamHParams), c_vp]),
    "glnn_mlp_dp_control_bytes": (c_i64, [C.POINTER(MlpDesc), c_i64, C.c_int]),
    "glnn_mlp_dp_flat_count": (c_i64, [C.POINTER(MlpDesc), C.c_int]),
    "glnn_mlp_dp_init": (C.c_int, [c_vp, c_i64, c_vp]),
    "glnn_mlp_train_pass_dp": (C.c_int, [C.POINTER(DpGroup), C.POINTER(MlpDesc), c_vp, c_vp, c_vp, c_vp,
                                         c_vp, c_vp, c_i64, C.POINTER(AdamHParams), c_vp, c_i64, c_vp,
                                         C.c_int, c_vp, c_i64, c_i64, c_vp, C.c_uint64, c_f32, c_vp,
                                         c_vp, c_i64, c_vp]),
    "glnn_mlp_eval": (C.c_int, [C.POINTER(MlpDesc), c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64,
                                C.c_int, c_i64, c_vp, c_i64, c_vp]),
    "glnn_gnn_forward_workspace_bytes": (c_i64, [c_i64, C.POINTER(GnnLayer), C.c_int]),
    "glnn_sage_forward": (C.c_int, [c_vp, C.c_int, c_vp, c_i64, c_vp, c_i64, C.POINTER(GnnLayer),
                                    C.c_int, c_vp, c_i64, C.c_int, c_vp, c_i64, c_vp]),
    "glnn_gcn_forward": (C.c_int, [c_vp, C.c_int, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64,
                                   C.POINTER(GnnLayer), C.c_int, c_vp, c_i64, C.c_int, c_vp, c_i64,
                                   c_vp]),
    "glnn_peer_push": (C.c_int, [c_vp, C.POINTER(c_vp), C.c_int, c_i64, C.c_int, c_vp]),
    "glnn_nll_loss_grad_f32": (C.c_int, [c_vp, c_i64, C.c_int, c_vp, c_vp, c_vp, c_i64, c_f32, c_vp, c_i64,
                                         c_vp, c_vp]),
    "glnn_act_train_fwd_f32": (C.c_int, [C.POINTER(ActDesc), c_vp, c_vp]),
    "glnn_act_train_bwd_f32": (C.c_int, [C.POINTER(ActDesc), c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp,
                                         c_vp, c_vp]),
    "glnn_spmm_csr_scatter_f32": (C.c_int, [c_vp, C.c_int, c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_i64,
                                            C.c_int, C.c_int, c_vp]),
    "glnn_sample_count": (C.c_int, [c_vp, C.c_int, c_vp, c_i64, C.c_int, c_vp, c_vp]),
    "glnn_sample_neighbors": (C.c_int, [c_vp, C.c_int, c_vp, c_vp, c_i64, C.c_int, C.c_uint64, c_vp, c_vp,
                                        c_vp]),
    "glnn_block_mark": (C.c_int, [c_vp, c_i64, c_vp, c_vp]),
    "glnn_block_relabel": (C.c_int, [c_vp, c_i64, c_vp, c_vp]),
    "glnn_csr_build_workspace_bytes": (C.c_size_t, [c_i64, c_i64]),
    "glnn_csr_from_coo": (C.c_int, [c_vp, c_vp, C.c_int, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp,
                                    C.c_size_t, c_vp]),
    "glnn_csr_subgraph": (C.c_int, [c_vp, C.c_int, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp,
                                    C.c_size_t, c_vp]),
    "glnn_sage_inference_host": (C.c_int, [c_vp, c_vp, c_i64, c_vp, C.POINTER(SageLayerHost),
                                           C.c_int, c_f32, c_vp]),
}


class GlnnError(RuntimeError):
    pass


def load():
    """Loads the shared library once.  Raises (never falls back) when it is missing."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise GlnnError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  glnn_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = lib
        return lib


def check(rc, what):
    if rc == 0:
        return
    msg = load().glnn_last_error().decode("utf-8", "replace")
    if rc < 0:
        raise ValueError(f"{what}: {msg} (code {rc})")
    raise GlnnError(f"{what}: {msg} (cudaError {rc})")


def ptr(t):
    """Device (or host) address of a tensor, None -> NULL."""
    return None if t is None else t.data_ptr()


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise GlnnError("glnn_b200 kernels need CUDA tensors; there is no CPU fallback")
