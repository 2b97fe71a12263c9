"""ctypes binding of liblagvae.so (the C-ABI declared in include/lagvae.h).

There is no CPU fallback: importing this module without the built library, or calling a compute
entry point without a B200 (sm_100) CUDA device, raises.  PyTorch is used by callers only for
device memory and streams; nothing here takes torch types.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblagvae.so")

OK, E_ARG, E_CUDA, E_WORKSPACE = 0, 1, 2, 3
ABI_VERSION = 2
NPARAM = 13
PLAN_DEFAULT, PLAN_FORCE_SIMT, PLAN_INFERENCE = 0, 1, 2
BWD_DEFAULT, BWD_DECODER_WGRAD_NORM_ONLY = 0, 1


class LagvaeError(RuntimeError):
    pass


class TextDims(C.Structure):
    _fields_ = [("B", C.c_int32), ("T", C.c_int32), ("ns", C.c_int32), ("V", C.c_int32),
                ("ni", C.c_int32), ("nh", C.c_int32), ("nz", C.c_int32)]


class TextParams(C.Structure):
    _fields_ = [("p", C.c_void_p * NPARAM)]


class PixelBlockDims(C.Structure):
    _fields_ = [("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("Cm", C.c_int32), ("k", C.c_int32),
                ("eps", C.c_float), ("momentum", C.c_float), ("eval", C.c_int32)]


class PixelBlockParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w1", "g1", "b1", "w2", "g2", "b2", "w3", "g3", "b3", "rm1", "rv1", "rm2", "rv2", "rm3", "rv3")]


class PixelBlockGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dw1", "dg1", "db1", "dw2", "dg2", "db2", "dw3", "dg3", "db3")]


class Dropout(C.Structure):
    _fields_ = [("mode", C.c_int32), ("p_in", C.c_float), ("p_out", C.c_float),
                ("mask_in", C.c_void_p), ("mask_out", C.c_void_p), ("seed", C.c_uint64), ("seed_dev", C.c_void_p)]


# name -> (restype, argtypes); must list every symbol of include/lagvae.h (tests check this)
_vp, _i, _f, _i64, _u32, _u64, _sz = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_uint32, C.c_uint64, C.c_size_t
PROTOTYPES = {
    "lagvae_abi_version": (_i, []),
    "lagvae_last_error": (C.c_char_p, []),
    "lagvae_device_check": (_i, []),
    "lagvae_launch_count": (_i64, []),
    "lagvae_text_workspace_bytes": (_sz, [C.POINTER(TextDims), _u32]),
    "lagvae_text_plan_create": (_i, [C.POINTER(TextDims), _u32, _vp, _sz, C.POINTER(_vp)]),
    "lagvae_text_plan_destroy": (None, [_vp]),
    "lagvae_text_loss_forward": (_i, [_vp, C.POINTER(TextParams), _vp, _vp, _f, C.POINTER(Dropout),
                                      _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lagvae_text_loss_backward": (_i, [_vp, C.POINTER(TextParams), _vp, _vp, _vp, _vp,
                                       C.POINTER(TextParams), _u32, _vp]),
    "lagvae_text_encode_stats": (_i, [_vp, C.POINTER(TextParams), _vp, _vp, _vp, _vp]),
    "lagvae_text_reconstruct_error": (_i, [_vp, C.POINTER(TextParams), _vp, _vp, C.POINTER(Dropout), _vp, _vp]),
    "lagvae_text_decode_logits": (_i, [_vp, C.POINTER(TextParams), _vp, _vp, C.POINTER(Dropout), _vp, _vp]),
    "lagvae_clip_sgd_step": (_i, [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64), _i, _i, _f, _f, _i,
                                  _vp, _vp, _vp]),
    "lagvae_mi_estimate": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "lagvae_adam_table_bytes": (_sz, [C.POINTER(_i64), _i]),
    "lagvae_adam_table_create": (_i, [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64), _i, _i, _i,
                                      _vp, _sz, _vp, C.POINTER(_vp)]),
    "lagvae_adam_table_destroy": (None, [_vp]),
    "lagvae_clip_adam_step": (_i, [_vp, _f, _f, _f, _f, _f, _i, _vp, _vp]),
    "lagvae_comm_load": (_i, [C.c_char_p]),
    "lagvae_comm_unique_id": (_i, [_vp]),
    "lagvae_comm_init": (_i, [_vp, _i, _i, C.POINTER(_vp)]),
    "lagvae_allreduce_bucket": (_i, [_vp, _vp, _i64, _vp]),
    "lagvae_comm_destroy": (None, [_vp]),
    "lagvae_text_param_count": (_i64, [C.POINTER(TextDims)]),
    "lagvae_text_inner_step": (_i, [_vp, C.POINTER(TextParams), _vp, _vp, _f, C.POINTER(Dropout), _f, _f,
                                    _vp, _vp, _vp, _vp]),
    "lagvae_text_outer_step": (_i, [_vp, C.POINTER(TextParams), _vp, _vp, _f, C.POINTER(Dropout), _f, _f, _i,
                                    _vp, _vp, _vp, _vp]),
    "lagvae_text_decoder_weights_epoch": (_i, [_vp, _u64]),
    "lagvae_text_decoder_grads_event": (_i, [_vp, _i]),
    "lagvae_text_wait_decoder_grads": (_i, [_vp, _vp]),
    "lagvae_lstm_workspace_bytes": (_sz, [_i, _i]),
    "lagvae_lstm_forward": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(Dropout), _vp, _sz, _vp]),
    "lagvae_lstm_backward": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(Dropout), _vp, _vp, _vp, _i,
                                  _vp, _sz, _vp]),
    "lagvae_debug_trace_buffer": (None, [_vp, _sz]),
    "lagvae_lstm_variant": (C.c_char_p, [_i]),
    "lagvae_im2col": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "lagvae_col2im": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "lagvae_bn_train_fwd": (_i, [_vp, _i64, _i, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lagvae_bn_train_bwd": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lagvae_bn_apply": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lagvae_elu_fwd": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "lagvae_elu_bwd": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "lagvae_add": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "lagvae_bernoulli_nll_fwd": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "lagvae_bernoulli_nll_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "lagvae_reparam_kl_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "lagvae_reparam_kl_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "lagvae_gemm_auto_scratch_bytes": (_sz, [_i, _i, _i]),
    "lagvae_gemm_auto": (_i, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i, _i, _i, _f, _f, _vp, _vp, _sz, _vp]),
    "lagvae_gemm_f32": (_i, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i, _i, _i, _f, _f, _vp, _vp, _i, _vp]),
    "lagvae_gemm_tc": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _i64, _i, _vp, _i64, _i, _i, _i, _i, _f, _f,
                            _vp, _vp, _i, _vp, _vp]),
    "lagvae_split_bf16": (_i, [_vp, _i64, _i, _i, _vp, _vp, _i64, _vp]),
    "lagvae_convtc_supported": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "lagvae_split_cat": (_i, [_vp, _i64, _i, _vp, _vp]),
    "lagvae_convtc_wbuf_bytes": (_sz, [_i, _i, _i, _i]),
    "lagvae_convtc_prepare_weights": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "lagvae_convtc_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "lagvae_convtc_dgrad": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "lagvae_convtc_wgrad_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "lagvae_convtc_wgrad": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "lagvae_bnact_fwd": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _f, _f, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lagvae_bn_eval_stats": (_i, [_vp, _vp, _i64, _i, _vp, _vp]),
    "lagvae_bnact_bwd": (_i, [_vp, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lagvae_pixelblock_stash_bytes": (_sz, [C.POINTER(PixelBlockDims)]),
    "lagvae_pixelblock_scratch_bytes": (_sz, [C.POINTER(PixelBlockDims)]),
    "lagvae_pixelblock_forward": (_i, [C.POINTER(PixelBlockDims), C.POINTER(PixelBlockParams), _vp, _vp, _vp, _vp, _vp, _vp]),
    "lagvae_pixelblock_backward": (_i, [C.POINTER(PixelBlockDims), C.POINTER(PixelBlockParams), _vp, _vp, _vp, _vp, _vp,
                                        C.POINTER(PixelBlockGrads), _vp, _vp]),
    "lagvae_dropout_mask": (_i, [_u64, _u32, _i64, _f, _vp, _vp]),
}

_lib = None


def lib():
    """Load (once) and return the ctypes library; raises LagvaeError when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LagvaeError(
                "liblagvae.so is not built (%s). Run `python vae-lagging-encoder_b200/build.py` "
                "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError if a declared symbol is missing
            fn.restype = res
            fn.argtypes = args
        if L.lagvae_abi_version() != ABI_VERSION:
            raise LagvaeError("liblagvae.so ABI %d != binding ABI %d" % (L.lagvae_abi_version(), ABI_VERSION))
        _lib = L
    return _lib


def check(status, what=""):
    if status != OK:
        msg = lib().lagvae_last_error()
        raise LagvaeError("%s failed (status %d): %s" % (what or "lagvae call", status,
                                                         msg.decode() if msg else "?"))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def launch_count():
    return int(lib().lagvae_launch_count())


def lstm_variant():
    """{'forward': ..., 'backward': ...}: the recurrence kernels the most recent LSTM launches used (lagvae.h)."""
    L = lib()
    return {"forward": L.lagvae_lstm_variant(0).decode(), "backward": L.lagvae_lstm_variant(1).decode()}
