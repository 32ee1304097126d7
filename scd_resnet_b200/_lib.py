"""ctypes binding of libscd_b200.so (the C ABI declared in include/scd_b200.h).

There is no fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libscd_b200.so")

c_void_p, c_int, c_float, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol of include/scd_b200.h
PROTOTYPES = {
    "scd_abi_version": (c_int, []),
    "scd_last_error": (ctypes.c_char_p, []),
    "scd_decode_topk": (c_int, [c_void_p] * 3 + [c_int] * 5 + [c_void_p] * 7 + [c_void_p]),
    "scd_decode_topk_impl": (c_int, [c_void_p] * 3 + [c_int] * 5 + [c_void_p] * 7 + [c_int, c_void_p]),
    "scd_selftest_decode_math": (c_int, [c_void_p, c_void_p]),
    "scd_render_targets": (c_int, [c_void_p, c_void_p, c_int] + [c_void_p] * 4 + [c_void_p]),
    "scd_render_targets_npos": (c_int, [c_void_p, c_void_p, c_int] + [c_void_p] * 5 + [c_void_p]),
    "scd_centernet_loss_sparse": (c_int, [c_void_p] * 8 + [c_int] * 4 + [c_float, c_float] + [c_void_p] * 4
                                  + [c_void_p, c_size_t, c_void_p]),
    "scd_centernet_loss_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "scd_centernet_loss": (c_int, [c_void_p] * 8 + [c_int] * 4 + [c_float, c_float] + [c_void_p] * 4
                           + [c_void_p, c_size_t, c_void_p]),
    "scd_stem_fwd": (c_int, [c_void_p] * 3 + [c_int] * 3 + [c_void_p, c_void_p]),
    "scd_conv_igemm_fwd": (c_int, [c_int] + [c_void_p] * 4 + [c_int] * 6 + [c_void_p, c_void_p]),
    "scd_heads_fwd": (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p] * 3 + [c_void_p]),
    "scd_infer_weights_bytes": (c_size_t, []),
    "scd_infer_weights_layout": (c_int, [c_void_p, c_void_p, c_int]),
    "scd_infer_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "scd_resnet10_infer": (c_int, [c_void_p, c_void_p] + [c_int] * 3 + [c_void_p] * 3
                           + [c_void_p, c_size_t, c_void_p, c_void_p]),
    "scd_stem_fwd_f16": (c_int, [c_void_p] * 3 + [c_int] * 3 + [c_void_p, c_void_p]),
    "scd_conv_igemm_fwd_f16": (c_int, [c_int] + [c_void_p] * 4 + [c_int] * 6 + [c_void_p, c_void_p]),
    "scd_heads_fwd_f16": (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p] * 3 + [c_void_p]),
    "scd_resnet10_infer_f16": (c_int, [c_void_p, c_void_p] + [c_int] * 3 + [c_void_p] * 3
                               + [c_void_p, c_size_t, c_void_p, c_void_p]),
    "scd_resnet_num_convs": (c_int, [c_int, c_void_p]),
    "scd_resnet_conv_specs": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "scd_resnet_weights_bytes": (c_size_t, [c_int, c_void_p]),
    "scd_resnet_weights_layout": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int]),
    "scd_resnet_workspace_bytes": (c_size_t, [c_int, c_void_p, c_int, c_int, c_int]),
    "scd_resnet_infer": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p] + [c_int] * 3 + [c_void_p] * 3
                         + [c_void_p, c_size_t, c_void_p, c_int, c_void_p]),
    "scd_heads_fwd_c": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p] * 3 + [c_void_p]),
    "scd_heads_fwd_c_f16": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p] * 3 + [c_void_p]),
    "scd_stem_fwd_fmt": (c_int, [c_int] + [c_void_p] * 3 + [c_int] * 3 + [c_void_p, c_void_p]),
    "scd_conv_igemm_fwd_fmt": (c_int, [c_int, c_int] + [c_void_p] * 4 + [c_int] * 6 + [c_void_p, c_void_p]),
    "scd_heads_fwd_fmt": (c_int, [c_int] + [c_void_p] * 5 + [c_int] * 4 + [c_void_p] * 3 + [c_void_p]),
    "scd_probe_umma": (c_int, [c_void_p, c_int, ctypes.c_ulonglong, ctypes.c_ulonglong, ctypes.c_uint, c_int, c_int,
                                ctypes.c_ulonglong, ctypes.c_ulonglong, c_void_p, c_void_p]),
    "scd_bn_stats": (c_int, [c_void_p, c_size_t, c_int, c_void_p, c_void_p]),
    "scd_bn_stats_finalize": (c_int, [c_void_p, c_size_t, c_int, c_void_p] + [c_void_p] * 5 + [ctypes.c_double, c_float, c_float]
                              + [c_void_p] * 4 + [c_void_p, c_int, c_int, c_int, ctypes.c_uint, ctypes.c_longlong, c_void_p,
                                 c_void_p]),
    "scd_conv_igemm_fwd_bn": (c_int, [c_int] + [c_void_p] * 3 + [c_int] * 5 + [c_void_p, c_void_p] + [c_void_p] * 5
                              + [ctypes.c_double, c_float, c_float] + [c_void_p] * 4
                              + [c_void_p, c_int, c_int, c_int, ctypes.c_uint, ctypes.c_longlong, c_void_p, c_void_p]),
    "scd_stem_bn_pool_bwd": (c_int, [c_void_p] * 6 + [c_int] * 3 + [ctypes.c_double, c_void_p, c_void_p, c_int]
                             + [c_void_p, c_int, c_int, c_int, ctypes.c_uint, ctypes.c_longlong, c_void_p]
                             + [c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "scd_bn_bwd_reduce": (c_int, [c_void_p] * 7 + [c_size_t, c_int, c_void_p, c_void_p]
                          + [c_void_p, c_int, c_int, c_int, ctypes.c_uint, ctypes.c_longlong, c_void_p, c_void_p]),
    "scd_bn_finalize": (c_int, [c_void_p] * 6 + [c_int, ctypes.c_double, c_float, c_float] + [c_void_p] * 4 + [c_void_p]),
    "scd_bn_apply": (c_int, [c_void_p] * 4 + [c_int, c_size_t, c_int, c_void_p, c_void_p]),
    "scd_bn_bwd": (c_int, [c_void_p] * 7 + [c_size_t, c_int, ctypes.c_double] + [c_void_p] * 6 + [c_int, c_void_p]),
    "scd_conv_igemm_dgrad": (c_int, [c_int] + [c_void_p] * 5 + [c_int] * 5 + [c_void_p, c_void_p]),
    "scd_conv_wgrad_out_floats": (c_size_t, [c_int, c_int, c_int]),
    "scd_conv_wgrad": (c_int, [c_int, c_void_p, c_void_p] + [c_int] * 5 + [c_void_p, c_void_p]),
    "scd_stem_conv_train": (c_int, [c_void_p, c_void_p] + [c_int] * 3 + [c_void_p] * 3),
    "scd_stem_bn_relu_pool": (c_int, [c_void_p] * 3 + [c_int] * 3 + [c_void_p, c_void_p, c_void_p]),
    "scd_stem_pool_bwd": (c_int, [c_void_p] * 2 + [c_int] * 3 + [c_void_p, c_void_p]),
    "scd_heads_fwd_train": (c_int, [c_void_p] * 5 + [c_int] * 4 + [c_void_p] * 4 + [c_void_p]),
    "scd_heads_bwd": (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p] * 4 + [c_void_p]),
    "scd_heads_bwd_sparse": (c_int, [c_void_p] * 6 + [c_int] * 4 + [c_void_p] * 5 + [c_void_p]),
    "scd_heads_wgrad_sparse": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_void_p, c_void_p]),
    "scd_heads_dgrad_sparse": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_void_p, c_void_p]),
    "scd_peer_allreduce_buffer_bytes": (c_size_t, [c_int, c_int]),
    "scd_peer_allreduce_f64": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, ctypes.c_uint, ctypes.c_longlong,
                                        c_void_p, c_void_p]),
    "scd_adam_step": (c_int, [c_void_p] * 5 + [c_size_t, c_int] + [c_float] * 5 + [c_void_p]),
    "scd_gather_cast_bf16": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "scd_gather_f32": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "scd_scale_inplace": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
    "scd_augment_batch": (c_int, [c_void_p] * 3 + [c_int] + [c_void_p] * 4 + [c_int, c_float, c_float] + [c_void_p] * 3 + [c_void_p]),
    "scd_augment_batch_philox": (c_int, [c_void_p] * 3 + [c_int, c_void_p, c_int, c_float, c_float, ctypes.c_ulonglong,
                                          ctypes.c_ulonglong] + [c_void_p] * 4 + [c_void_p]),
    "scd_centernet_eval_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "scd_centernet_eval": (c_int, [c_void_p] * 8 + [c_int] * 4 + [c_float] + [c_void_p] * 3 + [c_void_p, c_size_t, c_void_p]),
    "scd_slide_geometry": (c_int, [c_int, c_int, c_void_p]),
    "scd_slide_tiles": (c_int, [c_void_p] + [c_int] * 4 + [c_void_p, c_void_p]),
    "scd_slide_tiles_u8": (c_int, [c_void_p] + [c_int] * 4 + [c_void_p, c_void_p]),
    "scd_slide_tiles_strip": (c_int, [c_void_p] + [c_int] * 8 + [c_void_p, c_void_p]),
    "scd_slide_column_span": (c_int, [c_int, c_int, c_int, c_void_p]),
    "scd_grayscale_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_size_t, c_void_p, c_void_p, c_size_t, c_void_p]),
    "scd_tiles_normalize_u8": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "scd_slide_merge": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_int, c_void_p, c_void_p]),
    "scd_copy2d_h2d": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p]),
}


class ScdError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ScdError(
            "libscd_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `python scd_resnet_b200/build.py`. There is no CPU or PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.scd_abi_version() != 2:
        raise ScdError("libscd_b200.so ABI version %d, expected 2" % lib.scd_abi_version())
    return lib


lib = _load()


def check(code, what):
    """Raise on a non-zero return code, with the library's message."""
    if code != 0:
        raise ScdError("%s failed (%d): %s" % (what, code, lib.scd_last_error().decode("utf-8", "replace")))
