"""ctypes binding of libdocvision.so (include/docvision.h).  No fallback: if the CUDA library is missing
or no B200 is visible, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdocvision.so")

_lib = None


class DocVisionError(RuntimeError):
    pass


# name -> (restype, argtypes); keep in sync with include/docvision.h (tests/test_abi.py checks the header)
SIGNATURES = {
    "dv_version": (C.c_int, []),
    "dv_last_error": (C.c_char_p, [C.c_void_p]),
    "dv_create": (C.c_int, [C.c_char_p, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "dv_destroy": (C.c_int, [C.c_void_p]),
    "dv_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "dv_sync": (C.c_int, [C.c_void_p]),
    "dv_launch_count": (C.c_longlong, [C.c_void_p]),
    "dv_model_flops": (C.c_double, [C.c_void_p]),
    "dv_profile_begin": (C.c_int, [C.c_void_p]),
    "dv_profile_report": (C.c_longlong, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "dv_dbnet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "dv_dbnet_forward_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                      C.POINTER(C.c_float), C.c_float, C.c_int, C.c_void_p]),
    "dv_ctc_greedy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "dv_db_boxes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_float, C.c_double,
                              C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "dv_db_boxes_dbnet": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_float, C.c_double,
                                    C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "dv_lore_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "dv_lore_gather_logi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dv_lore_detect_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "dv_lore_detect_forward_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                            C.POINTER(C.c_float), C.c_int, C.c_void_p]),
    "dv_lore_cell_features": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.POINTER(C.c_int32)]),
    "dv_lore_add_position_embeddings": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "dv_lore_process_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "dv_picodet_decode": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                    C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float, C.c_double, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_void_p]),
    "dv_picodet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "dv_picodet_forward_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                        C.c_float, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "dv_picodet_num_classes": (C.c_int, [C.c_void_p]),
    "dv_rec_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dv_rec_forward_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dv_rec_time_steps": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "dv_cls_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "dv_rec_num_classes": (C.c_int, [C.c_void_p]),
    "dv_centernet_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "dv_centernet_forward_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                          C.POINTER(C.c_float), C.c_int, C.c_void_p]),
    "dv_centernet_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.c_double), C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    "dv_convnextvit_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dv_convnextvit_forward_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dv_convnextvit_labels": (C.c_int, [C.c_void_p]),
    "dv_crnn_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dv_crnn_labels": (C.c_int, [C.c_void_p]),
    "dv_match_cells": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "dv_convnextvit_set_pass_crops": (C.c_int, [C.c_void_p, C.c_int]),
    "dv_warp_perspective_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_int, C.c_void_p]),
    "dv_resize_linear_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p]),
    "dv_crop_quads_for_rec": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "dv_crop_boxes_for_rec": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "dv_warp_affine_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "dv_crop_tables_for_tsr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p]),
    "dv_pp_rec_normalise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "dv_ctc_collapse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "dv_debug_get_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int)]),
    "dv_conv2d_nhwc_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "dv_nchw_f32_to_nhwc_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "dv_nhwc_f16_to_nchw_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
}


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DocVisionError(
            f"{LIB_PATH} not found: build it with `python -m pdf_table_b200.build` "
            "(there is no CPU / PyTorch fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error(handle=None) -> str:
    msg = load().dv_last_error(handle)
    return msg.decode(errors="replace") if msg else ""


def check(rc: int, handle=None, what: str = "") -> None:
    if rc != 0:
        raise DocVisionError(f"{what or 'libdocvision'} failed (rc={rc}): {last_error(handle)}")
