"""ctypes binding of libdmf_b200.so (C-ABI declared in include/dmf_b200.h).

There is no fallback: if the shared library is missing this module raises ImportError, and every
compute entry point raises RuntimeError(dmf_last_error()) on failure.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libdmf_b200.so')

U8, U16, F32, F64 = 0, 1, 2, 3

vp, i32, i64, cstr = C.c_void_p, C.c_int, C.c_int64, C.c_char_p

# name -> (restype, argtypes); every symbol include/dmf_b200.h declares
SIGNATURES = {
    'dmf_abi_version': (i32, []),
    'dmf_last_error': (cstr, []),
    'dmf_launch_count': (i64, []),
    'dmf_normalize_pad': (i32, [vp, i32, i32, i32, i32, i32, vp, i32, vp]),
    'dmf_scene_create_raw': (i32, [C.POINTER(vp), vp, i32, vp, i32, i32, i32, i32, i32, vp]),
    'dmf_scene_update_raw': (i32, [vp, vp, i32, vp, i32, i32, vp]),
    'dmf_raster_minmax': (i32, [vp, i32, i64, vp, vp]),
    'dmf_scene_update_raw_range': (i32, [vp, vp, i32, vp, i32, i32, vp, vp, vp]),
    'dmf_scene_create_padded': (i32, [C.POINTER(vp), vp, vp, i32, i32, i32, i32, i32, vp]),
    'dmf_scene_set_mspan': (i32, [vp, vp, i32, i32, vp]),
    'dmf_scene_set_mspan_ihs': (i32, [vp, vp, i32, vp, i32, i32, vp, vp, vp, vp]),
    'dmf_scene_set_labels': (i32, [vp, vp, i32, vp]),
    'dmf_scene_destroy': (i32, [vp]),
    'dmf_scene_dims': (i32, [vp, C.POINTER(C.c_int32)]),
    'dmf_scene_export': (i32, [vp, i32, vp, vp]),
    'dmf_gather': (i32, [vp, vp, i64, vp, vp, vp, vp, vp]),
    'dmf_ihs_tran': (i32, [vp, vp, vp, vp, i32, i32, vp]),
    'dmf_pan2ms': (i32, [vp, i32, i32, i32, vp, vp]),
    'dmf_net_create': (i32, [C.POINTER(vp), i32, i32, i32]),
    'dmf_net_destroy': (i32, [vp]),
    'dmf_net_load_param': (i32, [vp, cstr, vp, i64]),
    'dmf_net_finalize': (i32, [vp, vp]),
    'dmf_net_flops_per_patch': (i64, [vp]),
    'dmf_net_forward_patches': (i32, [vp, vp, vp, i64, vp, vp]),
    'dmf_net_forward_scene': (i32, [vp, vp, vp, i64, i64, vp, vp, vp, vp, vp]),
    'dmf_infer_scene': (i32, [vp, vp, i32, i32, vp, vp, vp]),
    'dmf_infer_scene_dense': (i32, [vp, vp, i32, i32, vp, vp, vp, vp]),
    'dmf_net_set_dense': (i32, [vp, i32, i32]),
    'dmf_net_get_dense_timing': (i32, [vp, C.POINTER(C.c_float), i32]),
    'dmf_dense_class_table': (i32, [i32, i32, i32, vp, vp, vp, vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    'dmf_net_dense_buffer': (i32, [vp, cstr, C.POINTER(vp), C.POINTER(i64), C.POINTER(C.c_int32)]),
    'dmf_net_set_pan_source': (i32, [vp, i32]),
    'dmf_net_set_timing': (i32, [vp, i32]),
    'dmf_net_get_timing': (i32, [vp, C.POINTER(C.c_float)]),
    'dmf_net_debug_layer': (i32, [vp, i32, i32, vp, vp, i64, vp]),
    'dmf_net_debug_stem': (i32, [vp, i32, vp, vp, i64, vp]),
    'dmf_train_create': (i32, [C.POINTER(vp), i32, i32, i32]),
    'dmf_train_destroy': (i32, [vp]),
    'dmf_train_bind': (i32, [vp, cstr, vp, vp, i64]),
    'dmf_train_finalize': (i32, [vp]),
    'dmf_train_forward': (i32, [vp, vp, vp, i64, vp, vp]),
    'dmf_train_backward': (i32, [vp, vp, vp]),
    'dmf_softmax_ce': (i32, [vp, vp, i32, i64, i32, vp, vp, vp]),
    'dmf_adam_step': (i32, [vp, vp, vp, vp, i64, C.c_float, C.c_float, C.c_float, C.c_float, i64, vp]),
    'dmf_train_step_patches': (i32, [vp, vp, vp, vp, i32, i64, vp, vp]),
    'dmf_train_step_scene': (i32, [vp, vp, vp, i64, i32, vp, vp]),
    'dmf_train_buffer': (i32, [vp, cstr, C.POINTER(vp), C.POINTER(i64)]),
    'dmf_train_set_debug': (i32, [vp, i32]),
    'dmf_train_debug_op': (i32, [vp, i32, i32, i64, vp]),
    'dmf_argmax_confusion': (i32, [vp, vp, i32, i64, i32, vp, vp, vp]),
    'dmf_confusion_at': (i32, [vp, vp, vp, i64, i32, vp, vp]),
    'dmf_scatter_labels': (i32, [vp, vp, vp, i64, vp, i32, vp]),
    'dmf_paint_labels': (i32, [vp, i64, vp, i32, vp, vp]),
}

if not os.path.exists(LIB_PATH):
    raise ImportError('libdmf_b200.so is not built (%s); run `python __graft_entry__.py build` or '
                      '`python dual-modal-fusion_b200/dmf/_build.py` — there is no CPU fallback' % LIB_PATH)

lib = C.CDLL(LIB_PATH)
for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)          # AttributeError here = header and library out of sync
    _f.restype, _f.argtypes = _res, _args


def check(rc):
    if rc != 0:
        raise RuntimeError('libdmf_b200: %s (status %d)' % (lib.dmf_last_error().decode(errors='replace'), rc))
