"""ctypes binding of libinsmos_b200.so (C ABI declared in include/insmos_b200.h).

There is no CPU fallback: if the shared library is missing the import of any product op fails
with a RuntimeError that says how to build it.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libinsmos_b200.so")

OK = 0
ERRORS = {-1: "INVALID_ARG", -2: "CUDA", -3: "UNSUPPORTED"}
DEVERR_COORD_RANGE = 1
CNT_ROWS, CNT_AUX, CNT_ERR, CNT_TOTAL, NUM_COUNTERS = 0, 1, 2, 3, 4
ROW_BITS = 25


class MapSpec(C.Structure):
    _fields_ = [("mode", C.c_int32), ("ncol", C.c_int32), ("ndim", C.c_int32), ("first_fastest", C.c_int32),
                ("K", C.c_int32), ("ksize", C.c_int32 * 4), ("a", C.c_int32 * 4), ("b", C.c_int32 * 4),
                ("e", C.c_int32 * 4), ("q", C.c_int32 * 4), ("up_q", C.c_int32 * 4), ("up_ts", C.c_int32 * 4)]


class Epilogue(C.Structure):
    _fields_ = [("scale", C.c_void_p), ("shift", C.c_void_p), ("bias", C.c_void_p), ("residual", C.c_void_p),
                ("relu", C.c_int32), ("first_row", C.c_void_p)]


_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
_F = C.c_float

# name -> (restype, argtypes); kept in one table so tests can check every symbol of the header
PROTOTYPES = {
    "insmos_version": (C.c_char_p, []),
    "insmos_last_error": (C.c_char_p, []),
    "insmos_launch_count": (C.c_uint64, []),
    "insmos_hash_capacity": (_I64, [_I64]),
    "insmos_scan_scratch_bytes": (_I64, [_I64]),
    "insmos_table_clear": (C.c_int, [_P, _I64, _P]),
    "insmos_voxelize4d": (C.c_int, [_P, _I64, _I32, C.POINTER(_F), _P, _I64, _P, _P, _P, _P, _P, _P, _P]),
    "insmos_unique_coords": (C.c_int, [_P, _I64, _I32, C.POINTER(_I32), _P, _I64, _P, _P, _P, _P, _P, _P]),
    "insmos_spconv_out_scratch_bytes": (_I64, [_I64, _I32]),
    "insmos_spconv_out_coords": (C.c_int, [_P, _I64, C.POINTER(_I32), C.POINTER(_I32), C.POINTER(_I32),
                                           C.POINTER(_I32), _P, _I64, _P, _P, _P, _P]),
    "insmos_voxelize3d": (C.c_int, [_P, _I64, _I32, C.POINTER(_F), C.POINTER(_F), C.POINTER(_I32), _I32, _I32,
                                    _P, _I64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "insmos_rulebook_entries_capacity": (_I64, [_I64, _I32, _I32]),
    "insmos_rulebook_build": (C.c_int, [_P, _I64, _P, _I64, C.POINTER(MapSpec), _I32, _P, _P, _P, _P, _P]),
    "insmos_xblock_capacity": (_I64, [_I64]),
    "insmos_xblock_build": (C.c_int, [_P, _I64, _I32, _I32, _P, _I64, _P]),
    "insmos_rulebook_build_xb": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _I32, C.POINTER(MapSpec), _I32, _P, _P, _P, _P]),
    "insmos_leafgrid_capacity": (_I64, [_I64]),
    "insmos_leafgrid_bytes": (_I64, [_I64]),
    "insmos_leafgrid_build": (C.c_int, [_P, _I64, _I32, C.POINTER(_I32), _P, _I64, _P]),
    "insmos_rulebook_build_lg": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, C.POINTER(_I32), C.POINTER(MapSpec), _I32, _P, _P, _P, _P]),
    "insmos_rulebook_build_up": (C.c_int, [_P, _I64, _P, C.POINTER(MapSpec), _I32, _P, _P, _P, _P]),
    "insmos_sparse_conv_fwd": (C.c_int, [_P, _I64, _I32, _P, _I32, _I32, _P, _P, _I32, _P, _I64,
                                         C.POINTER(Epilogue), _I32, _P]),
    "insmos_sparse_conv_fwd_ffma": (C.c_int, [_P, _I64, _I32, _P, _I32, _I32, _P, _P, _I32, _P, _I64,
                                              C.POINTER(Epilogue), _P]),
    "insmos_sparse_conv_fma_supported": (C.c_int, [_I32, _I32, _I32]),
    "insmos_sparse_conv_fwd_fma": (C.c_int, [_P, _I64, _I32, _P, _I32, _I32, _P, _P, _I32, _P, _I64,
                                             C.POINTER(Epilogue), _P]),
    "insmos_conv_wfrag_elems": (_I64, [_I32, _I32, _I32]),
    "insmos_conv_prep_weights": (C.c_int, [_P, _I32, _I32, _I32, _P, _P]),
    "insmos_sparse_conv_fwd_tc": (C.c_int, [_P, _I64, _I32, _P, _I32, _I32, _P, _P, _I32, _P, _I64,
                                            C.POINTER(Epilogue), _P]),
    "insmos_conv_wimg_elems": (_I64, [_I32, _I32, _I32]),
    "insmos_conv_prep_weights_umma": (C.c_int, [_P, _I32, _I32, _I32, _P, _P]),
    "insmos_sparse_conv_fwd_umma": (C.c_int, [_P, _I64, _I32, _P, _I32, _I32, _P, _P, _I32, _P, _I64,
                                              C.POINTER(Epilogue), _P, _I64, _P]),
    "insmos_sparse_conv_umma_workspace_bytes": (_I64, [_I64, _I32]),
    "insmos_conv2d_nhwc_umma": (C.c_int, [_P, _I32, _I32, _I32, _P, _I32, _P, _I32, _P, _P, _I64, _P]),
    "insmos_linear_fwd": (C.c_int, [_P, _I64, _I32, _P, _I32, _P, C.POINTER(Epilogue), _P]),
    "insmos_affine_act": (C.c_int, [_P, _I64, _I32, _P, C.POINTER(Epilogue), _P]),
    "insmos_concat2": (C.c_int, [_P, _I32, _P, _I32, _I64, _P, _P]),
    "insmos_pairsum_add": (C.c_int, [_P, _P, _I64, _I32, _P, _P]),
    "insmos_gather_rows": (C.c_int, [_P, _I32, _P, _I64, _P, _P]),
    "insmos_segment_mean": (C.c_int, [_P, _I32, _P, _I64, _P, _P, _I64, _P]),
    "insmos_build_current_points": (C.c_int, [_P, _I32, _P, _I64, _P, _P, _I32, _I32, _P, _P]),
    "insmos_dense_scatter": (C.c_int, [_P, _P, _I64, _I32, _I32, _I32, _I32, _P, _P]),
    "insmos_center_decode": (C.c_int, [_P, _I64, _I64, _P, _I64, _I64, _I32, _I32, _I32, _F, _F, _F, _F, _F, _P, _P, _P, _P]),
    "insmos_dense_scatter_nhwc": (C.c_int, [_P, _P, _I64, _I32, _I32, _I32, _I32, _P, _P]),
    "insmos_bev_wimg_elems": (_I64, [_I32, _I32, _I32]),
    "insmos_bev_prep_weights_tcgen05": (C.c_int, [_P, _I32, _I32, _I32, _P, _P]),
    "insmos_conv2d_nhwc_tcgen05": (C.c_int, [_P, _I32, _I32, _I32, _P, _I32, _I32, _P, _I32, _P, _P]),
    "insmos_stage_scans": (C.c_int, [_P, _P, _I32, _P, _P, _I32, _P, _I64, _P]),
    "insmos_mos_labels": (C.c_int, [_P, _I64, _I32, C.c_uint32, _P, _P, _P, _P]),
    "insmos_nms_rotated": (C.c_int, [_P, _I32, _F, _I32, _P, _P, _P, _P]),
    "insmos_point_instance_ids": (C.c_int, [_P, _I64, _I32, _P, _I32, _F, _P, _I32, _P, _P]),
    "insmos_instance_stats": (C.c_int, [_P, _I32, _I32, _I64, _P, _I32, _P, _I32, _F, _I32, _P, _P]),
    "insmos_relabel_instances": (C.c_int, [_P, _I32, _I32, _I64, _P, _I32, _P, _P]),
    "insmos_nms_pair_capacity": (_I64, [_I32]),
    "insmos_nms_rotated_pairs": (C.c_int, [_P, _I32, _F, _I32, _P, _P, _I64, _P, _P, _P, _P]),
    "insmos_boxes_to_voxel_units": (C.c_int, [_P, _P, _I32, C.POINTER(_F), C.POINTER(_F), _F, _P, _P]),
    "insmos_box_membership": (C.c_int, [_P, _I64, _P, _I32, _F, _P, _I32, _P, _P]),
    "insmos_boxes_iou3d": (C.c_int, [_P, _I32, _P, _I32, _P, _P]),
    "insmos_time_row_starts": (C.c_int, [_P, _I64, _I32, _I32, _P, _P]),
    "insmos_rulebook_build_lg_from": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, C.POINTER(_I32), C.POINTER(MapSpec), _I32, _P, _P, _P, _P, _P]),
    # training step (N3)
    "insmos_sparse_conv_wgrad_slices": (_I32, [_I64, _I32, _I32, _I32]),
    "insmos_sparse_conv_wgrad": (C.c_int, [_P, _I64, _I32, _P, _I64, _I32, _P, _P, _I32, _I32, _P, _I32, _P, _P]),
    "insmos_column_moments": (C.c_int, [_P, _P, _P, _P, _P, _I64, _I32, _I32, _P, _P, _P]),
    "insmos_bn_train_finalize": (C.c_int, [_P, _P, _I64, _I32, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P]),
    "insmos_bn_bwd_apply": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _P, _P, _P, _P]),
    "insmos_scatter_add_rows": (C.c_int, [_P, _I32, _I32, _P, _I64, _P, _P]),
    "insmos_center_targets": (C.c_int, [_P, _I32, _I32, _I32, _I32, _I32, C.c_double, C.c_double, _I32, _F, _F, _I32, _F, _I32,
                                        _P, _P, _P, _P, _P]),
    "insmos_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _I64, _F, _P]),
}

_lib = None
LAUNCHES = 0          # kernels launched per the KERNELS_PER_CALL table (cross-check of the native counter)


def launch_count():
    """kernels launched by libinsmos_b200.so so far, counted inside the library at every launch site."""
    return int(load().insmos_launch_count())


def load():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "insmos_b200: %s is missing. Build it with `python -m insmos_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    # PyDLL: the GIL is HELD across a C-ABI call.  Every entry point only queues kernels (a few microseconds, never a device
    # synchronisation), and with two forwards in flight on two host threads a CDLL call -- which drops and re-takes the GIL --
    # hands the interpreter to the other thread ~280 times per step; each hand-off costs a condition-variable wake-up.  The
    # threads now switch where they should: at the blocking device->host reads of data-dependent sizes (torch releases the
    # GIL there).  INSMOS_HOLD_GIL=0 restores CDLL for A/B measurements.
    lib = (C.PyDLL if os.environ.get("INSMOS_HOLD_GIL", "1") != "0" else C.CDLL)(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != OK:
        detail = ""
        if rc == -2:
            detail = ": " + load().insmos_last_error().decode()
        raise RuntimeError("insmos_b200: %s failed with %s%s" % (what, ERRORS.get(rc, rc), detail))


# kernels launched per C-ABI call (memsets not counted) -- used for bench.py's gpu_launches claim
KERNELS_PER_CALL = {
    "insmos_table_clear": 1, "insmos_voxelize4d": 5, "insmos_unique_coords": 5, "insmos_spconv_out_coords": 4,
    "insmos_voxelize3d": 8, "insmos_rulebook_build": 1, "insmos_sparse_conv_fwd": 1, "insmos_sparse_conv_fwd_tc": 1, "insmos_sparse_conv_fwd_ffma": 1, "insmos_sparse_conv_fwd_fma": 1, "insmos_conv_prep_weights": 1,
    "insmos_linear_fwd": 1,
    "insmos_affine_act": 1, "insmos_concat2": 1, "insmos_pairsum_add": 1, "insmos_gather_rows": 1,
    "insmos_segment_mean": 2, "insmos_build_current_points": 1, "insmos_dense_scatter": 1, "insmos_center_decode": 1,
    "insmos_nms_rotated": 2, "insmos_point_instance_ids": 3, "insmos_nms_rotated_pairs": 4, "insmos_boxes_to_voxel_units": 1, "insmos_box_membership": 3,
    "insmos_xblock_build": 2, "insmos_leafgrid_build": 1, "insmos_rulebook_build_lg": 1, "insmos_rulebook_build_xb": 1, "insmos_rulebook_build_up": 1,
    "insmos_sparse_conv_fwd_umma": 1, "insmos_conv2d_nhwc_umma": 1, "insmos_conv2d_nhwc_tcgen05": 1,
    "insmos_conv_prep_weights_umma": 1, "insmos_bev_prep_weights_tcgen05": 1, "insmos_dense_scatter_nhwc": 1,
"insmos_boxes_iou3d": 1, "insmos_time_row_starts": 2, "insmos_rulebook_build_lg_from": 1, "insmos_sparse_conv_wgrad": 2, "insmos_column_moments": 1, "insmos_bn_train_finalize": 1, "insmos_bn_bwd_apply": 1, "insmos_scatter_add_rows": 1,
    "insmos_center_targets": 1, "insmos_adam_step": 1,
}
PROFILE = None        # list collecting (name, start_event, end_event, meta) when profiling is on
NEXT_META = None      # ops sets this right before a call to attach algorithmic bytes / flops


def profile_start():
    global PROFILE
    PROFILE = []


def profile_stop():
    """-> list of (name, milliseconds, meta); synchronises."""
    global PROFILE
    import torch
    torch.cuda.synchronize()
    out, prev = [], None
    for n, s, e, m in PROFILE:
        m = dict(m) if m else {}
        if prev is not None:
            m["gap_ms"] = round(prev.elapsed_time(s), 4)       # device idle (or torch kernels) between two C-ABI calls
        out.append((n, s.elapsed_time(e), m))
        prev = e
    PROFILE = None
    return out


_FN = {}             # name -> (bound ctypes function, kernels per call)


def call(name, *args):
    global LAUNCHES, NEXT_META
    ent = _FN.get(name)
    if ent is None:
        ent = _FN[name] = (getattr(load(), name), KERNELS_PER_CALL.get(name, 1))
    LAUNCHES += ent[1]
    if PROFILE is None:
        rc = ent[0](*args)
        if rc != OK:
            check(rc, name)
        return
    import torch
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    meta, NEXT_META = NEXT_META, None
    s.record()
    check(getattr(load(), name)(*args), name)
    e.record()
    PROFILE.append((name, s, e, meta))
