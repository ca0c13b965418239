"""ctypes access to oracle/native/liboracle_native.so and import helpers for oracle/_ref (test infrastructure)."""
import ctypes as C
import importlib.util
import os

import numpy as np

from . import build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build.NATIVE_LIB
        if not os.path.exists(path):
            path = build.build_native()
        L = C.CDLL(path)
        L.oracle_iou_bev.restype = C.c_float
        L.oracle_iou_bev.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_iou_matrix.restype = None
        L.oracle_iou_matrix.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_nms.restype = C.c_int
        L.oracle_nms.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p]
        L.oracle_find_features_by_bbox_with_yaw.restype = None
        L.oracle_find_features_by_bbox_with_yaw.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.oracle_find_point_in_instance_bbox_with_yaw.restype = None
        L.oracle_find_point_in_instance_bbox_with_yaw.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                                                  C.c_int, C.c_float]
        L.oracle_neighbor_table.restype = C.c_int
        L.oracle_neighbor_table.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                            C.c_void_p]
        _lib = L
    return _lib


def set_threads(n):
    """threads of the OpenMP loops in liboracle_native.so (kernel-map lookups, NMS mask rows); returns the count in effect."""
    L = lib()
    L.oracle_set_threads.restype = None
    L.oracle_set_threads.argtypes = [C.c_int]
    L.oracle_max_threads.restype = C.c_int
    L.oracle_max_threads.argtypes = []
    L.oracle_set_threads(int(n))
    return int(L.oracle_max_threads())


def neighbor_table(in_coords, q_coords, offs):
    """rows of `q + offs[k]` in in_coords for every kernel offset k: int32 [K, n_q], -1 = absent (hash map + OpenMP over the
    offsets; the numpy statement it must equal is oracle/me.py::kernel_map / oracle/sp.py::subm_maps with native=False)."""
    a = np.ascontiguousarray(in_coords, dtype=np.int64)
    q = np.ascontiguousarray(q_coords, dtype=np.int64)
    o = np.ascontiguousarray(offs, dtype=np.int64)
    if a.ndim != 2 or q.ndim != 2 or a.shape[1] != q.shape[1] or o.ndim != 2:
        raise ValueError("neighbor_table: [n,ncol] coordinate rows and [K,D] offsets expected")
    nbr = np.empty((o.shape[0], q.shape[0]), dtype=np.int32)
    rc = lib().oracle_neighbor_table(a.ctypes.data, a.shape[0], q.ctypes.data, q.shape[0], a.shape[1], o.ctypes.data, o.shape[0],
                                     o.shape[1], nbr.ctypes.data)
    if rc != 0:
        raise RuntimeError("oracle_neighbor_table failed (%d)" % rc)
    return nbr


def overlap_matrix(a, b):
    """rotated BEV overlap AREA of every pair (iou3d_nms_kernel.cu:104-225 box_overlap, :236-249 boxes_overlap_kernel)."""
    a = np.ascontiguousarray(a, dtype=np.float32); b = np.ascontiguousarray(b, dtype=np.float32)
    L = lib()
    L.oracle_box_overlap.restype = C.c_float
    L.oracle_box_overlap.argtypes = [C.c_void_p, C.c_void_p]
    out = np.zeros((len(a), len(b)), dtype=np.float32)
    for i in range(len(a)):
        for j in range(len(b)):
            out[i, j] = L.oracle_box_overlap(a[i].ctypes.data, b[j].ctypes.data)
    return out


def iou_matrix(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32); b = np.ascontiguousarray(b, dtype=np.float32)
    out = np.zeros((len(a), len(b)), dtype=np.float32)
    lib().oracle_iou_matrix(a.ctypes.data, len(a), b.ctypes.data, len(b), out.ctypes.data)
    return out


def nms(boxes_sorted, thresh):
    """boxes [n,7] sorted by descending score -> all kept indices (int64, ascending)."""
    b = np.ascontiguousarray(boxes_sorted, dtype=np.float32)
    keep = np.zeros(max(len(b), 1), dtype=np.int64)
    n = lib().oracle_nms(b.ctypes.data, len(b), C.c_float(thresh), keep.ctypes.data)
    return keep[:n]


def find_features_by_bbox_with_yaw(vox_xyz, boxes8, n_class=3):
    v = np.ascontiguousarray(vox_xyz, dtype=np.int32); b = np.ascontiguousarray(boxes8, dtype=np.float32)
    out = np.zeros((len(v), n_class), dtype=np.int32)
    lib().oracle_find_features_by_bbox_with_yaw(v.ctypes.data, len(v), b.ctypes.data, len(b), out.ctypes.data, n_class)
    return out


def find_point_in_instance_bbox_with_yaw(points, boxes8, out_ground, n_class=3):
    """refine.py:196: per-point instance ids [n,n_class] int32 (box index + 1 in column label-1), serial box order."""
    pts = np.ascontiguousarray(points, dtype=np.float32); b = np.ascontiguousarray(boxes8, dtype=np.float32)
    out = np.zeros((len(pts), n_class), dtype=np.int32)
    lib().oracle_find_point_in_instance_bbox_with_yaw(pts.ctypes.data, len(pts), pts.shape[1], b.ctypes.data, len(b),
                                                      out.ctypes.data, n_class, C.c_float(out_ground))
    return out


def _load_ext(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _KeepThreadCount:
    """proxy of the reference's Array_Index module.  Its functions call omp_set_num_threads(30) (Array_Index.cpp:19,88) and leave
    the process-wide OpenMP thread count there, which oversubscribes every later torch CPU operation of the process (measured
    here: the training-oracle fixture 3 s -> 36 s on 8 cores).  The proxy restores the count after each call."""

    def __init__(self, mod):
        self._mod = mod

    def __getattr__(self, name):
        fn = getattr(self._mod, name)
        if not callable(fn):
            return fn

        def call(*args, **kwargs):
            import torch
            n = torch.get_num_threads()
            try:
                return fn(*args, **kwargs)
            finally:
                torch.set_num_threads(n)
                set_threads(n)
        return call


def ref_array_index():
    """the reference's compiled Array_Index module (oracle/_ref), or None."""
    p = build.ref_paths()[0]
    return _KeepThreadCount(_load_ext("Array_Index", p)) if os.path.exists(p) else None


def ref_iou3d():
    """the reference's compiled iou3d_nms_cuda module (oracle/_ref), or None.  Needs `import torch` first."""
    p = build.ref_paths()[1]
    if not os.path.exists(p):
        return None
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    return _load_ext("iou3d_nms_cuda", p)
