"""Tensor-level wrappers over the C ABI (include/insmos_b200.h).

PyTorch is used for device memory (caching allocator), the current stream and the small host
read-backs of data-dependent sizes -- plumbing.  All arithmetic happens in libinsmos_b200.so.
Every function requires CUDA tensors; there is no CPU path.
"""
import ctypes as C

import os as _os

import torch

from . import _lib
from ._lib import MapSpec, Epilogue, call

I32 = torch.int32
F32 = torch.float32


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    # torch.cuda.current_stream() costs ~16 us of Python per call (cProfile on the GPU host: 2.2 ms per forward over
    # 136 C-ABI calls); the raw getter returns the same cudaStream_t handle in well under a microsecond
    if _raw_stream is not None:
        return _raw_stream(torch._C._cuda_getDevice())           # plain int: ctypes converts it for the c_void_p argument
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()                   # int / None -> c_void_p by the prototypes in _lib.PROTOTYPES


def _req(t, dtype, name):
    if not t.is_cuda:
        raise RuntimeError("insmos_b200.%s: CUDA tensor required (no CPU fallback)" % name)
    # the C ABI launches on the CURRENT device's current stream (_stream()): a tensor of another GPU would be read by the
    # wrong device's kernels, so refuse loudly instead of launching there (callers: `with torch.cuda.device(t.device)`)
    if t.device.index != torch._C._cuda_getDevice():
        raise RuntimeError("insmos_b200.%s: tensor lives on cuda:%d but the current device is cuda:%d; wrap the call in "
                           "`with torch.cuda.device(tensor.device):`" % (name, t.device.index, torch._C._cuda_getDevice()))
    if t.dtype != dtype:
        raise TypeError("insmos_b200.%s: expected %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()


def _arr(ctype, vals):
    return (ctype * len(vals))(*vals)


class _ZeroPool:
    """zero-initialised scratch handed out in slices: one fill kernel per `chunk` elements instead of one torch.zeros
    (a fill launch + ~5 us of host time) per rule book / coordinate op -- 35 of them per forward.  A slice is used once
    and never handed out again; the pool lives on the stream it was created on (the forward runs on one stream)."""

    def __init__(self, dtype, chunk=4096):
        self.dtype, self.chunk, self.state = dtype, chunk, {}

    def take(self, n, device):
        key = (device.type, device.index, torch._C._cuda_getCurrentRawStream(device.index if device.index is not None else torch._C._cuda_getDevice()))
        buf, used = self.state.get(key, (None, 0))
        if buf is None or used + n > self.chunk:
            buf, used = torch.zeros(self.chunk, dtype=self.dtype, device=device), 0
        self.state[key] = (buf, used + n)
        return buf[used:used + n]


_ZERO_I32 = _ZeroPool(torch.int32)
_ZERO_I64 = _ZeroPool(torch.int64)


def _read_counters(counters, what):
    c = counters.cpu().tolist()            # one small D2H read; the only sync of a coordinate op
    if c[_lib.CNT_ERR] & _lib.DEVERR_COORD_RANGE:
        raise RuntimeError("insmos_b200.%s: coordinate outside the packable range "
                           "(|x|,|y|,|z| < 32768 voxels, |t| < 128, batch < 255)" % what)
    return c


def _scan_scratch(n, device):
    return torch.empty(_lib.load().insmos_scan_scratch_bytes(int(n)), dtype=torch.uint8, device=device)


LAZY_SETS = __import__("os").environ.get("INSMOS_LAZY_SETS", "0") != "0"     # deferred count reads: measured no gain on B200 (124.5 vs 123 scans/s), off by default


class CoordSet:
    """Unique integer coordinates [n, ncol] (batch first) plus the hash table that maps a
    coordinate to its row.  Equivalent of an ME coordinate-map key / spconv indices.

    A set made with lazy=True keeps its row count on the device until `n` / `coords` is first read: the kernel that
    produces it is queued, the caller queues independent work behind it, and the one host read of the counter then
    finds the GPU busy instead of draining the queue (each eager read cost a ~65 us bubble in the C2 forward)."""

    def __init__(self, coords, n, table, cap, tensor_stride=None, pending=None):
        self._coords, self._n = coords, (None if n is None else int(n))
        self.table, self.cap = table, int(cap)
        self.tensor_stride = tensor_stride
        self.ncol = coords.shape[1]
        self._pending = pending                 # (counters tensor, name, clone flag) until resolved

    def _resolve(self):
        if self._pending is not None:
            counters, what, clone = self._pending
            self._pending = None
            c = _read_counters(counters, what)
            self._n = int(c[_lib.CNT_ROWS])
            self._coords = self._coords[:self._n].clone() if clone else self._coords[:self._n]

    @property
    def n(self):
        self._resolve()
        return self._n

    @property
    def coords(self):
        self._resolve()
        return self._coords


def _new_table(n, device):
    lib = _lib.load()
    cap = lib.insmos_hash_capacity(int(max(n, 1)))
    table = torch.empty((cap, 2), dtype=torch.int64, device=device)
    call("insmos_table_clear", _p(table), cap, _stream())
    return table, cap


def voxelize4d(points, quant):
    """a1+a2: points [N,>=5] f32 (x,y,z,intensity,t) -> (CoordSet (0,x,y,z,t), inverse [N] i32, cur_index [Nc] i32)."""
    points = _req(points, F32, "voxelize4d")
    n, stride = points.shape
    dev = points.device
    table, cap = _new_table(n, dev)
    slot = torch.empty(max(n, 1), dtype=I32, device=dev)
    coords = torch.empty((max(n, 1), 5), dtype=I32, device=dev)
    inverse = torch.empty(max(n, 1), dtype=I32, device=dev)
    cur = torch.empty(max(n, 1), dtype=I32, device=dev)
    counters = _ZERO_I32.take(_lib.NUM_COUNTERS, dev)
    call("insmos_voxelize4d", _p(points), n, stride, _arr(C.c_float, [float(q) for q in quant]), _p(table), cap,
         _p(slot), _p(coords), _p(inverse), _p(cur), _p(counters), _p(_scan_scratch(n, dev)), _stream())
    c = _read_counters(counters, "voxelize4d")
    nv, nc = c[_lib.CNT_ROWS], c[_lib.CNT_AUX]
    return CoordSet(coords[:nv], nv, table, cap), inverse[:n], cur[:nc]


def unique_coords(coords, q=None, lazy=False):
    """unique int32 rows [N,ncol] in first-occurrence order (optionally floored to multiples of q).
    lazy=True: the row count stays on the device until the returned set's n / coords is read."""
    coords = _req(coords, I32, "unique_coords")
    n, ncol = coords.shape
    dev = coords.device
    table, cap = _new_table(n, dev)
    slot = torch.empty(max(n, 1), dtype=I32, device=dev)
    out = torch.empty((max(n, 1), ncol), dtype=I32, device=dev)
    inverse = torch.empty(max(n, 1), dtype=I32, device=dev)
    counters = _ZERO_I32.take(_lib.NUM_COUNTERS, dev)
    qa = None if q is None else _arr(C.c_int32, [int(v) for v in q])
    call("insmos_unique_coords", _p(coords), n, ncol, qa, _p(table), cap, _p(slot), _p(out), _p(inverse),
         _p(counters), _p(_scan_scratch(n, dev)), _stream())
    if lazy and LAZY_SETS:
        return CoordSet(out, None, table, cap, pending=(counters, "unique_coords", False)), inverse[:n]
    c = _read_counters(counters, "unique_coords")
    nv = c[_lib.CNT_ROWS]
    return CoordSet(out[:nv], nv, table, cap), inverse[:n]


def spconv_out_coords(in_set, ksize, stride, pad, out_shape, lazy=False):
    """output coordinates of a strided spconv SparseConv3d (oracle order); lazy as in unique_coords."""
    lib = _lib.load()
    n = in_set.n
    K = int(ksize[0] * ksize[1] * ksize[2])
    dev = in_set.coords.device
    # an input voxel reaches at most prod(ceil(k/s)) distinct outputs: that bounds the table (the candidate list below still
    # has n*K rows); a table sized for n*K slots was 27x larger than needed and cost 0.1 ms of clears per forward
    reach = 1
    for k_, s_ in zip(ksize, stride):
        reach *= -(-int(k_) // int(s_))
    table, cap = _new_table(n * min(reach, K), dev)
    out = torch.empty((max(n * K, 1), 4), dtype=I32, device=dev)
    counters = _ZERO_I32.take(_lib.NUM_COUNTERS, dev)
    scratch = torch.empty(lib.insmos_spconv_out_scratch_bytes(n, K), dtype=torch.uint8, device=dev)
    call("insmos_spconv_out_coords", _p(in_set.coords), n, _arr(C.c_int32, list(ksize)), _arr(C.c_int32, list(stride)),
         _arr(C.c_int32, list(pad)), _arr(C.c_int32, list(out_shape)), _p(table), cap, _p(out), _p(counters),
         _p(scratch), _stream())
    if lazy and LAZY_SETS:
        return CoordSet(out, None, table, cap, pending=(counters, "spconv_out_coords", True))
    c = _read_counters(counters, "spconv_out_coords")
    nv = c[_lib.CNT_ROWS]
    return CoordSet(out[:nv].clone(), nv, table, cap)


def voxelize3d(points, pc_range, vsize, grid, max_voxels, max_points, want_voxels=False):
    """a6+a7: capped voxelisation with ids and fused mean.  Returns dict with CoordSet 'set' (0,z,y,x),
    'mean' [M,C], 'num_points' [M] i32, 'pc_voxel_id' [N] i32, optional 'voxels' [M,max_points,C]."""
    points = _req(points, F32, "voxelize3d")
    n, Cc = points.shape
    dev = points.device
    table, cap = _new_table(n, dev)
    slot = torch.empty(max(n, 1), dtype=I32, device=dev)
    mv = int(max_voxels)
    rows_cap = max(min(n, mv), 1)
    coords = torch.empty((rows_cap, 4), dtype=I32, device=dev)
    num = torch.empty(rows_cap, dtype=I32, device=dev)
    voxels = torch.empty((rows_cap, max_points, Cc), dtype=F32, device=dev) if want_voxels else None
    mean = torch.empty((rows_cap, Cc), dtype=F32, device=dev)
    ids = torch.empty(max(n, 1), dtype=I32, device=dev)
    work = torch.empty(mv * (1 + max_points), dtype=I32, device=dev)
    counters = _ZERO_I32.take(_lib.NUM_COUNTERS, dev)
    call("insmos_voxelize3d", _p(points), n, Cc, _arr(C.c_float, [float(v) for v in pc_range]),
         _arr(C.c_float, [float(v) for v in vsize]), _arr(C.c_int32, [int(v) for v in grid]), mv, int(max_points),
         _p(table), cap, _p(slot), _p(coords), _p(num), _p(voxels), _p(mean), _p(ids), _p(work), _p(counters),
         _p(_scan_scratch(n, dev)), _stream())
    c = _read_counters(counters, "voxelize3d")
    m = c[_lib.CNT_ROWS]
    out = {"set": CoordSet(coords[:m], m, table, cap), "mean": mean[:m], "num_points": num[:m],
           "pc_voxel_id": ids[:n], "n_distinct": c[_lib.CNT_TOTAL]}
    if want_voxels:
        out["voxels"] = voxels[:m]
    return out


# ---- rule books ---------------------------------------------------------------------------------
def _spec(mode, ncol, ndim, first_fastest, ksize, a=None, b=None, e=None, q=None, up_q=None, up_ts=None):
    s = MapSpec()
    s.mode, s.ncol, s.ndim, s.first_fastest = mode, ncol, ndim, first_fastest
    K = 1
    for d in range(4):
        s.ksize[d] = int(ksize[d]) if d < ndim else 1
        K *= s.ksize[d]
        s.a[d] = int(a[d]) if a is not None and d < ndim else 1
        s.b[d] = int(b[d]) if b is not None and d < ndim else 0
        s.e[d] = int(e[d]) if e is not None and d < ndim else 1
        s.q[d] = int(q[d]) if q is not None and d < ndim else 1
        s.up_q[d] = int(up_q[d]) if up_q is not None and d < ndim else 1
        s.up_ts[d] = int(up_ts[d]) if up_ts is not None and d < ndim else 1
    s.K = K
    return s


def spec_me_cube(ksize, in_stride):
    """MinkowskiConvolution hyper-cube kernel (any stride): in = out + offset*in_stride; odd sizes are
    centred, even sizes start at 0."""
    D = len(ksize)
    b = [-((k - 1) // 2) * t if k % 2 == 1 else 0 for k, t in zip(ksize, in_stride)]
    return _spec(0, D + 1, D, 1, ksize, a=[1] * D, b=b, e=list(in_stride), q=[1] * D)


def spec_me_up(ksize, stride, out_stride):
    """MinkowskiConvolutionTranspose with kernel == stride (2,2,2,1): out = fine rows, in = coarse parent."""
    D = len(ksize)
    up_q = [s * t for s, t in zip(stride, out_stride)]
    return _spec(1, D + 1, D, 1, ksize, up_q=up_q, up_ts=list(out_stride))


def spec_sp_subm(ksize):
    return _spec(0, 4, 3, 0, ksize, a=[1, 1, 1], b=[-(k // 2) for k in ksize], e=[1, 1, 1], q=[1, 1, 1])


def spec_sp_conv(ksize, stride, pad):
    return _spec(0, 4, 3, 0, ksize, a=list(stride), b=[-p for p in pad], e=[1, 1, 1], q=[1, 1, 1])


def spec_sp_inverse(ksize, stride, pad):
    return _spec(0, 4, 3, 0, ksize, a=[1, 1, 1], b=list(pad), e=[-1, -1, -1], q=list(stride))


class Rulebook:
    def __init__(self, seg, entries, TM, K, n_out, n_in, pair_count):
        self.seg, self.entries, self.TM, self.K = seg, entries, TM, K
        self.n_out, self.n_in, self.pair_count = n_out, n_in, pair_count
        self._pairs = None

    @property
    def num_pairs(self):
        if self._pairs is None:
            self._pairs = int(self.pair_count.item())
        return self._pairs

    def triples(self):
        """decode on the device to (k, in_row, out_row) int64 tensors in storage order -- tests / statistics only."""
        K, TM = self.K, self.TM
        n_tiles = (self.n_out + TM - 1) // TM
        dev = self.seg.device
        seg = (self.seg.view(torch.int16).to(torch.int64) & 0xFFFF).view(n_tiles, K + 1)
        pos = torch.arange(TM * K, device=dev, dtype=torch.int64)
        ks, ins, outs = [], [], []
        step = max(1, (1 << 24) // (TM * K))                                # bounded temporaries: ~16 M slots per pass
        ent = self.entries[:n_tiles * TM * K].view(n_tiles, TM * K)
        for t0 in range(0, n_tiles, step):
            sg = seg[t0:t0 + step]
            live = pos[None, :] < sg[:, K:K + 1]
            ti, pi = torch.nonzero(live, as_tuple=True)
            k = torch.searchsorted(sg[:, 1:].contiguous(), pos[None, :].expand(sg.shape[0], -1).contiguous(), right=True)[ti, pi]
            e = ent[t0:t0 + step][ti, pi].to(torch.int64) & 0xFFFFFFFF
            ks.append(k)
            ins.append(e & ((1 << _lib.ROW_BITS) - 1))
            outs.append((e >> _lib.ROW_BITS) + (ti + t0) * TM)
        if not ks:
            z = torch.zeros(0, dtype=torch.int64, device=dev)
            return z, z, z
        return torch.cat(ks), torch.cat(ins), torch.cat(outs)

    def to_coo(self):
        """decode to sorted (k, in_row, out_row) int64 triples on the host -- for tests / statistics only."""
        k, i, o = self.triples()
        trip = torch.stack([k, i, o], dim=1)
        key = (trip[:, 0] * (self.n_in + 1) + trip[:, 1]) * (self.n_out + 1) + trip[:, 2]
        return trip[torch.argsort(key)].cpu()


_TM_OVERRIDE = int(_os.environ.get("INSMOS_TM", "0"))


def choose_tile_rows(n_out, K):
    """rows per rule-book tile.  Larger tiles fill the 16-pair chunks of the mma.sync kernel (a bucket holds ~0.2*TM
    pairs in the 4D maps) but make the map build coarser; measured per map on B200 (C2 workload, rule book + its convs,
    profiles/r01_tile_rows_ab.txt): 5x5x5x1 -> 64; 3^4 -> 128 from 50 k rows, 64 below; 3^3 -> 64 from 20 k rows, 16 below."""
    if _TM_OVERRIDE in (16, 32, 64, 128) and K >= 27:
        tm = _TM_OVERRIDE
    elif K >= 100:
        tm = 64
    elif K >= 64:
        tm = 128 if n_out >= 50_000 else 64
    elif K >= 27:
        tm = 64 if n_out >= 20_000 else 16
    else:
        tm = 128 if n_out >= 300_000 else 64 if n_out >= 150_000 else 32 if n_out >= 75_000 else 16
    while tm > 16 and tm * K >= 65536:
        tm //= 2
    return tm


USE_XBLOCK = _os.environ.get("INSMOS_XBLOCK", "1") != "0"
# measured on B200 (C2 workload): 5-wide x runs 683 -> 404 us, 3-wide runs no gain (295 -> 277, 112 -> 124 us)
XBLOCK_MIN_KX = int(_os.environ.get("INSMOS_XBLOCK_MIN_KX", "5"))


def xblock_table(cs, xstep):
    """x-block table of a CoordSet whose x coordinates are multiples of xstep (cached on the set)."""
    cache = cs.__dict__.setdefault("_xblock", {})
    hit = cache.get(xstep)
    if hit is None:
        cap = _lib.load().insmos_xblock_capacity(max(cs.n, 1))
        table = torch.empty((cap, 4), dtype=torch.int64, device=cs.coords.device)
        call("insmos_xblock_build", _p(cs.coords), cs.n, cs.ncol, int(xstep), _p(table), cap, _stream())
        hit = cache[xstep] = (table, cap)
    return hit


USE_LEAFGRID = _os.environ.get("INSMOS_LEAFGRID", "1") != "0"
# measured on B200 (C2 maps, profiles/r02_leafgrid_ab.txt): the grid pays for maps with >= 27 offsets over sets of >= ~50 k voxels
# (3^4: 0.70 -> 0.33 ms at 496 k rows, 0.21 -> 0.13 at 226 k, 0.13 -> 0.08 at 94 k; 5^3: 0.41 -> 0.31); the 8-offset strided maps
# and the small spconv sets are faster on plain voxel-table probes (one probe per pair, no grid to build).
LEAFGRID_MIN_ROWS = int(_os.environ.get("INSMOS_LEAFGRID_MIN_ROWS", "50000"))
LEAFGRID_MIN_K = int(_os.environ.get("INSMOS_LEAFGRID_MIN_K", "27"))


def leafgrid(cs, step):
    """leaf grid of a CoordSet whose coordinates are multiples of step[d] (cached on the set)."""
    cache = cs.__dict__.setdefault("_leafgrid", {})
    key = tuple(int(v) for v in step)
    hit = cache.get(key)
    if hit is None:
        lib = _lib.load()
        cap = lib.insmos_leafgrid_capacity(max(cs.n, 1))
        grid = torch.empty(lib.insmos_leafgrid_bytes(cap), dtype=torch.uint8, device=cs.coords.device)
        call("insmos_leafgrid_build", _p(cs.coords), cs.n, cs.ncol, _arr(C.c_int32, list(key)), _p(grid), cap, _stream())
        hit = cache[key] = (grid, cap)
    return hit


def leafgrid_eligible(spec, step):
    """affine map whose kernel digits walk the input lattice (e == step, q == 1) and whose neighbourhood spans <= 32 leaves"""
    if spec.mode != 0 or step is None or len(step) != spec.ndim:
        return False
    ncand = 1
    for d in range(spec.ndim):
        if spec.q[d] != 1 or spec.e[d] != int(step[d]) or spec.ksize[d] > 60:
            return False
        ncand *= ((spec.ksize[d] + 2) >> 2) + 1 if d < 3 else spec.ksize[d]
    return ncand <= 32 and spec.K <= 128


def build_rulebook(out_set, in_set, spec, TM=None, parent=None, xstep=None, step=None, first_row=None):
    """tiled rule book of the map in_set -> out_set.  step (optional): lattice step of in_set per dimension (tensor stride
    of a MinkowskiEngine level, 1 for spconv indices): eligible maps then probe the set's leaf grid (leafgrid()).  parent (int32 [n_out], optional): for a transposed map (mode-1
    spec) the row of every output (fine) row's coarse cell in in_set, as returned by unique_coords(q): the map is then
    built without hash probes.  xstep (optional): tensor stride of in_set in x (its x coordinates are multiples of it):
    cube maps with >= 3 offsets in x then probe the set's x-block table."""
    lib = _lib.load()
    K = int(spec.K)
    n_out = out_set.n
    if TM is None:
        TM = choose_tile_rows(n_out, K)
    dev = out_set.coords.device
    n_tiles = max((n_out + TM - 1) // TM, 1)
    seg = torch.empty(n_tiles * (K + 1), dtype=torch.int16, device=dev)
    entries = torch.empty(max(lib.insmos_rulebook_entries_capacity(max(n_out, 1), K, TM), 1), dtype=I32, device=dev)
    pc = _ZERO_I64.take(1, dev)
    if in_set.n > (1 << _lib.ROW_BITS):
        raise RuntimeError("insmos_b200.build_rulebook: more than 2^25 input rows")
    use_xb = (USE_XBLOCK and parent is None and xstep is not None and spec.mode == 0 and spec.first_fastest == 1
              and spec.a[0] == 1 and spec.e[0] == int(xstep) and XBLOCK_MIN_KX <= spec.ksize[0] <= 8
              and all(spec.q[d] == 1 for d in range(spec.ndim)))
    use_lg = (USE_LEAFGRID and parent is None and step is not None and leafgrid_eligible(spec, step)
              and in_set.n >= LEAFGRID_MIN_ROWS and K >= LEAFGRID_MIN_K)
    if use_lg:
        use_xb = False
        lgrid, lgcap = leafgrid(in_set, step)            # (its own C-ABI call: before the profile meta is armed)
    if use_xb:
        xt, xcap = xblock_table(in_set, int(xstep))      # (its own C-ABI call: before the profile meta is armed)
    prof = _lib.PROFILE is not None
    if prof:
        _lib.NEXT_META = {"n_out": n_out, "K": K, "ncol": out_set.ncol}
    if use_lg and first_row is not None:
        # dead-row elimination: only the tiles holding rows >= *first_row (device-side value) are built; the book is marked
        call("insmos_rulebook_build_lg_from", _p(out_set.coords), n_out, _p(in_set.table), in_set.cap, _p(lgrid), lgcap,
             _arr(C.c_int32, [int(v) for v in step]), C.byref(spec), TM, _p(seg), _p(entries), _p(pc), _p(first_row), _stream())
    elif use_lg:
        call("insmos_rulebook_build_lg", _p(out_set.coords), n_out, _p(in_set.table), in_set.cap, _p(lgrid), lgcap,
             _arr(C.c_int32, [int(v) for v in step]), C.byref(spec), TM, _p(seg), _p(entries), _p(pc), _stream())
    elif use_xb:
        call("insmos_rulebook_build_xb", _p(out_set.coords), n_out, _p(in_set.table), in_set.cap, _p(xt), xcap, int(xstep),
             C.byref(spec), TM, _p(seg), _p(entries), _p(pc), _stream())
    elif parent is not None:
        parent = _req(parent, I32, "build_rulebook.parent")
        if parent.shape[0] != n_out:
            raise ValueError("build_rulebook: parent must have one entry per output row")
        call("insmos_rulebook_build_up", _p(out_set.coords), n_out, _p(parent), C.byref(spec), TM,
             _p(seg), _p(entries), _p(pc), _stream())
    else:
        call("insmos_rulebook_build", _p(out_set.coords), n_out, _p(in_set.table), in_set.cap, C.byref(spec), TM,
             _p(seg), _p(entries), _p(pc), None, _stream())
    rb = Rulebook(seg, entries, TM, K, n_out, in_set.n, pc)
    rb.rows_from = first_row if (use_lg and first_row is not None) else None     # tiles below it were NOT built
    if prof:                               # bytes_alg = coordinate rows read + 8 B per pair written
        _lib.PROFILE[-1][3]["bytes"] = 4 * out_set.ncol * n_out + 8 * rb.num_pairs
        _lib.PROFILE[-1][3]["pairs"] = rb.num_pairs
    return rb


# ---- feature ops --------------------------------------------------------------------------------
def _epilogue(scale=None, shift=None, bias=None, residual=None, relu=False, first_row=None):
    """kernel epilogue: v*scale+shift, +bias, +residual, relu.  A conv bias followed by a fused BatchNorm means
    BN(conv + bias) = acc*scale + (shift + bias*scale): the bias is folded into the shift here so that the kernels'
    order (affine first, bias second) cannot apply it after the normalisation."""
    if bias is not None and scale is not None:
        shift = (shift if shift is not None else torch.zeros_like(scale)) + bias.reshape(-1) * scale
        bias = None
    ep = Epilogue()
    keep = []
    for name, t in (("scale", scale), ("shift", shift), ("bias", bias), ("residual", residual)):
        if t is not None:
            t = _req(t, F32, "epilogue." + name)
            keep.append(t)
            setattr(ep, name, t.data_ptr())
        else:
            setattr(ep, name, None)
    ep.relu = 1 if relu else 0
    if first_row is not None:                       # device-side int32: rows below it may be left unwritten (dead-row hint)
        if not (first_row.is_cuda and first_row.dtype == I32 and first_row.numel() >= 1):
            raise TypeError("first_row: a CUDA int32 tensor with one element expected")
        keep.append(first_row)
        ep.first_row = first_row.data_ptr()
    else:
        ep.first_row = None
    return ep, keep


import os as _os

# sparse-conv algorithm used when algo=0: 1 = exact fp32 FFMA (thread per pair), 2 = tensor cores (3xTF32 mma.sync),
# 3 = first-generation SIMT kernel (any shape), 4 = tcgen05/TMEM output-stationary implicit GEMM (wide layers),
# 5 = exact fp32 block-cooperative FFMA kernel (narrow layers: Cin % 8 == 0, Cout in {8,16,32}; conv_fma.cu).
# INSMOS_CONV_ALGO overrides for A/B measurements.
DEFAULT_CONV_ALGO = int(_os.environ.get("INSMOS_CONV_ALGO", "2"))     # measured fastest on B200 in round 1
# wide layers (>= UMMA_MIN_C input AND output channels, K <= UMMA_MAX_K offsets) go to the tcgen05 kernel
# narrow layers on the exact-fp32 block-cooperative FFMA kernel (conv_fma.cu).  OFF by default: measured on B200 (C2 maps,
# profiles/r02_bench_convs_fma_vs_tc4.jsonl) it is 2.5-3.7x slower than the 3xTF32 mma.sync kernel -- broadcast LDS.128
# operands cap the FMA pipe at 47 % (profiles/r02_ffma_rate.txt) and the per-phase barriers cost the light layers more.
USE_FMA = _os.environ.get("INSMOS_FMA", "0") != "0"
USE_UMMA = _os.environ.get("INSMOS_UMMA", "1") != "0"
UMMA_MIN_CIN = int(_os.environ.get("INSMOS_UMMA_MIN_CIN", "32"))
UMMA_MIN_COUT = int(_os.environ.get("INSMOS_UMMA_MIN_COUT", "32"))
UMMA_MAX_K = int(_os.environ.get("INSMOS_UMMA_MAX_K", "32"))
_WIMG_UMMA_CACHE = {}


def fma_eligible(K, Cin, Cout):
    return Cin % 8 == 0 and 8 <= Cin <= 256 and Cout in (8, 16, 32)


def tc_eligible(K, Cin, Cout, TM):
    """shapes of the mma.sync kernel (conv_tc.cu): compile-time Cin in {8,16,24,32,48}, Cout < 64 (any: n-tiles are padded)."""
    return Cin in (8, 16, 24, 32, 48) and Cout < 64 and K <= 127 and TM * K < 65536


def umma_eligible(K, Cin, Cout):
    return Cout % 16 == 0 and 16 <= Cout <= 128 and K <= 128


def prepared_weight_images(weight):
    """pre-swizzled TF32 hi/lo operand images of a [K,Cin,Cout] weight for the tcgen05 sparse conv (cached)."""
    key = (weight.data_ptr(), weight._version, tuple(weight.shape), weight.device.index)
    hit = _WIMG_UMMA_CACHE.get(key)
    if hit is not None:
        return hit[0]
    K, Cin, Cout = weight.shape
    n = _lib.load().insmos_conv_wimg_elems(K, Cin, Cout)
    if n <= 0:
        raise ValueError("sparse_conv: shape K=%d Cin=%d Cout=%d is not supported by the tcgen05 path" % (K, Cin, Cout))
    img = torch.empty(n, dtype=F32, device=weight.device)
    call("insmos_conv_prep_weights_umma", _p(weight), K, Cin, Cout, _p(img), _stream())
    if len(_WIMG_UMMA_CACHE) > 512:
        _WIMG_UMMA_CACHE.clear()
    _WIMG_UMMA_CACHE[key] = (img, weight)
    return img

_WFRAG_CACHE = {}


def prepared_weights(weight):
    """tensor-core fragment-ordered, TF32 hi/lo pre-split copy of a [K,Cin,Cout] weight (insmos_conv_prep_weights),
    cached until the tensor is modified in place or freed (layers are constant at inference)."""
    key = (weight.data_ptr(), weight._version, tuple(weight.shape), weight.device.index)
    hit = _WFRAG_CACHE.get(key)
    if hit is not None:
        return hit[0]
    K, Cin, Cout = weight.shape
    wf = torch.empty(_lib.load().insmos_conv_wfrag_elems(K, Cin, Cout), dtype=I32, device=weight.device)
    call("insmos_conv_prep_weights", _p(weight), K, Cin, Cout, _p(wf), _stream())
    if len(_WFRAG_CACHE) > 512:
        _WFRAG_CACHE.clear()
    _WFRAG_CACHE[key] = (wf, weight)            # keep the source alive so the data_ptr cannot be recycled
    return wf


def sparse_conv(feat, weight, rb, scale=None, shift=None, bias=None, residual=None, relu=False, algo=0, first_row=None):
    """out[n_out,Cout] = sum over rule-book pairs of feat[in] @ weight[k] (+ fused epilogue).
    weight [K,Cin,Cout] f32.  algo: 0/2 = tensor cores (3xTF32, fp32-accurate), 1 = SIMT fp32 reference path."""
    feat = _req(feat, F32, "sparse_conv")
    weight = _req(weight, F32, "sparse_conv")
    K, Cin, Cout = weight.shape
    if K != rb.K or feat.shape[1] != Cin or feat.shape[0] != rb.n_in:
        raise ValueError("sparse_conv: shape mismatch (feat %s, weight %s, rule book K=%d n_in=%d)"
                         % (tuple(feat.shape), tuple(weight.shape), rb.K, rb.n_in))
    out = torch.empty((rb.n_out, Cout), dtype=F32, device=feat.device)
    ep, keep = _epilogue(scale, shift, bias, residual, relu, first_row)
    if algo == 0:
        algo = DEFAULT_CONV_ALGO
        if algo == 2 and Cin < 8 and Cout % 4 == 0 and Cout <= 16:
            algo = 1          # 1..7 input channels: a tensor-core k-step would be mostly padding; fp32 thread-per-pair
        if (algo == 2 and USE_UMMA and Cin >= UMMA_MIN_CIN and Cout >= UMMA_MIN_COUT and K <= UMMA_MAX_K
                and umma_eligible(K, Cin, Cout)):
            algo = 4
        if algo == 2 and USE_FMA and fma_eligible(K, Cin, Cout) and rb.TM * K < 65536:
            algo = 5
        if algo == 2 and not tc_eligible(K, Cin, Cout, rb.TM):
            # off the tuned shapes: tcgen05 kernel when it applies (wide layers with many offsets), else the general SIMT kernel
            algo = 4 if (USE_UMMA and umma_eligible(K, Cin, Cout) and Cin % 8 == 0 and Cin >= 16) else 3
    wf = prepared_weights(weight) if algo == 2 else None
    if _lib.PROFILE is not None:          # algorithmic bytes / flops of this launch (SURVEY 8d formula)
        P, rows_out, rows_in = rb.num_pairs, rb.n_out, rb.n_in
        if first_row is not None and algo == 2:
            # dead-row hint honoured by this kernel: only the tiles from *first_row on are processed -- count THEIR rows and
            # pairs (profiling mode only: this reads the bound back)
            t0 = min(int(first_row[0].item()) // rb.TM, (rb.n_out + rb.TM - 1) // rb.TM)
            n_tiles = (rb.n_out + rb.TM - 1) // rb.TM
            segv = (rb.seg.view(torch.int16)[:n_tiles * (K + 1)].view(n_tiles, K + 1)[t0:, K].to(torch.int64) & 0xFFFF)
            P = int(segv.sum().item())
            rows_out = max(rb.n_out - t0 * rb.TM, 0)
            rows_in = int(rb.n_in * (rows_out / max(rb.n_out, 1)))
        _lib.NEXT_META = {"bytes": 4 * (rows_in * Cin + rows_out * Cout) + 8 * P + 4 * K * Cin * Cout,
                          "flops": 2 * P * Cin * Cout, "pairs": P, "K": K, "Cin": Cin, "Cout": Cout, "n_out": rb.n_out,
                          "rows_done": rows_out}
    if algo == 1:
        call("insmos_sparse_conv_fwd_ffma", _p(feat), rb.n_in, Cin, _p(weight), K, Cout, _p(rb.seg), _p(rb.entries), rb.TM,
             _p(out), rb.n_out, C.byref(ep), _stream())
    elif algo == 4:
        wsb = _lib.load().insmos_sparse_conv_umma_workspace_bytes(rb.n_out, Cout)
        ws = torch.empty(wsb, dtype=torch.uint8, device=feat.device) if wsb > 0 else None
        call("insmos_sparse_conv_fwd_umma", _p(feat), rb.n_in, Cin, _p(prepared_weight_images(weight)), K, Cout, _p(rb.seg),
             _p(rb.entries), rb.TM, _p(out), rb.n_out, C.byref(ep), _p(ws) if ws is not None else None, wsb, _stream())
    elif algo == 5:
        call("insmos_sparse_conv_fwd_fma", _p(feat), rb.n_in, Cin, _p(weight), K, Cout, _p(rb.seg), _p(rb.entries), rb.TM,
             _p(out), rb.n_out, C.byref(ep), _stream())
    elif algo == 3:
        call("insmos_sparse_conv_fwd", _p(feat), rb.n_in, Cin, _p(weight), K, Cout, _p(rb.seg), _p(rb.entries), rb.TM,
             _p(out), rb.n_out, C.byref(ep), 1, _stream())
    else:
        call("insmos_sparse_conv_fwd_tc", _p(feat), rb.n_in, Cin, _p(wf), K, Cout, _p(rb.seg), _p(rb.entries), rb.TM,
             _p(out), rb.n_out, C.byref(ep), _stream())
    return out


def linear(feat, weight, scale=None, shift=None, bias=None, residual=None, relu=False, first_row=None):
    """out = feat[n,Cin] @ weight[Cin,Cout] (+ fused epilogue)."""
    feat = _req(feat, F32, "linear")
    weight = _req(weight, F32, "linear")
    Cin, Cout = weight.shape
    n = feat.shape[0]
    out = torch.empty((n, Cout), dtype=F32, device=feat.device)
    ep, keep = _epilogue(scale, shift, bias, residual, relu, first_row)
    call("insmos_linear_fwd", _p(feat), n, Cin, _p(weight), Cout, _p(out), C.byref(ep), _stream())
    return out


def affine_act(x, scale=None, shift=None, bias=None, residual=None, relu=False, out=None):
    x = _req(x, F32, "affine_act")
    n, Cc = x.shape
    if out is None:
        out = torch.empty_like(x)
    ep, keep = _epilogue(scale, shift, bias, residual, relu)
    call("insmos_affine_act", _p(x), n, Cc, _p(out), C.byref(ep), _stream())
    return out


def concat2(a, b):
    a = _req(a, F32, "concat2")
    b = _req(b, F32, "concat2")
    n = a.shape[0]
    out = torch.empty((n, a.shape[1] + b.shape[1]), dtype=F32, device=a.device)
    call("insmos_concat2", _p(a), a.shape[1], _p(b), b.shape[1], n, _p(out), _stream())
    return out


def pairsum_add(a, b):
    """a[n,C] + b[n,2C].view(n,C,2).sum(2); a may be None."""
    b = _req(b, F32, "pairsum_add")
    n, C2 = b.shape
    out = torch.empty((n, C2 // 2), dtype=F32, device=b.device)
    if a is not None:
        a = _req(a, F32, "pairsum_add")
    call("insmos_pairsum_add", _p(a), _p(b), n, C2 // 2, _p(out), _stream())
    return out


def gather_rows(src, idx):
    src = _req(src, F32, "gather_rows")
    idx = _req(idx, I32, "gather_rows")
    n = idx.shape[0]
    out = torch.empty((n, src.shape[1]), dtype=F32, device=src.device)
    call("insmos_gather_rows", _p(src), src.shape[1], _p(idx), n, _p(out), _stream())
    return out


def segment_mean(feat, inverse, n_rows):
    feat = _req(feat, F32, "segment_mean")
    inverse = _req(inverse, I32, "segment_mean")
    n, Cc = feat.shape
    out = torch.empty((max(n_rows, 1), Cc), dtype=F32, device=feat.device)
    cnt = torch.empty(max(n_rows, 1), dtype=I32, device=feat.device)
    call("insmos_segment_mean", _p(feat), Cc, _p(inverse), n, _p(out), _p(cnt), n_rows, _stream())
    return out[:n_rows]


def build_current_points(points, cur_index, inverse, vox_feat, n_motion):
    points = _req(points, F32, "build_current_points")
    vox_feat = _req(vox_feat, F32, "build_current_points")
    nc = cur_index.shape[0]
    out = torch.empty((nc, 4 + n_motion), dtype=F32, device=points.device)
    call("insmos_build_current_points", _p(points), points.shape[1], _p(cur_index), nc, _p(inverse), _p(vox_feat),
         vox_feat.shape[1], n_motion, _p(out), _stream())
    return out


def dense_scatter(feat, coords, D, H, W):
    feat = _req(feat, F32, "dense_scatter")
    coords = _req(coords, I32, "dense_scatter")
    n, Cc = feat.shape
    out = torch.empty((Cc, D, H, W), dtype=F32, device=feat.device)
    call("insmos_dense_scatter", _p(feat), _p(coords), n, Cc, D, H, W, _p(out), _stream())
    return out


# ---- detection head -----------------------------------------------------------------------------
def center_decode(cls, box, out_size_factor, vx, vy, x_min, y_min, hw=None):
    """head outputs -> boxes [HW,7], scores [HW], labels [HW] i32 (1-based).
    NCHW (hw=None): cls [ncls,H,W], box [8,H,W].  Channels-last (hw=(H,W)): cls [H*W,ncls], box [H*W,8] row-strided
    views (e.g. column slices of one [H*W,11] head output)."""
    if hw is None:
        cls = _req(cls, F32, "center_decode")
        box = _req(box, F32, "center_decode")
        ncls, H, W = cls.shape
        cs, ps, bcs, bps = H * W, 1, H * W, 1
    else:
        H, W = hw
        assert cls.is_cuda and cls.dtype == F32 and box.dtype == F32 and cls.stride(1) == 1 and box.stride(1) == 1
        ncls = cls.shape[1]
        cs, ps, bcs, bps = 1, cls.stride(0), 1, box.stride(0)
    dev = cls.device
    boxes = torch.empty((H * W, 7), dtype=F32, device=dev)
    scores = torch.empty(H * W, dtype=F32, device=dev)
    labels = torch.empty(H * W, dtype=I32, device=dev)
    call("insmos_center_decode", _p(cls), cs, ps, _p(box), bcs, bps, ncls, H, W, float(out_size_factor), float(vx),
         float(vy), float(x_min), float(y_min), _p(boxes), _p(scores), _p(labels), _stream())
    return boxes, scores, labels


# dense BEV conv implementation: "umma" (3x3 convs through the TMEM-operand kernel of spconv_umma.cu, the rest as "tcgen05")
# or "tcgen05" (UTCHMMA + TMEM + TMA, operands in shared memory, bev_tcgen05.cu)
BEV_IMPL = _os.environ.get("INSMOS_BEV_IMPL", "umma")
_WIMG_CACHE = {}


def bev_weight_images(weight):
    """pre-swizzled TF32 hi/lo tile images of a [taps,Cin,Cout] weight for the tcgen05 kernel (cached)."""
    key = (weight.data_ptr(), weight._version, tuple(weight.shape), weight.device.index)
    hit = _WIMG_CACHE.get(key)
    if hit is not None:
        return hit[0]
    taps, Cin, Cout = weight.shape
    img = torch.empty(_lib.load().insmos_bev_wimg_elems(taps, Cin, Cout), dtype=F32, device=weight.device)
    call("insmos_bev_prep_weights_tcgen05", _p(weight), taps, Cin, Cout, _p(img), _stream())
    if len(_WIMG_CACHE) > 64:
        _WIMG_CACHE.clear()
    _WIMG_CACHE[key] = (img, weight)
    return img


def conv2d_nhwc(x, H, W, weight, mode, bias=None, relu=False, impl=None):
    """dense conv on tensor cores (3xTF32): x [H*W,Cin] channels-last, weight [taps,Cin,Cout] (BN folded).
    mode 0: 3x3 pad 1; 1: 1x1; 2: 2x2 stride-2 transposed conv (output [2H*2W,Cout])."""
    x = _req(x, F32, "conv2d_nhwc")
    weight = _req(weight, F32, "conv2d_nhwc")
    taps, Cin, Cout = weight.shape
    assert x.shape == (H * W, Cin) and taps == (9, 1, 4)[mode]
    out = torch.empty(((4 if mode == 2 else 1) * H * W, Cout), dtype=F32, device=x.device)
    if bias is not None:
        bias = _req(bias, F32, "conv2d_nhwc")
    which = impl or BEV_IMPL
    if which == "umma" and mode == 0 and umma_eligible(taps, Cin, Cout) and Cout <= 128:
        call("insmos_conv2d_nhwc_umma", _p(x), H, W, Cin, _p(prepared_weight_images(weight)), Cout, _p(bias), 1 if relu else 0,
             _p(out), None, 0, _stream())
    elif which in ("tcgen05", "umma"):
        img = bev_weight_images(weight)
        call("insmos_conv2d_nhwc_tcgen05", _p(x), H, W, Cin, _p(img), mode, Cout, _p(bias), 1 if relu else 0, _p(out), _stream())
    else:
        raise ValueError("conv2d_nhwc: unknown implementation %r (umma | tcgen05)" % (which,))
    return out


def dense_scatter_nhwc(feat, coords, D, H, W):
    """SparseConvTensor.dense() + HeightCompression, channels-last: [H*W, C*D] with channel c*D+z."""
    feat = _req(feat, F32, "dense_scatter_nhwc")
    coords = _req(coords, I32, "dense_scatter_nhwc")
    n, Cc = feat.shape
    out = torch.empty((H * W, Cc * D), dtype=F32, device=feat.device)
    call("insmos_dense_scatter_nhwc", _p(feat), _p(coords), n, Cc, D, H, W, _p(out), _stream())
    return out


NMS_DENSE = _os.environ.get("INSMOS_NMS_DENSE", "0") != "0"        # A/B switch: the round-1 dense mask kernel


def nms_rotated(boxes_sorted, thresh, max_keep, dense=None):
    """boxes [n,7] sorted by descending score -> keep indices (ascending, <= max_keep) as i32 tensor.
    dense=True: every (i, j) of the upper triangle through the rotated-overlap code (the reference's kernel shape);
    default: pair list (distance test -> compact list -> one pair per thread), same keep list."""
    boxes_sorted = _req(boxes_sorted, F32, "nms_rotated")
    n = boxes_sorted.shape[0]
    dev = boxes_sorted.device
    cb = (n + 63) // 64
    mask = torch.empty(max(n * cb, 1), dtype=torch.int64, device=dev)
    keep = torch.empty(max(min(n, max_keep), 1), dtype=I32, device=dev)
    num = torch.zeros(1, dtype=I32, device=dev)
    if NMS_DENSE if dense is None else dense:
        call("insmos_nms_rotated", _p(boxes_sorted), n, float(thresh), int(max_keep), _p(mask), _p(keep), _p(num), _stream())
    else:
        cap = int(_lib.load().insmos_nms_pair_capacity(n))
        pairs = torch.empty((max(cap, 1), 2), dtype=I32, device=dev)
        count = torch.empty(1, dtype=I32, device=dev)
        call("insmos_nms_rotated_pairs", _p(boxes_sorted), n, float(thresh), int(max_keep), _p(mask), _p(pairs), cap, _p(count),
             _p(keep), _p(num), _stream())
    return keep[:int(num.item())]


def boxes_iou3d(boxes_a, boxes_b):
    """iou3d_nms_utils.boxes_iou3d_gpu: boxes [N,7], [M,7] -> IoU3D [N,M]."""
    boxes_a = _req(boxes_a, F32, "boxes_iou3d")
    boxes_b = _req(boxes_b, F32, "boxes_iou3d")
    if boxes_a.shape[1] != 7 or boxes_b.shape[1] != 7:
        raise ValueError("boxes_iou3d: [N,7] boxes (x,y,z,dx,dy,dz,heading) expected")
    out = torch.zeros((boxes_a.shape[0], boxes_b.shape[0]), dtype=F32, device=boxes_a.device)
    call("insmos_boxes_iou3d", _p(boxes_a), boxes_a.shape[0], _p(boxes_b), boxes_b.shape[0], _p(out), _stream())
    return out


def boxes_to_voxel_units(boxes7, labels, range_min, vsize, stride):
    boxes7 = _req(boxes7, F32, "boxes_to_voxel_units")
    labels = _req(labels, I32, "boxes_to_voxel_units")
    nb = boxes7.shape[0]
    out = torch.empty((nb, 8), dtype=F32, device=boxes7.device)
    call("insmos_boxes_to_voxel_units", _p(boxes7), _p(labels), nb, _arr(C.c_float, [float(v) for v in range_min]),
         _arr(C.c_float, [float(v) for v in vsize]), float(stride), _p(out), _stream())
    return out


def box_membership(coords, boxes8, mult, out=None, out_stride=None, col_offset=0, n_class=3):
    """Array_Index.find_features_by_bbox_with_yaw on device.  coords [n,4] (b,z,y,x) i32; boxes8 [nb,8].
    Writes 1.0 into out[j, col_offset + label-1]; out defaults to a fresh zero [n,n_class] f32."""
    coords = _req(coords, I32, "box_membership")
    n = coords.shape[0]
    dev = coords.device
    if out is None:
        out = torch.zeros((n, n_class), dtype=F32, device=dev)
        out_stride, col_offset = n_class, 0
    nb = boxes8.shape[0]
    if nb > 0 and n > 0:
        boxes8 = _req(boxes8, F32, "box_membership")
        first = torch.empty(nb, dtype=I32, device=dev)
        base = C.c_void_p(out.data_ptr() + 4 * col_offset)
        call("insmos_box_membership", _p(coords), n, _p(boxes8), nb, float(mult), base, int(out_stride), _p(first),
             _stream())
    return out


# ---- steps either side of the forward path -------------------------------------
def stage_scans(raw, offsets, transforms, stamps, out=None):
    """raw [total,4] f32 (x,y,z,intensity), offsets int64 [n+1], transforms f64 [n,4,4] or None, stamps f32 [n]
    (all CUDA) -> [total,5] f32 (x,y,z,intensity,t) in the newest scan's frame (predict_mos.py:114-159)."""
    raw = _req(raw, F32, "stage_scans")
    n_scans = int(stamps.shape[0])
    total = int(raw.shape[0])
    if offsets.dtype != torch.int64 or offsets.numel() != n_scans + 1 or stamps.dtype != F32:
        raise TypeError("stage_scans: offsets must be int64 [n_scans+1], stamps float32 [n_scans]")
    if transforms is not None and (transforms.dtype != torch.float64 or tuple(transforms.shape) != (n_scans, 4, 4)):
        raise TypeError("stage_scans: transforms must be float64 [n_scans,4,4]")
    if out is None:
        out = torch.empty((total, 5), dtype=F32, device=raw.device)
    call("insmos_stage_scans", _p(raw), _p(offsets.contiguous()), n_scans, _p(None if transforms is None else transforms.contiguous()),
         _p(stamps.contiguous()), 0 if transforms is None else 1, _p(out), total, _stream())
    return out


def mos_labels(logits, ignore_mask, label_map=None, want_confidence=True):
    """logits [n,C] f32 -> (labels int32 [n], confidence [n,C-1] f32 or None)  (predict_mos.py:440-454)."""
    logits = _req(logits, F32, "mos_labels")
    n, Cc = logits.shape
    labels = torch.empty(n, dtype=I32, device=logits.device)
    conf = torch.empty((n, Cc - 1), dtype=F32, device=logits.device) if want_confidence else None
    call("insmos_mos_labels", _p(logits), n, Cc, int(ignore_mask), _p(label_map), _p(labels), _p(conf), _stream())
    return labels, conf


USE_TPRUNE = _os.environ.get("INSMOS_TPRUNE", "1") != "0"


def time_row_starts(cs, tcol=None):
    """int32 [33] device tensor; element j (0..15) = smallest row of the CoordSet whose time coordinate is >= -j (n if none).
    Slices [j:j+1] are the `first_row` hints of the convolutions (dead-row elimination)."""
    out = torch.empty(33, dtype=I32, device=cs.coords.device)
    call("insmos_time_row_starts", _p(cs.coords), cs.n, cs.ncol, cs.ncol - 1 if tcol is None else int(tcol), _p(out), _stream())
    return out


# ---- training step: the pieces without a forward counterpart ------------------------------------------
def sparse_conv_wgrad(feat, dout, rb, K, Cin, Cout):
    """dW[K,Cin,Cout] = sum over rule-book pairs (k, i, o) of feat[i]^T dout[o]  (deterministic slice-ordered reduction)."""
    feat = _req(feat, F32, "sparse_conv_wgrad")
    dout = _req(dout, F32, "sparse_conv_wgrad")
    if feat.shape != (rb.n_in, Cin) or dout.shape != (rb.n_out, Cout) or K != rb.K:
        raise ValueError("sparse_conv_wgrad: shape mismatch (feat %s, dout %s, rule book K=%d n_in=%d n_out=%d)"
                         % (tuple(feat.shape), tuple(dout.shape), rb.K, rb.n_in, rb.n_out))
    S = int(_lib.load().insmos_sparse_conv_wgrad_slices(rb.n_out, rb.TM, K, Cin))
    partial = torch.empty((S, K, Cin, Cout), dtype=F32, device=feat.device)
    dw = torch.empty((K, Cin, Cout), dtype=F32, device=feat.device)
    if _lib.PROFILE is not None:
        P = rb.num_pairs
        _lib.NEXT_META = {"bytes": 4 * (rb.n_in * Cin + rb.n_out * Cout) + 8 * P + 4 * K * Cin * Cout,
                          "flops": 2 * P * Cin * Cout, "pairs": P, "K": K, "Cin": Cin, "Cout": Cout, "n_out": rb.n_out}
    call("insmos_sparse_conv_wgrad", _p(feat), rb.n_in, Cin, _p(dout), rb.n_out, Cout, _p(rb.seg), _p(rb.entries), rb.TM, K,
         _p(partial), S, _p(dw), _stream())
    return dw


def scatter_add_rows(src, idx, n_rows):
    """out[idx[i]] += src[i] over a zero [n_rows, C] matrix (idx < 0 skipped): backward of gather_rows."""
    idx = _req(idx, I32, "scatter_add_rows")
    if not src.is_cuda or src.dtype != F32 or src.stride(1) != 1:
        src = _req(src, F32, "scatter_add_rows")
    n, Cc = src.shape
    out = torch.zeros((n_rows, Cc), dtype=F32, device=src.device)
    call("insmos_scatter_add_rows", _p(src), Cc, src.stride(0) if n > 1 else Cc, _p(idx), n, _p(out), _stream())
    return out


def column_moments(a, mode, b=None, gate=None, mean=None, invstd=None):
    """fp64 column sums over [n,C]: mode 0 sum(a); 1 sum((a-mean)^2); 2 (sum(g), sum(g*(b-mean)*invstd)), g = a gated by gate>0;
    3 (sum(a), sum(a^2))."""
    a = _req(a, F32, "column_moments")
    n, Cc = a.shape
    out = torch.empty((2, Cc), dtype=torch.float64, device=a.device)
    call("insmos_column_moments", _p(a), _p(b), _p(gate), _p(mean), _p(invstd), n, Cc, int(mode), _p(out[0]), _p(out[1]) if mode >= 2 else None,
         _stream())
    return out[0] if mode < 2 else (out[0], out[1])


def bn_train_finalize(s, ss, n, gamma, beta, eps, momentum, running_mean=None, running_var=None):
    """-> consts [4, C] f32: rows mean, invstd, scale, shift; updates the running statistics in place."""
    Cc = s.shape[0]
    consts = torch.empty((4, Cc), dtype=F32, device=s.device)
    call("insmos_bn_train_finalize", _p(s), _p(ss), int(n), Cc, _p(gamma), _p(beta), float(eps), float(momentum), _p(consts[0]), _p(consts[1]),
         _p(consts[2]), _p(consts[3]), _p(running_mean), _p(running_var), _stream())
    return consts


def bn_bwd_apply(dy, x, gate, mean, invstd, gamma, s0, s1):
    """-> (dx [n,C], dgamma [C], dbeta [C])"""
    dy = _req(dy, F32, "bn_bwd_apply")
    n, Cc = dy.shape
    dx = torch.empty_like(dy)
    dgb = torch.empty((2, Cc), dtype=F32, device=dy.device)
    call("insmos_bn_bwd_apply", _p(dy), _p(x), _p(gate), _p(mean), _p(invstd), _p(gamma), _p(s0), _p(s1), n, Cc, _p(dx), _p(dgb[0]), _p(dgb[1]),
         _stream())
    return dx, dgb[0], dgb[1]


def center_targets(gt_boxes, max_objs, H, W, ncls, x_min, y_min, vx, vy, out_size_factor, min_overlap, min_radius, range_is_fp64=False):
    """CenterHead.get_targets_single on device: gt_boxes [M,8] -> heatmap [ncls,H,W], anno_boxes [max_objs,8],
    inds int64 [max_objs], masks uint8 [max_objs]."""
    gt_boxes = _req(gt_boxes, F32, "center_targets")
    dev = gt_boxes.device
    heat = torch.empty((ncls, H, W), dtype=F32, device=dev)
    anno = torch.empty((max_objs, 8), dtype=F32, device=dev)
    inds = torch.empty(max_objs, dtype=torch.int64, device=dev)
    masks = torch.empty(max_objs, dtype=torch.uint8, device=dev)
    call("insmos_center_targets", _p(gt_boxes), gt_boxes.shape[0], int(max_objs), int(H), int(W), int(ncls), float(x_min),
         float(y_min), 1 if range_is_fp64 else 0, float(vx), float(vy), int(out_size_factor), float(min_overlap), int(min_radius), _p(heat), _p(anno),
         _p(inds), _p(masks), _stream())
    return heat, anno, inds, masks


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    """in-place torch.optim.Adam step over flat fp32 buffers."""
    for t in (param, grad, exp_avg, exp_avg_sq):
        if not (t.is_cuda and t.dtype == F32 and t.is_contiguous() and t.numel() == param.numel()):
            raise ValueError("adam_step: flat contiguous float32 CUDA buffers of equal length required")
    call("insmos_adam_step", _p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), param.numel(), float(lr), float(beta1), float(beta2),
         float(eps), float(weight_decay), int(step), float(grad_scale), _stream())
