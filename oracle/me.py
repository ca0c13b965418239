"""MinkowskiEngine CPU semantics, restated (test infrastructure; see oracle/__init__.py).

Conventions frozen here (call sites motionnet.py:21-50,
minkunet.py:52-181, resnet.py:87-126):
  * coordinates are int rows (batch, x, y, z[, t]); unique voxels are numbered by FIRST OCCURRENCE
    in input order (ME's CPU coordinate map).
  * kernel offsets of a hyper-cube kernel: odd size k -> -(k-1)/2..(k-1)/2, even size -> 0..k-1, each
    times the input tensor stride; the offset index runs with dimension 0 (x) fastest.
  * strided convolution: output coordinates floor(c / (ts*s)) * (ts*s), first occurrence over inputs.
  * transposed convolution (kernel == stride): output key is the existing finer coordinate map; the
    kernel map is the strided map with in/out swapped and the same offset index.
  * convolution = for k: gather rows -> sgemm with W[k] -> scatter-add (k-major accumulation).
"""
import numpy as np
import torch

_B = 1 << 15


# dtype of the FEATURE arithmetic (coordinates are always quantised in fp32).  oracle/train.py switches it to float64 to obtain
# reference gradients free of fp32 rounding (tests/golden/make_golden_train_f64.py)
FDTYPE = torch.float32


def pack_keys(coords):
    """int64 key of int rows (batch, c0, c1, c2[, c3]); |c0..c2| < 2^15, |c3| < 128, batch < 127."""
    c = np.asarray(coords).astype(np.int64)
    key = c[:, 0]
    for d in range(1, 4):
        key = key * (2 * _B) + (c[:, d] + _B)
    c3 = c[:, 4] if c.shape[1] > 4 else np.zeros(len(c), dtype=np.int64)
    return key * 256 + (c3 + 128)


def unique_first(coords):
    """unique rows numbered by first occurrence.  Returns (unique_rows, inverse[int64])."""
    coords = np.asarray(coords)
    if len(coords) == 0:
        return coords.copy(), np.zeros(0, dtype=np.int64)
    keys = pack_keys(coords)
    _, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # sorted-unique id -> rank by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return coords[first[order]], rank[inv.reshape(-1)]


def sparse_collate(coords_list, feats_list):
    """ME.utils.sparse_collate: floor float coordinates, prepend the batch index (motionnet.py:33)."""
    cs, fs = [], []
    for b, (c, f) in enumerate(zip(coords_list, feats_list)):
        c = torch.as_tensor(c)
        ci = torch.floor(c).to(torch.int32) if c.is_floating_point() else c.to(torch.int32)
        cs.append(torch.cat([torch.full((len(ci), 1), b, dtype=torch.int32), ci], dim=1))
        fs.append(torch.as_tensor(f))
    return torch.cat(cs, 0), torch.cat(fs, 0)


def quantize_points(points_xyzt, quant):
    """motionnet.py:22-34: fp32 true division, floor, int32, batch column 0.  Returns (coords[N,5], t_is_zero[N])."""
    p = torch.as_tensor(points_xyzt, dtype=torch.float32)
    q = torch.as_tensor(quant, dtype=torch.float32)
    d = torch.div(p, q)
    coords = torch.cat([torch.zeros((len(d), 1), dtype=torch.int32), torch.floor(d).to(torch.int32)], dim=1)
    return coords.numpy(), (d[:, 3] == 0).numpy()


def segment_mean(feats, inverse, n_rows):
    """UNWEIGHTED_AVERAGE quantisation of TensorField.sparse()."""
    feats = torch.as_tensor(feats, dtype=FDTYPE)
    inv = torch.as_tensor(inverse, dtype=torch.int64)
    out = torch.zeros((n_rows, feats.shape[1]), dtype=FDTYPE)
    out.index_add_(0, inv, feats)
    cnt = torch.bincount(inv, minlength=n_rows).clamp_min(1).to(FDTYPE)
    return out / cnt[:, None]


def stride_coords(coords, new_stride):
    """coordinate-manager stride: floor to multiples of new_stride per dim, unique by first occurrence."""
    c = np.asarray(coords).copy()
    for d, q in enumerate(new_stride):
        if q > 1:
            c[:, 1 + d] = np.floor_divide(c[:, 1 + d], q) * q
    return unique_first(c)


def kernel_offsets(ksize, in_stride):
    """list of per-dim offsets, index k with dimension 0 fastest."""
    D = len(ksize)
    per_dim = []
    for k, t in zip(ksize, in_stride):
        base = -((k - 1) // 2) if k % 2 == 1 else 0
        per_dim.append([(base + i) * t for i in range(k)])
    K = int(np.prod(ksize))
    offs = np.zeros((K, D), dtype=np.int64)
    for k in range(K):
        rem = k
        for d in range(D):
            offs[k, d] = per_dim[d][rem % ksize[d]]
            rem //= ksize[d]
    return offs


class Lookup:
    """coordinate -> row lookup by sorted packed keys."""

    def __init__(self, coords):
        self.keys = pack_keys(coords)
        self.order = np.argsort(self.keys, kind="stable")
        self.sorted = self.keys[self.order]

    def find(self, coords):
        k = pack_keys(coords)
        if len(self.sorted) == 0:                      # empty input set: nothing can be found
            return np.full(len(k), -1, dtype=np.int64)
        pos = np.minimum(np.searchsorted(self.sorted, k), len(self.sorted) - 1)
        return np.where(self.sorted[pos] == k, self.order[pos], -1)


# Above this many (row, offset) probes the neighbour lookups of a kernel map run in oracle/native (hash map + OpenMP over the
# offsets -- how ME's and spconv's CPU backends do it); below it, and always with native=False, the plain numpy statement
# (stable sort + binary search).  Both give the same maps in the same order: tests/test_oracle_ops.py.
NATIVE_MIN_PROBES = 1 << 18


def maps_from_neighbor_table(nbr):
    """[K, n_out] table of input rows (-1 = absent) -> [(in_idx, out_idx)] per offset, output rows ascending."""
    maps = []
    for k in range(nbr.shape[0]):
        o = np.nonzero(nbr[k] >= 0)[0]
        maps.append((nbr[k][o].astype(np.int64), o))
    return maps


def kernel_map(in_coords, out_coords, ksize, in_stride, native=None):
    """[(in_idx, out_idx)] per offset k:  in = out + offset_k  (covers stride-1 and strided convs)."""
    offs = kernel_offsets(ksize, in_stride)
    if native is None:
        native = len(out_coords) * len(offs) >= NATIVE_MIN_PROBES and np.asarray(in_coords).shape[1] in (4, 5)
    if native:
        from . import native as _native
        return maps_from_neighbor_table(_native.neighbor_table(in_coords, out_coords, offs))
    lut = Lookup(in_coords)
    oc = np.asarray(out_coords).astype(np.int64)
    maps = []
    for k in range(len(offs)):
        q = oc.copy()
        q[:, 1:1 + offs.shape[1]] += offs[k]
        ok = np.all(np.abs(q[:, 1:4]) < _B, axis=1)
        if q.shape[1] > 4:
            ok &= np.abs(q[:, 4]) < 128
        idx = np.full(len(q), -1, dtype=np.int64)
        if ok.any():
            idx[ok] = lut.find(q[ok])
        o = np.nonzero(idx >= 0)[0]
        maps.append((idx[o], o))
    return maps


def transpose_map(maps):
    return [(o, i) for (i, o) in maps]


def conv(feats, weight, maps, n_out, bias=None):
    """gather -> mm -> scatter-add per kernel offset (ME CPU algorithm).  weight [K,Cin,Cout] or [Cin,Cout]."""
    feats = torch.as_tensor(feats, dtype=FDTYPE)
    W = torch.as_tensor(weight, dtype=FDTYPE)
    if W.dim() == 2:
        out = feats @ W
    else:
        out = torch.zeros((n_out, W.shape[2]), dtype=FDTYPE)
        for k, (i, o) in enumerate(maps):
            if len(i) == 0:
                continue
            out.index_add_(0, torch.from_numpy(np.asarray(o, dtype=np.int64)),
                           feats.index_select(0, torch.from_numpy(np.asarray(i, dtype=np.int64))) @ W[k])
    if bias is not None:
        out = out + torch.as_tensor(bias, dtype=FDTYPE).reshape(1, -1)
    return out


def maps_to_triples(maps, n_in, n_out):
    """canonical sorted (k, in, out) int64 triples of a kernel map."""
    ks, ins, outs = [], [], []
    for k, (i, o) in enumerate(maps):
        ks.append(np.full(len(i), k, dtype=np.int64)); ins.append(np.asarray(i, dtype=np.int64)); outs.append(np.asarray(o, dtype=np.int64))
    if not ks:
        return np.zeros((0, 3), dtype=np.int64)
    t = np.stack([np.concatenate(ks), np.concatenate(ins), np.concatenate(outs)], axis=1)
    key = (t[:, 0] * (n_in + 1) + t[:, 1]) * (n_out + 1) + t[:, 2]
    return t[np.argsort(key, kind="stable")]


def batch_norm_eval(x, weight, bias, mean, var, eps):
    x = torch.as_tensor(x, dtype=torch.float32)
    return torch.nn.functional.batch_norm(x, torch.as_tensor(mean), torch.as_tensor(var), torch.as_tensor(weight),
                                          torch.as_tensor(bias), False, 0.0, eps)
