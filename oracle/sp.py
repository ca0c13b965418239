"""spconv 2.3.6 semantics, restated (test infrastructure; see oracle/__init__.py).

Conventions frozen here (call sites voxel_generate.py:17-31,
spconv_unet.py:120-208,284-410, height_compression.py:26):
  * PointToVoxel: voxel index per axis floor((p - range_min) / vsize) in fp32; a point is dropped
    (id -1) when any index is outside [0, grid); voxels are numbered by first occurrence (CPU
    semantics), at most max_voxels; each voxel stores its first max_points points; coordinates are
    returned (z, y, x).
  * weights [Cout, kz, ky, kx, Cin]; offset index (kz*KY + ky)*KX + kx; cross-correlation
    out[o] += W[:, k, :] . in[o*s - p + k].
  * SubMConv3d: output set = input set, kernel centred at k//2 (padding ignored).
  * SparseConv3d: output set = all o in [0, out_shape) reached from some input, numbered in creation
    order scanning inputs ascending and offsets ascending; out_shape = (in + 2p - k)//s + 1.
  * SparseInverseConv3d(indice_key): the pairs of that key swapped, same offset index; output rows,
    indices and spatial shape are the key's input.
"""
import numpy as np
import torch

from . import me


def point_to_voxel(points, vsize, pc_range, max_points, max_voxels):
    """generate_voxel_with_id.  Returns voxels [M,max_points,C], coords [M,3] (z,y,x) int32,
    num_points [M] int32, pc_voxel_id [N] int64."""
    pts = torch.as_tensor(points, dtype=torch.float32)
    n, C = pts.shape
    lo = torch.tensor(pc_range[:3], dtype=torch.float32)
    vs = torch.tensor(vsize, dtype=torch.float32)
    grid = np.round((np.asarray(pc_range[3:], dtype=np.float64) - np.asarray(pc_range[:3], dtype=np.float64)) /
                    np.asarray(vsize, dtype=np.float64)).astype(np.int64)          # (gx, gy, gz)
    c = torch.floor((pts[:, :3] - lo) / vs).to(torch.int64).numpy()               # (ix, iy, iz)
    valid = np.all((c >= 0) & (c < grid[None, :]), axis=1)
    ids = np.full(n, -1, dtype=np.int64)
    vidx = np.nonzero(valid)[0]
    zyx = np.stack([np.zeros(len(vidx), dtype=np.int64), c[vidx, 2], c[vidx, 1], c[vidx, 0]], axis=1)
    uniq, inv = me.unique_first(zyx)
    keep = inv < max_voxels
    ids[vidx[keep]] = inv[keep]
    M = min(len(uniq), max_voxels)
    coords = uniq[:M, 1:4].astype(np.int32)
    # first max_points points of every voxel, in point order
    pv = vidx[keep]
    vid = inv[keep]
    order = np.argsort(vid, kind="stable")
    pv_s, vid_s = pv[order], vid[order]
    counts = np.bincount(vid_s, minlength=M)
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
    pos = np.arange(len(vid_s)) - starts[vid_s]
    sel = pos < max_points
    voxels = torch.zeros((M, max_points, C), dtype=torch.float32)
    voxels[torch.from_numpy(vid_s[sel]), torch.from_numpy(pos[sel])] = pts[torch.from_numpy(pv_s[sel])]
    num = np.minimum(counts, max_points).astype(np.int32)
    return voxels, coords, num, ids


def mean_vfe(voxels, num_points):
    """mean_vfe.py:47-52"""
    v = torch.as_tensor(voxels, dtype=torch.float32)
    s = v.sum(dim=1, keepdim=False)
    norm = torch.clamp_min(torch.as_tensor(num_points).view(-1, 1).to(torch.float32), min=1.0)
    return (s / norm).contiguous()


def _offsets3(ksize):
    """[K,3] (dz,dy,dx) kernel positions, x fastest."""
    kz, ky, kx = ksize
    o = np.zeros((kz * ky * kx, 3), dtype=np.int64)
    for k in range(len(o)):
        o[k] = (k // (ky * kx), (k // kx) % ky, k % kx)
    return o


def subm_maps(indices, ksize, native=None):
    """SubMConv3d pairs: in = out + (k - ksize//2).  indices [N,4] (b,z,y,x).  Large maps look their neighbours up in
    oracle/native (hash map, OpenMP over the offsets); native=False is the numpy statement both must equal."""
    ind = np.asarray(indices).astype(np.int64)
    offs = _offsets3(ksize) - np.asarray([k // 2 for k in ksize])[None, :]
    if native is None:
        native = len(ind) * len(offs) >= me.NATIVE_MIN_PROBES
    if native:
        from . import native as _native
        return me.maps_from_neighbor_table(_native.neighbor_table(ind, ind, offs))
    lut = me.Lookup(ind)
    maps = []
    for k in range(len(offs)):
        q = ind.copy()
        q[:, 1:4] += offs[k]
        ok = np.all(np.abs(q[:, 1:4]) < (1 << 15), axis=1)
        idx = np.full(len(q), -1, dtype=np.int64)
        if ok.any():
            idx[ok] = lut.find(q[ok])
        o = np.nonzero(idx >= 0)[0]
        maps.append((idx[o], o))
    return maps


def conv_out_shape(in_shape, ksize, stride, pad):
    return [(i + 2 * p - k) // s + 1 for i, k, s, p in zip(in_shape, ksize, stride, pad)]


def sparse_conv_indices(indices, in_shape, ksize, stride, pad):
    """SparseConv3d output indices (creation order) and pairs.  Returns (out_indices [M,4], maps, out_shape)."""
    ind = np.asarray(indices).astype(np.int64)
    out_shape = conv_out_shape(in_shape, ksize, stride, pad)
    offs = _offsets3(ksize)
    K = len(offs)
    n = len(ind)
    s = np.asarray(stride)[None, None, :]
    num = ind[:, None, 1:4] + np.asarray(pad)[None, None, :] - offs[None, :, :]          # [n,K,3]
    ok = np.all(num % s == 0, axis=2)
    o = num // s
    ok &= np.all((o >= 0) & (o < np.asarray(out_shape)[None, None, :]), axis=2)
    cand = np.concatenate([np.broadcast_to(ind[:, None, 0:1], (n, K, 1)), o], axis=2).reshape(n * K, 4)
    okf = ok.reshape(-1)
    vi = np.nonzero(okf)[0]                                                              # (i,k)-major order
    uniq, inv = me.unique_first(cand[vi])
    maps = []
    kk = vi % K
    ii = vi // K
    for k in range(K):
        m = kk == k
        maps.append((ii[m], inv[m]))
    return uniq.astype(np.int32), maps, out_shape


def conv(feats, weight, maps, n_out):
    """weight in spconv layout [Cout,kz,ky,kx,Cin]."""
    W = torch.as_tensor(weight, dtype=me.FDTYPE)
    Cout, Cin = W.shape[0], W.shape[-1]
    Wk = W.reshape(Cout, -1, Cin).permute(1, 2, 0).contiguous()                          # [K,Cin,Cout]
    return me.conv(feats, Wk, maps, n_out)


def dense(features, indices, spatial_shape, batch_size=1):
    f = torch.as_tensor(features, dtype=me.FDTYPE)
    ind = torch.as_tensor(np.asarray(indices), dtype=torch.int64)
    out = torch.zeros((batch_size, f.shape[1], *spatial_shape), dtype=me.FDTYPE)
    out[ind[:, 0], :, ind[:, 1], ind[:, 2], ind[:, 3]] = f
    return out


def gather_features_by_pc_voxel_id(seg, ids):
    seg = torch.as_tensor(seg, dtype=me.FDTYPE)
    ids = torch.as_tensor(np.asarray(ids), dtype=torch.int64)
    out = torch.zeros((len(ids), seg.shape[1]), dtype=me.FDTYPE)
    m = ids >= 0
    out[m] = seg[ids[m]]
    return out
