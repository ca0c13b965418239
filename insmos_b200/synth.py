"""Deterministic synthetic LiDAR sequences shaped like the reference's input.

HDL-64E-like scan: `n_elev` elevations linspace(-24.8deg, +2deg) x `n_azim` azimuths (64 x 1875 =
120 000 rays) cast from a sensor at z=0 over a ground plane (z=-1.73), two street walls, static
and moving boxes; range noise N(0, 0.02 m); max range 80 m; intensity U(0,1); ego speed 1 m/frame.
N scans are expressed in the frame of the last one and stacked oldest -> newest as rows
(x, y, z, intensity, t) with t_i = round((i-N+1)*dt, 3) -- the layout scripts/predict_mos.py:146-158
feeds to InsMOSNet.forward.  No network, no dataset: numpy.random.default_rng(seed) only.
"""
import numpy as np


def _ray_dirs(n_elev, n_azim):
    elev = np.deg2rad(np.linspace(-24.8, 2.0, n_elev))
    azim = np.linspace(-np.pi, np.pi, n_azim, endpoint=False)
    ce, se = np.cos(elev)[:, None], np.sin(elev)[:, None]
    d = np.stack([ce * np.cos(azim)[None, :], ce * np.sin(azim)[None, :], np.broadcast_to(se, (n_elev, n_azim))], axis=-1)
    return d.reshape(-1, 3)


def _hit_boxes(origin, dirs, centers, half, yaw):
    """nearest positive hit distance of rays with rotated boxes (slab test in the box frame), and box id."""
    best = np.full(len(dirs), np.inf)
    bid = np.full(len(dirs), -1, dtype=np.int64)
    for i in range(len(centers)):
        c, s = np.cos(-yaw[i]), np.sin(-yaw[i])
        o = origin - centers[i]
        ox, oy = o[0] * c - o[1] * s, o[0] * s + o[1] * c
        dx, dy = dirs[:, 0] * c - dirs[:, 1] * s, dirs[:, 0] * s + dirs[:, 1] * c
        lo = np.full(len(dirs), -np.inf)
        hi = np.full(len(dirs), np.inf)
        for oo, dd, h in ((ox, dx, half[i, 0]), (oy, dy, half[i, 1]), (o[2], dirs[:, 2], half[i, 2])):
            with np.errstate(divide="ignore", invalid="ignore"):
                t1 = (-h - oo) / dd
                t2 = (h - oo) / dd
            lo = np.maximum(lo, np.minimum(t1, t2))
            hi = np.minimum(hi, np.maximum(t1, t2))
        hit = (hi >= lo) & (lo > 0.5)
        better = hit & (lo < best)
        best[better] = lo[better]
        bid[better] = i
    return best, bid


def make_sequence(seed=0, n_scans=10, n_elev=64, n_azim=1875, dt=0.1, n_static=25, n_moving=3, return_labels=False):
    """-> float32 [n_scans*n_elev*n_azim, 5] (x,y,z,intensity,t); optionally per-point MOS labels of the
    current scan (1 static, 2 moving) and the moving boxes [n_moving,8] of the current frame."""
    rng = np.random.default_rng(seed)
    dirs = _ray_dirs(n_elev, n_azim)
    wall_y = 12.0 + rng.uniform(-2, 2)
    st_c = np.stack([rng.uniform(-55, 55, n_static), rng.choice([-1.0, 1.0], n_static) * rng.uniform(3, 10, n_static),
                     np.full(n_static, -1.73 + 0.75)], axis=1)
    st_yaw = rng.uniform(-0.2, 0.2, n_static)
    mv_c0 = np.stack([rng.uniform(-30, 30, n_moving), rng.uniform(-2.5, 2.5, n_moving), np.full(n_moving, -1.73 + 0.75)], axis=1)
    mv_v = np.stack([rng.choice([-1.0, 1.0], n_moving) * rng.uniform(0.5, 1.2, n_moving), np.zeros(n_moving),
                     np.zeros(n_moving)], axis=1)                       # metres per frame
    half = np.array([2.1, 0.9, 0.75])
    clouds, labels_cur = [], None
    for i in range(n_scans):
        k = i - (n_scans - 1)
        origin = np.array([1.0 * k, 0.0, 0.0])
        # ground
        with np.errstate(divide="ignore", invalid="ignore"):
            tg = np.where(dirs[:, 2] < 0, (-1.73 - origin[2]) / dirs[:, 2], np.inf)
            # walls (height up to z = 4.3)
            tw = np.where(dirs[:, 1] != 0, (np.sign(dirs[:, 1]) * wall_y - origin[1]) / dirs[:, 1], np.inf)
        zw = origin[2] + tw * dirs[:, 2]
        tw = np.where((tw > 0) & (zw < 4.3) & (zw > -1.73), tw, np.inf)
        centers = np.concatenate([st_c, mv_c0 + mv_v * k], axis=0)
        halves = np.tile(half, (len(centers), 1))
        yaws = np.concatenate([st_yaw, np.zeros(n_moving)])
        tb, bid = _hit_boxes(origin, dirs, centers, halves, yaws)
        t = np.minimum(np.minimum(tg, tw), tb)
        is_box = (tb <= t) & np.isfinite(tb)
        t = np.minimum(t, 80.0)
        t = t + rng.normal(0.0, 0.02, len(t))
        pts = origin[None, :] + dirs * t[:, None]
        inten = rng.uniform(0, 1, len(t))
        ts = round(k * dt, 3)
        clouds.append(np.concatenate([pts, inten[:, None], np.full((len(t), 1), ts)], axis=1).astype(np.float32))
        if i == n_scans - 1:
            labels_cur = np.where(is_box & (bid >= n_static), 2, 1).astype(np.int64)
    out = np.concatenate(clouds, axis=0)
    if return_labels:
        boxes = np.concatenate([mv_c0, np.tile(2 * half, (n_moving, 1)), np.zeros((n_moving, 1)), np.ones((n_moving, 1))], axis=1)
        return out, labels_cur, boxes.astype(np.float32)
    return out
