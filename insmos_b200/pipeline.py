"""Scan pipeline around InsMOSNet.forward: the steps immediately before and after the hot path.

Mirrors what scripts/predict_mos.py does per sample on the host -- DemoDataset.__getitem__ (:114-159: pose transform in
float64, timestamps, concatenation), `.cuda()` (:418), and the label step (:436-456: mask, softmax, argmax, label map,
`.cpu().numpy()`) -- as: ONE pinned host->device copy of the raw scans on a copy stream, one staging kernel, the forward,
one labelling kernel and two small asynchronous device->host copies into pinned buffers.  Two slots are in flight, each
driven by its own host thread and CUDA stream (insmos_b200.engine): the copies and the kernels of scan k+1 fill the
bubbles that the data-dependent host reads of scan k leave on the GPU; nothing here is a CPU fallback: the model and the
staging kernels are CUDA only.
"""
import numpy as np
import torch

from insmos_b200 import ops

DEFAULT_IGNORE = {0: True, 1: False, 2: False}            # config/semantic-kitti-mos.yaml:157-160
DEFAULT_MAP_INV = {0: 0, 1: 9, 2: 251}                    # config/semantic-kitti-mos.yaml:152-155


def scan_transforms(poses):
    """inv(to_pose) @ from_pose per scan in float64, newest scan = target (predict_mos.py:132-136,162)."""
    poses = np.asarray(poses, dtype=np.float64)
    to_inv = np.linalg.inv(poses[-1])
    return np.stack([to_inv @ p for p in poses])


class _Slot:
    def __init__(self, device, max_points, n_scans, n_class):
        self.raw_host = torch.empty((max_points, 4), dtype=torch.float32).pin_memory()
        self.aux_host = torch.empty(n_scans * 16 + n_scans, dtype=torch.float64).pin_memory()     # transforms | stamps
        self.off_host = torch.empty(n_scans + 1, dtype=torch.int64).pin_memory()
        self.raw_dev = torch.empty((max_points, 4), dtype=torch.float32, device=device)
        self.aux_dev = torch.empty(n_scans * 16 + n_scans, dtype=torch.float64, device=device)
        self.off_dev = torch.empty(n_scans + 1, dtype=torch.int64, device=device)
        self.labels_host = torch.empty(max_points, dtype=torch.int32).pin_memory()
        self.conf_host = torch.empty((max_points, n_class - 1), dtype=torch.float32).pin_memory()
        self.copied = torch.cuda.Event()
        self.done = torch.cuda.Event()
        self.meta = None


class ScanPipeline:
    """submit(scans, poses) -> ticket; result(ticket) -> dict(labels int32 [Nc], confidence [Nc,C-1], boxes dict)."""

    def __init__(self, model, dt_pred=0.1, n_scans=10, max_points=2_000_000, n_class=3, learning_ignore=None,
                 learning_map_inv=None, transform=True, device=None, gather_world=1, gather_pad_rows=None, gather_group=None,
                 workers=2):
        self.model, self.dt, self.n_scans, self.transform = model, float(dt_pred), int(n_scans), bool(transform)
        self.device = device if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("insmos_b200.ScanPipeline needs a CUDA model (no CPU fallback)")
        ign = DEFAULT_IGNORE if learning_ignore is None else learning_ignore
        inv = DEFAULT_MAP_INV if learning_map_inv is None else learning_map_inv
        self.n_class = n_class
        self.ignore_mask = sum(1 << int(k) for k, v in ign.items() if v)
        self.label_map = torch.tensor([int(inv[k]) for k in range(n_class)], dtype=torch.int32, device=self.device)
        # workers > 0: each of the two slots is driven by its own host thread and CUDA stream (insmos_b200.engine), so the
        # data-dependent host reads inside one forward do not stall the other sample; workers = 0: everything is queued
        # from the calling thread on its current stream (round-1 behaviour)
        self.workers = None
        if workers:
            from insmos_b200.engine import StreamWorkers
            self.workers = StreamWorkers(self.device, 2)
        self.copy_streams = [torch.cuda.Stream(device=self.device) for _ in range(2)]
        # multi-GPU: every rank labels and returns ITS OWN sample; the one exchange step -- the gather of the
        # per-point logits of all ranks -- is a single fixed-size NCCL all_gather queued behind the forward with no host
        # synchronisation (insmos_b200.distributed.gather_logits_padded); the gathered block stays on the device
        # (result()["gathered"]) for whoever consumes the whole batch.
        self.gather_world, self.gather_group = int(gather_world), gather_group
        self.gather_pad_rows = int(gather_pad_rows) if gather_pad_rows else max_points // max(self.n_scans, 1) + 4096
        self.slots = [_Slot(self.device, max_points, self.n_scans, n_class) for _ in range(2)]
        if self.gather_world > 1:
            for sl in self.slots:
                sl.gather_buf = torch.empty((self.gather_world, self.gather_pad_rows + 1, n_class), dtype=torch.float32,
                                            device=self.device)
        self.next_ticket = 0

    def submit(self, scans, poses=None):
        """scans: list of n_scans float32 arrays [Ni,4] (x,y,z,intensity), oldest first; poses: n_scans 4x4 float64.
        The scans are packed into the slot's pinned buffer (one host memcpy) and handed to submit_packed."""
        n = len(scans)
        if n != self.n_scans:
            raise ValueError("ScanPipeline: expected %d scans, got %d" % (self.n_scans, n))
        slot = self._free_slot()
        offs = np.zeros(n + 1, dtype=np.int64)
        for i, s in enumerate(scans):
            offs[i + 1] = offs[i] + s.shape[0]
        total = int(offs[-1])
        if total > slot.raw_host.shape[0]:
            raise ValueError("ScanPipeline: %d points exceed max_points=%d" % (total, slot.raw_host.shape[0]))
        raw = slot.raw_host.numpy()
        for i, s in enumerate(scans):
            raw[offs[i]:offs[i + 1]] = s[:, :4]
        return self.submit_packed(slot.raw_host[:total], offs, poses)

    def _free_slot(self):
        slot = self.slots[self.next_ticket % 2]
        if slot.meta is not None and not slot.meta.get("collected", False):
            raise RuntimeError("ScanPipeline: collect result(%d) before submitting two more samples" % slot.meta["ticket"])
        return slot

    def submit_packed(self, raw_pinned, offsets, poses=None):
        """raw_pinned: float32 [total,4] tensor in PINNED host memory holding the n_scans scans back to back (e.g. the
        buffer the .bin files were read into), offsets: int64 [n_scans+1] row offsets.  No host copy: the H2D transfer
        reads raw_pinned directly on the copy stream; the caller must not modify it until result(ticket) returned."""
        with torch.cuda.device(self.device):           # the C ABI launches on the current device's current stream
            return self._submit_packed(raw_pinned, offsets, poses)

    def _submit_packed(self, raw_pinned, offsets, poses):
        n = self.n_scans
        slot = self._free_slot()
        offs = np.asarray(offsets, dtype=np.int64)
        total = int(offs[-1])
        if offs.shape[0] != n + 1 or raw_pinned.shape[0] < total or raw_pinned.dtype != torch.float32:
            raise ValueError("ScanPipeline.submit_packed: need float32 [total,4] and %d offsets" % (n + 1))
        if not raw_pinned.is_pinned():
            raise ValueError("ScanPipeline.submit_packed: raw_pinned must live in pinned host memory (tensor.pin_memory())")
        if total > slot.raw_dev.shape[0]:
            raise ValueError("ScanPipeline: %d points exceed max_points=%d" % (total, slot.raw_dev.shape[0]))
        aux = slot.aux_host.numpy()
        use_T = self.transform and poses is not None
        if use_T:
            aux[:n * 16] = scan_transforms(poses).reshape(-1)
        # t_i = round((i - n + 1) * dt, 3) as float32 (predict_mos.py:147-148,177)
        aux[n * 16:] = np.array([np.float32(round((i - n + 1) * self.dt, 3)) for i in range(n)], dtype=np.float64)
        slot.off_host.numpy()[:] = offs
        slot.raw_src = raw_pinned
        ticket = self.next_ticket
        self.next_ticket += 1
        slot.meta = {"ticket": ticket, "total": total, "collected": False, "job": None}
        if self.workers is not None:
            slot.meta["job"] = self.workers.submit(self._run_slot, slot, ticket % 2, total, use_T, n, worker=ticket % 2)
        else:
            self._run_slot(slot, ticket % 2, total, use_T, n)
        return ticket

    def _run_slot(self, slot, which, total, use_T, n):
        """queue one sample on the CURRENT stream: H2D (copy stream) -> staging -> forward -> labels -> D2H -> gather."""
        cur = torch.cuda.current_stream(self.device)
        cs = self.copy_streams[which]
        with torch.cuda.stream(cs):
            cs.wait_event(slot.done)                                 # the slot's previous forward has consumed raw_dev
            slot.raw_dev[:total].copy_(slot.raw_src[:total], non_blocking=True)
            slot.aux_dev.copy_(slot.aux_host, non_blocking=True)
            slot.off_dev.copy_(slot.off_host, non_blocking=True)
            slot.copied.record(cs)
        cur.wait_event(slot.copied)
        T = slot.aux_dev[:n * 16].view(n, 4, 4) if use_T else None
        stamps = slot.aux_dev[n * 16:].to(torch.float32)
        pts = ops.stage_scans(slot.raw_dev[:total], slot.off_dev, T, stamps)
        with torch.no_grad():
            boxes, _, logits = self.model.forward([{"meta": None, "past_point_clouds": pts, "batch_size_npast": n}], "test")
        labels, conf = ops.mos_labels(logits[0], self.ignore_mask, self.label_map)
        nc = labels.shape[0]
        slot.labels_host[:nc].copy_(labels, non_blocking=True)
        slot.conf_host[:nc].copy_(conf, non_blocking=True)
        slot.done.record(cur)
        slot.meta.update({"nc": nc, "boxes": boxes[0][0], "logits": logits[0], "gathered": None})

    def result(self, ticket):
        slot = self.slots[ticket % 2]
        if slot.meta is None or slot.meta["ticket"] != ticket:
            raise KeyError("ScanPipeline: ticket %d is not in flight" % ticket)
        if slot.meta["job"] is not None:
            slot.meta["job"].finished.wait()                         # the worker has queued everything (or failed)
            if slot.meta["job"].error is not None:
                raise slot.meta["job"].error
        if self.gather_world > 1 and slot.meta["gathered"] is None:
            # the one exchange step, issued HERE -- on the caller's thread and stream, in ticket order -- so that every rank
            # queues its collectives in the same order whatever the interleaving of the worker threads
            from insmos_b200.distributed import gather_logits_padded
            with torch.cuda.device(self.device):
                main = torch.cuda.current_stream(self.device)
                main.wait_event(slot.done)
                lg = slot.meta["logits"]
                lg.record_stream(main)
                slot.meta["gathered"] = gather_logits_padded(lg, self.gather_world, self.gather_pad_rows,
                                                             group=self.gather_group, out=slot.gather_buf)
        slot.done.synchronize()
        nc = slot.meta["nc"]
        slot.meta["collected"] = True
        return {"labels": slot.labels_host[:nc].numpy(), "confidence": slot.conf_host[:nc].numpy(), "boxes": slot.meta["boxes"],
                "gathered": slot.meta["gathered"]}
