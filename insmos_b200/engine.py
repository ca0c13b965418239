"""Concurrent forwards: N host threads, each with its own CUDA stream, run InsMOSNet.forward on independent samples.

Why.  One forward is ~180 kernel launches with ~11 small device->host reads of data-dependent sizes (voxel counts per
level, candidate count); each read drains the stream and the GPU idles until the host has queued the next kernels
(measured: 7.0 ms per step for 6.2 ms of kernels), and many of the kernels on the coarse levels do not
fill 148 SMs.  Samples are independent (models/models.py:313 processes them one by one), so two or three forwards in
flight on separate streams fill each other's bubbles: the C-ABI calls only queue kernels (and keep the GIL: PyDLL, see _lib.load) and the
reads block outside the GIL.  Streams and threads instead of a static-shape graph capture: every sample has its own shapes.
"""
import queue
import threading

import torch


class _Job:
    __slots__ = ("fn", "args", "ready", "device", "done_event", "result", "error", "finished")

    def __init__(self, fn, args, ready, device):
        self.fn, self.args, self.ready, self.device = fn, args, ready, device
        self.done_event, self.result, self.error = None, None, None
        self.finished = threading.Event()

    def wait(self, stream=None):
        """block the HOST until the job's kernels are queued, then make `stream` (default: current) wait for them on the
        DEVICE; returns the job's result.  No device synchronisation on the host."""
        self.finished.wait()
        if self.error is not None:
            raise self.error
        (stream or torch.cuda.current_stream(self.device)).wait_event(self.done_event)
        return self.result


class StreamWorkers:
    """`workers` threads bound to `device`, one CUDA stream each; submit(fn, *args) runs fn on the next worker's stream."""

    def __init__(self, device, workers=2):
        self.device = torch.device(device)
        self.n = int(workers)
        self.queues = [queue.Queue() for _ in range(self.n)]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.n)]
        self.threads = [threading.Thread(target=self._loop, args=(i,), daemon=True) for i in range(self.n)]
        self._next = 0
        # Cold start: the first job fills the per-layer caches (prepared weight images, folded BatchNorm constants) with
        # kernels queued on ITS stream; a second job on another stream would find the cache entries and could read them before
        # those kernels have run.  Until one job has completed on the device the jobs therefore run one at a time.
        self._primed = False
        self._prime_lock = threading.Lock()
        for t in self.threads:
            t.start()

    def _loop(self, i):
        torch.cuda.set_device(self.device)
        stream = self.streams[i]
        while True:
            job = self.queues[i].get()
            if job is None:
                return
            cold = not self._primed
            if cold:
                self._prime_lock.acquire()
            try:
                with torch.cuda.stream(stream), torch.no_grad():
                    if job.ready is not None:
                        stream.wait_event(job.ready)                 # inputs produced on the submitter's stream
                    job.result = job.fn(*job.args)
                    job.done_event = torch.cuda.Event()
                    job.done_event.record(stream)
                    if cold and not self._primed:
                        stream.synchronize()                         # what this job cached is now complete on the device
                        self._primed = True
            except BaseException as e:                               # surfaced by wait()
                job.error = e
            finally:
                if cold:
                    self._prime_lock.release()
            job.finished.set()

    def submit(self, fn, *args, worker=None):
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        job = _Job(fn, args, ready, self.device)
        w = self._next if worker is None else int(worker) % self.n
        if worker is None:
            self._next = (self._next + 1) % self.n
        self.queues[w].put(job)
        return job

    def close(self):
        for q in self.queues:
            q.put(None)
        for t in self.threads:
            t.join(timeout=5)


class ForwardPool(StreamWorkers):
    """forward(points) of one model on `workers` concurrent streams: submit(points) -> job; job.wait() -> (logits, boxes)."""

    def __init__(self, net, workers=2, n_past=10):
        super().__init__(next(net.parameters()).device, workers)
        self.net, self.n_past = net, n_past

    def _forward(self, pts):
        boxes, _, logits = self.net.forward([{"meta": None, "past_point_clouds": pts, "batch_size_npast": self.n_past}], "test")
        return logits[0], boxes[0][0]

    def submit_points(self, pts):
        return self.submit(self._forward, pts)
