"""Concurrent forwards on several CUDA streams / host threads (insmos_b200.engine) and the threaded ScanPipeline give the
same bits as the plain single-stream forward: results must not depend on scheduling."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import golden_util  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small(cuda):
    from test_gpu_model import _net
    meta, shapes, sd, pts, gold = golden_util.load("small")
    return _net(cuda, sd), pts, gold


def test_forward_pool_matches_direct_forward_bit_for_bit(small, cuda):
    from insmos_b200 import synth
    from insmos_b200.engine import ForwardPool
    net, pts, gold = small
    clouds = [torch.from_numpy(pts).to(cuda)] + [torch.from_numpy(synth.make_sequence(seed=20 + i, n_scans=3, n_elev=32, n_azim=450)).to(cuda)
                                                  for i in range(3)]
    direct = []
    with torch.no_grad():
        for c in clouds:
            b, _, lg = net.forward([{"meta": None, "past_point_clouds": c, "batch_size_npast": 3}], "test")
            direct.append((lg[0].clone(), b[0][0]["pred_boxes"].clone()))
    torch.cuda.synchronize()
    for workers in (1, 2, 3):
        pool = ForwardPool(net, workers=workers, n_past=3)
        jobs = [pool.submit_points(clouds[i % 4]) for i in range(12)]
        for i, j in enumerate(jobs):
            lg, boxes = j.wait()
            torch.cuda.current_stream().synchronize()
            assert torch.equal(lg, direct[i % 4][0]), "workers=%d job %d: logits differ from the single-stream forward" % (workers, i)
            assert torch.equal(boxes["pred_boxes"], direct[i % 4][1])
        pool.close()
    err = (direct[0][0].cpu() - torch.from_numpy(gold["logits"])).abs().max().item()
    assert err < 1e-3 or gold["logits"].shape != tuple(direct[0][0].shape)


def test_forward_pool_surfaces_errors(small, cuda):
    from insmos_b200.engine import ForwardPool
    net, pts, _ = small
    pool = ForwardPool(net, workers=2, n_past=3)
    job = pool.submit_points(torch.zeros((8, 5)))                # CPU tensor: the product refuses loudly
    with pytest.raises(RuntimeError):
        job.wait()
    ok = pool.submit_points(torch.from_numpy(pts).to(cuda))      # the pool survives
    lg, _ = ok.wait()
    assert lg.shape[1] == 3
    pool.close()


def test_scan_pipeline_threaded_equals_inline(small, cuda):
    from insmos_b200.pipeline import ScanPipeline
    net, pts, _ = small
    stamps = np.unique(pts[:, 4])
    scans = [np.ascontiguousarray(pts[pts[:, 4] == t][:, :4]) for t in stamps]
    poses = [np.eye(4)] * len(scans)
    ref = None
    for workers in (0, 2):
        pipe = ScanPipeline(net, dt_pred=0.1, n_scans=len(scans), max_points=pts.shape[0] + 16, workers=workers)
        tickets, outs = [], []
        for i in range(5):
            tickets.append(pipe.submit(scans, poses))
            if len(tickets) == 2:
                outs.append(pipe.result(tickets.pop(0)))
        outs.append(pipe.result(tickets.pop(0)))
        labels = [o["labels"].copy() for o in outs]
        conf = [o["confidence"].copy() for o in outs]
        assert all(np.array_equal(labels[0], l) for l in labels)
        if ref is None:
            ref = (labels[0], conf[0])
        else:
            assert np.array_equal(ref[0], labels[0]) and np.array_equal(ref[1], conf[0])
