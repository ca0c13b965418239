"""Host-side logic that needs no GPU: tile-row policy, kernel selection tables, pose transforms of the scan pipeline,
loud failure on CPU tensors (the product has no CPU path)."""
import numpy as np
import pytest
import torch

from insmos_b200 import _lib, ops, pipeline
from oracle import staging


def test_tile_rows_policy_is_valid_for_every_map_family():
    for K in (1, 3, 8, 27, 81, 125):
        for n in (0, 1, 17, 6090, 28277, 75533, 172603, 376904, 3_000_000):
            tm = ops.choose_tile_rows(n, K)
            assert tm in (16, 32, 64, 128) and tm * K < 65536            # 16-bit bucket offsets, 7-bit row-in-tile
    assert ops.choose_tile_rows(376904, 125) == 64 and ops.choose_tile_rows(376904, 81) == 128
    assert ops.choose_tile_rows(6090, 27) == 16 and ops.choose_tile_rows(42098, 27) == 64


def test_tcgen05_eligibility_table():
    assert ops.umma_eligible(27, 128, 128) and ops.umma_eligible(27, 35, 32) and ops.umma_eligible(9, 256, 128)
    assert not ops.umma_eligible(27, 64, 8)            # N must be a multiple of 16
    assert not ops.umma_eligible(27, 64, 24)
    assert not ops.umma_eligible(200, 64, 64)          # neighbour table limited to 128 offsets
    assert not ops.umma_eligible(27, 64, 256) and not ops.umma_eligible(27, 64, 192)      # one TMEM tile: Cout <= 128
    assert ops.tc_eligible(81, 48, 32, 128) and not ops.tc_eligible(81, 19, 16, 64) and not ops.tc_eligible(27, 64, 64, 64)


def test_kernel_count_table_names_exist_in_the_abi():
    missing = [k for k in _lib.KERNELS_PER_CALL if k not in _lib.PROTOTYPES]
    assert not missing, missing


def test_pipeline_pose_transforms_match_the_oracle_bit_for_bit():
    rng = np.random.default_rng(5)
    poses = []
    for i in range(10):
        a = rng.normal(0, 0.3)
        T = np.eye(4)
        T[:3, :3] = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
        T[:3, 3] = rng.normal(0, 5, 3)
        poses.append(T)
    got = pipeline.scan_transforms(poses)
    exp = np.stack(staging.scan_transforms(poses))
    assert got.dtype == np.float64 and np.array_equal(got, exp)
    assert np.allclose(got[-1], np.eye(4), atol=1e-15)                  # newest scan is the target frame


def test_product_refuses_cpu_tensors_and_cpu_models():
    x = torch.zeros((4, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.linear(x, torch.zeros((8, 8)))
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.voxelize4d(torch.zeros((4, 5)), [0.1, 0.1, 0.1, 0.1])
    with pytest.raises(RuntimeError, match="CUDA"):
        pipeline.ScanPipeline(torch.nn.Linear(2, 2), n_scans=2, max_points=16)


# ---- training step host logic (N3): flat buffers, StepLR schedule, refusal of the CUDA-only pieces on CPU ------------------
def test_flat_parameters_are_views_of_one_buffer_and_gradients_accumulate_in_place():
    import torch
    from insmos_b200.train import FlatParameters, TrainStep
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.BatchNorm1d(3), torch.nn.Linear(3, 2))
    before = [p.detach().clone() for p in net.parameters()]
    flat = FlatParameters(net)
    assert flat.numel == sum(p.numel() for p in net.parameters())
    o = 0
    for p, b in zip(net.parameters(), before):
        assert torch.equal(p.detach(), b)                                   # values preserved
        assert p.data_ptr() == flat.data.data_ptr() + 4 * o and p.grad.data_ptr() == flat.grad.data_ptr() + 4 * o
        o += p.numel()
    net(torch.randn(8, 4)).sum().backward()
    g1 = flat.grad.clone()
    assert float(g1.abs().sum()) > 0
    net(torch.randn(8, 4)).sum().backward()                                 # autograd accumulates INTO the flat buffer
    assert not torch.equal(flat.grad, g1)
    net.zero_grad(set_to_none=True)                                         # a caller that drops .grad: re-attached by zero_grad()
    flat.zero_grad()
    assert all(p.grad is not None and p.grad.data_ptr() >= flat.grad.data_ptr() for p in net.parameters())
    assert float(flat.grad.abs().sum()) == 0.0
    ts = TrainStep.__new__(TrainStep)
    ts.base_lr, ts.lr_decay, ts.lr_epoch, ts.epoch = 1e-4, 0.99, 1, 0
    assert ts.lr == 1e-4
    ts.set_epoch(10)
    assert abs(ts.lr - 1e-4 * 0.99 ** 10) < 1e-18                           # StepLR(step_size=1, gamma=0.99), models.py:187-189
    ts.lr_epoch = 4
    assert abs(ts.lr - 1e-4 * 0.99 ** 2) < 1e-18


def test_training_kernels_refuse_cpu_tensors():
    import pytest
    import torch
    from insmos_b200 import autograd, ops
    x = torch.randn(16, 4, requires_grad=True)
    assert autograd.needs_grad(x) and not autograd.needs_grad(x.detach(), None)
    with torch.no_grad():
        assert not autograd.needs_grad(x)
    with pytest.raises(RuntimeError):
        autograd.batch_norm_train(torch.nn.BatchNorm1d(4), x)
    with pytest.raises(RuntimeError):
        autograd.gather_rows(x, torch.zeros(3, dtype=torch.int32))
    with pytest.raises((RuntimeError, ValueError)):
        ops.adam_step(torch.zeros(4), torch.zeros(4), torch.zeros(4), torch.zeros(4), 1e-3, 0.9, 0.999, 1e-8, 0.0, 1)
    with pytest.raises(RuntimeError):
        ops.center_targets(torch.zeros((1, 8)), 100, 250, 300, 3, -60, -50, 0.1, 0.1, 4, 0.1, 2)


def test_reference_arm_prints_one_contract_line(tmp_path):
    """`bench.py --impl reference` needs no GPU: it times the CPU port on the FULL C2 cloud and prints exactly one JSON line
    with the contract's keys, the counts of steps actually run, and the rule-book / convolution split."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, INSMOS_REF_BUDGET_S="1", OMP_NUM_THREADS="1")         # one timed step; torchrun-like OMP setting
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "scans_per_sec" and d["unit"] == "scans/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 0 and d["requested"] == {"steps": 3, "warmup": 0}
    assert d["value"] > 0 and abs(d["value"] - 1000.0 / d["ms_per_step"]) < 1e-9
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and cb["steps"] == 1
    assert cb["rulebook_s"] > 0 and cb["me_conv_s"] > 0 and "no extrapolation" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "C2" in d["config"]["workload"]


def test_stream_workers_serialise_the_cold_start_then_run_concurrently(monkeypatch):
    """engine.StreamWorkers control flow on fake CUDA objects: until one job has completed on the device (stream.synchronize
    after the first successful job) the jobs run one at a time -- the first forward fills the per-layer weight caches on its own
    stream --, afterwards the workers overlap; a failing first job is surfaced by wait() and does not prime the pool."""
    import contextlib
    import threading
    import time
    from insmos_b200 import engine

    class FakeStream:
        def __init__(self, device=None):
            self.syncs = 0

        def wait_event(self, ev):
            pass

        def synchronize(self):
            self.syncs += 1

    class FakeEvent:
        def __init__(self, enable_timing=False):
            pass

        def record(self, stream=None):
            pass

    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: FakeStream())
    log, lock = [], threading.Lock()

    def work(tag, fail=False):
        t0 = time.perf_counter()
        time.sleep(0.08)
        if fail:
            raise ValueError("boom %s" % tag)
        with lock:
            log.append((tag, t0, time.perf_counter()))
        return tag

    pool = engine.StreamWorkers("cuda:0", workers=2)
    try:
        bad = pool.submit(work, "bad", True)                     # first job fails: surfaced, pool stays cold
        with pytest.raises(ValueError):
            bad.wait(FakeStream())
        assert not pool._primed and sum(s.syncs for s in pool.streams) == 0
        a, b = pool.submit(work, "a"), pool.submit(work, "b")   # cold: one at a time, exactly one priming synchronize
        assert {a.wait(FakeStream()), b.wait(FakeStream())} == {"a", "b"}
        assert pool._primed and sum(s.syncs for s in pool.streams) == 1
        iv = {t: (s, e) for t, s, e in log}
        assert iv["a"][1] <= iv["b"][0] or iv["b"][1] <= iv["a"][0]            # no overlap
        c, d = pool.submit(work, "c"), pool.submit(work, "d")   # primed: the two workers overlap
        c.wait(FakeStream()); d.wait(FakeStream())
        iv = {t: (s, e) for t, s, e in log}
        assert iv["c"][0] < iv["d"][1] and iv["d"][0] < iv["c"][1]
        assert sum(s.syncs for s in pool.streams) == 1
    finally:
        pool.close()
