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
