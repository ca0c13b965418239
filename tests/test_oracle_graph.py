"""oracle/graph.py (the travelling restatement of the model graph) vs the golden fixtures produced by the
reference's OWN models/*.py (tests/golden/make_golden.py).  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import golden_util  # noqa: E402
from oracle import graph  # noqa: E402


@pytest.mark.parametrize("name", ["small_nodet", "small"])
def test_graph_matches_reference_code(name):
    meta, shapes, sd, pts, gold = golden_util.load(name)
    r = graph.forward(sd, pts)
    assert np.array_equal(r["voxel_coords"], gold["voxel_coords"])
    assert np.array_equal(r["pc_voxel_id"], gold["pc_voxel_id"])
    assert torch.allclose(r["current_point"], torch.from_numpy(gold["current_point"]), atol=1e-5, rtol=1e-5)
    assert r["n_cand"] == int(gold["n_cand"])
    assert torch.allclose(r["pred_boxes"], torch.from_numpy(gold["pred_boxes"]), atol=1e-5)
    assert torch.equal(r["pred_labels"], torch.from_numpy(gold["pred_labels"]))
    assert torch.allclose(r["logits"], torch.from_numpy(gold["logits"]), atol=1e-4, rtol=1e-5), \
        (r["logits"] - torch.from_numpy(gold["logits"])).abs().max()


def test_state_dict_contract_matches_reference():
    """the product mirror must expose exactly the reference's state_dict keys and shapes."""
    meta, shapes, _, _, _ = golden_util.load("small_nodet")
    import insmos_b200
    insmos_b200.install()
    from models.models import InsMOSNet
    from insmos_b200.config import default_config
    mine = {k: tuple(v.shape) for k, v in InsMOSNet(default_config()).state_dict().items()}
    assert set(mine) == set(shapes), (sorted(set(mine) ^ set(shapes))[:10])
    assert all(mine[k] == shapes[k] for k in shapes)
    assert meta["config"] == default_config()["MODEL"]["POST_PROCESSING"]
