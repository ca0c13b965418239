"""Drop-in proof for the caller-facing boundary: the reference's OWN scripts/predict_mos.py, imported
unchanged from /root/reference over insmos_b200/compat (MinkowskiEngine / spconv / pytorch_lightning / easydict / models
import names), driven on a small synthetic KITTI-layout directory with a Lightning-format checkpoint.

CPU part (runs wherever /root/reference exists -- the development container): import, DemoDataset over the synthetic
directory (poses.txt / calib.txt / velodyne/*.bin), the batch dict it yields vs the oracle's staging restatement, strict
load of {"state_dict", "hyper_parameters"} through InsMOSNet.load_from_checkpoint(ckpt, hparams=cfg), the label mapping.
GPU part (-m gpu): predict_mos.main() itself end to end -- .label / confidence / bbox files vs ScanPipeline on the same
scans.  The reference sources are never copied, so both parts skip where /root/reference is absent (the GPU box).
"""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch
import yaml

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "scripts", "predict_mos.py")),
                               reason="the reference repository is not present on this machine")

N_PAST, N_FRAMES = 10, 12


def _velo_pose(i):
    """pose of frame i (velodyne -> world): 1 m per frame along x with a slow yaw."""
    yaw = 0.01 * i
    T = np.eye(4)
    T[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
    T[0, 3] = 1.0 * i
    return T


def make_kitti_dir(root, seq=8):
    """sequences/<seq>/{velodyne/%06d.bin, poses.txt, calib.txt} from the synthetic generator; returns the raw scans."""
    from insmos_b200 import synth
    d = os.path.join(root, "%02d" % seq)
    os.makedirs(os.path.join(d, "velodyne"), exist_ok=True)
    Tr = np.array([[0.0, -1.0, 0.0, 0.1], [0.0, 0.0, -1.0, -0.2], [1.0, 0.0, 0.0, 0.3], [0, 0, 0, 1.0]])   # T_cam_velo
    scans, lines = [], []
    for i in range(N_FRAMES):
        pts = synth.make_sequence(seed=40 + i, n_scans=1, n_elev=16, n_azim=180)[:, :4].astype(np.float32)
        pts.tofile(os.path.join(d, "velodyne", "%06d.bin" % i))
        scans.append(pts)
        cam = Tr @ _velo_pose(i) @ np.linalg.inv(Tr)                                    # KITTI poses are camera poses
        lines.append(" ".join("%.12e" % v for v in cam[:3].reshape(-1)))
    open(os.path.join(d, "poses.txt"), "w").write("\n".join(lines) + "\n")
    open(os.path.join(d, "calib.txt"), "w").write("P0: 1 0 0 0 0 1 0 0 0 0 1 0\nTr: " + " ".join("%.12e" % v for v in Tr[:3].reshape(-1)) + "\n")
    return scans


@pytest.fixture(scope="module")
def dropin(tmp_path_factory):
    import insmos_b200
    compat = insmos_b200.install()
    if REF not in sys.path:
        sys.path.append(REF)                                   # AFTER compat: `models` resolves to the product, `dataloader` to the reference
    assert sys.path.index(compat) < sys.path.index(REF)
    spec = importlib.util.spec_from_file_location("ref_predict_mos", os.path.join(REF, "scripts", "predict_mos.py"))
    mod = importlib.util.module_from_spec(spec)
    cwd = os.getcwd()
    os.chdir(REF)                                              # the script appends '.' to sys.path and reads ./config/*.yaml
    try:
        spec.loader.exec_module(mod)
        cfg = yaml.safe_load(open(os.path.join(REF, "config", "config.yaml")))
    finally:
        os.chdir(cwd)
    root = str(tmp_path_factory.mktemp("kitti"))
    scans = make_kitti_dir(root)
    cfg["DATA"]["SEMANTIC_CONFIG_FILE"] = os.path.join(REF, "config", "semantic-kitti-mos.yaml")
    cfg["DATA"]["NUM_WORKER"] = 0
    cfg["DATA"]["SPLIT"]["TEST"] = [8]
    # Lightning-format checkpoint from the seeded weights
    from insmos_b200 import synth_weights
    import models.models as product_models
    assert product_models.__file__.startswith(os.path.join(ROOT, "insmos_b200")), "models.models must resolve to the product"
    assert mod.models is product_models
    net = product_models.InsMOSNet(cfg)
    sd = synth_weights.fill_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()})
    sd["model.unet.center_head.conv_cls.bias"] = torch.zeros(3)
    ckpt = os.path.join(root, "insmos_synth.ckpt")
    torch.save({"state_dict": sd, "hyper_parameters": cfg, "epoch": 0, "pytorch-lightning_version": "1.5.10"}, ckpt)
    return {"mod": mod, "cfg": cfg, "root": root, "scans": scans, "ckpt": ckpt, "sd": sd}


@needs_ref
def test_reference_script_imports_over_the_compat_surface(dropin):
    mod = dropin["mod"]
    import MinkowskiEngine
    import spconv.pytorch
    import pytorch_lightning
    for m in (MinkowskiEngine, spconv.pytorch, pytorch_lightning):
        assert m.__file__.startswith(os.path.join(ROOT, "insmos_b200", "compat"))
    assert mod.load_poses.__module__ == "dataloader.utils"                       # the reference's own loader
    assert sys.modules["dataloader.utils"].__file__.startswith(REF)
    assert hasattr(mod, "DemoDataset") and hasattr(mod, "main") and hasattr(mod, "to_original_labels")


@needs_ref
def test_demo_dataset_batch_matches_oracle_staging(dropin):
    """predict_mos.py:114-159 on the synthetic directory == oracle/staging.py == what ScanPipeline's kernel is tested against."""
    from oracle import staging
    mod, cfg = dropin["mod"], dropin["cfg"]
    ds = mod.DemoDataset(cfg, dropin["root"], split="test")
    assert len(ds) == N_FRAMES - (N_PAST - 1)
    for idx in (0, len(ds) - 1):
        item = ds[idx]
        seq, scan_idx, files = item["meta"]
        assert seq == 8 and scan_idx == N_PAST - 1 + idx and len(files) == N_PAST and item["batch_size_npast"] == N_PAST
        pc = item["past_point_clouds"]
        assert pc.dtype == torch.float32 and pc.shape[1] == 5
        poses = ds.poses[8][scan_idx - N_PAST + 1: scan_idx + 1]
        want = staging.stage_scans([dropin["scans"][i] for i in range(scan_idx - N_PAST + 1, scan_idx + 1)], list(poses), 0.1)
        assert np.array_equal(pc.numpy(), want), "DemoDataset.__getitem__ differs from the oracle staging restatement"
        assert np.allclose(np.unique(pc[:, 4].numpy()), np.round(np.arange(-(N_PAST - 1), 1) * 0.1, 3))
    batch = mod.DemoDataset.collate_batch_test([ds[0]])
    assert isinstance(batch, list) and set(batch[0]) == {"meta", "past_point_clouds", "batch_size_npast"}


@needs_ref
def test_lightning_checkpoint_loads_strictly_through_the_reference_call(dropin):
    """predict_mos.py:288,326: cfg = torch.load(ckpt)['hyper_parameters']; InsMOSNet.load_from_checkpoint(ckpt, hparams=cfg)"""
    mod, sd = dropin["mod"], dropin["sd"]
    cfg = torch.load(dropin["ckpt"])["hyper_parameters"]
    assert cfg["MODEL"]["N_PAST_STEPS"] == N_PAST
    model = mod.models.InsMOSNet.load_from_checkpoint(dropin["ckpt"], hparams=cfg)
    model.eval()
    got = model.state_dict()
    assert set(got) == set(sd) and len(got) == 394
    for k in ("model.motion_encoder.MinkUNet.conv0p1s1.kernel", "model.unet.conv_input.0.weight", "model.unet.mos_seg_layer.bias"):
        assert torch.equal(got[k], sd[k])
    assert model.hparams["MODEL"]["DENSE_HEAD"]["NUM_CLASS"] == 3
    with pytest.raises(ValueError):
        model.forward([], "validate")                          # the three modes of the reference: 'train', 'eval', 'test'
    with pytest.raises(RuntimeError):
        model.forward([], "train")                             # 'train' is implemented (N3) but needs model.train()
    with pytest.raises(RuntimeError):                          # no CPU fallback: the script's .cuda() is mandatory
        model.forward([{"meta": None, "past_point_clouds": torch.zeros((4, 5)), "batch_size_npast": N_PAST}], "test")


@needs_ref
def test_label_mapping_matches_the_reference_function(dropin):
    from insmos_b200 import pipeline
    sem = yaml.safe_load(open(dropin["cfg"]["DATA"]["SEMANTIC_CONFIG_FILE"]))
    lab = np.array([0, 1, 2, 2, 1, 0], dtype=np.int64)
    want = dropin["mod"].to_original_labels(lab, sem)
    got = np.array([pipeline.DEFAULT_MAP_INV[int(v)] for v in lab])
    assert np.array_equal(want, got)
    assert {k: bool(v) for k, v in sem["learning_ignore"].items()} == pipeline.DEFAULT_IGNORE


@needs_ref
@pytest.mark.gpu
def test_predict_mos_main_end_to_end(dropin, cuda, tmp_path, monkeypatch):
    """the unchanged script's main(): files under preb_out/<id>/... vs ScanPipeline on the same scans."""
    from insmos_b200.pipeline import ScanPipeline
    mod, cfg = dropin["mod"], dropin["cfg"]
    monkeypatch.chdir(tmp_path)
    monkeypatch.setattr(sys, "argv", ["predict_mos.py", "--ckpt", dropin["ckpt"], "--data_path", dropin["root"], "--split", "valid"])
    mod.main()
    out = os.path.join(str(tmp_path), "preb_out", cfg["EXPERIMENT"]["ID"])
    labels = sorted(os.listdir(os.path.join(out, "mos_preb", "sequences", "08", "predictions")))
    assert len(labels) == N_FRAMES                           # the first N_PAST scans (growing window) + the rest
    model = mod.models.InsMOSNet.load_from_checkpoint(dropin["ckpt"], hparams=cfg).cuda().eval()
    ds = mod.DemoDataset(cfg, dropin["root"], split="test")
    pipe = ScanPipeline(model, dt_pred=0.1, n_scans=N_PAST, max_points=200_000)
    for idx in range(len(ds)):
        scan_idx = N_PAST - 1 + idx
        r = pipe.result(pipe.submit([dropin["scans"][i] for i in range(scan_idx - N_PAST + 1, scan_idx + 1)],
                                    list(ds.poses[8][scan_idx - N_PAST + 1: scan_idx + 1])))
        got = np.fromfile(os.path.join(out, "mos_preb", "sequences", "08", "predictions", "%06d.label" % scan_idx), dtype=np.int32)
        assert np.array_equal(got, r["labels"])
        conf = np.load(os.path.join(out, "confidence", "sequences", "08", "predictions", "%06d.npy" % scan_idx))
        assert np.abs(conf - r["confidence"]).max() < 1e-6
