"""InsMOS_Model / InsMOSNet mirrors (models/models.py:27-59,269-376): same constructor, state_dict keys
and forward(batch_data, Model_mode) contract; the inference mode 'test' only ('train' / 'eval' raise)."""
import numpy as np
import torch
import torch.nn as nn
import yaml

from pytorch_lightning.core.lightning import LightningModule

from insmos_b200 import config as _config
from insmos_b200 import ops
from .motion import MotionNet
from .unet3d import UNetV2


class VoxelGenerate(nn.Module):
    """voxel_generate.py:8-31 + mean_vfe.py:47-52 in ONE pass: capped hashed voxelisation with ids and the
    per-voxel mean of the first <=5 points (the [100000,5,7] staging tensor of the reference is never built)."""

    def __init__(self, voxel_size, point_cloud_range, max_number_of_voxel, max_point_per_voxel, num_point_feature):
        super().__init__()
        self.voxel_size = [float(v) for v in voxel_size]
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.max_number_of_voxel, self.max_point_per_voxel = max_number_of_voxel, max_point_per_voxel
        self.num_point_feature = num_point_feature
        self.grid = [int(round((self.point_cloud_range[3 + d] - self.point_cloud_range[d]) / self.voxel_size[d])) for d in range(3)]

    def forward(self, batch_dict):
        r = ops.voxelize3d(batch_dict["current_point"], self.point_cloud_range, self.voxel_size, self.grid,
                           self.max_number_of_voxel, self.max_point_per_voxel, want_voxels=False)
        batch_dict["voxel_coords"] = r["set"].coords                     # [M,4] (0,z,y,x) int32
        batch_dict["_voxel_set"] = r["set"]
        batch_dict["voxel_num_points"] = r["num_points"]
        batch_dict["pc_voxel_id"] = r["pc_voxel_id"]
        batch_dict["list_pc_voxel_id"] = [r["pc_voxel_id"]]
        batch_dict["voxel_features"] = r["mean"]                          # MeanVFE output
        return batch_dict


class MeanVFE(nn.Module):
    """kept for the module tree; the mean is produced by VoxelGenerate's fused kernel."""

    def __init__(self, model_cfg, num_point_features, **kwargs):
        super().__init__()
        self.model_cfg, self.num_point_features = model_cfg, num_point_features

    def get_output_feature_dim(self):
        return self.num_point_features

    def forward(self, batch_dict, **kwargs):
        if "voxel_features" not in batch_dict:
            v, n = batch_dict["voxels"], batch_dict["voxel_num_points"]
            batch_dict["voxel_features"] = (v.sum(dim=1) / torch.clamp_min(n.view(-1, 1), 1.0).type_as(v)).contiguous()
        return batch_dict


class _NLLWeight(nn.Module):
    """holds the `weight` buffer of the reference's nn.NLLLoss (state_dict key model.MOSLoss.loss.weight)."""

    def __init__(self, weight):
        super().__init__()
        self.register_buffer("weight", weight)


class MOSLoss(nn.Module):
    def __init__(self, n_classes, ignore_index):
        super().__init__()
        self.n_classes, self.ignore_index = n_classes, ignore_index
        w = [0.0 if i in ignore_index else 1.0 for i in range(n_classes)]
        self.loss = _NLLWeight(torch.Tensor([x / sum(w) for x in w]))

    def compute_loss(self, out, past_labels):
        """loss.py:20-34: ignored classes to -inf (written into `out` in place, as the reference does), softmax, log of the
        clamped probabilities, class-weighted NLL."""
        logits = out
        logits[:, self.ignore_index] = -float("inf")
        log_softmax = torch.log(torch.softmax(logits, dim=1).clamp(min=1e-8))
        return torch.nn.functional.nll_loss(log_softmax, past_labels.long(), weight=self.loss.weight)


class InsMOS_Model(nn.Module):
    def __init__(self, cfg, n_mos_classes, ignore_index):
        super().__init__()
        M, D = cfg["MODEL"], cfg["DATA"]
        self.dt_prediction = M["DELTA_T_PREDICTION"]
        self.post_process = M["POST_PROCESSING"]
        self.num_class = M["DENSE_HEAD"]["NUM_CLASS"]
        self.point_cloud_range = np.array(D["POINT_CLOUD_RANGE"])
        self.voxel_size = D["VOXEL_SIZE"]
        self.grid_size = np.round((self.point_cloud_range[3:6] - self.point_cloud_range[0:3]) / np.array(self.voxel_size)).astype(np.int64)
        self.n_past_step = M["N_PAST_STEPS"]
        self.mos_class = n_mos_classes
        in_channel = len(M["POINT_FEATURE_ENCODING"]["src_feature_list"]) + 3      # + 3 motion logits
        self.voxel_generate = VoxelGenerate(self.voxel_size, self.point_cloud_range, 100000, 5, in_channel)
        self.vfe = MeanVFE(M["VFE"], in_channel)
        self.unet = UNetV2(cfg, in_channel, self.grid_size, self.voxel_size, self.point_cloud_range, self.mos_class)
        self.motion_encoder = MotionNet(self.dt_prediction, self.voxel_size, self.mos_class)
        self.MOSLoss = MOSLoss(self.mos_class, ignore_index)
        self.use_motion_loss = M["USE_MOTION_LOSS"]

    def forward(self, list_batch_dict, Model_mode):
        if Model_mode == "train":
            return self.forward_train(list_batch_dict)
        if Model_mode == "eval":
            return self.forward_eval(list_batch_dict)
        if Model_mode != "test":
            raise ValueError("Model_mode %r: 'train', 'eval' or 'test'" % (Model_mode,))
        boxes_out, recall_out, logits_out = [], [], []
        for batch_dict in list_batch_dict:
            batch_dict = self.motion_encoder(batch_dict)
            if not self.use_motion_loss:
                batch_dict["current_motion_feature"] = batch_dict["current_motion_feature"][:, :3]
            batch_dict = self.voxel_generate(batch_dict)
            batch_dict = self.vfe(batch_dict)
            point_seg, pred_dicts, recall_dicts = self.unet(batch_dict, Model_mode)
            boxes_out.append(pred_dicts)
            recall_out.append(recall_dicts)
            logits_out.append(point_seg)
        return boxes_out, recall_out, logits_out


def _forward_train(self, list_batch_dict):
    """models/models.py:297-347,366-368: per sample  loss_rpn + loss_mos + loss_motion_encoder, averaged over the list.
    The loss dictionary values stay 0-dim device tensors (the reference calls .item() on each: five host syncs per sample)."""
    if not self.training:
        raise RuntimeError("Model_mode 'train' needs model.train() (batch statistics, torch BEV modules)")
    dev = list_batch_dict[0]["past_point_clouds"].device
    loss = torch.zeros(1, device=dev)
    train_loss_dict, gt_list, pred_list = [], [], []
    for batch_dict in list_batch_dict:
        batch_dict = self.motion_encoder(batch_dict)
        if not self.use_motion_loss:
            batch_dict["current_motion_feature"] = batch_dict["current_motion_feature"][:, :3]
        gt = batch_dict["past_labels"][-1]
        loss_motion = self.MOSLoss.compute_loss(batch_dict["current_motion_feature"], gt)
        batch_dict = self.voxel_generate(batch_dict)
        batch_dict = self.vfe(batch_dict)
        (loss_rpn, tb), point_seg = self.unet(batch_dict, "train")
        loss_mos = self.MOSLoss.compute_loss(point_seg, gt)
        loss = loss + loss_rpn + loss_mos + (loss_motion if self.use_motion_loss else 0.0)
        train_loss_dict.append({"loss_mos": loss_mos.detach(), "loss_motion_encoder": loss_motion.detach(), **tb})
        gt_list.append(gt)
        pred_list.append(point_seg)
    return loss / len(list_batch_dict), train_loss_dict, gt_list, pred_list


def _forward_eval(self, list_batch_dict):
    """models/models.py:301-306,349-359,370-373: the inference forward plus the validation losses (MOS loss of both heads) and
    the per-sample recall record.  Returns (preb_dict_list, recall_dict_list, gt_list, pred_list, val_loss float,
    val_motion_loss tensor [1]) like the reference."""
    dev = list_batch_dict[0]["past_point_clouds"].device
    val_loss = torch.zeros(1, device=dev)
    val_motion_loss = torch.zeros(1, device=dev)
    boxes_out, recall_out, gt_list, pred_list = [], [], [], []
    for batch_dict in list_batch_dict:
        batch_dict = self.motion_encoder(batch_dict)
        if not self.use_motion_loss:
            batch_dict["current_motion_feature"] = batch_dict["current_motion_feature"][:, :3]
        gt = batch_dict["past_labels"][-1]
        loss_motion = self.MOSLoss.compute_loss(batch_dict["current_motion_feature"].clone(), gt)
        batch_dict = self.voxel_generate(batch_dict)
        batch_dict = self.vfe(batch_dict)
        batch_dict["_want_recall"] = True
        point_seg, pred_dicts, recall_dicts = self.unet(batch_dict, "eval")
        val_loss = val_loss + self.MOSLoss.compute_loss(point_seg, gt)
        val_motion_loss = val_motion_loss + loss_motion.detach()
        boxes_out.append(pred_dicts)
        recall_out.append(recall_dicts)
        gt_list.append(gt)
        pred_list.append(point_seg)
    n = len(list_batch_dict)
    return boxes_out, recall_out, gt_list, pred_list, (val_loss / n).item(), val_motion_loss / n


InsMOS_Model.forward_train = _forward_train
InsMOS_Model.forward_eval = _forward_eval


class ClassificationMetrics(nn.Module):
    """models/metrics.py:16-44 (the MOS-IoU acceptance metric is getIoU()[2])."""

    def __init__(self, n_classes, ignore_index):
        super().__init__()
        self.n_classes, self.ignore_index = n_classes, ignore_index

    def compute_confusion_matrix(self, pred_logits, gt_labels):
        pred_logits = pred_logits.clone()
        pred_logits[:, self.ignore_index] = -float("inf")
        pred = torch.argmax(torch.softmax(pred_logits, dim=1), dim=1).long()
        gt = gt_labels.long()
        cm = torch.zeros(self.n_classes, self.n_classes, dtype=gt.dtype, device=gt.device)
        return cm.index_put_((pred, gt), torch.ones_like(gt), accumulate=True)

    def getStats(self, cm):
        cm = cm.clone()
        cm[:, torch.tensor(self.ignore_index, dtype=torch.long)] = 0
        tp = cm.diag()
        return tp, cm.sum(dim=1) - tp, cm.sum(dim=0) - tp

    def getIoU(self, cm):
        tp, fp, fn = self.getStats(cm)
        return tp / (tp + fp + fn + 1e-15)


class InsMOSNet(LightningModule):
    def __init__(self, hparams: dict):
        super().__init__()
        self.save_hyperparameters(hparams)
        self.cfg = hparams
        self.id = hparams["EXPERIMENT"]["ID"]
        self.dt_prediction = hparams["MODEL"]["DELTA_T_PREDICTION"]
        self.n_past_steps = hparams["MODEL"]["N_PAST_STEPS"]
        tr = hparams.get("TRAIN", {})
        self.lr, self.lr_epoch, self.lr_decay = tr.get("LR"), tr.get("LR_EPOCH"), tr.get("LR_DECAY")
        self.weight_decay, self.batch_size = tr.get("WEIGHT_DECAY"), tr.get("BATCH_SIZE")
        try:
            with open(hparams["DATA"]["SEMANTIC_CONFIG_FILE"]) as f:
                self.semantic_config = yaml.safe_load(f)
        except (OSError, KeyError):
            self.semantic_config = _config.SEMANTIC_DEFAULT
        self.n_mos_classes = len(self.semantic_config["learning_map_inv"])
        self.ignore_index = [k for k, ign in self.semantic_config["learning_ignore"].items() if ign]
        self.model = InsMOS_Model(hparams, self.n_mos_classes, self.ignore_index)
        self.ClassificationMetrics = ClassificationMetrics(self.n_mos_classes, self.ignore_index)

    def forward(self, batch_data, Model_mode):
        return self.model(batch_data, Model_mode)
