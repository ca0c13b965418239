"""Detection post-processing (a10) and instance-feature fusion helpers (a11) on device.

Mirrors models/post_process.py:5-24,112-224, models/bbox_post_process/iou3d_nms_utils.py:64-79 and the
call sites of Array_Index.find_features_by_bbox_with_yaw in models/backbones_3d/spconv_unet.py:319-401.
Selection order follows the reference: score >= thresh, top-k (<= NMS_PRE_MAXSIZE) by score, sort
descending, rotated NMS (device-side sweep), first NMS_POST_MAXSIZE.
"""
import torch

from insmos_b200 import ops


def class_agnostic_nms(scores, boxes, nms_config, score_thresh):
    """returns indices (int64, into the unfiltered arrays) of the selected boxes, in NMS order."""
    cand = torch.nonzero(scores >= score_thresh).view(-1)              # one host sync (data-dependent size)
    if cand.numel() == 0:
        return cand
    s = scores[cand]
    k = min(int(nms_config["NMS_PRE_MAXSIZE"]), s.shape[0])
    top_s, top_i = torch.topk(s, k=k)
    order = top_s.sort(0, descending=True)[1]                          # iou3d_nms_utils.py:72 sorts again
    b = boxes[cand[top_i[order]]][:, 0:7].contiguous()
    keep = ops.nms_rotated(b, float(nms_config["NMS_THRESH"]), int(nms_config["NMS_POST_MAXSIZE"]))
    return cand[top_i[order[keep.long()]]]


def generate_recall_record(box_preds, recall_dict, batch_index, data_dict=None, thresh_list=None):
    """post_process.py:66-109: per IoU threshold, how many ground-truth boxes are matched by a prediction with IoU3D above
    it.  Same dictionary ('gt', 'roi_<t>', 'rcnn_<t>'); the pairwise IoU runs in insmos_boxes_iou3d and the per-threshold
    counts come back in ONE read (the reference reads one scalar per threshold)."""
    if "gt_boxes" not in data_dict:
        return recall_dict
    gt_boxes = data_dict["gt_boxes"][batch_index]
    if len(recall_dict) == 0:
        recall_dict = {"gt": 0}
        for t in thresh_list:
            recall_dict["roi_%s" % str(t)] = 0
            recall_dict["rcnn_%s" % str(t)] = 0
    k = len(gt_boxes) - 1
    nonzero = (gt_boxes.sum(dim=1) != 0).tolist() if len(gt_boxes) else []
    while k > 0 and not nonzero[k]:                          # trailing all-zero padding rows (post_process.py:82-85)
        k -= 1
    cur_gt = gt_boxes[:k + 1]
    if cur_gt.shape[0] > 0:
        if box_preds.shape[0] > 0:
            iou = ops.boxes_iou3d(box_preds[:, 0:7].contiguous(), cur_gt[:, 0:7].float().contiguous())
            best = iou.max(dim=0)[0]
            counts = torch.stack([(best > t).sum() for t in thresh_list]).tolist()
        else:
            counts = [0] * len(thresh_list)
        for t, c in zip(thresh_list, counts):
            recall_dict["rcnn_%s" % str(t)] += int(c)
        recall_dict["gt"] += int(cur_gt.shape[0])
    return recall_dict


def post_processing(batch_dict, cfg, num_class):
    """post_processing for batch 1 -> ([{'pred_boxes','pred_scores','pred_labels'}], recall dict); the recall record is
    filled whenever the sample carries 'gt_boxes', as in the reference (post_process.py:206-216)."""
    boxes, scores, labels = batch_dict["_decoded"]
    if cfg["NMS_CONFIG"]["MULTI_CLASSES_NMS"]:
        raise NotImplementedError("MULTI_CLASSES_NMS is disabled in the reference config (config.yaml:152)")
    sel = class_agnostic_nms(scores, boxes, cfg["NMS_CONFIG"], cfg["SCORE_THRESH"])
    final_scores = scores[sel]
    if cfg.get("OUTPUT_RAW_SCORE", False):
        final_scores = batch_dict["batch_cls_preds"][0].max(dim=-1)[0][sel]
    rec = {"pred_boxes": boxes[sel], "pred_scores": final_scores, "pred_labels": labels[sel].long()}
    recall = {}
    if "gt_boxes" in batch_dict and batch_dict.get("_want_recall", False):
        recall = generate_recall_record(rec["pred_boxes"], recall, 0, data_dict=batch_dict, thresh_list=cfg["RECALL_THRESH_LIST"])
    return [rec], recall


class InstanceBoxes:
    """the NMS survivors in voxel units of the stride-8 level; `bits(indices, level_mult)` is the
    device-side Array_Index.find_features_by_bbox_with_yaw for one resolution level."""

    def __init__(self, pred, range_min, voxel_size, stride, num_class):
        self.num_class = num_class
        b7 = pred["pred_boxes"].contiguous()
        self.nb = b7.shape[0]
        self.boxes8 = ops.boxes_to_voxel_units(b7, pred["pred_labels"].to(torch.int32).contiguous(), range_min, voxel_size,
                                               float(stride)) if self.nb else torch.zeros((0, 8), device=b7.device)

    PAD = 8                                            # the bits occupy 8 columns (3 used): C + 8 stays a multiple of 8

    def concat_bits(self, features, indices, mult):
        """cat([features, one-hot class membership, zero padding], dim=1) for the voxels `indices` [n,4] (b,z,y,x).
        The reference concatenates exactly num_class columns (spconv_unet.py:345-349); the fused graph appends 8 so that
        the next convolution's channel count stays a multiple of 8 (its weight gets zero rows, see
        SparseConvolution.kernel_major_weight) -- odd counts fall off the vectorised tensor-core paths."""
        n, c = features.shape
        bits = torch.zeros((n, self.PAD), dtype=torch.float32, device=features.device)
        ops.box_membership(indices, self.boxes8, float(mult), out=bits, out_stride=self.PAD, col_offset=0, n_class=self.num_class)
        if torch.is_grad_enabled() and features.requires_grad:
            return torch.cat([features, bits], 1), bits
        return ops.concat2(features, bits), bits
