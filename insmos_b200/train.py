"""Lightning-free training step for InsMOSNet (BASELINE config 5).

Mirrors, for the step itself (paths relative to the reference repository):
  models/models.py:61-98      training_step: forward(batch, 'train'), mean of the per-sample loss terms, confusion matrix
  models/models.py:185-190    Adam(lr, weight_decay) + StepLR(step_size=LR_EPOCH, gamma=LR_DECAY)
  scripts/train.py:74-79      Trainer(gpus=1, strategy="ddp"): replicated weights, gradients averaged over the ranks
What is B200-native about it:
  * forward and backward run on this repository's kernels (insmos_b200/autograd.py: sparse-conv dgrad through the forward
    tensor-core kernels over the transposed rule books, wgrad / BatchNorm / target-assignment / scatter kernels of
    csrc/train.cu); only the dense 2D BEV convolutions and the tiny loss arithmetic go through torch (cuDNN / elementwise);
  * every parameter lives in ONE flat fp32 buffer and every gradient in a second one (the tensors of the module are views):
    the data-parallel exchange is a single in-place NCCL all-reduce of 25.8 MB over NVLink/NVSwitch per step ("ZeRO-0":
    nothing is sharded) and the optimizer is one fused Adam kernel over the flat buffers (insmos_adam_step) instead of
    ~200 per-tensor launches;
  * the loss values stay on the device; one small read-back per step returns all of them.
BatchNorm statistics are per rank, as in the reference (plain BatchNorm, not SyncBatchNorm: spconv_unet.py:118, minkunet.py:52).
"""
import torch
import torch.distributed as dist

from . import ops


class FlatParameters:
    """re-homes the trainable parameters of `module` into one contiguous fp32 buffer (and their .grad into another)."""

    def __init__(self, module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        if any(p.dtype != torch.float32 or p.device != dev for p in self.params):
            raise RuntimeError("FlatParameters: float32 parameters on one device required")
        # (the buffers themselves are device-agnostic host logic -- the gloo tests exercise the exchange on CPU tensors;
        # the optimizer kernel and the model are CUDA-only and refuse anything else)
        n = sum(p.numel() for p in self.params)
        self.data = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        o = 0
        for p in self.params:
            k = p.numel()
            self.data[o:o + k].copy_(p.data.reshape(-1))
            p.data = self.data[o:o + k].view(p.shape)
            p.grad = self.grad[o:o + k].view(p.shape)
            o += k
        self.numel = n

    def zero_grad(self):
        self.grad.zero_()
        o = 0
        for p in self.params:                       # a backward that replaced .grad (e.g. after set_to_none) is re-attached
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + k].view(p.shape)
            o += k


class TrainStep:
    """model: insmos_b200.net.model.InsMOSNet (or its .model) on a CUDA device, in train() mode."""

    def __init__(self, model, lr=1e-4, weight_decay=1e-4, lr_epoch=1, lr_decay=0.99, betas=(0.9, 0.999), eps=1e-8,
                 process_group=None):
        self.net = model
        self.flat = FlatParameters(model)
        self.exp_avg = torch.zeros_like(self.flat.data)
        self.exp_avg_sq = torch.zeros_like(self.flat.data)
        self.base_lr, self.weight_decay, self.lr_epoch, self.lr_decay = lr, weight_decay, lr_epoch, lr_decay
        self.betas, self.eps = betas, eps
        self.steps, self.epoch = 0, 0
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1

    @classmethod
    def from_config(cls, model, cfg, process_group=None):
        t = cfg["TRAIN"]
        return cls(model, lr=t["LR"], weight_decay=t["WEIGHT_DECAY"], lr_epoch=t["LR_EPOCH"], lr_decay=t["LR_DECAY"],
                   process_group=process_group)

    @property
    def lr(self):                                    # StepLR: lr * gamma ** (epoch // step_size)
        return self.base_lr * self.lr_decay ** (self.epoch // max(int(self.lr_epoch), 1))

    def set_epoch(self, epoch):
        self.epoch = int(epoch)

    def forward_backward(self, batch):
        """forward(batch, 'train') + backward into the flat gradient buffer.  Returns (loss tensor [1], loss dicts, gts, preds)."""
        self.flat.zero_grad()
        model = self.net
        out = model(batch, "train") if not hasattr(model, "forward_train") else model.forward_train(batch)
        loss = out[0]
        loss.backward()
        return out

    def all_reduce_gradients(self):
        """ONE collective per step: in-place sum of the flat gradient buffer over the ranks (the average is folded into the
        optimizer kernel as grad_scale = 1 / world)."""
        if self.world > 1:
            dist.all_reduce(self.flat.grad, op=dist.ReduceOp.SUM, group=self.group)

    def optimizer_step(self):
        self.steps += 1
        ops.adam_step(self.flat.data, self.flat.grad, self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0], self.betas[1],
                      self.eps, self.weight_decay, self.steps, grad_scale=1.0 / self.world)
        # the kernel wrote the parameters behind autograd's back: bump their version counters so that every cache keyed on
        # (data_ptr, _version) -- fragment-ordered weight images, folded BatchNorm constants -- sees the change
        for p in self.flat.params:
            torch.autograd.graph.increment_version(p)

    def step(self, batch, want_confusion=False):
        """one training step on this rank's samples.  Returns a dict of python floats: loss, cls_loss, box_loss, mos_loss,
        motion_loss (means over the samples, models.py:69-82) [+ 'confusion_matrix' tensor]."""
        loss, dicts, gts, preds = self.forward_backward(batch)
        self.all_reduce_gradients()
        self.optimizer_step()
        n = len(dicts)
        vals = torch.stack([loss.detach().reshape(()),
                            sum(d["rpn_loss_cls"] for d in dicts) / n, sum(d["rpn_loss_loc"] for d in dicts) / n,
                            sum(d["loss_mos"] for d in dicts) / n, sum(d["loss_motion_encoder"] for d in dicts) / n])
        v = vals.tolist()                                                   # the step's one read-back
        res = {"loss": v[0], "cls_loss": v[1], "box_loss": v[2], "mos_loss": v[3], "motion_loss": v[4]}
        if want_confusion:
            cm = getattr(self.net, "ClassificationMetrics", None)
            if cm is not None:
                res["confusion_matrix"] = cm.compute_confusion_matrix(torch.cat(preds, 0).detach(), torch.cat(gts, 0))
        return res
