import torch.nn as nn

from .core import SparseConvTensor


class SparseModule(nn.Module):
    """marker base class: modules that take and return a SparseConvTensor."""


def _is_sparse(m):
    return isinstance(m, SparseModule)


class SparseSequential(SparseModule):
    """Sequential that applies plain nn.Modules to `.features`.  In eval mode the pattern
    [SparseConvolution, BatchNorm1d, ReLU] is fused into the convolution's epilogue."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], dict):
            for k, m in args[0].items():
                self.add_module(k, m)
        else:
            for i, m in enumerate(args):
                self.add_module(str(i), m)
        for k, m in kwargs.items():
            self.add_module(k, m)

    def __getitem__(self, idx):
        return list(self._modules.values())[idx]

    def __len__(self):
        return len(self._modules)

    def add(self, module, name=None):
        self.add_module(str(len(self._modules)) if name is None else name, module)

    def forward(self, x):
        from .conv import SparseConvolution
        mods = list(self._modules.values())
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, SparseConvolution) and not self.training and i + 1 < len(mods) and \
                    isinstance(mods[i + 1], nn.BatchNorm1d) and mods[i + 1].track_running_stats:
                relu = i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
                x = m(x, bn=mods[i + 1], relu=relu)
                i += 3 if relu else 2
                continue
            if self.training and isinstance(m, nn.BatchNorm1d) and isinstance(x, SparseConvTensor) and x.features.shape[0] != 0:
                # batch statistics on the library's column-moment kernels, ReLU fused when it follows (training step, N3)
                from insmos_b200 import autograd as _ag
                relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
                x = x.replace_feature(_ag.batch_norm_train(m, x.features, relu=relu))
                i += 2 if relu else 1
                continue
            if _is_sparse(m):
                x = m(x)
            elif isinstance(x, SparseConvTensor):
                if x.features.shape[0] != 0:
                    x = x.replace_feature(m(x.features))
            else:
                x = m(x)
            i += 1
        return x
