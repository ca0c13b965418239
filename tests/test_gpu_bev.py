"""Dense BEV tensor-core kernels (3xTF32 implicit GEMM, channels-last) vs torch fp32 on the CPU (MKL-DNN)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from insmos_b200 import ops
from oracle import sp

pytestmark = pytest.mark.gpu

# fp32 parity of a K = 9*256 = 2304 term dot product evaluated in a different order: 1e-5 relative to the
# magnitude of the output (outputs are O(1)); plain TF32 would be ~1e-3.
TOL = 2e-5


def _nhwc(x):                      # [1,C,H,W] -> [H*W, C]
    return x[0].permute(1, 2, 0).reshape(-1, x.shape[1]).contiguous()


@pytest.mark.parametrize("impl", ["tcgen05", "umma"])
@pytest.mark.parametrize("Cin,Cout,H,W", [(256, 128, 125, 150), (128, 128, 125, 150), (32, 128, 7, 9)])
def test_conv3x3_matches_torch(cuda, Cin, Cout, H, W, impl):
    g = torch.Generator().manual_seed(Cin + H)
    x = torch.randn((1, Cin, H, W), generator=g)
    w = torch.randn((Cout, Cin, 3, 3), generator=g) / np.sqrt(9 * Cin)
    b = torch.randn(Cout, generator=g)
    ref = torch.relu(F.conv2d(x, w, b, padding=1))
    wk = w.permute(2, 3, 1, 0).reshape(9, Cin, Cout).contiguous()
    out = ops.conv2d_nhwc(_nhwc(x).to(cuda), H, W, wk.to(cuda), 0, bias=b.to(cuda), relu=True, impl=impl).cpu()
    err = (out - _nhwc(ref)).abs().max().item()
    assert err < TOL * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("impl", ["tcgen05"])
def test_deconv2x2_and_1x1_match_torch(cuda, impl):
    g = torch.Generator().manual_seed(3)
    H, W, Cin, Cout = 125, 150, 128, 256
    x = torch.randn((1, Cin, H, W), generator=g)
    w = torch.randn((Cin, Cout, 2, 2), generator=g) / np.sqrt(Cin)
    b = torch.randn(Cout, generator=g)
    ref = torch.relu(F.conv_transpose2d(x, w, b, stride=2))                     # [1,Cout,2H,2W]
    wk = w.permute(2, 3, 0, 1).reshape(4, Cin, Cout).contiguous()
    out = ops.conv2d_nhwc(_nhwc(x).to(cuda), H, W, wk.to(cuda), 2, bias=b.to(cuda), relu=True, impl=impl).cpu()
    assert out.shape == (4 * H * W, Cout)
    assert (out - _nhwc(ref)).abs().max().item() < TOL * max(1.0, ref.abs().max().item())
    w1 = torch.randn((128, Cin, 1, 1), generator=g) / np.sqrt(Cin)
    ref1 = F.conv2d(x, w1)
    out1 = ops.conv2d_nhwc(_nhwc(x).to(cuda), H, W, w1.flatten(1).t().reshape(1, Cin, 128).contiguous().to(cuda), 1, impl=impl).cpu()
    assert (out1 - _nhwc(ref1)).abs().max().item() < TOL * max(1.0, ref1.abs().max().item())


def test_dense_scatter_nhwc_matches_dense_view(cuda):
    g = torch.Generator().manual_seed(6)
    zyx = torch.stack([torch.randint(0, 2, (3000,), generator=g), torch.randint(0, 125, (3000,), generator=g),
                       torch.randint(0, 150, (3000,), generator=g)], 1)
    zyx = torch.unique(zyx, dim=0)
    ind = torch.cat([torch.zeros((len(zyx), 1), dtype=torch.long), zyx], 1).to(torch.int32)
    f = torch.randn((len(ind), 128), generator=g)
    got = ops.dense_scatter_nhwc(f.to(cuda), ind.to(cuda), 2, 125, 150).cpu()
    dense = sp.dense(f, ind.numpy(), [2, 125, 150])                              # [1,C,D,H,W]
    ref = dense.view(1, 256, 125, 150)                                           # height_compression.py:29-30
    assert torch.equal(got, _nhwc(ref))


def test_decode_channels_last_equals_nchw(cuda):
    g = torch.Generator().manual_seed(1)
    H, W = 50, 60
    head = torch.randn((H * W, 11), generator=g).to(cuda)
    a = ops.center_decode(head[:, :3], head[:, 3:], 4, 0.1, 0.1, -60, -50, hw=(H, W))
    cls = head[:, :3].t().reshape(3, H, W).contiguous()
    box = head[:, 3:].t().reshape(8, H, W).contiguous()
    b = ops.center_decode(cls, box, 4, 0.1, 0.1, -60, -50)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
