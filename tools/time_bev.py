"""time the two dense BEV conv implementations (mma.sync vs tcgen05) on the BEV shapes; run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from insmos_b200 import ops

dev = torch.device("cuda:0")
H, W = 125, 150
for Cin, Cout, mode in ((256, 128, 0), (128, 128, 0), (128, 256, 2)):
    taps = (9, 1, 4)[mode]
    x = torch.randn(H * W, Cin, device=dev)
    w = torch.randn(taps, Cin, Cout, device=dev) / (taps * Cin) ** 0.5
    b = torch.randn(Cout, device=dev)
    ref = None
    for impl in ("tcgen05", "umma"):
        for _ in range(3):
            out = ops.conv2d_nhwc(x, H, W, w, mode, bias=b, relu=True, impl=impl)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            out = ops.conv2d_nhwc(x, H, W, w, mode, bias=b, relu=True, impl=impl)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        gf = 2.0 * H * W * taps * Cin * Cout / 1e9
        if ref is None:
            ref = out
        print("Cin %3d Cout %3d mode %d  %-8s %8.1f us  %6.1f TFLOP/s (fp32-equivalent)  max|diff vs mma| %.2e"
              % (Cin, Cout, mode, impl, ms * 1000, gf / ms, (out - ref).abs().max().item()))
