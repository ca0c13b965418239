"""two warm forwards, then ONE forward of the C2 workload inside the NVTX range 'fwd' (for ncu --nvtx --nvtx-include fwd/)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
clouds = [torch.from_numpy(c).to(dev) for c in bench.make_clouds(0, 1)]
net = bench.build_model(dev, clouds[0])
with torch.no_grad():
    for i in range(2):
        bench.step(net, clouds[0])
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("fwd")
    bench.step(net, clouds[0])
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("done")
