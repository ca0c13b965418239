"""cProfile of the host side of the forward (which Python frames the wall clock goes to); run on the GPU box."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
clouds = [torch.from_numpy(c).to(dev) for c in bench.make_clouds(0, 4)]
net = bench.build_model(dev, clouds[0])
with torch.no_grad():
    for i in range(6):
        bench.step(net, clouds[i % 4])
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    pr = cProfile.Profile()
    pr.enable()
    for i in range(20):
        bench.step(net, clouds[i % 4])
    torch.cuda.synchronize()
    pr.disable()
    print("wall per step under cProfile: %.2f ms" % ((time.perf_counter() - t0) * 1000 / 20))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(35)
print(s.getvalue()[:6000])
