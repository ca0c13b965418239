"""accuracy of the tcgen05 sparse conv vs an fp64 reference for 1/2/4 round-robin TMEM accumulators; run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from insmos_b200 import ops, synth
from oracle import me
dev = torch.device("cuda:0")
pts = synth.make_sequence(seed=7, n_scans=3, n_elev=32, n_azim=400)
cs, _, _ = ops.voxelize4d(torch.from_numpy(pts).to(dev), [0.1, 0.1, 0.1, 0.1])
c = cs.coords.cpu().numpy()
for ksize, K in (([3, 3, 3, 1], 27), ([3, 3, 3, 3], 81)):
    maps = me.kernel_map(c, c, ksize, [1, 1, 1, 1])
    rb = ops.build_rulebook(cs, cs, ops.spec_me_cube(ksize, [1, 1, 1, 1]))
    for Cin, Cout in ((64, 128), (128, 128), (48, 64)):
        g = torch.Generator().manual_seed(Cin * 1000 + Cout)
        feats = torch.randn((len(c), Cin), generator=g)
        W = torch.randn((K, Cin, Cout), generator=g) / np.sqrt(Cin * 10.0)
        ref = me.conv(feats.double(), W.double(), maps, len(c))
        line = "K=%d %d->%d |ref|max %.2f:" % (K, Cin, Cout, ref.abs().max())
        for algo, env in ((2, {}), (4, {"INSMOS_UMMA_NACC": "1"}), (4, {"INSMOS_UMMA_NACC": "2"}), (4, {"INSMOS_UMMA_NACC": "4"})):
            os.environ.update(env)
            out = ops.sparse_conv(feats.to(dev), W.to(dev), rb, algo=algo).cpu().double()
            err = (out - ref)
            line += "  algo%d%s max %.2e mean(signed*sgn(ref)) %.2e" % (algo, env.get("INSMOS_UMMA_NACC", ""), err.abs().max(), (err * torch.sign(ref)).mean())
        print(line)
