import sys, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from test_gpu_horizon import _pendulum, _run
ns = 2
for steps in (1, 2):
    params, eps, u = _pendulum(ns, steps, 11)
    fs, ts = _run(params, eps, u, False)
    ff, tf = _run(params, eps, u, True)
    print('steps', steps, 'traj diff', float((ts - tf).abs().max()))
    Xs, Ys = fs.engine.export_hallucinated(); Xf, Yf = ff.engine.export_hallucinated()
    print(' Y diff', float((Ys - Yf).abs().max()), 'X diff', float((Xs - Xf).abs().max()))
    g = torch.Generator().manual_seed(3)
    probe = (torch.rand(ns, 1, 2, 3, generator=g, dtype=torch.float64) - 0.5).expand(ns, 2, 2, 3).contiguous()
    pe = torch.randn(ns, 2, 2, 4, generator=g, dtype=torch.float64)
    for mma in (False, True):
        outs = []
        for fr in (ff, fs):
            fr.engine.set_block_kernels(mma)
            outs.append(fr.engine.posterior(probe, pe))
        print(' mma', mma, [float((a - b).abs().max()) for a, b in zip(*outs)])
    x1 = probe[:, :, :1].contiguous(); e1 = pe[:, :, :1].contiguous()
    print(' step', [float((a - b).abs().max()) for a, b in zip(ff.engine.step(x1, e1), fs.engine.step(x1, e1))])
