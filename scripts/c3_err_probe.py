"""Where does the value error of a bench workload come from?  Strict-tolerance ratio per engine, and with the
RFF part / the exact-kernel part of the paths zeroed (experiment aid).   python scripts/c3_err_probe.py c3"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from oracle import psbax_oracle as O
from psbax_b200 import SamplePathBatch

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
dev = torch.device("cuda:0")
paths, tau, _ = bench.draw_paths(wl, dev, "auto")
X = bench.candidates_for(wl, 0, 1, dev)
idx = torch.linspace(0, X.shape[0] - 1, 4096, device=dev).long()
Xs = X[idx]
f64 = lambda t: None if t is None else t.double().cpu()  # noqa: E731


def ratio(W, V, engine):
    b = SamplePathBatch(paths.omega, paths.bias, W, paths.kernel, paths.Xtrain, paths.inv_ls, V, paths.mean_const,
                        paths.y_mean, paths.y_std, device=dev, engine=engine)
    got = b.values(Xs).double().cpu()
    p = O.PathParams(omega=f64(paths.omega), bias=f64(paths.bias), W=f64(W), kernel=paths.kernel, Xtrain=f64(paths.Xtrain),
                     inv_ls=f64(paths.inv_ls), V=f64(V), mean_const=float(np.float32(paths.mean_const)),
                     y_mean=float(np.float32(paths.y_mean)), y_std=float(np.float32(paths.y_std)))
    ref = O.eval_paths(p, Xs.double().cpu())
    pfull = O.PathParams(omega=f64(paths.omega), bias=f64(paths.bias), W=f64(paths.W), kernel=paths.kernel,
                         Xtrain=f64(paths.Xtrain), inv_ls=f64(paths.inv_ls), V=f64(paths.V),
                         mean_const=p.mean_const, y_mean=p.y_mean, y_std=p.y_std)
    tol = 1e-4 * O.path_scale(pfull)  # always the tolerance of the full path
    return float(((got - ref).abs().max(0).values / tol).max())


print(f"{wl}: |W| max row norm {paths.W.norm(dim=1).max().item():.2f}, |V| max row norm {paths.V.norm(dim=1).max().item():.2f}, "
      f"max |omega.x| ~ {(Xs @ paths.omega[0].T).abs().max().item():.1f} rad")
for eng in ("auto", "tensor_bf16", "tensor", "simt"):
    try:
        print(f"{eng:12s} full {ratio(paths.W, paths.V, eng):.3f}   RFF only {ratio(paths.W, torch.zeros_like(paths.V), eng):.3f}   "
              f"exact-kernel only {ratio(torch.zeros_like(paths.W), paths.V, eng):.3f}", flush=True)
    except Exception as ex:
        print(eng, "n/a", str(ex)[:80])
