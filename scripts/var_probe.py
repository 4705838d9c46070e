"""Fused posterior variance (psbax_eval_sumsq over alpha L^-1 pseudo-paths) against float64 torch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from psbax_b200.models import SingleTaskGP
from psbax_b200.sample_paths import posterior_variance

dev = torch.device("cuda:0")
for (n, d, noise, N) in [(64, 2, 1e-4, 200_000), (512, 5, 1e-4, 50_000), (122, 10, 1e-3, 50_000), (38, 3, 1e-4, 100)]:
    g = torch.Generator().manual_seed(n)
    Xt = torch.rand(n, d, generator=g, dtype=torch.float64)
    y = torch.sin(6 * Xt.sum(1)) + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)
    model = SingleTaskGP(Xt, y[:, None], kernel="matern52", lengthscale=torch.full((1, d), 0.3), outputscale=1.3, noise=noise)
    Xq = torch.rand(N, d, generator=g, dtype=torch.float64)
    ref = torch.cat([model.posterior(c.unsqueeze(1)).variance.reshape(-1) for c in Xq.split(2000)])
    for eng in ["auto", "tensor", "simt"]:
        try:
            var = posterior_variance(model, Xq.float().to(dev), engine=eng).double().cpu()
        except Exception as ex:
            print(n, d, eng, "n/a", str(ex)[:60]); continue
        err = (var - ref).abs()
        print(f"n={n} d={d} noise={noise:g} {eng:7s}: var range [{ref.min():.3e}, {ref.max():.3e}]  max abs err {err.max():.3e}  "
              f"argmax same {int(var.argmax()) == int(ref.argmax())}  top-var rel err {((var.max() - ref.max()) / ref.max()).abs():.2e}", flush=True)
    pend = Xq[:3]
    refp = torch.cat([model.posterior(torch.cat([c.unsqueeze(1), pend.expand(c.shape[0], -1, -1)], 1)).covariance_matrix.logdet() for c in Xq[3:].split(2000)])
    covP = model.posterior(pend.unsqueeze(0)).covariance_matrix[0].logdet()
    var = posterior_variance(model, Xq[3:].float().to(dev), pending=pend).double().cpu().clamp_min(1e-300)
    got = var.log() + covP
    print(f"   with 3 pending points: max |logdet diff| on the top decile {((got - refp).abs()[refp > refp.quantile(0.9)]).max():.3e}; argmax same {int(got.argmax()) == int(refp.argmax())}")

# timing at the C2 shape: n = 64, 1.68e7 candidates
n, d = 64, 2
g = torch.Generator().manual_seed(0)
Xt = torch.rand(n, d, generator=g, dtype=torch.float64)
y = torch.sin(6 * Xt.sum(1))
model = SingleTaskGP(Xt, y[:, None], kernel="matern52", lengthscale=torch.full((1, d), 0.3), outputscale=1.0, noise=1e-4)
Xq = torch.rand(16_777_216, d, generator=g).to(dev)
for _ in range(2):
    posterior_variance(model, Xq)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    v = posterior_variance(model, Xq)
e1.record()
torch.cuda.synchronize()
print(f"posterior variance of 1.68e7 candidates (n = 64, d = 2), incl. Cholesky + pack: {e0.elapsed_time(e1) / 3:.2f} ms")
