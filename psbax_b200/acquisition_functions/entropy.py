"""Posterior-entropy acquisition (``/root/reference/src/acquisition_functions/entropy.py:14-41``)
and the greedy discrete optimiser PS-BAX pairs it with (botorch 0.10
``optimize_acqf_discrete``, call site ``posterior_sampling.py:78-81``)."""
import math

import torch


class EntropyAcquisitionFunction:
    """H[f(X)] = 0.5 log(2 pi) + 0.5 logdet Cov_post(X), X [b, q, d] -> [b].
    ``X_pending`` rows are appended to every q-batch (``@concatenate_pending_points``)."""

    def __init__(self, model, **kwargs):
        self.model = model
        self.X_pending = None

    def set_X_pending(self, X_pending=None):
        self.X_pending = X_pending

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        if X.dim() == 2:
            X = X.unsqueeze(0)
        if self.X_pending is not None:
            pend = self.X_pending.to(X).expand(X.shape[0], -1, -1)
            X = torch.cat([X, pend], dim=-2)
        cov = self.model.posterior(X).covariance_matrix
        return 0.5 * math.log(2 * math.pi) + 0.5 * torch.logdet(cov)

    __call__ = forward


def optimize_acqf_discrete(acq_function, q, choices, max_batch_size=100, unique=True):
    """Greedy arg max over ``choices`` [m, d]: evaluate in chunks of ``max_batch_size``,
    take the best row, add it to ``X_pending``, drop it from the choices, repeat q times."""
    choices = torch.as_tensor(choices)
    base_pending = acq_function.X_pending
    remaining = choices
    picked, values = [], []
    for _ in range(q):
        with torch.no_grad():
            vals = torch.cat([acq_function(c.unsqueeze(-2)) for c in remaining.split(max_batch_size)])
        j = int(torch.argmax(vals))
        picked.append(remaining[j])
        values.append(vals[j])
        pend = torch.stack(picked)
        acq_function.set_X_pending(pend if base_pending is None else torch.cat([base_pending, pend]))
        if unique:
            remaining = torch.cat([remaining[:j], remaining[j + 1:]])
        if remaining.shape[0] == 0:
            break
    acq_function.set_X_pending(base_pending)
    return torch.stack(picked), torch.stack(values)
