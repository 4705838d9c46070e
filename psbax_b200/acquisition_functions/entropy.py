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

    def q1_values(self, choices: torch.Tensor) -> torch.Tensor:
        """``forward(choices.unsqueeze(-2))`` for every row of ``choices`` [m, d] in ONE fused GPU
        launch (row f1 of SURVEY.md section 8): H = 0.5 log(2 pi) + 0.5 logdet Cov([x, P]) with
        logdet Cov([x, P]) = logdet Cov(P) + log var(x | P); the conditional variance comes from
        ``psbax_eval_sumsq`` over pseudo-paths built from the Cholesky factor of the training
        (+ pending) covariance.  Returns [m] float64 on the choices' device."""
        from ..sample_paths import posterior_variance

        pend = self.X_pending
        # alpha - sumsq is a float32 difference: floor it at its resolution (1e-6 of the prior variance)
        # instead of letting rounding noise below zero turn into log(1e-300)
        var = posterior_variance(self.model, choices, pending=pend).double().clamp_min(variance_floor(self.model))
        val = 0.5 * math.log(2 * math.pi) + 0.5 * torch.log(var)
        if pend is not None and pend.numel() > 0:
            covP = self.model.posterior(pend.reshape(1, -1, choices.shape[-1])).covariance_matrix[0]
            val = val + 0.5 * torch.logdet(covP).to(val.device)
        return val.to(torch.as_tensor(choices).device)


def variance_floor(model) -> float:
    """Smallest posterior variance (original units) the fused float32 path reports: 1e-6 of the prior
    variance.  The true variance is >= ~noise / 2 at observed rows and ~jitter at pending rows; the
    float32 difference alpha - |L^-1 k|^2 resolves neither below ~1e-6 alpha, so smaller (or negative)
    results are rounding noise and all tie at the floor."""
    if hasattr(model, "y_mean_std"):
        alpha, y_std = float(model.covar_module.outputscale), float(model.y_mean_std()[1])
        return 1e-6 * alpha * y_std * y_std
    return 1e-12


def _fusable(acq_function) -> bool:
    """The fused path needs the exact-GP stand-in (kernel_matrix / Cholesky access) and a GPU."""
    model = getattr(acq_function, "model", None)
    exact_gp = hasattr(model, "kernel_matrix") and hasattr(model, "y_mean_std")
    deep_kernel = getattr(model, "dkl", False) and hasattr(model, "weight_posterior")
    return hasattr(acq_function, "q1_values") and (exact_gp or deep_kernel) and torch.cuda.is_available()


def optimize_acqf_discrete(acq_function, q, choices, max_batch_size=100, unique=True, fused=None):
    """Greedy arg max over ``choices`` [m, d]: evaluate, take the best row, add it to
    ``X_pending``, drop it from the choices, repeat q times.  ``fused=None`` uses one fused GPU
    launch per greedy step (``EntropyAcquisitionFunction.q1_values``) when the model supports it,
    else the reference's chunks of ``max_batch_size`` through ``acq_function.forward``;
    ``fused=True`` insists on the GPU path (and raises if it is not available)."""
    choices = torch.as_tensor(choices)
    base_pending = acq_function.X_pending
    remaining = choices
    picked, values = [], []
    use_fused = _fusable(acq_function) if fused is None else bool(fused)
    if use_fused and not _fusable(acq_function):
        raise RuntimeError("fused max-entropy pick needs the psbax_b200 exact-GP model and a CUDA device")
    for _ in range(q):
        with torch.no_grad():
            if use_fused:
                vals = acq_function.q1_values(remaining)
            else:
                vals = torch.cat([acq_function(c.unsqueeze(-2)) for c in remaining.split(max_batch_size)])
        j = int(torch.argmax(vals))
        picked.append(remaining[j])
        values.append(vals[j])
        pend = torch.stack(picked)
        acq_function.set_X_pending(pend if base_pending is None else torch.cat([base_pending, pend]))
        if unique:
            if use_fused:  # every copy of the picked row goes: its conditional variance is now ~0 (float32 noise)
                remaining = remaining[~(remaining == remaining[j]).all(dim=-1)]
            else:
                remaining = torch.cat([remaining[:j], remaining[j + 1:]])
        if remaining.shape[0] == 0:
            break
    acq_function.set_X_pending(base_pending)
    return torch.stack(picked), torch.stack(values)
