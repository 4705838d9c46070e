"""``fit_model`` with the reference's signature and error behaviour
(``/root/reference/src/fit_model.py:20-91``) on the botorch-free model stand-ins.  Model fitting
is outside the accelerated path; this exists so that scripts written against the reference's
boundary (fit -> sample paths -> algorithm) run unchanged."""
import os

import torch

from .models import DKGP, ModelListGP, SingleTaskGP


def fit_model(inputs, outputs, model_type: str, **kwargs):
    if outputs.dim() == 1:
        outputs = outputs.reshape(-1, 1)
    try:
        if model_type == "gp":
            kernel_type = kwargs.pop("kernel_type", None)
            # None -> botorch default Matern-5/2 ARD; "rbf"/"matern" -> plain non-ARD kernels (:29-36)
            kernel = {"rbf": "rbf", "matern": "matern52", None: "matern52"}[kernel_type]
            if outputs.shape[-1] != 1:  # one model per output, each fitted, wrapped in a ModelListGP (:38-62)
                return ModelListGP(*[SingleTaskGP(inputs, outputs[:, j:j + 1], kernel=kernel, ard=kernel_type is None).fit()
                                     for j in range(outputs.shape[-1])])
            model = SingleTaskGP(inputs, outputs, kernel=kernel, ard=kernel_type is None)
            state = kwargs.pop("state_dict", None)
            if state is not None:  # fixed model without training (:50-54)
                model.covar_module.base_kernel.lengthscale = torch.as_tensor(state["lengthscale"], dtype=torch.float64).reshape(1, -1)
                model.covar_module.outputscale = torch.as_tensor(state["outputscale"], dtype=torch.float64)
                model.likelihood.noise = torch.as_tensor(state["noise"], dtype=torch.float64).reshape(1)
                model.mean_module.constant = torch.as_tensor(state.get("mean_const", 0.0), dtype=torch.float64)
                return model
            return model.fit()
        if model_type == "dkgp":
            architecture = kwargs.pop("architecture")
            if architecture is None:
                architecture = [32, 32, inputs.shape[-1]]
            architecture = [inputs.shape[-1]] + list(architecture) + [outputs.shape[-1]]
            std_outputs = (outputs - outputs.mean()) / outputs.std()
            model = DKGP(train_X=inputs, train_Y=std_outputs, architecture=architecture)
            model.train_model(X=inputs, Y=std_outputs.squeeze(-1), lr=1e-2, num_iter=kwargs.pop("epochs", 10000))
            return model
        raise ValueError(f"unknown model_type {model_type!r}")
    except Exception:  # the reference swallows everything, dumps the data and returns None (:83-91)
        file_path = kwargs.pop("file_path", None)
        if file_path is not None:
            print("Error in fitting model, saving inputs and outputs to files")
            os.makedirs(os.path.dirname(file_path) or ".", exist_ok=True)
            torch.save(inputs, f"{file_path}_inputs.pt")
            torch.save(outputs, f"{file_path}_outputs.pt")
        return None
