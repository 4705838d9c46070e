"""On-disk layout of a trial's results, as ``/root/reference/src/one_trial.py:344-372`` writes it and
``:69-113`` reads it back on restart (and ``demos/graph.py`` plots it):

    <results_folder>/inputs/inputs_<trial>.txt              [n, d] (np.savetxt of inputs.reshape(n, -1))
    <results_folder>/obj_vals/obj_vals_<trial>.txt          [n] or [n, m]
    <results_folder>/runtimes/runtimes_<trial>.txt          [iterations]
    <results_folder>/performance_metrics_<trial>.txt        [iterations + 1, n_metrics]

``results_folder`` is used as a string prefix, exactly as the reference does (it ends with "/").
"""
import os

import numpy as np
import torch


def save_trial(results_folder: str, trial: int, inputs, obj_vals, runtimes, performance_metrics_vals) -> None:
    for sub in ("", "inputs/", "obj_vals/", "runtimes/"):
        os.makedirs(results_folder + sub, exist_ok=True)
    inputs = torch.as_tensor(inputs).detach().cpu().numpy()
    np.savetxt(results_folder + "inputs/inputs_" + str(trial) + ".txt", inputs.reshape(inputs.shape[0], -1))
    np.savetxt(results_folder + "obj_vals/obj_vals_" + str(trial) + ".txt",
               torch.as_tensor(obj_vals).detach().cpu().numpy())
    np.savetxt(results_folder + "runtimes/runtimes_" + str(trial) + ".txt", np.atleast_1d(runtimes))
    np.savetxt(results_folder + "performance_metrics_" + str(trial) + ".txt", np.atleast_1d(performance_metrics_vals))


def load_trial(results_folder: str, trial: int):
    """-> (inputs [n, d] float64 tensor, obj_vals tensor, runtimes list, performance_metrics_vals list of
    1-D arrays, iteration) — the state ``one_trial`` resumes from (``one_trial.py:69-113``).  Raises
    ``OSError`` when the trial has not been saved (the reference then starts from initial data)."""
    inputs = np.atleast_2d(np.loadtxt(results_folder + "inputs/inputs_" + str(trial) + ".txt"))
    obj_vals = np.loadtxt(results_folder + "obj_vals/obj_vals_" + str(trial) + ".txt")
    runtimes = list(np.atleast_1d(np.loadtxt(results_folder + "runtimes/runtimes_" + str(trial) + ".txt")))
    arr = np.loadtxt(results_folder + "performance_metrics_" + str(trial) + ".txt")
    if arr.ndim == 1:
        vals = [np.array([arr[i]]) for i in range(arr.shape[0])]
    else:
        vals = [arr[i] for i in range(arr.shape[0])]
    if inputs.shape[0] != np.atleast_1d(obj_vals).shape[0]:  # a single d-dimensional row was read back as a column
        inputs = inputs.reshape(np.atleast_1d(obj_vals).shape[0], -1)
    return torch.from_numpy(inputs), torch.from_numpy(np.asarray(obj_vals)), runtimes, vals, len(runtimes)
