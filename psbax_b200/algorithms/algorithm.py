"""Base class of the BAX algorithms (interface of ``/root/reference/src/algorithms/algorithm.py:8-96``).

An algorithm sees a sample path either as an opaque callable ``f(X) -> Tensor``
(the reference protocol, kept for true objectives and foreign callables) or as a
``SamplePath`` / ``SamplePathBatch``, in which case evaluation and reduction run
fused on the GPU.
"""
import copy
from argparse import Namespace


def dict_to_namespace(params):
    return Namespace(**params) if isinstance(params, dict) else params


class Algorithm:
    def __init__(self, params=None, verbose=False):
        self.verbose_init_arg = verbose
        self.set_params(params)

    def set_params(self, params):
        ns = dict_to_namespace(params)
        self.params = Namespace()
        self.params.name = getattr(ns, "name", "Algorithm")

    def initialize(self):
        self.exe_path = Namespace(x=[], y=[])

    def get_copy(self):
        """Deep copy, as the reference does per acquisition (``src/one_trial.py:388``)."""
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        # device-side caches of X* are shared, not duplicated (the candidate set is read-only)
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for key, val in self.__dict__.items():
            setattr(new, key, val if key.startswith("_dev_") else copy.deepcopy(val, memo))
        return new

    def run_algorithm_on_f(self, f):
        raise NotImplementedError

    def run_algorithm_on_batch(self, paths):
        """All S paths of a ``SamplePathBatch`` at once -> list of (exe_path, output)."""
        return [self.get_copy().run_algorithm_on_f(p) for p in paths]

    def execute(self, f):
        _, output = self.run_algorithm_on_f(f)
        return output

    def execute_batch(self, paths):
        return [out for _, out in self.run_algorithm_on_batch(paths)]


def eval_in_batch(f, x_set, max_batch_size=100):
    """Chunked evaluation of a foreign callable (``src/algorithms/topk.py:48-49``)."""
    import torch
    return torch.cat([f(chunk) for chunk in x_set.split(max_batch_size)])
