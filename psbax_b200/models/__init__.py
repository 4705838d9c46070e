from .gp import ModelListGP, SingleTaskGP, Standardize  # noqa: F401
from .deep_kernel_gp import DKGP, FFNN  # noqa: F401
