from .gp import SingleTaskGP, Standardize  # noqa: F401
