from .entropy import EntropyAcquisitionFunction  # noqa: F401
from .posterior_sampling import gen_posterior_sampling_batch  # noqa: F401
from .bax_acquisition import BAXAcquisitionFunction  # noqa: F401
