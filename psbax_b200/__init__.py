"""psbax_b200 — B200-native implementation of the PS-BAX sample-path hot path.

Drop-in for the hot-path part of the reference's Python API (``src/algorithms``,
``src/utils.get_function_samples``, ``src/acquisition_functions/posterior_sampling``)
on top of the CUDA library ``libpsbax_b200.so`` (C ABI: ``include/psbax_b200.h``).
"""
from ._lib import PsbaxError  # noqa: F401
from .sample_paths import (  # noqa: F401
    LevelSetResult, MultiOutputSamplePathBatch, SamplePath, SamplePathBatch, draw_dkgp_linear_paths, draw_matheron_paths,
    draw_reference_paths, posterior_mean_path,
)

__all__ = ["PsbaxError", "LevelSetResult", "SamplePath", "SamplePathBatch", "draw_matheron_paths",
           "draw_reference_paths", "posterior_mean_path"]
