from .algorithm import Algorithm  # noqa: F401
from .topk import TopK  # noqa: F401
from .levelset import LevelSetEstimator  # noqa: F401
from .discobax import SubsetSelect  # noqa: F401
from .lbfgsb import LBFGSB  # noqa: F401
from .evolution import EvolutionStrategies  # noqa: F401
