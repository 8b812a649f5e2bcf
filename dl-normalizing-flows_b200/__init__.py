"""rnvp-b200: B200-native RealNVP coupling-stack hot path behind the reference's module API.

The directory is meant to be put on ``sys.path`` (so that ``flow_realnvp``,
``modules_realnvp`` and ``utils`` resolve to these drop-ins).
Importing it as a package (``importlib.import_module('dl-normalizing-flows_b200')``)
does that and re-exports the public names.
"""
import os as _os
import sys as _sys

_HERE = _os.path.dirname(_os.path.abspath(__file__))
if _HERE not in _sys.path:
    _sys.path.insert(0, _HERE)

import rnvp_cabi  # noqa: E402  (raises ImportError when librnvp_b200.so is missing: no fallback)
from flow_realnvp import RealNVP  # noqa: E402
from modules_realnvp import (  # noqa: E402
    AbstractCoupling, ChannelwiseAffineCoupling, CheckerboardAffineCoupling,
    ResidualBlock, ResidualModule, WeightNormConv2d)
from utils import Hyperparameters, logit_transform  # noqa: E402
from rnvp_engine import set_default_math  # noqa: E402
import rnvp_optim  # noqa: E402

__all__ = ["RealNVP", "AbstractCoupling", "ChannelwiseAffineCoupling", "CheckerboardAffineCoupling",
           "ResidualBlock", "ResidualModule", "WeightNormConv2d", "Hyperparameters", "logit_transform",
           "set_default_math", "rnvp_cabi", "rnvp_optim"]
