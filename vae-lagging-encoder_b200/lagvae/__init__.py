"""lagvae: host-side binding of the B200 kernels for the aggressive-inner-loop hot path."""
from ._backend import LagvaeError, lib, launch_count, lstm_variant  # noqa: F401
from .engine import DropoutSpec, TextEngine, PARAM_NAMES, param_shapes  # noqa: F401
from .graph import GraphedStep  # noqa: F401
from .optim import BucketComm, ClipAdam  # noqa: F401
