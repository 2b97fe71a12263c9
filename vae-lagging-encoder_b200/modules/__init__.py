"""Drop-in replacement for the reference `modules` package (reference modules/__init__.py:1-6):
`from modules import VAE, LSTMEncoder, LSTMDecoder, generate_grid` binds here when this directory's
parent (`vae-lagging-encoder_b200/`) precedes the reference on sys.path, so the reference's
text.py / toy.py run unmodified on the B200 kernels.  See INTEGRATION.md."""
from .utils import generate_grid, log_sum_exp  # noqa: F401
from .text import DecoderBase, GaussianEncoderBase, LSTMDecoder, LSTMEncoder  # noqa: F401
from .image import MaskedConv2d, PixelCNNDecoderV2, ResNetEncoderV2  # noqa: F401
from .vae import VAE  # noqa: F401
