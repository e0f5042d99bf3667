"""srb200 -- B200-native (sm_100a) engine for the conv + PixelShuffle hot path of
togheppi/pytorch-super-resolution-model-collection (base_networks.py blocks).

Importing this package loads libsrb200.so; if the CUDA library is not built the import fails
(no CPU / torch fallback exists on the product path)."""
from . import _lib  # noqa: F401  (raises ImportError when the engine is missing)
from ._lib import launch_count, SrbError, LIB_PATH
from . import functional
from .functional import conv2d, conv_transpose2d, prelu, set_math, get_math, set_grad_scale, set_fuse_relu_backward, mse_loss, l1_loss, image_to_tensor, conv2d_loss, enable_weight_cache, repack_weights, weight_cache_entries
from . import base_networks
from .base_networks import DenseBlock, ConvBlock, DeconvBlock, ResnetBlock, PSBlock, Upsample2xBlock, prepare, FusedLoss
from .convert import convert, PReLU, ConvTranspose2d, Conv2d
from .ddp import GradBucket
from .graphs import TrainStepGraphs
from .optim import FlatAdam
from . import nn_ops
from .nn_ops import batch_norm_act, linear, max_pool2, bce_loss
from . import models, host

__version__ = "0.1.0"
