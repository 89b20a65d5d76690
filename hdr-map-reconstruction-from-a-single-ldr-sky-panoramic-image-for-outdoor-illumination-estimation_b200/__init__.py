"""B200-native distortion-aware convolution path (sky-dome HDR reconstruction hot path).

Importing this package loads libskydome_b200.so; a missing library is an ImportError (no CPU fallback)."""
from . import _lib                      # noqa: F401  (fails loudly if the CUDA library is absent)
from . import distortion_aware_ops     # noqa: F401
from .distortion_aware_ops import conv2d, deconv2d   # noqa: F401
from . import generator                 # noqa: F401
from .generator import resBlock, resLayer, InstanceNormalization   # noqa: F401
from . import sharding                  # noqa: F401
from . import _flat                      # noqa: F401
from . import ops                        # noqa: F401
from .generator import model             # noqa: F401
from . import trunk_train              # noqa: F401
from . import sunpose_net              # noqa: F401
from . import sunrad_net               # noqa: F401
from . import grad_cam                 # noqa: F401
from . import inference                # noqa: F401
from . import tf_utils, discriminator, vgg16, train   # noqa: F401
from . import train_sun   # noqa: F401
from . import utils   # noqa: F401
from . import dataset   # noqa: F401
