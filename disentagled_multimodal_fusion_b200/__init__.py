"""B200-native hot path of Hassan-Sarwat/disentagled_multimodal_fusion: the encoder + InfoNCE + EDL
training step on hand-written sm_100a kernels (libdmf_b200.so, C ABI in include/dmf_b200.h).

Importing this package loads the shared library; it fails loudly when the library is missing and
every op raises on a non-B200 device.  There is no CPU / eager fallback.
"""
from . import _lib  # noqa: F401  (loads libdmf_b200.so or raises ImportError)
from . import ops  # noqa: F401
from .baselines import LateFusion  # noqa: F401
from .classifiers import MLP, EvidentialNN, IdentityEncoder, Linear  # noqa: F401
from .disentangledssl import DisentangledSSL  # noqa: F401
from .dmvae import DMVAE  # noqa: F401
from .evidential_probe import DisentangledEvidentialProbeModule, EvidentialProbeModule  # noqa: F401
from .losses import AvgTrustedLoss, SupConLoss, ortho_loss  # noqa: F401
from .optim import FusedAdam, FusedAdamW  # noqa: F401
from . import analysis  # noqa: F401

__version__ = "0.1.0"
