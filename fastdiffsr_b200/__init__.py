"""fastdiffsr_b200: B200-native FastDiffSR sampling path behind the reference's define_G API."""
from .config import load_config, dict_to_nonedict  # noqa: F401
from .networks import define_G  # noqa: F401
from .diffusion import GaussianDiffusion, make_beta_schedule  # noqa: F401
from .unet import SR3UNet, UNet  # noqa: F401
from .engine import Engine  # noqa: F401
from ._lib import FdsrError, FdsrOverflowError, build_library  # noqa: F401

__all__ = ["define_G", "GaussianDiffusion", "UNet", "SR3UNet", "Engine", "load_config", "dict_to_nonedict",
           "make_beta_schedule", "FdsrError", "FdsrOverflowError", "build_library"]
