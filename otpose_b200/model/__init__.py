from .ConvVideoTransformer import ConvTransformer  # noqa: F401
from .OTPose import OTPose, default_cfg  # noqa: F401
