from ...losses import MultipleLossWrapper  # noqa: F401
from ...metrics import DiceMetricWrapper  # noqa: F401
from ...unet import UNet  # noqa: F401
