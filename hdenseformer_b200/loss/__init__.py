from .combine_loss import CEPlusDice, DeepSuperloss  # noqa: F401
from .cross_entropy import CrossentropyLoss  # noqa: F401
from .dice_loss import BinaryDiceLoss, DiceLoss  # noqa: F401
