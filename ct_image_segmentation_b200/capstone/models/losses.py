from ...losses import (LOSSES, WEIGHT, BaseLossWrapper, CrossEntropyWrapper, DiceLossWrapper,  # noqa: F401
                       FocalLossWrapper, GeneralizedDiceLossWrapper, MultipleLossWrapper,
                       WeightedCrossEntropyWrapper, apply_missing_mask)
