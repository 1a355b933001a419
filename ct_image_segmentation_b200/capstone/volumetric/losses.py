from ...losses import (WEIGHT, CrossEntropyWrapper3D, DiceLossWrapper3D, FocalLossWrapper3D,  # noqa: F401
                       GeneralizedDiceLossWrapper3D, MultipleLossWrapper3D, WeightedCrossEntropyWrapper3D)
