from ...losses import (LOSSES, BaseLossWrapper, DiceLossWrapper, GeneralizedDiceLossWrapper,  # noqa: F401
                       MultipleLossWrapper, apply_missing_mask)
