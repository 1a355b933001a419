from ...losses import DiceLossWrapper3D, GeneralizedDiceLossWrapper3D, MultipleLossWrapper3D  # noqa: F401
