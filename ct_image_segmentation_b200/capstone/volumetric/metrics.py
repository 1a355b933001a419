from ...metrics import DiceMetricWrapper3D  # noqa: F401
