from ...metrics import DiceMetricWrapper  # noqa: F401
