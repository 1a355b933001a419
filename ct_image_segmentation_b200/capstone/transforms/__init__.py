from ...transforms import WINDOWING_CONFIG, window_normalize  # noqa: F401
