"""Data bridge (reference src/data_bridge): interpolated rows -> synthetic 1-minute OHLCV candles."""
from .ohlcv_converter import InterpolatedToOHLCVConverter  # noqa: F401
