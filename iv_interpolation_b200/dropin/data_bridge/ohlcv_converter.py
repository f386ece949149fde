"""Drop-in for the reference's ``src/data_bridge/ohlcv_converter.py`` (see dropin/interpolation/core.py):
``from data_bridge.ohlcv_converter import InterpolatedToOHLCVConverter`` (reference main.py:29) resolves to the B200
implementation.  Its database methods are out of scope; ``_generate_ohlcv_from_interpolated`` and ``convert_frames``
are the compute entries."""
from iv_interpolation_b200.data_bridge.ohlcv_converter import InterpolatedToOHLCVConverter  # noqa: F401
