"""Drop-in for the reference's ``src/candle_reconstruction/core.py`` (see dropin/interpolation/core.py).
``MultiSymbolCandleReconstructor`` (reference core.py:174-320) is database I/O and is out of scope."""
from iv_interpolation_b200.candle_reconstruction.core import CandleReconstructor  # noqa: F401
