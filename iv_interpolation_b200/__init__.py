"""iv_interpolation_b200 — B200-native (sm_100a) implementation of IV_INTERPOLATION's hot path.

Drop-in classes (same names/signatures as the reference):
    from iv_interpolation_b200.interpolation.core import IVInterpolator
    from iv_interpolation_b200.candle_reconstruction.core import CandleReconstructor
or put ``iv_interpolation_b200/dropin`` on ``sys.path`` in place of the reference's ``src/``.

There is no CPU fallback: every compute entry point calls the CUDA library
(``iv_interpolation_b200/csrc/libivx.so``) through its C ABI and raises if it is missing.
"""
__version__ = "0.1.0"
