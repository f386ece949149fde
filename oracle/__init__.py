"""CPU oracle for the IV_INTERPOLATION hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the shipped product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and there only as the checker or the timed CPU baseline.
The product package (``iv_interpolation_b200``) never imports this package and
fails loudly when its CUDA library is missing.

Parity status:
  * kernels A (time-linear interpolation + ffill) and B (OHLCV bucket reduce):
    PINNED - ``oracle/make_golden.py`` ran the reference's own classes
    (imported from /root/reference in the build container) on seeded tapes and
    the survey's known-answer cases; the outputs are committed under
    ``tests/golden/`` and ``tests/test_oracle_golden.py`` holds this oracle to
    them bit-for-bit.
  * the data bridge (``oracle/bridge_oracle.py``): PINNED - ``oracle/make_golden_bridge.py`` ran the reference's
    own ``InterpolatedToOHLCVConverter`` / ``_generate_ohlcv_candles`` under a seeded generator;
    ``tests/golden/bridge_cases.npz``, held bit-for-bit including the generator's end state.
  * kernels C (Black-Scholes implied vol) and D (natural cubic spline): the
    reference contains no such code ("parity unpinned" by the reference);
    the oracle is scipy (brentq / CubicSpline) and the Greeks formulas of
    reference ``src/interpolation/greeks.py:21-35`` are pinned by its KATs.
"""
