"""Drop-in for the reference's ``src/interpolation/core.py``: put ``iv_interpolation_b200/dropin``
on ``sys.path`` where the reference scripts put ``src/`` (reference main.py:21) and
``from interpolation.core import IVInterpolator`` (batch_processor.py:9, complete_pipeline.py:34)
resolves to the B200 implementation."""
from iv_interpolation_b200.interpolation.core import IVInterpolator  # noqa: F401
