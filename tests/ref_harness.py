"""tests-side alias of oracle/ref_harness.py (the unmodified-reference builder used for oracle validation)."""
import sys

from oracle import ref_harness as _rh

sys.modules[__name__] = _rh
