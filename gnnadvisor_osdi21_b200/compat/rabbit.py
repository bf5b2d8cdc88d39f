"""Drop-in module named `rabbit` (reference: rabbit_module/src/reorder.cpp:293-295 exports `reorder`)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from gnnadvisor_osdi21_b200.reorder import reorder  # noqa: E402,F401
