"""Drop-in module named `GNNAdvisor`: put this directory on sys.path and the reference's
`import GNNAdvisor as GNNA` (gnn_conv.py:4, GNNA_main.py:10, unitest.py:3) resolves to the B200
runtime.  Same six exports as the reference extension (GNNAdvisor.cpp:253-263)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from gnnadvisor_osdi21_b200.ops import SAG, forward, backward, forward_gin, backward_gin, build_part  # noqa: E402,F401
