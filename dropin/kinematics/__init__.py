"""Drop-in shim: with this directory ahead of the reference's own packages on sys.path, the reference's callers
(`from kinematics.inverse import ...`, cli.py:15-20, rpc_broker.py:13-16, tests/*_unit.py) resolve to the B200
engine.  `python dropin/launch.py cli.py ...` arranges that; see INTEGRATION.md section 1.

The repository root (home of `inversekinematicsann_b200`) is APPENDED to sys.path, so it can never shadow one of
the reference's own top-level names (`tests`, `plot`, `examples`, ...)."""
import os
import sys

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _REPO_ROOT not in sys.path:
    sys.path.append(_REPO_ROOT)
