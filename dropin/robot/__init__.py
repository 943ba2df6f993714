"""Drop-in shim for the reference's `robot` package (robot/robot.py constants and exception type,
robot/position_generator.py shapes).  See dropin/kinematics/__init__.py for the sys.path arrangement."""
import os
import sys

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _REPO_ROOT not in sys.path:
    sys.path.append(_REPO_ROOT)
