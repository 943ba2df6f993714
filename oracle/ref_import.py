"""Import the UNMODIFIED reference from /root/reference (this container only).

TEST INFRASTRUCTURE -- never imported by the product path.

The reference's ``kinematics/inverse.py`` imports ``kinematics/ann.py`` at module top
(reference inverse.py:11) which imports keras (reference ann.py:9-13).  keras/tensorflow are not
installed, so five empty stand-in modules are registered in ``sys.modules`` before the import;
nothing from them is touched on the FABRIK / FK path.  The reference tree is read-only, so byte
code writing is disabled.  ``/root/reference`` does not exist on the GPU box: callers must check
``available()`` and skip.
"""
import importlib
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get("IK_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "kinematics", "inverse.py"))


def _stub_keras():
    if "keras" in sys.modules and not getattr(sys.modules["keras"], "_ik_stub", False):
        return  # a real keras is present
    names = {
        "keras": ["activations"],
        "keras.models": ["load_model", "Sequential"],
        "keras.optimizers": ["Adam"],
        "keras.layers": ["Dense", "Input"],
        "keras.callbacks": ["EarlyStopping"],
    }
    for mod_name, attrs in names.items():
        mod = types.ModuleType(mod_name)
        mod._ik_stub = True
        for attr in attrs:
            setattr(mod, attr, None)
        sys.modules[mod_name] = mod


class _RefModules:
    """The reference's modules, imported under a private prefix-free namespace swap."""

    def __init__(self):
        if not available():
            raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
        sys.dont_write_bytecode = True
        _stub_keras()
        # The product repo ships drop-in packages also named `kinematics` / `robot`.  Import the
        # reference's under their own names with the reference root FIRST on sys.path, then
        # remove them from sys.modules again so nothing else resolves to them by accident.
        saved = {k: v for k, v in sys.modules.items()
                 if k == "kinematics" or k.startswith("kinematics.")
                 or k == "robot" or k.startswith("robot.")}
        for k in saved:
            del sys.modules[k]
        sys.path.insert(0, REFERENCE_ROOT)
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                self.point = importlib.import_module("kinematics.point")
                self.fabrik = importlib.import_module("kinematics.fabrik")
                self.forward = importlib.import_module("kinematics.forward")
                self.inverse = importlib.import_module("kinematics.inverse")
                self.robot = importlib.import_module("robot.robot")
                self.position_generator = importlib.import_module("robot.position_generator")
        finally:
            sys.path.remove(REFERENCE_ROOT)
            for k in [k for k in sys.modules
                      if k == "kinematics" or k.startswith("kinematics.")
                      or k == "robot" or k.startswith("robot.")]:
                del sys.modules[k]
            sys.modules.update(saved)


_CACHE = None


def load() -> _RefModules:
    global _CACHE
    if _CACHE is None:
        _CACHE = _RefModules()
    return _CACHE


def fresh_robot_constants(ref):
    """Deep copies of the robot constants (the reference mutates dh_matrix[0][0], inverse.py:125)."""
    r = ref.robot.SixDOFRobot
    return ([list(row) for row in r.dh_matrix], list(r.links_lengths),
            {k: list(v) for k, v in r.effector_workspace_limits.items()})


def fabrik_ikine_with_iterations(ref, points):
    """Run the reference FabrikInverseKinematics.ikine and also return per-target iteration counts
    by counting calls of the name-mangled private ``Fabrik._Fabrik__backward`` (fabrik.py:19)."""
    dh, links, limits = fresh_robot_constants(ref)
    ik = ref.inverse.FabrikInverseKinematics(dh, links, limits)
    counts = []
    orig = ref.fabrik.Fabrik._Fabrik__backward

    def counting(self, pts, goal):
        counts[-1] += 1
        return orig(self, pts, goal)

    ref.fabrik.Fabrik._Fabrik__backward = counting
    try:
        angles = []
        for p in points:
            counts.append(0)
            angles.extend(ik.ikine([p]))
    finally:
        ref.fabrik.Fabrik._Fabrik__backward = orig
    return angles, counts
