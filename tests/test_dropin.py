"""The zero-edit drop-in: the reference's own callers on top of dropin/.

CPU: with dropin/ ahead of the reference tree every `kinematics.*` / `robot.*` import of cli.py, rpc_broker.py and
the reference's tests resolves to the shims while the reference's own `tests`, `plot`, `cli`, `rpc_broker` stay its
own, and `cli.CLI()` constructs.  GPU: the reference's unit suites, cli.py and the broker callback run unmodified
(tools/run_reference_suite.py).  Both need a checkout of the reference (IK_REFERENCE_ROOT, /root/reference, or
an untracked .refscratch/ copy placed there for one run on the GPU box, which has no /root/reference: that run's log is
profiles/r02_dropin_suite.log); without one the tests skip."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_root():
    for cand in (os.environ.get("IK_REFERENCE_ROOT", ""), "/root/reference", os.path.join(ROOT, ".refscratch")):
        if cand and os.path.isfile(os.path.join(cand, "cli.py")):
            return cand
    return None


def _run(stage):
    ref = _reference_root()
    if ref is None:
        pytest.skip("no checkout of the reference available")
    env = dict(os.environ, IK_REFERENCE_ROOT=ref)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_suite.py"), "--stage", stage],
                       env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "ALL GREEN" in p.stdout, p.stdout[-4000:] + p.stderr[-2000:]
    return p.stdout


def test_reference_imports_resolve_to_the_dropin():
    out = _run("resolve")
    assert "ok  kinematics.inverse" in out and "ok  cli" in out and "LAUNCH" in out


def test_plain_pythonpath_is_not_enough():
    """Why INTEGRATION.md prescribes the launcher (or PYTHONSAFEPATH): `python cli.py` puts the script directory at
    sys.path[0], ahead of PYTHONPATH, so the reference's own packages would win."""
    ref = _reference_root()
    if ref is None:
        pytest.skip("no checkout of the reference available")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "dropin"), PYTHONDONTWRITEBYTECODE="1")
    env.pop("PYTHONSAFEPATH", None)
    probe = "import sys; sys.path.insert(0, sys.argv[1]); import kinematics.point as m; print(m.__file__)"
    # emulate `python <ref>/script.py`: the script directory goes to sys.path[0]
    p = subprocess.run([sys.executable, "-c", probe, ref], env=env, capture_output=True, text=True, cwd=ref)
    assert os.path.realpath(p.stdout.strip()).startswith(os.path.realpath(ref))


@pytest.mark.gpu
def test_reference_suites_cli_and_broker_run_on_the_engine():
    out = _run("suites,cli,broker")
    assert "SUITE inverse_fabrik: run=1 failures=0 errors=0" in out
