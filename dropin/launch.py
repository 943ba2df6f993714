#!/usr/bin/env python
"""Run one of the reference's scripts, unmodified, on top of the B200 engine.

    python /path/to/repo/dropin/launch.py cli.py --inverse-kine --method fabrik --points spring.csv --to-file out.csv
    python /path/to/repo/dropin/launch.py rpc_broker.py --method fabrik
    python /path/to/repo/dropin/launch.py runtests.py

`python cli.py` puts the script's own directory at sys.path[0], ahead of PYTHONPATH, so the reference's
`kinematics/` and `robot/` packages would win over the shims.  This launcher does what the interpreter does for
a script -- sys.argv, `__main__`, the script directory on sys.path -- except that `dropin/` comes first:

    sys.path = [<repo>/dropin, <script dir>, ...the usual entries..., <repo>]

(the repository root is appended by the shims themselves, last, so that it shadows nothing of the reference).
Equivalent without the launcher:  PYTHONSAFEPATH=1 PYTHONPATH=<repo>/dropin:<reference root> python cli.py ...
"""
import os
import runpy
import sys


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    script = os.path.abspath(sys.argv[1])
    here = os.path.dirname(os.path.abspath(__file__))
    # drop the launcher's own directory entry (sys.path[0]) and put dropin/ + the script's directory in front
    sys.path[:] = [here, os.path.dirname(script)] + [p for p in sys.path[1:] if os.path.abspath(p or ".") != here]
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
