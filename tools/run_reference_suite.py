#!/usr/bin/env python
"""Prove the drop-in against the reference's OWN callers: run its unit suites, cli.py and the broker callback,
unmodified, with dropin/ ahead of the reference's packages.

    IK_REFERENCE_ROOT=/path/to/InverseKinematicsANN python tools/run_reference_suite.py [--log FILE] [--stage ...]

Stages
  resolve : (no GPU) every module the reference's callers import resolves to dropin/ (kinematics.*, robot.*), the
            reference's own `tests`, `plot`, `cli`, `rpc_broker` resolve to the reference tree, `cli.CLI()` constructs
  suites  : the reference's tests/{point,fabrik,forward}_unit.py and InverseKinematicsFabrikTest (runtests.py:25-33)
            through unittest, as runtests.py does; the ANN suites run on stand-in weights (the reference's
            tests/test_model.h5 is not in its tree), so their golden vectors cannot be checked -- load/save only
  cli     : dropin/launch.py cli.py --generate-data --shape spring ... / --inverse-kine --method fabrik|ann ...
            --to-file, the CSV compared with the CPU oracle (and with the live reference when it can run here)
  broker  : rpc_broker.IkineRPCBroker.callback driven with a fake pika channel (JSON in, JSON out)

The reference tree is copied to a scratch directory first (its tests write files next to themselves); nothing under
IK_REFERENCE_ROOT is modified.  matplotlib and pika are replaced by inert stubs when they are not installed
(plotting and the AMQP transport are outside the hot path).  Test infrastructure: uses oracle/ as the checker.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "dropin")
LAUNCH = os.path.join(DROPIN, "launch.py")

MPL_STUB = '''
class _Inert:
    """Accepts any attribute access / call and returns itself: enough for plot/plot.py to run headless."""
    def __getattr__(self, name):
        return self
    def __call__(self, *a, **k):
        return self
    def __iter__(self):
        return iter(())
import sys
sys.modules[__name__ + ".pyplot"] = _Inert()
pyplot = sys.modules[__name__ + ".pyplot"]
'''

PIKA_STUB = '''
class ConnectionParameters:
    def __init__(self, **kw):
        self.kw = kw
class BasicProperties:
    def __init__(self, correlation_id=None, reply_to=None, **kw):
        self.correlation_id, self.reply_to = correlation_id, reply_to
class _Channel:
    def __init__(self):
        self.published, self.acked, self.consumers = [], [], []
    def queue_declare(self, queue):
        return queue
    def basic_qos(self, prefetch_count):
        self.prefetch_count = prefetch_count
    def basic_consume(self, queue, on_message_callback):
        self.consumers.append((queue, on_message_callback))
    def basic_publish(self, exchange, routing_key, properties, body):
        self.published.append({"routing_key": routing_key, "correlation_id": properties.correlation_id, "body": body})
    def basic_ack(self, delivery_tag):
        self.acked.append(delivery_tag)
    def start_consuming(self):
        pass
class BlockingConnection:
    def __init__(self, params):
        self.params, self._channel = params, _Channel()
    def channel(self):
        return self._channel
'''

RESOLVE_SNIPPET = r'''
import importlib, json, os, sys
out = {}
for name in ("kinematics.inverse", "kinematics.point", "kinematics.fabrik", "kinematics.forward", "kinematics.ann",
             "robot.robot", "robot.position_generator", "tests.point_unit", "plot.plot", "cli", "rpc_broker"):
    out[name] = os.path.realpath(importlib.import_module(name).__file__)
import cli
cli.CLI()            # constructs every Command, including RandomCommand (cli.py:312-319)
from kinematics.inverse import FabrikInverseKinematics
out["impl"] = os.path.realpath(sys.modules[FabrikInverseKinematics.__module__].__file__)
print("RESOLVE " + json.dumps(out))
'''

SUITES_SNIPPET = r'''
import sys, unittest
from tests.point_unit import point_test_suite
from tests.fabrik_unit import fabrik_test_suite
from tests.forward_unit import fwkine_test_suite
from tests.inverse_unit import InverseKinematicsFabrikTest, InverseKinematicsAnnTest
from tests.ann_unit import AnnTest
def fabrik_ikine_suite():
    s = unittest.TestSuite(); s.addTest(InverseKinematicsFabrikTest('inverse_kine')); return s
def ann_files_suite():      # stand-in weights: the file handling of ann_unit.py:23-35, not its golden vectors
    s = unittest.TestSuite(); s.addTest(AnnTest('load_model')); s.addTest(AnnTest('save_model')); return s
def ann_golden_suite():     # golden vectors of a model that is not in the reference tree: expected to differ
    s = unittest.TestSuite(); s.addTest(AnnTest('predict')); s.addTest(InverseKinematicsAnnTest('inverse_kine')); return s
for name, make in (("point", point_test_suite), ("fabrik", fabrik_test_suite), ("forward", fwkine_test_suite),
                   ("inverse_fabrik", fabrik_ikine_suite), ("ann_files_standin", ann_files_suite),
                   ("ann_golden_standin", ann_golden_suite)):
    res = unittest.TextTestRunner(verbosity=2, stream=sys.stdout).run(make())
    print(f"SUITE {name}: run={res.testsRun} failures={len(res.failures)} errors={len(res.errors)}")
'''

BROKER_SNIPPET = r'''
import json, sys, types
sys.argv = ["rpc_broker.py", "--method", "fabrik"]
import rpc_broker
from pika import BasicProperties
engine = rpc_broker.get_ikine_engine_cli()
broker = rpc_broker.IkineRPCBroker(engine)
chan = broker._IkineRPCBroker__channel
assert chan.prefetch_count == 1 and chan.consumers[0][0] == "ikine_queue"
method = types.SimpleNamespace(delivery_tag=7)
pts = json.loads(sys.stdin.read())
broker.callback(chan, method, BasicProperties(correlation_id="c-1", reply_to="reply_q"), json.dumps({"positions": pts}).encode())
broker.callback(chan, method, BasicProperties(correlation_id="c-2", reply_to="reply_q"),
                json.dumps({"positions": [[1.0, 2.1, 3.0], [1.567, 2.22, -3.123]]}).encode())
broker.callback(chan, method, BasicProperties(correlation_id="c-3", reply_to="reply_q"),
                json.dumps({"positions": [[1.0, 2.1]]}).encode())
print("BROKER " + json.dumps({"published": chan.published, "acked": chan.acked}))
'''


class Log:
    def __init__(self, path):
        self.lines, self.path = [], path

    def __call__(self, text=""):
        print(text, flush=True)
        self.lines.append(text)

    def save(self):
        if self.path:
            os.makedirs(os.path.dirname(os.path.abspath(self.path)), exist_ok=True)
            with open(self.path, "w") as f:
                f.write("\n".join(self.lines) + "\n")


def importable(name):
    return subprocess.run([sys.executable, "-c", f"import {name}"], capture_output=True).returncode == 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log", default="")
    ap.add_argument("--stage", default="resolve,suites,cli,broker")
    ap.add_argument("--keep", action="store_true", help="keep the scratch directory")
    args = ap.parse_args()
    ref = os.environ.get("IK_REFERENCE_ROOT", "")
    if not ref or not os.path.exists(os.path.join(ref, "cli.py")):
        sys.exit("IK_REFERENCE_ROOT must point at a checkout of lstar93/InverseKinematicsANN (cli.py not found)")
    log = Log(args.log)
    stages = args.stage.split(",")
    scratch = tempfile.mkdtemp(prefix="ikb_refsuite_")
    work, stubs = os.path.join(scratch, "reference"), os.path.join(scratch, "stubs")
    shutil.copytree(ref, work, ignore=shutil.ignore_patterns("*.gif", "__pycache__", ".git"))
    os.makedirs(stubs)
    stubbed = []
    if not importable("matplotlib"):
        os.makedirs(os.path.join(stubs, "matplotlib"))
        open(os.path.join(stubs, "matplotlib", "__init__.py"), "w").write(MPL_STUB)
        stubbed.append("matplotlib")
    if not importable("pika"):
        os.makedirs(os.path.join(stubs, "pika"))
        open(os.path.join(stubs, "pika", "__init__.py"), "w").write(PIKA_STUB)
        stubbed.append("pika")
    # stand-in for the reference's missing tests/test_model.h5: the network trained in this repo
    model = os.path.join(ROOT, "models", "roboarm_b200_r01")
    for suffix in (".npz", "_scaler_x.bin", "_scaler_y.bin"):
        shutil.copy(model + suffix, os.path.join(work, "tests", "test_model" + suffix))
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    env.pop("PYTHONSAFEPATH", None)
    launcher_env = dict(env, PYTHONPATH=stubs)
    safepath_env = dict(env, PYTHONSAFEPATH="1", PYTHONPATH=os.pathsep.join([DROPIN, work, stubs]))
    log(f"# reference suite through dropin/  (reference copy: {work}; stubs: {stubbed or 'none'})")
    failures = []

    def run(cmd, env, label, stdin=None, check=True):
        log(f"\n$ {label}")
        p = subprocess.run(cmd, cwd=work, env=env, input=stdin, capture_output=True, text=True)
        out = (p.stdout + p.stderr).rstrip()
        for line in out.splitlines():
            if "Warning" not in line and "warnings.warn" not in line:
                log("  " + line)
        if check and p.returncode != 0:
            failures.append(f"{label}: exit code {p.returncode}")
        return p

    def under(path, base):
        return os.path.realpath(path).startswith(os.path.realpath(base) + os.sep)

    if "resolve" in stages:
        # both documented ways of putting dropin/ first: PYTHONSAFEPATH=1 + PYTHONPATH, and the launcher
        p = run([sys.executable, "-c", RESOLVE_SNIPPET], safepath_env,
                "PYTHONSAFEPATH=1 PYTHONPATH=dropin:<reference> python -c <imports>")
        line = next((l for l in p.stdout.splitlines() if l.startswith("RESOLVE ")), None)
        if line is None:
            failures.append("resolve: no result")
        else:
            where = json.loads(line[8:])
            for name, path in where.items():
                if name.startswith(("kinematics.", "robot.")):
                    ok = under(path, DROPIN)
                elif name == "impl":
                    ok = under(path, os.path.join(ROOT, "inversekinematicsann_b200"))
                else:
                    ok = under(path, work)
                log(f"  {'ok ' if ok else 'BAD'} {name} -> {path}")
                if not ok:
                    failures.append(f"resolve: {name} -> {path}")
        probe = os.path.join(scratch, "where.py")
        open(probe, "w").write("import kinematics.inverse as m, tests.point_unit as t\nprint('LAUNCH', m.__file__, t.__file__)\n")
        shutil.copy(probe, os.path.join(work, "where_probe.py"))
        p = run([sys.executable, LAUNCH, "where_probe.py"], launcher_env, "python dropin/launch.py <script in the reference root>")
        got = next((l.split()[1:] for l in p.stdout.splitlines() if l.startswith("LAUNCH ")), None)
        if not got or not under(got[0], DROPIN) or not under(got[1], work):
            failures.append(f"launcher: resolved {got}")
        os.remove(os.path.join(work, "where_probe.py"))

    if "suites" in stages:
        p = run([sys.executable, "-c", SUITES_SNIPPET], safepath_env,
                "the reference's unit suites (runtests.py:25-33) on dropin/", check=False)
        seen = {}
        for l in p.stdout.splitlines():
            if l.startswith("SUITE "):
                name = l.split()[1].rstrip(":")
                seen[name] = dict(kv.split("=") for kv in l.split()[2:])
        for name in ("point", "fabrik", "forward", "inverse_fabrik", "ann_files_standin"):
            r = seen.get(name)
            if not r or int(r["failures"]) or int(r["errors"]) or not int(r["run"]):
                failures.append(f"suite {name}: {r}")
        log(f"  (ann_golden_standin compares stand-in weights with golden vectors of the reference's absent "
            f"tests/test_model.h5: {seen.get('ann_golden_standin')} -- informational)")

    if "cli" in stages:
        import numpy as np
        import pandas as pd
        sys.path.insert(0, ROOT)
        from oracle import c_oracle, np_oracle
        run([sys.executable, LAUNCH, "cli.py", "--generate-data", "--shape", "spring", "--samples", "50", "--dim", "2,3,6",
             "--to-file", "spring.csv"], launcher_env, "dropin/launch.py cli.py --generate-data --shape spring --samples 50 --dim 2,3,6 --to-file spring.csv")
        pts = pd.read_csv(os.path.join(work, "spring.csv")).values
        want_pts = np_oracle.spring(50, 2, 3, 6)
        d = float(np.abs(pts - want_pts).max())
        log(f"  spring.csv: {pts.shape[0]} points, max |d| vs position_generator.py:72-78 restatement = {d:.2e}")
        if pts.shape != (50, 3) or d > 1e-12:
            failures.append("cli: spring.csv differs from the reference shape")
        run([sys.executable, LAUNCH, "cli.py", "--inverse-kine", "--method", "fabrik", "--points", "spring.csv",
             "--to-file", "angles.csv", "--verbose"], launcher_env,
            "dropin/launch.py cli.py --inverse-kine --method fabrik --points spring.csv --to-file angles.csv --verbose")
        got = pd.read_csv(os.path.join(work, "angles.csv"))
        want = c_oracle.fabrik_ikine(pts)["angles"]
        d = float(np.abs(got.values - want).max())
        log(f"  angles.csv: columns {list(got.columns)}, max |dtheta| vs oracle = {d:.2e} rad")
        if list(got.columns) != ["theta1", "theta2", "theta3", "theta4"] or not d <= 1e-9:
            failures.append(f"cli: fabrik angles differ from the oracle by {d}")
        try:
            from oracle import ref_import
            ref_ik, _ = ref_import.fabrik_ikine_with_iterations(ref_import.load(), pts.tolist())
            d = float(np.abs(got.values - np.array(ref_ik)).max())
            log(f"  angles.csv vs the live reference's own FabrikInverseKinematics.ikine: max |dtheta| = {d:.2e} rad")
            if not d <= 1e-9:
                failures.append(f"cli: fabrik angles differ from the live reference by {d}")
        except (ImportError, RuntimeError) as exc:
            log(f"  (live reference not importable here: {type(exc).__name__}: {exc})")
        run([sys.executable, LAUNCH, "cli.py", "--inverse-kine", "--method", "ann", "--model", "tests/test_model.h5",
             "--points", "spring.csv", "--to-file", "ann.csv"], launcher_env,
            "dropin/launch.py cli.py --inverse-kine --method ann --model tests/test_model.h5 --points spring.csv --to-file ann.csv  (stand-in weights)")
        from joblib import load
        data = np.load(model + ".npz")
        nl = len([k for k in data.files if k.startswith("W")])
        sx, sy = load(model + "_scaler_x.bin"), load(model + "_scaler_y.bin")
        want = np_oracle.mlp_predict(pts, [data[f"W{i}"] for i in range(nl)], [data[f"b{i}"] for i in range(nl)],
                                     sx.mean_, sx.scale_, sy.mean_, sy.scale_)
        d = float(np.abs(pd.read_csv(os.path.join(work, "ann.csv")).values - want).max())
        log(f"  ann.csv: max |dtheta| vs the fp32 restatement of ann.py:70-76 = {d:.2e} rad")
        if not d <= 1e-5:
            failures.append(f"cli: ann angles differ from the oracle by {d}")
        bad = pts.copy()
        bad[3, 2] = -3.123                          # the out-of-reach row of tests/inverse_unit.py:33
        pd.DataFrame(bad, columns=["x", "y", "z"]).to_csv(os.path.join(work, "bad.csv"), index=False)
        p = run([sys.executable, LAUNCH, "cli.py", "--inverse-kine", "--method", "fabrik", "--points", "bad.csv",
                 "--to-file", "never.csv"], launcher_env, "dropin/launch.py cli.py ... --points bad.csv  (one row outside the workspace)")
        if "is out of manipulator reach area" not in p.stdout or os.path.exists(os.path.join(work, "never.csv")):
            failures.append("cli: out-of-reach CSV did not print the reference's message")

    if "broker" in stages:
        import numpy as np
        sys.path.insert(0, ROOT)
        from oracle import c_oracle, np_oracle
        pts = np_oracle.circle(2, 40, (2, 0, 2))
        p = run([sys.executable, "-c", BROKER_SNIPPET], safepath_env,
                "rpc_broker.IkineRPCBroker(engine).callback(...) with a fake pika channel", stdin=json.dumps(pts.tolist()))
        line = next((l for l in p.stdout.splitlines() if l.startswith("BROKER ")), None)
        if line is None:
            failures.append("broker: no result")
        else:
            res = json.loads(line[7:])
            r1, r2, r3 = (json.loads(m["body"]) for m in res["published"])
            want = c_oracle.fabrik_ikine(pts)["angles"]
            d = float(np.abs(np.array(r1["angles"]) - want).max())
            log(f"  reply 1: status {r1['status']}, {len(r1['angles'])} rows, max |dtheta| vs oracle = {d:.2e} rad")
            log(f"  reply 2: {r2}")
            log(f"  reply 3: {r3}")
            ok = (r1["status"] == "OK" and d <= 1e-9 and r2["status"] == "ERROR" and r2["correlation_id"] == "c-2"
                  and "is out of manipulator reach area" in r2["reason"] and r3["status"] == "ERROR"
                  and res["acked"] == [7, 7, 7] and all(m["routing_key"] == "reply_q" for m in res["published"]))
            if not ok:
                failures.append("broker: replies differ from rpc_broker.py:76-100 semantics")

    log("\n# result: " + ("ALL GREEN" if not failures else "FAILED: " + "; ".join(failures)))
    log.save()
    if not args.keep:
        shutil.rmtree(scratch, ignore_errors=True)
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
