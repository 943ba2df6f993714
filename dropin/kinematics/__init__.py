"""Drop-in shim: put this directory first on PYTHONPATH and the reference's callers
(`from kinematics.inverse import ...`, cli.py:15-20, rpc_broker.py:13-16) resolve to the B200 engine."""
