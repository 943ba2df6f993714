"""CPU oracle for the IK hot path -- TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  ``inversekinematicsann_b200`` never does: the product path fails
loudly when its CUDA library is missing instead of falling back to anything here.

* ``oracle.c_oracle``  -- ctypes binding of ``ik_oracle.c`` (plain C fp64 restatement of the
  reference's FABRIK + angle extraction + DH forward kinematics; parity PINNED against the
  unmodified reference and its golden vectors).
* ``oracle.np_oracle`` -- NumPy restatements: vectorised fp64 FABRIK (cross-check of the C one),
  the sklearn-StandardScaler -> Keras-Dense MLP -> StandardScaler composition of reference
  ``kinematics/ann.py:70-76`` in fp32 (and fp64), and the trajectory generators of
  ``robot/position_generator.py``.  ANN parity is UNPINNED: the shipped ``.h5`` weights and the
  reference tests' ``tests/test_model.h5`` are absent from the reference mount and keras is not
  installed, so no golden output of the real network exists to check against.
* ``oracle.ref_import`` -- imports the unmodified reference from ``/root/reference`` (this container
  only; absent on the GPU box) to validate the restatements and to generate ``tests/golden``.

The reference is pure Python: there is nothing to compile into ``oracle/_ref/``.
"""
