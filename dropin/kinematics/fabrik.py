from inversekinematicsann_b200.kinematics.fabrik import *  # noqa: F401,F403
