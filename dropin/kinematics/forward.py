from inversekinematicsann_b200.kinematics.forward import *  # noqa: F401,F403
