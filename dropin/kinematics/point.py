from inversekinematicsann_b200.kinematics.point import *  # noqa: F401,F403
