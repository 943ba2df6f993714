from inversekinematicsann_b200.kinematics.ann import *  # noqa: F401,F403
