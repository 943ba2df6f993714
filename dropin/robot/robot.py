from inversekinematicsann_b200.robot.robot import *  # noqa: F401,F403
