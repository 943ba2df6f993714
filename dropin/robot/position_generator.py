from inversekinematicsann_b200.robot.position_generator import *  # noqa: F401,F403
