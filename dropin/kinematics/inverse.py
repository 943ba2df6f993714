from inversekinematicsann_b200.kinematics.inverse import *  # noqa: F401,F403
from inversekinematicsann_b200.kinematics.inverse import (AnnInverseKinematics, FabrikInverseKinematics,  # noqa: F401
                                                          InverseKinematics)
