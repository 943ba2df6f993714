"""Model-side constants and synthetic weights for benchmarking the ANN path.

The reference ships only the two StandardScalers of ``models/roboarm_model_1674153800-982793`` (the ``.h5`` with
the Keras weights is missing from its tree), so throughput runs use seeded synthetic weights of the architecture in
reference ``kinematics/ann.py:46-56`` (Input(3), 12 x Dense(500, tanh), Dense(4)) together with the real scalers.
"""
import numpy as np

# sklearn StandardScaler statistics of the shipped model (x: effector position, y: the four joint angles)
SHIPPED_MEAN_X = np.array([2.2073088909641334, 0.19405985835497927, 1.494994275926956])
SHIPPED_SCALE_X = np.array([1.7144761363570307, 2.7973201512836416, 2.140079230865925])
SHIPPED_MEAN_Y = np.array([0.052229169532186454, 0.9236331507819656, -1.3859332319838136, -0.42092474514907724])
SHIPPED_SCALE_Y = np.array([0.8768847052996848, 0.6520665510519178, 1.0214536342625566, 0.4481255851377674])

REFERENCE_LAYER_DIMS = [3] + [500] * 12 + [4]


def synthetic_weights(seed=1234, dims=REFERENCE_LAYER_DIMS, gain=1.0, bias_range=0.1):
    """Seeded Glorot-uniform kernels (Keras' default initialiser, shape (in, out)) and small uniform biases."""
    rng = np.random.default_rng(seed)
    kernels, biases = [], []
    for fan_in, fan_out in zip(dims[:-1], dims[1:]):
        lim = gain * np.sqrt(6.0 / (fan_in + fan_out))
        kernels.append(rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32))
        biases.append(rng.uniform(-bias_range, bias_range, size=(fan_out,)).astype(np.float32))
    return kernels, biases
