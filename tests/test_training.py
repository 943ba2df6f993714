"""Training loop of the ANN solver (reference ann.py:27-68 recipe restated in kinematics/training.py)."""
import numpy as np
import pytest

from inversekinematicsann_b200.kinematics import training
from inversekinematicsann_b200.kinematics.ann import ANN
from inversekinematicsann_b200.robot.robot import SixDOFRobot as R
from oracle import np_oracle


def _forward(x_scaled, kernels, biases):
    return np_oracle.mlp_predict(x_scaled, kernels, biases, np.zeros(3), np.ones(3), np.zeros(4), np.ones(4))


def _toy_problem(n=3000, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-2, 2, (n, 3))
    y = np.stack([np.sin(x[:, 0]), x[:, 1] * 0.5, np.tanh(x[:, 2]), x[:, 0] * x[:, 1] * 0.1], axis=1)
    return x, y


def test_split_and_scalers_follow_the_reference():
    x, y = _toy_problem()
    xtr, ytr, xte, yte, xs, ys = training.fit_training_data(x, y)
    assert xtr.shape == (2010, 3) and xte.shape == (990, 3)  # test_size=0.33 (ann.py:30)
    from sklearn.model_selection import train_test_split
    want_train = train_test_split(x, y, test_size=0.33, random_state=42)[0]
    assert np.allclose(xs.mean_, want_train.mean(axis=0)) and np.allclose(xs.scale_, want_train.std(axis=0))
    assert np.allclose(xtr.mean(axis=0), 0, atol=1e-12) and np.allclose(ytr.std(axis=0), 1)
    assert np.allclose(ys.inverse_transform(yte), train_test_split(x, y, test_size=0.33, random_state=42)[3])


def test_train_model_learns_and_exports_keras_layout():
    x, y = _toy_problem()
    ann = ANN(R.effector_workspace_limits, R.dh_matrix)
    model = ann.train_model(30, x, y, batch_size=64, learning_rate=3e-3, hidden_layers=2, hidden_units=32,
                            device='cpu', seed=1, verbose=False)
    assert model.layer_dims == [3, 32, 32, 4]
    assert [k.shape for k in model.kernels] == [(3, 32), (32, 32), (32, 4)]
    hist = ann.history
    assert hist['val_loss'][-1] < 0.25 * hist['val_loss'][0]
    assert hist['best_val_loss'] == min(hist['val_loss'])
    # exported arrays reproduce the trained network: scaled inputs -> oracle forward -> val loss as recorded
    xtr, ytr, xte, yte, _, _ = training.fit_training_data(x, y)
    pred = _forward(xte, model.kernels, model.biases)
    assert abs(float(np.mean((pred - yte) ** 2)) - hist['best_val_loss']) < 1e-5
    assert type(ann.x_data_skaler).__name__ == 'StandardScaler'


def test_early_stopping_restores_best_weights():
    x, y = _toy_problem(600)
    xtr, ytr, xte, yte, _, _ = training.fit_training_data(x, y)
    # a huge step size makes the validation loss bounce, so patience=2 triggers well before 200 epochs
    kernels, biases, hist = training.train_dense_stack(xtr, ytr, xte, yte, 200, batch_size=32, learning_rate=0.3,
                                                       patience=2, hidden_layers=2, hidden_units=8, device='cpu',
                                                       seed=0, verbose=False)
    assert len(hist['val_loss']) < 200
    pred = _forward(xte, kernels, biases)
    assert abs(float(np.mean((pred - yte) ** 2)) - min(hist['val_loss'])) < 1e-5


@pytest.mark.gpu
def test_trained_model_runs_through_the_kernels():
    """train (torch) -> save_model/load_model files -> predict through csrc/mlp*.cu == oracle on the same weights."""
    from inversekinematicsann_b200.kinematics.inverse import FabrikInverseKinematics
    rng = np.random.default_rng(5)
    pts = (rng.uniform(0, 1, (20000, 3)) * [2, 4, 3] + [1, -2, 1]).astype(np.float32)
    fab = FabrikInverseKinematics(R.dh_matrix, R.links_lengths, R.effector_workspace_limits)
    labels = fab.ikine(pts, as_array=True)
    ann = ANN(R.effector_workspace_limits, R.dh_matrix)
    ann.train_model(3, pts, labels, batch_size=256, learning_rate=1e-3, hidden_layers=3, hidden_units=64,
                    verbose=False, seed=2)
    want = np_oracle.mlp_predict(pts[:4096], ann.model.kernels, ann.model.biases, ann.x_data_skaler.mean_,
                                 ann.x_data_skaler.scale_, ann.y_data_skaler.mean_, ann.y_data_skaler.scale_)
    for mode in ('fp32', 'fp16x3_ts'):
        ann.mode = mode
        got = ann.predict(pts[:4096])
        assert np.abs(got - want).max() <= 1e-5, mode
