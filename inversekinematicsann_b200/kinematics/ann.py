"""Neural-network IK -- inference API of reference kinematics/ann.py, computed by csrc/mlp.cu.

``ANN.predict`` = ``y_scaler.inverse_transform(model.predict(x_scaler.transform(points)))``
(ann.py:70-76) as ONE fused kernel.  ``load_model`` keeps the reference's file convention
(ann.py:78-85): ``<name>.h5`` plus ``<name>_scaler_x.bin`` / ``<name>_scaler_y.bin`` (joblib
pickles of sklearn StandardScaler).  Keras/TensorFlow and h5py are optional: when h5py is missing
the Dense kernels are read from the flat sibling ``<name>.npz`` (keys W0.., b0..) that ``save_model``
writes and ``tools/h5_to_npz.py`` produces.  ``train_model`` (ann.py:27-68) lives in training.py.
"""
import os
from datetime import datetime

import numpy as np

from ._shared import get_engine


class DenseStack:
    """The Sequential model as plain arrays: kernels[l] has shape (in, out) like Keras Dense."""

    def __init__(self, kernels, biases):
        self.kernels = [np.asarray(k, dtype=np.float32) for k in kernels]
        self.biases = [np.asarray(b, dtype=np.float32) for b in biases]

    @property
    def layer_dims(self):
        return [self.kernels[0].shape[0]] + [k.shape[1] for k in self.kernels]

    def save_npz(self, path):
        arrays = {}
        for i, (k, b) in enumerate(zip(self.kernels, self.biases)):
            arrays[f'W{i}'] = k
            arrays[f'b{i}'] = b
        np.savez(path, **arrays)

    @staticmethod
    def is_hdf5(path):
        """True for a real HDF5 file; `<name>.h5` written by save_model without h5py holds the npz container."""
        with open(path, 'rb') as f:
            return f.read(8) == b'\x89HDF\r\n\x1a\n'

    def save_h5(self, path):
        """Keras' legacy HDF5 layout (model_weights/<layer>/<layer>/{kernel:0,bias:0} + layer_names), needs h5py."""
        import h5py  # optional dependency
        with h5py.File(path, 'w') as f:
            group = f.create_group('model_weights')
            names = [f'dense_{i}' if i else 'dense' for i in range(len(self.kernels))]
            group.attrs['layer_names'] = [n.encode() for n in names]
            for name, k, b in zip(names, self.kernels, self.biases):
                inner = group.create_group(name).create_group(name)
                inner.create_dataset('kernel:0', data=k)
                inner.create_dataset('bias:0', data=b)

    @classmethod
    def load_npz(cls, path):
        data = np.load(path)
        n = len([k for k in data.files if k.startswith('W')])
        return cls([data[f'W{i}'] for i in range(n)], [data[f'b{i}'] for i in range(n)])

    @classmethod
    def load_h5(cls, path):
        import h5py  # optional dependency
        kernels, biases = [], []
        with h5py.File(path, 'r') as f:
            group = f['model_weights'] if 'model_weights' in f else f
            names = [n.decode() if isinstance(n, bytes) else n for n in group.attrs.get('layer_names', list(group))]
            for name in names:
                found = {}
                group[name].visititems(lambda key, obj: found.__setitem__(key.split('/')[-1], np.array(obj))
                                       if hasattr(obj, 'shape') else None)
                kern = next((v for k, v in found.items() if k.startswith('kernel')), None)
                bias = next((v for k, v in found.items() if k.startswith('bias')), None)
                if kern is not None:
                    kernels.append(kern)
                    biases.append(bias if bias is not None else np.zeros(kern.shape[1], np.float32))
        return cls(kernels, biases)


class _ScalerView:
    """Minimal stand-in with the two StandardScaler attributes the engine needs."""

    def __init__(self, mean, scale):
        self.mean_ = np.asarray(mean, dtype=np.float64)
        self.scale_ = np.asarray(scale, dtype=np.float64)


class ANN:
    """Neural-network IK approach (reference ann.py:18-95), inference only."""

    def __init__(self, effector_workspace_limits, dh_matrix, device=None):
        self.effector_workspace_limits = effector_workspace_limits
        self.dh_matrix = dh_matrix
        self.model = None
        self.x_data_skaler = None  # attribute names as upstream (ann.py:24-25)
        self.y_data_skaler = None
        # 'fp16x3_ts' (default) / 'fp16x3': tcgen05 split-fp16 tensor-core kernels; 'fp32': CUDA-core kernel
        self.mode = 'fp16x3_ts'
        self._device = device
        self._uploaded = False

    def _engine(self):
        return get_engine(dh_matrix=self.dh_matrix, workspace_limits=self.effector_workspace_limits,
                          device=self._device)

    def set_model(self, kernels, biases, mean_x, scale_x, mean_y, scale_y):
        """Install weights and scaler statistics directly (no files)."""
        self.model = DenseStack(kernels, biases)
        self.x_data_skaler = _ScalerView(mean_x, scale_x)
        self.y_data_skaler = _ScalerView(mean_y, scale_y)
        self._uploaded = False
        return self.model

    def load_model(self, model_h5):
        """Load model from file; scalers from `<name>_scaler_{x,y}.bin` (ann.py:78-85)."""
        from joblib import load
        modelname = model_h5[:-3]
        npz = modelname + '.npz'
        if os.path.exists(model_h5) and not DenseStack.is_hdf5(model_h5):
            self.model = DenseStack.load_npz(model_h5)   # written by save_model() where h5py is absent
        elif os.path.exists(model_h5):
            try:
                self.model = DenseStack.load_h5(model_h5)
            except ImportError:
                if not os.path.exists(npz):
                    raise ImportError(f'h5py is not installed and {npz} does not exist; convert the Keras '
                                      'file once with tools/h5_to_npz.py where h5py is available')
                self.model = DenseStack.load_npz(npz)
        elif os.path.exists(npz):
            self.model = DenseStack.load_npz(npz)
        else:
            raise FileNotFoundError(model_h5)
        self.x_data_skaler = load(f'{modelname}_scaler_x.bin')
        self.y_data_skaler = load(f'{modelname}_scaler_y.bin')
        self._uploaded = False
        return self.model

    def save_model(self, prefix='model'):
        """Save weights as `<prefix>_<timestamp>.h5` and the scalers next to it (ann.py:87-95).  With h5py the file
        is HDF5 in Keras' legacy weight layout; without it the same name holds the flat npz container (W0.., b0..),
        which `load_model` recognises by its magic bytes."""
        from joblib import dump
        stamp = str(datetime.timestamp(datetime.now())).replace('.', '-')
        try:
            self.model.save_h5(f'{prefix}_{stamp}.h5')
        except ImportError:
            with open(f'{prefix}_{stamp}.h5', 'wb') as f:
                self.model.save_npz(f)
        dump(self.x_data_skaler, f'{prefix}_{stamp}_scaler_x.bin', compress=True)
        dump(self.y_data_skaler, f'{prefix}_{stamp}_scaler_y.bin', compress=True)
        return f'{prefix}_{stamp}'

    def train_model(self, epochs, samples, features, **fit_options):
        """Train the Sequential model on (positions, joint angles) -- reference ann.py:40-68; see
        kinematics/training.py for the recipe and the extra `fit_options` (batch_size, learning_rate, device...)."""
        from . import training
        x_train, y_train, x_test, y_test, self.x_data_skaler, self.y_data_skaler = \
            training.fit_training_data(samples, features)
        kernels, biases, self.history = training.train_dense_stack(x_train, y_train, x_test, y_test, epochs,
                                                                   **fit_options)
        self.model = DenseStack(kernels, biases)
        self._uploaded = False
        return self.model

    def _ensure_uploaded(self):
        if self.model is None:
            raise RuntimeError('no model loaded: call load_model() or set_model() first')
        eng = self._engine()
        if not self._uploaded or getattr(eng, '_ann_owner', None) is not self:
            eng.mlp_load(self.model.kernels, self.model.biases, self.x_data_skaler.mean_,
                         self.x_data_skaler.scale_, self.y_data_skaler.mean_, self.y_data_skaler.scale_)
            eng._ann_owner = self
            self._uploaded = True
        return eng

    def predict_with_stats(self, position, out=None, return_fk_error=False):
        eng = self._ensure_uploaded()
        arr = np.asarray(position)
        if arr.dtype not in (np.float32, np.float64):
            arr = arr.astype(np.float64)
        return eng.ann_solve(arr.reshape(-1, 3), out=out, mode=self.mode, return_fk_error=return_fk_error,
                             fk_stats=return_fk_error)

    def predict(self, position):
        """Predict joint angles: (n, 4) float32 ndarray like Keras + sklearn return (ann.py:70-76).
        No workspace-limit check here, as upstream (tests/ann_unit.py:39 predicts an outside point)."""
        return self.predict_with_stats(position)[0]
