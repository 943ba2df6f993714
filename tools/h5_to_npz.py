#!/usr/bin/env python
"""Convert a Keras `.h5` model (reference kinematics/ann.py:78-85, e.g. models/roboarm_model_1674153800-982793.h5)
into the flat `<name>.npz` container (W0.., b0.. = the Dense kernels (in, out) and biases in layer order) that
`ANN.load_model` reads where h5py is not installed.  Needs h5py (not keras); run it once where that exists:

    python tools/h5_to_npz.py path/to/model.h5 [out.npz]

The scaler files `<name>_scaler_x.bin` / `<name>_scaler_y.bin` are used as they are.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(argv):
    if len(argv) not in (2, 3):
        sys.exit(__doc__)
    from inversekinematicsann_b200.kinematics.ann import DenseStack
    src = argv[1]
    dst = argv[2] if len(argv) == 3 else os.path.splitext(src)[0] + ".npz"
    if not DenseStack.is_hdf5(src):
        sys.exit(f"{src} is not an HDF5 file")
    try:
        stack = DenseStack.load_h5(src)
    except ImportError:
        sys.exit("h5py is not installed here; run this converter where it is")
    stack.save_npz(dst)
    print(f"{dst}: {len(stack.kernels)} Dense layers, dims {stack.layer_dims}")


if __name__ == "__main__":
    main(sys.argv)
