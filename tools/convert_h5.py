"""Convert a Keras `.h5` weight file of the reference into the `.npz` (Keras names) that CChessModel.save writes.
  python tools/convert_h5.py /root/reference/data/model/model_best_weight.h5 tests/golden/_local/model_best_192x10.npz"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import cczero_b200  # noqa: F401
from cczero_b200.keras_h5 import read_keras_weights

if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    w = read_keras_weights(src)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    np.savez(dst, **{k.replace("/", "__"): v for k, v in w.items()})
    print(len(w), "tensors,", sum(v.size for v in w.values()), "parameters ->", dst)
