"""Fixture generator (build container only): the reference's shipped, trained 192x10 network
(/root/reference/data/model/model_best_weight.h5, loaded by agent/model.py:95-107 through Keras) converted tensor for
tensor into tests/golden/model_best_192x10.npz — float32, bit-identical values, Keras weight names with '/' -> '__'.

The GPU box has neither /root/reference nor h5py, so the real-weight parity test (tests/test_keras_h5.py) reads this
file.  TEST INFRASTRUCTURE: nothing in the product package or bench.py's GPU arm reads it.

    python -m oracle.gen_golden_weights
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "model_best_192x10.npz")


def main():
    sys.path.insert(0, ROOT)
    from oracle import ref_import
    h5 = os.path.join(ref_import.REF_ROOT, "data", "model", "model_best_weight.h5")
    if not os.path.exists(h5):
        raise SystemExit("reference weights not present: " + h5)
    from cczero_b200.keras_h5 import read_keras_weights
    w = read_keras_weights(h5)
    assert len(w) == 121 and sum(v.size for v in w.values()) == 7519663
    np.savez_compressed(OUT, **{k.replace("/", "__"): np.ascontiguousarray(v, dtype=np.float32) for k, v in w.items()})
    print(OUT, os.path.getsize(OUT), "bytes,", len(w), "tensors")


if __name__ == "__main__":
    main()
