"""Generates tests/golden/nets_golden.npz: outputs of the fp32 PyTorch oracle networks (oracle/nets.py,
seeded weights 2 / 3) on seeded inputs.  The reference ships no network artefact to pin against
(TensorRT engines built from ONNX files that are not in the tree, README.md:72), so these vectors pin
the ORACLE against drift (PyTorch version, thread count) and give the CUDA path a committed target.

  python tests/golden/make_nets_golden.py        # CPU, ~1 min
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import fp_helpers as helpers  # noqa: E402
from oracle import nets  # noqa: E402


def golden_inputs(n: int = 3):
    """Seeded [n,160,160,6] blobs in the reference's layout: rgb in [0,1], normalised xyz in [-1,1],
    rounded to fp16 (what the packed network input holds)."""
    g = np.random.RandomState(123)
    a = g.rand(n, 160, 160, 6).astype(np.float32)
    b = g.rand(n, 160, 160, 6).astype(np.float32)
    a[..., 3:] = a[..., 3:] * 2 - 1
    b[..., 3:] = b[..., 3:] * 2 - 1
    for i in range(n):  # make the hypotheses differ: contrast, an empty quadrant, a shifted observation
        a[i] *= 0.4 + 0.3 * i
        a[i, : 40 * (i + 1), : 60 * i] = 0
        b[i] = np.roll(b[i], 7 * i, axis=1) * (1.0 - 0.25 * i)
    return a.astype(np.float16).astype(np.float32), b.astype(np.float16).astype(np.float32)


def main():
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    refiner, scorer = helpers.seeded_nets()
    a, b = golden_inputs()
    with torch.no_grad():
        ta, tb = nets.nhwc_to_inputs(torch.from_numpy(a)), nets.nhwc_to_inputs(torch.from_numpy(b))
        trans, rot = refiner(ta, tb)
        scores = scorer(ta, tb)
    np.savez_compressed(os.path.join(HERE, "nets_golden.npz"), trans=trans.numpy(), rot=rot.numpy(),
                        scores=scores.numpy().reshape(-1), input_checksum=np.float64(a.sum() + 2 * b.sum()))
    print("trans", trans.numpy(), "\nrot", rot.numpy(), "\nscores", scores.numpy().reshape(-1))


if __name__ == "__main__":
    main()
