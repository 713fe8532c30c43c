"""Warm timing of the tape's elementwise work: sigmoid(Z) forward + reverse sweep on a rows x cols container, recorded
once as a CUDA graph and replayed (so kernels run back to back with warm instruction caches, as in a training step).
Prints microseconds per replay for the fused groups and for the one-kernel-per-node path (EML_FUSE=0).

    python scripts/ew_bench.py [rows cols [replays]]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import easy_ml_b200 as eml  # noqa: E402


def sigmoid(z):
    return ((-z).exp() + 1.0).div_swapped(1.0)


def run(fuse, rows, cols, replays, dtype=np.float32):
    os.environ["EML_FUSE"] = "1" if fuse else "0"
    hist = eml.WengertList(dtype)
    z = np.random.default_rng(0).uniform(-2, 2, (rows, cols)).astype(dtype)
    Z = eml.RecordTensor.variables(hist, eml.Tensor([("r", rows), ("c", cols)], z))
    pinned = eml.PinnedBuffer((1,), dtype)

    def body():
        y = sigmoid(Z)
        d = y.derivatives_for([0, 0])
        g = d.at_tensor  # noqa: F841  (the derivatives stay on the device)
        del y, d
        hist.clear()
        Z.reset()

    body()
    hist.sync()
    before = eml.kernel_launches()
    step = hist.record(body)
    launches = eml.kernel_launches() - before
    for _ in range(5):
        step.launch()
    hist.sync()
    t0 = time.perf_counter()
    for _ in range(replays):
        step.launch()
    hist.sync()
    us = (time.perf_counter() - t0) / replays * 1e6
    del pinned
    return us, launches


if __name__ == "__main__":
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    cols = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    replays = int(sys.argv[3]) if len(sys.argv) > 3 else 300
    for fuse in (True, False):
        us, launches = run(fuse, rows, cols, replays)
        traffic = rows * cols * 4 * (2 + 3)  # forward: read z, write y; sweep: read dy, z, write dz
        print(f"fuse={int(fuse)} {rows}x{cols} f32: {us:8.2f} us per forward+sweep, {launches} kernels, "
              f"{traffic / us / 1e6:.2f} TB/s of algorithmic traffic")
