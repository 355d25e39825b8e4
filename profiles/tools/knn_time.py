#!/usr/bin/env python
"""Device time of the topology (kNN) launch sequence: the bench batch (53 structures, ~132 k atoms), one 8 192-atom and one
32 768-atom structure.    usage (under gpurun): python profiles/tools/knn_time.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench                                                                       # noqa: E402
from pesto_b200.data_encoding import batch_topology                                # noqa: E402


def timed(X, sizes, reps=10):
    for _ in range(3):
        batch_topology(X, sizes, 64)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        batch_topology(X, sizes, 64)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    from pesto_b200.synth import synth_structure
    wl = bench.load_workload()
    sizes = [int(n) for n in wl["sizes"]]
    X = wl["X"].cuda()
    print(f"batch {len(sizes)} structures, {X.shape[0]} atoms: {timed(X, sizes):.3f} ms")
    for n in (8192, 32768):
        Xn = synth_structure(n, 7)[0].cuda()
        print(f"one structure, {n} atoms: {timed(Xn, [n]):.3f} ms")


if __name__ == "__main__":
    main()
