"""Deterministic protein-like point clouds for measurement.

Pure host-side numpy/torch-CPU; used by bench.py, the tests and the golden-vector
generator so that every party sees bit-identical inputs for a given (N, seed).

Statistics match the reference's `pdbs_test/` structures: ~8 atoms per residue,
nearest-neighbour #1/#64 at ~1.5/7.2 A, element mix C/O/N/S, coordinates rounded
to PDB precision (3 decimals) so no pair is closer than the 1e-2 A mask of
`src/data_encoding.py:93`.
"""
import math

import numpy as np
import torch

BASE_SEED = 20230419
_ELEMENT_FREQ = (0.634, 0.189, 0.171, 0.005)     # std_elements[:4] = C, O, N, S
N_ELEMENT_CLASSES = 30                           # 29 std_elements + unknown


def synth_structure(n_atoms, seed=BASE_SEED):
    """Return (X[N,3] f32, el[N] int64 element class, rid[N] int64 residue index).

    Residue centres follow a snake (boustrophedon) walk through a cubic lattice with
    5.5 A spacing; each residue places 8 atoms on the corners of a 1.7 A cube plus
    N(0, 0.15 A) jitter.
    """
    g = torch.Generator().manual_seed(int(seed))
    n_res = math.ceil(n_atoms / 8)
    side = math.ceil(round(n_res ** (1.0 / 3.0), 9))
    while side ** 3 < n_res:
        side += 1
    idx = torch.arange(n_res)
    z = idx // (side * side)
    rem = idx % (side * side)
    y = rem // side
    x = rem % side
    y = torch.where(z % 2 == 1, side - 1 - y, y)
    x = torch.where(y % 2 == 1, side - 1 - x, x)
    centres = torch.stack([x, y, z], 1).float() * 5.5
    corners = torch.tensor([[sx, sy, sz] for sx in (-.85, .85) for sy in (-.85, .85) for sz in (-.85, .85)])
    X = (centres[:, None, :] + corners[None]).reshape(-1, 3)[:n_atoms]
    X = X + 0.15 * torch.randn(X.shape, generator=g)
    X = (torch.round(X * 1000) / 1000).float().contiguous()
    el = torch.multinomial(torch.tensor(_ELEMENT_FREQ), n_atoms, True, generator=g)
    rid = torch.arange(n_atoms) // 8
    return X, el, rid


def one_hot_features(el, n_classes=N_ELEMENT_CLASSES):
    """q0[N, n_classes] f32 one-hot, as `encode_features(...)[0]` gives for v4 models."""
    return torch.nn.functional.one_hot(el.long(), n_classes).float()


def dense_membership(rid, n_res=None):
    """M[N, R] f32 one-hot rows, as `encode_structure` builds from resids."""
    n_res = int(rid.max()) + 1 if n_res is None else n_res
    return torch.nn.functional.one_hot(rid.long(), n_res).float()


def interfaceome_sizes(n_structures, seed=BASE_SEED):
    """Residue counts for BASELINE config 5: clip(round(exp(N(5.8, 0.7))), 16, 2700)."""
    rng = np.random.default_rng(seed)
    n_res = np.clip(np.round(np.exp(rng.normal(5.8, 0.7, size=n_structures))), 16, 2700).astype(np.int64)
    return n_res
