"""The whole apply path on the GPU against the reference's shipped outputs (examples/issue_19_04_2023: the files the
reference wrote and their md5 sums; tests/golden/pdb, made by tests/golden/make_golden_pdb.py)."""
import gzip
import hashlib
import os

import pytest

from conftest import GOLDEN, BOTH_MODES, load_weights

pytestmark = pytest.mark.gpu
PDB = os.path.join(GOLDEN, "pdb")


def _gunzip_to(tmp_path, name):
    dst = os.path.join(tmp_path, name[:-3])
    with gzip.open(os.path.join(PDB, name), "rb") as fi, open(dst, "wb") as fo:
        fo.write(fi.read())
    return dst


def _bfactors(text):
    return [float(ln[60:66]) for ln in text.splitlines() if ln.startswith(("ATOM", "HETATM"))]


@pytest.mark.parametrize("mode", ["fp32", "f16x3"])
def test_apply_reproduces_shipped_output_files(tmp_path, cuda_models, mode):
    """PDB file -> C++ reader -> preprocessing -> CUDA kNN + forward -> sigmoid -> save_pdb for 2CUA_A: byte-identical
    files (md5check.txt) in fp32 mode; in f16x3 mode every record identical outside the two-decimal probability columns
    and every probability within one unit of the last printed digit."""
    from pesto_b200.apply import apply_to_pdb
    model = cuda_models("i_v4_1")
    src = _gunzip_to(str(tmp_path), "2CUA_A.pdb.gz")
    old = model.mode
    model.mode = mode
    try:
        paths, p = apply_to_pdb(model, src)
    finally:
        model.mode = old
    assert p.shape == (122, 5)
    md5 = dict(reversed(ln.split()) for ln in open(os.path.join(PDB, "md5check.txt")) if ln.strip())
    n_diff_lines = 0
    for i, path in enumerate(paths):
        got = open(path, "rb").read()
        exp = gzip.open(os.path.join(PDB, f"2CUA_A_i{i}.pdb.gz"), "rb").read()
        assert hashlib.md5(exp).hexdigest() == md5[f"2CUA_A_i{i}.pdb"]
        if mode == "fp32":
            assert hashlib.md5(got).hexdigest() == md5[f"2CUA_A_i{i}.pdb"], f"channel {i}"
        gl, el = got.decode().splitlines(), exp.decode().splitlines()
        assert len(gl) == len(el)
        for g, e in zip(gl, el):
            assert g[:54] == e[:54] and g[66:] == e[66:]
            n_diff_lines += g != e
        assert max(abs(a - b) for a, b in zip(_bfactors(got.decode()), _bfactors(exp.decode()))) <= 0.0100001
    assert n_diff_lines <= 4, n_diff_lines


@pytest.mark.parametrize("name", ["1ZNS", "7KHT_lipid"])
def test_apply_multichain_probabilities(tmp_path, cuda_models, name):
    """multi-chain (protein + DNA) and hetero (lipid) structures: the b-factor column of channel 0 matches the file the
    reference wrote to the printed precision (+-0.005 rounding, +-1e-3 logits)."""
    from pesto_b200.apply import apply_to_pdb
    model = cuda_models("i_v4_1")
    src = _gunzip_to(str(tmp_path), name + ".pdb.gz")
    paths, _ = apply_to_pdb(model, src)
    got = _bfactors(open(paths[0]).read())
    exp = [float(ln.split("|")[1]) for ln in gzip.open(os.path.join(PDB, name + "_i0.expected.gz"), "rt") if ln.startswith(("ATOM", "HETATM"))]
    assert len(got) == len(exp)
    assert max(abs(a - b) for a, b in zip(got, exp)) <= 0.0100001


@pytest.mark.parametrize("mode,n_atoms", [("fp32", 700), ("f16x3", 700), ("f16x3", 48)])
def test_trajectory_mode_equals_frame_by_frame(cuda_models, mode, n_atoms):
    """md_analysis/apply_model_md.ipynb cell 6: frame-0 topology reused for every frame; batching the frames as
    structures gives the logits of the frame-by-frame loop (also with a partial last batch) -- checked against the CPU
    ORACLE run frame by frame (stale frame-0 neighbour lists, src/model_operations.py:8 on the frame's coordinates), not
    against the model itself.  The 48-atom molecule has sink-padded neighbour slots, which read X[-1] of whatever is in
    the batch: such molecules must be run one frame per forward to keep the reference's numbers."""
    import torch
    from oracle import pesto_oracle as O
    from pesto_b200.data_encoding import extract_topology
    from pesto_b200.dataset import collate_batch_features
    from pesto_b200.md import predict_trajectory
    from pesto_b200.synth import synth_structure, one_hot_features, dense_membership
    model = cuda_models("i_v4_0", mode)
    tol = 2e-4 if mode == "fp32" else 3e-4
    X, el, rid = synth_structure(n_atoms, 99)
    g = torch.Generator().manual_seed(5)
    T = 7
    X_traj = X.unsqueeze(1) + 0.3 * torch.randn(X.shape[0], T, 3, generator=g)       # thermal motion around frame 0
    X_traj[:, 0] = X
    q, M = one_hot_features(el), dense_membership(rid)
    ids0 = extract_topology(X_traj[:, 0].cuda(), 64)[0]
    _, ids1, qc, Mc = collate_batch_features([[X_traj[:, 0].cuda(), ids0, q.cuda(), M.cuda()]])
    W = load_weights("i_v4_0")
    n_res = int(rid.max()) + 1
    ref = torch.stack([O.forward(W, X_traj[:, i].contiguous(), ids1.cpu(), q, rid, n_res) for i in range(T)])
    for per in (None, 3):
        z = predict_trajectory(model, X_traj, ids1, q, M, frames_per_batch=per).cpu()
        assert z.shape == ref.shape
        assert (z - ref).abs().max().item() <= tol
    z2 = predict_trajectory(model, X_traj.cuda(), ids1, q, rid, frames=[1, 5]).cpu()
    assert (z2 - ref[[1, 5]]).abs().max().item() <= tol


@pytest.mark.parametrize("mode", BOTH_MODES)
def test_structure_runner_equals_one_by_one(cuda_models, mode):
    """interfaceome/apply_model.py:49-82 pattern: packing structures into batches gives each structure the logits the CPU
    ORACLE computes for it alone (the reference's one-structure-at-a-time loop); batches respect the atom budget; order
    is preserved.  Structures with fewer than 64 atoms (sink-padded neighbour slots read X[-1] of the batch,
    src/model_operations.py:8) get a batch of their own, so they too keep the reference's numbers."""
    import numpy as np
    import torch
    from oracle import pesto_oracle as O
    from pesto_b200.data_encoding import std_elements
    from pesto_b200.runner import pack_batches, predict_structures
    from pesto_b200.synth import synth_structure, one_hot_features
    model = cuda_models("i_v4_0", mode)
    W = load_weights("i_v4_0")
    sizes = [300, 70, 900, 40, 513, 64, 1200]
    structures, singles = [], []
    for k, n in enumerate(sizes):
        X, el, rid = synth_structure(n, 1000 + k)
        structures.append({"xyz": X.numpy(), "element": std_elements[el.numpy()], "resid": rid.numpy() * 3 + 7})
        q0 = one_hot_features(el)
        n_res = int(rid.max()) + 1
        ids1 = O.collate([(X, O.extract_topology(X, 64)[0], q0, rid, n_res)])[1]
        singles.append(O.forward(W, X, ids1, q0, rid, n_res))
    assert pack_batches(sizes, 1300) == [[0, 1, 2], [3], [4, 5], [6]]
    got = list(predict_structures(model, structures, target_atoms=1300))
    assert [i for i, _ in got] == list(range(len(sizes)))
    for (i, z), ref in zip(got, singles):
        assert z.shape == ref.shape and (z - ref).abs().max().item() <= (2e-4 if mode == "fp32" else 3e-4), i


def test_one_process_two_devices():
    """One process, two GPUs: the dynamic shared-memory opt-in and the SM count of the persistent grids are per-device
    attributes (csrc `device_setup`), the model keeps one packed handle and one workspace per device, the runner's staging
    buffers are per device.  Skipped on a single-GPU box."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from conftest import load_config
    from pesto_b200.data_encoding import extract_topology
    from pesto_b200.model import Model
    from pesto_b200.runner import predict_structures
    from pesto_b200.synth import synth_structure, one_hot_features
    sd = {k: torch.from_numpy(v) for k, v in load_weights("i_v4_0").items()}
    model = Model.for_state_dict(load_config("i_v4_0"), sd, mode="f16x3").eval()
    X, el, rid = synth_structure(700, 33)
    q0, n_res = one_hot_features(el), int(rid.max()) + 1
    zs = []
    for d in (1, 0, 1):                                   # the second device first: nothing may be cached from device 0
        dev = torch.device("cuda", d)
        Xd = X.to(dev)
        ids1 = extract_topology(Xd, 64)[0] + 1
        z = model(Xd, ids1, q0.to(dev), rid.int().to(dev), n_res=n_res)
        model.raise_if_failed(dev)
        assert z.device == dev
        zs.append(z.cpu())
    assert torch.equal(zs[0], zs[2]) and (zs[0] - zs[1]).abs().max().item() < 1e-5
    elems = np.asarray(["C", "N", "O", "S", "H"] * 6)
    s = {"xyz": X.numpy(), "element": elems[np.minimum(el.numpy(), len(elems) - 1)], "resid": rid.numpy()}
    out = [list(predict_structures(model, [s, s], device=f"cuda:{d}")) for d in (0, 1)]
    assert all(len(o) == 2 for o in out) and (out[0][0][1] - out[1][0][1]).abs().max().item() < 1e-5
