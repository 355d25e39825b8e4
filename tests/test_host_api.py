"""Host-side logic and the C-ABI surface.  CPU only: no compute call is made without a GPU."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import REPO, load_case, load_weights, load_config
from oracle import pesto_oracle as O
from pesto_b200 import _lib
from pesto_b200.dataset import collate_batch_features
from pesto_b200.model import Model
from pesto_b200.sharding import lpt_partition
from pesto_b200.synth import synth_structure, one_hot_features, dense_membership, interfaceome_sizes


@pytest.fixture(scope="module")
def library():
    if not os.path.exists(_lib.LIB_PATH):
        from pesto_b200.build import build_library
        build_library()
    return _lib.load()


def test_library_exports_every_declared_symbol(library):
    header = open(os.path.join(REPO, "include", "pesto_b200.h")).read()
    declared = set(re.findall(r"\b(pesto_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in the header"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(library, name), f"libpesto_b200.so does not export {name}"
    assert library.pesto_abi_version() == 1


def test_size_queries_need_no_gpu(library):
    assert library.pesto_forward_workspace_bytes(8192, 1024) > 8192 * 128 * 4 * 2
    assert library.pesto_knn_scratch_bytes(100, 3) >= 100 * 4
    assert library.pesto_forward_workspace_bytes(0, 0) == 0


def test_model_create_rejects_bad_layers(library):
    nn = np.array([8, 12], dtype=np.int32)
    assert not library.pesto_model_create(2, nn.ctypes.data, 30)
    assert b"nn=12" in library.pesto_last_error()


@pytest.mark.parametrize("tag", ["i_v4_1", "i_v4_0"])
def test_model_accepts_reference_state_dict(tag):
    """load_state_dict with the checkpoint's own keys, strict."""
    m = Model(load_config(tag))
    sd = {k: torch.from_numpy(v) for k, v in load_weights(tag).items()}
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert len(sd) == {"i_v4_1": 1081, "i_v4_0": 553}[tag]
    if tag == "i_v4_1":      # parameter count printed by the reference's training log (slurm-731740.out:1387)
        assert sum(p.numel() for p in m.parameters()) == 1474957


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    m = Model(load_config("i_v4_0"))
    X, el, rid = synth_structure(64, 1)
    with pytest.raises(_lib.PestoError):
        m(X, torch.zeros((64, 64), dtype=torch.long), one_hot_features(el), dense_membership(rid))
    from pesto_b200.data_encoding import extract_topology
    with pytest.raises(_lib.PestoError):
        extract_topology(X, 64)


def test_collate_matches_oracle_and_golden():
    c = load_case("batch3")
    sizes = c["sizes"]
    off = np.concatenate([[0], np.cumsum(sizes)])
    X = torch.from_numpy(c["X"])
    rid = torch.from_numpy(c["rid"].astype(np.int64))
    el = torch.from_numpy(c["el"].astype(np.int64))
    parts, oparts = [], []
    r0 = 0
    for i in range(len(sizes)):
        sl = slice(off[i], off[i + 1])
        r_local = rid[sl] - r0
        n_res = int(r_local.max()) + 1
        ids0 = torch.from_numpy(c[f"ids0_{i}"]).long()
        parts.append([X[sl], ids0, one_hot_features(el[sl]), dense_membership(r_local, n_res)])
        oparts.append((X[sl], ids0, one_hot_features(el[sl]), r_local, n_res))
        r0 += n_res
    Xc, ids1, q, M = collate_batch_features(parts)
    assert torch.equal(ids1, torch.from_numpy(c["ids1"]).long())          # the reference's collate output
    assert torch.equal(M.argmax(1), rid) and M.shape == (X.shape[0], int(c["n_res"]))
    assert torch.equal(M.sum(1), torch.ones(X.shape[0]))
    Xo, ids1o, qo, rido, R = O.collate(oparts)
    assert torch.equal(ids1o, ids1) and torch.equal(rido, rid) and R == int(c["n_res"]) and torch.equal(Xo, Xc)
    _, _, _, Ms = collate_batch_features(parts, sparse_membership=True)
    assert torch.equal(Ms.long(), rid)


def test_synth_is_deterministic_and_protein_like():
    X1, el1, rid1 = synth_structure(1024, 20230419)
    X2, el2, rid2 = synth_structure(1024, 20230419)
    assert torch.equal(X1, X2) and torch.equal(el1, el2) and torch.equal(rid1, rid2)
    d = torch.cdist(X1.double(), X1.double())
    d.fill_diagonal_(1e9)
    assert d.min().item() > 0.05                     # nothing inside the 1e-2 mask of extract_topology
    nn64 = d.sort(1)[0][:, 63].mean().item()
    assert 6.0 < nn64 < 9.0                          # real structures: 7.6 A to the 64th neighbour
    sizes = interfaceome_sizes(1000)
    assert sizes.min() >= 16 and sizes.max() <= 2700


def test_lpt_partition_balances_cost():
    rng = np.random.default_rng(0)
    costs = rng.integers(100, 20000, size=500)
    for g in (1, 2, 4, 8):
        shards = lpt_partition(costs, g)
        assert sorted(i for s in shards for i in s) == list(range(500))
        loads = [sum(int(costs[i]) for i in s) for s in shards]
        assert max(loads) <= 1.02 * (sum(loads) / g) + costs.max()


def test_runner_packing_and_encoding():
    """host side of the structure runner: greedy packing under the atom budget, residue index = rank of the resid among
    the structure's sorted unique resids (src/data_encoding.py:73) offset per structure, element one-hot [N,30]."""
    import numpy as np
    from pesto_b200.runner import encode_batch, pack_batches
    assert pack_batches([500, 500, 500], 1000) == [[0, 1], [2]]
    assert pack_batches([5000], 1000) == [[0]] and pack_batches([], 10) == []
    # fewer atoms than neighbours: sink-padded slots read X[-1] of the batch, so such a structure is batched alone
    assert pack_batches([500, 63, 64, 500], 2000) == [[0], [1], [2, 3]]
    assert pack_batches([5, 5, 5], 10, num_nn=4) == [[0, 1], [2]]
    assert pack_batches([600, 300, 200], 1000) == [[0, 1, 2]]              # a tail below target / 4 rides along
    a = {"xyz": np.zeros((4, 3)), "element": np.array(["C", "N", "Zz", "O"]), "resid": np.array([7, 7, 3, 9])}
    b = {"xyz": np.ones((2, 3)), "element": np.array(["S", "C"]), "resid": np.array([1, 2])}
    X, q0, rid, n_at, n_rs = encode_batch([a, b])
    assert X.shape == (6, 3) and X.dtype == np.float32 and q0.shape == (6, 30)
    assert list(q0.argmax(1)) == [0, 2, 29, 1, 3, 0]
    assert list(rid) == [1, 1, 0, 2, 3, 4] and rid.dtype == np.int32 and n_at == [4, 2] and n_rs == [3, 2]
    from pesto_b200.data_encoding import onehot, std_elements
    from pesto_b200.runner import element_index
    els = np.array(list(std_elements) + ["X", "", "Zz", "H", "c"])
    assert list(element_index(els)) == list(onehot(els, std_elements).argmax(1))
    assert list(encode_batch([a, b], as_index=True)[1]) == [0, 2, 29, 1, 3, 0]


def test_empty_structure_is_rejected_before_any_device_work():
    import pytest
    import torch
    from pesto_b200.data_encoding import extract_topology
    with pytest.raises(ValueError):
        extract_topology(torch.zeros((0, 3)), 64)
    with pytest.raises(ValueError):
        extract_topology(torch.zeros((5, 2)), 64)


def test_feature_index_of_the_v3_models_and_head_depths():
    """123-feature q0 (element | residue name | atom name one-hots) travels as three bytes per atom; a Model built for a
    checkpoint takes the head depths from its keys (i_v3_1: single Linear em / dm, one logit)."""
    from conftest import load_case, load_config, load_weights
    from pesto_b200.data_encoding import encode_features
    from pesto_b200.model import Model
    from pesto_b200.runner import encode_batch, expand_features, feature_index
    c = load_case("v3_1gpw_A")
    s = {k: c[k] for k in ("element", "resname", "name", "resid")}
    s["xyz"] = c["X"]
    idx = feature_index(s, 123)
    assert idx.shape == (len(c["X"]), 3) and idx.dtype == np.uint8 and np.array_equal(idx, c["feat"])
    q0 = torch.cat(encode_features(s), dim=1)
    assert q0.shape[1] == 123 and torch.equal(expand_features(torch.from_numpy(idx), 123), q0)
    assert np.array_equal(encode_batch([s], n_features=123)[1], q0.numpy())
    assert feature_index(s, 30).shape == (len(c["X"]), 1)
    with pytest.raises(ValueError):
        feature_index(s, 64)
    for tag, depth, n_out in (("i_v3_0", 3, 5), ("i_v3_1", 1, 1)):
        sd = {k: torch.from_numpy(v) for k, v in load_weights(tag).items()}
        m = Model.for_state_dict(load_config(tag), sd)                     # strict load: the key sets agree
        assert len(m.em) == (5 if depth == 3 else 1) and len(m.dm) == (5 if depth == 3 else 1) and m.num_out == n_out
    with pytest.raises(RuntimeError):
        Model(load_config("i_v3_1")).load_state_dict({k: torch.from_numpy(v) for k, v in load_weights("i_v3_1").items()})
