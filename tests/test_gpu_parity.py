"""Parity of the CUDA path (through the C ABI) with the reference: golden vectors made by the unmodified
reference, the CPU oracle on fresh seeded inputs, and size-independent properties at full size.

Tolerances: kNN indices bit-exact (modulo permutations inside exact-distance tie groups, whose order torch.topk
leaves unspecified); logits max-abs <= 1e-3 as BASELINE.json's north_star states (fp32 mode lands at ~1e-5).
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, BOTH_MODES, load_case, load_config, load_weights, case_tensors
from oracle import pesto_oracle as O
from oracle import scoring
from pesto_b200.synth import synth_structure, one_hot_features, dense_membership, BASE_SEED

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-3          # north_star tolerance
FP32_EXPECTED = 2e-4      # what the fp32 path should actually achieve (fp32 re-association noise is ~1e-5)


def cuda(*ts):
    return [t.cuda() for t in ts]


def run_case(model, c, mode="fp32", dense=True):
    from pesto_b200.dataset import collate_batch_features
    X, el, rid, n_res = case_tensors(c)
    q0 = one_hot_features(el)
    if "ids1" in c:
        ids1 = torch.from_numpy(c["ids1"]).long()
    else:
        ids1 = collate_batch_features([[X, torch.from_numpy(c["ids0"]).long(), q0, dense_membership(rid, n_res)]])[1]
    Xd, idsd, qd = cuda(X, ids1, q0)
    if dense:
        return model(Xd, idsd, qd, dense_membership(rid, n_res).cuda(), mode=mode)
    return model(Xd, idsd, qd, rid.int().cuda(), n_res=n_res, mode=mode)


# ------------------------------------------------------------------------------------------------- topology
@pytest.mark.parametrize("name", ["2CUA_A", "1gpw_A", "1EWY", "tiny40", "synth517"])
def test_knn_matches_reference(name):
    from pesto_b200.data_encoding import extract_topology
    c = load_case(name)
    X = torch.from_numpy(c["X"])
    ids, d, r, D, R = extract_topology(X.cuda(), 64)
    ref = torch.from_numpy(c["ids0"]).long()
    assert ids.dtype == torch.int64 and tuple(ids.shape) == tuple(ref.shape) and D is None and R is None
    assert O.same_modulo_ties(ids.cpu(), ref, X)
    # D_topk / R_topk: bit-exact against the oracle's dense recipe
    oi, od, orr = O.extract_topology(X, 64)
    assert torch.equal(ids.cpu(), oi)                       # same canonical (distance, index) order
    assert torch.equal(d.cpu(), od)
    assert torch.equal(r.cpu(), orr)


def test_knn_synth8192_bit_exact():
    from pesto_b200.data_encoding import extract_topology
    g = dict(np.load(os.path.join(GOLDEN, "case_synth8192.npz")))
    X, _, _ = synth_structure(8192, BASE_SEED)
    ids = extract_topology(X.cuda(), 64)[0].cpu()
    ref = torch.from_numpy(g["ids0"].astype(np.int64))
    assert O.same_modulo_ties(ids, ref, X)
    assert (ids == ref).float().mean().item() > 0.9999


def test_knn_cpu_tensor_round_trip_and_small_k():
    from pesto_b200.data_encoding import extract_topology
    X, _, _ = synth_structure(300, 5)
    ids_cpu = extract_topology(X, 16)[0]                     # CPU tensor in -> staged through the GPU -> CPU out
    assert ids_cpu.device.type == "cpu" and ids_cpu.shape == (300, 16)
    assert torch.equal(ids_cpu, O.extract_topology(X, 16)[0])


def test_knn_duplicate_atoms_and_tiny_structures():
    """Masked entries (D < 1e-2 -> + max D): duplicates and structures with <= 64 atoms exercise the exact path."""
    from pesto_b200.data_encoding import extract_topology
    X, _, _ = synth_structure(200, 11)
    X[17] = X[3]                     # exact duplicate
    X[50] = X[51] + 0.004            # closer than the 1e-2 mask
    for n in (200, 66, 65, 64, 2, 1):
        Xn = X[:n].contiguous()
        ids, d, r, _, _ = extract_topology(Xn.cuda(), 64)
        oi, od, orr = O.extract_topology(Xn, 64)
        assert torch.equal(ids.cpu(), oi), n
        assert torch.equal(d.cpu(), od), n
        assert torch.equal(torch.nan_to_num(r.cpu()), torch.nan_to_num(orr)), n


def test_knn_box_pruning_is_exact_for_any_atom_order():
    """The chunk bounding boxes only prune: shuffled atoms (every box spans the structure), a chain folded back on itself
    and a far-away second domain give the oracle's rows bit for bit; structure boundaries fall inside 32-atom chunks."""
    from pesto_b200.data_encoding import batch_topology, extract_topology
    X, _, _ = synth_structure(1500, 21)
    g = torch.Generator().manual_seed(3)
    shuffled = X[torch.randperm(1500, generator=g)].contiguous()
    far = torch.cat([X[:700], X[700:] + 500.0]).contiguous()                 # second domain 500 A away: its chunks are skipped
    folded = torch.cat([X[:750], X[:750].flip(0) + 0.37]).contiguous()       # index-distant atoms are the nearest ones
    for name, Xn in (("shuffled", shuffled), ("far", far), ("folded", folded)):
        ids, d, r, _, _ = extract_topology(Xn.cuda(), 64)
        oi, od, orr = O.extract_topology(Xn, 64)
        assert torch.equal(ids.cpu(), oi), name
        assert torch.equal(d.cpu(), od), name
        assert torch.equal(r.cpu(), orr), name
    sizes = [333, 70, 1097]                                                    # boundaries inside chunks
    ids1 = batch_topology(shuffled.cuda(), sizes, 64).cpu()
    off = np.concatenate([[0], np.cumsum(sizes)])
    parts = [(shuffled[off[i]:off[i + 1]], O.extract_topology(shuffled[off[i]:off[i + 1]], 64)[0], torch.zeros((n, 1)),
              torch.zeros(n, dtype=torch.long), 1) for i, n in enumerate(sizes)]
    assert torch.equal(ids1, O.collate(parts)[1])


def test_knn_large_structure_against_brute_force():
    """32 768 atoms (1 024 chunk boxes, 32 box groups per query): the selected neighbours' distances equal the 64 smallest of
    a brute-force distance matrix computed block by block on the device (fp64), row by row."""
    from pesto_b200.data_encoding import extract_topology
    X, _, _ = synth_structure(32768, BASE_SEED + 9)
    Xd = X.cuda()
    ids, d, _, _, _ = extract_topology(Xd, 64)
    assert ids.shape == (32768, 64) and int(ids.min()) >= 0 and int(ids.max()) < 32768
    assert bool((ids != torch.arange(32768, device="cuda").unsqueeze(1)).all())          # the atom itself is masked (D < 1e-2)
    assert bool((d[:, 1:] >= d[:, :-1]).all())                                            # sorted by distance
    X64 = Xd.double()
    for r0 in range(0, 32768, 4096):
        D = torch.cdist(X64[r0:r0 + 4096], X64)                                           # [4096, 32768] fp64
        D[torch.arange(4096, device="cuda"), torch.arange(r0, r0 + 4096, device="cuda")] = float("inf")
        best = torch.topk(D, 64, dim=1, largest=False).values
        mine = torch.gather(D, 1, ids[r0:r0 + 4096])
        assert (mine - best).abs().max().item() < 1e-4, r0
        assert (d[r0:r0 + 4096].double() - best).abs().max().item() < 1e-4, r0


def test_batch_topology_equals_collate_of_per_structure_knn():
    """One launch over a batch of structures == per-structure topology + collate (index shift, sink padding)."""
    from pesto_b200.data_encoding import batch_topology
    c = load_case("batch3")
    sizes = [int(v) for v in c["sizes"]]
    X = torch.from_numpy(c["X"])
    ids1 = batch_topology(X.cuda(), sizes, 64).cpu()
    off = np.concatenate([[0], np.cumsum(sizes)])
    parts = []
    for i, n in enumerate(sizes):
        Xi = X[off[i]:off[i + 1]]
        oi = O.extract_topology(Xi, 64)[0]
        parts.append((Xi, oi, torch.zeros((n, 1)), torch.zeros(n, dtype=torch.long), 1))
        # against the reference's own per-structure topk (tie order unspecified there)
        mine = ids1[off[i]:off[i + 1], :min(64, n)] - (off[i] + 1)
        assert O.same_modulo_ties(mine, torch.from_numpy(c[f"ids0_{i}"]).long(), Xi)
        assert bool((ids1[off[i]:off[i + 1], min(64, n):] == 0).all())            # sink padding
    assert torch.equal(ids1, O.collate(parts)[1])                                     # canonical order: bit-exact


# ------------------------------------------------------------------------------------------------- forward
@pytest.mark.parametrize("name,tag", [("2CUA_A", "i_v4_1"), ("2CUA_A", "i_v4_0"), ("1gpw_A", "i_v4_1"),
                                      ("1EWY", "i_v4_1"), ("tiny40", "i_v4_1"), ("tiny40", "i_v4_0"),
                                      ("synth517", "i_v4_1"), ("batch3", "i_v4_1"), ("batch3", "i_v4_0"),
                                      ("stale257", "i_v4_0")])
def test_logits_match_reference(cuda_models, name, tag):
    c = load_case(name)
    z = run_case(cuda_models(tag), c).cpu()
    ref = torch.from_numpy(c[f"z_{tag}"])
    assert z.shape == ref.shape
    err = (z - ref).abs().max().item()
    assert err <= LOGIT_TOL, err
    assert err <= FP32_EXPECTED, err


def test_published_probabilities(cuda_models):
    """examples/*_i{0..4}.pdb of the reference repo: sigmoid(z) to the 2 decimals of the b-factor column."""
    for name in ("2CUA_A", "1gpw_A"):
        c = load_case(name)
        p = torch.sigmoid(run_case(cuda_models("i_v4_1"), c)).cpu().numpy()
        assert np.abs(p - c["published_prob"]).max() <= 0.0051


def test_sparse_residue_index_equals_dense_membership(cuda_models):
    c = load_case("batch3")
    m = cuda_models("i_v4_0")
    assert torch.equal(run_case(m, c, dense=True), run_case(m, c, dense=False))


def test_unsorted_residue_index(cuda_models):
    """Atoms of a residue need not be contiguous: permute the membership columns' atoms via a shuffled rid."""
    c = load_case("synth517")
    m = cuda_models("i_v4_0")
    X, el, rid, n_res = case_tensors(c)
    g = torch.Generator().manual_seed(3)
    rid2 = rid[torch.randperm(rid.shape[0], generator=g)]
    n_res2 = int(rid2.max()) + 1
    from pesto_b200.dataset import collate_batch_features
    q0 = one_hot_features(el)
    ids1 = collate_batch_features([[X, torch.from_numpy(c["ids0"]).long(), q0, dense_membership(rid2, n_res2)]])[1]
    z = m(X.cuda(), ids1.cuda(), q0.cuda(), dense_membership(rid2, n_res2).cuda()).cpu()
    zo = O.forward(load_weights("i_v4_0"), X, ids1, q0, rid2, n_res2)
    assert (z - zo).abs().max().item() <= FP32_EXPECTED


def test_per_layer_taps(cuda_models):
    """Layer-by-layer (q, p) against forward hooks on the reference's model.sum[L], through the staged C ABI."""
    from pesto_b200 import _lib
    from pesto_b200.dataset import collate_batch_features
    lib = _lib.load()
    c = load_case("2CUA_A")
    tag = "i_v4_1"
    model = cuda_models(tag)
    h = model._handle(0)
    X, el, rid, n_res = case_tensors(c)
    q0 = one_hot_features(el)
    ids1 = collate_batch_features([[X, torch.from_numpy(c["ids0"]).long(), q0, dense_membership(rid, n_res)]])[1]
    n = X.shape[0]
    Xd, idsd, qd = cuda(X, ids1, q0)
    st = [torch.empty((n + 1, 128), device="cuda") for _ in range(2)]
    ids32 = torch.empty((n, 64), dtype=torch.int32, device="cuda")
    geom = torch.empty((n, 64, 4), device="cuda")
    scratch = torch.zeros(16, dtype=torch.uint8, device="cuda")
    node = torch.empty(lib.pesto_node_scratch_bytes(n), dtype=torch.uint8, device="cuda")
    _lib.check(lib.pesto_prologue(h, Xd.data_ptr(), idsd.data_ptr(), 64, qd.data_ptr(), n, st[0].data_ptr(),
                                  ids32.data_ptr(), geom.data_ptr(), scratch.data_ptr(), None), "prologue")
    q = torch.empty((n + 1, 32), device="cuda")
    p = torch.empty((n + 1, 3, 32), device="cuda")
    tap_layers = sorted(int(k.split("_L")[1].split("_")[0]) for k in c if k.startswith(f"tap_{tag}_") and k.endswith("_q"))
    cur = 0
    for layer in range(lib.pesto_model_num_layers(h)):
        _lib.check(lib.pesto_state_update(h, layer, n, ids32.data_ptr(), geom.data_ptr(), st[cur].data_ptr(),
                                          st[1 - cur].data_ptr(), node.data_ptr(), 0, None), "state_update")
        cur = 1 - cur
        if layer in tap_layers:
            _lib.check(lib.pesto_unpack_state(st[cur].data_ptr(), n, q.data_ptr(), p.data_ptr(), None), "unpack")
            eq = (q.cpu() - torch.from_numpy(c[f"tap_{tag}_L{layer}_q"])).abs().max().item()
            ep = (p.cpu() - torch.from_numpy(c[f"tap_{tag}_L{layer}_p"])).abs().max().item()
            assert eq <= 2e-4 and ep <= 5e-4, (layer, eq, ep)
            assert float(q[0].abs().max()) == 0.0 and float(p[0].abs().max()) == 0.0     # sink row stays zero


@pytest.mark.parametrize("mode", BOTH_MODES)
def test_benchmark_table_53_structures(cuda_models, mode):
    """BASELINE config 2 (the workload bench.py times), in the FFMA mode and in the timed tensor-core mode: all 53
    pdbs_test structures through extract_topology -> collate -> Model.forward.

    21 of the 132 417 rows contain an exact fp32 distance tie whose order torch.topk leaves unspecified; in one row
    (XL/4XL5, atom 1829) the tie straddles the nn=32 prefix, so the neighbour SET of that atom is ambiguous and the
    logits of that structure move by ~5e-2 depending on the choice.  Parity is therefore checked on exactly the
    neighbour lists the reference used (our ids with the recorded tie rows substituted), and our own ids are checked
    to be identical modulo those tie groups.  The published 53 x 8 metric table
    (interface_ppi_benchmark.ipynb:168-220) must come out string-identical."""
    from pesto_b200.data_encoding import extract_topology
    from pesto_b200.dataset import collate_batch_features
    g = dict(np.load(os.path.join(GOLDEN, "pdbs_test_53.npz")))
    model = cuda_models("i_v4_1", mode)
    aoff = np.concatenate([[0], np.cumsum(g["sizes"])])
    roff = np.concatenate([[0], np.cumsum(g["n_res"])])
    worst, worst_own, mismatched = 0.0, 0.0, []
    for i, key in enumerate(g["keys"]):
        Xc_ = torch.from_numpy(g["X"][aoff[i]:aoff[i + 1]])
        X = Xc_.cuda()
        el = torch.from_numpy(g["el"][aoff[i]:aoff[i + 1]].astype(np.int64))
        rid = torch.from_numpy(g["rid"][aoff[i]:aoff[i + 1]].astype(np.int64))
        M = dense_membership(rid, int(g["n_res"][i])).cuda()
        q = one_hot_features(el).cuda()
        ids0 = extract_topology(X, 64)[0]
        ref_ids0 = ids0.clone()
        sel = np.nonzero(g["tie_struct"] == i)[0]
        for t in sel:
            ref_ids0[int(g["tie_row"][t])] = torch.from_numpy(g["tie_ids"][t].astype(np.int64)).cuda()
        if len(sel):
            assert O.same_modulo_ties(ids0.cpu(), ref_ids0.cpu(), Xc_)
        zref = torch.from_numpy(g["z_i_v4_1"][roff[i]:roff[i + 1]])
        z = model(*[t if j != 3 else t.float() for j, t in enumerate(collate_batch_features([[X, ref_ids0, q, M]]))]).cpu()
        worst = max(worst, (z - zref).abs().max().item())
        line = scoring.table_line(str(key), g["y"][roff[i]:roff[i + 1]], torch.sigmoid(z[:, 0]).numpy())
        if line != str(g["table_published"][i]):
            mismatched.append((line, str(g["table_published"][i])))
        if not any(g["tie_straddles_prefix"][sel]):
            z_own = model(*collate_batch_features([[X, ids0, q, M]])).cpu()
            worst_own = max(worst_own, (z_own - zref).abs().max().item())
    assert worst <= LOGIT_TOL, worst
    assert worst_own <= LOGIT_TOL, worst_own
    assert not mismatched, mismatched[:3]


@pytest.mark.parametrize("mode", BOTH_MODES)
def test_synth8192_full_size(cuda_models, mode):
    """Full-size (N = 8192, k = 64, i_v4_1: the north_star configuration) against the reference's logits (75 s of CPU, stored)."""
    from pesto_b200.data_encoding import extract_topology
    g = dict(np.load(os.path.join(GOLDEN, "case_synth8192.npz")))
    X, el, rid = synth_structure(8192, BASE_SEED)
    Xd = X.cuda()
    ids1 = extract_topology(Xd, 64)[0] + 1
    z = cuda_models("i_v4_1", mode)(Xd, ids1, one_hot_features(el).cuda(), rid.int().cuda(), n_res=1024).cpu()
    err = (z - torch.from_numpy(g["z_i_v4_1"])).abs().max().item()
    assert err <= LOGIT_TOL, err
    assert err <= (FP32_EXPECTED if mode == "fp32" else 3e-4), err


# ------------------------------------------------------------------------------------------------- properties
@pytest.mark.parametrize("mode", BOTH_MODES)
def test_batched_equals_separate(cuda_models, mode):
    """Structures are independent: a collated batch gives the same logits as separate forwards."""
    from pesto_b200.data_encoding import extract_topology
    from pesto_b200.dataset import collate_batch_features
    model = cuda_models("i_v4_0", mode)
    parts, zs = [], []
    for k, n in enumerate((700, 1301, 96)):
        X, el, rid = synth_structure(n, 100 + k)
        Xd = X.cuda()
        item = [Xd, extract_topology(Xd, 64)[0], one_hot_features(el).cuda(), dense_membership(rid).cuda()]
        parts.append(item)
        zs.append(model(*collate_batch_features([item])))
    zb = model(*collate_batch_features(parts))
    assert (zb - torch.cat(zs)).abs().max().item() <= 1e-4


@pytest.mark.parametrize("mode", BOTH_MODES)
def test_rigid_motion_invariance_full_size(cuda_models, mode):
    """Logits are invariant to rotation + translation of the input cloud (N = 8192)."""
    from pesto_b200.data_encoding import extract_topology
    model = cuda_models("i_v4_1", mode)
    X, el, rid = synth_structure(8192, 77)
    q0 = one_hot_features(el).cuda()
    ridd = rid.int().cuda()
    Xd = X.cuda()
    ids1 = extract_topology(Xd, 64)[0] + 1
    z0 = model(Xd, ids1, q0, ridd, n_res=1024)
    g = torch.Generator().manual_seed(1)
    Q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
    X2 = (X.double() @ Q + torch.tensor([12.5, -40.0, 7.25], dtype=torch.float64)).float().cuda()
    z1 = model(X2, ids1, q0, ridd, n_res=1024)      # same topology: rigid motion preserves neighbour sets
    assert torch.isfinite(z0).all()
    assert (z0 - z1).abs().max().item() <= 2e-3


@pytest.mark.parametrize("mode", BOTH_MODES)
def test_atom_permutation_equivariance(cuda_models, mode):
    """Relabelling atoms (and their ids) leaves the per-residue logits unchanged."""
    from pesto_b200.data_encoding import extract_topology
    model = cuda_models("i_v4_0", mode)
    X, el, rid = synth_structure(1500, 31)
    g = torch.Generator().manual_seed(2)
    perm = torch.randperm(1500, generator=g)
    n_res = int(rid.max()) + 1

    def fwd(X, el, rid):
        Xd = X.cuda()
        ids1 = extract_topology(Xd, 64)[0] + 1
        return model(Xd, ids1, one_hot_features(el).cuda(), rid.int().cuda(), n_res=n_res)
    za, zb = fwd(X, el, rid), fwd(X[perm].contiguous(), el[perm], rid[perm])
    assert (za - zb).abs().max().item() <= 2e-4


# ------------------------------------------------------------------------------------------------- errors
def test_bad_inputs_raise_or_poison(cuda_models):
    from pesto_b200 import _lib
    model = cuda_models("i_v4_0")
    X, el, rid = synth_structure(128, 9)
    q0 = one_hot_features(el).cuda()
    M = dense_membership(rid).cuda()
    ids = torch.randint(0, 129, (128, 64)).cuda()
    with pytest.raises(ValueError):
        model(X.cuda(), ids[:100], q0, M)
    with pytest.raises(ValueError):
        model(X.cuda(), ids, q0[:, :20], M)
    with pytest.raises(_lib.PestoError):
        model(X.cuda(), ids[:, :32].contiguous(), q0, M)           # fewer columns than the largest nn
    bad = ids.clone()
    bad[5, 5] = 4000                                               # out of range -> NaN logits, no crash, and a status flag
    assert torch.isnan(model(X.cuda(), bad, q0, M)).all()
    with pytest.raises(_lib.PestoError, match="neighbour id"):
        model.raise_if_failed()
    M2 = M.clone()
    M2[3] = 0.0                                                    # not one-hot -> NaN logits
    assert torch.isnan(model(X.cuda(), ids, q0, M2)).all()
    with pytest.raises(_lib.PestoError, match="one-hot"):
        model.raise_if_failed()
    rid_bad = torch.zeros(128, dtype=torch.int32, device="cuda")
    rid_bad[7] = 99                                                # residue index >= n_res
    assert torch.isnan(model(X.cuda(), ids, q0, rid_bad, n_res=4)).all()
    with pytest.raises(_lib.PestoError, match="residue index"):
        model.raise_if_failed()
    with pytest.raises(_lib.PestoError):                           # CPU tensors in: the result copy synchronises, so it raises directly
        model(X, bad.cpu(), q0.cpu(), M.cpu())
    assert torch.isfinite(model(X.cuda(), ids, q0, M)).all()       # and the model still works afterwards
    model.raise_if_failed()
    with pytest.raises(ValueError):                                # empty structure: the reference raises too (max of empty)
        model(X[:0].cuda(), ids[:0], q0[:0], M[:0])


def test_hung_tensor_core_stage_is_reported(cuda_models):
    """A tcgen05 stage whose completion never arrives (simulated: the edge kernels skip the commit of their third-layer
    GEMMs) must neither hang the GPU nor pass silently: the kernels' bounded waits give up, the forward's status word
    carries the stage id, every logit is NaN and raise_if_failed raises."""
    from pesto_b200 import _lib
    lib = _lib.load()
    model = cuda_models("i_v4_0", "f16x3")
    c = load_case("synth517")
    assert torch.isfinite(run_case(model, c, mode="f16x3")).all()
    _lib.check(lib.pesto_debug_force_watchdog(1), "force watchdog")
    try:
        z = run_case(model, c, mode="f16x3")
        torch.cuda.synchronize()
        assert torch.isnan(z).all()
        with pytest.raises(_lib.PestoError, match="tensor-core stage 3"):
            model.raise_if_failed()
    finally:
        _lib.check(lib.pesto_debug_force_watchdog(0), "force watchdog off")
    z = run_case(model, c, mode="f16x3")
    model.raise_if_failed()
    assert (z.cpu() - torch.from_numpy(c["z_i_v4_0"])).abs().max().item() <= 3e-4


def test_state_beyond_fp16_operand_range_is_flagged(cuda_models):
    """The tensor-core modes convert the state and the edge activations to fp16 planes with saturation; a state outside the
    validated range (|q| or |p| > 2^14; the shipped checkpoints stay below ~50) is flagged by the per-atom kernel instead of
    saturating silently: code 100 in the watchdog word (staged API: the device's fallback word)."""
    import ctypes
    from pesto_b200 import _lib
    model = cuda_models("i_v4_0", "f16x3")
    lib, h, n, st0, ids32, geom, node = _staged(model, load_case("synth517"))
    out = torch.empty_like(st0)
    word = ctypes.c_int32(0)
    _lib.check(lib.pesto_debug_watchdog(ctypes.byref(word)), "clear")
    _lib.check(lib.pesto_state_update(h, 0, n, ids32.data_ptr(), geom.data_ptr(), st0.data_ptr(), out.data_ptr(), node.data_ptr(), 1, None), "layer")
    torch.cuda.synchronize()
    _lib.check(lib.pesto_debug_watchdog(ctypes.byref(word)), "read")
    assert word.value == 0
    big = st0.clone()
    big[17, 40] = 3.0e4                                            # one vector-state entry beyond 2^14
    _lib.check(lib.pesto_state_update(h, 0, n, ids32.data_ptr(), geom.data_ptr(), big.data_ptr(), out.data_ptr(), node.data_ptr(), 1, None), "layer")
    torch.cuda.synchronize()
    _lib.check(lib.pesto_debug_watchdog(ctypes.byref(word)), "read")
    assert word.value == 100


# ------------------------------------------------------------------------------------------------- tensor-core modes
def _staged(model, c):
    """prologue through the staged C ABI; returns what pesto_state_update needs."""
    from pesto_b200 import _lib
    from pesto_b200.dataset import collate_batch_features
    lib = _lib.load()
    X, el, rid, n_res = case_tensors(c)
    q0 = one_hot_features(el)
    ids1 = collate_batch_features([[X, torch.from_numpy(c["ids0"]).long(), q0, dense_membership(rid, n_res)]])[1]
    n = X.shape[0]
    Xd, idsd, qd = cuda(X, ids1, q0)
    h = model._handle(0)
    st0 = torch.empty((n + 1, 128), device="cuda")
    ids32 = torch.empty((n, 64), dtype=torch.int32, device="cuda")
    geom = torch.empty((n, 64, 4), device="cuda")
    scratch = torch.zeros(16, dtype=torch.uint8, device="cuda")
    node = torch.empty(lib.pesto_node_scratch_bytes(n), dtype=torch.uint8, device="cuda")
    _lib.check(lib.pesto_prologue(h, Xd.data_ptr(), idsd.data_ptr(), 64, qd.data_ptr(), n, st0.data_ptr(),
                                  ids32.data_ptr(), geom.data_ptr(), scratch.data_ptr(), None), "prologue")
    return lib, h, n, st0, ids32, geom, node


@pytest.mark.parametrize("mode,tol", [(1, 2e-4), (2, 6e-2)])
def test_tensor_core_layer_against_fp32_layer(cuda_models, mode, tol):
    """Every layer of i_v4_1 (nn = 8, 16, 32, 64): the tcgen05 edge kernel against the FFMA kernel on the same
    input state (states grow to |q| ~ 40, so the tolerance is relative to the state's magnitude)."""
    from pesto_b200 import _lib
    model = cuda_models("i_v4_1")
    lib, h, n, st0, ids32, geom, node = _staged(model, load_case("2CUA_A"))
    cur = st0
    for layer in range(lib.pesto_model_num_layers(h)):
        ref = torch.empty_like(cur)
        out = torch.empty_like(cur)
        _lib.check(lib.pesto_state_update(h, layer, n, ids32.data_ptr(), geom.data_ptr(), cur.data_ptr(), ref.data_ptr(),
                                          node.data_ptr(), 0, None), "fp32 layer")
        _lib.check(lib.pesto_state_update(h, layer, n, ids32.data_ptr(), geom.data_ptr(), cur.data_ptr(), out.data_ptr(),
                                          node.data_ptr(), mode, None), "tc layer")
        torch.cuda.synchronize()
        scale = max(1.0, ref.abs().max().item())
        err = (out - ref).abs().max().item() / scale
        assert err <= tol, (layer, err, scale)
        cur = ref


@pytest.mark.parametrize("name,tag", [("2CUA_A", "i_v4_1"), ("1gpw_A", "i_v4_1"), ("tiny40", "i_v4_1"),
                                      ("batch3", "i_v4_0"), ("synth517", "i_v4_1"), ("1EWY", "i_v4_1")])
def test_logits_f16x3_within_north_star_tolerance(cuda_models, name, tag):
    """Parity mode on the tensor cores (default mode): 3-term split product over fp16 hi/lo planes.  The north_star bar is
    1e-3 on the logits; the fp16 planes measure ~1e-4 (bf16 planes: 5e-4 .. 7.5e-4), so a regression to bf16-plane accuracy
    fails here (3e-4) long before the bar is at risk.  The former mode name is an alias."""
    c = load_case(name)
    z = run_case(cuda_models(tag), c, mode="f16x3").cpu()
    err = (z - torch.from_numpy(c[f"z_{tag}"])).abs().max().item()
    assert err <= LOGIT_TOL and err <= 3e-4, err
    assert torch.equal(z, run_case(cuda_models(tag), c, mode="bf16x3").cpu())


def test_logits_f16_speed_mode_reports_its_error(cuda_models):
    """A single 16-bit pass cannot meet 1e-3 (tf32-like operands ~0.025 on logits, bf16 ~0.14; fp16 planes
    measure ~0.025); it must stay a sane approximation."""
    c = load_case("2CUA_A")
    z = run_case(cuda_models("i_v4_1"), c, mode="f16").cpu()
    ref = torch.from_numpy(c["z_i_v4_1"])
    err = (z - ref).abs().max().item()
    perr = (torch.sigmoid(z) - torch.sigmoid(ref)).abs().max().item()
    print(f"f16 speed mode: max|dz| = {err:.3e}, max|dp| = {perr:.3e}")
    assert err < 0.2 and perr < 0.03


def test_config4_large_chain_32768(cuda_models):
    """BASELINE config 4 (one N = 32 768 chain, i_v4_1) in the timed tensor-core mode: too large for the CPU oracle in
    test time, so checked through size-independent properties: finite logits, rigid-motion invariance, and agreement of
    the tcgen05 path (f16x3) with the FFMA path (fp32), which the golden cases pin to the reference."""
    from pesto_b200.data_encoding import extract_topology
    X, el, rid = synth_structure(32768, BASE_SEED + 4)
    Xd, q0, ridd = X.cuda(), one_hot_features(el).cuda(), rid.int().cuda()
    ids0 = extract_topology(Xd, 64)[0]
    assert ids0.shape == (32768, 64) and int(ids0.min()) >= 0 and int(ids0.max()) < 32768
    ids1 = ids0 + 1
    model = cuda_models("i_v4_1", "f16x3")
    assert model.mode == "f16x3"
    z = model(Xd, ids1, q0, ridd, n_res=4096)
    assert z.shape == (4096, 5) and torch.isfinite(z).all()
    g = torch.Generator().manual_seed(4)
    Q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
    X2 = (X.double() @ Q + torch.tensor([-3.0, 100.0, 55.5], dtype=torch.float64)).float().cuda()
    assert (z - model(X2, ids1, q0, ridd, n_res=4096)).abs().max().item() <= 2e-3
    z32 = cuda_models("i_v4_1", "fp32")(Xd, ids1, q0, ridd, n_res=4096)
    d = (z - z32).abs().max().item()
    assert 0.0 < d <= LOGIT_TOL, d          # two different arithmetic paths (d == 0 would mean the same kernel ran twice)


@pytest.mark.parametrize("mode", BOTH_MODES)
def test_config3_batch_of_8192_atom_structures(cuda_models, mode):
    """BASELINE config 3 (i_v4_0, synthetic N = 8192 structures collated into one batch): the batched forward over
    several structures equals the separate forwards, and the first structure matches the CPU oracle
    (the full 32-structure batch is timed by profiles/bench_configs.py)."""
    from pesto_b200.data_encoding import batch_topology, extract_topology
    model = cuda_models("i_v4_0", mode)
    n, n_struct = 8192, 4
    Xs, els, rids = zip(*[synth_structure(n, BASE_SEED + 100 + s) for s in range(n_struct)])
    Xb = torch.cat(Xs).cuda()
    q0b = one_hot_features(torch.cat(els)).cuda()
    ridb = torch.cat([r + s * (n // 8) for s, r in enumerate(rids)]).int().cuda()
    ids1b = batch_topology(Xb, [n] * n_struct, 64)
    zb = model(Xb, ids1b, q0b, ridb, n_res=n_struct * (n // 8))
    for s in (0, n_struct - 1):
        Xd = Xs[s].cuda()
        ids1 = extract_topology(Xd, 64)[0] + 1
        assert torch.equal(ids1b[s * n:(s + 1) * n] - s * n, ids1)          # 1-based global ids of collate_batch_features
        zs = model(Xd, ids1, one_hot_features(els[s]).cuda(), rids[s].int().cuda(), n_res=n // 8)
        assert (zb[s * (n // 8):(s + 1) * (n // 8)] - zs).abs().max().item() <= 1e-4
        if s == 0:          # 16 layers x 8192 atoms: ~20 s of CPU oracle
            zo = O.forward(load_weights("i_v4_0"), Xs[0], ids1.cpu(), one_hot_features(els[0]), rids[0], n // 8)
            assert (zs.cpu() - zo).abs().max().item() <= (FP32_EXPECTED if mode == "fp32" else 3e-4)


# ------------------------------------------------------------------------------------------------- v3 checkpoints
def _v3_inputs():
    from pesto_b200.runner import expand_features
    c = load_case("v3_1gpw_A")
    X, rid, n_res = torch.from_numpy(c["X"]), torch.from_numpy(c["rid"].astype(np.int64)), int(c["n_res"])
    q0 = expand_features(torch.from_numpy(c["feat"]), 123)
    ids1 = O.collate([(X, torch.from_numpy(c["ids0"]).long(), q0, rid, n_res)])[1]
    return c, X, ids1, q0, rid, n_res


@pytest.mark.parametrize("mode", BOTH_MODES)
def test_v3_0_checkpoint_matches_reference(cuda_models, mode):
    """i_v3_0: 123 input features, 16 layers (model/save/i_v3_0_2021-05-27_14-27)."""
    c, X, ids1, q0, rid, n_res = _v3_inputs()
    z = cuda_models("i_v3_0", mode)(X.cuda(), ids1.cuda(), q0.cuda(), rid.int().cuda(), n_res=n_res).cpu()
    assert z.shape == (n_res, 5)
    assert (z - torch.from_numpy(c["z_i_v3_0"])).abs().max().item() < 1e-3


@pytest.mark.parametrize("mode", BOTH_MODES)
def test_single_linear_heads_match_oracle(mode):
    """The one-Linear em / dm heads and the single logit of i_v3_1 (model/save/i_v3_1_2021-05-28_12-40/model.py:9-22) on a
    well-conditioned state: i_v3_1's heads around i_v3_0's layers and pooling (the oracle's fp32 and fp64 runs of this
    hybrid agree to 9e-7; the real i_v3_1 layers are ill-conditioned, see the next test)."""
    from pesto_b200.model import Model
    c, X, ids1, q0, rid, n_res = _v3_inputs()
    w0, w1 = load_weights("i_v3_0"), load_weights("i_v3_1")
    hybrid = {k: v for k, v in w0.items() if k.startswith(("sum.", "spl."))}
    hybrid.update({k: v for k, v in w1.items() if k.startswith(("em.", "dm."))})
    model = Model.for_state_dict(load_config("i_v3_1"), {k: torch.from_numpy(v) for k, v in hybrid.items()}, mode=mode).cuda()
    z = model(X.cuda(), ids1.cuda(), q0.cuda(), rid.int().cuda(), n_res=n_res).cpu()
    zo = O.forward(hybrid, X, ids1, q0, rid, n_res)
    assert z.shape == zo.shape == (n_res, 1)
    assert (z - zo).abs().max().item() < 1e-3


def test_v3_1_checkpoint_runs_in_fp32_and_is_refused_on_the_fp16_planes(cuda_models):
    """The shipped i_v3_1 state grows to ~4e5 (tests/test_oracle_golden.py): the fp32 mode reproduces the reference on the
    median residue (the logits are ill-conditioned beyond that: fp32 vs fp64 of the reference's own arithmetic differ by
    up to 8.6), the tensor-core mode reports the state leaving the fp16 operand range instead of returning numbers."""
    from pesto_b200 import _lib
    c, X, ids1, q0, rid, n_res = _v3_inputs()
    args = (X.cuda(), ids1.cuda(), q0.cuda(), rid.int().cuda())
    z = cuda_models("i_v3_1", "fp32")(*args, n_res=n_res).cpu()
    ref = torch.from_numpy(c["z_i_v3_1"])
    assert z.shape == ref.shape == (n_res, 1) and bool(torch.isfinite(z).all())
    assert (z - ref).abs().median().item() < 1e-3
    m = cuda_models("i_v3_1", "f16x3")
    zt = m(*args, n_res=n_res)
    with pytest.raises(_lib.PestoError, match="fp16 operand planes"):
        m.raise_if_failed()
    assert bool(torch.isnan(zt).all())


def test_runner_feeds_the_123_feature_models(cuda_models):
    """predict_structures on a structure dictionary with resname / name columns == Model.forward on encode_features."""
    from pesto_b200.runner import predict_structures
    c, X, ids1, q0, rid, n_res = _v3_inputs()
    s = {k: c[k] for k in ("element", "resname", "name", "resid")}
    s["xyz"] = c["X"]
    model = cuda_models("i_v3_0", "f16x3")
    (idx, z), = list(predict_structures(model, [s]))
    assert idx == 0 and z.shape == (n_res, 5)
    assert (z - torch.from_numpy(c["z_i_v3_0"])).abs().max().item() < 1e-3
