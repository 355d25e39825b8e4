"""CPU oracle for the PeSTo forward hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A plain torch-CPU restatement of the reference's algorithm (explicit per-edge formulas,
no nn.Module), written from the reference's sources and citing them line by line.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs
may import this file; the product (`pesto_b200/`) never does.

Pinned: `tests/test_oracle_golden.py` checks this file against outputs of the UNMODIFIED
reference run in the build container (`tests/golden/*.npz`, made by
`tests/golden/make_golden.py`): logits, per-layer (q, p) taps, kNN indices, the two
`examples/*_i{0..4}.pdb` published probability sets and the 53-row metric table of
`interface_ppi_benchmark.ipynb:168-220`.

All citations are file:line into the reference repository (LBM-EPFL/PeSTo).
"""
import math

import numpy as np
import torch

ELU = torch.nn.functional.elu


# ------------------------------------------------------------------------------------------------
# topology: src/data_encoding.py:87-102 (extract_topology)
# ------------------------------------------------------------------------------------------------
def masked_distances(X):
    """D' and R of src/data_encoding.py:89-95 for one structure (dense; oracle sizes only)."""
    R = X.unsqueeze(0) - X.unsqueeze(1)                        # :89   R[i,j] = X[j] - X[i]
    D = torch.norm(R, dim=2)                                   # :91
    D = D + torch.max(D) * (D < 1e-2).float()                  # :93   self / duplicates pushed to max(D)
    R = R / D.unsqueeze(2)                                     # :95
    return D, R


def extract_topology(X, num_nn=64):
    """ids_topk (0-based, int64), D_topk, R_topk as src/data_encoding.py:98-100.

    torch.topk leaves the order inside exact-distance tie groups unspecified;
    the oracle (and the CUDA kernel) fix it to (distance, index) lexicographic, which a
    stable ascending sort gives.
    """
    D, R = masked_distances(X)
    knn = min(num_nn, D.shape[0])                              # :98
    order = torch.sort(D, dim=1, stable=True)[1][:, :knn]      # :99  (largest=False, sorted)
    D_topk = torch.gather(D, 1, order)
    R_topk = torch.gather(R, 1, order.unsqueeze(2).repeat(1, 1, 3))   # :100
    return order, D_topk, R_topk


def same_modulo_ties(ids_a, ids_b, X):
    """True if two [N,k] index sets agree up to permutations inside equal-distance groups."""
    if ids_a.shape != ids_b.shape:
        return False
    D, _ = masked_distances(X)

    def canonical(ids):
        ids = torch.sort(ids.long(), dim=1)[0]                 # index ascending ...
        d = torch.gather(D, 1, ids)
        o = torch.sort(d, dim=1, stable=True)[1]               # ... then stable by distance = (d, idx) order
        return torch.gather(d, 1, o), torch.gather(ids, 1, o)

    da, ia = canonical(ids_a)
    db, ib = canonical(ids_b)
    # the as-given order must already be ascending in distance (only tie groups may be permuted)
    for ids in (ids_a, ids_b):
        d = torch.gather(D, 1, ids.long())
        if bool((d[:, 1:] < d[:, :-1]).any()):
            return False
    return torch.equal(da, db) and torch.equal(ia, ib)


# ------------------------------------------------------------------------------------------------
# batching: src/dataset.py:91-112 (collate_batch_features)
# ------------------------------------------------------------------------------------------------
def collate(structures, max_num_nn=64):
    """structures: list of (X, ids0, q0, rid, n_res).  Returns X, ids1 (1-based, 0 = sink), q0, rid."""
    X = torch.cat([s[0] for s in structures], 0)               # :93
    q0 = torch.cat([s[2] for s in structures], 0)              # :94
    ids1 = torch.zeros((X.shape[0], max_num_nn), dtype=torch.long)   # :100
    rid = torch.zeros(X.shape[0], dtype=torch.long)
    ix0, iy0 = 0, 0
    for Xs, ids0, _q, r, n_res in structures:
        n = Xs.shape[0]
        ids1[ix0:ix0 + n, :ids0.shape[1]] = ids0.long() + ix0 + 1     # :109
        rid[ix0:ix0 + n] = r.long() + iy0                              # :110 (block-diagonal M, kept sparse)
        ix0 += n
        iy0 += int(n_res)
    return X, ids1, q0, rid, iy0


# ------------------------------------------------------------------------------------------------
# geometry: src/model_operations.py:6-22 (unpack_state_features)
# ------------------------------------------------------------------------------------------------
def unpack_geometry(X, ids1):
    R = X[ids1 - 1] - X.unsqueeze(1)                           # :8   id 0 (sink) -> X[-1]
    D = torch.norm(R, dim=2)                                   # :10
    D = D + torch.max(D) * (D < 1e-2).to(D.dtype)              # :12  global max over the batch
    R = R / D.unsqueeze(2)                                     # :14
    return D, R


def mlp3(w, prefix, x):
    """Linear-ELU-Linear-ELU-Linear (src/model_operations.py:35-82); a single Linear when the checkpoint has no
    `<prefix>.2.weight` (the em / dm heads of model/save/i_v3_1_2021-05-28_12-40/model.py:9-22)."""
    if prefix + ".2.weight" not in w:
        return x @ w[prefix + ".0.weight"].T + w[prefix + ".0.bias"]
    h = ELU(x @ w[prefix + ".0.weight"].T + w[prefix + ".0.bias"])
    h = ELU(h @ w[prefix + ".2.weight"].T + w[prefix + ".2.bias"])
    return h @ w[prefix + ".4.weight"].T + w[prefix + ".4.bias"]


# ------------------------------------------------------------------------------------------------
# StateUpdate: src/model_operations.py:87-154, called from StateUpdateLayer.forward :225-242
# ------------------------------------------------------------------------------------------------
def state_update(w, pre, q, p, ids1, D, R, nn, chunk=4096):
    """One layer. q[N+1,S], p[N+1,3,S] include the sink row 0; D[N,64], R[N,64,3] do not.

    Returns new (q, p) with the sink row reset to zero (:239-240).
    """
    S = q.shape[1]
    Nh, Nk = 2, 3
    sdk = math.sqrt(Nk)                                        # :85
    q_out = torch.zeros_like(q)
    p_out = torch.zeros_like(p)
    N = q.shape[0] - 1
    for a0 in range(0, N, chunk):
        a1 = min(N, a0 + chunk)
        rows = slice(a0 + 1, a1 + 1)
        qi, pi = q[rows], p[rows]                              # [n,S], [n,3,S]
        ids = ids1[a0:a1, :nn]                                 # :230  prefix slice
        d = D[a0:a1, :nn]
        r = R[a0:a1, :nn]                                      # [n,nn,3]
        qj, pj = q[ids], p[ids]                                # :236  gathers [n,nn,S], [n,nn,3,S]
        n = qi.shape[0]
        pn_i = torch.sqrt((pi * pi).sum(1))                    # :105  |p_i|
        X_n = torch.cat([qi, pn_i], 1)                         # :103-106
        X_e = torch.cat([                                      # :109-116
            d.unsqueeze(2),
            X_n.unsqueeze(1).expand(n, nn, 2 * S),
            qj,
            torch.sqrt((pj * pj).sum(2)),
            (pi.unsqueeze(1) * r.unsqueeze(3)).sum(2),
            (pj * r.unsqueeze(3)).sum(2),
        ], 2)
        Q = mlp3(w, pre + "su.nqm", X_n).view(n, 2, Nh, Nk)    # :119
        Kq = mlp3(w, pre + "su.eqkm", X_e)                     # :122  [n,nn,Nk]
        Kp = mlp3(w, pre + "su.epkm", X_e).view(n, nn, 3, Nk)  # :125  chunk g <-> token group g
        V = mlp3(w, pre + "su.evm", X_e)                       # :128
        V0, V1 = V[..., :S], V[..., S:]
        lq = torch.einsum("nhk,njk->nhj", Q[:, 0], Kq) / sdk                       # :139
        lp = torch.einsum("nhk,njgk->nhgj", Q[:, 1], Kp) / sdk                     # :140
        Mq = torch.softmax(lq, dim=2)
        Mp = torch.softmax(lp.reshape(n, Nh, 3 * nn), dim=2).view(n, Nh, 3, nn)    # one softmax over 3*nn tokens
        Zq = torch.einsum("nhj,njs->nhs", Mq, V0).reshape(n, Nh * S)               # :143
        Zp = (torch.einsum("nhj,njs,njc->nchs", Mp[:, :, 0], V1, r)                # :131-136, :144
              + torch.einsum("nh,ncs->nchs", Mp[:, :, 1].sum(2), pi)
              + torch.einsum("nhj,njcs->nchs", Mp[:, :, 2], pj)).reshape(n, 3, Nh * S)
        q_out[rows] = qi + mlp3(w, pre + "su.qpm", Zq)                             # :147, :151
        p_out[rows] = pi + Zp @ w[pre + "su.ppm.0.weight"].T                        # :148, :152
    return q_out, p_out


# ------------------------------------------------------------------------------------------------
# StatePoolLayer + decoder: src/model_operations.py:197-213, model/model.py:46-50
# ------------------------------------------------------------------------------------------------
def pool_decode(w, q, p, rid, n_res):
    """q[N,S], p[N,3,S] (sink removed), rid[N] residue column -> logits z[R,5]."""
    N, S = q.shape
    Nh = w["spl.sam.4.weight"].shape[0] // 2
    zf = torch.cat([q, torch.sqrt((p * p).sum(1))], 1)                              # :202
    a = mlp3(w, "spl.sam", zf)                                                      # :205 [N, 2*Nh] index h*2+t
    amax = torch.full((n_res, 2 * Nh), -float("inf"), dtype=a.dtype)
    amax = amax.scatter_reduce(0, rid.unsqueeze(1).expand(N, 2 * Nh), a, "amax")
    e = torch.exp(a - amax[rid])                                                    # softmax over the atoms of a residue
    den = torch.zeros((n_res, 2 * Nh), dtype=a.dtype).index_add_(0, rid, e)
    wgt = (e / den[rid]).view(N, Nh, 2)
    qh = torch.zeros((n_res, S, Nh), dtype=a.dtype).index_add_(0, rid, q.unsqueeze(2) * wgt[:, None, :, 0])        # :206
    ph = torch.zeros((n_res, 3, S, Nh), dtype=a.dtype).index_add_(0, rid, p.unsqueeze(3) * wgt[:, None, None, :, 1])  # :207
    qr = mlp3(w, "spl.zdm", qh.reshape(n_res, S * Nh))                              # :210
    pr = ph.reshape(n_res, 3, S * Nh) @ w["spl.zdm_vec.0.weight"].T                 # :211
    zr = torch.cat([qr, torch.sqrt((pr * pr).sum(1))], 1)                           # model/model.py:49
    return mlp3(w, "dm", zr)                                                        # model/model.py:50


# ------------------------------------------------------------------------------------------------
# Model.forward: model/model.py:32-52
# ------------------------------------------------------------------------------------------------
def prepare_weights(weights, dtype=torch.float32):
    return {k: torch.as_tensor(np.asarray(v)).to(dtype) for k, v in weights.items()
            if not (k.endswith("m_nn") or k.endswith("sdk"))}


def layer_nn(weights):
    """nn per layer from the checkpoint's `sum.L.m_nn` buffers (src/model_operations.py:223)."""
    L = 1 + max(int(k.split(".")[1]) for k in weights if k.startswith("sum."))
    return [int(np.asarray(weights[f"sum.{i}.m_nn"]).shape[0]) for i in range(L)]


def forward(weights, X, ids1, q0, rid, n_res, dtype=torch.float32, taps=None, chunk=4096):
    """Logits z[R,N2] (5; 1 for i_v3_1).  X[N,3], ids1[N,64] 1-based (0 = sink), q0[N,N0] (30; 123 for the v3 models),
    rid[N] residue column.

    `taps`, if a dict, receives {layer: (q, p)} including the sink row, like forward hooks on
    `model.sum[layer]` of the reference.
    """
    nns = layer_nn(weights)
    w = prepare_weights(weights, dtype)
    X = X.to(dtype)
    q = mlp3(w, "em", q0.to(dtype))                                                 # model/model.py:34
    N, S = q.shape
    D, R = unpack_geometry(X, ids1.long())                                          # model/model.py:40
    q = torch.cat([torch.zeros((1, S), dtype=dtype), q], 0)                         # src/model_operations.py:17
    p = torch.zeros((N + 1, 3, S), dtype=dtype)                                     # model/model.py:37
    for li, nn in enumerate(nns):                                                   # model/model.py:43
        q, p = state_update(w, f"sum.{li}.", q, p, ids1.long(), D, R, nn, chunk=chunk)
        if taps is not None and li in taps:
            taps[li] = (q.clone(), p.clone())
    return pool_decode(w, q[1:], p[1:], rid.long(), int(n_res))                     # model/model.py:46-50
