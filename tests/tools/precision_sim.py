"""Precision-mix simulator for the tensor-core edge kernel -- TEST INFRASTRUCTURE (uses oracle/).

Emulates on the CPU, in fp32, which operand planes of each edge GEMM the tcgen05 kernel feeds to the tensor core:
every GEMM `x @ W^T` of the three edge MLPs (src/model_operations.py:122-128) can run
    "3"  : x and W both as fp16 hi + lo planes, 3-term product (~fp32 exact)          -> modelled as exact fp32
    "a"  : activations rounded to fp16 (hi plane only), W hi + lo (2 MMAs per K step)
    "w"  : weights rounded to fp16, activations hi + lo (2 MMAs per K step)
    "1"  : both rounded (1 MMA per K step)
and the attention-weighted reduction (:143-144) can round the weights, V and p_j operands to fp16 ("R on the tensor core").
Prints the max-abs logit error against the reference logits of the golden cases for a list of mixes, so that the
cheapest mix within the parity budget can be picked without a GPU (profiles/r2_precision_mix.md holds the table).

usage: python tests/tools/precision_sim.py [case ...]
"""
import math
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from conftest import load_case, load_weights, case_tensors   # noqa: E402
from oracle import pesto_oracle as O                         # noqa: E402
from pesto_b200.synth import one_hot_features                # noqa: E402

ELU = torch.nn.functional.elu


def r16(x):
    return x.half().float()


def gemm(x, W, b, mode):
    if mode in ("a", "1"):
        x = r16(x)
    if mode in ("w", "1"):
        W = r16(W)
    y = x @ W.T
    return y if b is None else y + b


def edge_mlp(w, prefix, X_e, cols_edge, m):
    """m = (mode1, mode2, mode3).  Layer 1: only the columns the kernel feeds through the MMA (d, p_i.r, p_j.r) see mode1;
    the q_i/pn_i/q_j/pn_j columns are the exact per-atom factors U_i, T_j (node kernel, 3-term)."""
    W1 = w[prefix + ".0.weight"]
    mask = torch.zeros(W1.shape[1], dtype=torch.bool)
    mask[cols_edge] = True
    exact = X_e[..., ~mask] @ W1[:, ~mask].T + w[prefix + ".0.bias"]
    if m[0].startswith("T"):            # T_j gathered as fp16 (hi only): round the neighbour-side factor
        S = 32
        tj = X_e[..., 1 + 2 * S:1 + 4 * S] @ W1[:, 1 + 2 * S:1 + 4 * S].T
        exact = exact - tj + r16(tj)
    h = ELU(exact + gemm(X_e[..., mask], W1[:, mask], None, m[0].lstrip("T")))
    h = ELU(gemm(h, w[prefix + ".2.weight"], w[prefix + ".2.bias"], m[1]))
    return gemm(h, w[prefix + ".4.weight"], w[prefix + ".4.bias"], m[2])


def state_update(w, pre, q, p, ids1, D, R, nn, cfg):
    S = q.shape[1]
    Nh, Nk = 2, 3
    sdk = math.sqrt(Nk)
    N = q.shape[0] - 1
    rows = slice(1, N + 1)
    qi, pi = q[rows], p[rows]
    ids = ids1[:, :nn]
    d = D[:, :nn]
    r = R[:, :nn]
    qj, pj = q[ids], p[ids]
    n = qi.shape[0]
    pn_i = torch.sqrt((pi * pi).sum(1))
    X_n = torch.cat([qi, pn_i], 1)
    pj_s0 = r16(pj) if cfg.get("s0_pj16") else pj
    X_e = torch.cat([d.unsqueeze(2), X_n.unsqueeze(1).expand(n, nn, 2 * S), qj, torch.sqrt((pj * pj).sum(2)),
                     (pi.unsqueeze(1) * r.unsqueeze(3)).sum(2), (pj_s0 * r.unsqueeze(3)).sum(2)], 2)
    cols_edge = [0] + list(range(1 + 4 * S, 1 + 6 * S))
    Q = O.mlp3(w, pre + "su.nqm", X_n).view(n, 2, Nh, Nk)
    Kq = edge_mlp(w, pre + "su.eqkm", X_e, cols_edge, cfg["q"])
    Kp = edge_mlp(w, pre + "su.epkm", X_e, cols_edge, cfg["p"]).view(n, nn, 3, Nk)
    V = edge_mlp(w, pre + "su.evm", X_e, cols_edge, cfg["v"])
    V0, V1 = V[..., :S], V[..., S:]
    lq = torch.einsum("nhk,njk->nhj", Q[:, 0], Kq) / sdk
    lp = torch.einsum("nhk,njgk->nhgj", Q[:, 1], Kp) / sdk
    Mq = torch.softmax(lq, dim=2)
    Mp = torch.softmax(lp.reshape(n, Nh, 3 * nn), dim=2).view(n, Nh, 3, nn)
    rw, rv, rp = cfg.get("r_w", False), cfg.get("r_v", False), cfg.get("r_pj", False)
    f = lambda x, on: r16(x) if on else x
    Wv = Mp[:, :, 0].unsqueeze(3) * r.unsqueeze(1)                     # [n,h,j,c] weights of the V1 (x) r tokens
    Zq = torch.einsum("nhj,njs->nhs", f(Mq, rw), f(V0, rv)).reshape(n, Nh * S)
    Zp = (torch.einsum("nhjc,njs->nchs", f(Wv, rw), f(V1, rv))
          + torch.einsum("nh,ncs->nchs", Mp[:, :, 1].sum(2), pi)
          + torch.einsum("nhj,njcs->nchs", f(Mp[:, :, 2], rw), f(pj, rp))).reshape(n, 3, Nh * S)
    q_out = torch.zeros_like(q)
    p_out = torch.zeros_like(p)
    q_out[rows] = qi + O.mlp3(w, pre + "su.qpm", Zq)
    p_out[rows] = pi + Zp @ w[pre + "su.ppm.0.weight"].T
    return q_out, p_out


def forward(weights, X, ids1, q0, rid, n_res, cfg):
    nns = O.layer_nn(weights)
    w = O.prepare_weights(weights)
    q = O.mlp3(w, "em", q0.float())
    N, S = q.shape
    D, R = O.unpack_geometry(X, ids1.long())
    q = torch.cat([torch.zeros((1, S)), q], 0)
    p = torch.zeros((N + 1, 3, S))
    for li, nn in enumerate(nns):
        q, p = state_update(w, f"sum.{li}.", q, p, ids1.long(), D, R, nn, cfg)
    return O.pool_decode(w, q[1:], p[1:], rid.long(), int(n_res))


def mix(q="333", p="333", v="333", **kw):
    d = {"q": tuple(q), "p": tuple(p), "v": tuple(v)}
    d.update(kw)
    return d


MIXES = {
    "exact (3-term everywhere)": mix(),
    "1-term everywhere (mode f16)": mix("111", "111", "111"),
    "a-only M2,M3 all": mix("3aa", "3aa", "3aa"),
    "a-only M1,M2,M3 all": mix("aaa", "aaa", "aaa"),
    "w-only M2,M3 all": mix("3ww", "3ww", "3ww"),
    "1-term M2,M3 all": mix("311", "311", "311"),
    "1-term M2,M3 of q,p; v 3-term": mix("311", "311", "333"),
    "1-term M2,M3 of q,p; v a-only": mix("311", "311", "3aa"),
    "1-term all of q,p; v a-only M2,M3": mix("111", "111", "3aa"),
    "a-only q,p; v 3-term": mix("3aa", "3aa", "333"),
    "v: a-only M2 only": mix("333", "333", "3a3"),
    "v: a-only M3 only": mix("333", "333", "33a"),
    "v: 1-term M2,M3": mix("333", "333", "311"),
    "R operands fp16 (w, V, p_j)": mix(r_w=True, r_v=True, r_pj=True),
    "R: weights fp16 only": mix(r_w=True),
    "R: V fp16 only": mix(r_v=True),
    "R: p_j fp16 only": mix(r_pj=True),
    "S0: p_j fp16": mix(s0_pj16=True),
    "T_j fp16": mix(("T3", "3", "3"), ("T3", "3", "3"), ("T3", "3", "3")),
}


def main():
    cases = sys.argv[1:] or ["2CUA_A"]
    only = os.environ.get("MIX")
    W = load_weights("i_v4_1")
    for name in cases:
        c = load_case(name)
        X, el, rid, n_res = case_tensors(c)
        ids1 = O.collate([(X, torch.from_numpy(c["ids0"]).long(), one_hot_features(el), rid, n_res)])[1]
        zref = torch.from_numpy(c["z_i_v4_1"])
        print(f"case {name}: {X.shape[0]} atoms", flush=True)
        for label, cfg in MIXES.items():
            if only and only not in label:
                continue
            t0 = time.time()
            z = forward(W, X, ids1, one_hot_features(el), rid, n_res, cfg)
            err = (z - zref).abs().max().item()
            perr = (torch.sigmoid(z) - torch.sigmoid(zref)).abs().max().item()
            print(f"  {label:42s} max|dz| = {err:.2e}   max|dp| = {perr:.2e}   ({time.time() - t0:.0f} s)", flush=True)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    main()
