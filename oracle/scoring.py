"""Binary-classification scores of the reference's benchmark table -- TEST INFRASTRUCTURE, NOT PRODUCT.

numpy restatement of `bc_scoring` (src/scoring.py:77-96 of LBM-EPFL/PeSTo): [acc, ppv, npv, tpr, tnr, mcc, auc, std]
for one label column, and the line format of interface_ppi_benchmark.ipynb cell 6 (:249 in the survey numbering).
Pinned by tests/test_oracle_golden.py against the 53 published lines.
"""
import numpy as np

SCORE_NAMES = ["acc", "ppv", "npv", "tpr", "tnr", "mcc", "auc", "std"]


def roc_auc(y, p):
    """Mann-Whitney U with average ranks for ties (= sklearn.metrics.roc_auc_score, src/scoring.py:59-67)."""
    y = np.asarray(y) > 0.5
    n_pos, n_neg = int(y.sum()), int((~y).sum())
    if n_pos == 0 or n_neg == 0:
        return np.nan
    order = np.argsort(p, kind="mergesort")
    ps = np.asarray(p, dtype=np.float64)[order]
    ranks = np.empty(len(ps), dtype=np.float64)
    i = 0
    while i < len(ps):
        j = i
        while j + 1 < len(ps) and ps[j + 1] == ps[i]:
            j += 1
        ranks[i:j + 1] = 0.5 * (i + j) + 1.0
        i = j + 1
    r = np.empty_like(ranks)
    r[order] = ranks
    return (r[y].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg)


def bc_scores(y, p):
    """y, p: 1-D arrays (labels in {0,1}, probabilities).  float32 arithmetic like the torch reference."""
    y = np.asarray(y, dtype=np.float32)
    p = np.asarray(p, dtype=np.float32)
    q = np.round(p)                                             # src/scoring.py:79
    TP, TN = np.sum(q * y), np.sum((1 - q) * (1 - y))           # :13-14
    FP, FN = np.sum(q * (1 - y)), np.sum((1 - q) * y)           # :15-16
    P, N = np.sum(y), np.sum(1 - y)                             # :17-18
    with np.errstate(divide="ignore", invalid="ignore"):
        f = np.float32
        acc = f(TP + TN) / f(TP + TN + FP + FN)                 # :25
        ppv = f(TP) / f(TP + FP) if P > 0 else np.nan           # :30-32
        npv = f(TN) / f(TN + FN) if N > 0 else np.nan           # :37-39
        tpr = f(TP) / f(TP + FN)                                # :44-46
        tnr = f(TN) / f(TN + FP)                                # :51-53
        mcc = f(TP * TN - FP * FN) / np.sqrt(f((TP + FP) * (TP + FN) * (TN + FP) * (TN + FN)))   # :56-58
    tpr = np.nan if np.isinf(tpr) else tpr
    tnr = np.nan if np.isinf(tnr) else tnr
    mcc = np.nan if np.isinf(mcc) else mcc
    auc = roc_auc(y, p) if (P > 0 and N > 0) else np.nan        # :61-67
    std = np.std(p.astype(np.float32), ddof=1)                  # :93  torch.std is unbiased
    return [acc, ppv, npv, tpr, tnr, mcc, auc, std]


def table_line(key, y, p):
    s = bc_scores(y, p)
    return ", ".join([key] + [f"{n}={float(v):.3f}" for n, v in zip(SCORE_NAMES, s)])
