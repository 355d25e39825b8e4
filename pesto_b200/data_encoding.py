"""Topology and feature encoding with the reference's names and return conventions
(src/data_encoding.py of LBM-EPFL/PeSTo).  `extract_topology` runs the exact-kNN CUDA kernel; the
encoders are thin host-side helpers (one-hot tables), kept so that callers of the reference find them.
"""
import ctypes

import numpy as np
import torch

from . import _lib

# vocabulary order fixes the one-hot columns the shipped checkpoints were trained on
# (src/data_encoding.py:6-29); the last column of every one-hot is "unknown".
std_elements = np.array("C O N S P Se Mg Cl Zn Fe Ca Na F Mn I K Br Cu Cd Ni Co Sr Hg W As B Mo Ba Pt".split())
std_resnames = np.array(("LEU GLU ARG LYS VAL ILE PHE ASP TYR ALA THR SER GLN ASN PRO GLY HIS TRP MET CYS "
                         "G A C U DG DA DT DC").split())
std_names = np.array(("CA N C O CB CG CD2 CD1 CG1 CG2 CD OE1 OE2 OG OG1 OD1 OD2 CE NZ NE CZ NH2 NH1 ND2 CE2 CE1 "
                      "NE2 OH ND1 SD SG NE1 CE3 CZ3 CZ2 CH2 P C3' C4' O3' C5' O5' O4' C1' C2' O2' OP1 OP2 N9 N2 O6 "
                      "N7 C8 N1 N3 C2 C4 C6 C5 N6 N4 O2 O4").split())
config_encoding = {"std_elements": std_elements, "std_resnames": std_resnames, "std_names": std_names}

# residue-name categories of the five interface types (src/data_encoding.py:32-46): a vocabulary the reference's own
# config.py imports at module level (`from src.data_encoding import categ_to_resnames`), so it has to exist for
# `from config import config_model` to run under pesto_b200.compat; the forward path never reads it
categ_to_resnames = {
    "protein": "GLU LEU ALA ASP SER VAL GLY THR ARG PHE TYR ILE PRO ASN LYS GLN HIS TRP MET CYS".split(),
    "rna": "A U G C".split(),
    "dna": "DA DT DG DC".split(),
    "ion": "MG ZN CL CA NA MN K IOD CD CU FE NI SR BR CO HG".split(),
    "ligand": ("SO4 NAG PO4 EDO ACT MAN HEM FMT BMA ADP FAD NAD NO3 GLC ATP NAP BGC GDP FUC FES FMN GAL GTP PLP MLI "
               "ANP H4B AMP NDP SAH OXY").split(),
    "lipid": "PLM CLR CDL RET".split(),
}
resname_to_categ = {rn: c for c, names in categ_to_resnames.items() for rn in names}


def onehot_index(x, v):
    """Column of every entry of x in `onehot(x, v)`: its position in vocabulary v, or len(v) ("not in v", the last column) --
    one binary search per entry instead of a [N, V] string comparison."""
    x, v = np.asarray(x).reshape(-1), np.asarray(v).reshape(-1)
    if not len(v):
        return np.zeros(len(x), dtype=np.int64)
    order = np.argsort(v, kind="stable")
    vs = v[order]
    pos = np.minimum(np.searchsorted(vs, x), len(vs) - 1)
    return np.where(vs[pos] == x, order[pos], len(v))


def onehot(x, v):
    """[len(x), len(v)+1] bool: membership in vocabulary v, last column = not in v (src/data_encoding.py:56-58)."""
    col = onehot_index(x, v)
    out = np.zeros((len(col), len(v) + 1), dtype=bool)
    out[np.arange(len(col)), col] = True
    return out


def encode_structure(structure, device=torch.device("cpu")):
    """X[N,3] float32 and the boolean residue membership M[N,R] (src/data_encoding.py:61-75)."""
    xyz, resid = structure["xyz"], structure["resid"]
    X = (xyz if isinstance(xyz, torch.Tensor) else torch.from_numpy(np.asarray(xyz, dtype=np.float32))).to(device)
    r = (resid if isinstance(resid, torch.Tensor) else torch.from_numpy(np.asarray(resid))).to(device)
    M = r.unsqueeze(1) == torch.unique(r).unsqueeze(0)
    return X, M


def encode_features(structure, device=torch.device("cpu")):
    """(qe[N,30], qr[N,29], qn[N,64]) float32 one-hots (src/data_encoding.py:78-84); v4 models use qe only."""
    return tuple(torch.from_numpy(onehot(structure[k], v).astype(np.float32)).to(device)
                 for k, v in (("element", std_elements), ("resname", std_resnames), ("name", std_names)))


def _knn(X, seg_off, k, base):
    lib = _lib.load()
    dev = X.device
    n_atoms, n_seg = X.shape[0], seg_off.numel() - 1
    ids = torch.empty((n_atoms, k), dtype=torch.int64, device=dev)
    d = torch.empty((n_atoms, k), dtype=torch.float32, device=dev)
    r = torch.empty((n_atoms, k, 3), dtype=torch.float32, device=dev)
    scratch = torch.empty(lib.pesto_knn_scratch_bytes(n_atoms, n_seg), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib.pesto_knn(X.data_ptr(), n_atoms, seg_off.data_ptr(), n_seg, k, base, ids.data_ptr(), d.data_ptr(),
                           r.data_ptr(), scratch.data_ptr(), ctypes.c_void_p(stream))
    _lib.check(rc, "pesto_knn")
    return ids, d, r


def extract_topology(X, num_nn):
    """(ids_topk, D_topk, R_topk, D, R) like src/data_encoding.py:87-102.

    ids_topk [N, min(num_nn, N)] int64, 0-based, ascending (masked distance, index).  The dense N x N tensors D
    and R are never built; no caller of the reference reads them, so they are returned as None.
    CPU tensors are staged through the GPU (results come back on X's device); without CUDA this raises.
    """
    if X.dim() != 2 or X.shape[1] != 3 or X.shape[0] == 0:
        # the reference fails on an empty structure as well (torch.max of an empty tensor, src/data_encoding.py:93)
        raise ValueError(f"extract_topology: X must be [N, 3] with N >= 1, got {tuple(X.shape)}")
    if not torch.cuda.is_available():
        raise _lib.PestoError("extract_topology needs a CUDA device: there is no CPU implementation")
    if num_nn > 64:
        raise ValueError("extract_topology: num_nn > 64 is not supported by the CUDA kernel")
    out_device = X.device
    dev = X.device if X.is_cuda else torch.device("cuda", torch.cuda.current_device())
    Xd = X.detach().to(device=dev, dtype=torch.float32).contiguous()
    n = Xd.shape[0]
    seg_off = torch.tensor([0, n], dtype=torch.int32, device=dev)
    knn = min(num_nn, n)
    ids, d, r = _knn(Xd, seg_off, num_nn, 0)
    if knn < num_nn:
        ids, d, r = ids[:, :knn].contiguous(), d[:, :knn].contiguous(), r[:, :knn].contiguous()
    if out_device != dev:
        ids, d, r = ids.to(out_device), d.to(out_device), r.to(out_device)
    return ids, d, r, None, None


def batch_topology(X, sizes, num_nn=64):
    """Extension: kNN + index shift + sink padding for a whole batch of structures in one launch.

    X [sum(sizes), 3] holds the structures back to back.  Returns ids_topk [N, num_nn] int64, 1-based global, 0 = sink
    -- identical to collate_batch_features over per-structure extract_topology outputs (src/dataset.py:100-109).
    """
    dev = X.device
    off = torch.zeros(len(sizes) + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(torch.as_tensor(sizes, dtype=torch.int64), 0)
    if int(off[-1]) != X.shape[0]:
        raise ValueError("sizes do not add up to the number of atoms")
    seg_off = off.to(torch.int32).to(dev)
    return _knn(X.contiguous(), seg_off, num_nn, 1)[0]
