"""Batch collation with the reference's signature (src/dataset.py:91-112 of LBM-EPFL/PeSTo).

Index plumbing only (concatenate, shift to 1-based global rows, pad with the sink id 0); works on any device.
"""
import torch


def collate_batch_features(batch_data, max_num_nn=64, sparse_membership=False):
    """batch_data: list of [X, ids_topk (0-based), q, M].  Returns (X, ids_topk, q, M).

    ids_topk [N, max_num_nn] int64 is 1-based over the concatenated atoms with 0 = sink in unfilled columns;
    M is the dense block-diagonal membership [N, R] float32 like the reference, or -- with
    sparse_membership=True -- the int32 residue column per atom (accepted by pesto_b200.Model.forward), which
    avoids the O(N*R) matrix for large batches.
    """
    X = torch.cat([d[0] for d in batch_data], dim=0)
    q = torch.cat([d[2] for d in batch_data], dim=0)
    n_atoms = [d[3].shape[0] for d in batch_data]
    n_res = [d[3].shape[1] for d in batch_data]
    ids_topk = torch.zeros((X.shape[0], max_num_nn), dtype=torch.long, device=X.device)
    if sparse_membership:
        M = torch.empty(X.shape[0], dtype=torch.int32, device=X.device)
    else:
        M = torch.zeros((sum(n_atoms), sum(n_res)), dtype=torch.float, device=X.device)
    a0 = r0 = 0
    for d, na, nr in zip(batch_data, n_atoms, n_res):
        ids_topk[a0:a0 + na, :d[1].shape[1]] = d[1] + (a0 + 1)
        if sparse_membership:
            M[a0:a0 + na] = d[3].to(torch.float32).argmax(dim=1).to(torch.int32) + r0
        else:
            M[a0:a0 + na, r0:r0 + nr] = d[3]
        a0 += na
        r0 += nr
    return X, ids_topk, q, M


class StructuresDataset(torch.utils.data.Dataset):
    """PDB files -> preprocessed subunits, like src/dataset.py:115-156: `dataset[i]` returns `(subunits, path)` -- or
    `(structure, path)` without preprocessing, or `(None, path)` (after printing a ReadError line) when the file
    cannot be parsed.  Parsing uses the gemmi-free C++ reader (pesto_b200.structure_io.read_pdb)."""

    def __init__(self, pdb_filepaths, with_preprocessing=True):
        super().__init__()
        self.pdb_filepaths = pdb_filepaths
        self.with_preprocessing = with_preprocessing

    def __len__(self):
        return len(self.pdb_filepaths)

    def __getitem__(self, i):
        from .structure_io import read_pdb
        from .structure import preprocess_structure
        pdb_filepath = self.pdb_filepaths[i]
        try:
            structure = read_pdb(pdb_filepath)
        except Exception as e:                                    # noqa: BLE001 -- reference behaviour: report and go on
            print(f"ReadError: {pdb_filepath}: {e}")
            return None, pdb_filepath
        if self.with_preprocessing:
            return preprocess_structure(structure), pdb_filepath
        return structure, pdb_filepath
