"""The apply path of the reference (apply_model.ipynb cells 2-6; same steps as profiling.py:68-119) on the CUDA kernels:
PDB file -> preprocessed subunits -> features + kNN topology -> Model.forward -> sigmoid -> per-residue probabilities
written to the b-factor column of `<name>_i{0..4}.pdb`.

    python -m pesto_b200.apply --model-dir <.../model/save/i_v4_1_2021-09-07_11-21> [--mode fp32] file.pdb [...]

Library use: `load_model(model_dir)`, `predict_structure(model, structure)`, `apply_to_pdb(model, path)`.
"""
import argparse
import importlib.util
import os

import numpy as np
import torch

from .data_encoding import encode_features, encode_structure, extract_topology
from .dataset import StructuresDataset, collate_batch_features
from .model import Model
from .structure import concatenate_chains, encode_bfactor, split_by_chain
from .structure_io import save_pdb


def load_model(model_dir, checkpoint="model_ckpt.pt", mode="f16x3", device="cuda"):
    """Model(config_model) + load_state_dict of the shipped checkpoint (apply_model.ipynb cells 2-4); the depth of the em / dm
    heads follows the checkpoint (the model.py saved with i_v3_1 has single Linear layers)."""
    from . import compat
    compat.install()          # the reference's config.py imports `src.data_encoding` at module level
    spec = importlib.util.spec_from_file_location("pesto_reference_config", os.path.join(model_dir, "config.py"))
    cfg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cfg)
    model = Model.for_state_dict(cfg.config_model, torch.load(os.path.join(model_dir, checkpoint), map_location="cpu"), mode=mode)
    return model.eval().to(device)


def predict_structure(model, structure, device="cuda"):
    """Per-residue probabilities p[n_res, N2] of one concatenated structure (apply_model.ipynb cell 6, lines 144-158)."""
    X, M = encode_structure(structure)
    feats = encode_features(structure)
    # v4 models: element one-hot; v3 models: element | residue name | atom name (model/save/i_v3_*/src/data_encoding.py:105-108)
    q = feats[0] if model.config["em"]["N0"] == feats[0].shape[1] else torch.cat(feats, dim=1)
    with torch.no_grad():
        X = X.to(device)
        ids_topk = extract_topology(X, 64)[0]
        X, ids_topk, q, M = collate_batch_features([[X, ids_topk, q.to(device), M.to(device)]])
        z = model(X, ids_topk, q, M.float())
        p = torch.sigmoid(z).cpu().numpy()
        model.raise_if_failed(X.device)        # (the copy above synchronised: device-side input / watchdog flags -> PestoError)
        return p


def apply_to_pdb(model, pdb_filepath, device="cuda", out_prefix=None):
    """One file through the whole path; writes `<prefix>_i{k}.pdb` for the 5 interface types and returns (paths, p)."""
    subunits, _ = StructuresDataset([pdb_filepath], with_preprocessing=True)[0]
    if subunits is None:
        raise ValueError(f"cannot read {pdb_filepath}")
    structure = concatenate_chains(subunits)
    p = predict_structure(model, structure, device)
    prefix = out_prefix or pdb_filepath[:-4]
    paths = []
    for i in range(p.shape[1]):
        structure = encode_bfactor(structure, p[:, i])
        paths.append(f"{prefix}_i{i}.pdb")
        save_pdb(split_by_chain(structure), paths[-1])
    return paths, p


def main():
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--model-dir", required=True, help="directory holding config.py and model_ckpt.pt")
    ap.add_argument("--mode", default="f16x3", choices=["fp32", "f16x3", "f16", "bf16x3", "bf16"])
    ap.add_argument("pdb", nargs="+")
    a = ap.parse_args()
    model = load_model(a.model_dir, mode=a.mode)
    for fp in a.pdb:
        try:                                                    # per-structure errors do not stop the run
            paths, p = apply_to_pdb(model, fp)
            print(f"{fp}: {p.shape[0]} residues -> {os.path.basename(paths[0])} .. {os.path.basename(paths[-1])}")
        except Exception as e:                                  # noqa: BLE001
            print(f"{fp}: FAILED: {e}")


if __name__ == "__main__":
    main()
