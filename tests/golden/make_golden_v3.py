#!/usr/bin/env python
"""Golden vectors for the two v3 checkpoints (model/save/i_v3_0_2021-05-27_14-27, i_v3_1_2021-05-28_12-40) by running
the UNMODIFIED reference in the build container -- each checkpoint with the model.py saved next to it (i_v3_1 has
single-Linear em / dm heads and one logit per residue), 123 input features (element | residue name | atom name one-hots,
model/save/i_v3_*/src/data_encoding.py:105-108), 16 layers.

    python tests/golden/make_golden_v3.py        ~1 min;  writes weights_i_v3_{0,1}.npz, config_i_v3_{0,1}.json, case_v3_1gpw_A.npz
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G                                                               # noqa: E402  (sets up the reference imports)

G.SAVE.update({"i_v3_0": "model/save/i_v3_0_2021-05-27_14-27", "i_v3_1": "model/save/i_v3_1_2021-05-28_12-40"})


def main():
    structure, _ = G.read_pdb_fixed(os.path.join(G.REF, "examples/double/1gpw_A.pdb"))
    structure = G.clean_structure(structure)
    X, M = G.encode_structure(structure)
    blocks = G.encode_features(structure)                                            # (qe [N,30], qr [N,29], qn [N,64])
    q = torch.cat(blocks, dim=1)
    ids0 = G.extract_topology(X, 64)[0]
    Xc, idsc, qc, Mc = G.collate_batch_features([[X, ids0, q, M]])
    d = dict(X=X.numpy().astype(np.float32), feat=np.stack([b.argmax(1).numpy() for b in blocks], 1).astype(np.uint8),
             rid=M.float().argmax(1).numpy().astype(np.int32), ids0=ids0.numpy().astype(np.int32), n_res=np.int32(M.shape[1]),
             element=np.asarray(structure["element"]), resname=np.asarray(structure["resname"]),
             name=np.asarray(structure["name"]), resid=np.asarray(structure["resid"]))
    assert all(bool((b.sum(1) == 1).all()) for b in blocks)                           # every block is a one-hot
    for tag in ("i_v3_0", "i_v3_1"):
        model, cfg, sd = G.load_reference_model(tag)
        np.savez_compressed(os.path.join(HERE, f"weights_{tag}.npz"), **{k: v.numpy() for k, v in sd.items()})
        with open(os.path.join(HERE, f"config_{tag}.json"), "w") as fh:
            json.dump(cfg, fh)
        with torch.no_grad():
            z = model(Xc, idsc, qc, Mc.float())               # (the layers themselves are pinned by the v4 cases' per-layer taps)
        d[f"z_{tag}"] = z.numpy()
        print(tag, "layers", len(cfg["sum"]), "N0", cfg["em"]["N0"], "z", tuple(z.shape))
    np.savez_compressed(os.path.join(HERE, "case_v3_1gpw_A.npz"), **d)


if __name__ == "__main__":
    main()
