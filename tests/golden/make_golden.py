#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference, CPU, fp32) in the build container.  The reference cannot travel to the
GPU box, so its inputs/outputs are committed here as small .npz fixtures.

    python tests/golden/make_golden.py small      # weights + small cases (+ per-layer taps)   ~2 min
    python tests/golden/make_golden.py bench53 bench53_ties   # the 53 pdbs_test structures (config 2)   ~25 min
    python tests/golden/make_golden.py synth8192  # synthetic N=8192 (config 3/4 shape)         ~3 min

Nothing from the reference's sources is copied: the script imports it in place.
PDB parsing is a fixed-column reader (gemmi is not installed), as SURVEY.md 8c describes.
"""
import importlib.util
import json
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
SAVE = {"i_v4_1": "model/save/i_v4_1_2021-09-07_11-21", "i_v4_0": "model/save/i_v4_0_2021-09-07_11-20"}

sys.path.insert(0, REPO)
sys.path.insert(0, REF)
sys.modules.setdefault("gemmi", types.ModuleType("gemmi"))       # src/dataset.py -> structure_io -> import gemmi
sys.modules["gemmi"].cif = types.ModuleType("gemmi.cif")
sys.modules.setdefault("gemmi.cif", sys.modules["gemmi"].cif)

from src.data_encoding import encode_structure, encode_features, extract_topology   # noqa: E402  (reference)
from src.dataset import collate_batch_features                                       # noqa: E402  (reference)
from src.structure import clean_structure                                            # noqa: E402  (reference)
from src.scoring import bc_scoring, bc_score_names                                   # noqa: E402  (reference)
from pesto_b200.synth import synth_structure, one_hot_features, dense_membership, BASE_SEED   # noqa: E402


def load_reference_model(tag):
    save = os.path.join(REF, SAVE[tag])
    spec = importlib.util.spec_from_file_location(f"refcfg_{tag}", os.path.join(save, "config.py"))
    cfg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cfg)
    spec = importlib.util.spec_from_file_location(f"refmodel_{tag}", os.path.join(save, "model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    model = mod.Model(cfg.config_model)
    sd = torch.load(os.path.join(save, "model_ckpt.pt"), map_location="cpu")
    model.load_state_dict(sd)
    return model.eval(), cfg.config_model, sd


def read_pdb_fixed(path):
    """Fixed-column ATOM/HETATM reader -> the dict `read_pdb` (src/structure_io.py:6-55) returns."""
    rows = []
    with open(path) as fh:
        for line in fh:
            if line[:4] == "ATOM" or line[:6] == "HETATM":
                rows.append((line[12:16].strip(), line[17:20].strip(), int(line[22:26]), line[26].strip(),
                             float(line[30:38]), float(line[38:46]), float(line[46:54]), float(line[60:66]),
                             line[76:78].strip().title(), "H" if line[:6] == "HETATM" else "A", line[21]))
    return {
        "xyz": np.array([[r[4], r[5], r[6]] for r in rows], dtype=np.float32),
        "name": np.array([r[0] for r in rows]),
        "element": np.array([r[8] for r in rows]),
        "resname": np.array([r[1] for r in rows]),
        "resid": np.array([r[2] for r in rows], dtype=np.int32),
        "het_flag": np.array([r[9] for r in rows]),
        "chain_name": np.array([r[10] + ":0" for r in rows]),
        "icode": np.array([r[3] for r in rows]),
    }, np.array([r[7] for r in rows], dtype=np.float32)


def structure_inputs(structure):
    """The apply_model.ipynb cell-6 preparation for one structure (reference code)."""
    X, M = encode_structure(structure)
    q = encode_features(structure)[0]
    ids0 = extract_topology(X, 64)[0]
    return X, ids0, q, M


def compact(X, ids0, q, M):
    """Inputs in compact form: element class and residue column instead of one-hots."""
    return dict(X=X.numpy().astype(np.float32), el=q.argmax(1).numpy().astype(np.uint8),
                rid=M.float().argmax(1).numpy().astype(np.int32), ids0=ids0.numpy().astype(np.int32),
                n_res=np.int32(M.shape[1]))


def forward_with_taps(model, Xc, idsc, qc, Mc, tap_layers):
    taps = {}
    hooks = []
    for li in tap_layers:
        def mk(li):
            def hook(_m, _inp, out):
                taps[li] = (out[0].detach().clone(), out[1].detach().clone())
            return hook
        hooks.append(model.sum[li].register_forward_hook(mk(li)))
    with torch.no_grad():
        z = model(Xc, idsc, qc, Mc.float())
    for h in hooks:
        h.remove()
    return z, taps


def part_small():
    t0 = time.time()
    out = {}
    models = {}
    for tag in SAVE:
        model, cfg, sd = load_reference_model(tag)
        models[tag] = (model, cfg)
        np.savez_compressed(os.path.join(HERE, f"weights_{tag}.npz"),
                            **{k: v.numpy() for k, v in sd.items()})
        with open(os.path.join(HERE, f"config_{tag}.json"), "w") as fh:
            json.dump(cfg, fh)
    print("weights saved", time.time() - t0)

    # ---- real structures with published outputs (examples/*_i{0..4}.pdb) --------------------------------
    for name, rel in (("2CUA_A", "examples/issue_19_04_2023/2CUA_A"), ("1gpw_A", "examples/double/1gpw_A")):
        structure, _ = read_pdb_fixed(os.path.join(REF, rel + ".pdb"))
        structure = clean_structure(structure)
        X, ids0, q, M = structure_inputs(structure)
        Xc, idsc, qc, Mc = collate_batch_features([[X, ids0, q, M]])
        d = compact(X, ids0, q, M)
        for tag, (model, cfg) in models.items():
            L = len(cfg["sum"])
            tl = sorted(set([0, 1, L // 4, L // 2, L - 1])) if name == "2CUA_A" else []
            z, taps = forward_with_taps(model, Xc, idsc, qc, Mc, tl)
            d[f"z_{tag}"] = z.numpy()
            for li, (qq, pp) in taps.items():
                d[f"tap_{tag}_L{li}_q"] = qq.numpy()
                d[f"tap_{tag}_L{li}_p"] = pp.numpy()
        # published per-residue probabilities: b-factor of the first atom of each residue, 5 channels
        first = np.concatenate([[True], np.diff(d["rid"]) != 0])
        bf = []
        for c in range(5):
            _, b = read_pdb_fixed(os.path.join(REF, f"{rel}_i{c}.pdb"))
            bf.append(b[first])
        d["published_prob"] = np.stack(bf, 1).astype(np.float32)
        p = torch.sigmoid(torch.from_numpy(d["z_i_v4_1"])).numpy()
        print(name, "N", X.shape[0], "R", M.shape[1], "max|sigmoid(z)-published|", np.abs(p - d["published_prob"]).max())
        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **d)

    # ---- config 1: pdbs_test/EW_1EWY_1_A:0.pdb ------------------------------------------------------------
    structure, _ = read_pdb_fixed(os.path.join(REF, "pdbs_test/EW_1EWY_1_A:0.pdb"))
    structure.pop("icode")
    X, ids0, q, M = structure_inputs(structure)
    Xc, idsc, qc, Mc = collate_batch_features([[X, ids0, q, M]])
    d = compact(X, ids0, q, M)
    with torch.no_grad():
        d["z_i_v4_1"] = models["i_v4_1"][0](Xc, idsc, qc, Mc.float()).numpy()
    np.savez_compressed(os.path.join(HERE, "case_1EWY.npz"), **d)
    print("1EWY done", time.time() - t0)

    # ---- synthetic small cases: sink-padded (<64 atoms), batch of 2, odd sizes ----------------------------
    def synth_case(n, seed):
        X, el, rid = synth_structure(n, seed)
        q = one_hot_features(el)
        M = dense_membership(rid)
        ids0 = extract_topology(X, 64)[0]
        return X, ids0, q, M

    for name, specs in (("tiny40", [(40, BASE_SEED + 1)]), ("synth517", [(517, BASE_SEED + 2)]),
                        ("batch3", [(300, BASE_SEED + 3), (33, BASE_SEED + 4), (129, BASE_SEED + 5)])):
        parts = [synth_case(n, s) for n, s in specs]
        Xc, idsc, qc, Mc = collate_batch_features([list(p) for p in parts])
        d = dict(sizes=np.array([p[0].shape[0] for p in parts], dtype=np.int32),
                 X=Xc.numpy(), el=qc.argmax(1).numpy().astype(np.uint8),
                 rid=Mc.argmax(1).numpy().astype(np.int32), n_res=np.int32(Mc.shape[1]),
                 ids1=idsc.numpy().astype(np.int32))
        for i, p in enumerate(parts):
            d[f"ids0_{i}"] = p[1].numpy().astype(np.int32)
        for tag, (model, cfg) in models.items():
            L = len(cfg["sum"])
            z, taps = forward_with_taps(model, Xc, idsc, qc, Mc, [0, L - 1])
            d[f"z_{tag}"] = z.numpy()
            for li, (qq, pp) in taps.items():
                d[f"tap_{tag}_L{li}_q"] = qq.numpy()
                d[f"tap_{tag}_L{li}_p"] = pp.numpy()
        np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **d)
        print(name, "done", time.time() - t0)

    # ---- MD-style stale topology: frame-0 ids reused on perturbed coordinates (md_analysis cell 6) ---------
    X, ids0, q, M = synth_case(257, BASE_SEED + 6)
    g = torch.Generator().manual_seed(7)
    X1 = (X + 0.3 * torch.randn(X.shape, generator=g)).contiguous()
    Xc, idsc, qc, Mc = collate_batch_features([[X1, ids0, q, M]])
    d = compact(X1, ids0, q, M)
    with torch.no_grad():
        d["z_i_v4_0"] = models["i_v4_0"][0](Xc, idsc, qc, Mc.float()).numpy()
    np.savez_compressed(os.path.join(HERE, "case_stale257.npz"), **d)
    print("small part done", time.time() - t0)


def part_bench53():
    t0 = time.time()
    model, cfg, _ = load_reference_model("i_v4_1")
    nb = json.load(open(os.path.join(REF, "interface_ppi_benchmark.ipynb")))
    table = None
    for c in nb["cells"]:
        for o in c.get("outputs", []):
            t = "".join(o.get("text", []))
            if "acc=" in t and "auc=" in t and t.count("\n") >= 50:
                table = [ln for ln in t.split("\n") if "acc=" in ln]
    assert table is not None and len(table) == 53, "metric table not found"
    keys = [ln.split(",")[0] for ln in table]
    Xs, els, rids, ys, zs, sizes, nres, lines = [], [], [], [], [], [], [], []
    for i, key in enumerate(keys):
        base = os.path.join(REF, "pdbs_test", key.replace("/", "_"))
        structure, _ = read_pdb_fixed(base + ".pdb")
        structure.pop("icode")
        _, btrue = read_pdb_fixed(base + "_T.pdb")
        X, ids0, q, M = structure_inputs(structure)
        Xc, idsc, qc, Mc = collate_batch_features([[X, ids0, q, M]])
        with torch.no_grad():
            z = model(Xc, idsc, qc, Mc.float())
        rid = M.float().argmax(1).numpy()
        first = np.concatenate([[True], np.diff(rid) != 0])
        y = torch.from_numpy((btrue[first] > 0.5).astype(np.float32)).unsqueeze(1)
        p = torch.sigmoid(z)
        scores = bc_scoring(y, p[:, :1])[:, 0]
        line = ", ".join([key] + [f"{bc_score_names[j]}={scores[j]:.3f}" for j in range(scores.shape[0])])
        ok = (line == table[i])
        print(f"[{i:2d}] {time.time() - t0:6.0f}s match={ok} {line}", flush=True)
        Xs.append(X.numpy()); els.append(q.argmax(1).numpy().astype(np.uint8)); rids.append(rid.astype(np.int32))
        ys.append(y.numpy()[:, 0].astype(np.uint8)); zs.append(z.numpy()); sizes.append(X.shape[0]); nres.append(M.shape[1])
        lines.append(line)
    np.savez_compressed(os.path.join(HERE, "pdbs_test_53.npz"),
                        keys=np.array(keys), sizes=np.array(sizes, np.int32), n_res=np.array(nres, np.int32),
                        X=np.concatenate(Xs), el=np.concatenate(els), rid=np.concatenate(rids),
                        y=np.concatenate(ys), z_i_v4_1=np.concatenate(zs),
                        table_published=np.array(table), table_reference_here=np.array(lines))
    print("bench53 done", time.time() - t0)


def part_bench53_ties():
    """torch.topk leaves the order inside exact-distance tie groups unspecified; the CUDA kernel and the oracle
    use (distance, index).  Record the rows of the 53 structures where the reference's CPU topk order differs,
    so that forward parity can be tested on exactly the neighbour lists the reference used."""
    from oracle import pesto_oracle as O
    path = os.path.join(HERE, "pdbs_test_53.npz")
    g = dict(np.load(path))
    aoff = np.concatenate([[0], np.cumsum(g["sizes"])])
    st, rows, ids, straddle = [], [], [], []
    for i in range(len(g["keys"])):
        X = torch.from_numpy(g["X"][aoff[i]:aoff[i + 1]])
        r = extract_topology(X, 64)[0]
        o = O.extract_topology(X, 64)[0]
        for row in (r != o).any(1).nonzero()[:, 0].tolist():
            st.append(i); rows.append(row); ids.append(r[row].numpy().astype(np.int32))
            straddle.append(any(set(r[row, :n].tolist()) != set(o[row, :n].tolist()) for n in (8, 16, 32, 64)))
    g.update(tie_struct=np.array(st, np.int32), tie_row=np.array(rows, np.int32), tie_ids=np.stack(ids),
             tie_straddles_prefix=np.array(straddle))
    np.savez_compressed(path, **g)
    print("tie rows:", len(rows), "straddling a prefix boundary:", int(np.sum(straddle)))


def part_synth8192():
    t0 = time.time()
    model, cfg, _ = load_reference_model("i_v4_1")
    X, el, rid = synth_structure(8192, BASE_SEED)
    q = one_hot_features(el)
    M = dense_membership(rid)
    ids0 = extract_topology(X, 64)[0]
    print("knn", time.time() - t0)
    Xc, idsc, qc, Mc = collate_batch_features([[X, ids0, q, M]])
    with torch.no_grad():
        z = model(Xc, idsc, qc, Mc.float())
    np.savez_compressed(os.path.join(HERE, "case_synth8192.npz"), ids0=ids0.numpy().astype(np.int16),
                        z_i_v4_1=z.numpy(), n_atoms=np.int32(8192), seed=np.int64(BASE_SEED))
    print("synth8192 done", time.time() - t0)


if __name__ == "__main__":
    torch.set_num_threads(int(os.environ.get("GOLDEN_THREADS", os.cpu_count())))
    for part in sys.argv[1:]:
        {"small": part_small, "bench53": part_bench53, "bench53_ties": part_bench53_ties, "synth8192": part_synth8192}[part]()
