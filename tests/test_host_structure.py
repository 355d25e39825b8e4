"""Host side of the apply path (no GPU): the gemmi-free PDB reader, the preprocessing chain and save_pdb, against the
reference's own example inputs / outputs (tests/golden/pdb, made by tests/golden/make_golden_pdb.py) and small crafted
records for the parser's edge cases (src/structure_io.py:6-55, src/structure.py:14-146)."""
import gzip
import os

import numpy as np
import pytest

from conftest import GOLDEN

PDB = os.path.join(GOLDEN, "pdb")


def _gunzip_to(tmp_path, name):
    dst = os.path.join(tmp_path, name[:-3])
    with gzip.open(os.path.join(PDB, name), "rb") as fi, open(dst, "wb") as fo:
        fo.write(fi.read())
    return dst


def _rec(kind, serial, name, alt, resn, chain, resid, icode, x, y, z, b=0.0, elem="", occ=1.0):
    return f"{kind:<6s}{serial:>5d} {name:<4s}{alt:1s}{resn:>3s} {chain:1s}{resid:>4d}{icode:1s}   {x:8.3f}{y:8.3f}{z:8.3f}{occ:6.2f}{b:6.2f}          {elem:>2s}\n"


@pytest.mark.parametrize("name", ["1ZNS", "2VGO_A", "7KHT_lipid", "6I9F", "2CUA_A"])
def test_host_chain_reproduces_reference_outputs(tmp_path, name):
    """read_pdb -> clean -> tag hetero -> split -> filter -> dedupe -> concatenate -> save_pdb gives, record by
    record, the atoms / names / chains / renumbered residue ids / coordinates / elements of the file the reference wrote."""
    from pesto_b200.dataset import StructuresDataset
    from pesto_b200.structure import concatenate_chains, encode_bfactor, split_by_chain
    from pesto_b200.structure_io import save_pdb
    src = _gunzip_to(str(tmp_path), name + ".pdb.gz")
    subunits, path = StructuresDataset([src], with_preprocessing=True)[0]
    assert path == src and subunits is not None
    s = concatenate_chains(subunits)
    s = encode_bfactor(s, np.zeros(np.unique(s["resid"]).shape[0], dtype=np.float32))
    out = os.path.join(str(tmp_path), "out.pdb")
    save_pdb(split_by_chain(s), out)
    got = open(out).read().splitlines()
    exp = gzip.open(os.path.join(PDB, name + "_i0.expected.gz"), "rt").read().splitlines()
    assert len(got) == len(exp)
    for g, e in zip(got, exp):
        if e.startswith(("ATOM", "HETATM")):
            head, _bf, elem = e.split("|")
            assert g[:54] == head and g[76:78].strip() == elem.strip(), (g, e)
            assert len(g) == 80
        else:
            assert g == e
    assert got[-1] == "END" and not open(out).read().endswith("\n")


def test_parser_edge_cases(tmp_path):
    """altloc duplicates (reference key: chain_resnum_name, first kept), insertion codes, separated chain parts merged
    behind the chain's first part, models -> '<chain>:<model>', blank element columns, END stops the read, 80-column cut."""
    from pesto_b200.structure_io import parse_pdb_text
    t = "HEADER    TEST\nMODEL        1\n"
    t += _rec("ATOM", 1, " N", " ", "ALA", "A", 1, " ", 0, 0, 0, elem="N")
    t += _rec("ATOM", 2, " CA", "A", "ALA", "A", 1, " ", 1, 0, 0, elem="C")
    t += _rec("ATOM", 3, " CA", "B", "ALA", "A", 1, " ", 1.1, 0, 0, elem="C")          # altloc duplicate: dropped
    t += _rec("ATOM", 4, " N", " ", "GLY", "A", 1, "A", 2, 0, 0, elem="N")              # insertion code
    t += "TER\n"
    t += _rec("ATOM", 5, " N", " ", "SER", "B", 7, " ", 3, 0, 0, elem="")               # element from the name columns
    t += _rec("HETATM", 6, "ZN", " ", " ZN", "A", 201, " ", 4, 0, 0, elem="ZN")         # chain A again: moved before B
    t += _rec("HETATM", 7, " O", " ", "HOH", "A", 301, " ", 5, 0, 0, elem="O").rstrip("\n") + "   trailing beyond col 80\n"
    t += "ENDMDL\nMODEL        2\n"
    t += _rec("ATOM", 1, " CA", "A", "ALA", "A", 1, " ", 9, 0, 0, elem="C")             # key seen in model 1: dropped
    t += _rec("ATOM", 2, " CB", " ", "ALA", "A", 1, " ", 9, 1, 0, elem="C")
    t += "ENDMDL\nEND\n"
    t += _rec("ATOM", 9, " X", " ", "ALA", "Z", 1, " ", 0, 0, 0, elem="C")              # after END: ignored
    s = parse_pdb_text(t)
    assert list(s["chain_name"]) == ["A:0"] * 5 + ["B:0"] + ["A:1"]
    assert list(s["name"]) == ["N", "CA", "N", "ZN", "O", "N", "CB"]
    assert list(s["element"]) == ["N", "C", "N", "Zn", "O", "N", "C"]
    assert list(s["icode"]) == ["", "", "A", "", "", "", ""]
    assert list(s["het_flag"]) == ["A", "A", "A", "H", "H", "A", "A"]
    assert list(s["resid"]) == [1, 1, 1, 201, 301, 7, 1] and s["resid"].dtype == np.int32
    assert s["xyz"].dtype == np.float32 and np.allclose(s["xyz"][:, 0], [0, 1, 2, 4, 5, 3, 9])
    empty = parse_pdb_text("REMARK nothing\nEND\n")
    assert empty["xyz"].shape == (0, 3) and empty["name"].shape == (0,)


def test_preprocessing_semantics():
    """clean_structure renumbers along the atom order over chain / number / icode changes and drops H, D, HOH, DOD;
    tag_hetatm_chains names hetero residues '<chain>:<model>:<k>'; one-atom-per-residue subunits and duplicated hetero
    subunits are removed (src/structure.py:14-146)."""
    from pesto_b200.structure import (clean_structure, filter_non_atomic_subunits, remove_duplicate_tagged_subunits,
                                      split_by_chain, tag_hetatm_chains)
    n = 10
    s = {
        "xyz": np.arange(3 * n, dtype=np.float32).reshape(n, 3),
        "name": np.array(["N", "CA", "H", "N", "CA", "O", "ZN", "ZN", "C1", "C2"]),
        "element": np.array(["N", "C", "H", "N", "C", "O", "Zn", "Zn", "C", "C"]),
        "resname": np.array(["ALA", "ALA", "ALA", "GLY", "GLY", "HOH", "ZN", "ZN", "LIG", "LIG"]),
        "resid": np.array([5, 5, 5, 5, 5, 9, 20, 21, 30, 30], dtype=np.int32),
        "het_flag": np.array(["A"] * 5 + ["H"] * 5),
        "chain_name": np.array(["A:0"] * 10),
        "icode": np.array(["", "", "", "A", "A", "", "", "", "", ""]),
    }
    s["xyz"][7] = s["xyz"][6] + 0.05                        # the second zinc sits on the first
    c = clean_structure({k: v.copy() for k, v in s.items()})
    assert "icode" not in c and list(c["name"]) == ["N", "CA", "N", "CA", "ZN", "ZN", "C1", "C2"]
    assert list(c["resid"]) == [1, 1, 2, 2, 3, 4, 5, 5]
    t = tag_hetatm_chains(c)
    assert list(t["chain_name"]) == ["A:0"] * 4 + ["A:0:0", "A:0:1", "A:0:2", "A:0:2"]
    sub = split_by_chain(t)
    assert list(sub) == ["A:0", "A:0:0", "A:0:1", "A:0:2"] and "chain_name" not in sub["A:0"]
    sub = remove_duplicate_tagged_subunits(filter_non_atomic_subunits(sub))
    assert list(sub) == ["A:0", "A:0:0", "A:0:2"]
    w = clean_structure({k: v.copy() for k, v in s.items()}, rm_wat=False)
    assert "HOH" in w["resname"] and len(w["xyz"]) == 9


def test_encode_bfactor_and_compat_imports():
    """per-residue / per-atom values -> 'bfactor' (src/structure.py:185-223); the reference's import lines resolve after
    compat.install() (apply_model.ipynb:21-24, 66-73)."""
    import pesto_b200.compat as compat
    from pesto_b200.structure import encode_bfactor
    s = {"name": np.array(["N", "CA", "CA", "O"]), "element": np.array(["N", "C", "C", "O"]),
         "het_flag": np.array(["A"] * 4), "resid": np.array([1, 1, 2, 2])}
    assert list(encode_bfactor(dict(s), np.array([0.25, 0.5], dtype=np.float32))["bfactor"]) == [0.25, 0.25, 0.5, 0.5]
    assert list(encode_bfactor(dict(s), np.arange(4.0))["bfactor"]) == [0, 1, 2, 3]
    compat.install()
    from src.dataset import StructuresDataset, collate_batch_features, select_by_sid          # noqa: F401
    from src.data_encoding import encode_structure, encode_features, extract_topology          # noqa: F401
    from src.structure import data_to_structure, encode_bfactor as eb, concatenate_chains      # noqa: F401
    from src.structure_io import save_pdb, read_pdb                                            # noqa: F401
    from model import Model                                                                    # noqa: F401
    from src.scoring import bc_scoring, bc_score_names                                         # noqa: F401  (apply_model.ipynb:25)
    assert len(bc_score_names) == 8
    with pytest.raises(NotImplementedError):
        select_by_sid(None, None)


REF_MODEL_DIR = "/root/reference/model/save/i_v4_1_2021-09-07_11-21"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_MODEL_DIR, "model_ckpt.pt")), reason="reference checkout not present")
def test_load_model_from_the_reference_checkout():
    """apply_model.ipynb cells 2-4 against the reference's own files: its config.py (which imports src.data_encoding at module
    level) and the shipped checkpoint load strictly into pesto_b200.Model (1 474 957 parameters)."""
    from pesto_b200.apply import load_model
    from pesto_b200.data_encoding import categ_to_resnames, resname_to_categ
    model = load_model(REF_MODEL_DIR, device="cpu")
    assert sum(p.numel() for p in model.parameters()) == 1474957
    assert len(model.config["sum"]) == 32 and model.mode == "f16x3"
    assert resname_to_categ["ZN"] == "ion" and len(categ_to_resnames["protein"]) == 20 and len(resname_to_categ) == 79


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_MODEL_DIR, "model_ckpt.pt")), reason="reference checkout not present")
def test_apply_notebook_cells_run_unchanged_up_to_the_device_call(monkeypatch):
    """apply_model.ipynb cells 0 and 2-6 (the lines that touch this path), executed as written after compat.install(): imports,
    `from config import config_model` from the reference's model directory, Model + load_state_dict, StructuresDataset,
    encoding -- and extract_topology either runs (GPU) or fails loudly (no GPU): there is no CPU fallback."""
    import sys
    import torch as pt
    import pesto_b200.compat as compat
    from pesto_b200 import _lib
    compat.install()
    monkeypatch.syspath_prepend(REF_MODEL_DIR)
    sys.modules.pop("config", None)
    from src.dataset import StructuresDataset, collate_batch_features                          # noqa: F401
    from src.data_encoding import encode_structure, encode_features, extract_topology
    from src.structure import concatenate_chains
    from config import config_model
    from model import Model
    model = Model(config_model)
    res = model.load_state_dict(pt.load(os.path.join(REF_MODEL_DIR, "model_ckpt.pt"), map_location=pt.device("cpu")))
    assert str(res) == "<All keys matched successfully>"
    model = model.eval().to(pt.device("cpu"))
    fp = "/root/reference/examples/issue_19_04_2023/2CUA_A.pdb"
    subunits, filepath = StructuresDataset([fp], with_preprocessing=True)[0]
    structure = concatenate_chains(subunits)
    X, M = encode_structure(structure)
    q = encode_features(structure)[0]
    assert X.shape == (955, 3) and M.shape == (955, 122) and q.shape == (955, 30)
    if not pt.cuda.is_available():
        with pytest.raises(_lib.PestoError):
            extract_topology(X, 64)
    sys.modules.pop("config", None)
