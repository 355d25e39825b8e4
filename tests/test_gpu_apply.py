"""The whole apply path on the GPU against the reference's shipped outputs (examples/issue_19_04_2023: the files the
reference wrote and their md5 sums; tests/golden/pdb, made by tests/golden/make_golden_pdb.py)."""
import gzip
import hashlib
import os

import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
PDB = os.path.join(GOLDEN, "pdb")


def _gunzip_to(tmp_path, name):
    dst = os.path.join(tmp_path, name[:-3])
    with gzip.open(os.path.join(PDB, name), "rb") as fi, open(dst, "wb") as fo:
        fo.write(fi.read())
    return dst


def _bfactors(text):
    return [float(ln[60:66]) for ln in text.splitlines() if ln.startswith(("ATOM", "HETATM"))]


@pytest.mark.parametrize("mode", ["fp32", "bf16x3"])
def test_apply_reproduces_shipped_output_files(tmp_path, cuda_models, mode):
    """PDB file -> C++ reader -> preprocessing -> CUDA kNN + forward -> sigmoid -> save_pdb for 2CUA_A: byte-identical
    files (md5check.txt) in fp32 mode; in bf16x3 mode every record identical outside the two-decimal probability columns
    and every probability within one unit of the last printed digit."""
    from pesto_b200.apply import apply_to_pdb
    model = cuda_models("i_v4_1")
    src = _gunzip_to(str(tmp_path), "2CUA_A.pdb.gz")
    old = model.mode
    model.mode = mode
    try:
        paths, p = apply_to_pdb(model, src)
    finally:
        model.mode = old
    assert p.shape == (122, 5)
    md5 = dict(reversed(ln.split()) for ln in open(os.path.join(PDB, "md5check.txt")) if ln.strip())
    n_diff_lines = 0
    for i, path in enumerate(paths):
        got = open(path, "rb").read()
        exp = gzip.open(os.path.join(PDB, f"2CUA_A_i{i}.pdb.gz"), "rb").read()
        assert hashlib.md5(exp).hexdigest() == md5[f"2CUA_A_i{i}.pdb"]
        if mode == "fp32":
            assert hashlib.md5(got).hexdigest() == md5[f"2CUA_A_i{i}.pdb"], f"channel {i}"
        gl, el = got.decode().splitlines(), exp.decode().splitlines()
        assert len(gl) == len(el)
        for g, e in zip(gl, el):
            assert g[:54] == e[:54] and g[66:] == e[66:]
            n_diff_lines += g != e
        assert max(abs(a - b) for a, b in zip(_bfactors(got.decode()), _bfactors(exp.decode()))) <= 0.0100001
    assert n_diff_lines <= 10, n_diff_lines


@pytest.mark.parametrize("name", ["1ZNS", "7KHT_lipid"])
def test_apply_multichain_probabilities(tmp_path, cuda_models, name):
    """multi-chain (protein + DNA) and hetero (lipid) structures: the b-factor column of channel 0 matches the file the
    reference wrote to the printed precision (+-0.005 rounding, +-1e-3 logits)."""
    from pesto_b200.apply import apply_to_pdb
    model = cuda_models("i_v4_1")
    src = _gunzip_to(str(tmp_path), name + ".pdb.gz")
    paths, _ = apply_to_pdb(model, src)
    got = _bfactors(open(paths[0]).read())
    exp = [float(ln.split("|")[1]) for ln in gzip.open(os.path.join(PDB, name + "_i0.expected.gz"), "rt") if ln.startswith(("ATOM", "HETATM"))]
    assert len(got) == len(exp)
    assert max(abs(a - b) for a, b in zip(got, exp)) <= 0.0100001
