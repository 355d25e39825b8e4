#!/usr/bin/env python
"""Fixtures for the host side of the apply path (PDB read -> preprocessing -> save_pdb), taken from the reference's own
example files: inputs `examples/<dir>/<name>.pdb` and the files `<name>_i{0..4}.pdb` the reference wrote for them with
gemmi + apply_model.ipynb + i_v4_1/model_ckpt.pt.  They pin atom order, chain merging, altloc handling, hetero tagging
and residue renumbering (columns 1-54 and 77-78 of every record) and, through the b-factor column, the probabilities.

    python tests/golden/make_golden_pdb.py          # needs /root/reference; writes tests/golden/pdb/*.gz

For all cases the expected file keeps columns [0:54] + [60:66] + [76:78] of `_i0` (older example outputs carry a
different occupancy column, src/structure_io.py:117 writes the b-factor there since); for 2CUA_A
(examples/issue_19_04_2023, the md5check.txt fixture) the five output files are kept whole.
"""
import gzip
import os
import shutil

REF = "/root/reference/examples"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pdb")
CASES = ["endonuclease/1ZNS", "kinase/2VGO_A", "lipids/7KHT_lipid", "lipids/6I9F", "issue_19_04_2023/2CUA_A"]


def main():
    os.makedirs(OUT, exist_ok=True)
    for case in CASES:
        name = os.path.basename(case)
        with open(os.path.join(REF, case + ".pdb"), "rb") as fi, gzip.GzipFile(os.path.join(OUT, name + ".pdb.gz"), "wb", mtime=0) as fo:
            shutil.copyfileobj(fi, fo)
        with open(os.path.join(REF, case + "_i0.pdb")) as fi:
            keep = "".join((ln[:54] + "|" + ln[60:66] + "|" + ln[76:78]).rstrip("\n") + "\n" if ln.startswith(("ATOM", "HETATM"))
                           else ln for ln in fi)
        with gzip.GzipFile(os.path.join(OUT, name + "_i0.expected.gz"), "wb", mtime=0) as fo:
            fo.write(keep.encode())
    for i in range(5):
        src = os.path.join(REF, f"issue_19_04_2023/2CUA_A_i{i}.pdb")
        with open(src, "rb") as fi, gzip.GzipFile(os.path.join(OUT, f"2CUA_A_i{i}.pdb.gz"), "wb", mtime=0) as fo:
            shutil.copyfileobj(fi, fo)
    shutil.copy(os.path.join(REF, "issue_19_04_2023/md5check.txt"), os.path.join(OUT, "md5check.txt"))


if __name__ == "__main__":
    main()
