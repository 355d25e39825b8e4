"""PDB input/output with the reference's names and dictionaries (src/structure_io.py of LBM-EPFL/PeSTo), gemmi-free.

`read_pdb` parses the file with the C++ fixed-column parser of the shared library (`pesto_pdb_parse_host`); `save_pdb`
writes exactly the record layout of src/structure_io.py:96-123, so that the files of the apply path come out byte
for byte (fixture: examples/issue_19_04_2023/md5check.txt).
"""
import ctypes

import numpy as np

from . import _lib

# one ATOM/HETATM record of save_pdb (src/structure_io.py:117): serial, name, resname, chain, resid, xyz, occupancy
# and b-factor (both carry the value), element
_RECORD = "{:<6s}{:>5d} {:<4s} {:>3s} {:1s}{:>4d}    {:8.3f}{:8.3f}{:8.3f}{:6.2f}{:6.2f}          {:<2s}  \n"


def _chars(buf, n, width):
    """fixed-width NUL-padded character fields -> numpy unicode array"""
    return np.frombuffer(buf, dtype=f"S{width}", count=n).astype(f"U{width}")      # (C loop; np.char.decode goes through Python per element)


def parse_pdb_text(text):
    """Structure dictionary of a PDB file's text: the keys and dtypes read_pdb of the reference returns."""
    lib = _lib.load()
    raw = text.encode("ascii", errors="replace") if isinstance(text, str) else bytes(text)
    cap = lib.pesto_pdb_count_atoms_host(raw, len(raw))
    cap1 = max(cap, 1)
    xyz = np.empty((cap1, 3), dtype=np.float32)
    bfac = np.empty(cap1, dtype=np.float32)
    resid = np.empty(cap1, dtype=np.int32)
    model = np.empty(cap1, dtype=np.int32)
    name, elem, resn = (ctypes.create_string_buffer(w * cap1) for w in (4, 2, 3))
    het, chain, icode = (ctypes.create_string_buffer(cap1) for _ in range(3))
    n = ctypes.c_int(0)
    rc = lib.pesto_pdb_parse_host(raw, len(raw), cap, xyz.ctypes.data, ctypes.addressof(name), ctypes.addressof(elem),
                                  ctypes.addressof(resn), resid.ctypes.data, ctypes.addressof(het),
                                  ctypes.addressof(chain), model.ctypes.data, ctypes.addressof(icode),
                                  bfac.ctypes.data, ctypes.byref(n))
    _lib.check(rc, "pesto_pdb_parse_host")
    n = n.value
    # '<chain>:<model index>' per atom: format once per distinct (chain, model) pair, not once per atom
    pair = np.frombuffer(chain.raw, dtype=np.uint8, count=n).astype(np.int64) * (1 << 32) + model[:n].astype(np.int64)
    upair, inv = np.unique(pair, return_inverse=True)
    chain_names = np.array([f"{chr(int(u >> 32))}:{int(u & 0xffffffff)}" for u in upair], dtype=str)[inv] if n else np.array([], dtype=str)
    return {
        "xyz": xyz[:n].copy(),
        "name": _chars(name.raw, n, 4),
        "element": _chars(elem.raw, n, 2),
        "resname": _chars(resn.raw, n, 3),
        "resid": resid[:n].copy(),
        "het_flag": _chars(het.raw, n, 1),
        "chain_name": chain_names,
        "icode": _chars(icode.raw, n, 1),
        "bfactor": bfac[:n].copy(),          # extension: the reference drops it; the apply path never reads it
    }


def read_pdb(pdb_filepath):
    """src/structure_io.py:6-55: dict with xyz float32 [N,3], name, element, resname, resid int32, het_flag ('A'/'H'),
    chain_name '<chain>:<model index>', icode ('' when absent)."""
    with open(pdb_filepath, "rb") as fh:
        structure = parse_pdb_text(fh.read())
    structure.pop("bfactor")
    return structure


def save_pdb(subunits, filepath):
    """src/structure_io.py:96-123: one ATOM/HETATM record per atom, serials restart per subunit, the subunit key's
    first character is the chain id, TER after every subunit, END without a trailing newline."""
    out = []
    for cn, su in subunits.items():
        c = cn.split(":")[0][0]
        bf = su["bfactor"] if "bfactor" in su else None
        for i in range(su["xyz"].shape[0]):
            b = bf[i] if bf is not None else 0.0
            x = su["xyz"][i]
            out.append(_RECORD.format("ATOM" if su["het_flag"][i] == "A" else "HETATM", i + 1, su["name"][i], su["resname"][i],
                                      c, su["resid"][i], x[0], x[1], x[2], b, b, su["element"][i]))
        out.append("TER\n")
    out.append("END")
    with open(filepath, "w") as fs:
        fs.write("".join(out))


def save_traj_pdb(subunits, filepath):
    """src/structure_io.py:126-159: MODEL k ... ENDMDL per frame of xyz [frames, N, 3]; occupancy 0, b-factor column."""
    frames = {su["xyz"].shape[0] for su in subunits.values()}
    assert all(su["xyz"].ndim == 3 for su in subunits.values()), "no time dimension"
    assert len(frames) == 1, "mismatching number of frames"
    out = []
    for k in range(frames.pop()):
        out.append("MODEL    {:>4d}\n".format(k))
        for cn, su in subunits.items():
            bf = su["bfactor"] if "bfactor" in su else None
            for i in range(su["xyz"].shape[1]):
                x = su["xyz"][k][i]
                out.append(_RECORD.format("ATOM" if su["het_flag"][i] == "A" else "HETATM", i + 1, su["name"][i],
                                          su["resname"][i], cn, su["resid"][i], x[0], x[1], x[2], 0.0,
                                          bf[i] if bf is not None else 0.0, su["element"][i]))
            out.append("TER\n")
        out.append("ENDMDL\n")
    out.append("END")
    with open(filepath, "w") as fs:
        fs.write("".join(out))
