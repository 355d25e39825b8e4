"""Structure preprocessing of the apply path with the reference's names (src/structure.py of LBM-EPFL/PeSTo).

Plain numpy over the structure dictionaries `read_pdb` returns; O(N) host work on either side of the CUDA path.
"""
import numpy as np

res3to1 = dict(CYS="C", ASP="D", SER="S", GLN="Q", LYS="K", ILE="I", PRO="P", THR="T", PHE="F", ASN="N",
               GLY="G", HIS="H", LEU="L", ARG="R", TRP="W", ALA="A", VAL="V", GLU="E", TYR="Y", MET="M")
res1to3 = {v: k for k, v in res3to1.items()}


def _changes(values):
    """1 where values[i] != values[i-1] (0 for the first element)"""
    v = np.asarray(values)
    out = np.zeros(v.shape[0], dtype=np.int64)
    if v.shape[0] > 1:
        out[1:] = v[1:] != v[:-1]
    return out


def clean_structure(structure, rm_wat=True):
    """src/structure.py:14-56: drop hydrogens / deuterium / heavy water (and water unless rm_wat=False, which tags it
    with resid -999 instead), renumber residues 1.. along the atom order (a new residue starts wherever chain, residue
    number or insertion code changes), drop the 'icode' key."""
    resname, element = structure["resname"], structure["element"]
    water = resname == "HOH"
    drop = (element == "H") | (element == "D") | (resname == "DOD")
    if rm_wat:
        drop = drop | water
    else:
        structure["resid"][water] = -999
    keep = ~drop
    s = {k: v[keep] for k, v in structure.items()}
    new_res = (_changes(s["chain_name"]) + _changes(s["resid"]) + _changes(s["icode"])) > 0
    s["resid"] = np.cumsum(new_res.astype(np.int64)) + 1
    s.pop("icode")
    return s


def atom_select(structure, sel):
    return {k: v[sel] for k, v in structure.items()}


def split_by_chain(structure):
    """src/structure.py:63-80: {chain name: atoms of that chain}, keys in sorted order, 'chain_name' removed."""
    names = structure["chain_name"]
    chains = {}
    for cn in np.unique(names):
        chain = atom_select(structure, names == cn)
        chain.pop("chain_name")
        chains[cn] = chain
    return chains


def concatenate_chains(chains):
    """src/structure.py:83-93: concatenation over the keys all chains share, plus 'chain_name'."""
    keys = set.intersection(*[set(c) for c in chains.values()])
    s = {k: np.concatenate([c[k] for c in chains.values()]) for k in keys}
    s["chain_name"] = np.concatenate([np.array([cid] * c["xyz"].shape[0]) for cid, c in chains.items()])
    return s


def tag_hetatm_chains(structure):
    """src/structure.py:96-110: every run of HETATM atoms with one residue id becomes its own chain
    '<chain>:<model>:<running index over the structure's hetero residues>'."""
    het = structure["het_flag"] == "H"
    run = np.cumsum(_changes(structure["resid"][het]))
    cids = structure["chain_name"].astype("<U10")
    cids[het] = np.array([f"{c}:{h}" for c, h in zip(structure["chain_name"][het], run)], dtype="<U10") if het.any() else cids[het]
    structure["chain_name"] = np.array(list(cids)).astype(str)
    return structure


def remove_duplicate_tagged_subunits(subunits):
    """src/structure.py:113-135: of two tagged (hetero) subunits with equal atom counts whose closest pair of
    corresponding atoms is nearer than 0.2 A, the later one is removed."""
    tagged = [cid for cid in subunits if len(cid.split(":")) == 3]
    for a, ci in enumerate(tagged):
        for cj in tagged[a + 1:]:
            if ci in subunits and cj in subunits:
                x0, x1 = subunits[ci]["xyz"], subunits[cj]["xyz"]
                if x0.shape[0] == x1.shape[0] and np.min(np.linalg.norm(x0 - x1, axis=1)) < 0.2:
                    subunits.pop(cj)
    return subunits


def filter_non_atomic_subunits(subunits):
    """src/structure.py:138-146: subunits with exactly one atom per residue (and more than one atom) are dropped."""
    for sname in list(subunits):
        n_atm = subunits[sname]["xyz"].shape[0]
        if n_atm > 1 and n_atm == np.unique(subunits[sname]["resid"]).shape[0]:
            subunits.pop(sname)
    return subunits


def encode_bfactor(structure, p):
    """src/structure.py:185-223: spread per-residue (or per-C-alpha, or per-atom) values p over the atoms as 'bfactor'."""
    resids = structure["resid"]
    ca = (structure["name"] == "CA") & (structure["element"] == "C") & (structure["het_flag"] == "A")
    p = np.asarray(p)
    if p.shape[0] == ca.shape[0]:
        structure["bfactor"] = p
    elif p.shape[0] == int(ca.sum()):
        bf = np.zeros(len(resids), dtype=np.float32)
        res_of_ca = resids[ca]
        for r in np.unique(resids):
            hit = np.where(res_of_ca == r)[0]
            if len(hit):
                bf[resids == r] = float(np.max(p[hit]))
        structure["bfactor"] = bf
    elif p.shape[0] == np.unique(resids).shape[0]:
        ures, inv = np.unique(resids, return_inverse=True)
        structure["bfactor"] = np.asarray(p, dtype=np.float32).reshape(len(ures), -1).max(axis=1)[inv].astype(np.float32)
    else:
        print("WARNING: bfactor not saved")
    return structure


def preprocess_structure(structure):
    """The preprocessing chain of StructuresDataset.__getitem__ (src/dataset.py:140-154): clean, tag hetero residues as
    their own chains, split by chain, drop one-atom-per-residue subunits and duplicated hetero subunits."""
    structure = tag_hetatm_chains(clean_structure(structure))
    return remove_duplicate_tagged_subunits(filter_non_atomic_subunits(split_by_chain(structure)))
