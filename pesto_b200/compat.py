"""Import shim: makes the reference's own import lines resolve to pesto_b200, so that `apply_model.ipynb`,
`profiling.py` and `interfaceome/apply_model.py` run unchanged after one call:

    import pesto_b200.compat; pesto_b200.compat.install()
    from src.dataset import StructuresDataset, collate_batch_features          # apply_model.ipynb:21
    from src.data_encoding import encode_structure, encode_features, extract_topology
    from src.structure import encode_bfactor, concatenate_chains, split_by_chain
    from src.structure_io import save_pdb, read_pdb
    sys.path.insert(0, save_path); from config import config_model; from model import Model   # apply_model.ipynb:66-73

`config` stays the reference's own file (a plain dictionary, found through sys.path as before); `model` and `src.*`
come from here.  Names of the reference that belong to training / dataset building (out of scope)
import fine and raise NotImplementedError when called.
"""
import sys
import types


def _out_of_scope(name):
    def fn(*_a, **_k):
        raise NotImplementedError(f"{name}: training / dataset tooling of the reference is out of scope of pesto_b200")
    fn.__name__ = name
    return fn


def install():
    from . import data_encoding, dataset, model, structure, structure_io
    src = types.ModuleType("src")
    src.__path__ = []                                     # a package, so that `import src.dataset` resolves
    mods = {"src": src, "src.dataset": dataset, "src.data_encoding": data_encoding, "src.structure": structure,
            "src.structure_io": structure_io, "model": model}
    for name in ("select_by_sid", "select_by_max_ba", "select_by_interface_types", "select_complete_assemblies"):
        if not hasattr(dataset, name):
            setattr(dataset, name, _out_of_scope(name))
    if not hasattr(structure, "data_to_structure"):
        structure.data_to_structure = _out_of_scope("data_to_structure")
    scoring = types.ModuleType("src.scoring")             # src/scoring.py: evaluation metrics (imported by the notebooks' first cell)
    scoring.bc_score_names = ["acc", "ppv", "npv", "tpr", "tnr", "mcc", "auc", "std"]
    for name in ("bc_scoring", "nanmean", "acc", "ppv", "npv", "tpr", "tnr", "mcc", "roc_auc"):
        setattr(scoring, name, _out_of_scope("src.scoring." + name))
    mods["src.scoring"] = scoring
    handler = types.ModuleType("data_handler")            # model/save/*/data_handler.py: HDF5 training dataset
    handler.Dataset = _out_of_scope("data_handler.Dataset")
    mods["data_handler"] = handler
    for k, m in mods.items():
        sys.modules[k] = m
        if k.startswith("src."):
            setattr(src, k.split(".", 1)[1], m)
    return mods
