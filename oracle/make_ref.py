#!/usr/bin/env python
"""Stage the UNMODIFIED reference implementation of the hot path under oracle/_ref/ -- TEST / BENCH INFRASTRUCTURE.

The reference (LBM-EPFL/PeSTo) is pure Python: there is nothing to compile.  The files that implement the path
 are copied verbatim, from where they lie under /root/reference, into oracle/_ref/, which is
git-ignored (the history never holds reference sources) but travels to the GPU box with the repo snapshot, so that
`bench.py --impl reference` can time the reference's own forward on the box's host cores (`cpu_baseline.kind` =
"reference").  Run by `__graft_entry__.build()` when /root/reference is present; nothing in the product imports it.

    python oracle/make_ref.py
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PESTO_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = {
    "src/model_operations.py": "src/model_operations.py",          # StateUpdate, StateUpdateLayer, StatePoolLayer, unpack_state_features
    "src/data_encoding.py": "src/data_encoding.py",                # extract_topology (+ the tables config.py imports)
    "model/save/i_v4_1_2021-09-07_11-21/model.py": "i_v4_1/model.py",
    "model/save/i_v4_1_2021-09-07_11-21/config.py": "i_v4_1/config.py",
}


def make_ref(verbose=True):
    if not os.path.isdir(REF):
        if verbose:
            print(f"oracle/_ref: {REF} not present, nothing staged (the prebuilt copy, if any, is kept)")
        return False
    for src, dst in FILES.items():
        out = os.path.join(DST, dst)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(REF, src), out)
    if verbose:
        print(f"oracle/_ref: staged {len(FILES)} reference files from {REF}")
    return True


def load_reference_model(weights):
    """The reference's `Model(config_model)` with the shipped i_v4_1 checkpoint (given as the golden fixture's arrays),
    and its `extract_topology`; raises FileNotFoundError if oracle/_ref has not been staged."""
    import importlib.util
    import torch
    if not os.path.exists(os.path.join(DST, "src", "model_operations.py")):
        raise FileNotFoundError("oracle/_ref is not staged (run python oracle/make_ref.py where /root/reference exists)")
    if DST not in sys.path:
        sys.path.insert(0, DST)                                    # `from src.model_operations import ...` inside model.py / config.py
    mods = {}
    for name in ("config", "model"):
        spec = importlib.util.spec_from_file_location(f"pesto_ref_{name}", os.path.join(DST, "i_v4_1", f"{name}.py"))
        mods[name] = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mods[name])
    model = mods["model"].Model(mods["config"].config_model)
    model.load_state_dict({k: torch.as_tensor(v) for k, v in weights.items()})
    from src.data_encoding import extract_topology                 # (oracle/_ref/src)
    return model.eval(), extract_topology


if __name__ == "__main__":
    make_ref()
