"""pesto_b200 -- B200-native (sm_100a) forward path of the PeSTo geometric transformer.

Only the hot path lives here: exact kNN topology, the StateUpdate
layer stack, residue pooling + decoding, behind the reference's Python call boundary.
All arithmetic runs in hand-written CUDA reached through the C ABI in
`include/pesto_b200.h` (`pesto_b200/csrc`); there is no CPU fallback.
"""
from .synth import synth_structure, one_hot_features, dense_membership  # noqa: F401

__all__ = ["Model", "extract_topology", "collate_batch_features", "lib"]


def __getattr__(name):
    # lazy: importing the package must not need the CUDA library (CPU-only tooling imports synth)
    if name == "Model":
        from .model import Model
        return Model
    if name in ("extract_topology", "encode_structure", "encode_features"):
        from . import data_encoding
        return getattr(data_encoding, name)
    if name == "collate_batch_features":
        from .dataset import collate_batch_features
        return collate_batch_features
    if name == "lib":
        from . import _lib
        return _lib
    raise AttributeError(name)
