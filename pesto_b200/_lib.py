"""ctypes binding of libpesto_b200.so (the C ABI declared in include/pesto_b200.h).

There is no fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PESTO_B200_LIB: experiment aid (profiles/variants.py builds kernel variants side by side); the default is the in-tree build
LIB_PATH = os.environ.get("PESTO_B200_LIB") or os.path.join(_HERE, "libpesto_b200.so")

MODE_FP32, MODE_F16X3, MODE_F16 = 0, 1, 2
MODE_BF16X3, MODE_BF16 = MODE_F16X3, MODE_F16          # former names (see include/pesto_b200.h)
MODES = {"fp32": MODE_FP32, "f16x3": MODE_F16X3, "f16": MODE_F16, "bf16x3": MODE_F16X3, "bf16": MODE_F16}

_c = ctypes
_vp, _i, _sz, _i64 = _c.c_void_p, _c.c_int, _c.c_size_t, _c.c_int64

# name -> (restype, argtypes); mirrors include/pesto_b200.h one to one
SIGNATURES = {
    "pesto_abi_version": (_i, []),
    "pesto_last_error": (_c.c_char_p, []),
    "pesto_knn_scratch_bytes": (_sz, [_i, _i]),
    "pesto_knn": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "pesto_knn_launch_count": (_i, []),
    "pesto_model_create": (_vp, [_i, _vp, _i]),
    "pesto_model_set_tensor": (_i, [_vp, _c.c_char_p, _vp, _i64]),
    "pesto_model_finalize": (_i, [_vp]),
    "pesto_model_destroy": (None, [_vp]),
    "pesto_model_num_layers": (_i, [_vp]),
    "pesto_model_num_out": (_i, [_vp]),
    "pesto_model_layer_nn": (_i, [_vp, _i]),
    "pesto_prologue": (_i, [_vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "pesto_node_scratch_bytes": (_sz, [_i]),
    "pesto_state_update": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "pesto_state_update_timed": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "pesto_edge_kernel_timed": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "pesto_residue_index": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "pesto_pool_scratch_bytes": (_sz, [_i, _i]),
    "pesto_pool_decode": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "pesto_unpack_state": (_i, [_vp, _i, _vp, _vp, _vp]),
    "pesto_forward_workspace_bytes": (_sz, [_i, _i]),
    "pesto_forward": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _sz, _i, _vp]),
    "pesto_forward_launch_count": (_i, [_vp, _i, _i]),
    "pesto_forward_status": (_i, [_vp, _i, _i, _vp, _vp]),
    "pesto_forward_status_offset": (_sz, [_i, _i]),
    "pesto_debug_force_watchdog": (_i, [_i]),
    "pesto_debug_watchdog": (_i, [_vp]),
    "pesto_pdb_count_atoms_host": (_i, [_c.c_char_p, _sz]),
    "pesto_pdb_parse_host": (_i, [_c.c_char_p, _sz, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "pesto_debug_umma_probe": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "pesto_debug_edge_timeline": (_i, [_vp, _i]),
    "pesto_debug_mma_time": (_i, [_i] * 13 + [_vp, _vp]),
    "pesto_debug_rmma_probe": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
}

_lib = None


class PestoError(RuntimeError):
    pass


def load():
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PestoError(
            f"{LIB_PATH} not found: the CUDA library has not been built. Run `python -m pesto_b200.build` "
            "(needs nvcc); pesto_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


STATUS_MESSAGES = {1: "a neighbour id in ids_topk is out of range [0, n_atoms]", 2: "a row of the membership matrix M is not one-hot",
                   4: "a residue index is out of range [0, n_res)"}


def raise_status(words, what="pesto_forward"):
    """Raise PestoError if the status words of a forward (PESTO_STATUS_WORDS int32, word 0 is scratch) flag an error."""
    for k, msg in STATUS_MESSAGES.items():
        if int(words[k]):
            raise PestoError(f"{what}: {msg} (the logits are NaN)")
    if int(words[3]) >= 100:
        raise PestoError(f"{what}: the state left the range of the fp16 operand planes (|q| or |p| > 2^14, or NaN): run this model / "
                         "input in mode fp32 (the logits are NaN)")
    if int(words[3]):
        raise PestoError(f"{what}: tensor-core stage {int(words[3])} never completed (watchdog): the logits are NaN")


def check(rc, what):
    if rc != 0:
        msg = load().pesto_last_error()
        raise PestoError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device/host pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
