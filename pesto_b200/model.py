"""`Model` with the reference's constructor, state-dict keys and `forward(X, ids_topk, q0, M)` signature
(model/model.py:6-52 of LBM-EPFL/PeSTo), computing in the hand-written CUDA library.

The torch modules below only HOLD parameters so that `load_state_dict(torch.load('model_ckpt.pt'))` works with the
shipped checkpoints unchanged (key map: SURVEY.md A.5).  `forward` packs them once per device into the C-ABI model
handle and enqueues `pesto_forward` on the current CUDA stream.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib


def _mlp(n_in, n_hidden, n_out, bias=True):
    """Linear-ELU-Linear-ELU-Linear container: parameter names `<name>.{0,2,4}.{weight,bias}`."""
    return torch.nn.Sequential(
        torch.nn.Linear(n_in, n_hidden, bias=bias), torch.nn.ELU(),
        torch.nn.Linear(n_hidden, n_hidden, bias=bias), torch.nn.ELU(),
        torch.nn.Linear(n_hidden, n_out, bias=bias))


class StateUpdate(torch.nn.Module):
    """Parameter holder for src/model_operations.py:26-85."""

    def __init__(self, Ns, Nh, Nk):
        super().__init__()
        self.Ns, self.Nh, self.Nk = Ns, Nh, Nk
        self.nqm = _mlp(2 * Ns, Ns, 2 * Nk * Nh)
        self.eqkm = _mlp(6 * Ns + 1, Ns, Nk)
        self.epkm = _mlp(6 * Ns + 1, Ns, 3 * Nk)
        self.evm = _mlp(6 * Ns + 1, 2 * Ns, 2 * Ns)
        self.qpm = _mlp(Nh * Ns, Ns, Ns)
        self.ppm = torch.nn.Sequential(torch.nn.Linear(Nh * Ns, Ns, bias=False))
        self.sdk = torch.nn.Parameter(torch.tensor(math.sqrt(Nk), dtype=torch.float32), requires_grad=False)


class StateUpdateLayer(torch.nn.Module):
    """Parameter holder for src/model_operations.py:217-223."""

    def __init__(self, layer_params):
        super().__init__()
        self.su = StateUpdate(layer_params["Ns"], layer_params["Nh"], layer_params["Nk"])
        self.m_nn = torch.nn.Parameter(torch.arange(layer_params["nn"], dtype=torch.int64), requires_grad=False)


class StatePoolLayer(torch.nn.Module):
    """Parameter holder for src/model_operations.py:171-195."""

    def __init__(self, N0, N1, Nh):
        super().__init__()
        self.sam = _mlp(2 * N0, N0, 2 * Nh)
        self.zdm = torch.nn.Sequential(
            torch.nn.Linear(Nh * N0, N0), torch.nn.ELU(), torch.nn.Linear(N0, N0), torch.nn.ELU(), torch.nn.Linear(N0, N1))
        self.zdm_vec = torch.nn.Sequential(torch.nn.Linear(Nh * N0, N1, bias=False))


class Model(torch.nn.Module):
    """Drop-in for the reference `Model` (inference only): model/model.py:6-52 and the per-checkpoint copies under
    model/save/*/model.py.  Those copies differ in one respect only: i_v3_1 (model/save/i_v3_1_2021-05-28_12-40/model.py:9-22)
    has a single Linear as embedding and as decoder where the other checkpoints have Linear-ELU-Linear-ELU-Linear; pass
    `em_layers=1, dm_layers=1` for it (or build the model with `Model.for_state_dict`, which reads the depth off the keys).

    forward(X, ids_topk, q0, M) -> z[R, N2] logits on the input device (N2 = config['dm']['N2']: 5 for i_v4_*, 1 for i_v3_1).
      X         [N, 3] float32 coordinates
      ids_topk  [N, K<=64] int64, 1-based, 0 = sink   (what collate_batch_features returns)
      q0        [N, N0] float32 features
      M         [N, R] float membership (rows one-hot), as the reference; or, as an extension that avoids the dense
                matrix, a 1-D integer tensor with the residue column of every atom (pass n_res= to skip a sync).
    `mode`: 'f16x3' (default: tcgen05 tensor cores, 3-term split product over fp16 hi/lo planes, logits within ~1e-4 of
    the reference), 'fp32' (FFMA, exact mode) or 'f16' (tensor cores, single fp16 pass, speed mode with ~2e-2 logit error).
    'bf16x3' / 'bf16' are accepted as former names of the two tensor-core modes.
    """

    def __init__(self, config, mode="f16x3", em_layers=3, dm_layers=3):
        super().__init__()
        for lp in config["sum"]:
            if (lp["Ns"], lp["Nh"], lp["Nk"]) != (32, 2, 3) or lp["nn"] not in (8, 16, 32, 64):
                raise ValueError(f"unsupported layer parameters {lp}: kernels are built for Ns=32, Nh=2, Nk=3, nn in 8/16/32/64")
        if config["em"]["N1"] != 32 or config["spl"] != {"N0": 32, "N1": 32, "Nh": 4} or \
                (config["dm"]["N0"], config["dm"]["N1"]) != (32, 32) or not 1 <= config["dm"]["N2"] <= 8 or \
                not 1 <= config["em"]["N0"] <= 128 or em_layers not in (1, 3) or dm_layers not in (1, 3):
            raise ValueError("unsupported em/spl/dm configuration for the CUDA kernels")
        self.config = config
        self.mode = mode
        self.num_out = int(config["dm"]["N2"])
        if em_layers == 3:
            self.em = _mlp(config["em"]["N0"], config["em"]["N1"], config["em"]["N1"])
        else:
            self.em = torch.nn.Sequential(torch.nn.Linear(config["em"]["N0"], config["em"]["N1"]))
        self.sum = torch.nn.Sequential(*[StateUpdateLayer(lp) for lp in config["sum"]])
        self.spl = StatePoolLayer(config["spl"]["N0"], config["spl"]["N1"], config["spl"]["Nh"])
        if dm_layers == 3:
            self.dm = _mlp(2 * config["dm"]["N0"], config["dm"]["N1"], config["dm"]["N2"])
        else:
            self.dm = torch.nn.Sequential(torch.nn.Linear(2 * config["dm"]["N0"], config["dm"]["N2"]))
        self._handles = {}        # device index -> C model handle
        self._workspaces = {}     # device index -> uint8 tensor
        self._last = {}           # device index -> (aligned workspace pointer, n_atoms, n_res) of the last forward

    @classmethod
    def for_state_dict(cls, config, state_dict, mode="f16x3"):
        """Model with the head depths the checkpoint has (`em.2.weight` / `dm.2.weight` present -> three layers), loaded."""
        model = cls(config, mode=mode, em_layers=3 if "em.2.weight" in state_dict else 1,
                    dm_layers=3 if "dm.2.weight" in state_dict else 1)
        model.load_state_dict(state_dict)
        return model

    # ---- packed-weight handle management ------------------------------------------------------------------
    def _invalidate(self):
        lib = _lib.load() if self._handles else None
        for h in self._handles.values():
            lib.pesto_model_destroy(h)
        self._handles = {}

    def repack(self):
        """Call after modifying parameters in place (load_state_dict / .to() do it automatically)."""
        self._invalidate()

    def load_state_dict(self, *args, **kwargs):
        self._invalidate()
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._invalidate()
        return super()._apply(fn, *args, **kwargs)

    def __del__(self):
        try:
            self._invalidate()
        except Exception:
            pass

    def _handle(self, dev_index):
        h = self._handles.get(dev_index)
        if h is not None:
            return h
        lib = _lib.load()
        nns = np.asarray([lp["nn"] for lp in self.config["sum"]], dtype=np.int32)
        h = lib.pesto_model_create(len(nns), nns.ctypes.data, int(self.config["em"]["N0"]))
        if not h:
            _lib.check(-1, "pesto_model_create")
        try:
            for key, t in self.state_dict().items():
                if not t.dtype.is_floating_point:
                    continue
                a = np.ascontiguousarray(t.detach().to("cpu", torch.float32).numpy())
                _lib.check(lib.pesto_model_set_tensor(h, key.encode(), a.ctypes.data, a.size), f"set_tensor({key})")
            with torch.cuda.device(dev_index):
                _lib.check(lib.pesto_model_finalize(h), "pesto_model_finalize")
            if lib.pesto_model_num_out(h) != self.num_out:
                raise _lib.PestoError(f"the decoder packs {lib.pesto_model_num_out(h)} logits per residue, the configuration says {self.num_out}")
        except Exception:
            lib.pesto_model_destroy(h)
            raise
        self._handles[dev_index] = h
        return h

    def _workspace(self, dev, nbytes):
        ws = self._workspaces.get(dev.index)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=dev)
            self._workspaces[dev.index] = ws
        return ws

    # ---- forward --------------------------------------------------------------------------------------------
    def forward(self, X, ids_topk, q0, M, n_res=None, mode=None):
        if not torch.cuda.is_available():
            raise _lib.PestoError("pesto_b200.Model needs a CUDA device: there is no CPU implementation")
        out_device = X.device
        dev = X.device if X.is_cuda else torch.device("cuda", torch.cuda.current_device())
        lib = _lib.load()

        def prep(t, dtype):
            return t.detach().to(device=dev, dtype=dtype, non_blocking=True).contiguous()

        Xd = prep(X, torch.float32)
        ids = prep(ids_topk, torch.int64)
        qd = prep(q0, torch.float32)
        if Xd.dim() != 2 or Xd.shape[1] != 3 or ids.dim() != 2 or qd.dim() != 2 or \
                ids.shape[0] != Xd.shape[0] or qd.shape[0] != Xd.shape[0]:
            raise ValueError(f"bad input shapes X{tuple(Xd.shape)} ids_topk{tuple(ids.shape)} q0{tuple(qd.shape)}")
        if Xd.shape[0] == 0:
            raise ValueError("Model.forward: empty structure (0 atoms)")
        if qd.shape[1] != self.config["em"]["N0"]:
            raise ValueError(f"q0 has {qd.shape[1]} features, the model expects {self.config['em']['N0']}")
        n_atoms = Xd.shape[0]
        Md = rid = None
        if M.dim() == 2:
            if M.shape[0] != n_atoms:
                raise ValueError(f"M has {M.shape[0]} rows for {n_atoms} atoms")
            Md = prep(M, torch.float32)
            n_res = M.shape[1]
        else:
            rid = prep(M, torch.int32)
            if rid.shape[0] != n_atoms:
                raise ValueError(f"residue index has {rid.shape[0]} entries for {n_atoms} atoms")
            n_res = int(rid.max()) + 1 if n_res is None else int(n_res)
        with torch.cuda.device(dev):
            h = self._handle(dev.index)
            nbytes = lib.pesto_forward_workspace_bytes(n_atoms, n_res)
            ws = self._workspace(dev, nbytes)
            z = torch.empty((n_res, self.num_out), dtype=torch.float32, device=dev)
            base = ws.data_ptr()
            aligned = (base + 255) // 256 * 256
            stream = torch.cuda.current_stream(dev).cuda_stream
            rc = lib.pesto_forward(h, Xd.data_ptr(), ids.data_ptr(), ids.shape[1], qd.data_ptr(), _lib.ptr(Md),
                                   _lib.ptr(rid), n_atoms, n_res, z.data_ptr(), aligned, ws.numel() - (aligned - base),
                                   _lib.MODES[mode or self.mode], ctypes.c_void_p(stream))
            _lib.check(rc, "pesto_forward")
            self._last[dev.index] = (aligned, n_atoms, n_res, ws, aligned - base)
        if out_device != dev:
            z = z.to(out_device)
            self.raise_if_failed(dev)          # the copy synchronised anyway
        return z

    def status_words(self, device=None):
        """int32[8] device view of the status words of the last forward on `device` (stream-ordered: valid until the next
        forward on that device starts).  Callers that keep several forwards in flight copy it out together with the logits
        and hand the host copy to `_lib.raise_status`."""
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        _, n_atoms, n_res, ws, shift = self._last[idx]
        off = shift + _lib.load().pesto_forward_status_offset(n_atoms, n_res)
        return ws[off:off + 32].view(torch.int32)

    def raise_if_failed(self, device=None):
        """Errors only the device can detect (a neighbour id or residue index out of range, a row of M that is not one-hot,
        a tensor-core stage that never completed) make every logit of that forward NaN without a host synchronisation.
        This reads the status words of the last forward on `device` (synchronising its current stream) and raises
        PestoError if one is set.  Callers that synchronise anyway -- the structure runner, the trajectory and apply paths --
        call it there, as the reference's callers catch exceptions per structure (interfaceome/apply_model.py:81-82)."""
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        last = self._last.get(idx)
        if last is None:
            return
        status = (ctypes.c_int32 * 8)()
        with torch.cuda.device(idx):
            stream = torch.cuda.current_stream(idx).cuda_stream
            rc = _lib.load().pesto_forward_status(ctypes.c_void_p(last[0]), last[1], last[2], status, ctypes.c_void_p(stream))
        _lib.check(rc, "pesto_forward")

    def launches_per_forward(self, dense_m=True):
        return len(self.config["sum"]) * 2 + 3 + 5 + (1 if dense_m else 0) + (0 if self.mode == "fp32" else 1)
