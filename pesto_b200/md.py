"""Trajectory mode: many frames of one molecule with a fixed topology (md_analysis/apply_model_md.ipynb cell 6 of
LBM-EPFL/PeSTo: `extract_topology(X_traj[:,0], 64)` once, then `model(X_traj[:,i], ids_topk, q, M)` frame by frame).

Frames are independent, so they are batched like structures (src/dataset.py:91-112): F frames become one
F*N-atom forward whose neighbour indices are the frame-0 indices shifted by the frame's atom offset.  One launch
sequence then carries ~128 k atoms instead of a few thousand, which takes the per-frame forward out of the
launch-bound regime (profiles/README.md: 1.1 M atoms/s at 2 386 atoms vs 3.2 M atoms/s on a full batch).
"""
import torch

TARGET_ATOMS_PER_BATCH = 131072


def _residue_index(M):
    return M if M.dim() == 1 else M.to(torch.float32).argmax(dim=1)


def predict_trajectory(model, X_traj, ids_topk, q, M, frames=None, frames_per_batch=None, device="cuda"):
    """Logits z[len(frames), R, N2] (N2 = 5 for the i_v4 models) for the frames of X_traj [N, T, 3] (atoms, frames, xyz -- the notebook's layout).

    ids_topk [N, 64] int64 is what collate_batch_features returns (1-based, 0 = sink) and is reused for every
    frame; q [N, 30]; M dense [N, R] or a residue index [N].  `frames`: iterable of frame numbers (default: all).
    """
    n_atoms, n_frames = X_traj.shape[0], X_traj.shape[1]
    frames = list(range(n_frames)) if frames is None else list(frames)
    per = frames_per_batch or max(1, TARGET_ATOMS_PER_BATCH // max(n_atoms, 1))
    per = min(per, max(len(frames), 1))
    if n_atoms < ids_topk.shape[1] or bool((ids_topk == 0).any()):
        per = 1        # sink-padded neighbour slots read X[-1] of the batch (src/model_operations.py:8): one frame per forward
    dev = torch.device(device)
    ids = ids_topk.to(dev)
    rid = _residue_index(M.to(dev)).to(torch.int64)
    n_res = int(M.shape[1]) if M.dim() == 2 else int(rid.max()) + 1
    qd = q.to(dev, torch.float32)
    shift = (torch.arange(per, device=dev) * n_atoms).view(per, 1, 1)
    ids_b = torch.where(ids.unsqueeze(0) > 0, ids.unsqueeze(0) + shift, torch.zeros_like(ids).unsqueeze(0)).reshape(per * n_atoms, -1)
    rid_b = (rid.unsqueeze(0) + (torch.arange(per, device=dev) * n_res).view(per, 1)).reshape(-1).to(torch.int32)
    q_b = qd.repeat(per, 1)
    n_out = getattr(model, "num_out", 5)
    out = torch.empty((len(frames), n_res, n_out), dtype=torch.float32, device=dev)
    Xd = X_traj if X_traj.is_cuda else None
    with torch.no_grad():
        for b0 in range(0, len(frames), per):
            sel = frames[b0:b0 + per]
            f = len(sel)
            idx = torch.as_tensor(sel, device=X_traj.device)
            Xb = (Xd if Xd is not None else X_traj).index_select(1, idx).permute(1, 0, 2).reshape(f * n_atoms, 3)
            Xb = Xb.to(dev, torch.float32, non_blocking=True).contiguous()
            z = model(Xb, ids_b[:f * n_atoms], q_b[:f * n_atoms], rid_b[:f * n_atoms], n_res=f * n_res)
            out[b0:b0 + f] = z.view(f, n_res, n_out)
            if b0 == 0 or b0 + per >= len(frames):
                model.raise_if_failed(dev)     # input flags show on the first batch (topology and membership are shared by all frames)
    return out
