"""Many structures -> logits, the caller pattern of interfaceome/apply_model.py:49-82 (one structure at a time in the
reference) re-laid for the GPU: structures are packed back to back into batches of ~128 k atoms (one kNN launch
sequence + one forward per batch, residue membership as an index instead of the dense block-diagonal M of
src/dataset.py:101-110), and -- across GPUs -- sharded by cost with no data-path collective.

    shards = rank_shard([len(s["xyz"]) for s in structures], rank, world)          # LPT by atoms
    for idx, z in predict_structures(model, [structures[i] for i in shards]): ...   # z[n_res, 5] logits per structure
"""
import numpy as np
import torch

from .data_encoding import batch_topology, onehot, std_elements

TARGET_ATOMS_PER_BATCH = 131072


def pack_batches(sizes, target=TARGET_ATOMS_PER_BATCH):
    """Greedy packing in the given order: lists of structure indices whose atom counts add up to <= target
    (a structure larger than the target gets a batch of its own)."""
    batches, cur, load = [], [], 0
    for i, n in enumerate(sizes):
        if cur and load + int(n) > target:
            batches.append(cur)
            cur, load = [], 0
        cur.append(i)
        load += int(n)
    if cur:
        batches.append(cur)
    return batches


def encode_batch(structures):
    """Host side of one batch: X [N,3] f32, q0 [N,30] f32 (element one-hot, src/data_encoding.py:78-84), residue index
    [N] int32 (position of the atom's resid among the structure's sorted unique resids, src/data_encoding.py:73, plus
    the batch's residue offset), per-structure atom and residue counts."""
    X = np.concatenate([np.asarray(s["xyz"], dtype=np.float32) for s in structures], axis=0)
    q0 = np.concatenate([onehot(s["element"], std_elements) for s in structures], axis=0).astype(np.float32)
    rids, n_res, r0 = [], [], 0
    for s in structures:
        u, inv = np.unique(np.asarray(s["resid"]), return_inverse=True)
        rids.append(inv.astype(np.int64) + r0)
        n_res.append(len(u))
        r0 += len(u)
    return X, q0, np.concatenate(rids).astype(np.int32), [len(s["xyz"]) for s in structures], n_res


def predict_structures(model, structures, device="cuda", target_atoms=TARGET_ATOMS_PER_BATCH, num_nn=64):
    """Yield (index, z[n_res, 5] float32 on the host) for every structure dictionary (keys xyz, element, resid), in order.
    The next batch's inputs are staged (pinned, non-blocking) on a copy stream while the current batch computes."""
    dev = torch.device(device)
    sizes = [len(s["xyz"]) for s in structures]
    batches = pack_batches(sizes, target_atoms)
    copy_stream = torch.cuda.Stream(dev)

    def stage(b):
        X, q0, rid, n_at, n_rs = encode_batch([structures[i] for i in b])
        host = [torch.from_numpy(a).pin_memory() for a in (X, q0, rid)]
        with torch.cuda.stream(copy_stream):
            on_dev = [t.to(dev, non_blocking=True) for t in host]
            ready = torch.cuda.Event()
            ready.record(copy_stream)
        return on_dev, host, ready, n_at, n_rs

    nxt = stage(batches[0]) if batches else None
    with torch.no_grad():
        for k, b in enumerate(batches):
            (Xd, q0d, ridd), _host, ready, n_at, n_rs = nxt
            torch.cuda.current_stream(dev).wait_event(ready)
            for t in (Xd, q0d, ridd):
                t.record_stream(torch.cuda.current_stream(dev))
            nxt = stage(batches[k + 1]) if k + 1 < len(batches) else None
            ids1 = batch_topology(Xd, n_at, num_nn)
            z = model(Xd, ids1, q0d, ridd, n_res=int(sum(n_rs))).cpu()
            r0 = 0
            for i, nr in zip(b, n_rs):
                yield i, z[r0:r0 + nr]
                r0 += nr
