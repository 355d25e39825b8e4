"""Many structures -> logits, the caller pattern of interfaceome/apply_model.py:49-82 (one structure at a time in the
reference) re-laid for the GPU: structures are packed back to back into batches of ~128 k atoms (one kNN launch
sequence + one forward per batch, residue membership as an index instead of the dense block-diagonal M of
src/dataset.py:101-110), and -- across GPUs -- sharded by cost with no data-path collective.

    shards = rank_shard([len(s["xyz"]) for s in structures], rank, world)          # LPT by atoms
    for idx, z in predict_structures(model, [structures[i] for i in shards]): ...   # z[n_res, 5] logits per structure
"""
import numpy as np
import torch

from . import _lib
from .data_encoding import batch_topology, onehot, onehot_index, std_elements, std_names, std_resnames

TARGET_ATOMS_PER_BATCH = 131072


def pack_batches(sizes, target=TARGET_ATOMS_PER_BATCH, num_nn=64):
    """Greedy packing in the given order: lists of structure indices whose atom counts add up to <= target
    (a structure larger than the target gets a batch of its own).  So does a structure with fewer than num_nn atoms:
    its neighbour lists are padded with sink slots, and a sink slot reads X[-1] -- the last atom of whatever the batch
    holds (src/model_operations.py:8) -- so only alone does it get the numbers of the reference's one-structure-at-a-time
    loop (interfaceome/apply_model.py:49-82)."""
    batches, cur, load = [], [], 0
    for i, n in enumerate(sizes):
        if int(n) < num_nn:
            if cur:
                batches.append(cur)
                cur, load = [], 0
            batches.append([i])
            continue
        if cur and load + int(n) > target:
            batches.append(cur)
            cur, load = [], 0
        cur.append(i)
        load += int(n)
    if cur:
        # a small tail (< target / 4) rides along with the previous batch: a forward over a few thousand atoms is launch-bound
        prev_ok = batches and all(int(sizes[j]) >= num_nn for j in batches[-1])
        if prev_ok and load < target // 4 and all(int(sizes[j]) >= num_nn for j in cur):
            batches[-1].extend(cur)
        else:
            batches.append(cur)
    return batches


# feature blocks of q0 (src/data_encoding.py:78-84): the v4 models take the element one-hot (30 columns), the v3 models the
# concatenation element | residue name | atom name (30 + 29 + 64 = 123 columns, model/save/i_v3_*/src/data_encoding.py:105-108)
FEATURE_BLOCKS = {30: (("element", std_elements),),
                  123: (("element", std_elements), ("resname", std_resnames), ("name", std_names))}


def element_index(elements):
    """Column of every atom in the element one-hot of src/data_encoding.py:56-58,78-84: position in std_elements, or
    len(std_elements) (= "unknown", the last column) -- as uint8."""
    return onehot_index(elements, std_elements).astype(np.uint8)


def feature_index(structure, n_features=30):
    """uint8 [N, number of one-hot blocks]: the hot column of each block of q0 (one byte per block crosses PCIe instead of
    n_features floats; `expand_features` rebuilds q0 on the device)."""
    if n_features not in FEATURE_BLOCKS:
        raise ValueError(f"no feature encoding with {n_features} columns (known: {sorted(FEATURE_BLOCKS)})")
    return np.stack([onehot_index(structure[key], vocab) for key, vocab in FEATURE_BLOCKS[n_features]], axis=1).astype(np.uint8)


def expand_features(index, n_features=30):
    """q0 [N, n_features] float32 on the device of `index` (uint8 [N, blocks] from feature_index)."""
    blocks = FEATURE_BLOCKS[n_features]
    if index.dim() == 1:
        index = index.unsqueeze(1)
    hot = [torch.nn.functional.one_hot(index[:, b].long(), len(vocab) + 1) for b, (_, vocab) in enumerate(blocks)]
    return (hot[0] if len(hot) == 1 else torch.cat(hot, dim=1)).to(torch.float32)


def encode_batch(structures, as_index=False, n_features=30):
    """Host side of one batch: X [N,3] f32, q0 [N,n_features] f32 (one-hot blocks, src/data_encoding.py:78-84; with
    as_index the uint8 hot columns instead, expanded on the device), residue index [N] int32 (position of the atom's
    resid among the structure's sorted unique resids, src/data_encoding.py:73, plus the batch's residue offset),
    per-structure atom and residue counts."""
    X = np.concatenate([np.asarray(s["xyz"], dtype=np.float32) for s in structures], axis=0)
    if as_index:
        q0 = np.concatenate([feature_index(s, n_features) for s in structures], axis=0)
        if q0.shape[1] == 1:
            q0 = q0[:, 0]                                       # one block: a flat index vector
    else:
        q0 = np.concatenate([np.concatenate([onehot(s[key], vocab) for key, vocab in FEATURE_BLOCKS[n_features]], axis=1)
                             for s in structures], axis=0).astype(np.float32)
    rids, n_res, r0 = [], [], 0
    for s in structures:
        u, inv = np.unique(np.asarray(s["resid"]), return_inverse=True)
        rids.append(inv.astype(np.int64) + r0)
        n_res.append(len(u))
        r0 += len(u)
    return X, q0, np.concatenate(rids).astype(np.int32), [len(s["xyz"]) for s in structures], n_res


class _PinnedSlot:
    """Reusable pinned staging buffers of one pipeline slot (allocating pinned memory per batch costs milliseconds)."""

    def __init__(self):
        self.buf = {}
        self.copied = None          # event: the slot's last H2D copies have completed

    def reserve(self, name, numel, dtype):
        """Size a buffer once for the largest batch of the job (growing it batch by batch costs a pinned allocation each time)."""
        t = self.buf.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            self.buf[name] = torch.empty(max(int(numel), 1), dtype=dtype).pin_memory()

    def put(self, name, arr):
        t = self.buf.get(name)
        if t is None or t.numel() < arr.size or t.dtype != torch.from_numpy(arr[:0]).dtype:
            t = torch.empty(max(arr.size, 1) * 5 // 4, dtype=torch.from_numpy(arr[:0]).dtype).pin_memory()
            self.buf[name] = t
        v = t[:arr.size].view(arr.shape)
        v.numpy()[...] = arr
        return v


class _Staging:
    """Pinned staging buffers and the copy stream of one device, kept for the life of the process: pinned allocations cost
    milliseconds each, a job of a few dozen structures only tens of milliseconds of GPU time.  (One job per device at a
    time: predict_structures is a generator that owns these buffers until it is exhausted.)"""
    _per_device = {}

    def __init__(self, dev):
        self.copy_stream = torch.cuda.Stream(dev)
        self.slots = [_PinnedSlot(), _PinnedSlot()]
        self.zpin = [None, None]
        self.spin = [torch.zeros(8, dtype=torch.int32).pin_memory() for _ in range(2)]

    @classmethod
    def get(cls, dev):
        key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
        if key not in cls._per_device:
            cls._per_device[key] = cls(dev)
        return cls._per_device[key]

    def reserve(self, n_atoms_max, n_blocks=1, n_out=5):
        for slot in self.slots:
            slot.reserve("X", 3 * n_atoms_max, torch.float32)
            slot.reserve("el", n_blocks * n_atoms_max, torch.uint8)
            slot.reserve("rid", n_atoms_max, torch.int32)
        for k in range(2):
            if self.zpin[k] is None or self.zpin[k].shape[0] < n_atoms_max // 4 or self.zpin[k].shape[1] != n_out:
                self.zpin[k] = torch.empty((max(n_atoms_max // 4, 1024), n_out), dtype=torch.float32).pin_memory()


def predict_structures(model, structures, device="cuda", target_atoms=TARGET_ATOMS_PER_BATCH, num_nn=64):
    """Yield (index, z[n_res, N2] float32 on the host) for every structure dictionary (keys xyz, element, resid -- and
    resname, name for the 123-feature v3 models), in order.

    Pipeline per batch k: (1) its kNN + forward + D2H of logits and status words are enqueued, (2) the host encodes batch
    k+1 and enqueues its pinned H2D copies on a copy stream while the GPU works, (3) the host collects batch k-1's results
    -- two batches are in flight, so the GPU never waits for the host between batches.  Errors only the device can see
    (a neighbour id or residue index out of range, a hung tensor-core stage) raise PestoError for the batch they occur in."""
    dev = torch.device(device)
    sizes = [len(s["xyz"]) for s in structures]
    batches = pack_batches(sizes, target_atoms, num_nn)
    n_features = int(model.config["em"]["N0"])
    n_out = int(getattr(model, "num_out", 5))
    staging = _Staging.get(dev)
    staging.reserve(max((sum(sizes[i] for i in b) for b in batches), default=0), len(FEATURE_BLOCKS[n_features]), n_out)
    copy_stream, slots = staging.copy_stream, staging.slots
    for slot in slots:
        slot.copied = None

    def stage(k):
        slot = slots[k % 2]
        if slot.copied is not None:
            slot.copied.synchronize()                       # the buffers' previous copies are long done
        X, el, rid, n_at, n_rs = encode_batch([structures[i] for i in batches[k]], as_index=True, n_features=n_features)
        host = [slot.put(n, a) for n, a in (("X", X), ("el", el), ("rid", rid))]
        with torch.cuda.stream(copy_stream):
            on_dev = [t.to(dev, non_blocking=True) for t in host]
            # one byte per atom and one-hot block crosses PCIe; the [N, N0] one-hots Model.forward takes are expanded on the device
            on_dev[1] = expand_features(on_dev[1], n_features)
            slot.copied = torch.cuda.Event()
            slot.copied.record(copy_stream)
        return on_dev, slot.copied, n_at, n_rs

    nxt = stage(0) if batches else None
    zpin, spin = staging.zpin, staging.spin                     # (zpin is grown below if a batch has more residues than atoms / 4)

    def finish(job):            # wait for a batch's logits and status words, hand out per-structure results
        k, b, n_rs, zh, done = job
        done.synchronize()
        _lib.raise_status(spin[k % 2].numpy(), "predict_structures")     # bad ids / hung tensor-core stage of THIS batch -> PestoError
        zc = zh.clone()
        r0 = 0
        for i, nr in zip(b, n_rs):
            yield i, zc[r0:r0 + nr]
            r0 += nr

    pending = None              # two batches in flight: batch k's results are collected while batch k + 1 runs
    with torch.no_grad():
        for k, b in enumerate(batches):
            (Xd, q0d, ridd), ready, n_at, n_rs = nxt
            main = torch.cuda.current_stream(dev)
            main.wait_event(ready)
            for t in (Xd, q0d, ridd):
                t.record_stream(main)
            ids1 = batch_topology(Xd, n_at, num_nn)
            z = model(Xd, ids1, q0d, ridd, n_res=int(sum(n_rs)))
            if zpin[k % 2].shape[0] < z.shape[0]:
                zpin[k % 2] = torch.empty((z.shape[0] * 5 // 4, n_out), dtype=torch.float32).pin_memory()
            zh = zpin[k % 2][:z.shape[0]]
            zh.copy_(z, non_blocking=True)
            spin[k % 2].copy_(model.status_words(dev), non_blocking=True)
            done = torch.cuda.Event()
            done.record(main)
            nxt = stage(k + 1) if k + 1 < len(batches) else None      # host work of the next batch under this batch's GPU work
            if pending is not None:
                yield from finish(pending)
            pending = (k, b, n_rs, zh, done)
        if pending is not None:
            yield from finish(pending)
