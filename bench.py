#!/usr/bin/env python
"""bench.py -- atoms/sec of the i_v4_1 forward (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode fp32|f16x3|f16] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU, no data-path collective)

One "step" = one forward pass (em -> 32 StateUpdate layers -> residue pool -> decoder) over the workload:
BASELINE configs[1], the 53 `pdbs_test` structures (132 417 atoms, 16 632 residues) collated into one batch
(src/dataset.py:91-112 semantics, sparse membership).  Every rank processes the full workload (weak scaling).

  value     forward only, inputs (X, ids_topk, q0, residue index) resident in HBM, CUDA events, max over ranks
  e2e       per step: pinned host X / element index / residue index -> device, kNN topology on the device, forward, logits -> host
  roofline  the fused per-edge StateUpdate kernel of the nn=64 layers: algorithmic bytes N*(64*536+1024) / measured
            kernel time (CUDA events on the launching stream) vs the measured HBM peak of MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (torch, all host threads) on a bounded sample (one structure of the workload)

`--impl reference` times the CPU implementation (oracle port of the reference's forward) on the same sample.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")
TAG = "i_v4_1"
METRIC = "atoms/sec i_v4_1 forward"


def load_workload():
    g = dict(np.load(os.path.join(GOLDEN, "pdbs_test_53.npz")))
    sizes = g["sizes"].astype(np.int64)
    n_res = g["n_res"].astype(np.int64)
    roff = np.concatenate([[0], np.cumsum(n_res)])[:-1]
    rid = g["rid"].astype(np.int64) + np.repeat(roff, sizes)
    return dict(X=torch.from_numpy(g["X"]), el=torch.from_numpy(g["el"].astype(np.int64)),
                rid=torch.from_numpy(rid.astype(np.int32)), sizes=sizes, n_res=n_res)


def load_weights():
    return dict(np.load(os.path.join(GOLDEN, f"weights_{TAG}.npz")))


CPU_SAMPLE_STRUCTURES = 16      # bounded CPU sample: the first 16 of the 53 structures, one forward each (~10 s)


def cpu_sample(wl):
    """Bounded CPU sample of the same workload: the first structures, forwarded one by one like the reference does."""
    out, a0, r0 = [], 0, 0
    for i in range(CPU_SAMPLE_STRUCTURES):
        a1, r1 = a0 + int(wl["sizes"][i]), r0 + int(wl["n_res"][i])
        out.append(dict(index=i, X=wl["X"][a0:a1].contiguous(), el=wl["el"][a0:a1], rid=(wl["rid"][a0:a1].long() - r0),
                        n_res=r1 - r0, n_atoms=a1 - a0, r0=r0))
        a0, r0 = a1, r1
    return out


def cpu_topology(sample):
    from oracle import pesto_oracle as O
    t0 = time.perf_counter()
    ids = [O.extract_topology(s["X"], 64)[0] + 1 for s in sample]
    return ids, time.perf_counter() - t0


def cpu_forward_seconds(weights, sample, ids1):
    from oracle import pesto_oracle as O
    from pesto_b200.synth import one_hot_features
    t0 = time.perf_counter()
    zs = [O.forward(weights, s["X"], i1, one_hot_features(s["el"]), s["rid"], s["n_res"]) for s, i1 in zip(sample, ids1)]
    return time.perf_counter() - t0, zs


def sample_desc(sample):
    return (f"first {len(sample)} of the 53 structures ({sum(s['n_atoms'] for s in sample)} atoms, "
            f"{sum(s['n_res'] for s in sample)} residues), one forward per structure")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t_begin <= t <= t_end + 0.2 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        sm, reasons, smax, pw = [], set(), None, []
        for r in rows:
            try:
                sm.append(float(r[0])); smax = float(r[1]); pw.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


def structure_dicts(wl):
    """The workload as the structure dictionaries the reference's loaders produce (xyz, element, resid) -- what the
    many-structures runner (pesto_b200.runner, the caller pattern of interfaceome/apply_model.py:49-82) consumes."""
    from pesto_b200.data_encoding import std_elements
    names = np.append(std_elements, "X")
    out, a0, r0 = [], 0, 0
    for n, nr in zip(wl["sizes"], wl["n_res"]):
        a1 = a0 + int(n)
        out.append({"xyz": wl["X"][a0:a1].numpy(), "element": names[wl["el"][a0:a1].numpy()],
                    "resid": (wl["rid"][a0:a1].numpy() - r0) * 2 + 3})
        a0, r0 = a1, r0 + int(nr)
    return out


def reference_arm(args, rank):
    """The reference's own CPU implementation of the path, rank 0 only: the UNMODIFIED reference staged under oracle/_ref
    (kind "reference") when present, else the oracle port (kind "port").  Each step forwards a bounded sample of the
    workload, one structure at a time like the reference's callers; the sample is sized by a probe so that
    warm-up + timed steps stay within ~150 s (the unmodified reference runs at ~30-100 atoms/s: its norm over the
    strided [N, n, 3, S] dimension dominates)."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count())
    from pesto_b200.synth import one_hot_features, dense_membership
    wl = load_workload()
    weights = load_weights()
    sample = cpu_sample(wl)
    kind, note = "port", ("torch-CPU oracle port of the reference forward (oracle/_ref not staged on this box); it avoids the "
                          "reference's slow strided torch.norm, so it is faster than the unmodified reference")
    fwd = None
    try:
        from oracle.make_ref import load_reference_model
        ref_model, ref_topology = load_reference_model(weights)
        kind, note = "reference", "unmodified reference files (oracle/_ref: model.py, config.py, src/model_operations.py, src/data_encoding.py), torch CPU"

        def fwd(s):
            with torch.no_grad():
                ids0 = ref_topology(s["X"], 64)[0]
                return ref_model(s["X"], ids0 + 1, one_hot_features(s["el"]), dense_membership(s["rid"], s["n_res"]))
    except FileNotFoundError:
        from oracle import pesto_oracle as O

        def fwd(s):
            ids1 = O.extract_topology(s["X"], 64)[0] + 1
            return O.forward(weights, s["X"], ids1, one_hot_features(s["el"]), s["rid"], s["n_res"])

    def head_of(s, n_res):              # the first n_res residues of a structure: a bounded piece of the same workload
        n = int((s["rid"] < n_res).sum())
        return dict(X=s["X"][:n].contiguous(), el=s["el"][:n], rid=s["rid"][:n], n_res=n_res, n_atoms=n)

    probe = head_of(sample[0], 24)
    t0 = time.perf_counter()
    fwd(probe)
    rate = probe["n_atoms"] / (time.perf_counter() - t0)                 # atoms/s (pessimistic: the cost grows ~linearly in atoms)
    budget_atoms = rate * 150.0 / max(args.steps + args.warmup, 1)
    chosen, total = [], 0
    for s in sample:
        if total + s["n_atoms"] <= budget_atoms:
            chosen.append(s)
            total += s["n_atoms"]
    if not chosen:                                                       # not even one whole structure fits: its first residues
        frac = max(budget_atoms / sample[0]["n_atoms"], 0.02)
        chosen = [head_of(sample[0], max(8, int(sample[0]["n_res"] * frac)))]
        total = chosen[0]["n_atoms"]
        desc = f"the first {chosen[0]['n_res']} residues ({total} atoms) of structure 0 of the 53"
    else:
        desc = f"first {len(chosen)} of the 53 structures ({total} atoms), one forward per structure"
    for _ in range(args.warmup):
        [fwd(s) for s in chosen]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        [fwd(s) for s in chosen]
    dt = time.perf_counter() - t0
    value = total * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "atoms/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "fixture: pdbs_test coordinates, shipped i_v4_1 checkpoint",
        "config": {"workload": "configs[1]: i_v4_1 (32 layers, k=64, Ns=32) over pdbs_test; CPU arm runs a bounded sample: " + desc
                               + ", kNN + forward"},
        "cpu_baseline": {"value": value, "unit": "atoms/s", "cores": torch.get_num_threads(), "kind": kind, "sample": desc, "note": note},
        "e2e": {"value": value, "unit": "atoms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def load_peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "MEASURED_PEAKS.json hbm_gbs (measured)"
    except (OSError, ValueError):          # B200_PROFILING.md's stated fallback
        return {"hbm_gbs": 6650.0}, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def staged_layer_times(lib, model, dev, Xd, ids1, q0d, mode, reps):
    """Per-layer kernel times of one forward through the staged C ABI (CUDA events on the launching stream inside
    pesto_state_update_timed): {nn: [edge kernel ms]}, [per-atom kernel ms]."""
    from pesto_b200 import _lib
    h = model._handle(dev.index)
    n = int(Xd.shape[0])
    st = [torch.empty((n + 1, 128), device=dev) for _ in range(2)]
    ids32 = torch.empty((n, 64), dtype=torch.int32, device=dev)
    geom = torch.empty((n, 64, 4), device=dev)
    scratch = torch.zeros(16, dtype=torch.uint8, device=dev)
    node = torch.empty(lib.pesto_node_scratch_bytes(n), dtype=torch.uint8, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    per_nn, node_ms = {}, []
    for _ in range(reps):
        _lib.check(lib.pesto_prologue(h, Xd.data_ptr(), ids1.data_ptr(), 64, q0d.data_ptr(), n, st[0].data_ptr(),
                                      ids32.data_ptr(), geom.data_ptr(), scratch.data_ptr(), stream), "prologue")
        cur = 0
        for layer in range(lib.pesto_model_num_layers(h)):
            a, b = ctypes.c_float(), ctypes.c_float()
            _lib.check(lib.pesto_state_update_timed(h, layer, n, ids32.data_ptr(), geom.data_ptr(), st[cur].data_ptr(),
                                                    st[1 - cur].data_ptr(), node.data_ptr(), mode, stream,
                                                    ctypes.byref(a), ctypes.byref(b)), "state_update_timed")
            cur = 1 - cur
            per_nn.setdefault(lib.pesto_model_layer_nn(h, layer), []).append(b.value)
            node_ms.append(a.value)
    return per_nn, node_ms


def edge_kernel_ms(lib, model, dev, Xd, ids1, q0d, mode, nn=64, reps=8):
    """Average duration of the fused edge kernel of the nn-neighbour layers: for each such layer, `reps` launches back to
    back between two CUDA events on the launching stream (pesto_edge_kernel_timed), on the state that layer sees in a
    real forward (the staged pipeline advances the state layer by layer)."""
    from pesto_b200 import _lib
    h = model._handle(dev.index)
    n = int(Xd.shape[0])
    st = [torch.empty((n + 1, 128), device=dev) for _ in range(2)]
    ids32 = torch.empty((n, 64), dtype=torch.int32, device=dev)
    geom = torch.empty((n, 64, 4), device=dev)
    scratch = torch.zeros(16, dtype=torch.uint8, device=dev)
    node = torch.empty(lib.pesto_node_scratch_bytes(n), dtype=torch.uint8, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(lib.pesto_prologue(h, Xd.data_ptr(), ids1.data_ptr(), 64, q0d.data_ptr(), n, st[0].data_ptr(),
                                  ids32.data_ptr(), geom.data_ptr(), scratch.data_ptr(), stream), "prologue")
    cur, out = 0, []
    for layer in range(lib.pesto_model_num_layers(h)):
        if lib.pesto_model_layer_nn(h, layer) == nn:
            ms = ctypes.c_float()
            _lib.check(lib.pesto_edge_kernel_timed(h, layer, n, ids32.data_ptr(), geom.data_ptr(), st[cur].data_ptr(), node.data_ptr(),
                                                   mode, reps, stream, ctypes.byref(ms)), "edge_kernel_timed")
            out.append(ms.value)
        _lib.check(lib.pesto_state_update(h, layer, n, ids32.data_ptr(), geom.data_ptr(), st[cur].data_ptr(), st[1 - cur].data_ptr(),
                                          node.data_ptr(), mode, stream), "state_update")
        cur = 1 - cur
    return float(np.mean(out))


def event_ms(fn, steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("PESTO_MODE", "f16x3"), choices=["fp32", "f16x3", "f16", "bf16x3", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other-configuration rows (roofline.by_config, config5)")
    ap.add_argument("--config5-structures", type=int, default=1500, help="synthetic AlphaFold-sized structures per GPU in the config5 row")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch.distributed as dist
    from pesto_b200 import _lib
    from pesto_b200.model import Model
    from pesto_b200.data_encoding import batch_topology, extract_topology, std_elements
    from pesto_b200.runner import predict_structures, encode_batch
    from pesto_b200.sharding import rank_shard
    from pesto_b200.synth import one_hot_features, synth_structure, interfaceome_sizes, BASE_SEED

    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner does, from C) are sent
    # to stderr for the duration of the run, and the line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([float(x)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([float(x)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- model and workload ---------------------------------------------------------------------------------
    # Weak scaling over structures, really sharded: the job is `world` copies of configs[1] (53 * world structures);
    # sharding.rank_shard (LPT by atoms) gives every rank its structures, no data-path collective.  At world = 1 the
    # shard is the 53 structures themselves.
    weights = load_weights()
    with open(os.path.join(GOLDEN, f"config_{TAG}.json")) as fh:
        model = Model(json.load(fh), mode=args.mode)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    model = model.eval().to(dev)
    wl = load_workload()
    all_structs = structure_dicts(wl)
    job_sizes = [len(s["xyz"]) for s in all_structs] * world
    mine = rank_shard(job_sizes, rank, world)
    structs = [all_structs[i % len(all_structs)] for i in mine]
    sizes = [len(s["xyz"]) for s in structs]
    Xn, eln, ridn, _, n_rs = encode_batch(structs, as_index=True)
    n_atoms, n_res = int(Xn.shape[0]), int(sum(n_rs))
    Xh = torch.from_numpy(Xn).pin_memory()
    elh = torch.from_numpy(eln).pin_memory()          # element column per atom: one byte over PCIe instead of the 30-float
    ridh = torch.from_numpy(ridn).pin_memory()        # one-hot row, which is expanded on the device (as pesto_b200.runner does)
    Xd, ridd = Xh.to(dev), ridh.to(dev)
    q0d = torch.nn.functional.one_hot(elh.to(dev).long(), len(std_elements) + 1).to(torch.float32)
    ids1 = batch_topology(Xd, sizes, 64)
    lib = _lib.load()
    launches_fwd = lib.pesto_forward_launch_count(model._handle(local_rank), 0, _lib.MODES[args.mode])

    def step_resident():
        return model(Xd, ids1, q0d, ridd, n_res=n_res)

    zbuf = torch.empty((n_res, 5), dtype=torch.float32).pin_memory()

    def step_e2e():
        X = Xh.to(dev, non_blocking=True)
        q0 = torch.nn.functional.one_hot(elh.to(dev, non_blocking=True).long(), q0d.shape[1]).to(torch.float32)
        rid = ridh.to(dev, non_blocking=True)
        ids = batch_topology(X, sizes, 64)
        z = model(X, ids, q0, rid, n_res=n_res)
        zbuf.copy_(z, non_blocking=True)
        return z

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        return max_over_ranks(ms)

    with torch.no_grad():
        for _ in range(args.warmup):
            z_res = step_resident()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        t_begin = time.perf_counter()
        ms_total = timed(step_resident, args.steps)
        t_end = time.perf_counter()
        clocks = sampler.stop(t_begin, t_end) if sampler else None
        for _ in range(2):
            z_e2e = step_e2e()
        ms_e2e = timed(step_e2e, args.steps)
        torch.cuda.synchronize()
        model.raise_if_failed(dev)
        same = bool(torch.equal(z_res, z_e2e))
        total_atoms = sum_over_ranks(n_atoms)

        # ---- the same job from structure dictionaries through the public many-structures runner: host encoding (element
        #      strings -> index, residue index), pinned H2D on a copy stream, kNN, forward, D2H, per-structure results
        list(predict_structures(model, structs, device=dev))                          # warm-up (pinned staging buffers sized for this shard)
        barrier()
        t0 = time.perf_counter()
        n_out = sum(int(z.shape[0]) for _, z in predict_structures(model, structs, device=dev))
        torch.cuda.synchronize()
        s_runner = max_over_ranks(time.perf_counter() - t0)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            encode_batch(structs, as_index=True)
        host_atoms_per_s = n_atoms * 3 / (time.perf_counter() - t0)
        assert n_out == n_res

        # ---- roofline of the dominant kernel: fused per-edge StateUpdate kernel, nn = 64 layers -----------------
        roof = None
        peaks, peak_src = load_peaks()
        if rank == 0:
            mode = _lib.MODES[args.mode]
            per_nn, node_ms = staged_layer_times(lib, model, dev, Xd, ids1, q0d, mode, min(args.steps, 3))
            tc_mode = mode != _lib.MODE_FP32
            # the dominant kernel's average launch duration: launches back to back between two events (tensor-core modes);
            # per_nn holds the durations of single launches bracketed by events inside a staged forward (launch gap included)
            ms64 = edge_kernel_ms(lib, model, dev, Xd, ids1, q0d, mode) if tc_mode else float(np.mean(per_nn[64]))
            alg_bytes = n_atoms * (64 * 536 + 1024)
            achieved = alg_bytes / (ms64 * 1e-3) / 1e9
            edge_ms_per_fwd = sum(float(np.mean(v)) * 8 for v in per_nn.values())
            traffic, traffic_src = None, None
            tpath = os.path.join(REPO, "profiles", "edge64_traffic.json")
            if os.path.exists(tpath):              # dram__bytes_read + write of this kernel on this workload (one ncu capture)
                with open(tpath) as fh:
                    tj = json.load(fh)
                traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
            ms_fwd = ms_total / args.steps
            roof = {"bound": "hbm", "kernel": "edge_kernel_tc (fused StateUpdate edge kernel, nn=64)", "achieved": achieved,
                    "peak": peaks["hbm_gbs"], "peak_source": peak_src, "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src, "ms_per_launch": ms64,
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "timing": "CUDA events on the launching stream around 8 back-to-back launches of each nn=64 layer's edge kernel "
                              "(pesto_edge_kernel_timed), mean over the 8 layers; edge_kernel_ms_by_nn = single launches bracketed by "
                              "events inside a staged forward",
                    "edge_kernel_ms_by_nn": {str(k): float(np.mean(v)) for k, v in sorted(per_nn.items())},
                    "node_kernel_ms": float(np.mean(node_ms)),
                    "edge_kernel_share_of_step": edge_ms_per_fwd / ms_fwd,
                    "whole_forward": {"algorithmic_bytes": n_atoms * 547328, "ms": ms_fwd,
                                      "frac": n_atoms * 547328 / (ms_fwd * 1e-3) / 1e9 / peaks["hbm_gbs"]}}

            # ---- the other BASELINE configurations on this GPU (synthetic point clouds of SURVEY.md A.6, shipped weights):
            #      the north_star target (N = 8192, k = 64) and the large chain (N = 32 768), plus config 1's single structure
            if not args.no_extras:
                by = {}
                for label, n_syn in (("north_star N=8192", 8192), ("configs[3] N=32768", 32768)):
                    Xs, els, rids = synth_structure(n_syn, BASE_SEED)
                    Xs_d = Xs.to(dev)
                    ids_s = extract_topology(Xs_d, 64)[0] + 1
                    q0_s, rid_s = one_hot_features(els).to(dev), rids.int().to(dev)
                    nres_s = int(rids.max()) + 1
                    for _ in range(3):
                        model(Xs_d, ids_s, q0_s, rid_s, n_res=nres_s)
                    reps = max(10, 400000 // n_syn)         # the 8192-atom working set fits in L2 (as SURVEY 8d notes); timed back to back
                    ms_f = event_ms(lambda: model(Xs_d, ids_s, q0_s, rid_s, n_res=nres_s), reps)
                    pn, nm = staged_layer_times(lib, model, dev, Xs_d, ids_s, q0_s, mode, 5)
                    k64 = edge_kernel_ms(lib, model, dev, Xs_d, ids_s, q0_s, mode, reps=16) if tc_mode else float(np.mean(pn[64]))
                    by[label] = {"n_atoms": n_syn, "nn64_edge_kernel_ms": k64,
                                 "nn64_frac": n_syn * (64 * 536 + 1024) / (k64 * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                 "nn64_target_ms_at_0.70": n_syn * (64 * 536 + 1024) / (0.70 * peaks["hbm_gbs"] * 1e9) * 1e3,
                                 "edge_kernel_ms_by_nn": {str(k): float(np.mean(v)) for k, v in sorted(pn.items())},
                                 "node_kernel_ms": float(np.mean(nm)), "forward_ms": ms_f, "atoms_per_s": n_syn / (ms_f * 1e-3),
                                 "forward_frac": n_syn * 547328 / (ms_f * 1e-3) / 1e9 / peaks["hbm_gbs"]}
                # configs[2]: i_v4_0 (16 layers) over a batch of 32 synthetic 8192-atom structures (262 144 atoms), one topology
                # call + one forward per step
                with open(os.path.join(GOLDEN, "config_i_v4_0.json")) as fh:
                    m40 = Model.for_state_dict(json.load(fh), {k: torch.from_numpy(v) for k, v in
                                                                 np.load(os.path.join(GOLDEN, "weights_i_v4_0.npz")).items()},
                                               mode=args.mode).eval().to(dev)
                parts = [synth_structure(8192, BASE_SEED + 100 + i) for i in range(32)]
                Xb = torch.cat([p_[0] for p_ in parts]).to(dev)
                qb = one_hot_features(torch.cat([p_[1] for p_ in parts])).to(dev)
                nres_each = [int(p_[2].max()) + 1 for p_ in parts]
                roff_b = np.concatenate([[0], np.cumsum(nres_each)])
                ridb = torch.cat([p_[2] + int(roff_b[i]) for i, p_ in enumerate(parts)]).int().to(dev)
                nres_b = int(roff_b[-1])
                ids_b = batch_topology(Xb, [8192] * 32, 64)
                for _ in range(2):
                    m40(Xb, ids_b, qb, ridb, n_res=nres_b)
                ms_knn = event_ms(lambda: batch_topology(Xb, [8192] * 32, 64), 5)
                ms_b = event_ms(lambda: m40(Xb, ids_b, qb, ridb, n_res=nres_b), 5)
                m40.raise_if_failed(dev)
                by["configs[2] i_v4_0 (16 layers), batch of 32 x N=8192"] = {
                    "n_atoms": int(Xb.shape[0]), "structures": 32, "forward_ms": ms_b, "topology_ms": ms_knn,
                    "atoms_per_s": int(Xb.shape[0]) / (ms_b * 1e-3), "atoms_per_s_with_topology": int(Xb.shape[0]) / ((ms_b + ms_knn) * 1e-3),
                    "forward_frac": int(Xb.shape[0]) * (547328 // 2) / (ms_b * 1e-3) / 1e9 / peaks["hbm_gbs"]}
                del m40, Xb, qb, ridb, ids_b, parts
                s0 = structs[0]
                X1 = torch.from_numpy(s0["xyz"]).to(dev)
                e1, r1 = encode_batch([s0], as_index=True)[1:3]
                q1 = torch.nn.functional.one_hot(torch.from_numpy(e1).to(dev).long(), q0d.shape[1]).float()
                r1 = torch.from_numpy(r1).to(dev)
                i1 = extract_topology(X1, 64)[0] + 1
                nr1 = int(r1.max()) + 1
                for _ in range(3):
                    model(X1, i1, q1, r1, n_res=nr1)
                ms_1 = event_ms(lambda: model(X1, i1, q1, r1, n_res=nr1), 50)
                by["configs[0] single structure"] = {"n_atoms": int(X1.shape[0]), "forward_ms": ms_1,
                                                     "atoms_per_s": int(X1.shape[0]) / (ms_1 * 1e-3), "launches": int(launches_fwd)}
                # trajectory mode (md_analysis/apply_model_md.ipynb cell 6): 64 frames of that structure around its own coordinates,
                # frame-0 topology, frames batched as structures (pesto_b200.md)
                from pesto_b200.md import predict_trajectory
                gtr = torch.Generator().manual_seed(5)
                Xt = torch.from_numpy(s0["xyz"]).unsqueeze(1) + 0.3 * torch.randn(X1.shape[0], 64, 3, generator=gtr)
                Xt = Xt.to(dev)
                predict_trajectory(model, Xt, i1, q1, r1, device=dev)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                predict_trajectory(model, Xt, i1, q1, r1, device=dev)
                torch.cuda.synchronize()
                s_tr = time.perf_counter() - t0
                by["trajectory mode (64 frames of that structure)"] = {"n_atoms": int(X1.shape[0]), "frames": 64, "seconds": s_tr,
                                                                       "frames_per_s": 64 / s_tr, "atoms_per_s": 64 * int(X1.shape[0]) / s_tr}
                roof["by_config"] = by

            # ---- host side of the apply path: PDB text -> C++ parser -> preprocessing -> features, per file;
            #      the reference publishes 53 ms load (gemmi) + 69 ms processing per structure on its own hardware
            host = None
            if not args.no_extras:
                import glob
                import gzip
                import tempfile
                from pesto_b200.dataset import StructuresDataset
                from pesto_b200.structure import concatenate_chains
                from pesto_b200.data_encoding import encode_structure, encode_features
                rows_h = []
                with tempfile.TemporaryDirectory() as tmp:
                    for gz in sorted(glob.glob(os.path.join(GOLDEN, "pdb", "*.pdb.gz"))):
                        if "_i" in os.path.basename(gz):
                            continue
                        path = os.path.join(tmp, os.path.basename(gz)[:-3])
                        with gzip.open(gz, "rb") as fi, open(path, "wb") as fo:
                            fo.write(fi.read())
                        best = None
                        for _ in range(3):
                            t0 = time.perf_counter()
                            subunits, _p = StructuresDataset([path], with_preprocessing=True)[0]
                            t1 = time.perf_counter()
                            st_ = concatenate_chains(subunits)
                            Xe, Me = encode_structure(st_)
                            qe = encode_features(st_)[0]
                            t2 = time.perf_counter()
                            if best is None or t2 - t0 < best[0]:
                                best = (t2 - t0, t1 - t0, t2 - t1, int(Xe.shape[0]))
                        rows_h.append({"file": os.path.basename(path), "atoms": best[3], "read_and_preprocess_ms": 1e3 * best[1],
                                       "encode_ms": 1e3 * best[2]})
                tot_a = sum(r["atoms"] for r in rows_h)
                tot_ms = sum(r["read_and_preprocess_ms"] + r["encode_ms"] for r in rows_h)
                host = {"files": rows_h, "ms_per_structure": tot_ms / len(rows_h), "atoms_per_s_per_process": tot_a / (tot_ms * 1e-3),
                        "reference_published_ms_per_structure": {"load": 53, "process": 69, "note": "BASELINE.md / SURVEY.md section 6, the reference's own hardware"},
                        "includes": "read_pdb (C++ fixed-column parser) + clean/split/filter preprocessing, then concatenate_chains + encode_structure + encode_features"}
                roof["host_pipeline"] = host

        # ---- BASELINE configs[4] ("interfaceome scale") as a sharded job: synthetic AlphaFold-sized structures (residue counts
        #      clip(round(exp(N(5.8, 0.7))), 16, 2700), 8 atoms per residue), `--config5-structures` per GPU, LPT-sharded over the
        #      ranks, through runner.predict_structures; wall clock between barriers, max over ranks
        cfg5 = None
        if not args.no_extras:
            n5 = args.config5_structures * world
            sz5 = (interfaceome_sizes(n5) * 8).tolist()
            mine5 = rank_shard(sz5, rank, world)
            st5 = []
            for i in mine5:
                Xs, els, rids = synth_structure(int(sz5[i]), BASE_SEED + i)
                st5.append({"xyz": Xs.numpy(), "element": std_elements[els.numpy()], "resid": rids.numpy() + 1})
            atoms5 = sum(int(sz5[i]) for i in mine5)
            list(predict_structures(model, st5[:4], device=dev))
            barrier()
            t0 = time.perf_counter()
            for _ in predict_structures(model, st5, device=dev):
                pass
            torch.cuda.synchronize()
            s5 = max_over_ranks(time.perf_counter() - t0)
            t0 = time.perf_counter()
            encode_batch(st5, as_index=True)
            h5 = time.perf_counter() - t0
            tot5 = sum_over_ranks(atoms5)
            imb = max_over_ranks(atoms5) / (tot5 / world)
            cfg5 = {"workload": f"configs[4] subset: {n5} synthetic AlphaFold-sized structures ({args.config5_structures} per GPU, "
                                f"the 20 000-structure job scaled to a bench run), i_v4_1, LPT-sharded over {world} GPU(s)",
                    "structures": n5, "atoms": tot5, "seconds": s5, "atoms_per_s_e2e": tot5 / s5,
                    "lpt_max_over_mean_atoms": imb, "host_encode_share_of_wall": h5 / s5,
                    "includes": "host encoding from structure dictionaries, pinned H2D (copy stream, one batch ahead), segmented kNN, "
                                "forward, D2H, per-structure logits; wall clock between barriers, max over ranks"}

    # ---- CPU baseline (rank 0, N = 1 only): oracle on a bounded sample ------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count())
        sample = cpu_sample(wl)
        ids1_s, t_knn = cpu_topology(sample)
        secs, z_cpu = cpu_forward_seconds(weights, sample, ids1_s)
        # the GPU logits of the same structures: the shard is in LPT order, so look each sample structure up by its index
        roff = np.concatenate([[0], np.cumsum(n_rs)])
        pos = {int(i) % len(all_structs): k for k, i in enumerate(mine)}
        z_gpu = torch.cat([z_res[int(roff[pos[s["index"]]]):int(roff[pos[s["index"]] + 1])] for s in sample]).cpu()
        cpu = {"value": sum(s["n_atoms"] for s in sample) / secs, "unit": "atoms/s", "cores": torch.get_num_threads(),
               "kind": "port", "sample": sample_desc(sample) + f", {secs:.2f} s; CPU kNN {t_knn:.2f} s",
               "note": "torch-CPU oracle port (faster than the unmodified reference, which `--impl reference` times from oracle/_ref)",
               "max_abs_logit_diff_vs_gpu": float((torch.cat(z_cpu) - z_gpu).abs().max())}

    if rank == 0:
        out = {
            "metric": METRIC, "value": total_atoms * args.steps / (ms_total * 1e-3), "unit": "atoms/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32", "f16x3": "f16x3 (3-term split product over fp16 hi/lo planes on tcgen05, fp32 accumulate/state)", "f16": "f16"}[{"bf16x3": "f16x3", "bf16": "f16"}.get(args.mode, args.mode)],
            "data": "fixture: pdbs_test coordinates/elements/residue ids (tests/golden/pdbs_test_53.npz), shipped i_v4_1 checkpoint",
            "config": {"workload": "configs[1]: i_v4_1 (32 layers, k=64, Ns=32) over the 53 pdbs_test structures; the job is one copy of "
                                   "the set per GPU, LPT-sharded by structure over the ranks (sharding.rank_shard), one collated batch per step and rank",
                       "atoms_per_step_per_gpu": n_atoms, "residues_per_step_per_gpu": n_res, "structures_per_gpu": len(sizes),
                       "structures_total": len(job_sizes), "mode": args.mode,
                       "parallelism": f"sharded by structure (LPT) over {world} rank(s), no data-path collective",
                       "l2": "no flush needed: per-step working set (state 136 MB + geometry 136 MB + node factors 348 MB) exceeds the 126 MB L2"},
            "clocks": clocks,
            "e2e": {"value": total_atoms * args.steps / (ms_e2e * 1e-3), "unit": "atoms/s",
                    "h2d_bytes_per_step": int(Xh.numel() * 4 + elh.numel() + ridh.numel() * 4),
                    "d2h_bytes_per_step": int(zbuf.numel() * 4), "ms_per_step": ms_e2e / args.steps,
                    "includes": "pinned H2D of X / element index (uint8, one-hot expanded on the device) / residue index, kNN topology (4 launches), forward, D2H of logits",
                    "logits_equal_resident_run": same,
                    "runner": {"atoms_per_s": total_atoms / s_runner, "seconds": s_runner, "host_atoms_per_s_per_process": host_atoms_per_s,
                               "includes": "the same shard as structure dictionaries through runner.predict_structures: host encoding, "
                                           "pinned H2D on a copy stream, kNN, forward, D2H, per-structure logits (wall clock, max over ranks)"}},
            "gpu_launches": int(launches_fwd * args.steps),
            "roofline": roof,
            "cpu_baseline": cpu,
            "config5": cfg5,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
