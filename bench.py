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


def reference_arm(args, rank):
    """CPU implementation of the path (oracle port of the reference's PyTorch forward), rank 0 only."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count())
    wl = load_workload()
    weights = load_weights()
    sample = cpu_sample(wl)
    # bounded sample: time one structure, then keep as many of the 16 as fit ~150 s over all warm-up + timed steps
    ids_probe, _ = cpu_topology(sample[:1])
    t_probe, _ = cpu_forward_seconds(weights, sample[:1], ids_probe)
    n_keep = max(2, min(len(sample), int(150.0 / (max(args.steps + args.warmup, 1) * max(t_probe, 1e-3)))))
    sample = sample[:n_keep]
    ids1, t_knn = cpu_topology(sample)
    for _ in range(args.warmup):
        cpu_forward_seconds(weights, sample, ids1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_forward_seconds(weights, sample, ids1)
    dt = time.perf_counter() - t0
    value = sum(s["n_atoms"] for s in sample) * args.steps / dt
    desc = sample_desc(sample) + ", forward only"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "atoms/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "fixture: pdbs_test coordinates, shipped i_v4_1 checkpoint",
        "config": {"workload": "configs[1]: i_v4_1 (32 layers, k=64, Ns=32) over pdbs_test; CPU arm runs a bounded sample: " + desc},
        "cpu_baseline": {"value": value, "unit": "atoms/s", "cores": torch.get_num_threads(), "kind": "port", "sample": desc,
                         "note": "torch-CPU oracle port of the reference forward (the Python reference cannot travel to the GPU box); "
                                 "it avoids the reference's slow strided torch.norm, so it is faster than the unmodified reference",
                         "knn_seconds": t_knn},
        "e2e": {"value": value, "unit": "atoms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default=os.environ.get("PESTO_MODE", "f16x3"), choices=["fp32", "f16x3", "f16", "bf16x3", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    from pesto_b200.data_encoding import batch_topology
    from pesto_b200.synth import one_hot_features

    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "INFO"):
        os.environ["NCCL_DEBUG"] = "WARN"
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

    # ---- model and workload ---------------------------------------------------------------------------------
    weights = load_weights()
    with open(os.path.join(GOLDEN, f"config_{TAG}.json")) as fh:
        model = Model(json.load(fh), mode=args.mode)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()})
    model = model.eval().to(dev)
    wl = load_workload()
    n_atoms, n_res = int(wl["X"].shape[0]), int(wl["n_res"].sum())
    sizes = [int(s) for s in wl["sizes"]]
    Xh = wl["X"].pin_memory()
    q0h = one_hot_features(wl["el"]).pin_memory()
    ridh = wl["rid"].pin_memory()
    Xd, q0d, ridd = Xh.to(dev), q0h.to(dev), ridh.to(dev)
    ids1 = batch_topology(Xd, sizes, 64)
    lib = _lib.load()
    launches_fwd = lib.pesto_forward_launch_count(model._handle(local_rank), 0, _lib.MODES[args.mode])

    def step_resident():
        return model(Xd, ids1, q0d, ridd, n_res=n_res)

    zbuf = torch.empty((n_res, 5), dtype=torch.float32).pin_memory()

    elh = wl["el"].to(torch.uint8).pin_memory()       # element column per atom: one byte over PCIe instead of the 30-float
                                                      # one-hot row, which is expanded on the device (as pesto_b200.runner does)

    def step_e2e():
        X = Xh.to(dev, non_blocking=True)
        q0 = torch.nn.functional.one_hot(elh.to(dev, non_blocking=True).long(), q0h.shape[1]).to(torch.float32)
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
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        barrier()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

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
        same = bool(torch.equal(z_res, z_e2e))

        # ---- roofline of the dominant kernel: fused per-edge StateUpdate kernel, nn = 64 layers -----------------
        roof = None
        if rank == 0:
            h = model._handle(local_rank)
            n = n_atoms
            st = [torch.empty((n + 1, 128), device=dev) for _ in range(2)]
            ids32 = torch.empty((n, 64), dtype=torch.int32, device=dev)
            geom = torch.empty((n, 64, 4), device=dev)
            scratch = torch.zeros(16, dtype=torch.uint8, device=dev)
            node = torch.empty(lib.pesto_node_scratch_bytes(n), dtype=torch.uint8, device=dev)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            per_nn = {}
            node_ms = []
            mode = _lib.MODES[args.mode]
            for rep in range(min(args.steps, 3)):
                _lib.check(lib.pesto_prologue(h, Xd.data_ptr(), ids1.data_ptr(), 64, q0d.data_ptr(), n, st[0].data_ptr(),
                                              ids32.data_ptr(), geom.data_ptr(), scratch.data_ptr(), stream), "prologue")
                cur = 0
                for layer in range(lib.pesto_model_num_layers(h)):
                    a, b = ctypes.c_float(), ctypes.c_float()
                    _lib.check(lib.pesto_state_update_timed(h, layer, n, ids32.data_ptr(), geom.data_ptr(),
                                                            st[cur].data_ptr(), st[1 - cur].data_ptr(), node.data_ptr(),
                                                            mode, stream, ctypes.byref(a), ctypes.byref(b)), "state_update_timed")
                    cur = 1 - cur
                    per_nn.setdefault(lib.pesto_model_layer_nn(h, layer), []).append(b.value)
                    node_ms.append(a.value)
            peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)"
            try:
                with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as fh:
                    peaks = json.load(fh)
            except (OSError, ValueError):          # B200_PROFILING.md's stated fallback
                peaks, peak_src = {"hbm_gbs": 6650.0}, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"
            ms64 = float(np.mean(per_nn[64]))
            alg_bytes = n * (64 * 536 + 1024)
            achieved = alg_bytes / (ms64 * 1e-3) / 1e9
            edge_ms_per_fwd = sum(float(np.mean(v)) * 8 for v in per_nn.values())
            traffic, traffic_src = None, None
            tpath = os.path.join(REPO, "profiles", "edge64_traffic.json")
            if os.path.exists(tpath):              # dram__bytes_read + write of this kernel on this workload (one ncu capture)
                with open(tpath) as fh:
                    tj = json.load(fh)
                traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
            roof = {"bound": "hbm", "kernel": "edge_kernel_tc (fused StateUpdate edge kernel, nn=64)", "achieved": achieved,
                    "peak": peaks["hbm_gbs"], "peak_source": peak_src, "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src, "ms_per_launch": ms64,
                    "algorithmic_bytes_per_launch": alg_bytes,
                    "edge_kernel_ms_by_nn": {str(k): float(np.mean(v)) for k, v in sorted(per_nn.items())},
                    "node_kernel_ms": float(np.mean(node_ms)),
                    "edge_kernel_share_of_step": edge_ms_per_fwd / (ms_total / args.steps)}

    # ---- CPU baseline (rank 0, N = 1 only): oracle on a bounded sample ------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count())
        sample = cpu_sample(wl)
        ids1_s, t_knn = cpu_topology(sample)
        secs, z_cpu = cpu_forward_seconds(weights, sample, ids1_s)
        n_r = sum(s["n_res"] for s in sample)
        z_gpu = z_res[:n_r].cpu()
        cpu = {"value": sum(s["n_atoms"] for s in sample) / secs, "unit": "atoms/s", "cores": torch.get_num_threads(),
               "kind": "port", "sample": sample_desc(sample) + f", {secs:.2f} s; CPU kNN {t_knn:.2f} s",
               "max_abs_logit_diff_vs_gpu": float((torch.cat(z_cpu) - z_gpu).abs().max())}

    if rank == 0:
        total_atoms = n_atoms * world
        out = {
            "metric": METRIC, "value": total_atoms * args.steps / (ms_total * 1e-3), "unit": "atoms/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32", "f16x3": "f16x3 (3-term split product over fp16 hi/lo planes on tcgen05, fp32 accumulate/state)", "f16": "f16"}[{"bf16x3": "f16x3", "bf16": "f16"}.get(args.mode, args.mode)],
            "data": "fixture: pdbs_test coordinates/elements/residue ids (tests/golden/pdbs_test_53.npz), shipped i_v4_1 checkpoint",
            "config": {"workload": "configs[1]: i_v4_1 (32 layers, k=64, Ns=32) over the 53 pdbs_test structures, one collated batch per step",
                       "atoms_per_step_per_gpu": n_atoms, "residues_per_step_per_gpu": n_res, "structures": len(sizes),
                       "mode": args.mode, "parallelism": f"structures replicated per rank x{world}, no collective",
                       "l2": "no flush needed: per-step working set (state 136 MB + geometry 136 MB + node factors 348 MB) exceeds the 126 MB L2"},
            "clocks": clocks,
            "e2e": {"value": total_atoms * args.steps / (ms_e2e * 1e-3), "unit": "atoms/s",
                    "h2d_bytes_per_step": int(Xh.numel() * 4 + elh.numel() + ridh.numel() * 4),
                    "d2h_bytes_per_step": int(zbuf.numel() * 4), "ms_per_step": ms_e2e / args.steps,
                    "includes": "pinned H2D of X / element index (uint8, one-hot expanded on the device) / residue index, kNN topology (3 launches), forward, D2H of logits",
                    "logits_equal_resident_run": same},
            "gpu_launches": int(launches_fwd * args.steps),
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
