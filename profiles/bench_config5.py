#!/usr/bin/env python
"""BASELINE config 5 in miniature: S synthetic AlphaFold-sized structures (residue counts clip(round(exp(N(5.8, 0.7))), 16,
2700), 8 atoms per residue) through the many-structures runner -- host encoding, pinned H2D on a copy
stream, segmented kNN, forward, logits back to the host -- sharded over the ranks by cost (LPT), no data-path collective.

    python profiles/bench_config5.py [--structures 2000]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/bench_config5.py

Wall clock around the whole job between barriers (+ cuda synchronize), max over ranks; structures are generated before the
timed region.  Prints one JSON line (rank 0).  Context numbers for profiles/README.md, not the bench line."""
import argparse, json, os, sys, time
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from pesto_b200.model import Model                                   # noqa: E402
from pesto_b200.data_encoding import std_elements                    # noqa: E402
from pesto_b200.runner import predict_structures                     # noqa: E402
from pesto_b200.sharding import rank_shard                           # noqa: E402
from pesto_b200.synth import synth_structure, interfaceome_sizes, BASE_SEED   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--structures", type=int, default=2000)
ap.add_argument("--mode", default="f16x3")
a = ap.parse_args()
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = os.path.join(REPO, "tests", "golden")
model = Model(json.load(open(os.path.join(g, "config_i_v4_1.json"))), mode=a.mode)
model.load_state_dict({k: torch.from_numpy(v) for k, v in np.load(os.path.join(g, "weights_i_v4_1.npz")).items()})
model = model.eval().to(f"cuda:{local}")
sizes = interfaceome_sizes(a.structures) * 8
mine = rank_shard(sizes.tolist(), rank, world)
structures = []
for i in mine:
    X, el, rid = synth_structure(int(sizes[i]), BASE_SEED + i)
    structures.append({"xyz": X.numpy(), "element": std_elements[el.numpy()], "resid": rid.numpy() + 1})
list(predict_structures(model, structures[:8], device=f"cuda:{local}"))          # warm-up


def barrier():
    if world > 1:
        dist.barrier(device_ids=[local])
    torch.cuda.synchronize()


barrier()
t0 = time.perf_counter()
n_res = sum(z.shape[0] for _, z in predict_structures(model, structures, device=f"cuda:{local}"))
barrier()
dt = torch.tensor([time.perf_counter() - t0], device=f"cuda:{local}")
atoms = torch.tensor([float(sum(int(sizes[i]) for i in mine))], device=f"cuda:{local}")
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dist.all_reduce(atoms, op=dist.ReduceOp.SUM)
if rank == 0:
    print(json.dumps({"config": f"configs[4] in miniature: {a.structures} synthetic structures, i_v4_1, sharded by LPT over {world} GPU(s)",
                      "n_gpus": world, "structures": a.structures, "atoms": atoms.item(), "seconds": dt.item(),
                      "atoms_per_s_e2e": atoms.item() / dt.item(), "mode": a.mode,
                      "includes": "host encoding, pinned H2D (copy stream), kNN, forward, D2H; wall clock, max over ranks"}), flush=True)
if world > 1:
    dist.destroy_process_group()
