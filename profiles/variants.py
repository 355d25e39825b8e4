#!/usr/bin/env python
"""Experiment aid: build kernel variants of the tensor-core edge kernel side by side (here, on CPU) and compare them in
ONE gpurun call (there).

    python profiles/variants.py build base= profall=PESTO_PROF_ALL_NN                                # here (name=DEFINE[,DEFINE...])
    python profiles/variants.py run base profall                                                    # under gpurun

`run` prints, per variant, the bench line's value / e2e / per-nn edge-kernel times and the logit error against the
CPU oracle sample is NOT taken (--no-cpu-baseline); parity is checked separately with pytest on the chosen variant."""
import json, os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    cmd, args = sys.argv[1], sys.argv[2:]
    if cmd == "build":
        from pesto_b200.build import build_variant
        for a in args:
            name, _, defs = a.partition("=")
            print(build_variant(name, [d for d in defs.split(",") if d]))
    elif cmd == "run":
        os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
        for name in args:
            env = dict(os.environ, PESTO_B200_LIB=os.path.join(REPO, "pesto_b200", f"libpesto_b200.{name}.so"))
            r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "5", "--warmup", "3", "--no-cpu-baseline", "--no-extras"],
                               env=env, capture_output=True, text=True, timeout=600)
            try:
                d = json.loads(r.stdout.strip().splitlines()[-1])
                rf = d["roofline"]
                print(f"{name:12s} value {d['value']/1e6:6.3f} M  e2e {d['e2e']['value']/1e6:6.3f} M  edge ms "
                      + " ".join(f"{k}:{v:.3f}" for k, v in rf["edge_kernel_ms_by_nn"].items())
                      + f"  node {rf['node_kernel_ms']:.3f}  frac {rf['frac']:.3f}", flush=True)
                with open(os.path.join(REPO, "gpurun_out", f"variant_{name}.json"), "w") as fh:
                    fh.write(r.stdout)
            except Exception as ex:                                   # noqa: BLE001
                print(name, "FAILED", ex, r.stdout[-500:], r.stderr[-1500:], flush=True)


if __name__ == "__main__":
    main()
