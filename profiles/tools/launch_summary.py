#!/usr/bin/env python
"""Per-kernel summary (markdown) of an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
    python profiles/tools/launch_summary.py gpurun_out/<tag>_launches.csv > profiles/<name>_summary.md"""
import collections
import csv
import re
import sys


def main(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    head = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[head]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for r in rows[head + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(r[mu], 1e-6)
        name = re.sub(r"\(.*", "", r[kn])
        name = re.sub(r"pesto::\(anonymous namespace\)::|pesto::|<unnamed>::|void ", "", name)
        name = re.sub(r"<\(int\)(\d+), \(bool\)(\d), \(bool\)(\d)>", r"<\1, \2, \3>", name)
        tot[name] += ms
        cnt[name] += 1
    total = sum(tot.values())
    print("| kernel | launches | total ms | share | mean us |\n|---|---|---|---|---|")
    for name, ms in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"| `{name[:100]}` | {cnt[name]} | {ms:.3f} | {100 * ms / total:.1f} % | {1e3 * ms / cnt[name]:.1f} |")
    edge = sum(ms for n, ms in tot.items() if n.startswith("edge_kernel_tc"))
    print(f"\nedge kernels together: {100 * edge / total:.1f} % of the device time of the captured launches "
          f"({total:.1f} ms over {sum(cnt.values())} launches).")


if __name__ == "__main__":
    main(sys.argv[1])
