#!/bin/bash
# Copy the outputs of `bash profiles/sanitize.sh r2; bash profiles/ncu_r2.sh r2z` (run under gpurun) from gpurun_out/ into the
# tracked evidence files under profiles/ and print the numbers DESIGN.md quotes.   usage: bash profiles/tools/refresh_evidence.sh
cd "$(dirname "$0")/../.."
{ for t in memcheck racecheck synccheck; do echo "== compute-sanitizer --tool $t python profiles/run_forward.py --atoms 700 --mode f16x3 (profiles/sanitize.sh r2, B200; final round-2 kernels)"; grep -E "COMPUTE-SANITIZER|ERROR SUMMARY|RACECHECK SUMMARY|^ok" gpurun_out/r2_sanitize_$t.log | sort -u | head -4; done; } > profiles/r2_sanitizer_summary.txt
{ echo '# ncu launch list of `bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras` (profiles/ncu_r2.sh, round 2, final kernels)'; echo; echo '`gpu__time_duration.sum`, `--clock-control none`; per-launch times under ncu are serialised and cold-cache -- the SHARE of the step is what must agree with the bench line (`roofline.edge_kernel_share_of_step`).  Made by `profiles/tools/launch_summary.py`.'; echo; python profiles/tools/launch_summary.py gpurun_out/r2z_launches.csv; } > profiles/r2_launches_bench_summary.md
cp gpurun_out/r2z_launches.csv profiles/r2_launches_bench.csv
for f in edge64_bench edge64_n8192 node_bench; do cp gpurun_out/r2z_${f}_details.txt profiles/r2_${f}_ncu_details.txt; done
python - <<'PY'
import csv, json
rows = list(csv.reader(open('gpurun_out/r2z_edge64_bench_raw.csv')))
h = [r for r in rows if 'dram__bytes_read.sum' in r][0]
i0 = rows.index(h); units = rows[i0 + 1]; vals = rows[i0 + 2]
def val(n):
    i = h.index(n); return float(vals[i].replace(',', '')), units[i]
def tobytes(n):
    v, u = val(n); return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
r, w = tobytes('dram__bytes_read.sum'), tobytes('dram__bytes_write.sum')
t = json.load(open('profiles/edge64_traffic.json'))
t.update(dram_bytes_read=int(r), dram_bytes_write=int(w), dram_bytes_per_launch=int(r + w))
json.dump(t, open('profiles/edge64_traffic.json', 'w'), indent=1)
print('dram read / write MB', r / 1e6, w / 1e6)
for n in ('gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
          'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
          'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct'):
    if n in h:
        print(n, *val(n))
for n in h:
    if 'smsp__average_warps_issue_stalled' in n and n.endswith('per_issue_active.ratio'):
        v, _ = val(n)
        if v >= 0.3:
            print(n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), round(v, 2))
PY
grep -E "Duration|Issue Slots Busy|No Eligible|Executed Ipc Active" profiles/r2_edge64_bench_ncu_details.txt | head -5
