"""Summarise an .ncu-rep of the window kernel: key metrics + dynamic opcode mix per warp-attempt.
usage: python tools/ncu_summary.py REPORT.ncu-rep ATTEMPTS_PER_LAUNCH [out.csv]"""
import collections, csv, io, re, subprocess, sys

rep, attempts = sys.argv[1], float(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keep = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg',
        'smsp__warps_eligible.avg.per_cycle_active', 'sass__inst_executed_register_spilling',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'] + \
       [f'smsp__average_warps_issue_stalled_{k}_per_issue_active.ratio' for k in
        ('wait', 'no_instruction', 'math_pipe_throttle', 'branch_resolving', 'not_selected', 'short_scoreboard',
         'dispatch_stall', 'long_scoreboard')]
lines = ["metric,unit," + ",".join(f"launch_{i}" for i in range(len(data)))]
for i, h in enumerate(hdr):
    if h in keep:
        lines.append(f"{h},{units[i]}," + ",".join(r[i] for r in data))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
i0 = starts[0]
h2 = rows[i0]
body = rows[i0 + 1:(starts[1] - 1 if len(starts) > 1 else len(rows))]
ci, cs, cn = h2.index('Instructions Executed'), h2.index('Source'), h2.index('# Samples')
tot, samp, static = collections.Counter(), collections.Counter(), collections.Counter()
for r in body:
    if len(r) <= ci:
        continue
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[cs].strip())
    if not m:
        continue
    op = m.group(2).split('.')[0]
    tot[op] += int(r[ci]); samp[op] += int(r[cn]); static[op] += 1
total = sum(tot.values())
wa = attempts / 32.0
lines.append(f"# static SASS instructions,{len(body)},dynamic warp instructions,{total},per warp-attempt,{total / wa:.1f}")
lines.append("# opcode,share_pct,per_warp_attempt,static,stall_sample_share_pct")
for op, n in tot.most_common(24):
    lines.append(f"# {op},{100 * n / total:.2f},{n / wa:.1f},{static[op]},{100 * samp[op] / max(sum(samp.values()), 1):.1f}")
text = "\n".join(lines)
print(text)
if out:
    open(out, "w").write(text + "\n")
