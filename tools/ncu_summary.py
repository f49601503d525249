#!/usr/bin/env python
"""Text summary of one kernel of an .ncu-rep (--set full): headline metrics, stall-reason totals and the hottest
instructions of the source page.   python tools/ncu_summary.py gpurun_out/prof_mb.ncu-rep > profiles/xxx.txt"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed"]
print("# %s" % rep)
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print("%-82s %-10s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S, I = ix["# Samples"], ix["Instructions Executed"]
stalls = [h for h in hdr if h.startswith("stall_") and "Not" not in h]
tot = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
ns = sum(int(r[S]) for r in data)
print("\n# warp-state samples: %d over %d SASS instructions (%d executed warp-instructions)" % (ns, len(data), sum(int(r[I]) for r in data)))
print("# stall reasons: " + ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(1, ns)) for k, v in sorted(tot.items(), key=lambda t: -t[1])[:9]))
ops = {}
for r in data:
    t = r[ix["Source"]].split()
    o = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c = ops.setdefault(o, [0, 0]); c[0] += int(r[I]); c[1] += int(r[S])
ti = sum(c[0] for c in ops.values())
print("# instruction mix (executed): " + ", ".join("%s %.1f%%" % (o, 100.0 * c[0] / ti) for o, c in sorted(ops.items(), key=lambda t: -t[1][0])[:14]))
print("\n# hottest instructions by samples")
for i, r in sorted(sorted(enumerate(data), key=lambda t: -int(t[1][S]))[:24]):
    st = sorted(((s[6:], int(r[ix[s]])) for s in stalls if int(r[ix[s]]) > 0), key=lambda t: -t[1])[:2]
    print("%5d  %-64s samples %5s  executed %9s  %s" % (i, r[ix["Source"]].strip()[:64], r[S], r[I], st))
