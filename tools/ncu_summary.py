"""Summarise an ncu report exported with `--page raw --csv` and `--page source --csv`."""
import collections
import csv
import sys

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__cycles_elapsed.avg',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']
for k in keys:
    for h in d:
        if h == k:
            print(f"{h:78s} {d[h][0]:16s} {d[h][1]}")
tot = 0
st = {}
for h in d:
    if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued'):
        v = int(float(d[h][1] or 0))
        st[h.replace('smsp__pcsamp_warps_issue_stalled_', '')] = v
        tot += v
print("stall samples:", ", ".join(f"{k}={v/tot*100:.1f}%" for k, v in sorted(st.items(), key=lambda x: -x[1]) if v / tot > 0.01))
rows = list(csv.reader(open(src)))
hdr = rows[1]
data = rows[2:]
ie = hdr.index("Instructions Executed")
isamp = hdr.index("# Samples")
tot = sum(int(r[ie]) for r in data)
print("SASS instrs:", len(data), "(%.1f KB)" % (len(data) * 16 / 1024), "warp-instrs executed:", tot)
s = sorted(data, key=lambda r: -int(r[ie]))
for frac in (0.5, 0.9, 0.99):
    acc = n = 0
    for r in s:
        acc += int(r[ie]); n += 1
        if acc >= frac * tot:
            break
    print(f"  {frac:.2f} of executed instrs from {n} SASS instrs ({n*16/1024:.1f} KB)")
cnt = collections.Counter(int(r[ie]) for r in data)
print("top (exec count x #instrs):")
for v, n in sorted(cnt.items(), key=lambda x: -x[0] * x[1])[:12]:
    print(f"  exec={v:>12d} x {n:5d} -> {v*n/tot*100:5.1f}%")

# ---- executed-instruction mix from the SASS page (thread-level) ----
ith = hdr.index("Predicated-On Thread Instructions Executed")
isrc = hdr.index("Source")
mix = collections.Counter()
for r in data:
    toks = r[isrc].replace("@", " @").split()
    op = next((t for t in toks if not t.startswith("@") and not t.startswith("!")), "?").split(".")[0]
    mix[op] += int(r[ith])
tot_th = sum(mix.values())
dur_ms = float(d['gpu__time_duration.sum'][1]) if 'gpu__time_duration.sum' in d else 0.0
flop = mix["FADD"] + mix["FMUL"] + 2 * mix["FFMA"]
print(f"thread-instr mix (top): " + ", ".join(f"{k}={v/tot_th*100:.1f}%" for k, v in mix.most_common(14)))
if dur_ms:
    print(f"executed fp32 arithmetic: FFMA={mix['FFMA']:.3e} FADD={mix['FADD']:.3e} FMUL={mix['FMUL']:.3e} -> {flop/(dur_ms*1e-3)/1e12:.2f} TFLOP/s "
          f"(FMA=2) in {dur_ms:.3f} ms; all fp32-pipe ops (incl. FSETP/FMNMX/FSEL) = "
          f"{(mix['FFMA']+mix['FADD']+mix['FMUL']+mix['FSETP']+mix['FMNMX']+mix['FMNMX3']+mix['FSEL'])/tot_th*100:.1f}% of thread instrs")
