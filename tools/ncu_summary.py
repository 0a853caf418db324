"""Summarise an ncu report: key raw metrics + per-source-line instruction shares.  Usage: tools/ncu_summary.py rep [n_lines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'l1tex__throughput.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
 'l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','lts__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_local_op_st.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio']
print("| metric | value | unit |\n|---|---|---|")
for n in want:
    if n in h: i = h.index(n); print("| %s | %s | %s |" % (n, v[i], u[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; hh = None; agg = {}
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == 'Line No': hh = r; ix = hh.index('Instructions Executed'); tx = hh.index('Thread Instructions Executed'); sx = hh.index('# Samples'); continue
    if hh and len(r) > ix and r[0] != '':
        try: n = int(r[ix]); t = int(r[tx]); s = int(r[sx])
        except ValueError: continue
        a = agg.setdefault((cur, int(r[0]), r[1].strip()), [0, 0, 0]); a[0] += n; a[1] += t; a[2] += s
tot = sum(a[0] for a in agg.values()) or 1; tots = sum(a[2] for a in agg.values()) or 1
byf = {}
for (f, l, s), a in agg.items():
    b = byf.setdefault(f, [0, 0]); b[0] += a[0]; b[1] += a[2]
print("\nwarp instructions (source-attributed): %d" % tot)
for f, b in sorted(byf.items(), key=lambda kv: -kv[1][0]): print("  %-28s %5.1f %% inst  %5.1f %% samples" % (f, 100 * b[0] / tot, 100 * b[1] / tots))
print()
for (f, l, s), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print("%5.2f%% inst %5.2f%% smp thr %4.1f  %s:%d | %s" % (100 * a[0] / tot, 100 * a[2] / tots, a[1] / max(a[0], 1), f, l, s[:90]))
