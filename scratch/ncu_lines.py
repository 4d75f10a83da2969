"""Aggregate an ncu report's stall samples / instructions per CUDA source line (dev tool)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; data = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) >= 8 and r[0] != '':
        try: data.append((cur, int(r[0]), r[1], int(r[7]), int(r[6])))
        except Exception: pass
tot = sum(d[3] for d in data); tots = sum(d[4] for d in data)
print("total inst", tot, "samples", tots)
for d in sorted(data, key=lambda d: -d[4])[:top]:
    print(f"{d[0][:14]:>14}:{d[1]:<4} inst {100*d[3]/tot:5.1f}% samp {100*d[4]/tots:5.1f}%  {d[2].strip()[:105]}")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); h = rows[0]; v = rows[2]
keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'launch__registers_per_thread', 'launch__grid_size', 'launch__cluster_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio']
for a, b in zip(h, v):
    if a in keys or ('issue_stalled' in a and 'per_issue_active' in a and float(b or 0) > 0.5): print(a, b)
