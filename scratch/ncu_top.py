"""Top source lines by executed instructions for one kernel of an ncu report (dev tool): ncu_top.py report.ncu-rep kernel-regex [n]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; data = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) >= 8 and r[0] != '':
        try: data.append((cur, int(r[0]), r[1], int(r[7]), int(r[6])))
        except Exception: pass
tot = sum(d[3] for d in data); ts = sum(d[4] for d in data)
print("total inst", tot, "samples", ts)
for d in sorted(data, key=lambda d: -(d[4] if "--samp" in sys.argv else d[3]))[:top]:
    print(f"{d[0][:14]:>14}:{d[1]:<4} inst {d[3]/1e6:6.2f}M {100*d[3]/max(tot,1):5.1f}% samp {100*d[4]/max(ts,1):5.1f}%  {d[2].strip()[:100]}")
