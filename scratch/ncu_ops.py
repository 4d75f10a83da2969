"""Dev tool: executed-instruction histogram by SASS opcode for one kernel of an ncu report: ncu_ops.py report.ncu-rep kernel-regex"""
import csv, collections, re, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = None
cnt = collections.Counter(); smp = collections.Counter(); tot = 0
for r in rows:
    if len(r) > 5 and r[0] == 'Address':
        if h is not None: break          # first kernel instance only
        h = r; si = h.index('Source'); ii = h.index('Instructions Executed'); ss = h.index('# Samples'); continue
    if h is None or len(r) <= ii: continue
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[si].strip())
    if not m: continue
    op = m.group(2); base = op.split('.')[0]; key = base
    if base in ('MUFU', 'F2F', 'I2F', 'F2I', 'LDG', 'STG', 'LDL', 'STL', 'LDS', 'STS', 'I2FP', 'F2FP'): key = '.'.join(op.split('.')[:3])
    try: n = int(r[ii])
    except ValueError: continue
    cnt[key] += n; smp[key] += int(r[ss]); tot += n
print('total', tot)
for k, v in cnt.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 45):
    print(f'{k:28} {v/1e6:7.3f}M {100*v/tot:5.1f}%  samples {smp[k]}')
