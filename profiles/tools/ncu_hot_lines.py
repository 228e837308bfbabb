"""Top SASS lines by warp-stall samples, from `ncu --set full --import-source on` reports.
usage: ncu_hot_lines.py <report.ncu-rep> <kernel-name regex> [nth launch] [top N]"""
import csv, sys, subprocess, collections
rep, kre = sys.argv[1], sys.argv[2]
nth = int(sys.argv[3]) if len(sys.argv) > 3 else 0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 22
cmd = ['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre, '--launch-skip', str(nth), '--launch-count', '1']
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
if len(rows) < 3:
    sys.exit("no rows: " + out[:300])
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= 12 and r[0].startswith('0x')]
col = idx['Warp Stall Sampling (All Samples)']
f = lambda v: float(v) if v not in ('', '-') else 0.0
tot = sum(f(r[col]) for r in data)
print(rows[0][1][:80], 'total samples', tot, 'instructions', len(data))
top = sorted(enumerate(data), key=lambda t: -f(t[1][col]))[:topn]
for i, r in sorted(top):
    st = {k: f(r[idx[k]]) for k in hdr if k.startswith('stall_') and '(Not' not in k}
    st = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(str(i).rjust(5), r[col].rjust(6), r[idx['Instructions Executed']].rjust(9), r[idx['Source']].strip()[:70].ljust(70), st)
b = collections.Counter()
for i, r in enumerate(data):
    b[i // 200] += f(r[col])
print('per-200-instr buckets:', [(k, int(v)) for k, v in sorted(b.items()) if v > 0])
