"""Opcode histogram + top stall lines of one kernel from an ncu report's source page (SASS view).
usage: sass_hist.py report.ncu-rep launch_index [top_n]"""
import csv, io, subprocess, sys, collections
rep, idx = sys.argv[1], int(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', str(idx), '--launch-count', '1'],
                     capture_output=True, text=True).stdout
lines = raw.splitlines()
print(lines[0][:150])
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
iS, iE, iSamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
ops = collections.Counter(); samp = collections.Counter(); tot = 0; tots = 0
body = []
for r in rows[1:]:
    if len(r) <= iE or not r[iE]: continue
    try: ex = int(r[iE]); sm = int(r[iSamp] or 0)
    except ValueError: continue
    op = r[iS].split()[0] if r[iS] else '?'
    if op.startswith('@'): op = r[iS].split()[1]
    op = op.split('.')[0] + ('.' + op.split('.')[1] if '.' in op and op.split('.')[0] in ('LDS', 'STS', 'LDCU', 'LDG', 'STG', 'LDL', 'STL') else '')
    ops[op] += ex; samp[op] += sm; tot += ex; tots += sm
    body.append((sm, ex, r[iS], [(hdr[i], int(r[i] or 0)) for i in stall_cols if r[i] and int(r[i] or 0) > 0]))
print(f"total warp instructions {tot}, samples {tots}")
for op, c in ops.most_common(30):
    print(f"  {op:12s} {c:12d} {100.0*c/tot:6.2f} %   samples {100.0*samp[op]/max(tots,1):6.2f} %")
print("top stall lines:")
for sm, ex, src, st in sorted(body, key=lambda b: -b[0])[:topn]:
    st = sorted(st, key=lambda kv: -kv[1])[:3]
    print(f"  {sm:7d} {ex:10d}  {src[:70]:70s} {st}")
