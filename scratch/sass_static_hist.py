"""Static SASS opcode histogram of kernels in an object file (cuobjdump -sass), no GPU needed.
usage: sass_static_hist.py file.o 'regex on the demangled kernel name' [top_n]"""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], re.compile(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 24
sass = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
cur, hist = None, {}
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['cu++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = name if pat.search(name) else None
        if cur: hist[cur] = collections.Counter()
        continue
    if cur is None: continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)', line)
    if m:
        op = m.group(1).split('.')
        key = op[0]
        if key in ('LDS', 'STS', 'LDG', 'STG', 'LD', 'ST', 'LDCU', 'LDC', 'LDL', 'STL', 'UBLKCP', 'SYNCS', 'ATOMS', 'DFMA', 'DADD', 'DMUL'):
            key = '.'.join(o for o in op if o in (key, '64', '128', 'E', 'S', 'G')) if key not in ('DFMA', 'DADD', 'DMUL') else key
        hist[cur][key] += 1
for name, h in hist.items():
    tot = sum(h.values())
    print(f"{name}\n  static instructions: {tot}")
    for op, c in h.most_common(topn):
        print(f"    {op:16s} {c:7d} {100.0 * c / tot:6.2f} %")
