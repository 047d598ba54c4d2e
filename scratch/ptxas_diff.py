"""Compare registers / spills of the march3 kernels between two ptxas log directories."""
import re, sys, glob, os, subprocess
def parse(path):
    out = {}
    txt = open(path).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n(?:.*\n)*?ptxas info\s+: Function properties for \1\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s+: Used (\d+) registers", txt):
        out[m.group(1)] = (int(m.group(5)), int(m.group(3)), int(m.group(4)))
    return out
a, b = sys.argv[1], sys.argv[2]
for f in sorted(glob.glob(os.path.join(b, 'sem_march_p*.ptxas.log')), key=lambda x: int(re.search(r'_p(\d+)', x).group(1))):
    old = parse(os.path.join(a, os.path.basename(f))); new = parse(f)
    for k in new:
        if 'march3' not in subprocess.run(['c++filt', k], capture_output=True, text=True).stdout: continue
        if old.get(k) != new[k]:
            name = subprocess.run(['c++filt', k], capture_output=True, text=True).stdout.strip().split('(')[0]
            print(os.path.basename(f)[:14], name, 'old', old.get(k), 'new', new[k])
