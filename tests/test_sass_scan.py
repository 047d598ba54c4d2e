"""Build check on the compiled marching kernels (no GPU: cuobjdump on the objects nvcc cross-compiled): every DFMA of the
sum-factorised contractions must take its table operand from the uniform datapath (`LDCU ... c[0x3][UR+imm]`).  ptxas falls
back to per-thread `LDC` fetches (ADU pipe: the round-2 finding behind P >= 10 running at 28-54 % of the roofline, DESIGN.md
section 3.1) whenever it decides to keep the table offset in a vector register, silently and per instantiation -- so the scan
runs over all of them."""
import collections
import glob
import os
import re
import shutil
import subprocess

import pytest

from tests.conftest import ROOT

BUILD = os.path.join(ROOT, "sem_b200", "csrc", "build")


def _scan(obj):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    ldc, ldcu, tma2d, bulk = (collections.Counter() for _ in range(4))
    fn = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        if fn is None or "sem_march3_kernel" not in fn:
            continue
        if "c[0x3][R" in line:
            ldc[fn] += 1
        elif "c[0x3][UR" in line:
            ldcu[fn] += 1
        elif "UTMALDG" in line:
            tma2d[fn] += 1
        elif "UBLKCP" in line:
            bulk[fn] += 1
    return ldc, ldcu, tma2d, bulk


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="needs the CUDA toolkit's cuobjdump")
def test_table_operands_stay_on_the_uniform_datapath():
    objs = sorted(glob.glob(os.path.join(BUILD, "sem_march_p*.o")))
    if len(objs) < 16:
        pytest.skip("the library has not been built in this tree (python -c 'import __graft_entry__ as g; g.build()')")
    bad, kernels = [], 0
    for obj in objs:
        ldc, ldcu, tma2d, bulk = _scan(obj)
        for fn in ldcu:
            kernels += 1
            total = ldc[fn] + ldcu[fn]
            if total >= 40 and ldc[fn] > 0.2 * total:   # shipped build: at most 9 % (P = 15 NS); P = 1 has 9 fetches in all
                bad.append(f"{os.path.basename(obj)} {fn}: {ldc[fn]} LDC of {total} table fetches")
            assert (tma2d[fn] > 0) != (bulk[fn] > 0), f"{fn}: a kernel stages either by tensor maps or by bulk copies"
    assert kernels >= 180, kernels               # 16 orders x 12 variants (5 modes, pointwise / exchange flags); P = 1 folds some away
    assert not bad, "\n".join(bad)
    # the BASELINE order stages every mode by one 2-D tensor-map copy per field and step
    _, _, tma2d, bulk = _scan(os.path.join(BUILD, "sem_march_p8.o"))
    assert sum(bulk.values()) == 0 and all(v > 0 for v in tma2d.values())
