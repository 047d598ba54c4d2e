"""Build libsem_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(jobs=None, verbose=False):
    jobs = jobs or os.cpu_count() or 4
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), f"-j{jobs}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libsem_b200.so failed")
    return os.path.join(_HERE, "libsem_b200.so")


if __name__ == "__main__":
    print(build(verbose=True))
