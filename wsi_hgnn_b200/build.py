"""Build libwsi_hgnn.so in-tree (wsi_hgnn_b200/lib/) with nvcc for sm_100a.

    python -m wsi_hgnn_b200.build

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libwsi_hgnn.so")


def build(verbose: bool = False, jobs: int = 0) -> str:
    jobs = jobs or max(1, (os.cpu_count() or 2))
    cmd = ["make", "-C", CSRC, f"-j{jobs}"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stdout.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building libwsi_hgnn.so failed (see output above)")
    if not os.path.exists(LIB):
        raise RuntimeError(f"{LIB} was not produced")
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
