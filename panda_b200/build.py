"""Build libpanda-cuda.{so,a} for sm_100a with nvcc (cross-compiles without a GPU)."""
import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")


def build(jobs: int | None = None, verbose: bool = False) -> str:
    jobs = jobs or min(8, os.cpu_count() or 1)
    proc = subprocess.run(["make", "-C", CSRC, f"-j{jobs}", "all"], capture_output=True, text=True)
    if proc.returncode != 0 or verbose:
        sys.stderr.write(proc.stdout[-4000:] + proc.stderr[-4000:])
    if proc.returncode != 0:
        raise RuntimeError("building libpanda-cuda failed")
    so = os.path.join(CSRC, "libpanda-cuda.so")
    if not os.path.exists(so):
        raise RuntimeError("libpanda-cuda.so was not produced")
    return so


if __name__ == "__main__":
    print(build(verbose=True))
