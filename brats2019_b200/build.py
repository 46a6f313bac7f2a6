"""Build libbrats_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python -m brats2019_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libbrats_b200.so")
SOURCES = [os.path.join(CSRC, "api.cu")]
DEPS = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), "include", "brats_b200.h")]


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in DEPS)


PROBES_OUT = os.path.join(HERE, "libbrats_b200_probes.so")


def build(force=False, verbose=False, probes=False):
    """probes=True: a second library with -DB200_PROBES (in-kernel cycle accounting and skip-a-stage switches for the
    profiling scripts under tools/; select it with B200_LIB_PATH).  The default library carries none of that code."""
    out = PROBES_OUT if probes else OUT
    if not force and not probes and up_to_date():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC", "-o", out] + (["-DB200_PROBES"] if probes else []) + SOURCES
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print(r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, probes="--probes" in sys.argv))
