"""Build tests/hostsim/_hostsim.so (g++, no CUDA).  TEST INFRASTRUCTURE ONLY."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_hostsim.so")


def build(force=False):
    src = os.path.join(HERE, "hostsim.cpp")
    csrc = os.path.join(os.path.dirname(os.path.dirname(HERE)), "advanced-lane-detection_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return SO
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", SO, src])
    return SO


if __name__ == "__main__":
    print(build(True))
