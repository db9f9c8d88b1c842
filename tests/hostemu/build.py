"""TEST ONLY: compiles tests/hostemu/hostemu.cpp (host emulation of the device kernels) with g++."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "libbqa_b200_hostemu.so")
SRC = os.path.join(HERE, "hostemu.cpp")
DEPS = [SRC, os.path.join(HERE, "..", "..", "bqa_b200", "csrc", "bqa_core.cuh"),
        os.path.join(HERE, "..", "..", "include", "bqa_b200.h")]


def build() -> str:
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", SRC, "-o", OUT])
    return OUT


if __name__ == "__main__":
    print(build())
