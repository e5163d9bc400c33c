"""Builds the drop-in CLI `soapnuke_b200/bin/SOAPnuke` (host C++ driver linked against the engine)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
SRCS = [os.path.join(HERE, f) for f in ("main.cpp", "cli_params.cpp", "process.cpp", "gz_members.cpp", "fast_deflate.cpp")]
import glob
DEPS = SRCS + glob.glob(os.path.join(HERE, "*.h")) + glob.glob(os.path.join(PKG, "csrc", "*.cuh")) + \
    glob.glob(os.path.join(os.path.dirname(PKG), "include", "*.h")) + [os.path.abspath(__file__)]


def build(force=False):
    out_dir = os.path.join(PKG, "bin")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "SOAPnuke")
    lib = os.path.join(PKG, "lib", "libsnk_engine.so")
    deps = DEPS + [lib]
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps if os.path.exists(d)):
        return out
    cmd = ["g++", "-std=c++17", "-O3", "-Wall", "-Wno-unknown-pragmas", "-pthread", "-I/usr/local/cuda/include", "-o", out] + SRCS + \
          ["-L" + os.path.join(PKG, "lib"), "-lsnk_engine", "-lz", "-ldl", "-Wl,-rpath,$ORIGIN/../lib"]
    subprocess.check_call(cmd)
    return out
