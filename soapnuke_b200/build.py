"""Build helper: compiles the in-tree engine library (CUDA, sm_100a) and the CLI.

    python -m soapnuke_b200.build            # engine + CLI
Everything is built IN-TREE (soapnuke_b200/lib, soapnuke_b200/bin) so the artefacts travel with the
repo snapshot to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBDIR = os.path.join(HERE, "lib")
BINDIR = os.path.join(HERE, "bin")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]

ENGINE_SRCS = [os.path.join(HERE, "csrc", "engine.cu"),
               os.path.join(HERE, "host", "report.cpp"),
               os.path.join(HERE, "host", "host_common.cpp")]
def _engine_deps():
    """Every file the engine library is compiled from: its sources plus all headers under csrc/, host/ and include/
    (globbed, so that a new header can never be forgotten and ship a stale .so to the GPU box)."""
    import glob
    deps = list(ENGINE_SRCS)
    for pat in (os.path.join(HERE, "csrc", "*"), os.path.join(HERE, "host", "*.h"), os.path.join(ROOT, "include", "*.h")):
        deps += glob.glob(pat)
    return deps + [os.path.abspath(__file__)]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_engine(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    out = os.path.join(LIBDIR, os.environ.get("SNK_ENGINE_LIB_NAME", "libsnk_engine.so"))
    if not force and not _stale(out, _engine_deps()):
        return out
    cmd = [NVCC] + ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-Wall,-Wno-unknown-pragmas",
                           "-shared", "-cudart", "shared", "-o", out] + ENGINE_SRCS + ["-ldl"]
    cmd += os.environ.get("SNK_CXXFLAGS", "").split()       # tuning builds, e.g. -DSNK_WS_J=2
    if verbose:
        cmd += ["-Xptxas", "-v"]
    subprocess.check_call(cmd)
    return out


def build_all(force=False, verbose=False):
    lib = build_engine(force, verbose)
    cli = None
    cli_build = os.path.join(HERE, "host", "build_cli.py")
    if os.path.exists(cli_build):
        from .host import build_cli
        cli = build_cli.build(force)
    return lib, cli


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
