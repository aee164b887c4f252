"""Build tools/emu/_gen/libemu_jacobi[_<tag>].so: the Jacobi kernels' own source compiled for the host (see gen_emu.py).
usage: build_emu.py [tag] [-DNAME=VALUE ...]     e.g.  build_emu.py il2e -DSF_INNER_LOOP=2 -DSF_EDGE_SPLIT=1"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import gen_emu   # noqa: E402


def build(tag="", defs=()):
    gen_emu.main()
    out = os.path.join(HERE, "_gen", f"libemu_jacobi{'_' + tag if tag else ''}.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = [cxx, "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I" + cuda_inc, "-I" + HERE,
           *defs, os.path.join(HERE, "emu_harness.cpp"), "-o", out]
    subprocess.check_call(cmd)
    return out


def build_stages():
    gen_emu.gen_stages()
    out = os.path.join(HERE, "_gen", "libemu_stages.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.check_call([cxx, "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I" + cuda_inc,
                           "-I" + HERE, os.path.join(HERE, "emu_stages.cpp"), "-o", out])
    return out


if __name__ == "__main__":
    args = sys.argv[1:]
    tag = args[0] if args and not args[0].startswith("-D") else ""
    print(build(tag, [a for a in args if a.startswith("-D")]))
