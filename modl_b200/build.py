"""Build libmodl_b200.so (sm_100a) in-tree with nvcc.  `python -m modl_b200.build`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmodl_b200.so")
OBJ = os.path.join(HERE, "csrc", "_obj")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "--extended-lambda"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu") or f.endswith(".cpp"))


def headers():
    inc = os.path.join(os.path.dirname(HERE), "include", "modl_b200.h")
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh") or f.endswith(".h")] + [inc]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, variant=None, defines=()):
    """variant / defines: an A/B build (`libmodl_b200_<variant>.so`, objects in `_obj_<variant>`) compiled with extra -D
    flags; modl_b200._lib loads it when MODL_B200_LIB points at it."""
    nvcc = os.environ.get("NVCC", "nvcc")
    OBJ = os.path.join(HERE, "csrc", "_obj" + ("_" + variant if variant else ""))
    OUT = os.path.join(HERE, "libmodl_b200" + ("_" + variant if variant else "") + ".so")
    os.makedirs(OBJ, exist_ok=True)
    objs, procs = [], []
    hdr = headers()
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdr):
            cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src, out))
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("modl_b200: CUDA build failed")
    if force or procs or _stale(OUT, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    kw = {}
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        kw = dict(variant=sys.argv[i + 1], defines=sys.argv[i + 2].split(","))
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, **kw))
