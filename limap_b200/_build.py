"""In-tree build of the CUDA engine (nvcc, sm_100a only). The .so travels to the GPU box with the
repo snapshot; nothing is JIT-compiled at run time. Translation units are compiled in parallel into
limap_b200/lib/obj/ and only when they (or a header) changed."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "liblimap_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".cu")]
    out.append(os.path.join(os.path.dirname(HERE), "include", "limap_b200.h"))
    out.append(os.path.abspath(__file__))
    return out


def _obj(src):
    return os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(p) > t for p in deps)


def needs_build():
    hdr = _headers()
    return _stale(LIB, sources() + hdr)


def build_native(force=False, verbose=False, extra_flags=()):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    hdr = _headers()
    extra = list(extra_flags) + os.environ.get("LIMAP_B200_NVCC_EXTRA", "").split()
    todo = [s for s in sources() if force or extra or _stale(_obj(s), [s] + hdr)]

    def cc(src):
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", "-o", _obj(src), src]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True, env=env)

    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        list(ex.map(cc, todo))
    cmd = [nvcc] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", LIB] + [_obj(s) for s in sources()]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True, env=env)
    return LIB
