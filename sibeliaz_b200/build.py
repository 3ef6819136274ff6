"""In-tree build of the sm_100a shared library and the sibeliaz-lcb CLI (explicit nvcc, no JIT cache).

    python -m sibeliaz_b200.build            # build if sources are newer than the artefacts
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
BINDIR = os.path.join(HERE, "bin")
LIB = os.path.join(LIBDIR, "libsibeliaz_lcb.so")
LIB_TINY = os.path.join(LIBDIR, "libsibeliaz_lcb_tiny.so")
CLI = os.path.join(BINDIR, "sibeliaz-lcb")
CLI_GRAPH = os.path.join(BINDIR, "twopaco")
CLI_ALIGN = os.path.join(BINDIR, "sibeliaz-align")
SOURCES = ("lcb_device.cu", "lcb_host.cpp", "graph_device.cu", "graph_host.cpp", "poa_device.cu")
HEADERS = ("lcb_traverse.cuh", "lcb_lean.cuh", "device_prims.cuh", "graph_kmer.cuh", "host_common.h", "graph_internal.h", "poa_core.cuh", "lcb_internal.h", "cli_common.h")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-O3,-pthread"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _host_cxx():
    # the image exports CXX=/opt/gcc/bin/g++ (a wrapper); the system compiler is the safe choice
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")


def _nccl():
    """(include_dir, lib_path) of a usable NCCL: the one bundled with torch, else the system one."""
    cands = []
    try:
        import nvidia.nccl as n  # wheel layout used by torch
        base = os.path.dirname(n.__file__) if getattr(n, "__file__", None) else list(n.__path__)[0]
        cands.append((os.path.join(base, "include"), os.path.join(base, "lib", "libnccl.so.2")))
    except Exception:
        pass
    cands.append(("/usr/include", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"))
    for inc, lib in cands:
        if os.path.exists(os.path.join(inc, "nccl.h")) and os.path.exists(lib):
            return inc, lib
    return None


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build_selfcheck(verbose=False, defines=("-DLCB_CHECK_MPV",), name="libsibeliaz_lcb_check.so"):
    """Developer build: every shortcut path of the traversal also runs the general path and compares (-DLCB_CHECK_MPV);
    a mismatch makes lcb_find_blocks fail with code 90.  Load it with LCB_LIB_PATH=<returned path>.
    Also used for other experimental -D variants (`--variant name -DX=Y ...`)."""
    os.makedirs(LIBDIR, exist_ok=True)
    inc = os.path.join(ROOT, "include")
    srcs = [os.path.join(CSRC, f) for f in SOURCES]
    out = os.path.join(LIBDIR, name)
    cmd = [_nvcc()] + ARCH + NVCC_FLAGS + list(defines) + ["-ccbin", _host_cxx(), "-I", inc, "-shared", "-o", out] + srcs
    nccl = _nccl()
    if nccl:
        cmd += ["-DLCB_WITH_NCCL", "-I", nccl[0], '-DLCB_NCCL_PATH="%s"' % nccl[1], "-ldl"]
    cmd += ["-cudart", "shared"]
    o = _run(cmd)
    if verbose:
        print(o)
    return out


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(BINDIR, exist_ok=True)
    inc = os.path.join(ROOT, "include")
    srcs = [os.path.join(CSRC, f) for f in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.join(inc, "sibeliaz_lcb.h"), os.path.join(inc, "sibeliaz_graph.h"), os.path.join(inc, "sibeliaz_align.h"), __file__]
    nvcc = _nvcc()
    if force or _newer(LIB, deps):
        cmd = [nvcc] + ARCH + NVCC_FLAGS + ["-ccbin", _host_cxx(), "-I", inc, "-shared", "-o", LIB] + srcs
        nccl = _nccl()
        if nccl:
            # NCCL is dlopen'ed on first multi-GPU use: headers for the types, no link-time dependency
            cmd += ["-DLCB_WITH_NCCL", "-I", nccl[0], '-DLCB_NCCL_PATH="%s"' % nccl[1], "-ldl"]
        cmd += ["-cudart", "shared"]
        out = _run(cmd)
        if verbose:
            print(out)
    main_src = os.path.join(CSRC, "sibeliaz_lcb_main.cpp")
    if os.path.exists(main_src) and (force or _newer(CLI, [main_src, LIB])):
        _run([_host_cxx(), "-O2", "-std=c++17", "-I", inc, main_src, "-o", CLI, "-L", LIBDIR, "-lsibeliaz_lcb",
              "-Wl,-rpath,$ORIGIN/../lib"])
    # test build with tiny per-warp arenas (tests/test_gpu_parity.py::test_big_arena_rerun): same sources, -DLCB_TINY_ARENA
    if force or _newer(LIB_TINY, deps):
        build_selfcheck(verbose=verbose, defines=("-DLCB_TINY_ARENA",), name=os.path.basename(LIB_TINY))
    graph_src = os.path.join(CSRC, "twopaco_main.cpp")
    if os.path.exists(graph_src) and (force or _newer(CLI_GRAPH, [graph_src, LIB])):
        _run([_host_cxx(), "-O2", "-std=c++17", "-I", inc, graph_src, "-o", CLI_GRAPH, "-L", LIBDIR, "-lsibeliaz_lcb",
              "-Wl,-rpath,$ORIGIN/../lib"])
    align_src = os.path.join(CSRC, "sibeliaz_align_main.cpp")
    if os.path.exists(align_src) and (force or _newer(CLI_ALIGN, [align_src, LIB])):
        _run([_host_cxx(), "-O2", "-std=c++17", "-I", inc, align_src, "-o", CLI_ALIGN, "-L", LIBDIR, "-lsibeliaz_lcb",
              "-Wl,-rpath,$ORIGIN/../lib"])
    return LIB


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_selfcheck(verbose=True, defines=sys.argv[i + 2:], name="libsibeliaz_lcb_%s.so" % sys.argv[i + 1]))
    elif "--selfcheck" in sys.argv:
        print(build_selfcheck(verbose=True))
    else:
        print(build(force="--force" in sys.argv, verbose=True))
