"""In-tree build of the native libraries (no JIT cache: the built .so files travel with the repo
snapshot to the GPU box).

  libshimmer_gpu.so  : CUDA kernels + C ABI, sm_100a only, -fmad=false (arithmetic contract)
  libshimmer_host.so : host-side BVH build (stands in for shimmer's Rust BvhAggregate::new)
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "--shared", "-Xcompiler", "-fPIC", "--compress-mode=size"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


# translation units of libshimmer_gpu.so: the C ABI + traversal kernels, and ten groups of shade-kernel instantiations
# (csrc/sg_kernels.h).  They are compiled in parallel (one nvcc process each) and linked into one shared library.
GPU_UNITS = [("shimmer_gpu", "shimmer_gpu.cu", [])] + [("shade_tu%d" % i, "shade_tu.cu", ["-DSG_TU=%d" % i]) for i in range(1, 11)]


# units an A/B variant re-compiles (the C ABI + traversal unit, the lean shade kernels, the staged-shading unit); the rest is
# linked from the default build's objects
VARIANT_UNITS = ("shimmer_gpu", "shade_tu1", "shade_tu6", "shade_tu10")


def build_gpu(force=False, verbose=False, variant=None, defs=()):
    """variant / defs: an A/B build with extra -D flags -> ab/libshimmer_gpu_<variant>.so (loaded through SHIMMER_GPU_LIB)"""
    out = os.path.join(HERE, "libshimmer_gpu.so")
    if variant:
        os.makedirs(os.path.join(HERE, "ab"), exist_ok=True)
        out = os.path.join(HERE, "ab", "libshimmer_gpu_%s.so" % variant)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    srcs.append(os.path.join(HERE, "..", "include", "shimmer_gpu.h"))
    if not force and not _stale(out, srcs):
        return out
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    objdir = os.path.join(HERE, "build", variant) if variant else os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "--shared"] + (["-Xptxas", "-v"] if verbose else []) + list(defs)
    procs = []
    reused = []
    for name, src, udefs in GPU_UNITS:
        obj = os.path.join(objdir, name + ".o")
        if variant and name not in VARIANT_UNITS:       # A/B builds only re-compile the units their macros reach
            reused.append(os.path.join(HERE, "build", name + ".o"))
            continue
        procs.append((name, obj, subprocess.Popen([nvcc] + flags + udefs + ["-c", "-o", obj, os.path.join(CSRC, src)])))
    failed = [name for name, _, p in procs if p.wait() != 0]
    if failed:
        raise RuntimeError("nvcc failed for: " + ", ".join(failed))
    # NCCL is dlopen'ed at run time (csrc/shimmer_gpu.cu: nccl_load), so the library links against nothing but cudart and libdl
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", out] + [obj for _, obj, _ in procs] + reused + ["-ldl"], check=True)
    return out


def build_host(force=False):
    out = os.path.join(HERE, "libshimmer_host.so")
    srcs = [os.path.join(CSRC, "host_bvh.cpp"), os.path.join(CSRC, "sg_host_tables.h"), os.path.join(HERE, "..", "include", "shimmer_gpu.h")]
    if not force and not _stale(out, srcs):
        return out
    cxx = shutil.which("g++") or "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", out, srcs[0]], check=True)
    return out


def build_all(force=False, verbose=False):
    return build_host(force), build_gpu(force, verbose)


if __name__ == "__main__":
    import sys
    if "--variant" in sys.argv:          # python -m shimmer_b200.build --variant t512 -DSG_SHADE_THREADS=512
        name = sys.argv[sys.argv.index("--variant") + 1]
        print(build_gpu(force=True, verbose="-v" in sys.argv, variant=name, defs=[a for a in sys.argv if a.startswith("-D")]))
    else:
        print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
