"""Builds liblumilly_b200.so (CUDA kernels + C ABI + host front end) in-tree for sm_100a.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box with
the gpurun snapshot.  Flags:
  -gencode arch=compute_100a,code=sm_100a   B200 only, no PTX for other targets
  -fmad=false                               no FMA contraction in device code (numerics contract,
                                            csrc/device_path.cuh)
  -Xcompiler -ffp-contract=off              same for the host fp32 set-up code
  -lineinfo                                 so ncu's source page maps to our code
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblumilly_b200.so")
CLI = os.path.join(HERE, "bin", "lumilly")
SOURCES = ["kernels.cu", "api.cpp", "bvh_build.cpp", "bvh_build_gpu.cu", "toml_obj.cpp", "host_scene.cpp", "image_io.cpp"]
HEADERS = ["device_scene.h", "device_path.cuh", "persistent.cuh", "pool.cuh", "path_vertex.inc", "kernels.h", "common.h", "host_scene.h", "../../include/lumilly.h"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: liblumilly_b200 cannot be built (there is no CPU fallback)")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _compile(job):
    cmd, verbose = job
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)


def build_variant(name, defines):
    """Development: builds variants/lib_<name>.so with extra -D flags (A/B experiments, tools/ab.py)."""
    global LIB
    out = os.path.join(HERE, "variants", "lib_%s.so" % name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    keep = LIB
    try:
        LIB = out
        build_library(force=True, extra=["-D" + d for d in defines], obj_dir=os.path.join(HERE, "build", "variant_" + name))
    finally:
        LIB = keep
    return out


def build_library(force=False, verbose=False, extra=(), obj_dir=None):
    """Compiles every translation unit to build/*.o in parallel (the render kernel is instantiated in eight
    units, integrator x scene-has-a-BVH x scene-has-GGX, see csrc/persistent_inst.cu) and links liblumilly_b200.so."""
    from concurrent.futures import ThreadPoolExecutor
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS + ["persistent_inst.cu"]] + [os.path.abspath(__file__)]
    if not force and not _stale(LIB, deps):
        return LIB
    obj_dir = obj_dir or os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
             "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
             "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-fvisibility=default"] + list(extra)
    if verbose:
        flags.insert(0, "-Xptxas=-v")
    jobs, objs = [], []
    for src in SOURCES:
        o = os.path.join(obj_dir, os.path.splitext(src)[0] + ".o")
        objs.append(o)
        jobs.append(([_nvcc()] + flags + ["-c", os.path.join(CSRC, src), "-o", o], verbose))
    for integ in (0, 1):
        for tree in (0, 1):
            for ggx in (0, 1):
                o = os.path.join(obj_dir, "persistent_i%d_t%d_g%d.o" % (integ, tree, ggx))
                objs.append(o)
                defs = ["-DLR_INST_INTEGRATOR=%d" % integ, "-DLR_INST_TREE=%d" % tree, "-DLR_INST_GGX=%d" % ggx, "-DLR_OUTLINE_COLD"]
                if tree and integ == 1:
                    # vector / scalar quotients out of line: +4.6 % for the one-path-per-lane kernel over a BVH (it is
                    # instruction-fetch bound), -4.5 % for the pool kernel (pt over a BVH): profiles/r01_e_ab_s50.txt
                    defs.append("-DLR_DIV_OUT_OF_LINE")
                if not ggx or (tree and integ == 1 and os.environ.get("LR_BUILD_PTD_GGX_OUT")):
                    defs.append("-DLR_GGX_OUT_OF_LINE")
                jobs.append(([_nvcc()] + flags + defs + ["-c", os.path.join(CSRC, "persistent_inst.cu"), "-o", o], verbose))
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        list(ex.map(_compile, jobs))
    subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + ["-lz"], check=True)
    return LIB


def build_cli(force=False):
    src = os.path.join(CSRC, "cli_main.cpp")
    if not os.path.exists(src):
        return None
    build_library(force=force)
    if not force and not _stale(CLI, [src, LIB]):
        return CLI
    os.makedirs(os.path.dirname(CLI), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-o", CLI, src, "-L" + HERE, "-llumilly_b200",
           "-Wl,-rpath,$ORIGIN/..", "-I" + os.path.join(HERE, "..", "include")]
    subprocess.run(cmd, check=True)
    return CLI


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
    print(build_cli(force="--force" in sys.argv))
