"""Build libg3d_b200.so (and the dfgen CLI) in-tree with nvcc for sm_100a.

    python -m generate3d_b200.build [--force]

Cross-compiles without a GPU. The faithful distance-field translation unit is
compiled with -fmad=false on top of its explicit round-to-nearest intrinsics so
that no float multiply-add is ever contracted (the CPU reference has no FMA).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libg3d_b200.so")
CLI = os.path.join(HERE, "bin", "dfgen")
VOX2MESH = os.path.join(HERE, "bin", "vox2mesh")
CLIS = {CLI: "dfgen_main.cpp", VOX2MESH: "vox2mesh_main.cpp"}
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function"]
UNITS = {
    "g3d_tdf.cu": ["-fmad=false"],
    "g3d_raycast.cu": ["-fmad=false"],
    "g3d_next.cu": ["-fmad=false"],
    "g3d_capi.cu": [],
    "g3d_mesh_io.cpp": ["-Xcompiler", "-ffp-contract=off"],      # replays tinyobj's double / float arithmetic
}
HEADERS = ["g3d_internal.h", "g3d_ptd.cuh", "g3d_tdf_exact.cuh", os.path.join("..", "..", "include", "g3d.h")]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "nvcc")
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    objs, cmds = [], []
    for src, extra in UNITS.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        if force or _stale(o, [s] + hdrs):
            cmds.append([nvcc] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])
        objs.append(o)
    if cmds:                                                   # the translation units are independent: compile side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(len(cmds)) as ex:
            list(ex.map(subprocess.check_call, cmds))
    if force or _stale(LIB, objs):
        subprocess.check_call([nvcc] + ARCH + ["-shared", "-o", LIB] + objs)
    for exe, src_name in CLIS.items():                         # the reference's two executables, same argv
        cli_src = os.path.join(CSRC, src_name)
        if os.path.exists(cli_src) and (force or _stale(exe, [cli_src, LIB] + hdrs)):
            os.makedirs(os.path.dirname(exe), exist_ok=True)
            subprocess.check_call([nvcc] + ARCH + ["-O2", "-std=c++17", "-o", exe, cli_src, "-L" + HERE, "-lg3d_b200",
                                                   "-Xcompiler", "-pthread", "-Xlinker", "-rpath,$ORIGIN/.."])
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
