"""In-tree build of libacro_b200.so (nvcc, sm_100a only).

    python -m acrobotics_b200.build [--force] [--verbose]

The shared library is written next to this file so that it travels with the
repository snapshot to the GPU box; it is git-ignored (history stays source
only).  Translation units are compiled in parallel and relinked only when a
source or header is newer than the library.
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
OBJ_DIR = os.path.join(HERE, "_obj")
LIB_PATH = os.path.join(HERE, "libacro_b200.so")

SOURCES = ["abi.cu", "fk.cu", "fk_cc.cu", "fk_cc_wave.cu", "fk_cc_filt.cu", "jac_ik.cu", "path.cu", "swept.cu", "planner.cu", "dp.cu"]
NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-std=c++17",
    "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler",
    "-fPIC",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libacro_b200.so")
    return exe


def _deps():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return files


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(f) > t for f in _deps())


def _compile(src, verbose):
    obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
    cmd = [_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{p.stdout}\n{p.stderr}")
    return obj, p.stderr


def build_variant(tag, defines):
    """Tuning aid: build libacro_b200_<tag>.so with extra -D defines (selected at run
    time with ACRO_B200_LIB=<path>); objects go to _obj/<tag>/."""
    global OBJ_DIR
    saved_obj, saved_flags = OBJ_DIR, list(NVCC_FLAGS)
    OBJ_DIR = os.path.join(saved_obj, tag)
    NVCC_FLAGS.extend(f"-D{d}" for d in defines)
    try:
        os.makedirs(OBJ_DIR, exist_ok=True)
        with cf.ThreadPoolExecutor(max_workers=8) as ex:
            objs = [o for o, _ in ex.map(lambda s: _compile(s, False), SOURCES)]
        out = os.path.join(HERE, f"libacro_b200_{tag}.so")
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *objs,
               "-lcudart"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
        return out
    finally:
        OBJ_DIR = saved_obj
        NVCC_FLAGS[:] = saved_flags


def build(force=False, verbose=False):
    """Compile every CUDA translation unit for sm_100a and link the C-ABI library."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    objs = [o for o, _ in results]
    cmd = [
        _nvcc(),
        "-shared",
        "-gencode",
        "arch=compute_100a,code=sm_100a",
        "-o",
        LIB_PATH,
        *objs,
        "-lcudart",
    ]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    if "--variant" in sys.argv:
        k = sys.argv.index("--variant")
        print(build_variant(sys.argv[k + 1], sys.argv[k + 2:]))
        sys.exit(0)
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
