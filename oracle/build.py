"""Builds the oracle's native pieces.  TEST INFRASTRUCTURE.

* ``oracle/_build/liboracle.so``      <- oracle/c/uyd_oracle.c (our C restatement)
* ``oracle/_ref/libref_postprocess.so`` <- the reference's own postprocess.hpp compiled
  *where it lies* under /root/reference (only when that mount exists; the GPU box uses
  the prebuilt file that travels with the snapshot).
* ``oracle/_ref/libref_preprocess.so``  <- the reference's own cuda_preprocess.cu compiled with nvcc for
  sm_100a where it lies (its kernels are the parity oracle of the pre-processing row; they run on the GPU
  box only, from the prebuilt file).
"""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_INC = Path("/root/reference/unina_yolo_dla/ros2_ws/src/perception/include")
ORACLE_SO = HERE / "_build" / "liboracle.so"
REF_SO = HERE / "_ref" / "libref_postprocess.so"


def _stale(out: Path, srcs) -> bool:
    return (not out.exists()) or any(Path(s).stat().st_mtime > out.stat().st_mtime for s in srcs)


def build_oracle(force: bool = False) -> Path:
    src = HERE / "c" / "uyd_oracle.c"
    if force or _stale(ORACLE_SO, [src]):
        ORACLE_SO.parent.mkdir(exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c11", str(src), "-o", str(ORACLE_SO), "-lm"]
        )
    return ORACLE_SO


def build_ref(force: bool = False):
    """Returns the path of the compiled reference header shim, or None if unavailable."""
    src = HERE / "c" / "ref_postprocess_harness.cpp"
    if REF_INC.exists() and (force or _stale(REF_SO, [src, REF_INC / "postprocess.hpp"])):
        REF_SO.parent.mkdir(exist_ok=True)
        subprocess.check_call(
            ["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", f"-I{REF_INC}", str(src), "-o", str(REF_SO)]
        )
    return REF_SO if REF_SO.exists() else None


REF_PRE_SRC = Path("/root/reference/unina_yolo_dla/ros2_ws/src/perception/src/cuda_preprocess.cu")
REF_PRE_SO = HERE / "_ref" / "libref_preprocess.so"


def build_ref_preprocess(force: bool = False):
    """The reference's CUDA pre-processing kernels as a shared library (nvcc, sm_100a), or None."""
    if REF_PRE_SRC.exists() and (force or _stale(REF_PRE_SO, [REF_PRE_SRC])):
        REF_PRE_SO.parent.mkdir(exist_ok=True)
        subprocess.check_call(
            ["/usr/local/cuda/bin/nvcc", "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-O2",
             f"-I{REF_INC}", str(REF_PRE_SRC), "-o", str(REF_PRE_SO), "-cudart", "static"],
            stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return REF_PRE_SO if REF_PRE_SO.exists() else None


COMPAT_TEST_SRC = HERE.parent / "tests" / "compat" / "node_sequence.cpp"
COMPAT_TEST_BIN = HERE / "_ref" / "compat_node_test"


def build_compat_test(force: bool = False):
    """tests/compat/node_sequence.cpp compiled against the REFERENCE's headers (where they lie) and linked with
    unina-yolo-dla_b200/libuyd_compat.so; a prebuilt binary on the GPU box.  None when the mount is missing and
    no prebuilt binary exists."""
    pkg = HERE.parent / "unina-yolo-dla_b200"
    lib = pkg / "libuyd_compat.so"
    if REF_INC.exists() and lib.exists() and (force or _stale(COMPAT_TEST_BIN, [COMPAT_TEST_SRC, lib, HERE.parent / "include" / "uyd_compat.h"])):
        COMPAT_TEST_BIN.parent.mkdir(exist_ok=True)
        subprocess.check_call(
            ["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-Xcompiler", "-ffp-contract=off", f"-I{REF_INC}", f"-I{HERE.parent / 'include'}",
             str(COMPAT_TEST_SRC), "-o", str(COMPAT_TEST_BIN), f"-L{pkg}", "-luyd_compat", "-luyd", "-ldl",
             "-Xlinker", "-rpath=$ORIGIN/../../unina-yolo-dla_b200", "-cudart", "static"])
    return COMPAT_TEST_BIN if COMPAT_TEST_BIN.exists() else None


if __name__ == "__main__":
    print(build_ref_preprocess(True))
    print(build_oracle(True))
    print(build_ref(True))
    print(build_compat_test(True))
