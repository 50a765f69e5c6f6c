"""Build recipes of the CHECKERS (test infrastructure; see oracle/Makefile).

    python oracle/build_checkers.py

    oracle/_build/libyalla_oracle.so   CPU oracle (g++, OpenMP)
    oracle/_ref/libyalla_ref.so        the unmodified reference headers compiled
                                       for sm_100a through yalla_b200/csrc/capi.cu
    tests/_bin/<test>                  the reference's own tests/*.cu compiled
                                       UNCHANGED against this repo's include/
The last two need /root/reference (read in place, never copied); on the GPU box
the prebuilt files that travelled with the snapshot are used.
"""
import os
import shutil
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from yalla_b200.build import NVCC_FLAGS, _newer, _run  # noqa: E402

REFERENCE = os.environ.get("YALLA_REFERENCE", "/root/reference")
UPSTREAM_TESTS = ["test_dtypes", "test_solvers", "test_links", "test_polarity",
                  "test_inits", "test_vtk", "test_mesh"]


def build_oracle():
    _run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return os.path.join(ROOT, "oracle", "_build", "libyalla_oracle.so")


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE, "include"))


def build_reference():
    """The reference's own headers -> oracle/_ref (only where it is mounted)."""
    if not have_reference():
        return None
    _run(["make", "-C", os.path.join(ROOT, "oracle"), "ref",
          f"REFERENCE={REFERENCE}"])
    return os.path.join(ROOT, "oracle", "_ref", "libyalla_ref.so")


def build_upstream_tests():
    """Compile the reference's tests/*.cu, unmodified, against include/.

    The sources include "../include/x.cuh" relative to themselves, so they are
    symlinked into a staging tree whose include/ is this repo's. Nothing is
    copied into the repository; only the binaries land in tests/_bin/.
    """
    if not have_reference():
        return []
    stage = os.path.join(ROOT, "build", "upstream_stage")
    shutil.rmtree(stage, ignore_errors=True)
    os.makedirs(os.path.join(stage, "tests"))
    os.symlink(os.path.join(ROOT, "include"), os.path.join(stage, "include"))
    for name in os.listdir(os.path.join(REFERENCE, "tests")):
        os.symlink(os.path.join(REFERENCE, "tests", name),
                   os.path.join(stage, "tests", name))
    out_dir = os.path.join(ROOT, "tests", "_bin")
    os.makedirs(out_dir, exist_ok=True)

    def compile_one(test):
        out = os.path.join(out_dir, test)
        if _newer(out, [os.path.join(ROOT, "include")]):
            return out
        _run(["nvcc"] + NVCC_FLAGS + ["-o", out, f"tests/{test}.cu"], cwd=stage)
        return out

    with ThreadPoolExecutor(max_workers=6) as pool:
        return list(pool.map(compile_one, UPSTREAM_TESTS))


def build_grid_ab():
    """scripts/grid_build_ab.cu against both header sets (the sort A/B)."""
    source = os.path.join(ROOT, "scripts", "grid_build_ab.cu")
    out_dir = os.path.join(ROOT, "tests", "_bin")
    os.makedirs(out_dir, exist_ok=True)
    built = []
    variants = [("grid_ab_product", os.path.join(ROOT, "include"))]
    if have_reference():
        variants.append(("grid_ab_reference", os.path.join(REFERENCE, "include")))
    for name, include in variants:
        out = os.path.join(out_dir, name)
        if not _newer(out, [include, source]):
            _run(["nvcc"] + NVCC_FLAGS + ["-I", include, "-o", out, source])
        built.append(out)
    return built


def build_checkers():
    with ThreadPoolExecutor(max_workers=4) as pool:
        jobs = [pool.submit(build_oracle), pool.submit(build_reference),
                pool.submit(build_upstream_tests), pool.submit(build_grid_ab)]
        return [job.result() for job in jobs]


if __name__ == "__main__":
    print(build_checkers())
