"""Build recipes (nvcc cross-compiles sm_100a without a GPU).

    python -m yalla_b200.build            product library
    python -m yalla_b200.build all        + oracle, reference library, upstream
                                            test binaries (where /root/reference
                                            is mounted)

Everything is built in-tree so that it travels with the repository snapshot:
    yalla_b200/_lib/libyalla_b200.so   the product (include/ + csrc/capi.cu)
    oracle/_build/libyalla_oracle.so   CPU oracle               (oracle/Makefile)
    oracle/_ref/libyalla_ref.so        reference headers, sm_100a (oracle/Makefile)
    tests/_bin/<test>                  the reference's own tests/*.cu compiled
                                       UNCHANGED against include/ (drop-in proof)
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("YALLA_REFERENCE", "/root/reference")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo"] + ARCH
UPSTREAM_TESTS = ["test_dtypes", "test_solvers", "test_links", "test_polarity",
                  "test_inits", "test_vtk"]


def _run(cmd, **kw):
    result = subprocess.run(cmd, capture_output=True, text=True, **kw)
    if result.returncode != 0:
        raise RuntimeError(
            "build failed: " + " ".join(cmd) + "\n" + result.stdout[-4000:] +
            result.stderr[-4000:])
    return result


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    stamp = os.path.getmtime(target)
    for source in sources:
        if os.path.isdir(source):
            for base, _, files in os.walk(source):
                if any(os.path.getmtime(os.path.join(base, f)) > stamp
                       for f in files):
                    return False
        elif os.path.getmtime(source) > stamp:
            return False
    return True


def build_product(force=False):
    out = os.path.join(ROOT, "yalla_b200", "_lib", "libyalla_b200.so")
    sources = [os.path.join(ROOT, "include"), os.path.join(ROOT, "yalla_b200", "csrc")]
    if not force and _newer(out, sources):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    _run(["nvcc"] + NVCC_FLAGS + ["-Xcompiler", "-fPIC", "-shared",
         "-I", os.path.join(ROOT, "include"),
         "-I", os.path.join(ROOT, "yalla_b200", "csrc"),
         "-o", out, os.path.join(ROOT, "yalla_b200", "csrc", "capi.cu")])
    return out


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


def build_all():
    with ThreadPoolExecutor(max_workers=4) as pool:
        jobs = [pool.submit(build_product), pool.submit(build_oracle),
                pool.submit(build_reference), pool.submit(build_upstream_tests)]
        return [job.result() for job in jobs]


if __name__ == "__main__":
    print(build_all() if "all" in sys.argv[1:] else build_product(force=True))
