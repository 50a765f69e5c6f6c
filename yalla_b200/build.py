"""Build recipe of the product library (nvcc cross-compiles sm_100a without a
GPU).

    python -m yalla_b200.build

Built in-tree so that it travels with the repository snapshot:
    yalla_b200/_lib/libyalla_b200.so   include/ + yalla_b200/csrc/capi.cu

The checkers (CPU oracle, reference build, upstream test programs) are test
infrastructure and have their own recipes in oracle/build_checkers.py.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo"] + ARCH


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


def build_variant(name, defines):
    """A tuning variant of the product library (-D overrides of the kernel
    budgets) -> yalla_b200/_lib/variants/<name>.so; picked up through the
    YALLA_B200_LIB environment variable (scripts/tune_variants.sh)."""
    out = os.path.join(ROOT, "yalla_b200", "_lib", "variants", name + ".so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    _run(["nvcc"] + NVCC_FLAGS + [f"-D{d}" for d in defines] +
         ["-Xcompiler", "-fPIC", "-shared",
          "-I", os.path.join(ROOT, "include"),
          "-I", os.path.join(ROOT, "yalla_b200", "csrc"),
          "-o", out, os.path.join(ROOT, "yalla_b200", "csrc", "capi.cu")])
    return out


def build_extension_tests(force=False):
    """tests/cuda/*.cu -- GPU test programs for header-level extensions that
    have no C-ABI entry -- into tests/_bin/ (run by tests/test_extensions_gpu.py)."""
    out_dir = os.path.join(ROOT, "tests", "_bin")
    os.makedirs(out_dir, exist_ok=True)
    built = []
    source_dir = os.path.join(ROOT, "tests", "cuda")
    for name in sorted(os.listdir(source_dir)):
        if not name.endswith(".cu"):
            continue
        source = os.path.join(source_dir, name)
        out = os.path.join(out_dir, name[:-3])
        if force or not _newer(out, [os.path.join(ROOT, "include"), source]):
            _run(["nvcc"] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"),
                 "-o", out, source])
        built.append(out)
    return built


if __name__ == "__main__":
    print(build_product(force=True))
    print(build_extension_tests(force=True))
