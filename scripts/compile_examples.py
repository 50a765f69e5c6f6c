"""Compile the reference's examples/*.cu UNCHANGED against this repo's include/
(drop-in check of the header API). Needs /root/reference; nothing is copied:
the sources are symlinked into a staging tree whose include/ is ours.

    python scripts/compile_examples.py [name ...]
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("YALLA_REFERENCE", "/root/reference")


def stage():
    path = os.path.join(ROOT, "build", "examples_stage")
    shutil.rmtree(path, ignore_errors=True)
    os.makedirs(os.path.join(path, "examples"))
    os.makedirs(os.path.join(path, "bin"))
    os.symlink(os.path.join(ROOT, "include"), os.path.join(path, "include"))
    for name in os.listdir(os.path.join(REFERENCE, "examples")):
        os.symlink(os.path.join(REFERENCE, "examples", name),
                   os.path.join(path, "examples", name))
    return path


def compile_one(path, name):
    result = subprocess.run(
        ["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a",
         "-o", f"bin/{name}", f"examples/{name}.cu"], cwd=path,
        capture_output=True, text=True)
    errors = [l for l in (result.stdout + result.stderr).splitlines() if "error" in l]
    return name, result.returncode == 0, errors[:2]


def compile_examples(names=None, workers=8):
    path = stage()
    if not names:
        names = sorted(f[:-3] for f in os.listdir(os.path.join(REFERENCE, "examples"))
                       if f.endswith(".cu"))
    with ThreadPoolExecutor(max_workers=workers) as pool:
        return list(pool.map(lambda n: compile_one(path, n), names))


if __name__ == "__main__":
    results = compile_examples(sys.argv[1:])
    for name, ok, errors in results:
        print(f"{'ok  ' if ok else 'FAIL'} {name}" + ("" if ok else f"  {errors}"))
    print(f"{sum(ok for _, ok, _ in results)} of {len(results)} examples compile")
