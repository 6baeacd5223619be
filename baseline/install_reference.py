#!/usr/bin/env python
"""Recipe that places the UNMODIFIED reference package under baseline/_ref (git-ignored; travels to the GPU box with the
gpurun snapshot) for bench.py's `--impl reference` arm and the `cpu_baseline.reference` figure.

    python baseline/install_reference.py [/root/reference]

1. `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of the reference>`.
2. The reference's setup.py does not survive a current setuptools (`install_requires=['python_version>="3.7"', ...]` is not a
   requirement specifier: "metadata-generation-failed").  The package is pure Python with no build step, so what pip would
   have placed under --target is exactly the `brancher/` directory: it is copied as is.  No file is edited.
Nothing under baseline/_ref is tracked by git; no reference source enters the repository's history.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")


def install(src="/root/reference"):
    if not os.path.isdir(os.path.join(src, "brancher")):
        return "reference tree %s not present" % src
    if os.path.isdir(os.path.join(TARGET, "brancher")):
        return "present"
    os.makedirs(TARGET, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(src, work, ignore=shutil.ignore_patterns(".git", "*.ipynb", "*.pdf"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", TARGET, work]
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if p.returncode == 0 and os.path.isdir(os.path.join(TARGET, "brancher")):
            return "pip"
        shutil.copytree(os.path.join(src, "brancher"), os.path.join(TARGET, "brancher"),
                        ignore=shutil.ignore_patterns("__pycache__"))
        return "copied (pip: metadata-generation-failed, see the module docstring)"


if __name__ == "__main__":
    print(install(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
