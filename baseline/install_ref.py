"""Install the UNMODIFIED upstream package (Danderson123/Amira, /root/reference) into baseline/_ref/.

baseline/_ref/ is git-ignored but travels to the GPU box with the repository snapshot, so that
`bench.py --impl reference` and the `cpu_baseline` leg can time upstream's own Python
`GeneMerGraph(readDict, k)` on the box's host cores.

Upstream builds with poetry-core, which is not in this image (`pip install /root/reference` dies with
"No module named 'poetry'").  So the sources are copied to a scratch directory, the build-system table
of pyproject.toml is swapped for setuptools there (packaging metadata only -- no source file is touched)
and pip installs from that copy:

    python baseline/install_ref.py

Nothing under amira_b200/ imports baseline/_ref; tests/ use it only through oracle/ref_harness.py."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("AMIRA_REFERENCE", "/root/reference")

SHIM = """[build-system]
requires = ["setuptools"]
build-backend = "setuptools.build_meta"

[project]
name = "amira-amr"
version = "0.11.0"

[tool.setuptools]
packages = ["amira"]
include-package-data = true

[tool.setuptools.package-data]
amira = ["assets/*", "assets/**/*"]
"""


def install(force: bool = False) -> str | None:
    if not os.path.isdir(os.path.join(SOURCE, "amira")):
        return TARGET if os.path.isdir(os.path.join(TARGET, "amira")) else None
    if os.path.isdir(os.path.join(TARGET, "amira")) and not force:
        return TARGET
    tmp = tempfile.mkdtemp(prefix="amira_ref_")
    try:
        src = os.path.join(tmp, "src")
        shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns(".git", "tests", "*.pyc", "__pycache__"))
        with open(os.path.join(src, "pyproject.toml"), "w") as f:
            f.write(SHIM)
        shutil.rmtree(TARGET, ignore_errors=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout[-2000:] + res.stderr[-2000:])
            raise RuntimeError("pip install of the reference failed")
        # the one upstream fixture BASELINE.json names (configs[0]); data, not source
        fx = os.path.join(SOURCE, "tests", "complex_gene_calls_one.json")
        if os.path.exists(fx):
            os.makedirs(os.path.join(TARGET, "fixtures"), exist_ok=True)
            shutil.copy(fx, os.path.join(TARGET, "fixtures", "complex_gene_calls_one.json"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return TARGET


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
