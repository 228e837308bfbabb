"""Install the UNMODIFIED reference into baseline/_ref/ (git-ignored; it ships to the GPU box with the snapshot).

    python baseline/install_ref.py            # build container only: needs /root/reference

First choice is the install the task statement prescribes,
    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref /root/reference
which fails in this image: the reference's build backend (hatchling, pyproject.toml:22-24) is in neither the
environment nor /opt/wheelhouse.  The fallback reproduces what that wheel would contain —
``[tool.hatch.build.targets.wheel] packages = ["./mmidas"]``: the ``mmidas`` package directory, .py files only — by
copying it file for file.  Nothing is edited; nothing under baseline/_ref/ is tracked by git, imported by the product
or read by the tests.  Users: ``bench.py --impl reference`` (the reference's own CPU path, timed) and bench.py's
``reference_gpu`` object (the same unmodified classes on the B200 — the "reference GPU path" of the north star).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = "/root/reference"


def install() -> str:
    if not os.path.isdir(SRC):
        return "present" if os.path.isdir(os.path.join(DST, "mmidas")) else "unavailable: /root/reference is not on this machine"
    rc = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--find-links",
                         "/opt/wheelhouse", "--target", DST, SRC], capture_output=True, text=True)
    if rc.returncode == 0:
        return "pip"
    shutil.rmtree(os.path.join(DST, "mmidas"), ignore_errors=True)
    shutil.copytree(os.path.join(SRC, "mmidas"), os.path.join(DST, "mmidas"),
                    ignore=shutil.ignore_patterns("*.ipynb", "__pycache__", "*.pyc"))
    return "copied (pip: " + (rc.stderr.strip().splitlines() or ["failed"])[-1][:120] + ")"


if __name__ == "__main__":
    print("baseline/_ref:", install())
