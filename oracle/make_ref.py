"""Recipe for ``oracle/_ref``: the UNMODIFIED reference (openai/gym 0.26.2), installed from ``/root/reference``.

Test / measurement infrastructure, not product code.  The reference is pure Python; ``/root/reference`` exists only
in the build container, so this recipe installs it -- with pip, from a scratch copy because the source tree is
read-only -- into ``oracle/_ref/`` (git-ignored, but it travels to the GPU box with the snapshot, like the built
``.so`` files).  Nothing is copied into the repository's history, and nothing under ``gym_b200/`` imports it.

    python oracle/make_ref.py            # (re)build oracle/_ref if /root/reference is present

Consumers (``tests/``, ``bench.py``'s ``cpu_baseline_python`` leg) go through ``oracle.ref_gym.import_reference()``.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"
TARGET = os.path.join(HERE, "_ref")


def build(force=False):
    """Returns the path of the installed reference, or None when it cannot be built here."""
    marker = os.path.join(TARGET, "gym", "version.py")
    if os.path.exists(marker) and not force:
        return TARGET
    if not os.path.isdir(REFERENCE):
        return TARGET if os.path.exists(marker) else None
    tmp = tempfile.mkdtemp(prefix="gymref_")
    try:
        src = os.path.join(tmp, "gym-src")
        shutil.copytree(REFERENCE, src, symlinks=True, ignore=shutil.ignore_patterns(".git"))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        subprocess.check_call(cmd)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return TARGET


if __name__ == "__main__":
    print(build(force=True))
