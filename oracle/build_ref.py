"""TEST / BENCH INFRASTRUCTURE ONLY — installs the UNMODIFIED reference (pure Python) into ``oracle/_ref`` so that
``bench.py --impl reference`` and the ``cpu_baseline`` leg can time the reference's OWN ``CTRTrainer`` on the GPU box, where
``/root/reference`` does not exist.  ``oracle/_ref`` is git-ignored (never committed) but travels with gpurun snapshots.

Recipe, run by ``__graft_entry__.build()`` whenever ``/root/reference`` is present (this container only):
  1. ``pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference>``;
  2. the reference's build backend (hatchling) is not in this image, so step 1 fails here; the fallback does what installing
     the wheel of a pure-Python project does: the package directory is copied as-is and a ``dist-info/METADATA`` is written
     (the reference reads its own metadata at import, ``torch_rechub/__init__.py:3-9``).
Nothing in the product (``torch-rechub_b200/``) imports from ``oracle/``.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"
TARGET = os.path.join(HERE, "_ref")


def build(verbose=True):
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "torch_rechub")):
        return os.path.isdir(os.path.join(TARGET, "torch_rechub"))  # GPU box: use what travelled
    if os.path.isdir(os.path.join(TARGET, "torch_rechub")):
        return True
    how = "pip"
    tmp = tempfile.mkdtemp(prefix="rechub_ref_src_")
    src = os.path.join(tmp, "src")
    shutil.copytree(REFERENCE_ROOT, src, ignore=shutil.ignore_patterns("docs", "tutorials", "node_modules", ".git"))
    rc = subprocess.call([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse", "--target", TARGET, src],
                         stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if rc != 0 or not os.path.isdir(os.path.join(TARGET, "torch_rechub")):
        how = "copy (pip failed: build backend hatchling not installed)"
        shutil.rmtree(TARGET, ignore_errors=True)
        os.makedirs(TARGET)
        shutil.copytree(os.path.join(REFERENCE_ROOT, "torch_rechub"), os.path.join(TARGET, "torch_rechub"), ignore=shutil.ignore_patterns("__pycache__"))
        di = os.path.join(TARGET, "torch_rechub-0.8.0.dist-info")
        os.makedirs(di)
        with open(os.path.join(di, "METADATA"), "w") as f:
            f.write("Metadata-Version: 2.1\nName: torch-rechub\nVersion: 0.8.0\nLicense: MIT\n")
        with open(os.path.join(di, "INSTALLER"), "w") as f:
            f.write("oracle/build_ref.py\n")
    shutil.rmtree(tmp, ignore_errors=True)
    if verbose:
        print("[oracle/build_ref] reference installed into %s via %s" % (TARGET, how))
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
