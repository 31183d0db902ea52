"""Copies the UNMODIFIED reference tree into the git-ignored ``baseline/_ref/`` so that it travels to the GPU box with the
repository snapshot (``/root/reference`` exists only in the build container).

The reference is a script tree without setup.py / pyproject.toml, so ``pip install --target baseline/_ref /root/reference``
has nothing to build; this script is the install.  Nothing is edited: the files are byte copies,
``baseline/_ref/MANIFEST.sha256`` records their hashes.  Users: ``oracle/ref_loader.py`` (falls back to this copy when
``/root/reference`` is absent) and through it ``bench.py --impl reference`` / ``cpu_baseline`` (kind "reference") and the
live-reference tests.  Product code (``fusion_gcn_b200/``) never imports it.

    python baseline/install_ref.py [--source /root/reference]
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
KEEP = ("torch_src", "util", "datasets", "config", "requirements.txt", "README.md")


def install(source="/root/reference", dest=DEST) -> bool:
    if not os.path.isfile(os.path.join(source, "torch_src", "models", "mmargcn", "agcn.py")):
        return False
    if os.path.isdir(dest):
        shutil.rmtree(dest)
    os.makedirs(dest)
    lines = []
    for name in KEEP:
        src = os.path.join(source, name)
        if os.path.isdir(src):
            shutil.copytree(src, os.path.join(dest, name), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.isfile(src):
            shutil.copy2(src, os.path.join(dest, name))
    for root, _, files in sorted(os.walk(dest)):
        for f in sorted(files):
            path = os.path.join(root, f)
            with open(path, "rb") as fh:
                lines.append(f"{hashlib.sha256(fh.read()).hexdigest()}  {os.path.relpath(path, dest)}")
    with open(os.path.join(dest, "MANIFEST.sha256"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return True


if __name__ == "__main__":
    src = sys.argv[sys.argv.index("--source") + 1] if "--source" in sys.argv else "/root/reference"
    ok = install(src)
    print("installed" if ok else "reference tree not found at " + src, DEST)
