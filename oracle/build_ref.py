"""Vendor the UNMODIFIED reference package into ``oracle/_ref/`` so that it can run on the GPU box.

TEST / BENCH INFRASTRUCTURE.  ``/root/reference`` exists only in the build container; ``oracle/_ref/`` is listed
in ``.gitignore`` (reference sources never enter the history) but NOT in ``.gpurunignore``, so the copy travels with
the repository snapshot like the built ``.so`` files do.  Run by ``__graft_entry__.build()`` whenever
``/root/reference`` is present:

    python oracle/build_ref.py

What it is used for (and nothing else): ``bench.py --impl reference`` / ``cpu_baseline`` time the reference's own
``Hamiltonian.local_energy`` on the host cores (``kind: "reference"``), and ``tests/test_dropin_gpu.py`` runs the
reference's own ``train.py`` loop on top of this package to prove the drop-in claim.  The product package never
imports it.

The reference is seven pure-Python modules (no build step): the recipe is a verbatim file copy plus a manifest with
the sha256 of every file, so a test can tell that the copy is unmodified.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src/psiformer_torch"
REF_DIR = os.path.join(HERE, "_ref")
FILES = ("__init__.py", "config.py", "psiformer.py", "logdet_matmul.py", "jastrow.py", "hamiltonian.py", "mcmc.py",
         "train.py")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "MANIFEST.json"))


def build(verbose: bool = False) -> bool:
    """Copy the reference package; returns False (and leaves an existing copy alone) when /root/reference is absent."""
    if not os.path.isdir(REF_SRC):
        return available()
    dst = os.path.join(REF_DIR, "psiformer_torch")
    os.makedirs(dst, exist_ok=True)
    manifest = {"source": REF_SRC, "files": {}}
    for f in FILES:
        shutil.copyfile(os.path.join(REF_SRC, f), os.path.join(dst, f))
        with open(os.path.join(dst, f), "rb") as fh:
            manifest["files"][f] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(REF_DIR, "MANIFEST.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    if verbose:
        print(f"vendored {len(FILES)} files of the reference into {dst}")
    return True


def verify() -> bool:
    """The vendored copy is byte-identical to what the manifest recorded."""
    if not available():
        return False
    with open(os.path.join(REF_DIR, "MANIFEST.json")) as fh:
        manifest = json.load(fh)
    for f, digest in manifest["files"].items():
        with open(os.path.join(REF_DIR, "psiformer_torch", f), "rb") as fh:
            if hashlib.sha256(fh.read()).hexdigest() != digest:
                return False
    return True


def import_reference():
    """Import the vendored reference as the top-level package ``psiformer_torch`` (wandb disabled)."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` in the build container")
    os.environ.setdefault("WANDB_MODE", "disabled")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import psiformer_torch  # noqa: F401
    return psiformer_torch


if __name__ == "__main__":
    ok = build(verbose=True)
    print("oracle/_ref", "ready" if ok and verify() else "NOT available")
