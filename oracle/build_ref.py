#!/usr/bin/env python
"""Recipe for oracle/_ref/: the reference's OWN model classes, taken from where they lie under /root/reference.

TEST / BENCH INFRASTRUCTURE ONLY.  The reference implementation of the hot path is three small Python files per
scale (CODON_x4.py / CODON_x16.py, CAC_module.py, attention/ResCBAM.py).  `build()` packs exactly those files, unmodified,
into one importable archive per scale, oracle/_ref/CODON_X{4,8,16}.zip (Python imports straight from zip archives) -- a
build output like the compiled .so: git-ignored, never committed, no reference source file lands in the tree, but the
archives ship to the GPU box with the repository snapshot -- so that `bench.py --impl reference` and the `cpu_baseline`
leg can time the UNMODIFIED reference classes on the box's host cores (kind "reference") instead of the functional
restatement in oracle/codon_oracle.py (kind "port", the fallback when oracle/_ref is absent).  Nothing under codon_b200/
imports it.

  python oracle/build_ref.py          # needs /root/reference (or $CODON_REFERENCE); run by __graft_entry__.build()
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_OUT = os.path.join(HERE, "_ref")
FILES = {
    "CODON_X4": ["CODON_x4.py", "CAC_module.py", os.path.join("attention", "ResCBAM.py")],
    "CODON_X8": ["CODON_x8.py", "CAC_module.py", os.path.join("attention", "ResCBAM.py")],
    "CODON_X16": ["CODON_x16.py", "CAC_module.py"],
}


def _archive(sub: str) -> str:
    return os.path.join(REF_OUT, sub + ".zip")


def build(reference_root=None) -> bool:
    """Packs the reference's model files into oracle/_ref/*.zip.  Returns False (and leaves any earlier archives alone)
    when the reference checkout is not present -- e.g. on the GPU box, which only uses the prebuilt archives."""
    import zipfile
    root = reference_root or os.environ.get("CODON_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(root, "CODON_X4")):
        return False
    os.makedirs(REF_OUT, exist_ok=True)
    for sub, files in FILES.items():
        tmp = _archive(sub) + ".tmp"
        with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
            for f in files:
                z.write(os.path.join(root, sub, f), arcname=f.replace(os.sep, "/"))
            if any(os.sep in f or "/" in f for f in files):
                z.writestr("attention/__init__.py", "")        # (the reference relies on an implicit namespace package)
        os.replace(tmp, _archive(sub))
    return True


def available() -> bool:
    return all(os.path.exists(_archive(sub)) for sub in FILES)


def load_model_class(scale: int):
    """The reference's `CODONNet` class for a scale, imported from its archive under oracle/_ref (None if that was never
    built).  The three archives reuse module names (CAC_module, attention), so they are purged between imports."""
    if not available():
        return None
    for m in ("CAC_module", "attention", "attention.ResCBAM", "CODON_x4", "CODON_x8", "CODON_x16"):
        sys.modules.pop(m, None)
    d = _archive(f"CODON_X{scale}")
    sys.path.insert(0, d)
    try:
        importlib.invalidate_caches()
        return importlib.import_module(f"CODON_x{scale}").CODONNet
    finally:
        sys.path.remove(d)


if __name__ == "__main__":
    ok = build()
    print("oracle/_ref built from the reference checkout" if ok else "reference checkout not found; oracle/_ref unchanged",
          "| available:", available())
