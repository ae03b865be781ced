"""TEST INFRASTRUCTURE ONLY -- recipe that builds ``oracle/_ref/`` from the reference where it lies under ``/root/reference``.

The reference is pure Python, so its "binary" is byte code: every module of ``/root/reference/reactranker`` is compiled with
``py_compile`` straight from its source file into ``oracle/_ref/reactranker/<same relative path>.rrc`` (CPython byte code; the neutral extension keeps file-sync filters that drop
``*.pyc`` from leaving it behind; oracle/ref_loader.py imports these files with importlib's SourcelessFileLoader).
No reference source enters this repository: ``oracle/_ref/`` is git-ignored and holds compiler OUTPUT only -- like the ``.so`` the CUDA
sources compile to it is not gpurun-ignored, so it travels to the GPU box, where ``/root/reference`` does not exist and
``bench.py --impl reference`` / ``cpu_baseline`` then time the reference's own modules (``kind: "reference"``) instead of the oracle port.

Run by ``__graft_entry__.build()`` whenever ``/root/reference`` is present (the build container); a no-op elsewhere.
"""
import os
import py_compile
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("RR_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")


def build(verbose: bool = False) -> int:
    """Compile the reference package; returns the number of modules written (0 when the reference is absent)."""
    pkg = os.path.join(SRC, "reactranker")
    if not os.path.isdir(pkg):
        return 0
    out_pkg = os.path.join(DST, "reactranker")
    if os.path.isdir(out_pkg):
        shutil.rmtree(out_pkg)
    n = 0
    for root, dirs, files in os.walk(pkg):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        rel = os.path.relpath(root, pkg)
        for f in files:
            if not f.endswith(".py"):
                continue
            dst_dir = os.path.join(out_pkg, rel) if rel != "." else out_pkg
            os.makedirs(dst_dir, exist_ok=True)
            py_compile.compile(os.path.join(root, f), cfile=os.path.join(dst_dir, f[:-3] + ".rrc"), dfile=os.path.join("reactranker", rel, f),
                               doraise=True, optimize=0)
            n += 1
    with open(os.path.join(DST, "BUILT_FROM"), "w") as fh:
        fh.write(f"byte code of {SRC}/reactranker compiled by oracle/build_ref.py with CPython {sys.version.split()[0]}; {n} modules\n")
    if verbose:
        print(f"oracle/_ref: {n} reference modules compiled from {pkg}")
    return n


if __name__ == "__main__":
    build(verbose=True)
