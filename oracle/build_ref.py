"""Build recipe for oracle/_ref: the reference itself, compiled where it lies.  TEST / BENCH INFRASTRUCTURE.

The reference is pure Python, so "compiling its own source files" means byte-compiling them: every module of
/root/reference/cchess_alphazero that the self-play path imports is compiled with py_compile straight from the read-only
tree into oracle/_ref/cchess_alphazero/**/<module>.pyc (sourceless-import layout).  No reference SOURCE is copied into
this repository: oracle/_ref/ holds only compiler output, is listed in .gitignore (so it stays out of history) and not in
.gpurunignore (so it travels to the GPU box like the built .so files).  `__graft_entry__.build()` runs this whenever
/root/reference is present; the GPU box only ever uses the prebuilt files.

What uses it: bench.py's CPU arm (`--impl reference` and the `cpu_baseline` leg) — the UNMODIFIED manager -> worker/
self_play.start -> SelfPlayWorker -> CChessPlayer <-> Pipe <-> CChessModelAPI.predict_batch_worker plumbing
(oracle/ref_selfplay_bench.py).  Nothing in the product package may import it.

    python -m oracle.build_ref
"""
import hashlib
import json
import os
import py_compile
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("CZ_REFERENCE_ROOT", "/root/reference")
PKG = "cchess_alphazero"
# sub-packages on (or imported by) the self-play / evaluation path; GUI, play_games and colaboratory are not needed
SKIP_DIRS = {"play_games", "__pycache__"}


def have_sources():
    return os.path.isdir(os.path.join(REF_ROOT, PKG, "worker"))


def have_build():
    return os.path.exists(os.path.join(OUT, "MANIFEST.json")) and os.path.exists(os.path.join(OUT, PKG, "worker", "self_play.pyc"))


def _tree_digest(files):
    h = hashlib.sha256()
    for rel, src in files:
        h.update(rel.encode())
        with open(src, "rb") as f:
            h.update(f.read())
    h.update(sys.version.encode())
    return h.hexdigest()


def build_ref(force=False):
    """Returns the output directory, or None when neither the sources nor a previous build exist."""
    if not have_sources():
        return OUT if have_build() else None
    files = []
    top = os.path.join(REF_ROOT, PKG)
    for dirpath, dirnames, filenames in os.walk(top):
        dirnames[:] = sorted(d for d in dirnames if d not in SKIP_DIRS)
        for fn in sorted(filenames):
            if fn.endswith(".py"):
                src = os.path.join(dirpath, fn)
                files.append((os.path.relpath(src, REF_ROOT), src))
    digest = _tree_digest(files)
    man_path = os.path.join(OUT, "MANIFEST.json")
    if not force and os.path.exists(man_path):
        try:
            with open(man_path) as f:
                if json.load(f).get("digest") == digest and have_build():
                    return OUT
        except Exception:
            pass
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    for rel, src in files:
        dst = os.path.join(OUT, rel[:-3] + ".pyc")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile=rel, doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(man_path, "w") as f:
        json.dump({"what": "byte-compiled modules of the reference (compiler output only, no sources)", "reference_root": REF_ROOT,
                   "python": sys.version, "modules": [rel for rel, _ in files], "digest": digest}, f, indent=1)
    return OUT


if __name__ == "__main__":
    out = build_ref(force="--force" in sys.argv)
    print(out if out else "reference sources not present and no previous build")
