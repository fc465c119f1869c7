"""Import the real reference (read-only, /root/reference) when it is present.

Used only to validate the restatements in this directory and to generate tests/golden/.  The
reference tree does not exist on the GPU box; callers must handle `available() == False`.
Recipe from SURVEY.md Appendix B: PYTHONPATH root + the package dir itself (config.py does
`import configs.mini`), data/log dirs redirected to a scratch directory.
"""
import os
import sys
import tempfile

REF_ROOT = os.environ.get("CZ_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "cchess_alphazero", "environment"))


_done = False


def setup():
    global _done
    if _done:
        return
    if not available():
        raise RuntimeError("reference tree not present at " + REF_ROOT)
    sys.dont_write_bytecode = True
    scratch = tempfile.mkdtemp(prefix="cz_ref_")
    os.environ.setdefault("PROJECT_DIR", scratch)
    os.environ.setdefault("DATA_DIR", os.path.join(scratch, "data"))
    for p in (REF_ROOT, os.path.join(REF_ROOT, "cchess_alphazero")):
        if p not in sys.path:
            sys.path.insert(0, p)
    _done = True


def senv():
    setup()
    import cchess_alphazero.environment.static_env as m
    return m


def lookup_tables():
    setup()
    import cchess_alphazero.environment.lookup_tables as m
    return m


def player_module():
    setup()
    import cchess_alphazero.agent.player as m
    return m


def config(kind="mini"):
    setup()
    from cchess_alphazero.config import Config
    return Config(kind)
