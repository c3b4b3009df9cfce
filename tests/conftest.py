import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "zk-research-implementations_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import c_oracle

    c_oracle.build()
    return c_oracle


def _stale(so: str) -> bool:
    """True if the prebuilt library was not built from the sources in the tree (the Makefile bakes csrc/src_hash.py's
    hash into zkb_version())."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("_zkb_src_hash", os.path.join(ROOT, PKG, "csrc", "src_hash.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    with open(so, "rb") as f:
        return ("src:" + mod.src_hash()).encode() not in f.read()


@pytest.fixture(scope="session")
def zkb():
    """The product package (hyphenated directory name -> importlib).  The shared library is a build artefact
    (git-ignored): compile it first if this is a fresh checkout.  There is still no fallback: if the build
    fails the tests fail."""
    so = os.path.join(ROOT, PKG, "libzkb200.so")
    if not os.path.exists(so) or _stale(so):
        import __graft_entry__

        __graft_entry__.build()
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def ctxs(zkb):
    """One device context per (field, mode), created lazily on cuda:0.  Fails loudly without a GPU."""
    cache = {}

    def get(field=0, mode=0):
        key = (field, mode)
        if key not in cache:
            cache[key] = zkb.Context(field, 0, mode)
        return cache[key]

    yield get
    for c in cache.values():
        c.close()
