import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


_graph_cache: dict = {}


@pytest.fixture(scope="session")
def case_inputs():
    """Build (and cache) the seeded workload + oracle graph of a named parity case."""
    import cases
    import oracle_lib

    def get(name: str):
        if name not in _graph_cache:
            c = cases.CASE_BY_NAME[name]
            w = cases.build(c)
            g = oracle_lib.OracleGraph.from_reads(w.k, w.solid, w.reads)
            _graph_cache[name] = (c, w, g)
        return _graph_cache[name]

    return get
