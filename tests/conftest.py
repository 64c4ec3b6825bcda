import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def lib_built():
    from hephaestus_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.lib()


@pytest.fixture(scope="session")
def ctx(lib_built):
    from hephaestus_b200 import operators
    return operators.Context(0)


def pytest_terminal_summary(terminalreporter):
    """One line per first-run group (tests/test_zz_first_run_gpu.py) whatever the verbosity: XPASS = the code that had
    never executed on a device is parity-green on this GPU, XFAIL = not yet (with the tail of the child's output)."""
    rows = []
    for key, label in (("xpassed", "XPASS"), ("xfailed", "XFAIL")):
        for rep in terminalreporter.stats.get(key, []):
            if "test_zz_first_run_gpu" in rep.nodeid:
                rows.append((rep.nodeid.split("::")[-1], label, rep))
    if not rows:
        return
    terminalreporter.section("first GPU runs of host-verified code")
    for name, label, rep in rows:
        terminalreporter.write_line("%-5s %s" % (label, name))
        if label == "XFAIL":
            tail = str(getattr(rep, "longrepr", ""))[-600:].strip().splitlines()[-6:]
            for ln in tail:
                terminalreporter.write_line("        " + ln)
