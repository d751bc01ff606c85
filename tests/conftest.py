import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

PKG_NAME = "jsvm-code-notes_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.hookimpl(hookwrapper=True)
def pytest_runtest_makereport(item, call):
    """Keep the full text of every failure in gpurun_out/test_failures.log (the -q summary line is cut at the terminal width)."""
    outcome = yield
    rep = outcome.get_result()
    if rep.failed:
        try:
            d = os.path.join(ROOT, "gpurun_out")
            os.makedirs(d, exist_ok=True)
            with open(os.path.join(d, "test_failures.log"), "a") as f:
                f.write(f"==== {rep.nodeid} [{rep.when}]\n{rep.longreprtext}\n")
        except OSError:
            pass


@pytest.fixture(scope="session")
def pkg():
    """The product package (its directory name has a hyphen, so it is imported by string)."""
    return importlib.import_module(PKG_NAME)


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module(PKG_NAME + ".synth")


@pytest.fixture(scope="session")
def oracle():
    """C restatement of the reference algorithm (oracle/me_oracle.c) -- the checker."""
    import oracle_lib
    path = os.path.join(ROOT, "oracle", "libme_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libme_oracle.so"])
    return oracle_lib.CpuBackend(path, "jsvm_ora_")


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference compiled by oracle/Makefile (only where oracle/_ref was built)."""
    import oracle_lib
    path = os.path.join(ROOT, "oracle", "_ref", "libjsvm_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libjsvm_ref.so not built (needs /root/reference; run `make -C oracle ref`)")
    return oracle_lib.CpuBackend(path, "jsvm_ref_")
