import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "pde-odesolver_b200", "python"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle_lib as ol
    ol.build_oracle()
    return ol


@pytest.fixture(scope="session")
def pb():
    """The python binding over the C ABI, with the CUDA library built."""
    import __graft_entry__ as ge
    import oracle_lib as ol
    ge.build_cuda()
    ol.build_oracle()
    import pdeode_b200
    return pdeode_b200


ENGINE_KEYS = ("PDEODE_PERSIST", "PDEODE_VARIANT", "PDEODE_DEFER_X", "PDEODE_WHOLE_STEP")


@pytest.fixture
def engine_env(monkeypatch):
    """Selects the CG engine of contexts created afterwards (read at pdeode_create)."""
    def select(engine, sum_order=None):
        for k in ENGINE_KEYS:
            monkeypatch.delenv(k, raising=False)
        # small grids default to the whole step in one launch; the other engines are asked for
        monkeypatch.setenv("PDEODE_WHOLE_STEP", "1" if engine == "whole-step" else "0")
        if engine == "kernels":                      # kernel per phase, 88-byte iteration
            monkeypatch.setenv("PDEODE_PERSIST", "0")
        elif engine == "kernels-96":                 # kernel per phase, x updated by K_B
            monkeypatch.setenv("PDEODE_PERSIST", "0")
            monkeypatch.setenv("PDEODE_DEFER_X", "0")
        elif engine == "persistent":
            monkeypatch.setenv("PDEODE_PERSIST", "1")
        elif engine == "register-staged":
            monkeypatch.setenv("PDEODE_PERSIST", "0")
            monkeypatch.setenv("PDEODE_VARIANT", "reg")
        if sum_order:
            monkeypatch.setenv("PDEODE_SUM_ORDER", sum_order)
        else:
            monkeypatch.delenv("PDEODE_SUM_ORDER", raising=False)
    return select
