"""Host-side pieces of bench.py that need no GPU: the clock sampler's parsing and windowing
(against a stand-in nvidia-smi), the slab form of the synthetic initial condition, the thread
count of the CPU arm, and that both arms print the same `config`."""
import argparse
import importlib.util
import os
import stat
import sys
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


FAKE = """#!/usr/bin/env python3
import sys, time, datetime
loop = int(sys.argv[sys.argv.index("-lms") + 1]) if "-lms" in sys.argv else None
def line():
    t = datetime.datetime.now()
    return (t.strftime("%Y/%m/%d %H:%M:%S.") + "%03d" % (t.microsecond // 1000) +
            ", 0, 1965, 1980, 512.33, 0x0000000000000004, Not Active, Not Active, Not Active, Active")
time.sleep(0.3)            # the start-up a real sampler spends attaching to the GPUs
while True:
    print(line(), flush=True)
    if loop is None:
        break
    time.sleep(loop / 1000)
"""


@pytest.fixture
def fake_smi(tmp_path, monkeypatch):
    exe = tmp_path / "nvidia-smi"
    exe.write_text(FAKE)
    exe.chmod(exe.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])


def test_sampler_line_parsing(bench):
    p = bench.ClockSampler.parse_line
    r = p("2026/10/20 13:22:01.250, 3, 1830, 1965, 900.1, 0x4, Not Active, Active, Not Active, Active")
    assert r[1:] == (1830.0, 1965.0, ["hw_thermal_slowdown", "sw_power_cap"])
    assert abs(r[0] - (time.mktime(time.strptime("2026/10/20 13:22:01", "%Y/%m/%d %H:%M:%S")) + 0.25)) < 1e-6
    assert p("garbage") is None and p("") is None
    assert p("a, 0, [N/A], 1965, 1, 0, Not Active, Not Active, Not Active, Not Active") is None
    # an unreadable time stamp keeps the sample (it then counts as outside any window)
    assert p("??, 0, 1965, 1965, 1, 0, Not Active, Not Active, Not Active, Not Active") == (None, 1965.0, 1965.0, [])


def test_sampler_reports_only_the_timed_window(bench, fake_smi):
    c = bench.ClockSampler(0, 50)
    c.start()
    time.sleep(0.9)                 # "warm-up": the sampler is up and has printed several lines
    c.begin()
    time.sleep(0.5)
    r = c.stop()
    assert r["sm_mhz"] == 1965.0 and r["sm_max_mhz"] == 1980.0 and r["reasons"] == ["sw_power_cap"]
    assert 2 <= r["samples"] <= 14, r          # ~0.5 s / 50 ms (fewer on a loaded box), not the ~1.1 s since the start (>= 16)
    assert "inside the timed region" in r["window"]


def test_sampler_without_nvidia_smi(bench, tmp_path, monkeypatch):
    monkeypatch.setenv("PATH", str(tmp_path))          # no nvidia-smi anywhere
    c = bench.ClockSampler(0)
    c.start()
    c.begin()
    assert c.stop() == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    assert bench.clock_snapshot(0)["samples"] == 0


def test_clock_snapshot(bench, fake_smi):
    r = bench.clock_snapshot(0)
    assert (r["sm_mhz"], r["sm_max_mhz"], r["samples"], r["reasons"]) == (1965.0, 1980.0, 1, ["sw_power_cap"])


def test_ic1_slabs_are_slices_of_the_grid(bench):
    row = col = 96
    M, C = bench.ic1(row, col, 0.9, 0.3)
    assert M.shape == (row * col,) and M.max() <= 0.9 and (C == 1).all()
    import ctypes
    import __graft_entry__ as ge
    lib = ctypes.CDLL(ge.build_cuda())
    for n in (2, 5, 8):
        parts = []
        for r in range(n):
            r0, nr = ctypes.c_int(), ctypes.c_int()
            assert lib.pdeode_slab_rows(row, n, r, ctypes.byref(r0), ctypes.byref(nr)) == 0
            parts.append(bench.ic1(row, col, 0.9, 0.3, r0.value, nr.value)[0])
        assert np.array_equal(np.concatenate(parts), M)


def test_both_arms_print_the_same_config(bench):
    a = argparse.Namespace(grid=4096, height=0.9, depth=0.3)
    for n in (1, 2, 8):
        assert bench.workload_config(a, n) == bench.workload_config(a, max(1, n))
    assert "single GPU" in bench.workload_config(a, 1)["parallelism"]
    assert bench.workload_config(a, 8)["parallelism"] == "row-slabs x8"
    assert bench.workload_config(a, 1)["l2"].startswith("inputs larger than L2")


def test_cpu_arm_uses_every_host_thread(bench, monkeypatch):
    monkeypatch.setenv("OMP_NUM_THREADS", "1")          # what torch.distributed.run exports
    n = bench.force_omp_threads()
    assert n == bench.host_threads() and os.environ["OMP_NUM_THREADS"] == str(n)
