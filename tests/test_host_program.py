"""Host program (pde-odesolver_b200/host): parameter file, initial conditions,
gfortran-style formatting, frame writers and the whole time loop, driven on the
CPU through the oracle adapter (tests/host_oracle_adapter.c)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import __graft_entry__ as ge
import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(HERE, "data")
BUILD = os.path.join(HERE, "_build")


class HostParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("pSize", "true2D", "numInnocu", "alpha", "beta", "nOuts",
                                       "fSelect", "dSelect", "gSelect", "MinitialCond", "row", "col")] + \
               [("n", C.c_longlong), ("nit", C.c_longlong)] + \
               [(n, C.c_double) for n in ("depth", "height", "nu", "kappa", "gama", "delta", "tEnd",
                                          "tDel", "e1", "e2", "eSoln", "xDel")]


@pytest.fixture(scope="module")
def host():
    ge.build_cuda()
    ge.build_host()
    ol.build_oracle()
    lib = C.CDLL(os.path.join(ge.LIBDIR, "libpdeode_hostlogic.so"))
    lib.pdeode_host_read_parameters.argtypes = [C.c_char_p, C.POINTER(HostParams), C.c_char_p, C.c_int]
    lib.pdeode_host_initial_conditions.argtypes = [C.POINTER(HostParams), C.c_ulonglong, C.c_void_p, C.c_void_p]
    lib.pdeode_host_format_real.argtypes = [C.c_double, C.c_char_p, C.c_int]
    lib.pdeode_host_write_frame.argtypes = [C.c_char_p, C.POINTER(HostParams), C.c_void_p, C.c_void_p, C.c_int]
    lib.pdeode_host_run.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_int]
    return lib


@pytest.fixture(scope="module")
def adapter():
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "libhost_oracle_adapter.so")
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-fPIC", "-shared", "-o", out,
                           os.path.join(HERE, "host_oracle_adapter.c"),
                           "-L" + ol.ORACLE_DIR, "-loracle", "-Wl,-rpath," + ol.ORACLE_DIR])
    lib = C.CDLL(out)
    lib.oracle_stepper_table.restype = C.c_void_p
    return lib


def read_params(host, path):
    p = HostParams()
    err = C.create_string_buffer(256)
    rc = host.pdeode_host_read_parameters(path.encode(), C.byref(p), err, 256)
    return rc, p, err.value.decode()


def fmt(host, x):
    b = C.create_string_buffer(64)
    host.pdeode_host_format_real(x, b, 64)
    return b.value.decode()


def test_parameter_file_is_positional(host, tmp_path):
    rc, p, _ = read_params(host, os.path.join(DATA, "wave2d.txt"))
    assert rc == 0
    assert (p.pSize, p.true2D, p.row, p.col, p.n) == (129, 1, 129, 4, 516)
    assert (p.alpha, p.beta, p.nOuts, p.fSelect, p.dSelect, p.gSelect, p.MinitialCond) == (4, 4, 4, 1, 3, 1, 1)
    assert (p.depth, p.height, p.nu, p.kappa, p.gama, p.delta) == (0.3, 0.9, 0.30, 0.01, 0.60, 0.0001)
    assert (p.tEnd, p.tDel, p.e1, p.e2, p.eSoln) == (0.04, 0.001, 1e-8, 1e-8, 1e-8)
    assert p.xDel == 1.0 / 129 and p.nit == 51600          # recomputed, file lines skipped (P:110-111)
    # names are ignored, positions are not: rename every variable, same result
    txt = open(os.path.join(DATA, "wave2d.txt")).read().splitlines()
    renamed = [("zz" + l[2:]) if "=" in l and not l.startswith("!") else l for l in txt]
    q = tmp_path / "renamed.txt"
    q.write_text("\n".join(renamed) + "\n")
    rc2, p2, _ = read_params(host, str(q))
    assert rc2 == 0 and bytes(p2) == bytes(p)
    # a missing line shifts everything: the reader reports it rather than guessing
    q.write_text("\n".join(txt[:20]) + "\n")
    rc3, _, err = read_params(host, str(q))
    assert rc3 != 0 and "29" in err
    rc4, _, err4 = read_params(host, str(tmp_path / "nope.txt"))
    assert rc4 != 0 and "cannot open" in err4


def test_square_mode_sets_col_to_psize(host):
    rc, p, _ = read_params(host, os.path.join(DATA, "square3d.txt"))
    assert rc == 0 and (p.row, p.col, p.n) == (33, 33, 1089)


def test_gfortran_list_directed_real_format(host):
    # strings as gfortran >= 4.9 prints REAL(8) with write(*,*)
    assert fmt(host, 1.0) == "   1.0000000000000000     "
    assert fmt(host, 0.0) == "   0.0000000000000000     "
    assert fmt(host, 0.1) == "  0.10000000000000001     "
    assert fmt(host, 123.456) == "   123.45600000000000     "
    assert fmt(host, 1.0e-3) == "   1.0000000000000000E-003"
    assert fmt(host, -2.5) == "  -2.5000000000000000     "
    assert fmt(host, -0.0625) == "  -6.2500000000000000E-002"
    assert fmt(host, 1.0e17) == "   1.0000000000000000E+017"
    assert fmt(host, 0.02) == "   2.0000000000000000E-002"
    for x in (0.3, 7.25e-5, 12345.678, 0.999999999, 1 / 3):
        assert float(fmt(host, x)) == x and len(fmt(host, x)) == 26


@pytest.mark.parametrize("ic", [1, 2, 3, 5, 6, 8, 9])
def test_initial_conditions(host, ic):
    rc, p, _ = read_params(host, os.path.join(DATA, "square3d.txt"))
    p.MinitialCond = ic
    p.numInnocu = 5
    p.depth, p.height = 0.2, 0.3
    M = np.empty(p.n); Cc = np.empty(p.n)
    host.pdeode_host_initial_conditions(C.byref(p), 42, M.ctypes.data, Cc.ctypes.data)
    assert np.all(Cc == 1.0)                                  # P:276
    if ic == 1:
        Mw, _ = ol.ic1(p.row, p.col, p.height, p.depth)
        assert np.array_equal(M, Mw)
    elif ic == 2:
        assert np.all(M == p.height)
    elif ic == 5:
        nz = M > 0
        assert M[nz].mean() == pytest.approx(p.height, rel=1e-12)   # normalisation, total = 0 (D5)
        M2 = np.empty(p.n)
        host.pdeode_host_initial_conditions(C.byref(p), 42, M2.ctypes.data, Cc.ctypes.data)
        assert np.array_equal(M, M2)                           # seeded: reproducible
    elif ic == 6:
        rows = np.arange(p.row) * p.xDel <= p.depth
        assert np.array_equal(M.reshape(p.row, p.col)[:, 0] == p.height, rows)
    elif ic == 8:
        i = np.arange(1, p.n + 1)
        want = i / p.n / 4.0
        assert np.array_equal(M, np.where(want >= 0.1, 0.0, want))   # K4 ramp (P:384-387)
    assert np.all(M >= 0) and M.max() > 0


def split_frames(path):
    """seperateOutput.py / seperateOutput2D.py of the reference: a new frame
    starts at every line that is exactly a newline."""
    frames, cur = [], []
    for line in open(path).readlines():
        if line == "\n":
            frames.append(cur); cur = []
        cur.append(line)
    return frames + [cur]


def test_frame_writers(host, tmp_path):
    rc, p, _ = read_params(host, os.path.join(DATA, "square3d.txt"))
    n = p.n
    M = np.arange(n) / n; Cc = 1 - M
    f = str(tmp_path / "output.dat")
    host.pdeode_host_write_frame(f.encode(), C.byref(p), M.ctypes.data, Cc.ctypes.data, 0)
    host.pdeode_host_write_frame(f.encode(), C.byref(p), M.ctypes.data, Cc.ctypes.data, 1)
    fr = split_frames(f)
    assert len(fr) == 3 and fr[2] == ["\n"]                   # 2 frames + the trailing terminator
    rows = [l for l in fr[0] if l.strip()]
    assert len(rows) == n                                     # row <= 257: no decimation (P:829-832)
    x, y, m, c = (float(v) for v in rows[p.col + 2].split())  # cell i=2, j=3 (1-based)
    assert (x, y) == (2 / (p.col - 1), 1 / (p.row - 1))
    assert m == M[p.col + 2] and c == Cc[p.col + 2]
    # 2-D writer: y, column means, f20.12
    rc, p2, _ = read_params(host, os.path.join(DATA, "wave2d.txt"))
    M2 = np.repeat(np.arange(p2.row) / p2.row, p2.col); C2 = np.ones(p2.n)
    f2 = str(tmp_path / "2D_output.dat")
    host.pdeode_host_write_frame(f2.encode(), C.byref(p2), M2.ctypes.data, C2.ctypes.data, 0)
    lines = open(f2).readlines()
    assert len(lines) == p2.row + 1 and lines[-1] == "\n" and len(lines[0]) == 61
    y, am, ac = (float(v) for v in lines[5].split())
    assert y == pytest.approx(5 / (p2.row - 1), abs=1e-12) and am == pytest.approx(5 / p2.row, abs=1e-12)


@pytest.mark.parametrize("size", [33, 257, 600])
def test_decimated_frame_writer_equals_full_field_writer(host, tmp_path, size):
    """print_frame (device-decimated data, rows formatted by several threads above
    16384 points) must write the bytes print_to_file writes from the whole fields."""
    rc, p, _ = read_params(host, os.path.join(DATA, "square3d.txt"))
    p.pSize = p.row = p.col = size
    p.n = size * size
    rng = np.random.RandomState(size)
    M = rng.uniform(-1e-3, 1.0, p.n) * (rng.uniform(size=p.n) > 0.3)      # zeros, tiny and negative values
    M[:5] = [0.0, -0.0, 1e-300, 123456.789, -2.5e-7]
    Cc = rng.uniform(0, 1, p.n)
    a, b = str(tmp_path / "a.dat"), str(tmp_path / "b.dat")
    host.pdeode_host_write_frame_decimated.argtypes = host.pdeode_host_write_frame.argtypes
    for k in (0, 1):
        assert host.pdeode_host_write_frame(a.encode(), C.byref(p), M.ctypes.data, Cc.ctypes.data, k) == 0
        assert host.pdeode_host_write_frame_decimated(b.encode(), C.byref(p), M.ctypes.data, Cc.ctypes.data, k) == 0
    assert open(a, "rb").read() == open(b, "rb").read()


@pytest.mark.parametrize("tEnd,tDel,nOuts", [(30.0, 1e-3, 90), (0.05, 1e-3, 5), (1.0, 1e-2, 7), (2.0, 1e-3, 3)])
def test_batched_steps_visit_the_same_counters(host, tEnd, tDel, nOuts):
    """The GPU program advances in batches (pdeode_step_many); the counters at which it
    looks at the state -- diagnostics P:665, frames P:683, loop end P:663 -- must be the
    ones the reference's one-step-at-a-time loop visits."""
    host.pdeode_host_steps_until_next_look.argtypes = [C.c_longlong, C.c_int, C.c_double, C.c_double]
    filt = int(tEnd / (nOuts * tDel))                      # P:638
    cad = filt // 100 + 1
    # the reference loop
    looks, counter = [], 0
    while counter * tDel <= tEnd:
        if counter % cad == 0 or counter % filt == 0:
            looks.append(counter)
        counter += 1
    steps_ref = counter
    # the batched loop
    got, counter = [], 0
    while counter * tDel <= tEnd:
        assert counter % cad == 0 or counter % filt == 0 or not got or counter - got[-1] >= 4096
        got.append(counter)
        counter += host.pdeode_host_steps_until_next_look(counter, filt, tDel, tEnd)
    assert counter == steps_ref
    assert [c for c in got if c % cad == 0 or c % filt == 0] == looks


def test_whole_program_with_oracle_stepper(host, adapter, tmp_path):
    """`echo FILE | a.out` end to end: console text, files, cadences (S6)."""
    out = tmp_path / "run"; out.mkdir()
    con = str(tmp_path / "console.txt")
    os.environ["PDEODE_SEED"] = "1"
    rc = host.pdeode_host_run(os.path.join(DATA, "wave2d.txt").encode(), str(out).encode(),
                              con.encode(), adapter.oracle_stepper_table(), 0)
    assert rc == 0
    names = sorted(os.listdir(out))
    assert names == ["2D_output.dat", "COprod.dat", "peakInfo.dat", "statReport.dat", "total.dat"]
    # tEnd/(nOuts*tDel) = 10 -> frames at steps 0,10,20,30,40; diagnostics every filter/100+1 = 1 step
    frames = split_frames(str(out / "2D_output.dat"))
    assert len(frames) == 5 + 1
    total = np.loadtxt(out / "total.dat")
    assert total.shape == (41, 3) and total[0, 0] == 0.0 and total[-1, 0] == pytest.approx(0.04)
    peak = np.loadtxt(out / "peakInfo.dat")
    assert peak.shape == (41, 4)
    co = np.loadtxt(out / "COprod.dat")
    assert co[0, 1] == pytest.approx(1 - total[0, 2]) and co[3, 2] == pytest.approx(1 - total[3, 2])
    text = open(con).read()
    assert "Enter parameter file name" in text and "Problem is 2D" in text
    assert "   time    avgIter  maxIter      avgNit  maxNit        avgM        avgC" in text
    lines = text.splitlines()
    hdr = next(i for i, l in enumerate(lines) if "avgIter" in l)
    frame_lines = lines[hdr + 1:hdr + 6]
    assert [float(l.split()[0]) for l in frame_lines] == [0.0, 0.01, 0.02, 0.03, 0.04]
    assert all(len(l) == 72 for l in frame_lines)             # (F8.2,F12.2,I8,F12.2,I8,F12.6,F12.6)
    assert "Execution completed" in lines[-1]
    stat = open(out / "statReport.dat").read()
    assert "Time to compute" in stat and "Max Iters for linear solver" in stat

    # the same run driven directly from python with the oracle: same numbers
    _, p, _ = read_params(host, os.path.join(DATA, "wave2d.txt"))
    M0, C0 = ol.ic1(p.row, p.col, p.height, p.depth)
    o = ol.Oracle(p.row, p.col, ol.default_params(p.row))
    o.set_state(M0, C0)
    lib = ol.load()
    sumP = sumN = 0
    for k in range(41):
        M, Cc = o.get_state()
        assert total[k, 1] == lib.oracle_calc_mass(M, M.size)
        assert total[k, 2] == lib.oracle_calc_mass(Cc, Cc.size)
        if k == 40:
            # frame 5 = state before step 40
            fr = np.array([[float(v) for v in l.split()] for l in frames[4] if l.strip()])
            assert np.allclose(fr[:, 1], M.reshape(p.row, p.col).mean(axis=1), atol=1e-12)
        r = o.step()
        sumP += r["countIters"]; sumN += r["nitSum"]
    assert f"{sumP / 41:.16g}"[:10] in stat.replace("E+000", "") or True
    avg_iters = float([l for l in stat.splitlines() if "Avg Iters for iterating" in l][0].split("=")[1])
    avg_nit = float([l for l in stat.splitlines() if "Avg Iters for linear" in l][0].split("=")[1])
    assert avg_iters == pytest.approx(sumP / 41, rel=1e-15)
    assert avg_nit == pytest.approx(sumN / sumP, rel=1e-15)


def test_final_state_and_restart(host, adapter, tmp_path, monkeypatch):
    """PDEODE_FINAL_STATE writes the final fields (the print the reference has
    commented out, P:747-755); MinitialCond = 11 continues from such a file (the
    restart the reference could not get working, P:423-435)."""
    base = open(os.path.join(DATA, "wave2d.txt")).read()
    full = tmp_path / "full"; a = tmp_path / "a"; b = tmp_path / "b"
    for d in (full, a, b):
        d.mkdir()
    # 41 steps in one go
    (tmp_path / "full.txt").write_text(base)
    monkeypatch.setenv("PDEODE_FINAL_STATE", "final.dat")
    assert host.pdeode_host_run(str(tmp_path / "full.txt").encode(), str(full).encode(),
                                str(tmp_path / "c0.txt").encode(), adapter.oracle_stepper_table(), 0) == 0
    # 21 steps, then 20 more from the written state
    (tmp_path / "first.txt").write_text(base.replace("tEnd      = 0.04", "tEnd      = 0.0205")
                                        .replace("nOuts     = 4", "nOuts     = 2"))
    assert host.pdeode_host_run(str(tmp_path / "first.txt").encode(), str(a).encode(),
                                str(tmp_path / "c1.txt").encode(), adapter.oracle_stepper_table(), 0) == 0
    fin = np.loadtxt(a / "final.dat")
    assert fin.shape == (129 * 4, 4)
    os.rename(a / "final.dat", b / "initialCond.dat")
    (tmp_path / "second.txt").write_text(base.replace("tEnd      = 0.04", "tEnd      = 0.0195")
                                         .replace("nOuts     = 4", "nOuts     = 1")
                                         .replace("MinitialCond = 1", "MinitialCond = 11"))
    assert host.pdeode_host_run(str(tmp_path / "second.txt").encode(), str(b).encode(),
                                str(tmp_path / "c2.txt").encode(), adapter.oracle_stepper_table(), 0) == 0
    assert "continue from initialCond.dat" in open(tmp_path / "c2.txt").read()
    want = np.loadtxt(full / "final.dat")
    got = np.loadtxt(b / "final.dat")
    # same trajectory; the restart only loses the CG warm start of its first solve
    assert np.allclose(got[:, 2:], want[:, 2:], rtol=0, atol=2e-7)
    assert np.array_equal(got[:, :2], want[:, :2])
    # a wrong-size file is refused
    (b / "initialCond.dat").write_text("0 0 0.5 1\n")
    assert host.pdeode_host_run(str(tmp_path / "second.txt").encode(), str(b).encode(),
                                str(tmp_path / "c3.txt").encode(), adapter.oracle_stepper_table(), 0) == 2


def test_gfortran_shim_creates_a_out(tmp_path):
    ge.build_cuda(); ge.build_host()
    shim = os.path.join(ge.PKG, "scripts", "gfortran")
    rc = subprocess.call(["bash", shim, "-O3", "-fdefault-real-8", "PDE-ODEsolver.f90", "CGdiag.f90",
                          "init_random_seed.f90"], cwd=tmp_path)
    assert rc == 0 and os.access(tmp_path / "a.out", os.X_OK)
    assert "pdeode_solver" in open(tmp_path / "a.out").read()
