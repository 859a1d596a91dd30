#!/usr/bin/env python
"""Runs the host program with the CPU oracle as its stepper (tests' adapter):
the closest thing to `echo FILE | ./a.out` of the reference that this image
can run (no Fortran compiler).  Test infrastructure (it drives the oracle).
Usage: run_cpu_program.py PARAM.txt OUTDIR      (ORACLE_OMP=1: the OpenMP build of the oracle)"""
import ctypes as C, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
omp = os.environ.get("ORACLE_OMP") == "1"
import __graft_entry__ as ge
import oracle_lib as ol
ge.build_host(); ol.build_oracle()
build = os.path.join(ROOT, "tests", "_build"); os.makedirs(build, exist_ok=True)
tag = os.environ.get("ADAPTER_TAG", "")          # a private copy while another run has the library mapped
so = os.path.join(build, "libhost_oracle_adapter" + ("_omp" if omp else "") + tag + ".so")
subprocess.check_call(["/usr/bin/gcc", "-O2", "-fPIC", "-shared", "-o", so, os.path.join(ROOT, "tests", "host_oracle_adapter.c"),
                       "-L" + ol.ORACLE_DIR, "-loracle_omp" if omp else "-loracle", "-Wl,-rpath," + ol.ORACLE_DIR])
ad = C.CDLL(so); ad.oracle_stepper_table.restype = C.c_void_p
host = C.CDLL(os.path.join(ge.LIBDIR, "libpdeode_hostlogic.so"))
host.pdeode_host_run.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_void_p, C.c_int]
out = sys.argv[2]; os.makedirs(out, exist_ok=True)
t0 = time.time()
rc = host.pdeode_host_run(os.path.abspath(sys.argv[1]).encode(), out.encode(), os.path.join(out, "console.txt").encode(), ad.oracle_stepper_table(), 0)
print("rc", rc, "wall %.1f s" % (time.time() - t0))
print(open(os.path.join(out, "statReport.dat")).read())
