"""Builds the reference's OWN Fortran host with its time step on the C ABI (TEST INFRASTRUCTURE).

north_star: "the Fortran host code stays and calls CUDA through a thin ISO_C_BINDING C-ABI
layer".  pde-odesolver_b200/fortran/patch_reference.py applies INTEGRATION.md's route 2 to the
reference's PDE-ODEsolver.f90 where it lies; a maintainer compiles the result and
pdeode_b200_mod.f90 with gfortran.  This image has no Fortran compiler, so here the two files go
through the Fortran-to-C translation that also builds oracle/_ref (oracle/f90c/f90toc.py) and
are linked

  fortran_host_cpu   against a test-only library that implements the ABI with the CPU oracle
                     (tests/host_oracle_adapter.c, -DADAPTER_EXPORT_ABI)
  fortran_host_b200  against pde-odesolver_b200/lib/libpdeode_b200.so (rpath relative to the file)

into tests/_build/ (git-ignored; travels to the GPU box, which has no /root/reference).
Generated sources are never committed."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(HERE, "_build")
REF_SRC = "/root/reference"
FDIR = os.path.join(ROOT, "pde-odesolver_b200", "fortran")
F90C = os.path.join(ROOT, "oracle", "f90c")
CC = "/usr/bin/gcc"
CFLAGS = ["-O3", "-ffp-contract=off", "-fno-fast-math", "-fwrapv", "-std=gnu11", "-Wno-unknown-pragmas", "-I" + F90C]


def have_sources():
    return os.path.isfile(os.path.join(REF_SRC, "PDE-ODEsolver.f90"))


def exe(kind):
    return os.path.join(BUILD, "fortran_host_" + kind)


def available(kind):
    return os.path.exists(exe(kind)) or have_sources()


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _sources():
    return [os.path.join(FDIR, "patch_reference.py"), os.path.join(FDIR, "pdeode_b200_mod.f90"),
            os.path.join(F90C, "f90toc.py"), os.path.join(F90C, "f90rt.c"), os.path.join(F90C, "f90rt.h"),
            os.path.join(REF_SRC, "PDE-ODEsolver.f90"), os.path.join(REF_SRC, "init_random_seed.f90")]


def generate(workdir):
    """patched Fortran -> C inside workdir (derived from the reference: removed after the build,
    only the executables stay under tests/_build/)."""
    patched = os.path.join(workdir, "PDE-ODEsolver_b200.f90")
    gen = os.path.join(workdir, "fortran_host_gen.c")
    subprocess.check_call([sys.executable, os.path.join(FDIR, "patch_reference.py"),
                           os.path.join(REF_SRC, "PDE-ODEsolver.f90"), "-o", patched], stderr=subprocess.DEVNULL)
    subprocess.check_call([sys.executable, os.path.join(F90C, "f90toc.py"), "-o", gen,
                           os.path.join(FDIR, "pdeode_b200_mod.f90"), patched,
                           os.path.join(REF_SRC, "init_random_seed.f90")])
    return gen


def build(kind):
    """kind = 'cpu' | 'b200'.  Returns the executable (prebuilt one where the sources are absent)."""
    import shutil
    import tempfile
    out = exe(kind)
    if not have_sources():
        if not os.path.exists(out):
            raise FileNotFoundError(out + " not prebuilt and /root/reference absent")
        return out
    os.makedirs(BUILD, exist_ok=True)
    rt = os.path.join(F90C, "f90rt.c")
    if kind == "cpu":
        if HERE not in sys.path:
            sys.path.insert(0, HERE)
        import oracle_lib as ol
        ol.build_oracle()
        so = os.path.join(BUILD, "libabi_oracle.so")
        ad = os.path.join(HERE, "host_oracle_adapter.c")
        if _newer(so, [ad, os.path.join(ol.ORACLE_DIR, "liboracle.so")]):
            subprocess.check_call([CC, "-O2", "-fPIC", "-shared", "-DADAPTER_EXPORT_ABI", "-o", so, ad,
                                   "-L" + ol.ORACLE_DIR, "-loracle", "-Wl,-rpath," + ol.ORACLE_DIR])
        deps, link = _sources() + [so], ["-L" + BUILD, "-labi_oracle", "-Wl,-rpath,$ORIGIN"]
    else:
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        import __graft_entry__ as ge
        lib = ge.build_cuda()
        deps, link = _sources() + [lib], ["-L" + ge.LIBDIR, "-lpdeode_b200",
                                          "-Wl,-rpath,$ORIGIN/../../pde-odesolver_b200/lib"]
    if _newer(out, deps):
        work = tempfile.mkdtemp(prefix="fortran_host_")
        try:
            gen = generate(work)
            subprocess.check_call([CC] + CFLAGS + ["-DF90_MAIN", "-o", out, gen, rt] + link + ["-lm"])
        finally:
            shutil.rmtree(work, ignore_errors=True)
    return out


if __name__ == "__main__":
    for k in sys.argv[1:] or ["cpu", "b200"]:
        print(build(k))
