#!/usr/bin/env python
"""patch_reference.py -- route 2 of INTEGRATION.md as a tool: keep the reference's Fortran host,
run its time step on the B200.

    python patch_reference.py /path/to/PDE-ODEsolver.f90 -o PDE-ODEsolver_b200.f90

Reads the reference's PDE-ODEsolver.f90 where it lies and writes a copy in which the body of
`solveOrder2` calls libpdeode_b200.so through the ISO_C_BINDING module next to this file
(pdeode_b200_mod.f90) instead of GenMatrixM / solveLSDIAG / solveC / calcDiff / calcMass /
calcPeakInterface.  Nothing else changes: parameter file, initial conditions, cadences, console
and output files stay the reference's own code.  Build (a machine with a Fortran compiler):

    gfortran -O3 -fdefault-real-8 pdeode_b200_mod.f90 PDE-ODEsolver_b200.f90 init_random_seed.f90 \\
             -L$REPO/pde-odesolver_b200/lib -lpdeode_b200 -Wl,-rpath,$REPO/pde-odesolver_b200/lib

(CGdiag.f90 is no longer needed.)  This image has no Fortran compiler; how the same two files are
built and run here nevertheless: tests/test_fortran_host.py.

The edits are anchored on statements of solveOrder2 (PDE-ODEsolver.f90:600-763); an anchor that is
missing or ambiguous means the source is not the one this was written for, and nothing is written.
The output is a derived file: it is never committed to this repository.
"""
import argparse
import re
import sys

USE = "  use pdeode_b200"

DECLS = """\
  type(c_ptr) :: ctx
  type(pdeode_params) :: pp
  integer(c_int) :: rc, cnt
  integer(c_long_long) :: nsum, nmax
"""

CREATE = """\
  pp = pdeode_params(delta, nu, kappa, gama, tDel, xDel, e1, e2, eSoln, alpha, beta, &
                     dSelect, fSelect, gSelect, 0)
  rc = pdeode_create(ctx, row, col, pp, 0)
  if (rc /= 0) then
    write(*,*) "[!] pdeode_create failed"
    stop
  end if
  rc = pdeode_set_state(ctx, M, C)
"""

MASS = "      rc = pdeode_mass(ctx, totalMassM, totalMassC)\n"
PEAK = "        rc = pdeode_peak(ctx, peak, height, intfac)\n"
FRAME = "      rc = pdeode_get_state(ctx, M, C)\n"

STEP = """\
    rc = pdeode_step(ctx, cnt, nsum, nmax)
    if (iand(rc, PDEODE_ERR_MASK) == PDEODE_ERR_PICARD_CAP) then
      write(*,*) "[!] Over 10000 iterations in one timestep"
      write(*,*) "[!] Solutions not converging. Exit!"
      stop
    end if
    if (nmax > maxNit) maxNit = nmax
    avgNit = avgNit + nsum
    countIters = cnt
"""

FINISH = """\
  rc = pdeode_get_state(ctx, M, C)
  rc = pdeode_destroy(ctx)
"""


def norm(line):
    return re.sub(r"\s+", "", line.split("!")[0]).lower()


def find_one(lines, lo, hi, pattern, what):
    hits = [k for k in range(lo, hi) if re.match(pattern, norm(lines[k]))]
    if len(hits) != 1:
        raise SystemExit(f"patch_reference: anchor {what!r} found {len(hits)} times in solveOrder2: "
                         "this is not the source the patch was written for")
    return hits[0]


def patch(text):
    lines = text.splitlines(keepends=True)
    n = len(lines)
    lo = find_one(lines, 0, n, r"subroutinesolveorder2\(", "subroutine solveOrder2")
    hi = find_one(lines, lo, n, r"endsubroutinesolveorder2$", "end subroutine solveOrder2")
    head_end = lo
    while norm(lines[head_end]).endswith("&"):
        head_end += 1
    a_decl = find_one(lines, lo, hi, r"real::peak,height,intfac$", "real :: peak, height, intfac")
    a_loop = find_one(lines, lo, hi, r"dowhile\(\(counter\)\*tdel<=tend\)$", "do while((counter) * tDel <= tEnd)")
    a_massM = find_one(lines, lo, hi, r"callcalcmass\(m,totalmassm,", "call calcMass(M, ...)")
    a_massC = find_one(lines, lo, hi, r"callcalcmass\(c,totalmassc,", "call calcMass(C, ...)")
    a_peak = find_one(lines, lo, hi, r"callcalcpeakinterface\(m,", "call calcPeakInterface")
    a_frame = find_one(lines, lo, hi, r"if\(mod\(counter,filter\)==0\)then$", "if (MOD(counter, filter) == 0) then")
    a_step0 = find_one(lines, lo, hi, r"diffc=1$", "diffC = 1")
    a_step1 = find_one(lines, lo, hi, r"if\(countiters>maxiters\)maxiters=countiters$",
                       "if(countIters > maxIters) maxIters = countIters")
    a_fin = find_one(lines, lo, hi, r"avgnit=avgnit/\(avgiters\)$", "avgNit = avgNit/(avgIters)")
    # the automatic work arrays of the CPU path (P:623-627) are no longer used: 9 n doubles of host memory
    unused = [find_one(lines, lo, hi, pat, pat) for pat in
              (r"real,dimension\(n\)::mnew$", r"real,dimension\(n\)::cnew$", r"real,dimension\(n,ndiag\)::matrixm$",
               r"real,dimension\(n\)::mrhs$", r"real,dimension\(n\)::cprev,mprev$")]
    if not (head_end < a_decl < a_loop < a_massM < a_massC < a_peak < a_frame < a_step0 < a_step1 < a_fin < hi):
        raise SystemExit("patch_reference: anchors out of order: not the source the patch was written for")
    if a_massC != a_massM + 1:
        raise SystemExit("patch_reference: the two calcMass calls are not adjacent")
    out = []
    for k, line in enumerate(lines):
        if k in unused:
            pass
        elif k == a_massM:
            out.append(MASS)
        elif k == a_massC:
            pass
        elif k == a_peak:
            out.append(PEAK)
        elif a_step0 <= k < a_step1:
            if k == a_step0:
                out.append(STEP)          # P:700-738; the two accumulations of P:739-740 stay
        elif k == a_loop:
            out.append(CREATE)
            out.append(line)
        elif k == a_fin:
            out.append(FINISH)
            out.append(line)
        else:
            out.append(line)
            if k == head_end:
                out.append(USE + "\n")
            elif k == a_decl:
                out.append(DECLS)
            elif k == a_frame:
                out.append(FRAME)
    return "".join(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("source", help="the reference's PDE-ODEsolver.f90")
    ap.add_argument("-o", "--output", required=True)
    args = ap.parse_args()
    with open(args.source) as f:
        text = f.read()
    new = patch(text)
    with open(args.output, "w") as f:
        f.write(new)
    print(f"patch_reference: wrote {args.output}", file=sys.stderr)


if __name__ == "__main__":
    main()
