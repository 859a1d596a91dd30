"""The row-slab protocol of the multi-GPU path, emulated on the CPU with two
`gloo` ranks.  Each rank holds only its slab plus one ghost row
above and below and follows the device protocol step by step:

  * ghost rows of d once per step, of x0 once per solve, of r after INIT and
    after every update (the peer-memory stores / ncclSend-Recv of the GPU path);
  * p's ghost rows are never exchanged: each rank recomputes r + beta*p on them,
    which must stay bit-identical to the owner's rows;
  * all-reduce of the four dot products per CG iteration; identical scalars,
    hence identical stop decisions on every rank.

The result must match the serial oracle: same CG iteration count, fields to
rounding.  The partition is the library's own (pdeode_slab_rows, no GPU).
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "pde-odesolver_b200", "python"))


def slab_rows(row, world, rank):
    import pdeode_b200 as pb
    r0, nr = C.c_int(), C.c_int()
    assert pb._lib().pdeode_slab_rows(row, world, rank, C.byref(r0), C.byref(nr)) == 0
    return r0.value, nr.value


def exchange_ghosts(X, rank, world):
    """X: (nrows + 2, col) with ghost rows 0 and -1; swap boundary rows with the neighbours."""
    reqs = []
    up, dn = rank - 1, rank + 1
    bufs = {}
    if up >= 0:
        reqs.append(dist.isend(torch.from_numpy(X[1].copy()), up))
        bufs["up"] = torch.empty(X.shape[1], dtype=torch.float64)
        reqs.append(dist.irecv(bufs["up"], up))
    if dn < world:
        reqs.append(dist.isend(torch.from_numpy(X[-2].copy()), dn))
        bufs["dn"] = torch.empty(X.shape[1], dtype=torch.float64)
        reqs.append(dist.irecv(bufs["dn"], dn))
    for q in reqs:
        q.wait()
    if "up" in bufs:
        X[0] = bufs["up"].numpy()
    if "dn" in bufs:
        X[-1] = bufs["dn"].numpy()


def allsum(*vals):
    t = torch.tensor(list(vals), dtype=torch.float64)
    dist.all_reduce(t)
    return [float(v) for v in t]


def worker(rank, world, port, row, col, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import np_model
    import oracle_lib as ol
    p = ol.default_params(row)
    M0, C0 = ol.ic1(row, col, 0.9, 0.3)
    rng = np.random.RandomState(4)
    M0 = (M0 * (1 + 0.03 * rng.rand(M0.size))).reshape(row, col)
    C0 = C0.reshape(row, col)
    r0, nr = slab_rows(row, world, rank)
    n = row * col

    def ext(field):                       # my slab with zeroed ghost rows
        X = np.zeros((nr + 2, col))
        X[1:-1] = field[r0:r0 + nr]
        return X

    def to_global(X):                     # slab + ghosts placed into a global-size vector
        G = np.zeros((row, col))
        lo, hi = max(r0 - 1, 0), min(r0 + nr + 1, row)
        G[lo:hi] = X[1 - (r0 - lo):1 + nr + (hi - r0 - nr)]
        return G.ravel()

    Mprev = ext(M0); Cc = ext(C0)
    exchange_ghosts(Mprev, rank, world)   # stands for the ghost rows of d = D(Mprev)
    exchange_ghosts(Cc, rank, world)
    # my rows of the global operator, built from slab + ghost data only
    A = np_model.dense_operator(to_global(Mprev), np.maximum(to_global(Cc), 1e-300), row, col, p)
    mine = slice(r0 * col, (r0 + nr) * col)
    A = A[mine]

    def apply(V):                         # q = A v on my rows; V has valid ghost rows
        return (A @ to_global(V)).reshape(nr, col)

    x = np.zeros((nr + 2, col))           # first solve ever: x0 = 0 (D8)
    r = np.zeros((nr + 2, col)); pvec = np.zeros((nr + 2, col))
    rhs = Mprev[1:-1] / p.tDel
    exchange_ghosts(x, rank, world)
    r[1:-1] = rhs - apply(x)
    rho, xx = allsum(float((r[1:-1] ** 2).sum()), float((x[1:-1] ** 2).sum()))
    exchange_ghosts(r, rank, world)
    nit, rho_old = 0, 0.0
    ghost_ok = True
    while True:
        nit += 1
        normx = np.sqrt(xx) / n
        # K_A: p = r + beta p on my rows AND on the ghost rows, never exchanged
        pvec = r.copy() if nit == 1 else r + (rho / rho_old) * pvec
        check = pvec.copy()
        exchange_ghosts(check, rank, world)               # what the owners hold
        ghost_ok = ghost_ok and np.array_equal(check, pvec)
        q = apply(pvec)
        pq, pp = allsum(float((pvec[1:-1] * q).sum()), float((pvec[1:-1] ** 2).sum()))
        alfa = rho / pq
        err2 = abs(alfa) * np.sqrt(pp) / n
        stop = (nit > 100 * n) or (err2 < p.e2 * normx)
        # K_B
        x[1:-1] += alfa * pvec[1:-1]
        r[1:-1] -= alfa * q
        rho_old = rho
        rho, xx = allsum(float((r[1:-1] ** 2).sum()), float((x[1:-1] ** 2).sum()))
        exchange_ghosts(r, rank, world)
        if stop:
            break
    np.save(os.path.join(out, f"x{rank}.npy"), x[1:-1])
    with open(os.path.join(out, f"info{rank}.txt"), "w") as f:
        f.write(f"{nit} {int(ghost_ok)} {r0} {nr}\n")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("row,col", [(24, 10), (19, 7)])
def test_slab_protocol_matches_serial_oracle(tmp_path, row, col):
    import __graft_entry__ as ge
    ge.build_cuda()
    import np_model
    import oracle_lib as ol
    ol.build_oracle()
    world = 2
    port = 29600 + (os.getpid() % 200)
    mp.spawn(worker, args=(world, port, row, col, str(tmp_path)), nprocs=world, join=True)
    parts, infos = [], []
    for r in range(world):
        parts.append(np.load(tmp_path / f"x{r}.npy"))
        infos.append([int(v) for v in open(tmp_path / f"info{r}.txt").read().split()])
    assert infos[0][0] == infos[1][0]                     # same decisions on every rank
    assert all(i[1] == 1 for i in infos)                  # self-maintained p ghosts are exact
    assert infos[0][2] == 0 and infos[1][2] == infos[0][3] and infos[0][3] + infos[1][3] == row
    x = np.concatenate(parts).ravel()
    # serial oracle, same system
    p = ol.default_params(row)
    M0, C0 = ol.ic1(row, col, 0.9, 0.3)
    rng = np.random.RandomState(4)
    M0 = M0 * (1 + 0.03 * rng.rand(M0.size))
    o = ol.Oracle(row, col, p)
    o.set_state(M0, C0)
    nit, _ = o.first_solve_trace(1)
    xs = o.get_mnew()
    assert abs(nit - infos[0][0]) <= 2
    tol = 1e-9 if nit == infos[0][0] else 2e-7
    assert np.linalg.norm(x - xs) <= tol * np.linalg.norm(xs)


def test_slab_partition_covers_the_grid():
    import __graft_entry__ as ge
    ge.build_cuda()
    for row in (3, 17, 257, 4096, 16384):
        for world in (1, 2, 3, 4, 8):
            if row // world < 2:
                continue
            edges = [slab_rows(row, world, r) for r in range(world)]
            assert edges[0][0] == 0
            for a, b in zip(edges, edges[1:]):
                assert a[0] + a[1] == b[0]
            assert edges[-1][0] + edges[-1][1] == row
            assert max(e[1] for e in edges) - min(e[1] for e in edges) <= 1
