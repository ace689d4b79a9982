"""CPU, world_size = 2, gloo: the N>1 host logic of the sharded finite-difference driver
(partition of the points, the one amplitude exchange, element partition, final tensor gather)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _code(p):
    return float(ord(p[0]) + 7 * p[1] + 3 * (p[2] + 1))


def _worker(rank, world, port, natom, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from apyib_b200.parallel import partition, exchange_points, owned_elements, gather_tensor, _dist
        from apyib_b200.fin_diff import aat_points, point_cost
        d, r, w = _dist()
        assert (r, w) == (rank, world)
        # the unperturbed point is one more point of the partition: solved by one rank, exchanged with the others
        pts = [("U", 0, 0)] + aat_points(natom)
        own = partition(pts, [point_cost(p[0]) for p in pts], world)
        # payload shaped like the real one: (C, [1, t1, t2]); real for R points, complex for B points,
        # odd byte counts (alignment padding), a scalar 0 standing for "no singles"
        def payload(p):
            dt = np.complex128 if p[0] == "B" else np.float64
            c = _code(p)
            return (rank, np.full((3, 3), c, dtype=dt), [1, 0 if p[1] % 2 else np.arange(5, dtype=dt) * c,
                                                        np.full((2, 2, 3, 3), c + 1, dtype=dt)])
        mine = {p: payload(p) for p, o in zip(pts, own) if o == rank}
        allp = exchange_points(d, mine, world)
        assert sorted(allp) == sorted(pts)
        for p, (src, C, T) in allp.items():
            dt = np.complex128 if p[0] == "B" else np.float64
            assert src == own[pts.index(p)] and C.dtype == dt and C.shape == (3, 3) and C[0, 0] == _code(p)
            assert T[0] == 1 and T[2].shape == (2, 2, 3, 3) and T[2].dtype == dt and T[2][1, 1, 2, 2] == _code(p) + 1
            if p[1] % 2:
                assert isinstance(T[1], int) and T[1] == 0
            else:
                assert np.array_equal(np.asarray(T[1]), np.arange(5, dtype=dt) * _code(p))
        n3 = 3 * natom
        I = np.zeros((n3, 3))
        for a, b in owned_elements(n3, rank, world):
            I[a, b] = 100 * a + b + 0.5
        I = gather_tensor(d, I, world)
        want = np.array([[100 * a + b + 0.5 for b in range(3)] for a in range(n3)])
        # energy-only drivers (APT): disjointly filled energies, one all-reduce, identical on every rank
        from apyib_b200.parallel import gather_energies
        from apyib_b200.fin_diff import finite_difference
        fd = finite_difference.__new__(finite_difference)
        fd.natom = natom
        apts = fd.apt_points()
        aown = partition(apts, [1.0] * len(apts), world)
        ecode = lambda p: -76.0 + 1e-3 * p[0] + 1e-4 * p[1] + 1e-5 * p[2] + 1e-6 * p[3]
        E = gather_energies(d, [ecode(p) if o == rank else None for p, o in zip(apts, aown)], aown, rank, world)
        ok_apt = len(apts) == 36 * natom and E == [ecode(p) for p in apts] and abs(aown.count(0) - aown.count(1)) <= 1
        apt = fd.compute_APT(1e-3, 1e-4, energies=dict(zip(apts, E)))
        ok_apt = ok_apt and apt.shape == (n3, 3)
        q.put((rank, bool(np.array_equal(I, want)) and ok_apt, len(mine)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_driver_logic_world2():
    world, natom = 2, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, natom, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=90) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert sum(n for _, _, n in res) == 6 * natom + 7          # 6N + 6 displaced / field points + the unperturbed one
    loads = sorted(n for _, _, n in res)
    assert loads[-1] - loads[0] <= 12         # complex points weigh 5.5x, counts may differ


def test_single_process_degenerates():
    from apyib_b200.parallel import exchange_points, owned_elements, gather_tensor
    assert exchange_points(None, {1: 2}, 1) == {1: 2}
    assert owned_elements(2, 0, 1) == [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (1, 2)]
    assert owned_elements(3, 1, 2) == [(1, 0), (1, 1), (1, 2)]
    I = np.ones((2, 3))
    assert gather_tensor(None, I, 1) is I
