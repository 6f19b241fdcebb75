"""Host-side multi-rank logic on CPU: world_size 2 over torch.distributed `gloo`.

What the multi-GPU path relies on from the host -- the Morton-chunk partition with one ghost layer,
the per-neighbour send / receive lists (the role of GhostCommunicator's exchange lists,
src/communications.cpp:621-630), the box decomposition of the uniform path, the order of exchanges
and the MAX all-reduce of the stage-1 eigenvalue (src/main.cpp:393, 425-430, 461-466, 497-502) -- is
exercised with the CPU oracle standing in for the kernels: every rank runs the reference-shaped stage
sequence on its chunk, ghosts travel through gloo send/recv following the lists, and the assembled
result must equal the SERIAL run bit for bit (the reference expects one golden string from 1 and 3
ranks, test/<case>/CMakeLists.txt:36-40)."""
import os
import socket
import sys

import numpy as np
import pytest

from common import bits_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _exchange(dist, torch, rank, comm, field):
    """startAllExchanges + completeAllExchanges: post sends (5 doubles per listed cell, list order),
    receive into the ghost cells."""
    reqs, bufs = [], []
    for q, nb in enumerate(comm["neighbours"]):
        send = torch.from_numpy(np.ascontiguousarray(field[comm["send"][q]]))
        recv = torch.empty((len(comm["recv"][q]), 5), dtype=torch.float64)
        reqs.append(dist.isend(send, nb))
        reqs.append(dist.irecv(recv, nb))
        bufs.append((q, recv, send))
    for r in reqs:
        r.wait()
    for q, recv, _ in bufs:
        field[comm["recv"][q]] = recv.numpy()


def _worker(rank, world, port, problem, dim, n, steps, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    import oracle_lib
    from minimmerflow_b200.partition import morton_chunk_partition
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oracle_lib.load()
    m = orc.problem_mesh(problem, dim, n)
    U0 = orc.init_state(m)
    loc, comm = morton_chunk_partition(m, world, rank)
    gids, n_int = comm["global_ids"], comm["n_internal"]
    U = U0[gids].copy()
    W = np.zeros_like(U)
    mask = (loc["solved"] & loc["internal"]).astype(np.uint8)
    h = float(m["size"].min())
    t = 0.0
    for _ in range(steps):
        R, me = orc.compute_rhs(loc, U)
        me_t = torch.tensor([me], dtype=torch.float64)
        dist.all_reduce(me_t, op=dist.ReduceOp.MAX)                       # src/main.cpp:393
        dt = orc.choose_dt(0.45, h, float(me_t[0]), t, 1e30)
        orc.rk_stage(loc, 1, dt, U, W, R, mask=mask)
        _exchange(dist, torch, rank, comm, W)                             # :425-430
        R, _ = orc.compute_rhs(loc, W)
        orc.rk_stage(loc, 2, dt, U, W, R, mask=mask)
        _exchange(dist, torch, rank, comm, W)                             # :461-466
        R, _ = orc.compute_rhs(loc, W)
        orc.rk_stage(loc, 3, dt, U, W, R, mask=mask)
        _exchange(dist, torch, rank, comm, U)                             # :497-502
        t += dt
    np.save(os.path.join(out_dir, f"U_{rank}.npy"), U[:n_int])
    np.save(os.path.join(out_dir, f"ids_{rank}.npy"), gids[:n_int])
    np.save(os.path.join(out_dir, f"t_{rank}.npy"), np.array([t]))
    dist.destroy_process_group()


@pytest.mark.parametrize("problem,dim,n", [("vortex_xy", 2, 16), ("radsod", 3, 8)])
def test_two_rank_partitioned_run_equals_serial(oracle, tmp_path, problem, dim, n):
    import torch.multiprocessing as mp
    world, steps = 2, 3
    mp.spawn(_worker, args=(world, _free_port(), problem, dim, n, steps, str(tmp_path)), nprocs=world, join=True)
    m = oracle.problem_mesh(problem, dim, n)
    Uo = oracle.init_state(m)
    Wo, Ro = np.zeros_like(Uo), np.zeros_like(Uo)
    t = 0.0
    for _ in range(steps):
        dt, _ = oracle.step(m, 0.45, t, 1e30, Uo, Wo, Ro)
        t += dt
    got = np.full_like(Uo, np.nan)
    for r in range(world):
        got[np.load(tmp_path / f"ids_{r}.npy")] = np.load(tmp_path / f"U_{r}.npy")
        assert float(np.load(tmp_path / f"t_{r}.npy")[0]) == t
    assert bits_equal(got, Uo)


def test_partition_lists_are_mutually_consistent(oracle):
    """What rank a sends to rank b is exactly what b expects from a, in the same order (global ids)."""
    from minimmerflow_b200.partition import chunk_ranges, morton_chunk_partition
    m = oracle.problem_mesh("radsod", 3, 8)
    world = 3   # uneven chunks: 512 = 171 + 171 + 170 (remainder to the first ranks)
    assert [b - a for a, b in chunk_ranges(512, world)] == [171, 171, 170]
    parts = [morton_chunk_partition(m, world, r) for r in range(world)]
    for a in range(world):
        loc_a, comm_a = parts[a]
        assert loc_a["solved"][comm_a["n_internal"]:].sum() == 0       # ghosts are never solved (src/main.cpp:231-235)
        for q, b in enumerate(comm_a["neighbours"]):
            comm_b = parts[b][1]
            qb = comm_b["neighbours"].index(a)
            sent = comm_a["global_ids"][comm_a["send"][q]]
            expected = comm_b["global_ids"][comm_b["recv"][qb]]
            assert np.array_equal(sent, expected)


def test_box_decomposition_matches_morton_chunks(oracle):
    """For 2/4/8 ranks the equal Morton chunks of a cube are the boxes of box_decomposition (z split
    first, then y, then x), and neighbour ranks are symmetric."""
    from minimmerflow_b200.partition import box_decomposition, box_of_rank, chunk_ranges
    m = oracle.problem_mesh("radsod", 3, 8)
    ijk = m["cell_ijk"]
    for world in (2, 4, 8):
        grid = box_decomposition(world)
        dims = (8 // grid[0], 8 // grid[1], 8 // grid[2])
        for rank, (c0, c1) in enumerate(chunk_ranges(512, world)):
            offset, nbrs = box_of_rank(rank, grid, dims)
            lo, hi = ijk[c0:c1].min(0), ijk[c0:c1].max(0)
            assert tuple(lo) == tuple(offset) and tuple(hi - lo + 1) == dims
            for side, nb in enumerate(nbrs):
                if nb >= 0:
                    assert box_of_rank(nb, grid, dims)[1][side ^ 1] == rank
