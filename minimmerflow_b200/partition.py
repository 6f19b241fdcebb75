"""Host-side domain decomposition for the multi-GPU runs (one process per GPU).

Two decompositions, both following the reference's `mesh.partition(false, true)` (src/main.cpp:159,
183 -> PABLO: equal contiguous chunks of the Morton-ordered cell list, remainder to the first ranks):

* :func:`morton_chunk_partition` -- generic path.  From a global mesh description it builds each
  rank's local mesh (interior chunk + one layer of face-neighbour ghost cells, local raw ids) and the
  per-neighbour send / receive lists, in the role of `getGhostCellExchangeSources/Targets`
  (src/communications.cpp:621-630).  Ghost cells are flagged not-internal, hence not solved
  (src/main.cpp:231-235).  The relative order of interfaces is preserved, so every interior cell
  accumulates its faces in the same order as in the serial run and results stay bitwise identical.
* :func:`box_decomposition` / :func:`box_of_rank` -- uniform path.  For 1/2/4/8 ranks the Morton
  chunks of a cube are boxes (halves split z, quarters z then y, octants z, y, x, because x is the
  lowest Morton bit); each rank owns one box and exchanges whole face layers.

Pure numpy index manipulation; no numerics.
"""
import numpy as np


def chunk_ranges(n_cells, n_ranks):
    """Equal contiguous chunks, remainder to the first ranks."""
    base, rem = divmod(int(n_cells), int(n_ranks))
    starts = [0]
    for r in range(n_ranks):
        starts.append(starts[-1] + base + (1 if r < rem else 0))
    return [(starts[r], starts[r + 1]) for r in range(n_ranks)]


def morton_chunk_partition(mesh, n_ranks, rank):
    """Local mesh of `rank` plus exchange lists.

    mesh: global description (owner, neigh, bc, area, normal, volume, solved, optional icentroid,
    ccentroid, size); cells are assumed to be stored in partition (Morton) order.
    Returns (local_mesh, comm) where comm = dict(neighbours=[ranks], send=[local ids per neighbour],
    recv=[local ids per neighbour], global_ids=array of the local cells' global ids, n_internal=int).
    """
    owner, neigh = np.asarray(mesh["owner"]), np.asarray(mesh["neigh"])
    nc = int(np.asarray(mesh["volume"]).shape[0])
    ranges = chunk_ranges(nc, n_ranks)
    c0, c1 = ranges[rank]
    bounds = np.array([r[0] for r in ranges] + [nc])

    def rank_of(cells):
        return np.searchsorted(bounds, cells, side="right") - 1

    own_o = (owner >= c0) & (owner < c1)
    own_n = (neigh >= c0) & (neigh < c1)
    keep = own_o | own_n                      # every interface touching an interior cell
    f_ids = np.nonzero(keep)[0]               # ascending -> relative order preserved

    # ghost cells: the other side of kept interfaces
    other = np.concatenate([owner[f_ids][~own_o[f_ids]], neigh[f_ids][(~own_n[f_ids]) & (neigh[f_ids] >= 0)]])
    ghosts = np.unique(other)
    global_ids = np.concatenate([np.arange(c0, c1), ghosts])
    n_int = c1 - c0
    lookup = {int(g): n_int + q for q, g in enumerate(ghosts)}

    def to_local(cells):
        out = np.empty(cells.shape, np.int64)
        for q, c in enumerate(cells):
            c = int(c)
            out[q] = -1 if c < 0 else (c - c0 if c0 <= c < c1 else lookup[c])
        return out

    internal = np.zeros(global_ids.shape[0], np.uint8)
    internal[:n_int] = 1
    solved_g = np.asarray(mesh["solved"])[global_ids].astype(np.uint8)
    local = dict(dim=mesh["dim"], owner=to_local(owner[f_ids]), neigh=to_local(neigh[f_ids]),
                 bc=np.asarray(mesh["bc"])[f_ids].copy(), area=np.asarray(mesh["area"])[f_ids].copy(),
                 normal=np.asarray(mesh["normal"])[f_ids].copy(), volume=np.asarray(mesh["volume"])[global_ids].copy(),
                 internal=internal, solved=(solved_g & internal).astype(np.uint8))   # src/main.cpp:231-235
    for key in ("icentroid",):
        if key in mesh:
            local[key] = np.asarray(mesh[key])[f_ids].copy()
    for key in ("ccentroid", "size", "fluid"):
        if key in mesh:
            local[key] = np.asarray(mesh[key])[global_ids].copy()
    if "problem" in mesh:
        local["problem"] = mesh["problem"]
    if "h" in mesh:
        local["h"] = mesh["h"]

    # exchange lists: what I receive = my ghosts grouped by owner rank (ascending global id);
    # what I send to rank q = my interior cells that are ghosts of q = interior cells adjacent,
    # through a kept interface, to a cell of q (ascending global id) -- both sides derive the same
    # ordering independently, like bitpit's sorted exchange lists
    ghost_rank = rank_of(ghosts)
    neighbours = sorted(set(int(r) for r in ghost_rank))
    recv = [np.array([lookup[int(g)] for g in ghosts[ghost_rank == q]], np.int64) for q in neighbours]
    send = []
    fo, fn = owner[f_ids], neigh[f_ids]
    for q in neighbours:
        q0, q1 = ranges[q]
        mine = np.concatenate([fo[own_o[f_ids] & (fn >= q0) & (fn < q1)], fn[own_n[f_ids] & (fo >= q0) & (fo < q1)]])
        send.append(np.unique(mine) - c0)
    comm = dict(neighbours=neighbours, send=send, recv=recv, global_ids=global_ids, n_internal=n_int)
    return local, comm


def box_decomposition(n_ranks):
    """Process grid (px, py, pz) in which the Morton chunks of a cube are boxes."""
    grids = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}
    if n_ranks not in grids:
        raise ValueError("box decomposition is defined for 1, 2, 4 or 8 ranks")
    return grids[n_ranks]


def box_of_rank(rank, grid, box_dims):
    """(offset, neighbour ranks for -x,+x,-y,+y,-z,+z) of `rank`; ranks are numbered x fastest."""
    px, py, pz = grid
    cx, cy, cz = rank % px, (rank // px) % py, rank // (px * py)
    offset = (cx * box_dims[0], cy * box_dims[1], cz * box_dims[2])

    def nb(dx, dy, dz):
        x, y, z = cx + dx, cy + dy, cz + dz
        if not (0 <= x < px and 0 <= y < py and 0 <= z < pz):
            return -1
        return (z * py + y) * px + x

    return offset, [nb(-1, 0, 0), nb(1, 0, 0), nb(0, -1, 0), nb(0, 1, 0), nb(0, 0, -1), nb(0, 0, 1)]
