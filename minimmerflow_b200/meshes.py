"""Host descriptions of full uniform boxes for the C-ABI (`EulerSolver.from_mesh`), vectorised: what a host
code hands to `mmf_create` when its cells are numbered lexicographically (x fastest) and its interfaces are
created while visiting the cells in that order, faces in the order -x, +x, -y, +y, -z, +z (an interface is created
by the lower of its two cells).  Used by the development tools to build benchmark-sized inputs; the kernels'
parity tests use independent generators.

`with_bodies` sets flags and boundary codes the way the reference does for body boxes (src/main.cpp:221-237,
251-277; src/body.cpp:80-95)."""
import numpy as np


def box_mesh(nx, ny, nz, h, bc_code, origin=(0.0, 0.0, 0.0)):
    nc = nx * ny * nz
    c = np.arange(nc, dtype=np.int64)
    i, j, k = c % nx, (c // nx) % ny, c // (nx * ny)
    ijk = np.stack([i, j, k], axis=1).astype(np.int32)
    dims = (nx, ny, nz)
    stride = (1, nx, nx * ny)
    exists = np.ones((nc, 6), bool)
    neigh = np.full((nc, 6), -1, np.int64)
    for a in range(3):
        exists[:, 2 * a] = ijk[:, a] == 0                       # a low face is created by the lower cell unless on the border
        inner_hi = ijk[:, a] < dims[a] - 1
        neigh[inner_hi, 2 * a + 1] = c[inner_hi] + stride[a]
    sel = exists.ravel()
    owner = np.repeat(c, 6)[sel]
    nb = neigh.ravel()[sel]
    slot = np.tile(np.arange(6), nc)[sel]
    normal = np.zeros((owner.shape[0], 3))
    normal[np.arange(owner.shape[0]), slot // 2] = np.where(slot % 2 == 1, 1.0, -1.0)
    bc = np.where(nb < 0, bc_code, -1).astype(np.int32)
    nf = owner.shape[0]
    cc = np.asarray(origin)[None, :] + (ijk + 0.5) * h
    return dict(dim=3, owner=owner, neigh=nb, bc=bc, area=np.full(nf, h * h), normal=normal,
                icentroid=np.zeros((nf, 3)), volume=np.full(nc, h * h * h), size=np.full(nc, h), ccentroid=cc,
                cell_ijk=ijk, box_dims=(nx, ny, nz), solved=np.ones(nc, np.uint8), internal=np.ones(nc, np.uint8),
                fluid=np.ones(nc, np.uint8), h=h)


def with_bodies(mesh, boxes):
    """Cells whose centroid lies in a closed box (xMin, yMin, zMin, xMax, yMax, zMax) are not solved; interfaces
    between a solved and an unsolved cell get BC_WALL (2); border interfaces keep their code."""
    m = dict(mesh)
    cc = m["ccentroid"]
    solid = np.zeros(cc.shape[0], bool)
    for b in boxes:
        solid |= np.all((cc >= np.array(b[:3])) & (cc <= np.array(b[3:])), axis=1)
    fluid = (~solid).astype(np.uint8)
    bc = m["bc"].copy()
    inner = m["neigh"] >= 0
    o, n = m["owner"][inner], m["neigh"][inner]
    bc[inner] = np.where(fluid[o] != fluid[n], 2, -1)
    m.update(fluid=fluid, solved=fluid.copy(), bc=bc)
    return m


def vortex_state(mesh, gamma=1.4, beta=5.0):
    """Isentropic vortex of src/problem.cpp:252-327 (vortex_xy: axis z, p_inf = T_inf = 1, u_inf = (1, 1, 0)) at
    the cell centroids, conservative AoS [cell][rho, rho u, rho v, rho w, rho E]."""
    x, y = mesh["ccentroid"][:, 0], mesh["ccentroid"][:, 1]
    shape = beta / (2 * np.pi) * np.exp(0.5 * (1 - (x * x + y * y)))
    T = 1.0 - (gamma - 1) / (2 * gamma) * shape * shape
    p = T ** (gamma / (gamma - 1))
    r = p / T
    u, v = 1.0 - y * shape, 1.0 + x * shape
    return np.stack([r, r * u, r * v, 0 * r, p / (gamma - 1) + 0.5 * r * (u * u + v * v)], axis=1)
