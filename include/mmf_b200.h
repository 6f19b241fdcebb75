/*
 * mmf_b200.h -- C-ABI of libmmf_b200.so: the B200 (sm_100a) implementation of minimmerflow's
 * explicit finite-volume Euler residual-and-update path.
 *
 * The reference has no plugin/FFI layer; its de-facto operator API is the set of free functions
 * that src/main.cpp calls inside the time loop.  Every entry point below names the reference
 * interface it replaces (paths relative to the reference tree).  INTEGRATION.md shows the C++
 * adapters a maintainer adds so that main.cpp keeps calling the reference signatures.
 *
 * Conventions
 *   - extern "C", opaque handle, int status return (0 = MMF_OK), no exceptions cross the boundary;
 *     mmf_last_error() returns the message of the last failure.
 *   - Host buffers are caller-owned and use the reference's layout: AoS, value (raw cell c,
 *     field k) at [c*5 + k] (bitpit::PiercedStorage<double,long> with 5 fields,
 *     src/storage.hpp:31-41, src/constants.hpp:35-52).  Device buffers are library-owned (SoA FP64).
 *   - Cells and interfaces are addressed by the host's RAW ids; connectivity, geometry, flags and
 *     the interface processing order are INPUTS, so indexing is identical to the host's by
 *     construction.
 *   - One handle drives one GPU (one process per GPU; see mmf_comm_init for multi-GPU).
 *     Calls on a handle must be serialised by the caller (the reference is single-threaded).
 *   - There is no CPU fallback: every compute entry point fails with MMF_ERR_NO_DEVICE when no
 *     sm_100 device is usable.
 */
#ifndef MMF_B200_H
#define MMF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMF_N_FIELDS 5 /* src/constants.hpp:35 */

/* status codes */
enum {
    MMF_OK                    = 0,
    MMF_ERR_INVALID           = 1, /* bad argument / inconsistent mesh description            */
    MMF_ERR_CUDA              = 2, /* a CUDA runtime call failed                               */
    MMF_ERR_UNSUPPORTED_ORDER = 3, /* order != 1: the reference calls exit(2)                  */
                                   /* (src/reconstruction.cpp:76); adapters map this to exit(2) */
    MMF_ERR_NO_DEVICE         = 4, /* no usable sm_100 GPU: there is no CPU fallback           */
    MMF_ERR_NCCL              = 5,
    MMF_ERR_STATE             = 6  /* call sequence error (e.g. step before set_state)         */
};

/* boundary-condition codes: src/constants.hpp:58-62 (values are part of the contract) */
enum {
    MMF_BC_NONE       = -1,
    MMF_BC_FREE_FLOW  =  0,
    MMF_BC_REFLECTING =  1,
    MMF_BC_WALL       =  2,
    MMF_BC_DIRICHLET  =  3
};

/* cell fields held on the device; mirror the storages allocated at src/main.cpp:216-219 */
enum {
    MMF_FIELD_U   = 0, /* cellConservatives     */
    MMF_FIELD_W   = 1, /* cellConservativesWork */
    MMF_FIELD_RHS = 2  /* cellRHS               */
};

/* execution paths (reported by mmf_get_info) */
enum {
    MMF_PATH_GENERIC = 0, /* connectivity-driven cell-gather kernels, any mesh                 */
    MMF_PATH_UNIFORM = 1  /* fused brick kernels for a full uniform 3-D box                    */
};

/* flags for mmf_mesh_desc.flags */
enum {
    MMF_FLAG_FORCE_GENERIC = 1u << 0, /* never select the uniform fast path                    */
    MMF_FLAG_ORDER_AXIS    = 1u << 1  /* uniform path: accumulate faces in fixed axis order     */
                                      /* instead of the host's interface-id order              */
};

/*
 * Mesh description = what MeshGeometricalInfo::_extract caches (src/mesh_info.cpp:86-118) plus the
 * flag / BC tables main.cpp builds (src/main.cpp:221-237, 251-277).  All arrays are indexed by RAW
 * id and are only read during mmf_create (the library copies what it needs).
 */
typedef struct mmf_mesh_desc {
    size_t   struct_size;        /* = sizeof(mmf_mesh_desc), for ABI versioning                 */
    int32_t  dim;                /* 2 or 3 (MeshGeometricalInfo::getDimension)                   */
    int32_t  problem_type;       /* problem::ProblemType (src/problem.hpp:33-42); informational  */
    uint32_t flags;              /* MMF_FLAG_*                                                  */
    int32_t  reserved0;

    int64_t  n_cells;            /* number of raw cell slots (interior + ghost)                 */
    int64_t  n_interfaces;       /* number of raw interface slots                               */

    /* interface processing order = MeshGeometricalInfo::getInterfaceRawIds()
     * (src/euler.cpp:153); NULL means 0..n_interfaces-1 */
    const int64_t *interface_order;
    int64_t        n_interfaces_listed;

    /* per interface (raw id): Interface::getOwner/getNeigh mapped to raw cell ids
     * (src/euler.cpp:158-178); neigh < 0 for a border interface */
    const int64_t *owner;
    const int64_t *neigh;
    const int32_t *bc;           /* interfaceBCs (src/main.cpp:251-277), MMF_BC_*               */
    const double  *area;         /* rawGetInterfaceArea                                         */
    const double  *normal;       /* rawGetInterfaceNormal, AoS [f*3+e]                          */

    /* per cell (raw id) */
    const double  *volume;       /* rawGetCellVolume                                            */
    const uint8_t *solved;       /* cellSolvedFlag, copied element-wise (PiercedStorage<bool>)  */
    const uint8_t *internal;     /* 1 if in getInternalCellRawIds(); NULL = all internal        */

    /* BC_DIRICHLET data in primitive order {p,u,v,w,T} (problem::getBorderBCInfo,
     * src/problem.cpp:450-477) */
    double dirichlet_info[MMF_N_FIELDS];

    /* Optional structured hint enabling MMF_PATH_UNIFORM: integer lattice coordinates of every
     * cell inside this rank's box, AoS [c*3+d]; NULL = generic path.  The library verifies that
     * the description really is a full, conforming, uniform, all-solved 3-D box before using it. */
    const int32_t *cell_ijk;
    int32_t        box_dims[3];     /* local box size in cells                                   */
    int32_t        global_dims[3];  /* global lattice size (== box_dims on one GPU)              */
    int32_t        box_offset[3];   /* lattice coordinate of this box's first cell               */
    int32_t        reserved1;
} mmf_mesh_desc;

/* how the host numbers the cells / interfaces of a uniform box */
enum {
    MMF_NUMBERING_MORTON        = 0, /* bitpit VolOctree / PABLO: Z-order, x lowest interleaved bit; */
                                     /* interfaces created while visiting cells in that order        */
    MMF_NUMBERING_LEXICOGRAPHIC = 1, /* x fastest, then y, then z; interfaces created likewise       */
    MMF_NUMBERING_AXIS          = 2  /* accumulation order only: fixed -x,+x,-y,+y,-z,+z             */
};

/*
 * Compact description of a full uniform 3-D box (what `VolOctree mesh(3, origin, length, dh)` of
 * src/main.cpp:146-150 builds) for meshes too large to describe interface by interface
 * (256^3 cells = 50.5 M interfaces).  Connectivity is implied by the numbering convention.
 */
typedef struct mmf_uniform_desc {
    size_t   struct_size;          /* = sizeof(mmf_uniform_desc)                                   */
    int32_t  problem_type;
    uint32_t flags;
    int32_t  box_dims[3];          /* cells of this rank's box                                     */
    int32_t  global_dims[3];       /* global lattice                                               */
    int32_t  box_offset[3];        /* lattice coordinate of the box's first cell                   */
    int32_t  cell_numbering;       /* raw cell id <-> lattice coordinate inside the box            */
    int32_t  interface_numbering;  /* decides the per-cell face accumulation order                 */
    int32_t  bc_side[6];           /* MMF_BC_* on the global -x,+x,-y,+y,-z,+z sides               */
    double   h;                    /* cell size                                                    */
    double   dirichlet_info[MMF_N_FIELDS];
    /* What meshInfo.rawGetInterfaceArea / rawGetCellVolume return for this mesh (src/mesh_info.cpp:86-118).
     * 0 = built as h*h and h*h*h, which equals a host's own evalInterfaceArea / evalCellVolume only to
     * the last ulp: a host that wants the bits of ITS geometry passes its two values here.  A caller
     * compiled against the header without these two fields (struct_size = offset of `area`) gets 0.  */
    double   area;
    double   volume;
} mmf_uniform_desc;

typedef struct mmf_ctx mmf_ctx;

typedef struct mmf_info {
    int32_t path;              /* MMF_PATH_*                                                     */
    int32_t device;            /* CUDA device ordinal                                            */
    int32_t sm_count;
    int32_t cc_major, cc_minor;
    int32_t order_exact;       /* 1 if face accumulation follows the host's interface-id order   */
    int64_t n_cells, n_interfaces;
    int64_t kernel_launches;   /* number of library kernels launched so far on this handle       */
    int64_t device_bytes;      /* device memory owned by the handle                              */
} mmf_info;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* Replaces: MeshGeometricalInfo construction + storage allocation (src/main.cpp:197, 214-219). */
int mmf_create(const mmf_mesh_desc *desc, int device, mmf_ctx **out);
/* Same, from the compact uniform-box description (always MMF_PATH_UNIFORM). */
int mmf_create_uniform(const mmf_uniform_desc *desc, int device, mmf_ctx **out);
int mmf_destroy(mmf_ctx *ctx);
const char *mmf_last_error(const mmf_ctx *ctx); /* ctx may be NULL (error from mmf_create)        */
int mmf_get_info(const mmf_ctx *ctx, mmf_info *info);
int mmf_device_count(void); /* number of usable sm_100 devices (0 on a CPU-only box)            */

/* ---- state transfer (host AoS raw order <-> device SoA) ------------------------------------ */
/* Replaces: direct rawData() access to cellConservatives / cellConservativesWork / cellRHS.     */
int mmf_set_state(mmf_ctx *ctx, int field, const double *host_aos);
int mmf_get_state(mmf_ctx *ctx, int field, double *host_aos);
/* Replaces: the per-cell utils::conservative2primitive loop main.cpp runs before every mesh.write()
 * (src/main.cpp:511-518, :531-538; src/utils.cpp:48-63): the primitive fields {p,u,v,w,T}
 * (src/constants.hpp:37-47) of conservative field `field`, evaluated on the device with IEEE
 * divisions, AoS in raw cell order like cellPrimitives.rawData(0).  Every cell is converted,
 * solved or not, as in the reference.                                                            */
int mmf_get_primitives(mmf_ctx *ctx, int field, double *host_aos);

/* ---- operators with the reference call shape ----------------------------------------------- */
/* reconstruction::computePolynomials (src/reconstruction.hpp:41-42, reconstruction.cpp:47-55):
 * a no-op at order 1, kept so the call sequence of main.cpp:388,432,468 is preserved.            */
int mmf_compute_polynomials(mmf_ctx *ctx, int field);

/* euler::computeRHS (src/euler.hpp:41-43, src/euler.cpp:127-249): residual of device field
 * `field` into the device RHS field, *max_eig = max face eigenvalue.                            */
int mmf_compute_rhs(mmf_ctx *ctx, int field, int order, double *max_eig);

/* Strict drop-in for euler::computeRHS with HOST storages: uploads cons_aos, computes, downloads
 * rhs_aos.  PCIe-bound by construction; for adapters that keep main.cpp's RK loops on the host. */
int mmf_compute_rhs_host(mmf_ctx *ctx, const double *cons_aos, int order, double *rhs_aos, double *max_eig);

/* The RK loops written inline in main.cpp: stage 1 (:409-423) W = U + dt*RHS/V,
 * stage 2 (:445-459) W = 0.75*U + 0.25*(W + dt*RHS/V), stage 3 (:481-495)
 * U = (1./3)*U + (2./3)*(W + dt*RHS/V); internal AND solved cells only.                          */
int mmf_rk_stage(mmf_ctx *ctx, int stage, double dt);

/* One whole SSP-RK3 step, device resident (replaces src/main.cpp:383-502 by one call):
 * dt = 0.9*cfl*min_cell_size/maxEig(stage 1), clamped so that t+dt <= t_max (:398-402).
 * max_eig_out[3] receives the three per-stage values main.cpp logs (:399, :440, :476).          */
int mmf_step(mmf_ctx *ctx, double cfl, double min_cell_size, double t, double t_max,
             double *dt_out, double max_eig_out[3]);

/* The whole `while (t < tMax)` loop (src/main.cpp:377-524 without the VTK branch), device
 * resident with no host round trip per step: runs until *t >= t_max or max_steps steps
 * (max_steps < 0: unlimited).  Updates *t, returns the number of steps in *steps_out.            */
int mmf_run(mmf_ctx *ctx, double cfl, double min_cell_size, double *t, double t_max,
            int max_steps, int *steps_out);

/* ---- multi-GPU: one process per GPU -------------------------------------------------------- */
/* Replaces GhostCommunicator construction (src/main.cpp:307-328): joins an NCCL communicator.
 * nccl_unique_id is the 128-byte ncclUniqueId created by mmf_comm_unique_id on rank 0 and
 * broadcast by the host (torch.distributed / MPI).                                              */
int mmf_comm_unique_id(void *id_out_128);
int mmf_comm_init(mmf_ctx *ctx, int rank, int n_ranks, const void *nccl_unique_id);
/* Ghost lists for the generic path: per neighbour rank the raw ids to send / receive, in the
 * order of GhostCommunicator's exchange lists (src/communications.cpp:621-630).                  */
int mmf_comm_set_ghost_lists(mmf_ctx *ctx, int n_neighbours, const int32_t *neighbour_ranks,
                             const int64_t *send_offsets, const int64_t *send_ids,
                             const int64_t *recv_offsets, const int64_t *recv_ids);
/* Uniform path: ranks owning the boxes across the -x,+x,-y,+y,-z,+z sides (-1 = physical border). */
int mmf_comm_set_box_neighbours(mmf_ctx *ctx, const int32_t neighbour_ranks[6]);
/* Uniform path, optional: exchange halos by DIRECT PEER STORES over NVLink instead of NCCL
 * send/recv + pack/unpack (the role of the send / receive buffers of bitpit's DataCommunicator,
 * src/communications.cpp:282-322, disappears).  Every rank exports MMF_IPC_BLOB_BYTES bytes (CUDA IPC
 * handles of its state arrays, arrival counters and x ghost columns), the host gathers the blobs of all ranks in rank
 * order (MPI_Allgather / torch.distributed) and hands the concatenation to every rank.  Requires one
 * process per GPU on one node, equal box dimensions on all ranks, and mmf_comm_set_box_neighbours
 * before the import.  mmf_destroy of such a handle is COLLECTIVE: all ranks must call it (a neighbour
 * may still be storing into this rank's ghost cells), like freeing a communicator. */
#define MMF_IPC_BLOB_BYTES 512
int mmf_comm_ipc_export(mmf_ctx *ctx, void *blob_out);
int mmf_comm_ipc_import(mmf_ctx *ctx, const void *all_ranks_blobs);
/* startAllExchanges + completeAllExchanges (src/communications.cpp:375-470; call sites
 * src/main.cpp:427-428, 463-464, 499-500): refresh the ghost cells of `field`.                  */
int mmf_exchange(mmf_ctx *ctx, int field);
/* MPI_Allreduce(MAX) of main.cpp:393 -- exposed for adapters that keep the host loop.           */
int mmf_allreduce_max(mmf_ctx *ctx, double *value);

/* ---- measurement helpers (CUDA events on the handle's own stream) -------------------------- */
int mmf_timer_start(mmf_ctx *ctx);
int mmf_timer_stop(mmf_ctx *ctx, float *milliseconds);
int mmf_synchronize(mmf_ctx *ctx);
/* Per-launch timing of the residual kernels: between begin and end every residual launch is
 * bracketed by CUDA events on the launching stream.  end returns, per kernel kind (0 = RHS only /
 * generic gather, 1..3 = fused RK stage kernels), the summed device time and the launch count. */
int mmf_profile_begin(mmf_ctx *ctx);
int mmf_profile_end(mmf_ctx *ctx, double total_ms[4], int64_t launches[4]);
/* Writes a scratch buffer larger than L2 on the handle's stream (L2 flush between timed runs).  */
int mmf_flush_l2(mmf_ctx *ctx);
/* GPU self-test of the shared-reciprocal division used by the kernels: counts bitwise mismatches against IEEE `/`
 * over ~n_samples random operand pairs AND a directed sweep (structured mantissas for numerator and denominator --
 * all zeros / ones, single bits, ends of the range --, exponents up to 2^+-300 on either side, both signs, products
 * b*q next to representable quotients, a = +0); must be 0. */
int mmf_selftest_division(int device, long long n_samples, unsigned long long seed, unsigned long long *mismatches);
/* Pinned host memory for the e2e leg. */
int mmf_host_alloc(void **ptr, size_t bytes);
int mmf_host_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif
