// tools/emu/emu_stage.cpp -- DEVELOPMENT TOOL (see shim/cuda_runtime.h): the fiber runtime and a C entry
// point that runs ONE launch of a fused stage kernel, compiled from the product's kernel source, on padded
// host arrays.  tools/emu/run_emu.py drives it and compares with the CPU oracle.
// standard headers that spell attributes the shim's CUDA keywords would rewrite come first
#include <stdint.h>
#include <stdio.h>
#include <time.h>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include <sys/mman.h>

#include <functional>
#include <thread>
#include <vector>

// ---- runtime --------------------------------------------------------------------------------------
namespace emu {

thread_local Lane *tl_cur = nullptr;
uint3e g_blockIdx, g_blockDim, g_gridDim;
int g_chaos = 0;
long long g_spin_limit = 20000000;

struct Cta {
    int n_threads = 0;
    std::atomic<int> bar_count{0};
    std::atomic<unsigned> bar_gen{0};
    std::function<void()> body;
};

void yield_lane()
{
    Lane *me = tl_cur;
    Warp *w = me->warp;
    for (int d = 1; d < 32; ++d) {
        Lane *n = &w->lanes[(me->lane + d) & 31];
        if (!n->done) {
            tl_cur = n;
            swapcontext(&me->ctx, &n->ctx);
            return;
        }
    }
}

void spin_pause(long long &spins, const char *what)
{
    // the watchdog is a wall-clock one: a warp that only keeps hand-shakes alive (kernel form 't': rows beyond the box)
    // spins through a whole plane of its neighbours' arithmetic, however long the host takes for it
    static thread_local struct timespec t0;
    ++spins;
    if ((spins & 63) == 0) sched_yield();
    if (spins == 1) clock_gettime(CLOCK_MONOTONIC, &t0);
    if (spins > 100000) usleep(20);
    if ((spins & 1023) == 0) {
        struct timespec t1;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        const double waited = (double) (t1.tv_sec - t0.tv_sec) + 1e-9 * (double) (t1.tv_nsec - t0.tv_nsec);
        if (waited > (double) g_spin_limit * 3e-6) { // the default limit of 2e7 "yields" = 60 s
            Lane *me = tl_cur;
            fprintf(stderr, "emu: HANG in %s: block (%u,%u,%u) warp %d lane %d gave up after %.0f s\n", what,
                    g_blockIdx.x, g_blockIdx.y, g_blockIdx.z, me->warp->index, me->lane, waited);
            fflush(stderr);
            _exit(3);
        }
    }
}

void chaos_delay()
{
    if (g_chaos <= 0) return;
    Warp *w = tl_cur->warp;
    w->rng = w->rng * 1664525u + 1013904223u;
    const unsigned r = w->rng >> 8;
    if (r % 16 == 0) usleep(r % (unsigned) g_chaos);
    else if (r % 4 == 0) sched_yield();
}

void cta_barrier()
{
    Cta *c = tl_cur->warp->cta;
    const unsigned g = c->bar_gen.load();
    if (c->bar_count.fetch_add(1) + 1 == c->n_threads) {
        c->bar_count.store(0);
        c->bar_gen.fetch_add(1);
    } else {
        long long spins = 0;
        while (c->bar_gen.load() == g) { yield_lane(); spin_pause(spins, "__syncthreads"); }
    }
}

static void lane_entry()
{
    Lane *me = tl_cur;
    me->warp->cta->body();
    me->done = true;
    Warp *w = me->warp;
    for (int d = 1; d < 32; ++d) {
        Lane *n = &w->lanes[(me->lane + d) & 31];
        if (!n->done) { tl_cur = n; setcontext(&n->ctx); }
    }
    // last lane of the warp: returning resumes uc_link = the warp's main context
}

constexpr size_t STACK_BYTES = 256 * 1024;

static void warp_main(Warp *w, char *stacks)
{
    for (int l = 0; l < 32; ++l) {
        Lane &L = w->lanes[l];
        getcontext(&L.ctx);
        L.ctx.uc_stack.ss_sp = stacks + (size_t) l * STACK_BYTES;
        L.ctx.uc_stack.ss_size = STACK_BYTES;
        L.ctx.uc_link = &w->main;
        makecontext(&L.ctx, lane_entry, 0);
    }
    tl_cur = &w->lanes[0];
    swapcontext(&w->main, &w->lanes[0].ctx);
    for (int l = 0; l < 32; ++l) {
        if (!w->lanes[l].done) { fprintf(stderr, "emu: warp %d ended with lane %d unfinished\n", w->index, l); _exit(4); }
    }
}

// runs the CTAs of a grid one after the other; `body` is the kernel call with its arguments bound
static void launch(unsigned gx, unsigned gy, unsigned gz, int n_threads, void *smem, size_t smem_bytes,
                   const std::function<void()> &body, unsigned seed)
{
    const int nw = n_threads / 32;
    static char *stacks = nullptr;
    static size_t stacks_bytes = 0;
    const size_t need = (size_t) nw * 32 * STACK_BYTES;
    if (need > stacks_bytes) {
        if (stacks) munmap(stacks, stacks_bytes);
        stacks = (char *) mmap(nullptr, need, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (stacks == MAP_FAILED) { perror("mmap"); _exit(5); }
        stacks_bytes = need;
    }
    g_gridDim = { gx, gy, gz };
    g_blockDim = { (unsigned) n_threads, 1, 1 };
    std::vector<Warp> warps((size_t) nw);
    for (unsigned bz = 0; bz < gz; ++bz)
        for (unsigned by = 0; by < gy; ++by)
            for (unsigned bx = 0; bx < gx; ++bx) {
                g_blockIdx = { bx, by, bz };
                memset(smem, 0xff, smem_bytes); // shared memory starts as NaNs: reads of unwritten slots show up
                Cta cta;
                cta.n_threads = n_threads;
                cta.body = body;
                for (int wi = 0; wi < nw; ++wi) {
                    Warp &w = warps[wi];
                    memset(w.arrived, 0, sizeof w.arrived);
                    memset(w.departed, 0, sizeof w.departed);
                    w.cta = &cta;
                    w.index = wi;
                    w.rng = seed * 2654435761u + (unsigned) wi * 40503u + bx * 7919u + by * 104729u + bz * 1299709u;
                    for (int l = 0; l < 32; ++l) {
                        Lane &L = w.lanes[l];
                        L.tid = { (unsigned) (wi * 32 + l), 0, 0 };
                        L.warp = &w;
                        L.lane = l;
                        L.gen = 0;
                        L.done = false;
                    }
                }
                std::vector<std::thread> threads;
                for (int wi = 0; wi < nw; ++wi) threads.emplace_back(warp_main, &warps[wi], stacks + (size_t) wi * 32 * STACK_BYTES);
                for (auto &t : threads) t.join();
            }
}

} // namespace emu

// ---- the product's kernel source --------------------------------------------------------------------
namespace mmf {
alignas(128) double smem[40 * 1024]; // the kernels' `extern __shared__ double smem[]` (320 KB: room for variants)
}

#include "uniform_stage_v5r.cuh"
#include "uniform_body_cells.cuh"
#include "uniform_stage_t.cuh"
#include "uniform_eligibility.h"
#include "generic_kernels.cuh"
#include "generic_tables.h"

namespace {

using namespace mmf;

struct Args {
    UniformGeom g;
    const double *Sin, *Un;
    double *Out;
    StepControl *ctl;
    double *max_eig;
    int lz;
    float *cta_est;
    LoadClamp lc;
    HaloWait hw;
    XGhost xg;
    const unsigned char *solid; // form 'b': flag array of a box with bodies
};

// the emulated descriptor of a padded state array for the bulk tensor loads of form 't' (uniform_in_map)
static TmaDesc in_desc(const Args &a, const double *arr, int rows)
{
    TmaDesc d{};
    d.base = arr;
    d.stride[0] = 1; d.stride[1] = a.g.px; d.stride[2] = (long long) a.g.px * a.g.py; d.stride[3] = a.g.fs;
    d.dim[0] = a.g.px; d.dim[1] = a.g.py; d.dim[2] = a.g.pz; d.dim[3] = NF;
    d.box[0] = 32; d.box[1] = rows; d.box[2] = 1; d.box[3] = NF;
    return d;
}

// compact x ghost columns (XG = true instantiations): the 12-warp shapes only, to bound the build time
template <int STAGE, int ORDER>
std::function<void()> bind_kernel_xg(int form, const Args &a)
{
    switch (form) {
    case 'r': return [a] { uniform_stage_kernel_v5r<STAGE, ORDER, 12, true>(a.g, a.Sin, a.Un, a.Out, a.ctl, a.max_eig, a.lz, a.cta_est, a.lc, a.hw, a.xg); };
    default: return nullptr;
    }
}

template <int STAGE, int ORDER, int NW>
std::function<void()> bind_kernel(int form, const Args &a)
{
    if (a.xg.lo || a.xg.hi) return (NW == 12) ? bind_kernel_xg<STAGE, ORDER>(form, a) : nullptr;
    switch (form) {
    case 'r': return [a] { uniform_stage_kernel_v5r<STAGE, ORDER, NW, false>(a.g, a.Sin, a.Un, a.Out, a.ctl, a.max_eig, a.lz, a.cta_est, a.lc, a.hw, a.xg); };
    case 't': return [a] { uniform_stage_kernel_t<STAGE, ORDER, NW, T_DEPTH, false>(a.g, a.Out, a.ctl, a.max_eig, a.lz, a.cta_est, a.lc, a.hw, in_desc(a, a.Sin, NW), in_desc(a, a.Un, NW - 2)); };
    case 'h': return [a] { uniform_stage_kernel_t<STAGE, ORDER, NW, T_DEPTH, true>(a.g, a.Out, a.ctl, a.max_eig, a.lz, a.cta_est, a.lc, a.hw, in_desc(a, a.Sin, NW + 1), in_desc(a, a.Un, NW - 1)); };
    case 'b': return a.solid ? std::function<void()>([a] { uniform_stage_kernel_t<STAGE, ORDER, NW, T_DEPTH, true, true>(a.g, a.Out, a.ctl, a.max_eig, a.lz, a.cta_est, a.lc, a.hw, in_desc(a, a.Sin, NW + 1), in_desc(a, a.Un, NW - 1), a.solid); }) : nullptr;
    default: return nullptr;
    }
}

template <int STAGE, int ORDER>
std::function<void()> bind_nw(int form, int nw, const Args &a)
{
    switch (nw) {
    case 8:  return bind_kernel<STAGE, ORDER, 8>(form, a);
    case 12: return bind_kernel<STAGE, ORDER, 12>(form, a);
    case 16: return bind_kernel<STAGE, ORDER, 16>(form, a);
    default: return nullptr;
    }
}

template <int STAGE>
std::function<void()> bind_order(int order, int form, int nw, const Args &a)
{
    switch (order) {
    case NUM_MORTON: return bind_nw<STAGE, NUM_MORTON>(form, nw, a);
    case NUM_LEXI:   return bind_nw<STAGE, NUM_LEXI>(form, nw, a);
    case NUM_AXIS:   return bind_nw<STAGE, NUM_AXIS>(form, nw, a);
    default: return nullptr;
    }
}

} // namespace

namespace {
template <int STAGE>
double wall_cells_order(int order, const UniformGeom &g, const LoadClamp &lc, const double *Sin, const double *Un,
                        const unsigned char *flag, const int *list, int n_list, double dt, double *compact)
{
    double m = 0.0;
    for (int q = 0; q < n_list; ++q) {
        double out[NF];
        double l;
        if (order == NUM_MORTON)    l = wall_cell_update<STAGE, NUM_MORTON>(g, lc, Sin, Un, flag, list[q], dt, out);
        else if (order == NUM_LEXI) l = wall_cell_update<STAGE, NUM_LEXI>(g, lc, Sin, Un, flag, list[q], dt, out);
        else                        l = wall_cell_update<STAGE, NUM_AXIS>(g, lc, Sin, Un, flag, list[q], dt, out);
        for (int f = 0; f < NF; ++f) compact[(size_t) f * n_list + q] = out[f];
        m = (l < m) ? m : l;
    }
    return m;
}
} // namespace

static mmf::XGhost g_xghost{};
static const unsigned char *g_solid = nullptr;

extern "C" {

// compact x ghost columns [field][k+1][j+1] for the NEXT emu_stage calls (nullptr, nullptr: padded array)
void emu_set_xghost(const double *lo, const double *hi, long long fs, int pitch)
{
    g_xghost.lo = lo; g_xghost.hi = hi; g_xghost.fs = fs; g_xghost.pitch = pitch;
}

// flag array (padded layout of one field, 1 = not solved, 2 = wall cell) for the NEXT emu_stage calls of form 'b'
void emu_set_solid(const unsigned char *solid) { g_solid = solid; }

// padded extents of a box, as uniform_alloc (uniform_path.cuh) lays them out
void emu_padded(const int dims[3], int pad[3], long long *fs)
{
    pad[0] = mmf::padded_row(dims[0]);
    pad[1] = dims[1] + 2;
    pad[2] = dims[2] + 2;
    *fs = ((long long) pad[0] * pad[1] * pad[2] + 15) / 16 * 16;
}

// One launch of a fused stage kernel.  Arrays are padded SoA [field][k+1][j+1][i+1] as on the device.
// smem_doubles: dynamic shared memory the host side would request for this form (checked against the
// tool's static array).  Returns 0, or a negative code for an unknown configuration.
int emu_stage(int form, int stage, int order, int nw, int lz, const int dims[3], const int goff[3], const int gdims[3],
              const int bc[6], double h, double area, double volume, const double dirichlet[5], const int clamp[6],
              const double *Sin, const double *Un, double *Out, double dt, double *max_eig, float *cta_est,
              int smem_doubles, int chaos, unsigned seed)
{
    Args a{};
    UniformGeom &g = a.g;
    g.nx = dims[0]; g.ny = dims[1]; g.nz = dims[2];
    g.gx0 = goff[0]; g.gy0 = goff[1]; g.gz0 = goff[2];
    g.gnx = gdims[0]; g.gny = gdims[1]; g.gnz = gdims[2];
    int pad[3];
    emu_padded(dims, pad, &g.fs);
    g.px = pad[0]; g.py = pad[1]; g.pz = pad[2];
    g.h = h; g.area = area; g.volume = volume;
    for (int s = 0; s < 6; ++s) g.bc[s] = bc[s];
    for (int k = 0; k < NF; ++k) g.dirichlet[k] = dirichlet[k];
    a.lc = { clamp[0], clamp[1], clamp[2], clamp[3], clamp[4], clamp[5] };
    StepControl ctl{};
    ctl.dt = dt;
    ctl.active = 1.0;
    a.ctl = &ctl;
    a.Sin = Sin; a.Un = Un; a.Out = Out;
    a.max_eig = max_eig;
    a.lz = lz;
    a.cta_est = cta_est;
    const int rows = (form == 'h' || form == 'b') ? nw - 1 : nw - 2; // rows a tile updates
    const unsigned gx = (g.nx + XW - 1) / XW, gy = (g.ny + rows - 1) / rows, gz = (g.nz + lz - 1) / lz;
    a.hw = HaloWait{};
    a.hw.tx = (int) gx; a.hw.ty = (int) gy; a.hw.tz = (int) gz;
    a.xg = g_xghost;
    a.solid = g_solid;
    if ((size_t) smem_doubles * sizeof(double) > sizeof(mmf::smem)) return -2;

    std::function<void()> body;
    switch (stage) {
    case 0: body = bind_order<0>(order, form, nw, a); break;
    case 1: body = bind_order<1>(order, form, nw, a); break;
    case 2: body = bind_order<2>(order, form, nw, a); break;
    case 3: body = bind_order<3>(order, form, nw, a); break;
    default: break;
    }
    if (!body) return -1;
    emu::g_chaos = chaos;
    emu::launch(gx, gy, gz, nw * 32, mmf::smem, (size_t) smem_doubles * sizeof(double), body, seed);
    return 0;
}

// max eigenvalue over the processed interfaces of a box with bodies: eig_body_cell (the per-thread body of
// uniform_eig_body_kernel) over all cells, plain loops
double emu_eig_body(const int dims[3], const int bc[6], const double *Sin, const unsigned char *solid)
{
    UniformGeom g{};
    g.nx = dims[0]; g.ny = dims[1]; g.nz = dims[2];
    int pad[3];
    emu_padded(dims, pad, &g.fs);
    g.px = pad[0]; g.py = pad[1]; g.pz = pad[2];
    for (int s = 0; s < 6; ++s) g.bc[s] = bc[s];
    double m = 0.0;
    for (int k = 0; k < g.nz; ++k)
        for (int j = 0; j < g.ny; ++j)
            for (int i = 0; i < g.nx; ++i) {
                const double l = eig_body_cell(g, Sin, solid, i, j, k);
                m = (l < m) ? m : l;
            }
    return m;
}

// kernel form 'b': the wall cells of a box with bodies (wall_cell_update, the per-thread body of
// uniform_wall_cells_kernel) into the compact buffer [field][n_list]; returns the largest interface eigenvalue

double emu_wall_cells(int stage, int order, const int dims[3], const int bc[6], double area, double volume, const int clamp[6],
                      const double *Sin, const double *Un, const unsigned char *flag, const int *list, int n_list, double dt,
                      double *compact)
{
    UniformGeom g{};
    g.nx = dims[0]; g.ny = dims[1]; g.nz = dims[2];
    g.gnx = dims[0]; g.gny = dims[1]; g.gnz = dims[2];
    int pad[3];
    emu_padded(dims, pad, &g.fs);
    g.px = pad[0]; g.py = pad[1]; g.pz = pad[2];
    g.area = area; g.volume = volume;
    for (int s = 0; s < 6; ++s) g.bc[s] = bc[s];
    const LoadClamp lc = { clamp[0], clamp[1], clamp[2], clamp[3], clamp[4], clamp[5] };
    switch (stage) {
    case 0: return wall_cells_order<0>(order, g, lc, Sin, Un, flag, list, n_list, dt, compact);
    case 1: return wall_cells_order<1>(order, g, lc, Sin, Un, flag, list, n_list, dt, compact);
    case 2: return wall_cells_order<2>(order, g, lc, Sin, Un, flag, list, n_list, dt, compact);
    default: return wall_cells_order<3>(order, g, lc, Sin, Un, flag, list, n_list, dt, compact);
    }
}

// body_flags (uniform_device.cuh: the host code uniform_try_create runs for a box with bodies) on a box of the
// given extents; flag_out has room for fs bytes, walls_out for n_cells offsets; returns the number of wall cells
int emu_body_flags(const int dims[3], long long n_cells, const int *cell_ijk, const unsigned char *solved, int mark_walls,
                   unsigned char *flag_out, int *walls_out)
{
    UniformGeom g{};
    g.nx = dims[0]; g.ny = dims[1]; g.nz = dims[2];
    int pad[3];
    emu_padded(dims, pad, &g.fs);
    g.px = pad[0]; g.py = pad[1]; g.pz = pad[2];
    std::vector<unsigned char> flag;
    std::vector<int> walls;
    body_flags(g, n_cells, cell_ijk, solved, mark_walls != 0, flag, walls);
    memcpy(flag_out, flag.data(), flag.size());
    for (size_t q = 0; q < walls.size(); ++q) walls_out[q] = walls[q];
    return (int) walls.size();
}

// analyze_uniform_box (uniform_eligibility.h): the decision mmf_create takes between the fused uniform path
// and the generic one.  out = { numbering, order_exact, bodies, bc_side[6] }; returns 1 if eligible.
int emu_analyze_box(const mmf_mesh_desc *d, int allow_bodies, int out[9])
{
    const UniformBoxAnalysis r = analyze_uniform_box(d, allow_bodies != 0);
    out[0] = r.numbering; out[1] = r.order_exact; out[2] = r.bodies ? 1 : 0;
    for (int s = 0; s < 6; ++s) out[3 + s] = r.bc_side[s];
    return r.eligible ? 1 : 0;
}

// ---- the generic path ------------------------------------------------------------------------------------
// cell -> interface lists of a host mesh description (generic_tables.h: what create_generic uploads).  ptr_out
// has room for n_cells + 1 entries, ent_out for two per interface (either may be null); returns the number of
// entries or -1 (message in err).
long long emu_generic_tables(const mmf_mesh_desc *d, long long *ptr_out, int *ent_out, unsigned char *update_out, char *err,
                             int err_len)
{
    GenericTables t;
    std::string msg;
    if (build_generic_tables(d, t, msg)) {
        if (err && err_len > 0) snprintf(err, (size_t) err_len, "%s", msg.c_str());
        return -1;
    }
    if (ptr_out) for (size_t c = 0; c < t.ptr.size(); ++c) ptr_out[c] = t.ptr[c];
    if (ent_out) for (size_t e = 0; e < t.ent.size(); ++e) ent_out[e] = t.ent[e];
    if (update_out) memcpy(update_out, t.update.data(), t.update.size());
    return (long long) t.ent.size();
}

long long emu_generic_stride(long long n_cells) { return (n_cells + 31) / 32 * 32; }

// Steps of the generic path, the kernel sequence of step_enqueue (mmf_b200.cu) run on the emulator: the
// product's generic kernels, compiled from their source, on SoA arrays [field * stride + cell] the caller
// provides (stride = emu_generic_stride).  fused = 0: residual and RK kernels of every stage; fused = 1: stages 2
// and 3 by generic_stage_kernel with the two work arrays swapping roles -- on return W is whichever array the
// last stage 2 wrote, copied back into the caller's W.  One warp per block here (the block-wide maximum is
// order independent; the tool's `__shared__` is per lane, which a one-warp block never notices).
// Loop: `while (t < t_max)` and at most max_steps (< 0: unbounded) like mmf_run.  Returns the steps taken.
int emu_generic_run(const mmf_mesh_desc *d, int fused, double *U, double *W, double *RHS, double cfl, double min_h,
                    double *t, double t_max, int max_steps, double *dt_last, double eig_last[3])
{
    GenericTables tb;
    std::string msg;
    if (build_generic_tables(d, tb, msg)) return -1;
    if (tb.bc_between_solved) fused = 0; // as create_generic decides
    GenericMesh m{};
    m.n_cells = tb.n_cells; m.n_ifaces = tb.n_ifaces; m.stride = tb.stride;
    m.cf_ptr = tb.ptr.data(); m.cf_ent = tb.ent.data();
    m.f_owner = tb.owner.data(); m.f_neigh = tb.neigh.data(); m.f_bc = tb.bc.data();
    m.f_area = tb.area.data(); m.f_normal = tb.normal.data();
    m.c_solved = tb.solved.data(); m.c_update = tb.update.data(); m.c_volume = tb.volume.data();
    memcpy(m.dirichlet_info, d->dirichlet_info, sizeof m.dirichlet_info);

    StepControl ctl{};
    ctl.t = *t; ctl.t_max = t_max; ctl.cfl = cfl; ctl.min_h = min_h;
    std::vector<double> alt((size_t) NF * m.stride, 0.0);
    double *Wc = W, *Wa = alt.data();
    const unsigned grid = (unsigned) ((m.n_cells + 31) / 32);
    emu::g_chaos = 0;
    auto run = [&](const std::function<void()> &body) { emu::launch(grid, 1, 1, 32, mmf::smem, 0, body, 1u); };
    auto rhs = [&](const double *S, int slot) {
        ctl.max_eig[slot] = 0.0;
        double *mx = &ctl.max_eig[slot];
        if (fused) run([=] { generic_rhs_derived_kernel<>(m, S, RHS, mx); });
        else       run([=] { generic_rhs_kernel(m, S, RHS, mx); });
    };
    StepControl *c = &ctl;
    int steps = 0;
    while (ctl.t < t_max && (max_steps < 0 || steps < max_steps)) {
        rhs(U, 0);
        choose_dt_kernel(c);
        run([=] { generic_rk_kernel<1>(m.n_cells, m.stride, m.c_update, m.c_volume, c, U, Wc, RHS); });
        if (fused) {
            ctl.max_eig[1] = 0.0;
            { double *mx = &ctl.max_eig[1]; double *Si = Wc, *So = Wa; run([=] { generic_stage_kernel<2>(m, Si, U, So, RHS, c, mx); }); }
            std::swap(Wc, Wa);
            ctl.max_eig[2] = 0.0;
            { double *mx = &ctl.max_eig[2]; double *Si = Wc; run([=] { generic_stage_kernel<3>(m, Si, U, U, RHS, c, mx); }); }
        } else {
            rhs(Wc, 1);
            run([=] { generic_rk_kernel<2>(m.n_cells, m.stride, m.c_update, m.c_volume, c, U, Wc, RHS); });
            rhs(Wc, 2);
            run([=] { generic_rk_kernel<3>(m.n_cells, m.stride, m.c_update, m.c_volume, c, U, Wc, RHS); });
        }
        advance_time_kernel(c, 0);
        ++steps;
    }
    if (Wc != W) memcpy(W, Wc, sizeof(double) * (size_t) NF * m.stride);
    *t = ctl.t;
    if (dt_last) *dt_last = ctl.dt;
    if (eig_last) for (int i = 0; i < 3; ++i) eig_last[i] = ctl.max_eig[i];
    return steps;
}

void emu_set_spin_limit(long long n) { emu::g_spin_limit = n; }

// phase bookkeeping of the emulated mbarrier word (no fibers involved: arrivals only, waits inspected)
int emu_mbar_selftest()
{
    auto waiting = [](unsigned long long w, unsigned parity) { return ((w >> 32) & 1u) == (parity & 1u); };
    unsigned long long bar = 0;
    emu::Lane lane{};
    emu::Warp warp{};
    lane.warp = &warp;
    emu::tl_cur = &lane; // chaos_delay() looks at the current lane's warp
    const int chaos = emu::g_chaos;
    emu::g_chaos = 0;
    int bad = 0;
    mmf::mbar_init(&bar, 2);
    if (!waiting(bar, 0)) bad |= 1;          // phase 0 still running: a wait on parity 0 blocks
    mmf::mbar_arrive(&bar);
    if (!waiting(bar, 0)) bad |= 2;          // one of two arrivals: still blocked
    mmf::mbar_arrive(&bar);
    if (waiting(bar, 0)) bad |= 4;           // phase 0 complete: released
    if (!waiting(bar, 1)) bad |= 8;          // phase 1 running
    mmf::mbar_arrive(&bar);
    mmf::mbar_arrive(&bar);
    if (waiting(bar, 1)) bad |= 16;          // phase 1 complete
    if (!waiting(bar, 0)) bad |= 32;         // ... and a late waiter of phase 0 would now block for good
    emu::g_chaos = chaos;
    emu::tl_cur = nullptr;
    return bad;
}

} // extern "C"
