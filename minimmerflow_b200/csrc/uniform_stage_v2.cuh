// uniform_stage_v2.cuh -- second generation of the fused residual + RK-stage kernel.
//
// Same decomposition and arithmetic as uniform_stage_kernel (uniform_kernels.cuh) -- a warp owns a
// 32-cell x window (30 updated), the CTA's warps are consecutive y rows with two halo rows, the
// CTA marches along z keeping plane k-1 in registers -- but restructured after the first ncu
// capture (profiles/r01a_*), which showed the kernel issue-bound: ~835 thread instructions per
// cell of which only ~300 FP64, because of 64-bit shuffles + register-pairing moves, select-heavy
// masking and register-rotation copies.  Changes:
//   * neighbour records travel through shared memory as 16-byte pairs (LDS.128 / STS.128): the
//     same record serves the +x neighbour (same warp: only __syncwarp) and the +y neighbour;
//   * the 0.5 of the LLF average is folded into the area factor (exact: scaling by 2 commutes
//     with rounding);
//   * domain-border cells take a rare, separate accumulation path; the common path decides the
//     Morton creation order with one warp-uniform branch and one per-lane select;
//   * the z loop is unrolled twice over two register sets (no rotation copies) and the next
//     plane is loaded straight into the set that just retired.
// Results are bit-identical to v1 and to the oracle (tests/test_uniform_gpu.py).
#pragma once

#include "uniform_kernels.cuh"

namespace mmf {

struct PlaneRegs {
    double U[NF];   // residual-input state of the plane
    double Fz[NF];  // z flux of the cell
    double lz;      // |w| + a
    double S[NF];   // partial RHS: everything except the +z face
    double Un[NF];  // U^n of the plane (stages 2, 3)
};

// record layout per row: 9 double2 slots x 32 lanes
//   0:(U0,U1) 1:(U2,U3) 2:(U4,lx) 3:(Fx0,Fx1) 4:(Fx2,Fx3) 5:(Fx4,ly) 6:(Fy0,Fy1) 7:(Fy2,Fy3) 8:(Fy4,-)
constexpr int REC_SLOTS = 9;
constexpr int FLX_SLOTS = 3; // (AF0,AF1) (AF2,AF3) (AF4,-)

__device__ __forceinline__ double llf_half_area_flux(const double *UL, const double *FL, double lamL,
                                                     const double *UR, const double *FR, double lamR,
                                                     double Ah, double *AF)
{
    const double lam = (lamR < lamL) ? lamL : lamR;
#pragma unroll
    for (int k = 0; k < NF; ++k) {
        // A*(0.5*x) == (0.5*A)*x exactly
        AF[k] = Ah * ((FR[k] + FL[k]) - lam * (UR[k] - UL[k]));
    }
    return lam;
}

template <int STAGE, int ORDER, int NW>
struct StageV2 {
    const UniformGeom &g;
    const double *__restrict__ Sin;
    const double *Un;
    double *Out;
    double2 *rec, *fx, *fy;
    int lane, row, z0, z1;
    bool upd_row, upd, xf_ok, yf_ok, zf_ok, blo_x, blo_y;
    int key_x, key_y;
    long long col, plane, fs;
    double Ah, dt;
    DivConsts dc;
    double lmax;

    __device__ __forceinline__ void load_plane(double *dst, const double *base, int kz) const
    {
        const double *p = base + col + (long long) (kz + 1) * plane;
#pragma unroll
        for (int k = 0; k < NF; ++k) dst[k] = p[k * fs];
    }

    // one plane: `cur` holds the loaded state of plane kz, `prev` the finished derived data of kz-1
    __device__ __forceinline__ void step(PlaneRegs &prev, PlaneRegs &cur, const int kz)
    {
        // U^n of plane kz-1 is consumed ~150 instructions further down (stages 2, 3)
        if (STAGE >= 2 && upd && kz > z0) load_plane(prev.Un, Un, kz - 1);

        CellPrim q;
        derive_cell(cur.U, dc, q);
        axis_flux<2>(q, cur.Fz, cur.lz);

        // ---- z interface (kz-1 | kz) and completion of cell (i,j,kz-1) --------------------------
        double AFz[NF];
        if (upd_row && kz >= z0) {
            const double lam = llf_half_area_flux(prev.U, prev.Fz, prev.lz, cur.U, cur.Fz, cur.lz, Ah, AFz);
            if (zf_ok) lmax = (lam < lmax) ? lmax : lam;
            if (upd && kz > z0) {
                double *op = Out + col + (long long) kz * plane; // plane kz-1
#pragma unroll
                for (int k = 0; k < NF; ++k) {
                    const double rhs = prev.S[k] - AFz[k];
                    double out;
                    if (STAGE == 0) {
                        out = rhs;
                    } else {
                        const double dq = div_nr(dt * rhs, g.volume, dc.y_vol); // dt * RHS[k] / cellVolume
                        if (STAGE == 1)      out = prev.U[k] + dq;
                        else if (STAGE == 2) out = 0.75 * prev.Un[k] + 0.25 * (prev.U[k] + dq);
                        else                 out = (1. / 3) * prev.Un[k] + (2. / 3) * (prev.U[k] + dq);
                    }
                    op[k * fs] = out;
                }
            }
        }
        if (kz == z1) return;

        // ---- publish this cell's record -----------------------------------------------------------
        double cFx[NF], clx;
        {
            double cFy[NF], cly;
            axis_flux<1>(q, cFy, cly);
            double2 *d = rec + (row * REC_SLOTS) * 32 + lane;
            d[0 * 32] = make_double2(cur.U[0], cur.U[1]);
            d[1 * 32] = make_double2(cur.U[2], cur.U[3]);
            if (upd_row) {
                axis_flux<0>(q, cFx, clx);
                d[2 * 32] = make_double2(cur.U[4], clx);
                d[3 * 32] = make_double2(cFx[0], cFx[1]);
                d[4 * 32] = make_double2(cFx[2], cFx[3]);
                d[5 * 32] = make_double2(cFx[4], cly);
            } else {
                d[2 * 32] = make_double2(cur.U[4], 0.0);
                d[5 * 32] = make_double2(0.0, cly);
            }
            d[6 * 32] = make_double2(cFy[0], cFy[1]);
            d[7 * 32] = make_double2(cFy[2], cFy[3]);
            d[8 * 32] = make_double2(cFy[4], 0.0);
        }

        // the retired register set receives the next plane now: the loads fly during the exchange
        load_plane(prev.U, Sin, kz + 1);

        // ---- x direction: same warp, shared-memory records, no block barrier ---------------------
        double AFxhi[NF], AFxlo[NF];
        if (upd_row && kz >= z0) {
            __syncwarp();
            const double2 *d = rec + (row * REC_SLOTS) * 32 + ((lane + 1) & 31);
            const double2 a0 = d[0 * 32], a1 = d[1 * 32], a2 = d[2 * 32], a3 = d[3 * 32], a4 = d[4 * 32], a5 = d[5 * 32];
            const double nU[NF] = { a0.x, a0.y, a1.x, a1.y, a2.x };
            const double nF[NF] = { a3.x, a3.y, a4.x, a4.y, a5.x };
            const double lam = llf_half_area_flux(cur.U, cFx, clx, nU, nF, a2.y, Ah, AFxhi);
            if (xf_ok) lmax = (lam < lmax) ? lmax : lam;
            double2 *f = fx + (row * FLX_SLOTS) * 32 + lane;
            f[0 * 32] = make_double2(AFxhi[0], AFxhi[1]);
            f[1 * 32] = make_double2(AFxhi[2], AFxhi[3]);
            f[2 * 32] = make_double2(AFxhi[4], 0.0);
            __syncwarp();
            const double2 *fl = fx + (row * FLX_SLOTS) * 32 + ((lane + 31) & 31);
            const double2 b0 = fl[0 * 32], b1 = fl[1 * 32], b2 = fl[2 * 32];
            AFxlo[0] = b0.x; AFxlo[1] = b0.y; AFxlo[2] = b1.x; AFxlo[3] = b1.y; AFxlo[4] = b2.x;
        }

        // ---- y direction: neighbouring warps ------------------------------------------------------
        __syncthreads();
        double AFyhi[NF], AFylo[NF];
        if (row <= NW - 2 && kz >= z0) {
            const double2 *o = rec + (row * REC_SLOTS) * 32 + lane;      // own Fy, ly (not kept in registers)
            const double2 o5 = o[5 * 32], o6 = o[6 * 32], o7 = o[7 * 32], o8 = o[8 * 32];
            const double cFy[NF] = { o6.x, o6.y, o7.x, o7.y, o8.x };
            const double2 *d = rec + ((row + 1) * REC_SLOTS) * 32 + lane;
            const double2 a0 = d[0 * 32], a1 = d[1 * 32], a2 = d[2 * 32], a5 = d[5 * 32], a6 = d[6 * 32], a7 = d[7 * 32], a8 = d[8 * 32];
            const double nU[NF] = { a0.x, a0.y, a1.x, a1.y, a2.x };
            const double nF[NF] = { a6.x, a6.y, a7.x, a7.y, a8.x };
            const double lam = llf_half_area_flux(cur.U, cFy, o5.y, nU, nF, a5.y, Ah, AFyhi);
            if (yf_ok) lmax = (lam < lmax) ? lmax : lam;
            double2 *f = fy + (row * FLX_SLOTS) * 32 + lane;
            f[0 * 32] = make_double2(AFyhi[0], AFyhi[1]);
            f[1 * 32] = make_double2(AFyhi[2], AFyhi[3]);
            f[2 * 32] = make_double2(AFyhi[4], 0.0);
        }
        __syncthreads();

        if (upd_row && kz >= z0) {
            const double2 *f = fy + ((row - 1) * FLX_SLOTS) * 32 + lane;
            const double2 b0 = f[0 * 32], b1 = f[1 * 32], b2 = f[2 * 32];
            AFylo[0] = b0.x; AFylo[1] = b0.y; AFylo[2] = b1.x; AFylo[3] = b1.y; AFylo[4] = b2.x;

            // ---- ordered accumulation (see uniform_kernels.cuh for the derivation) ----------------
            const int gk = g.gz0 + kz;
            const bool blo_z = (gk == 0);
            if (ORDER == NUM_AXIS) {
#pragma unroll
                for (int k = 0; k < NF; ++k) cur.S[k] = ((((0.0 + AFxlo[k]) - AFxhi[k]) + AFylo[k]) - AFyhi[k]) + AFz[k];
            } else if (blo_x | blo_y | blo_z) {
                // rare: a low face on the domain border moves from the "created by a lower cell"
                // group to the cell's own group
                int kx = key_x, ky = key_y, kzz = blo_z ? -1 : 3 * (__ffs(gk) - 1) + 2;
                if (ORDER == NUM_LEXI) { kx = blo_x ? -1 : 0; ky = blo_y ? -1 : 1; kzz = blo_z ? -1 : 2; }
                const int last = (kx < ky) ? ((kx < kzz) ? 0 : 2) : ((ky < kzz) ? 1 : 2);
#pragma unroll
                for (int k = 0; k < NF; ++k) {
                    const double cx = blo_x ? 0.0 : AFxlo[k];
                    const double cy = blo_y ? 0.0 : AFylo[k];
                    const double cz = blo_z ? 0.0 : AFz[k];
                    const double p = (last == 0) ? cy : cx;
                    const double r = (last == 0) ? cx : (last == 1) ? cy : cz;
                    const double t = (last == 2) ? cy : cz;
                    double s = ((0.0 + p) + t) + r;
                    if (blo_x) s += AFxlo[k];
                    s -= AFxhi[k];
                    if (blo_y) s += AFylo[k];
                    s -= AFyhi[k];
                    if (blo_z) s += AFz[k];
                    cur.S[k] = s;
                }
            } else if (ORDER == NUM_LEXI) {
#pragma unroll
                for (int k = 0; k < NF; ++k) cur.S[k] = ((((AFz[k] + AFylo[k]) + AFxlo[k])) - AFxhi[k]) - AFyhi[k];
            } else {
                // Morton: low faces ordered by creator <=> keys 3*ctz(coord)+axis descending; a+b is
                // commutative so only the LAST term matters.  key_y / key_z are warp-uniform.
                const int key_z = 3 * (__ffs(gk) - 1) + 2;
                if (key_y < key_z) {
                    const bool xl = key_x < key_y; // x-low last, else y-low last
#pragma unroll
                    for (int k = 0; k < NF; ++k) {
                        const double a = xl ? AFylo[k] : AFxlo[k];
                        const double r = xl ? AFxlo[k] : AFylo[k];
                        cur.S[k] = (((a + AFz[k]) + r) - AFxhi[k]) - AFyhi[k];
                    }
                } else {
                    const bool xl = key_x < key_z; // x-low last, else z-low last
#pragma unroll
                    for (int k = 0; k < NF; ++k) {
                        const double a = xl ? AFz[k] : AFxlo[k];
                        const double r = xl ? AFxlo[k] : AFz[k];
                        cur.S[k] = (((a + AFylo[k]) + r) - AFxhi[k]) - AFyhi[k];
                    }
                }
            }
        }
    }
};

template <int STAGE, int ORDER, int NW, bool UNROLL2>
__global__ void __launch_bounds__(NW * 32, 1)
uniform_stage_kernel_v2(const UniformGeom g, const double *__restrict__ Sin, const double *Un, double *Out,
                        const StepControl *__restrict__ ctl, double *__restrict__ max_eig, const int lz)
{
    extern __shared__ double2 smem2[];
    if (STAGE >= 1 && ctl->active == 0.0) return;

    StageV2<STAGE, ORDER, NW> s{ g, Sin, Un, Out };
    s.rec = smem2;
    s.fx  = smem2 + NW * REC_SLOTS * 32;
    s.fy  = s.fx + NW * FLX_SLOTS * 32;
    s.lane = threadIdx.x & 31;
    s.row  = threadIdx.x >> 5;
    const int i = blockIdx.x * XW - 1 + s.lane;
    const int j = blockIdx.y * (NW - 2) - 1 + s.row;
    s.z0 = blockIdx.z * lz;
    s.z1 = min(s.z0 + lz, g.nz);
    const int ic = min(max(i, -1), g.nx);
    const int jc = min(max(j, -1), g.ny);
    const bool in_x = (i >= 0 && i < g.nx), in_y = (j >= 0 && j < g.ny);
    s.upd_row = (s.row >= 1 && s.row <= NW - 2);
    s.upd     = s.upd_row && s.lane >= 1 && s.lane <= XW && in_x && in_y;
    s.xf_ok   = s.upd_row && in_y && s.lane <= XW && i >= -1 && i < g.nx;
    s.yf_ok   = s.row <= NW - 2 && in_x && s.lane >= 1 && s.lane <= XW && j >= -1 && j < g.ny;
    s.zf_ok   = s.upd_row && in_x && in_y;
    const int gi = g.gx0 + i, gj = g.gy0 + j;
    s.blo_x = (gi == 0);
    s.blo_y = (gj == 0);
    s.key_x = s.blo_x ? -1 : 3 * (__ffs(gi) - 1);
    s.key_y = s.blo_y ? -1 : 3 * (__ffs(gj) - 1) + 1;
    s.plane = (long long) g.py * g.px;
    s.col   = (long long) (jc + 1) * g.px + (ic + 1);
    s.fs    = g.fs;
    s.Ah    = 0.5 * g.area;
    s.dt    = (STAGE >= 1) ? ctl->dt : 0.0;
    s.dc.y_gm1 = rcp_nr(GM1);
    s.dc.y_c1  = rcp_nr(TWO_OVER_GM1);
    s.dc.y_vol = rcp_nr(g.volume);
    s.lmax = 0.0;

    PlaneRegs A, B;
#pragma unroll
    for (int k = 0; k < NF; ++k) { A.Fz[k] = 0.0; A.S[k] = 0.0; A.Un[k] = 0.0; B.Fz[k] = 0.0; B.S[k] = 0.0; B.Un[k] = 0.0; B.U[k] = 0.0; }
    A.lz = B.lz = 0.0;
    s.load_plane(A.U, Sin, s.z0 - 1);

    if (UNROLL2) {
        // two planes per trip over two register sets: B is "previous" for A and vice versa
        for (int kz = s.z0 - 1; kz <= s.z1; kz += 2) {
            s.step(B, A, kz);
            if (kz + 1 > s.z1) break;
            s.step(A, B, kz + 1);
        }
    } else {
        // one plane per trip; the two register sets swap by copy
#pragma unroll 1
        for (int kz = s.z0 - 1; kz <= s.z1; ++kz) {
            s.step(B, A, kz);
            if (kz == s.z1) break;
            // after step: A = derived data of plane kz (+ its S, Un), B.U = loaded plane kz+1
#pragma unroll
            for (int k = 0; k < NF; ++k) {
                const double u = A.U[k]; A.U[k] = B.U[k]; B.U[k] = u;
                B.Fz[k] = A.Fz[k]; B.S[k] = A.S[k];
            }
            B.lz = A.lz;
        }
    }

    // ---- max eigenvalue: warp shuffle, block reduction, one atomic per CTA ----------------------
    double lmax = s.lmax;
    __syncthreads();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, lmax, o);
        lmax = (lmax < other) ? other : lmax;
    }
    double *red = reinterpret_cast<double *>(smem2);
    if (s.lane == 0) red[s.row] = lmax;
    __syncthreads();
    if (s.row == 0) {
        double v = (s.lane < NW) ? red[s.lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, v, o);
            v = (v < other) ? other : v;
        }
        if (s.lane == 0) atomic_max_nonneg(max_eig, v);
    }
}

} // namespace mmf
