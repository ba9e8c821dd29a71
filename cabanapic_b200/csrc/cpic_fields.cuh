// Field-side kernels: interpolator load, accumulator unload, periodic ghost fold/copy,
// Yee EM advance_b/advance_e, 1-D ES advance_e, field energies.
// All are bandwidth-bound O(cells) stencils: one thread per cell, x fastest so a warp
// touches consecutive addresses of each struct-of-arrays field member.
// This file is compiled with -fmad=false: every expression below rounds exactly like
// the reference's host build (results are bit-identical to the oracle).
#pragma once
#include "cpic_common.cuh"

namespace cpic {

// Decompose a flat thread index into (x,y,z) inside the box [x0,x0+wx) x [y0,y0+wy) x [z0,..).
struct Box {
    int x0, y0, z0, wx, wy, wz;
    __host__ __device__ long long count() const { return (long long)wx * wy * wz; }
};
__device__ __forceinline__ bool box_coords(const Box& b, long long t, int& x, int& y, int& z) {
    if (t >= b.count()) return false;
    const long long row = t / b.wx;
    x = b.x0 + (int)(t - row * b.wx);
    const int zz = (int)(row / b.wy);
    y = b.y0 + (int)(row - (long long)zz * b.wy);
    z = b.z0 + zz;
    return true;
}

// Reference: load_interpolator_array, src/interpolator.cpp:48-111.  Interior cells only;
// ghost-cell records keep whatever initialize_interpolator left there (zeros).
template <class R>
__global__ void __launch_bounds__(256) k_load_interpolator(Fields<R> f, R* __restrict__ ip, Grid g, Box b) {
    int x, y, z;
    if (!box_coords(b, blockIdx.x * 256LL + threadIdx.x, x, y, z)) return;
    const long long i = x + (long long)g.sy * y + (long long)g.sz * z;
    const int sx = 1, sy = g.sy, sz = g.sz;
    const R fourth = R(1.0 / 4.0), half = R(1.0 / 2.0);
    R o[IpStride<R>::value];
    R w0, w1, w2, w3;
    const R* __restrict__ ex = f.c[F_EX]; const R* __restrict__ ey = f.c[F_EY]; const R* __restrict__ ez = f.c[F_EZ];
    const R* __restrict__ bx = f.c[F_CBX]; const R* __restrict__ by = f.c[F_CBY]; const R* __restrict__ bz = f.c[F_CBZ];
    w0 = ex[i]; w1 = ex[i + sy]; w2 = ex[i + sz]; w3 = ex[i + sy + sz];
    o[I_EX] = fourth * ((w3 + w0) + (w1 + w2));
    o[I_DEXDY] = fourth * ((w3 - w0) + (w1 - w2));
    o[I_DEXDZ] = fourth * ((w3 - w0) - (w1 - w2));
    o[I_D2EXDYDZ] = fourth * ((w3 + w0) - (w1 + w2));
    w0 = ey[i]; w1 = ey[i + sz]; w2 = ey[i + sx]; w3 = ey[i + sx + sz];
    o[I_EY] = fourth * ((w3 + w0) + (w1 + w2));
    o[I_DEYDZ] = fourth * ((w3 - w0) + (w1 - w2));
    o[I_DEYDX] = fourth * ((w3 - w0) - (w1 - w2));
    o[I_D2EYDZDX] = fourth * ((w3 + w0) - (w1 + w2));
    w0 = ez[i]; w1 = ez[i + sx]; w2 = ez[i + sy]; w3 = ez[i + sx + sy];
    o[I_EZ] = fourth * ((w3 + w0) + (w1 + w2));
    o[I_DEZDX] = fourth * ((w3 - w0) + (w1 - w2));
    o[I_DEZDY] = fourth * ((w3 - w0) - (w1 - w2));
    o[I_D2EZDXDY] = fourth * ((w3 + w0) - (w1 + w2));
    w0 = bx[i]; w1 = bx[i + sx];
    o[I_CBX] = half * (w1 + w0); o[I_DCBXDX] = half * (w1 - w0);
    w0 = by[i]; w1 = by[i + sy];
    o[I_CBY] = half * (w1 + w0); o[I_DCBYDY] = half * (w1 - w0);
    w0 = bz[i]; w1 = bz[i + sz];
    o[I_CBZ] = half * (w1 + w0); o[I_DCBZDZ] = half * (w1 - w0);
    constexpr int S = IpStride<R>::value;
#pragma unroll
    for (int k = I_N; k < S; ++k) o[k] = R(0);
    // 16-byte stores of the whole (padded) record
    using V = typename std::conditional<sizeof(R) == 4, float4, double2>::type;
    constexpr int PER = 16 / sizeof(R);
    V* dst = reinterpret_cast<V*>(ip + i * S);
#pragma unroll
    for (int k = 0; k < S / PER; ++k) dst[k] = *reinterpret_cast<V*>(&o[k * PER]);
}

// Reference: unload_accumulator_array, src/accumulator.cpp:71-110 (box [ng, n+ng] per axis,
// assignment).  cx,cy,cz are computed on the host exactly as src/accumulator.cpp:66-68.
template <class R>
__global__ void __launch_bounds__(256) k_unload_accumulator(Fields<R> f, const R* __restrict__ acc, Grid g, Box b,
                                                            R cx, R cy, R cz) {
    int x, y, z;
    if (!box_coords(b, blockIdx.x * 256LL + threadIdx.x, x, y, z)) return;
    const long long i = x + (long long)g.sy * y + (long long)g.sz * z;
    const long long xd = i - 1, yd = i - g.sy, zd = i - g.sz;
    const long long xzd = zd - 1, xyd = yd - 1, yzd = zd - g.sy;
#define ACC(c, comp, q) acc[((c) * 3 + (comp)) * 4 + (q)]
    f.c[F_JFX][i] = cx * (ACC(i, 0, 0) + ACC(yd, 0, 1) + ACC(zd, 0, 2) + ACC(yzd, 0, 3));
    f.c[F_JFY][i] = cy * (ACC(i, 1, 0) + ACC(zd, 1, 1) + ACC(xd, 1, 2) + ACC(xzd, 1, 3));
    f.c[F_JFZ][i] = cz * (ACC(i, 2, 0) + ACC(xd, 2, 1) + ACC(yd, 2, 2) + ACC(xyd, 2, 3));
#undef ACC
}

// Reference: serial_update_ghosts_B, src/fields.h:33-98 -- periodic ghost COPY, x faces then
// y faces (incl. x ghosts) then z faces (incl. all).  The three ordered sweeps leave every
// ghost cell equal to the interior cell at the wrapped coordinates, so one pass suffices:
// each ghost thread reads interior(wrap(x),wrap(y),wrap(z)); interior cells are untouched.
// One thread per GHOST cell (O(surface), not O(cells): at 256^3 the ghost layer is 2.3 % of the grid): the two z ghost
// planes, then the y ghost rows of the interior planes, then the x ghost cells of the interior rows.
inline long long ghost_cell_count(const Grid& g) {
    return 2ll * g.gx * g.gy + 2ll * g.nz * g.gx + 2ll * g.nz * g.ny;
}
template <class R>
__global__ void __launch_bounds__(256) k_ghost_copy3(R* __restrict__ a, R* __restrict__ b3, R* __restrict__ c,
                                                     Grid g) {
    long long t = blockIdx.x * 256LL + threadIdx.x;
    const long long plane = (long long)g.gx * g.gy;
    int x, y, z;
    if (t < 2 * plane) {
        z = t < plane ? 0 : g.nz + 1;
        const long long r = t < plane ? t : t - plane;
        y = (int)(r / g.gx); x = (int)(r - (long long)y * g.gx);
    } else {
        t -= 2 * plane;
        const long long nrow = 2ll * g.nz * g.gx;
        if (t < nrow) {
            z = 1 + (int)(t / (2 * g.gx));
            const int r = (int)(t % (2 * g.gx));
            y = r < g.gx ? 0 : g.ny + 1;
            x = r < g.gx ? r : r - g.gx;
        } else {
            t -= nrow;
            if (t >= 2ll * g.nz * g.ny) return;
            z = 1 + (int)(t / (2 * g.ny));
            const int r = (int)(t % (2 * g.ny));
            y = 1 + (r >> 1);
            x = (r & 1) ? g.nx + 1 : 0;
        }
    }
    const int wx = (x == 0) ? g.nx : (x == g.nx + 1 ? 1 : x);
    const int wy = (y == 0) ? g.ny : (y == g.ny + 1 ? 1 : y);
    int wz = (z == 0) ? g.nz : (z == g.nz + 1 ? 1 : z);
    if (!(g.per & 4)) {           // z ghosts belong to the neighbouring slab: filled by the plane exchange
        if (wz != z) return;
        wz = z;
    }
    const long long d = x + (long long)g.sy * y + (long long)g.sz * z;
    const long long s = wx + (long long)g.sy * wy + (long long)g.sz * wz;
    a[d] = a[s]; b3[d] = b3[s]; c[d] = c[s];
}

// Field side of Boundary::Reflect (the reference exit(1)s, src/fields.h:21-25,113-117; parity unpinned): a perfectly
// conducting box -- "anti_symmetric_fields: E_tang = 0", src/grid.h:4-17.  The walls are the node planes 1 and n+1 of
// each axis; the E components tangential to a wall are zeroed after every E update.  One thread per cell of the six
// wall planes would do; the box is small next to the particle work, so one thread per cell it is.
template <class R>
__global__ void __launch_bounds__(256) k_pec_walls(R* __restrict__ ex, R* __restrict__ ey, R* __restrict__ ez, Grid g) {
    const long long t = blockIdx.x * 256LL + threadIdx.x;
    if (t >= g.nc) return;
    const long long row = t / g.gx;
    const int x = (int)(t - row * g.gx);
    const int z = (int)(row / g.gy);
    const int y = (int)(row - (long long)z * g.gy);
    const bool wx = x == 1 || x == g.nx + 1, wy = y == 1 || y == g.ny + 1, wz = z == 1 || z == g.nz + 1;
    if (wy || wz) ex[t] = R(0);
    if (wz || wx) ey[t] = R(0);
    if (wx || wy) ez[t] = R(0);
}

// Reference: serial_update_ghosts, src/fields.h:126-183 -- periodic ghost FOLD of J.  Each
// component does two ordered sweeps (the corner reaches cell 1 through two hops); PHASE 0 is
// the first sweep of all three components, PHASE 1 the second.  blockIdx.y = component.
template <class R, int PHASE>
__global__ void __launch_bounds__(256) k_ghost_fold(R* __restrict__ jx, R* __restrict__ jy, R* __restrict__ jz,
                                                    Grid g) {
    const int comp = blockIdx.y;
    const long long t = blockIdx.x * 256LL + threadIdx.x;
    const int nx = g.nx, ny = g.ny, nz = g.nz;
    // slab mode (z not periodic here): the z folds are done by the host from the neighbour's plane
    if (!(g.per & 4) && ((comp == 0 && PHASE == 1) || (comp == 1 && PHASE == 0))) return;
    const long long sy = g.sy, sz = g.sz;
    long long to, from;
    R* v;
    if (comp == 0) {          // jfx, x in [1,nx]
        v = jx;
        if (PHASE == 0) {     // z in [1,nz+1]: (x,1,z) += (x,ny+1,z)
            if (t >= (long long)nx * (nz + 1)) return;
            const int x = 1 + (int)(t % nx), z = 1 + (int)(t / nx);
            to = x + sy * 1 + sz * z; from = x + sy * (ny + 1) + sz * z;
        } else {              // y in [1,ny+1]: (x,y,1) += (x,y,nz+1)
            if (t >= (long long)nx * (ny + 1)) return;
            const int x = 1 + (int)(t % nx), y = 1 + (int)(t / nx);
            to = x + sy * y + sz * 1; from = x + sy * y + sz * (nz + 1);
        }
    } else if (comp == 1) {   // jfy, y in [1,ny]
        v = jy;
        if (PHASE == 0) {     // x in [1,nx+1]: (x,y,1) += (x,y,nz+1)
            if (t >= (long long)ny * (nx + 1)) return;
            const int x = 1 + (int)(t % (nx + 1)), y = 1 + (int)(t / (nx + 1));
            to = x + sy * y + sz * 1; from = x + sy * y + sz * (nz + 1);
        } else {              // z in [1,nz+1]: (1,y,z) += (nx+1,y,z)
            if (t >= (long long)ny * (nz + 1)) return;
            const int y = 1 + (int)(t % ny), z = 1 + (int)(t / ny);
            to = 1 + sy * y + sz * z; from = (nx + 1) + sy * y + sz * z;
        }
    } else {                  // jfz, z in [1,nz]
        v = jz;
        if (PHASE == 0) {     // y in [1,ny+1]: (1,y,z) += (nx+1,y,z)
            if (t >= (long long)nz * (ny + 1)) return;
            const int y = 1 + (int)(t % (ny + 1)), z = 1 + (int)(t / (ny + 1));
            to = 1 + sy * y + sz * z; from = (nx + 1) + sy * y + sz * z;
        } else {              // x in [1,nx+1]: (x,1,z) += (x,ny+1,z)
            if (t >= (long long)nz * (nx + 1)) return;
            const int x = 1 + (int)(t % (nx + 1)), z = 1 + (int)(t / (nx + 1));
            to = x + sy * 1 + sz * z; from = x + sy * (ny + 1) + sz * z;
        }
    }
    v[to] += v[from];
}

// Reference: EM_Field_Solver::advance_b, src/fields.h:692-717 (interior box).
template <class R>
__global__ void __launch_bounds__(256) k_advance_b(Fields<R> f, Grid g, Box b, R px, R py, R pz) {
    int x, y, z;
    if (!box_coords(b, blockIdx.x * 256LL + threadIdx.x, x, y, z)) return;
    const long long f0 = x + (long long)g.sy * y + (long long)g.sz * z;
    const long long fx = f0 + 1, fy = f0 + g.sy, fz = f0 + g.sz;
    const R* __restrict__ ex = f.c[F_EX]; const R* __restrict__ ey = f.c[F_EY]; const R* __restrict__ ez = f.c[F_EZ];
    const R e0x = ex[f0], e0y = ey[f0], e0z = ez[f0];
    f.c[F_CBX][f0] -= (py * (ez[fy] - e0z) - pz * (ey[fz] - e0y));
    f.c[F_CBY][f0] -= (pz * (ex[fz] - e0x) - px * (ez[fx] - e0z));
    f.c[F_CBZ][f0] -= (px * (ey[fx] - e0y) - py * (ex[fy] - e0x));
}

// Reference: EM_Field_Solver::advance_e, src/fields.h:646-664 (box [1, n+1] per axis: the upper
// ghost is computed, not copied).
template <class R>
__global__ void __launch_bounds__(256) k_advance_e_em(Fields<R> f, Grid g, Box b, R px, R py, R pz, R cj) {
    int x, y, z;
    if (!box_coords(b, blockIdx.x * 256LL + threadIdx.x, x, y, z)) return;
    const long long f0 = x + (long long)g.sy * y + (long long)g.sz * z;
    const long long fx = f0 - 1, fy = f0 - g.sy, fz = f0 - g.sz;
    const R* __restrict__ cbx = f.c[F_CBX]; const R* __restrict__ cby = f.c[F_CBY]; const R* __restrict__ cbz = f.c[F_CBZ];
    const R b0x = cbx[f0], b0y = cby[f0], b0z = cbz[f0];
    R* ex = f.c[F_EX]; R* ey = f.c[F_EY]; R* ez = f.c[F_EZ];
    ex[f0] = ex[f0] + (-cj * f.c[F_JFX][f0]) + (py * (b0z - cbz[fy]) - pz * (b0y - cby[fz]));
    ey[f0] = ey[f0] + (-cj * f.c[F_JFY][f0]) + (pz * (b0x - cbx[fz]) - px * (b0z - cbz[fx]));
    ez[f0] = ez[f0] + (-cj * f.c[F_JFZ][f0]) + (px * (b0y - cby[fx]) - py * (b0x - cbx[fy]));
}

// Reference: ES_Field_Solver_1D::advance_e, src/fields.h:534-543 -- every cell incl. ghosts.
template <class R>
__global__ void __launch_bounds__(256) k_advance_e_es(Fields<R> f, long long nc, R cj) {
    const long long i = blockIdx.x * 256LL + threadIdx.x;
    if (i >= nc) return;
    f.c[F_EX][i] = f.c[F_EX][i] + (-cj * f.c[F_JFX][i]);
    f.c[F_EY][i] = f.c[F_EY][i] + (-cj * f.c[F_JFY][i]);
    f.c[F_EZ][i] = f.c[F_EZ][i] + (-cj * f.c[F_JFZ][i]);
}

// Reference: e_energy/b_energy, src/fields.h:556-615 (EM: interior) and :484-509 (ES_1D: all
// cells).  The reference sums in real_t in backend order; we sum in double (block tree +
// one atomic per block), so agreement is to rounding, not bitwise.  out[0]=sum E^2, out[1]=sum cB^2.
template <class R>
__global__ void __launch_bounds__(256) k_energy(Fields<R> f, Grid g, Box b, int with_b, double* __restrict__ out) {
    int x, y, z;
    double e = 0.0, m = 0.0;
    for (long long t = blockIdx.x * 256LL + threadIdx.x; t < b.count(); t += (long long)gridDim.x * 256) {
        box_coords(b, t, x, y, z);
        const long long i = x + (long long)g.sy * y + (long long)g.sz * z;
        const double a0 = f.c[F_EX][i], a1 = f.c[F_EY][i], a2 = f.c[F_EZ][i];
        e += a0 * a0 + a1 * a1 + a2 * a2;
        if (with_b) {
            const double b0 = f.c[F_CBX][i], b1 = f.c[F_CBY][i], b2 = f.c[F_CBZ][i];
            m += b0 * b0 + b1 * b1 + b2 * b2;
        }
    }
    __shared__ double se[8], sm[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        e += __shfl_xor_sync(0xffffffffu, e, o);
        m += __shfl_xor_sync(0xffffffffu, m, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { se[w] = e; sm[w] = m; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double te = 0.0, tm = 0.0;
        for (int k = 0; k < 8; ++k) { te += se[k]; tm += sm[k]; }
        atomicAdd(out, te);
        atomicAdd(out + 1, tm);
    }
}

}  // namespace cpic
