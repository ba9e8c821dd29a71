/*
 * TEST INFRASTRUCTURE -- the parity oracle.  NOT part of the product: only
 * tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may load it.
 *
 * Plain-C restatement of CabanaPIC's per-timestep hot path, scalar and serial,
 * written from the reference's algorithm (file:line cited per function,
 * relative to /root/reference).  Compiled twice: -DREAL=float and -DREAL=double
 * (oracle/Makefile), with -O2 -ffp-contract=off so every operation rounds once,
 * in the order the reference's expressions round.
 *
 * Parity is PINNED: tests/test_oracle.py checks every function here bit-for-bit
 * against the reference's own sources compiled from /root/reference
 * (oracle/_ref/libcpic_ref_*.so, built by oracle/Makefile from
 * oracle/ref_driver.cpp), and that build reproduces
 * tests/energy_comparison/energies_gold.2stream-em.double on all 6000 lines.
 *
 * Layouts (all plain arrays):
 *   particles     8 arrays of np: dx dy dz ux uy uz w (REAL) and cell (int)
 *   fields        9 arrays of nc: ex ey ez cbx cby cbz jfx jfy jfz
 *   interpolators [nc][18] in the reference's InterpolatorFields order
 *   accumulators  [nc][3][4]  (jx[4] jy[4] jz[4])
 * Cell index = x + (nx+2ng)*(y + (ny+2ng)*z)           (src/types.h:195)
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#ifndef REAL
#define REAL float
#endif
typedef REAL real;

typedef struct {
    double qdt_2mc, cdt_dx, cdt_dy, cdt_dz, qsp;
    double dx, dy, dz, dt;
    double px, py, pz, dt_eps0;
} orc_consts;

enum { I_EX, I_DEXDY, I_DEXDZ, I_D2EXDYDZ, I_EY, I_DEYDZ, I_DEYDX, I_D2EYDZDX, I_EZ, I_DEZDX, I_DEZDY,
       I_D2EZDXDY, I_CBX, I_DCBXDX, I_CBY, I_DCBYDY, I_CBZ, I_DCBZDZ, I_N };
enum { F_EX, F_EY, F_EZ, F_CBX, F_CBY, F_CBZ, F_JFX, F_JFY, F_JFZ, F_N };

int orc_real_bytes(void) { return (int)sizeof(real); }

static inline long vox(long x, long y, long z, long nx, long ny, long ng) {
    return x + (nx + 2 * ng) * (y + (ny + 2 * ng) * z);
}

/* src/interpolator.cpp:48-111 -- interior cells only; ghost records untouched */
void orc_load_interpolator(real* const* f, real* ip, long nx, long ny, long nz, long ng) {
    const long sx = 1, sy = nx + 2 * ng, sz = (nx + 2 * ng) * (ny + 2 * ng);
    const real fourth = 1.0 / 4.0, half = 1.0 / 2.0;
    const real *ex = f[F_EX], *ey = f[F_EY], *ez = f[F_EZ], *bx = f[F_CBX], *by = f[F_CBY], *bz = f[F_CBZ];
    for (long x = ng; x < nx + ng; ++x)
        for (long y = ng; y < ny + ng; ++y)
            for (long z = ng; z < nz + ng; ++z) {
                const long i = vox(x, y, z, nx, ny, ng);
                real* o = ip + i * I_N;
                real w0, w1, w2, w3;
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
            }
}

/* src/accumulator.cpp:13-42 */
void orc_clear_accumulator(real* acc, long nc) { memset(acc, 0, (size_t)nc * 12 * sizeof(real)); }

/* One streak's 12 quadrant currents into cell `a` -- the arithmetic shared by
 * CALC_J (src/push.h:218-232) and accumulate_j (src/move_p.h:156-170).
 * d = streak midpoint, u = half displacement (cell units), v5 = correction. */
static inline void deposit(real* a, real q, real ux, real uy, real uz, real dx, real dy, real dz, real v5,
                           real one) {
    real v0, v1, v2, v3, v4;
#define QUAD(U, DA, DB, OUT)            \
    v4 = q * (U);                       \
    v1 = v4 * (DA);                     \
    v0 = v4 - v1;                       \
    v1 += v4;                           \
    v4 = one + (DB);                    \
    v2 = v0 * v4;                       \
    v3 = v1 * v4;                       \
    v4 = one - (DB);                    \
    v0 *= v4;                           \
    v1 *= v4;                           \
    v0 += v5;                           \
    v1 -= v5;                           \
    v2 -= v5;                           \
    v3 += v5;                           \
    (OUT)[0] += v0;                     \
    (OUT)[1] += v1;                     \
    (OUT)[2] += v2;                     \
    (OUT)[3] += v3;
    QUAD(ux, dy, dz, a + 0)
    QUAD(uy, dz, dx, a + 4)
    QUAD(uz, dx, dy, a + 8)
#undef QUAD
}

/* src/move_p.h:9-54: which ghost layer (if any) the neighbour index sits in.
 * Later tests override earlier ones; one ghost layer is hard-wired. */
static inline int leaving_domain(long nx, long ny, long nz, long ix, long iy, long iz) {
    int leaving = -1;
    if (ix == 0) leaving = 0;
    if (iy == 0) leaving = 1;
    if (iz == 0) leaving = 2;
    if (ix == nx + 1) leaving = 3;
    if (iy == ny + 1) leaving = 4;
    if (iz == nz + 1) leaving = 5;
    return leaving;
}

/* src/move_p.h:93-371 -- cell-crossing mover for one particle.  p* point at the
 * particle's stored position, disp* is the remaining half-displacement.
 * In move_p the bare literals 3.4e38, 0.5, 2 and (1./3.) are doubles/ints mixed
 * into real arithmetic; the expressions below keep those promotions.
 * Returns the number of faces crossed (the reference returns 0; diagnostics). */
/* `periodic`: bits 0-2 = per-axis periodic wrap (7 = the reference at HEAD); bits 4-6 = per-axis REFLECTING walls:
 * Boundary::Reflect, which the reference declares (src/input/deck.h:9-12) but only carries as a commented block
 * (src/move_p.h:298-324, VPIC's move_p): a particle whose streak ends on a domain face keeps its cell, stays exactly
 * on the face, and has the momentum component and the remaining displacement along that axis reversed.  No
 * reference output exists for it: parity for Reflect is UNPINNED by the reference. */
static int move_particle(real* px, real* py, real* pz, int* pcell, real* acc, real q, real dispx, real dispy,
                         real dispz, long nx, long ny, long nz, long ng, int periodic, real* pux, real* puy, real* puz) {
    int crossings = 0;
    for (;;) {
        real mx = *px, my = *py, mz = *pz;
        real sx = dispx, sy = dispy, sz = dispz;
        real dir[3];
        real v0, v1, v2, v3;
        int axis;
        dir[0] = (sx > 0) ? 1 : -1;
        dir[1] = (sy > 0) ? 1 : -1;
        dir[2] = (sz > 0) ? 1 : -1;
        /* twice the fractional distance to each face (src/move_p.h:115-117) */
        v0 = (sx == 0) ? 3.4e38 : (dir[0] - mx) / sx;
        v1 = (sy == 0) ? 3.4e38 : (dir[1] - my) / sy;
        v2 = (sz == 0) ? 3.4e38 : (dir[2] - mz) / sz;
        v3 = 2; axis = 3;                               /* :125-129 */
        if (v0 < v3) { v3 = v0; axis = 0; }
        if (v1 < v3) { v3 = v1; axis = 1; }
        if (v2 < v3) { v3 = v2; axis = 2; }
        v3 *= 0.5;
        sx *= v3; sy *= v3; sz *= v3;                   /* :132-137 */
        mx += sx; my += sy; mz += sz;
        {
            const int ii = *pcell;                      /* :144 */
            const real v5 = q * sx * sy * sz * (1. / 3.); /* :154, final multiply in double */
            deposit(acc + (size_t)ii * 12, q, sx, sy, sz, mx, my, mz, v5, 1);
            dispx -= sx; dispy -= sy; dispz -= sz;      /* :195-197 */
            *px += sx + sx; *py += sy + sy; *pz += sz + sz; /* :201-203 */
            if (axis == 3) break;                       /* :209 */
            v0 = dir[axis];                             /* :218-227 snap onto the face */
            if (axis == 0) *px = v0;
            if (axis == 1) *py = v0;
            if (axis == 2) *pz = v0;
            {
                int face = axis;
                long ix, iy, iz;
                const long gx = nx + 2 * ng, gy = ny + 2 * ng;
                if (v0 > 0) face += 3;
                iy = ii / gx; ix = ii - iy * gx;        /* RANK_TO_INDEX, src/types.h:184-193 */
                iz = iy / gy; iy -= iz * gy;
                if (face == 0) ix--;
                if (face == 1) iy--;
                if (face == 2) iz--;
                if (face == 3) ix++;
                if (face == 4) iy++;
                if (face == 5) iz++;
                {
                    const int refl = periodic >> 4;
                    const long coord = axis == 0 ? ix : (axis == 1 ? iy : iz), top = axis == 0 ? nx : (axis == 1 ? ny : nz);
                    if ((refl & (1 << axis)) && (coord == 0 || coord == top + 1)) {
                        if (axis == 0) { dispx = -dispx; if (pux) *pux = -*pux; }
                        if (axis == 1) { dispy = -dispy; if (puy) *puy = -*puy; }
                        if (axis == 2) { dispz = -dispz; if (puz) *puz = -*puz; }
                        ++crossings;
                        continue;                       /* same cell, position snapped on the face */
                    }
                }
                {
                    /* periodic = per-axis bit mask (7 = the reference).  An axis whose bit is clear
                     * (slab mode: that ghost layer belongs to a neighbour) neither wraps nor hides
                     * the wrap of another axis, so its tests are skipped. */
                    int lv = leaving_domain(nx, ny, nz, ix, iy, iz);
                    if ((periodic & 7) != 7) {
                        lv = -1;
                        if ((periodic & 1) && ix == 0) lv = 0;
                        if ((periodic & 2) && iy == 0) lv = 1;
                        if ((periodic & 4) && iz == 0) lv = 2;
                        if ((periodic & 1) && ix == nx + 1) lv = 3;
                        if ((periodic & 2) && iy == ny + 1) lv = 4;
                        if ((periodic & 4) && iz == nz + 1) lv = 5;
                    }
                    if (lv >= 0) {                      /* :257-288 */
                        if (lv == 0) ix = (nx - 1) + ng;
                        else if (lv == 1) iy = (ny - 1) + ng;
                        else if (lv == 2) iz = (nz - 1) + ng;
                        else if (lv == 3) ix = ng;
                        else if (lv == 4) iy = ng;
                        else if (lv == 5) iz = ng;
                    }
                }
                *pcell = (int)vox(ix, iy, iz, nx, ny, ng); /* :351-352 */
            }
            if (axis == 0) *px = -v0;                   /* :368-370 re-enter from the other side */
            if (axis == 1) *py = -v0;
            if (axis == 2) *pz = -v0;
            ++crossings;
        }
    }
    return crossings;
}

/* src/push.h:65-295 (+ move_p).  Returns the number of particles that took the
 * mover path; *ncross (optional) gets the total faces crossed. */
long orc_push(real* dx, real* dy, real* dz, real* ux, real* uy, real* uz, const real* w, int* cell, long np,
              const real* ip, real* acc, const orc_consts* k, long nx, long ny, long nz, long ng, int periodic,
              long* ncross) {
    const real qdt_2mc = (real)k->qdt_2mc, cdt_dx = (real)k->cdt_dx, cdt_dy = (real)k->cdt_dy,
               cdt_dz = (real)k->cdt_dz, qsp = (real)k->qsp;
    const real one = 1., one_third = 1. / 3., two_fifteenths = 2. / 15.;
    long movers = 0, crossings = 0;
    for (long n = 0; n < np; ++n) {
        const int ii = cell[n];
        const real* f = ip + (size_t)ii * I_N;
        real x = dx[n], y = dy[n], z = dz[n];
        real hax = qdt_2mc * ((f[I_EX] + y * f[I_DEXDY]) + z * (f[I_DEXDZ] + y * f[I_D2EXDYDZ]));
        real hay = qdt_2mc * ((f[I_EY] + z * f[I_DEYDZ]) + x * (f[I_DEYDX] + z * f[I_D2EYDZDX]));
        real haz = qdt_2mc * ((f[I_EZ] + x * f[I_DEZDX]) + y * (f[I_DEZDY] + x * f[I_D2EZDXDY]));
        real cbx = f[I_CBX] + x * f[I_DCBXDX];
        real cby = f[I_CBY] + y * f[I_DCBYDY];
        real cbz = f[I_CBZ] + z * f[I_DCBZDZ];
        real vx = ux[n], vy = uy[n], vz = uz[n];
        real v0, v1, v2, v3, v4, v5, q;
        vx += hax; vy += hay; vz += haz;                              /* half E kick */
        v0 = qdt_2mc / sqrtf(one + (vx * vx + (vy * vy + vz * vz)));   /* sqrtf even for double, :148 */
        v1 = cbx * cbx + (cby * cby + cbz * cbz);
        v2 = (v0 * v0) * v1;
        v3 = v0 * (one + v2 * (one_third + v2 * two_fifteenths));
        v4 = v3 / (one + v1 * (v3 * v3));
        v4 += v4;
        v0 = vx + v3 * (vy * cbz - vz * cby);
        v1 = vy + v3 * (vz * cbx - vx * cbz);
        v2 = vz + v3 * (vx * cby - vy * cbx);
        vx += v4 * (v1 * cbz - v2 * cby);
        vy += v4 * (v2 * cbx - v0 * cbz);
        vz += v4 * (v0 * cby - v1 * cbx);
        vx += hax; vy += hay; vz += haz;                              /* second half kick */
        ux[n] = vx; uy[n] = vy; uz[n] = vz;                           /* :165-167 */
        v0 = one / sqrtf(one + (vx * vx + (vy * vy + vz * vz)));       /* :169 */
        vx *= cdt_dx; vy *= cdt_dy; vz *= cdt_dz;                     /* order matters, :171-176 */
        vx *= v0; vy *= v0; vz *= v0;
        v0 = x + vx; v1 = y + vy; v2 = z + vz;                        /* streak midpoint */
        v3 = v0 + vx; v4 = v1 + vy; v5 = v2 + vz;                     /* new position */
        q = w[n] * qsp;
        if (v3 <= one && v4 <= one && v5 <= one && -v3 <= one && -v4 <= one && -v5 <= one) {
            dx[n] = v3; dy[n] = v4; dz[n] = v5;
            v5 = q * vx * vy * vz * one_third;                        /* :203, all in real */
            deposit(acc + (size_t)ii * 12, q, vx, vy, vz, v0, v1, v2, v5, one);
        } else {
            ++movers;
            crossings += move_particle(dx + n, dy + n, dz + n, cell + n, acc, q, vx, vy, vz, nx, ny, nz, ng,
                                       periodic, ux + n, uy + n, uz + n);
        }
    }
    if (ncross) *ncross = crossings;
    return movers;
}

/* src/uncenter_p.h:27-98 -- backward half rotation, then one +half E kick */
void orc_uncenter(const real* dx, const real* dy, const real* dz, real* ux, real* uy, real* uz, const int* cell,
                  long np, const real* ip, double qdt_2mc_) {
    const real qdt_2mc = (real)qdt_2mc_;
    const real qdt_4mc = -0.5 * qdt_2mc;
    const real one = 1., one_third = 1. / 3., two_fifteenths = 2. / 15.;
    for (long n = 0; n < np; ++n) {
        const real* f = ip + (size_t)cell[n] * I_N;
        real x = dx[n], y = dy[n], z = dz[n];
        real hax = qdt_2mc * ((f[I_EX] + y * f[I_DEXDY]) + z * (f[I_DEXDZ] + y * f[I_D2EXDYDZ]));
        real hay = qdt_2mc * ((f[I_EY] + z * f[I_DEYDZ]) + x * (f[I_DEYDX] + z * f[I_D2EYDZDX]));
        real haz = qdt_2mc * ((f[I_EZ] + x * f[I_DEZDX]) + y * (f[I_DEZDY] + x * f[I_D2EZDXDY]));
        real cbx = f[I_CBX] + x * f[I_DCBXDX];
        real cby = f[I_CBY] + y * f[I_DCBYDY];
        real cbz = f[I_CBZ] + z * f[I_DCBZDZ];
        real vx = ux[n], vy = uy[n], vz = uz[n];
        real v0 = qdt_4mc / (real)sqrt(one + (vx * vx + (vy * vy + vz * vz)));  /* sqrt, not sqrtf, :72 */
        real v1 = cbx * cbx + (cby * cby + cbz * cbz);
        real v2 = (v0 * v0) * v1;
        real v3 = v0 * (one + v2 * (one_third + v2 * two_fifteenths));
        real v4 = v3 / (one + v1 * (v3 * v3));
        v4 += v4;
        v0 = vx + v3 * (vy * cbz - vz * cby);
        v1 = vy + v3 * (vz * cbx - vx * cbz);
        v2 = vz + v3 * (vx * cby - vy * cbx);
        vx += v4 * (v1 * cbz - v2 * cby);
        vy += v4 * (v2 * cbx - v0 * cbz);
        vz += v4 * (v0 * cby - v1 * cbx);
        vx += hax; vy += hay; vz += haz;
        ux[n] = vx; uy[n] = vy; uz[n] = vz;
    }
}

/* src/accumulator.cpp:66-110 -- assignment into jf over [ng, n+ng] per axis */
void orc_unload_accumulator(real* const* f, const real* acc, long nx, long ny, long nz, long ng,
                            const orc_consts* k) {
    const real dx = (real)k->dx, dy = (real)k->dy, dz = (real)k->dz, dt = (real)k->dt;
    const real cx = 0.25 / (dy * dz * dt);  /* 0.25 is a double: divide in double, then narrow */
    const real cy = 0.25 / (dz * dx * dt);
    const real cz = 0.25 / (dx * dy * dt);
    real *jfx = f[F_JFX], *jfy = f[F_JFY], *jfz = f[F_JFZ];
#define A(c, comp, q) acc[((size_t)(c) * 3 + (comp)) * 4 + (q)]
    for (long x = ng; x < nx + ng + 1; ++x)
        for (long y = ng; y < ny + ng + 1; ++y)
            for (long z = ng; z < nz + ng + 1; ++z) {
                const long i = vox(x, y, z, nx, ny, ng);
                const long xd = vox(x - 1, y, z, nx, ny, ng), yd = vox(x, y - 1, z, nx, ny, ng),
                           zd = vox(x, y, z - 1, nx, ny, ng), xzd = vox(x - 1, y, z - 1, nx, ny, ng),
                           xyd = vox(x - 1, y - 1, z, nx, ny, ng), yzd = vox(x, y - 1, z - 1, nx, ny, ng);
                jfx[i] = cx * (A(i, 0, 0) + A(yd, 0, 1) + A(zd, 0, 2) + A(yzd, 0, 3));
                jfy[i] = cy * (A(i, 1, 0) + A(zd, 1, 1) + A(xd, 1, 2) + A(xzd, 1, 3));
                jfz[i] = cz * (A(i, 2, 0) + A(xd, 2, 1) + A(yd, 2, 2) + A(xyd, 2, 3));
            }
#undef A
}

/* src/fields.h:33-98 -- periodic ghost COPY of three components, x then y then z */
static void ghost_copy(real* a, real* b, real* c, long nx, long ny, long nz, long ng) {
    real* s[3] = {a, b, c};
    for (int m = 0; m < 3; ++m) {
        real* v = s[m];
        for (long z = 1; z < nz + 1; ++z)
            for (long y = 1; y < ny + 1; ++y) {
                v[vox(nx + 1, y, z, nx, ny, ng)] = v[vox(1, y, z, nx, ny, ng)];
                v[vox(0, y, z, nx, ny, ng)] = v[vox(nx, y, z, nx, ny, ng)];
            }
        for (long x = 0; x < nx + 2; ++x)
            for (long z = 1; z < nz + 1; ++z) {
                v[vox(x, ny + 1, z, nx, ny, ng)] = v[vox(x, 1, z, nx, ny, ng)];
                v[vox(x, 0, z, nx, ny, ng)] = v[vox(x, ny, z, nx, ny, ng)];
            }
        for (long y = 0; y < ny + 2; ++y)
            for (long x = 0; x < nx + 2; ++x) {
                v[vox(x, y, nz + 1, nx, ny, ng)] = v[vox(x, y, 1, nx, ny, ng)];
                v[vox(x, y, 0, nx, ny, ng)] = v[vox(x, y, nz, nx, ny, ng)];
            }
    }
}

/* src/fields.h:126-183 -- periodic ghost FOLD of J: upper ghost += into cell 1,
 * two sequential sweeps per component (so the corner arrives via two hops). */
static void ghost_fold(real* jx, real* jy, real* jz, long nx, long ny, long nz, long ng) {
    for (long x = 1; x <= nx; ++x) {
        for (long z = 1; z <= nz + 1; ++z) jx[vox(x, 1, z, nx, ny, ng)] += jx[vox(x, ny + 1, z, nx, ny, ng)];
        for (long y = 1; y <= ny + 1; ++y) jx[vox(x, y, 1, nx, ny, ng)] += jx[vox(x, y, nz + 1, nx, ny, ng)];
    }
    for (long y = 1; y <= ny; ++y) {
        for (long x = 1; x <= nx + 1; ++x) jy[vox(x, y, 1, nx, ny, ng)] += jy[vox(x, y, nz + 1, nx, ny, ng)];
        for (long z = 1; z <= nz + 1; ++z) jy[vox(1, y, z, nx, ny, ng)] += jy[vox(nx + 1, y, z, nx, ny, ng)];
    }
    for (long z = 1; z <= nz; ++z) {
        for (long y = 1; y <= ny + 1; ++y) jz[vox(1, y, z, nx, ny, ng)] += jz[vox(nx + 1, y, z, nx, ny, ng)];
        for (long x = 1; x <= nx + 1; ++x) jz[vox(x, 1, z, nx, ny, ng)] += jz[vox(x, ny + 1, z, nx, ny, ng)];
    }
}
void orc_ghost_copy(real* a, real* b, real* c, long nx, long ny, long nz, long ng) { ghost_copy(a, b, c, nx, ny, nz, ng); }
void orc_ghost_fold(real* a, real* b, real* c, long nx, long ny, long nz, long ng) { ghost_fold(a, b, c, nx, ny, nz, ng); }

/* src/fields.h:692-718 -- EM advance_b over the interior, then ghost copy of cB */
void orc_advance_b(real* const* f, double px_, double py_, double pz_, long nx, long ny, long nz, long ng) {
    const real px = (real)px_, py = (real)py_, pz = (real)pz_;
    const real *ex = f[F_EX], *ey = f[F_EY], *ez = f[F_EZ];
    real *cbx = f[F_CBX], *cby = f[F_CBY], *cbz = f[F_CBZ];
    for (long x = 1; x < nx + 1; ++x)
        for (long y = 1; y < ny + 1; ++y)
            for (long z = 1; z < nz + 1; ++z) {
                const long f0 = vox(x, y, z, nx, ny, ng), fx = vox(x + 1, y, z, nx, ny, ng),
                           fy = vox(x, y + 1, z, nx, ny, ng), fz = vox(x, y, z + 1, nx, ny, ng);
                cbx[f0] -= (py * (ez[fy] - ez[f0]) - pz * (ey[fz] - ey[f0]));
                cby[f0] -= (pz * (ex[fz] - ex[f0]) - px * (ez[fx] - ez[f0]));
                cbz[f0] -= (px * (ey[fx] - ey[f0]) - py * (ex[fy] - ex[f0]));
            }
    ghost_copy(cbx, cby, cbz, nx, ny, nz, ng);
}

/* src/fields.h:618-665 -- EM advance_e: fold J, copy J ghosts, update [1, n+1] */
void orc_advance_e_em(real* const* f, double px_, double py_, double pz_, long nx, long ny, long nz, long ng,
                      double dt_eps0) {
    const real px = (real)px_, py = (real)py_, pz = (real)pz_;
    const real cj = (real)dt_eps0;
    real *ex = f[F_EX], *ey = f[F_EY], *ez = f[F_EZ];
    const real *cbx = f[F_CBX], *cby = f[F_CBY], *cbz = f[F_CBZ];
    real *jfx = f[F_JFX], *jfy = f[F_JFY], *jfz = f[F_JFZ];
    ghost_fold(jfx, jfy, jfz, nx, ny, nz, ng);
    ghost_copy(jfx, jfy, jfz, nx, ny, nz, ng);
    for (long x = 1; x < nx + 2; ++x)
        for (long y = 1; y < ny + 2; ++y)
            for (long z = 1; z < nz + 2; ++z) {
                const long f0 = vox(x, y, z, nx, ny, ng), fx = vox(x - 1, y, z, nx, ny, ng),
                           fy = vox(x, y - 1, z, nx, ny, ng), fz = vox(x, y, z - 1, nx, ny, ng);
                ex[f0] = ex[f0] + (-cj * jfx[f0]) + (py * (cbz[f0] - cbz[fy]) - pz * (cby[f0] - cby[fz]));
                ey[f0] = ey[f0] + (-cj * jfy[f0]) + (pz * (cbx[f0] - cbx[fz]) - px * (cbz[f0] - cbz[fx]));
                ez[f0] = ez[f0] + (-cj * jfz[f0]) + (px * (cby[f0] - cby[fx]) - py * (cbx[f0] - cbx[fy]));
            }
}

/* src/fields.h:511-544 -- ES_1D advance_e: fold J (no ghost copy), every cell */
void orc_advance_e_es1d(real* const* f, long nx, long ny, long nz, long ng, double dt_eps0) {
    const real cj = (real)dt_eps0;
    const long nc = (nx + 2 * ng) * (ny + 2 * ng) * (nz + 2 * ng);
    real *ex = f[F_EX], *ey = f[F_EY], *ez = f[F_EZ];
    real *jfx = f[F_JFX], *jfy = f[F_JFY], *jfz = f[F_JFZ];
    ghost_fold(jfx, jfy, jfz, nx, ny, nz, ng);
    for (long i = 0; i < nc; ++i) {
        ex[i] = ex[i] + (-cj * jfx[i]);
        ey[i] = ey[i] + (-cj * jfy[i]);
        ez[i] = ez[i] + (-cj * jfz[i]);
    }
}

/* ---- slab-mode helpers (multi-GPU tests only; NOT in the reference, which is single-domain).
 * They are the pieces of the functions above with the z-direction ghost work left out, so a
 * test can interleave a z-plane exchange between ranks exactly where the periodic z copy/fold
 * would have happened.  per = axis bit mask (1 x, 2 y, 4 z) of the axes handled locally. */
void orc_ghost_copy_axes(real* a, real* b, real* c, long nx, long ny, long nz, long ng, int per) {
    real* s[3] = {a, b, c};
    for (int m = 0; m < 3; ++m) {
        real* v = s[m];
        if (per & 1)
            for (long z = 1; z < nz + 1; ++z)
                for (long y = 1; y < ny + 1; ++y) {
                    v[vox(nx + 1, y, z, nx, ny, ng)] = v[vox(1, y, z, nx, ny, ng)];
                    v[vox(0, y, z, nx, ny, ng)] = v[vox(nx, y, z, nx, ny, ng)];
                }
        if (per & 2)
            for (long x = 0; x < nx + 2; ++x)
                for (long z = 1; z < nz + 1; ++z) {
                    v[vox(x, ny + 1, z, nx, ny, ng)] = v[vox(x, 1, z, nx, ny, ng)];
                    v[vox(x, 0, z, nx, ny, ng)] = v[vox(x, ny, z, nx, ny, ng)];
                }
        if (per & 4)
            for (long y = 0; y < ny + 2; ++y)
                for (long x = 0; x < nx + 2; ++x) {
                    v[vox(x, y, nz + 1, nx, ny, ng)] = v[vox(x, y, 1, nx, ny, ng)];
                    v[vox(x, y, 0, nx, ny, ng)] = v[vox(x, y, nz, nx, ny, ng)];
                }
    }
}
/* first (phase 0) or second (phase 1) sweep of every component's fold; z folds only if per & 4 */
void orc_ghost_fold_phase(real* jx, real* jy, real* jz, long nx, long ny, long nz, long ng, int phase, int per) {
    if (phase == 0) {
        for (long x = 1; x <= nx; ++x)
            for (long z = 1; z <= nz + 1; ++z) jx[vox(x, 1, z, nx, ny, ng)] += jx[vox(x, ny + 1, z, nx, ny, ng)];
        if (per & 4)
            for (long y = 1; y <= ny; ++y)
                for (long x = 1; x <= nx + 1; ++x) jy[vox(x, y, 1, nx, ny, ng)] += jy[vox(x, y, nz + 1, nx, ny, ng)];
        for (long z = 1; z <= nz; ++z)
            for (long y = 1; y <= ny + 1; ++y) jz[vox(1, y, z, nx, ny, ng)] += jz[vox(nx + 1, y, z, nx, ny, ng)];
    } else {
        if (per & 4)
            for (long x = 1; x <= nx; ++x)
                for (long y = 1; y <= ny + 1; ++y) jx[vox(x, y, 1, nx, ny, ng)] += jx[vox(x, y, nz + 1, nx, ny, ng)];
        for (long y = 1; y <= ny; ++y)
            for (long z = 1; z <= nz + 1; ++z) jy[vox(1, y, z, nx, ny, ng)] += jy[vox(nx + 1, y, z, nx, ny, ng)];
        for (long z = 1; z <= nz; ++z)
            for (long x = 1; x <= nx + 1; ++x) jz[vox(x, 1, z, nx, ny, ng)] += jz[vox(x, ny + 1, z, nx, ny, ng)];
    }
}
void orc_advance_b_stencil(real* const* f, double px_, double py_, double pz_, long nx, long ny, long nz, long ng) {
    const real px = (real)px_, py = (real)py_, pz = (real)pz_;
    const real *ex = f[F_EX], *ey = f[F_EY], *ez = f[F_EZ];
    real *cbx = f[F_CBX], *cby = f[F_CBY], *cbz = f[F_CBZ];
    for (long x = 1; x < nx + 1; ++x)
        for (long y = 1; y < ny + 1; ++y)
            for (long z = 1; z < nz + 1; ++z) {
                const long f0 = vox(x, y, z, nx, ny, ng), fx = vox(x + 1, y, z, nx, ny, ng),
                           fy = vox(x, y + 1, z, nx, ny, ng), fz = vox(x, y, z + 1, nx, ny, ng);
                cbx[f0] -= (py * (ez[fy] - ez[f0]) - pz * (ey[fz] - ey[f0]));
                cby[f0] -= (pz * (ex[fz] - ex[f0]) - px * (ez[fx] - ez[f0]));
                cbz[f0] -= (px * (ey[fx] - ey[f0]) - py * (ex[fy] - ex[f0]));
            }
}
void orc_advance_e_stencil(real* const* f, double px_, double py_, double pz_, long nx, long ny, long nz, long ng,
                           double dt_eps0) {
    const real px = (real)px_, py = (real)py_, pz = (real)pz_;
    const real cj = (real)dt_eps0;
    real *ex = f[F_EX], *ey = f[F_EY], *ez = f[F_EZ];
    const real *cbx = f[F_CBX], *cby = f[F_CBY], *cbz = f[F_CBZ];
    const real *jfx = f[F_JFX], *jfy = f[F_JFY], *jfz = f[F_JFZ];
    for (long x = 1; x < nx + 2; ++x)
        for (long y = 1; y < ny + 2; ++y)
            for (long z = 1; z < nz + 2; ++z) {
                const long f0 = vox(x, y, z, nx, ny, ng), fx = vox(x - 1, y, z, nx, ny, ng),
                           fy = vox(x, y - 1, z, nx, ny, ng), fz = vox(x, y, z - 1, nx, ny, ng);
                ex[f0] = ex[f0] + (-cj * jfx[f0]) + (py * (cbz[f0] - cbz[fy]) - pz * (cby[f0] - cby[fz]));
                ey[f0] = ey[f0] + (-cj * jfy[f0]) + (pz * (cbx[f0] - cbx[fz]) - px * (cbz[f0] - cbz[fx]));
                ez[f0] = ez[f0] + (-cj * jfz[f0]) + (px * (cby[f0] - cby[fx]) - py * (cbx[f0] - cbx[fy]));
            }
}

/* Field side of Boundary::Reflect (unpinned by the reference, which exit(1)s in src/fields.h:21-25,113-117): the box is
 * a perfect conductor -- "anti_symmetric_fields: E_tang = 0" of src/grid.h:4-17, the wall VPIC pairs with reflecting
 * particles.  The walls are the node planes 1 and n+1 of each axis; the E components tangential to a wall are zeroed
 * after every E update.  Nothing else is needed: no ghost copy or fold (no particle ever deposits into a ghost cell),
 * every E value the B update or the interpolator reads at index n+1 is tangential to that wall, and the normal cB on
 * a wall is never updated (dB_n/dt = -(curl E)_n = 0 there). */
void orc_pec_walls(real* const* f, long nx, long ny, long nz, long ng) {
    real *ex = f[F_EX], *ey = f[F_EY], *ez = f[F_EZ];
    for (long z = 0; z < nz + 2; ++z)
        for (long y = 0; y < ny + 2; ++y)
            for (long x = 0; x < nx + 2; ++x) {
                const long i = vox(x, y, z, nx, ny, ng);
                const int wx = x == 1 || x == nx + 1, wy = y == 1 || y == ny + 1, wz = z == 1 || z == nz + 1;
                if (wy || wz) ex[i] = 0;
                if (wz || wx) ey[i] = 0;
                if (wx || wy) ez[i] = 0;
            }
}

/* src/fields.h:556-615 (EM: interior) and :484-509 (ES_1D: every cell);
 * accumulated in real like the reference's parallel_reduce, x outermost. */
void orc_energies(real* const* f, int solver, long nx, long ny, long nz, long ng, double* e, double* b) {
    real es = 0, bs = 0;
    if (solver == 0) {
        for (long x = 1; x < nx + 1; ++x)
            for (long y = 1; y < ny + 1; ++y)
                for (long z = 1; z < nz + 1; ++z) {
                    const long i = vox(x, y, z, nx, ny, ng);
                    es += f[F_EX][i] * f[F_EX][i] + f[F_EY][i] * f[F_EY][i] + f[F_EZ][i] * f[F_EZ][i];
                }
        for (long x = 1; x < nx + 1; ++x)
            for (long y = 1; y < ny + 1; ++y)
                for (long z = 1; z < nz + 1; ++z) {
                    const long i = vox(x, y, z, nx, ny, ng);
                    bs += f[F_CBX][i] * f[F_CBX][i] + f[F_CBY][i] * f[F_CBY][i] + f[F_CBZ][i] * f[F_CBZ][i];
                }
    } else {
        const long nc = (nx + 2 * ng) * (ny + 2 * ng) * (nz + 2 * ng);
        for (long i = 0; i < nc; ++i)
            es += f[F_EX][i] * f[F_EX][i] + f[F_EY][i] * f[F_EY][i] + f[F_EZ][i] * f[F_EZ][i];
    }
    *e = (real)(es * 0.5f);
    *b = (real)(bs * 0.5f);
}

/* One time step in the reference's call order, example/example.cpp:221-266 */
void orc_step(real* dx, real* dy, real* dz, real* ux, real* uy, real* uz, const real* w, int* cell, long np,
              real* const* f, real* ip, real* acc, const orc_consts* k, int solver, long nx, long ny, long nz,
              long ng, long nsteps, double* en) {
    const long nc = (nx + 2 * ng) * (ny + 2 * ng) * (nz + 2 * ng);
    const real px = (real)k->px, py = (real)k->py, pz = (real)k->pz;
    const real hpx = (real)0.5 * px, hpy = (real)0.5 * py, hpz = (real)0.5 * pz;
    for (long s = 0; s < nsteps; ++s) {
        orc_load_interpolator(f, ip, nx, ny, nz, ng);
        orc_clear_accumulator(acc, nc);
        orc_push(dx, dy, dz, ux, uy, uz, w, cell, np, ip, acc, k, nx, ny, nz, ng, 7, 0);
        orc_unload_accumulator(f, acc, nx, ny, nz, ng, k);
        if (solver == 0) {
            orc_advance_b(f, hpx, hpy, hpz, nx, ny, nz, ng);
            orc_advance_e_em(f, px, py, pz, nx, ny, nz, ng, k->dt_eps0);
            orc_advance_b(f, hpx, hpy, hpz, nx, ny, nz, ng);
        } else {
            orc_advance_e_es1d(f, nx, ny, nz, ng, k->dt_eps0);
        }
        if (en) orc_energies(f, solver, nx, ny, nz, ng, en + 2 * s, en + 2 * s + 1);
    }
}
