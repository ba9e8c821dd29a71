// TEST INFRASTRUCTURE -- not part of the product.  Only tests/, bench.py's
// cpu_baseline / --impl reference legs and __graft_entry__.smoke() may use it.
//
// C-callable driver around the REFERENCE'S OWN hot-path sources, compiled where
// they lie under /root/reference (nothing is copied): src/push.h, src/move_p.h,
// src/interpolator.cpp, src/accumulator.cpp, src/fields.h, src/uncenter_p.h and
// src/input/deck.h, against the host stand-in headers in include/compat/.
// The recipe is oracle/Makefile; outputs go to oracle/_ref/ only.
//
// What is ours here: the state container, the getters/setters, and a
// restatement of the driver-side constants (example/example.cpp:77-113,179-181)
// and call order (example/example.cpp:221-266).  That restatement is pinned by
// comparing ref_run()'s energies with the energies.txt written by the
// reference's unmodified main() (oracle/_ref/gold_*), which in turn reproduces
// tests/energy_comparison/energies_gold.2stream-em.double on all 6000 lines.
//
// Built once per (deck, precision[, openmp]); without -DUSER_INPUT_DECK the
// global `deck` is the reference's built-in default deck (src/input/deck.h:431).

#include <Cabana_Core.hpp>

#include "types.h"
#include "fields.h"
#include "accumulator.h"
#include "interpolator.h"
#include "uncenter_p.h"
#include "push.h"
#include "input/deck.h"

Input_Deck deck;  // example/example.cpp:39

namespace {

struct Consts {  // all doubles on the wire, narrowed to real_t on entry
    double qdt_2mc, cdt_dx, cdt_dy, cdt_dz, qsp;
    double dx, dy, dz, dt;
    double px, py, pz, dt_eps0;
};

struct RefSim {
    size_t nx, ny, nz, ng, num_cells, np;
    int solver;  // 0 = EM, 1 = ES_1D
    particle_list_t particles;
    interpolator_array_t interpolators;
    accumulator_array_t accumulators;
    accumulator_array_sa_t scatter_add;
    field_array_t fields;
    grid_t* grid;
    RefSim(size_t nx_, size_t ny_, size_t nz_, size_t ng_, size_t np_, int solver_)
        : nx(nx_), ny(ny_), nz(nz_), ng(ng_), num_cells((nx_ + 2 * ng_) * (ny_ + 2 * ng_) * (nz_ + 2 * ng_)),
          np(np_), solver(solver_), particles("particles", np_), interpolators("interpolator", num_cells),
          accumulators("accumulator", num_cells), fields("fields", num_cells), grid(new grid_t()) {
        scatter_add = Kokkos::Experimental::create_scatter_view(accumulators);  // example.cpp:139
        initialize_interpolator(interpolators);                                  // example.cpp:148
        Field_Solver<EM_Field_Solver> zero_them(fields);                         // example.cpp:154 (ctor zeroes)
    }
    ~RefSim() { delete grid; }
};

template <int M, class A, class T>
void put(A& a, const T* src, size_t n) {
    auto s = Cabana::slice<M>(a);
    for (size_t i = 0; i < n; ++i) s(i) = src[i];
}
template <int M, class A, class T>
void get(const A& a, T* dst, size_t n) {
    auto s = Cabana::slice<M>(a);
    for (size_t i = 0; i < n; ++i) dst[i] = s(i);
}

void one_step(RefSim& S, const Consts& k) {
    const real_t px = k.px, py = k.py, pz = k.pz;
    load_interpolator_array(S.fields, S.interpolators, S.nx, S.ny, S.nz, S.ng);  // example.cpp:221
    clear_accumulator_array(S.fields, S.accumulators, S.nx, S.ny, S.nz);         // :223
    push(S.particles, S.interpolators, (real_t)k.qdt_2mc, (real_t)k.cdt_dx, (real_t)k.cdt_dy, (real_t)k.cdt_dz,
         (real_t)k.qsp, S.scatter_add, S.grid, S.nx, S.ny, S.nz, S.ng, deck.BOUNDARY_TYPE);  // :231
    Kokkos::Experimental::contribute(S.accumulators, S.scatter_add);             // :248
    S.scatter_add.reset_except(S.accumulators);                                  // :251
    unload_accumulator_array(S.fields, S.accumulators, S.nx, S.ny, S.nz, S.ng, (real_t)k.dx, (real_t)k.dy,
                             (real_t)k.dz, (real_t)k.dt);                        // :257
    if (S.solver == 0) {
        EM_Field_Solver em;
        em.advance_b(S.fields, real_t(0.5) * px, real_t(0.5) * py, real_t(0.5) * pz, S.nx, S.ny, S.nz, S.ng);  // :260
        em.advance_e(S.fields, px, py, pz, S.nx, S.ny, S.nz, S.ng, (real_t)k.dt_eps0);                          // :263
        em.advance_b(S.fields, real_t(0.5) * px, real_t(0.5) * py, real_t(0.5) * pz, S.nx, S.ny, S.nz, S.ng);  // :266
    } else {
        ES_Field_Solver_1D es;
        es.advance_b(S.fields, real_t(0.5) * px, real_t(0.5) * py, real_t(0.5) * pz, S.nx, S.ny, S.nz, S.ng);
        es.advance_e(S.fields, px, py, pz, S.nx, S.ny, S.nz, S.ng, (real_t)k.dt_eps0);
        es.advance_b(S.fields, real_t(0.5) * px, real_t(0.5) * py, real_t(0.5) * pz, S.nx, S.ny, S.nz, S.ng);
    }
}

void energies(RefSim& S, double* e, double* b) {
    if (S.solver == 0) {
        EM_Field_Solver em;
        *e = em.e_energy(S.fields, 0, 0, 0, S.nx, S.ny, S.nz, S.ng);
        *b = em.b_energy(S.fields, 0, 0, 0, S.nx, S.ny, S.nz, S.ng);
    } else {
        ES_Field_Solver_1D es;
        *e = es.e_energy(S.fields, 0, 0, 0, S.nx, S.ny, S.nz, S.ng);
        *b = 0.0;
    }
}

}  // namespace

namespace {
template <int M>
struct InterpIO {
    static void get_all(const interpolator_array_t& a, real_t* out, size_t n) {
        auto s = Cabana::slice<M>(a);
        for (size_t i = 0; i < n; ++i) out[i * 18 + M] = s(i);
        InterpIO<M + 1>::get_all(a, out, n);
    }
    static void put_all(interpolator_array_t& a, const real_t* in, size_t n) {
        auto s = Cabana::slice<M>(a);
        for (size_t i = 0; i < n; ++i) s(i) = in[i * 18 + M];
        InterpIO<M + 1>::put_all(a, in, n);
    }
};
template <>
struct InterpIO<18> {
    static void get_all(const interpolator_array_t&, real_t*, size_t) {}
    static void put_all(interpolator_array_t&, const real_t*, size_t) {}
};
}  // namespace

extern "C" {

int ref_real_bytes() { return (int)sizeof(real_t); }
int ref_vector_length() { return particle_list_t::vector_length; }
int ref_num_threads() { return Kokkos::DefaultExecutionSpace::concurrency(); }

void* ref_create(long nx, long ny, long nz, long ng, long np, int solver) {
    return new RefSim(nx, ny, nz, ng, np, solver);
}
void ref_destroy(void* h) { delete static_cast<RefSim*>(h); }
long ref_num_cells(void* h) { return (long)static_cast<RefSim*>(h)->num_cells; }
long ref_num_particles(void* h) { return (long)static_cast<RefSim*>(h)->np; }

void ref_set_particles(void* h, const real_t* dx, const real_t* dy, const real_t* dz, const real_t* ux,
                       const real_t* uy, const real_t* uz, const real_t* w, const int* cell) {
    RefSim& S = *static_cast<RefSim*>(h);
    put<PositionX>(S.particles, dx, S.np);
    put<PositionY>(S.particles, dy, S.np);
    put<PositionZ>(S.particles, dz, S.np);
    put<VelocityX>(S.particles, ux, S.np);
    put<VelocityY>(S.particles, uy, S.np);
    put<VelocityZ>(S.particles, uz, S.np);
    put<Weight>(S.particles, w, S.np);
    put<Cell_Index>(S.particles, cell, S.np);
}
void ref_get_particles(void* h, real_t* dx, real_t* dy, real_t* dz, real_t* ux, real_t* uy, real_t* uz, real_t* w,
                       int* cell) {
    RefSim& S = *static_cast<RefSim*>(h);
    get<PositionX>(S.particles, dx, S.np);
    get<PositionY>(S.particles, dy, S.np);
    get<PositionZ>(S.particles, dz, S.np);
    get<VelocityX>(S.particles, ux, S.np);
    get<VelocityY>(S.particles, uy, S.np);
    get<VelocityZ>(S.particles, uz, S.np);
    get<Weight>(S.particles, w, S.np);
    get<Cell_Index>(S.particles, cell, S.np);
}

// fields: 9 arrays of num_cells in FieldFields order (src/types.h:138-149)
void ref_set_fields(void* h, const real_t* const* f) {
    RefSim& S = *static_cast<RefSim*>(h);
    const size_t n = S.num_cells;
    put<FIELD_EX>(S.fields, f[0], n);  put<FIELD_EY>(S.fields, f[1], n);  put<FIELD_EZ>(S.fields, f[2], n);
    put<FIELD_CBX>(S.fields, f[3], n); put<FIELD_CBY>(S.fields, f[4], n); put<FIELD_CBZ>(S.fields, f[5], n);
    put<FIELD_JFX>(S.fields, f[6], n); put<FIELD_JFY>(S.fields, f[7], n); put<FIELD_JFZ>(S.fields, f[8], n);
}
void ref_get_fields(void* h, real_t* const* f) {
    RefSim& S = *static_cast<RefSim*>(h);
    const size_t n = S.num_cells;
    get<FIELD_EX>(S.fields, f[0], n);  get<FIELD_EY>(S.fields, f[1], n);  get<FIELD_EZ>(S.fields, f[2], n);
    get<FIELD_CBX>(S.fields, f[3], n); get<FIELD_CBY>(S.fields, f[4], n); get<FIELD_CBZ>(S.fields, f[5], n);
    get<FIELD_JFX>(S.fields, f[6], n); get<FIELD_JFY>(S.fields, f[7], n); get<FIELD_JFZ>(S.fields, f[8], n);
}

// interpolators as [cell][18] in InterpolatorFields order (src/types.h:64-84)
void ref_get_interpolators(void* h, real_t* out) {
    RefSim& S = *static_cast<RefSim*>(h);
    InterpIO<0>::get_all(S.interpolators, out, S.num_cells);
}
void ref_set_interpolators(void* h, const real_t* in) {
    RefSim& S = *static_cast<RefSim*>(h);
    InterpIO<0>::put_all(S.interpolators, in, S.num_cells);
}
// accumulators as [cell][3][4] (src/types.h:116-120)
void ref_get_accumulators(void* h, real_t* out) {
    RefSim& S = *static_cast<RefSim*>(h);
    for (size_t i = 0; i < S.num_cells; ++i)
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 4; ++k) out[(i * 3 + j) * 4 + k] = S.accumulators(i, j, k);
}
void ref_set_accumulators(void* h, const real_t* in) {
    RefSim& S = *static_cast<RefSim*>(h);
    for (size_t i = 0; i < S.num_cells; ++i)
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 4; ++k) S.accumulators(i, j, k) = in[(i * 3 + j) * 4 + k];
}

// ---- the reference's hot-path entry points, one call each -----------------
void ref_load_interpolator(void* h) {
    RefSim& S = *static_cast<RefSim*>(h);
    load_interpolator_array(S.fields, S.interpolators, S.nx, S.ny, S.nz, S.ng);
}
void ref_clear_accumulator(void* h) {
    RefSim& S = *static_cast<RefSim*>(h);
    clear_accumulator_array(S.fields, S.accumulators, S.nx, S.ny, S.nz);
}
void ref_push(void* h, const Consts* k) {
    RefSim& S = *static_cast<RefSim*>(h);
    push(S.particles, S.interpolators, (real_t)k->qdt_2mc, (real_t)k->cdt_dx, (real_t)k->cdt_dy, (real_t)k->cdt_dz,
         (real_t)k->qsp, S.scatter_add, S.grid, S.nx, S.ny, S.nz, S.ng, deck.BOUNDARY_TYPE);
    Kokkos::Experimental::contribute(S.accumulators, S.scatter_add);
    S.scatter_add.reset_except(S.accumulators);
}
void ref_unload_accumulator(void* h, const Consts* k) {
    RefSim& S = *static_cast<RefSim*>(h);
    unload_accumulator_array(S.fields, S.accumulators, S.nx, S.ny, S.nz, S.ng, (real_t)k->dx, (real_t)k->dy,
                             (real_t)k->dz, (real_t)k->dt);
}
void ref_advance_b(void* h, double px, double py, double pz) {
    RefSim& S = *static_cast<RefSim*>(h);
    if (S.solver == 0) {
        EM_Field_Solver em;
        em.advance_b(S.fields, (real_t)px, (real_t)py, (real_t)pz, S.nx, S.ny, S.nz, S.ng);
    }
}
void ref_advance_e(void* h, double px, double py, double pz, double dt_eps0) {
    RefSim& S = *static_cast<RefSim*>(h);
    if (S.solver == 0) {
        EM_Field_Solver em;
        em.advance_e(S.fields, (real_t)px, (real_t)py, (real_t)pz, S.nx, S.ny, S.nz, S.ng, (real_t)dt_eps0);
    } else {
        ES_Field_Solver_1D es;
        es.advance_e(S.fields, (real_t)px, (real_t)py, (real_t)pz, S.nx, S.ny, S.nz, S.ng, (real_t)dt_eps0);
    }
}
void ref_uncenter(void* h, double qdt_2mc) {
    RefSim& S = *static_cast<RefSim*>(h);
    uncenter_particles(S.particles, S.interpolators, (real_t)qdt_2mc);
}
void ref_energies(void* h, double* e, double* b) { energies(*static_cast<RefSim*>(h), e, b); }

// n steps of example.cpp:216-271 without its ASCII dumps; if `en` is non-null
// it receives (e,b) after every step: en[2*s], en[2*s+1].
void ref_run(void* h, const Consts* k, long nsteps, double* en) {
    RefSim& S = *static_cast<RefSim*>(h);
    for (long s = 0; s < nsteps; ++s) {
        one_step(S, *k);
        if (en) energies(S, en + 2 * s, en + 2 * s + 1);
    }
}

// ---- deck seam: the linked deck's parameters and initial conditions --------
// out[0..23]: nx ny nz ng nppc num_steps num_particles num_cells | dt c eps qsp
// me n0 Npe Ne v0 | len_x len_y len_z dx dy dz | perform_uncenter
void ref_deck_params(double* out) {
    deck.derive_params();  // example.cpp:61
    int i = 0;
    out[i++] = (double)deck.nx; out[i++] = (double)deck.ny; out[i++] = (double)deck.nz;
    out[i++] = (double)deck.num_ghosts; out[i++] = (double)deck.nppc; out[i++] = (double)deck.num_steps;
    out[i++] = (double)deck.num_particles; out[i++] = (double)deck.num_cells;
    out[i++] = deck.dt; out[i++] = deck.c; out[i++] = deck.eps; out[i++] = deck.qsp;
    out[i++] = deck.me; out[i++] = deck.n0; out[i++] = deck.Npe; out[i++] = deck.Ne; out[i++] = deck.v0;
    out[i++] = deck.len_x; out[i++] = deck.len_y; out[i++] = deck.len_z;
    out[i++] = deck.dx; out[i++] = deck.dy; out[i++] = deck.dz;
    out[i++] = deck.perform_uncenter ? 1.0 : 0.0;
}

// Driver-side constants exactly as example/example.cpp:77-113,179-181 derives
// them (in real_t), widened to double for transport.
void ref_deck_consts(Consts* k, double* dxp_out, double* we_out) {
    deck.derive_params();
    const int npc = deck.nppc;
    const int nx = deck.nx, ny = deck.ny, nz = deck.nz;
    real_t dxp = 2.f / (npc);
    const real_t dx = deck.dx, dy = deck.dy, dz = deck.dz;
    real_t dt = deck.dt, c = deck.c, eps0 = deck.eps;
    real_t Npe = deck.Npe;
    size_t Ne = deck.Ne;
    real_t qsp = deck.qsp, me = deck.me;
    real_t qdt_2mc = qsp * dt / (2 * me * c);
    real_t cdt_dx = c * dt / dx, cdt_dy = c * dt / dy, cdt_dz = c * dt / dz;
    real_t dt_eps0 = dt / eps0;
    real_t frac = 1.0f;
    real_t we = (real_t)Npe / (real_t)Ne;
    const real_t px = (nx > 1) ? frac * c * dt / dx : 0;
    const real_t py = (ny > 1) ? frac * c * dt / dy : 0;
    const real_t pz = (nz > 1) ? frac * c * dt / dz : 0;
    k->qdt_2mc = qdt_2mc; k->cdt_dx = cdt_dx; k->cdt_dy = cdt_dy; k->cdt_dz = cdt_dz; k->qsp = qsp;
    k->dx = dx; k->dy = dy; k->dz = dz; k->dt = dt;
    k->px = px; k->py = py; k->pz = pz; k->dt_eps0 = dt_eps0;
    *dxp_out = dxp;
    *we_out = we;
}

// Build a simulation from the linked deck and run its initializers the way
// main() does (example.cpp:121-168, 204-213).
void* ref_create_from_deck(int solver) {
    Consts k; double dxp, we;
    ref_deck_consts(&k, &dxp, &we);
    RefSim* S = new RefSim(deck.nx, deck.ny, deck.nz, deck.num_ghosts, deck.num_particles, solver);
    deck.initialize_particles(S->particles, deck.nx, deck.ny, deck.nz, deck.num_ghosts, (real_t)dxp, deck.nppc,
                              (real_t)we, deck.v0);
    deck.initialize_fields(S->fields, deck.nx, deck.ny, deck.nz, deck.num_ghosts, deck.len_x, deck.len_y,
                           deck.len_z, deck.dx, deck.dy, deck.dz);
    if (deck.perform_uncenter) {
        load_interpolator_array(S->fields, S->interpolators, S->nx, S->ny, S->nz, S->ng);
        uncenter_particles(S->particles, S->interpolators, (real_t)k.qdt_2mc);
    }
    return S;
}

}  // extern "C"
