// cabanapic_b200 C++ host facade -- the deck interface.
//
// Source-compatible with the reference's src/input/deck.h (:9-12 Boundary, :15-159 the three
// hook classes, :161-348 _Input_Deck, :415-468 Input_Deck and the global `deck`): a deck written for
// CabanaPIC (decks/*.cxx, tests/energy_comparison/2stream-em.cxx) compiles against this header
// unchanged.  Initialisers run on the HOST mirror of the device-backed arrays, serially and in
// (s,i) order (decks may call rand(), decks/dioctron_3d.cxx:112-131); the arrays are uploaded when
// the first hot-path call needs them.
#ifndef CABANAPIC_B200_INPUT_DECK_H
#define CABANAPIC_B200_INPUT_DECK_H

#include <cmath>
#include <cstddef>
#include <iostream>

#include "types.h"

enum Boundary { Reflect = 0, Periodic };

// End-of-run hook (correctness checks, timing dumps).
class Run_Finalizer {
   public:
    virtual ~Run_Finalizer() {}
    virtual void finalize() {}
};

// Default field initialiser: E = cB = 0 everywhere.
class Field_Initializer {
   public:
    using real_ = real_t;
    Field_Initializer() {}
    virtual ~Field_Initializer() {}
    virtual void init(field_array_t& fields, size_t, size_t, size_t, size_t, real_, real_, real_, real_, real_, real_) {
        std::cout << "Default field init" << std::endl;
        auto ex = Cabana::slice<FIELD_EX>(fields);
        auto ey = Cabana::slice<FIELD_EY>(fields);
        auto ez = Cabana::slice<FIELD_EZ>(fields);
        auto bx = Cabana::slice<FIELD_CBX>(fields);
        auto by = Cabana::slice<FIELD_CBY>(fields);
        auto bz = Cabana::slice<FIELD_CBZ>(fields);
        Kokkos::parallel_for("zero_fields()", fields.size(), KOKKOS_LAMBDA(const int i) {
            ex(i) = 0.0; ey(i) = 0.0; ez(i) = 0.0; bx(i) = 0.0; by(i) = 0.0; bz(i) = 0.0;
        });
    }
};

// Default particle initialiser: two cold counter-streaming beams along y, every pair of particles
// sharing a slot, with a 1e-4 sinusoidal perturbation of ux (reference src/input/deck.h:111-153;
// the double-precision intermediates below follow the reference's literal types, which is what
// makes the resulting float state bit-identical -- checked in tests/test_decks.py).
class Particle_Initializer {
   public:
    using real_ = real_t;
    Particle_Initializer() {}
    virtual ~Particle_Initializer() {}
    virtual void init(particle_list_t& particles, size_t nx, size_t ny, size_t, size_t, real_ dxp, size_t nppc, real_ w,
                      real_ v0, real_, real_, real_) {
        std::cout << "Default particle init" << std::endl;
        auto px = Cabana::slice<PositionX>(particles);
        auto py = Cabana::slice<PositionY>(particles);
        auto pz = Cabana::slice<PositionZ>(particles);
        auto ux = Cabana::slice<VelocityX>(particles);
        auto uy = Cabana::slice<VelocityY>(particles);
        auto uz = Cabana::slice<VelocityZ>(particles);
        auto weight = Cabana::slice<Weight>(particles);
        auto cell = Cabana::slice<Cell_Index>(particles);
        printf("dxp = %e \n", dxp);
        printf("part list len = %ld \n", (long)particles.size());
        auto fill = KOKKOS_LAMBDA(const int s, const int i) {
            const size_t k = size_t(s) * particle_list_t::vector_length + i;   // particle number
            const size_t pair = k / 2;                                          // both beams share a slot
            const int sign = (k % 2 == 0) ? 1 : -1;
            const int slot = int((2 * pair) % nppc);                            // position slot inside the cell
            const int cell_no = int(2 * pair / nppc);                           // cell along the y line, ghosts not counted
            const real_ y = slot * dxp + 0.5 * dxp - 1.0;
            px.access(s, i) = 0.0;
            py.access(s, i) = y;
            pz.access(s, i) = 0.0;
            weight.access(s, i) = w;
            cell.access(s, i) = cell_no * (nx + 2) + (nx + 2) * (ny + 2) + (nx + 2) + 1;   // one ghost layer hard-wired
            const real_ gam = 1.0 / sqrt(1.0 - v0 * v0);
            const real_t ripple = 0.0001 * sin(2.0 * 3.1415926 * ((y + 1.0 + cell_no * 2) / (2 * ny)));
            ux.access(s, i) = sign * v0 * gam * (1.0 + ripple * sign);
            uy.access(s, i) = 0;
            uz.access(s, i) = 0;
        };
        Cabana::SimdPolicy<particle_list_t::vector_length, ExecutionSpace> policy(0, particles.size());
        Cabana::simd_parallel_for(policy, fill, "init()");
    }
};

class _Input_Deck {
   public:
    using real_ = real_t;
    // heap-allocated hooks, replaced (and the defaults leaked) by custom decks, as in the reference
    Particle_Initializer* particle_initer;
    Field_Initializer* field_initer;
    Run_Finalizer* run_finalizer;

    _Input_Deck() : particle_initer(new Particle_Initializer), field_initer(new Field_Initializer), run_finalizer(new Run_Finalizer) {}

    // Courant length of the grid; axes with a single cell do not count.
    static real_ courant_length(real_ lx, real_ ly, real_ lz, size_t nx, size_t ny, size_t nz) {
        real_ inv, sum = 0;
        if (nx > 1) inv = nx / lx, sum += inv * inv;
        if (ny > 1) inv = ny / ly, sum += inv * inv;
        if (nz > 1) inv = nz / lz, sum += inv * inv;
        return sqrt(1 / sum);
    }

    void finalize() { run_finalizer->finalize(); }

    void initialize_particles(particle_list_t& particles, size_t nx, size_t ny, size_t nz, size_t ng, real_ dxp, size_t nppc,
                              real_ w, real_ v0) {
        particle_initer->init(particles, nx, ny, nz, ng, dxp, nppc, w, v0, len_x_global, len_y_global, len_z_global);
    }
    void initialize_fields(field_array_t& fields, size_t nx, size_t ny, size_t nz, size_t ng, real_ Lx, real_ Ly, real_ Lz,
                           real_ dx, real_ dy, real_ dz) {
        field_initer->init(fields, nx, ny, nz, ng, Lx, Ly, Lz, dx, dy, dz);
    }

    // normalisation
    real_ de = 1.0, ec = 1.0, me = 1.0, mu = 1.0, c = 1.0, eps = 1.0;
    real_ qsp = -ec;
    real_ n0 = 1.0;
    size_t num_species = 1;
    // grid
    size_t nx = 16, ny = 1, nz = 1;
    size_t num_ghosts = 1;
    size_t nppc = 1;
    real_ dt = 1.0;
    int num_steps = 2;
    real_ len_x_global = 1.0, len_y_global = 1.0, len_z_global = 1.0;
    real_ Npe = -1;     // physical electrons in the box (derived from n0 when negative)
    real_ Ne = -1;      // macro-particles (derived when negative)
    real_ v0 = 1.0;     // drift velocity
    Boundary BOUNDARY_TYPE = Boundary::Periodic;

    // derived by derive_params()
    real_ dx, dy, dz;
    real_ len_x, len_y, len_z;
    size_t num_cells;            // includes the ghost cells
    long num_particles = -1;
    bool perform_uncenter = false;

    void print_run_details() {
        std::cout << "#~~~ Run Specifications ~~~ " << std::endl;
        std::cout << "#Nx: " << nx << " Ny: " << ny << " Nz: " << nz << " Num Ghosts: " << num_ghosts
                  << ". Cells Total: " << num_cells << std::endl;
        std::cout << "#Len X: " << len_x << " Len Y: " << len_y << " Len Z: " << len_z
                  << " number of ghosts: " << num_ghosts << std::endl;
        std::cout << "#Approx Particle Count: " << num_particles << " (nppc: " << nppc << ")" << std::endl;
        std::cout << "#~~~~~~~~~~~~~~~~~~~~~~~~~~ " << std::endl << std::endl;
    }

    void derive_params() {
        len_x = len_x_global; len_y = len_y_global; len_z = len_z_global;
        dx = len_x / nx; dy = len_y / ny; dz = len_z / nz;
        const size_t g2 = 2 * num_ghosts;
        num_cells = (nx + g2) * (ny + g2) * (nz + g2);
        if (num_particles < 0) {
            num_particles = nx * ny * nz * nppc;
            if (Ne < 0) Ne = num_particles;
        }
        if (Npe < 0) Npe = n0 * len_x_global * len_y_global * len_z_global;
    }
};

#ifdef USER_INPUT_DECK
#define STRINGIFY(s) #s
#define EXPAND(s) STRINGIFY(s)
// the deck's constructor lives in a separately compiled decks/*.cxx
class Input_Deck : public _Input_Deck {
   public:
    Input_Deck();
};
#else
// built-in deck: the 1x32x1 electromagnetic two-stream problem of the regression test
class Input_Deck : public _Input_Deck {
   public:
    Input_Deck() {
        nx = 1; ny = 32; nz = 1;
        num_steps = 6000;
        nppc = 100;
        v0 = 0.0866025403784439;
        real_ gam = 1.0 / sqrt(1.0 - v0 * v0);
        len_x_global = 1.0;
        len_y_global = 0.628318530717959 * (gam * sqrt(gam));
        len_z_global = 1.0;
        dt = 0.99 * courant_length(len_x_global, len_y_global, len_z_global, nx, ny, nz) / c;
        n0 = 2.0;   // two beams, each with plasma frequency 1
    }
};
#endif

extern Input_Deck deck;

#include "device.h"   // the runtime behind the device-backed arrays (needs `deck`)

#endif  // CABANAPIC_B200_INPUT_DECK_H
