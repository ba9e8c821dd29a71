// cbnpic -- the CabanaPIC driver on the B200-native hot path.
//
// Same start-up and time loop as the reference's example/example.cpp (:44-290): derive the deck
// parameters and step constants in real_t, run the deck's initialisers, then per step
//   load_interpolator_array -> clear_accumulator_array -> push -> contribute -> unload_accumulator_array
//   -> advance_b(1/2) -> advance_e -> advance_b(1/2) -> dump_energies
// through the facade headers in include/cabanapic/src (every call lands in a hand-written sm_100a
// kernel through the C ABI).  Differences from the reference driver, all about I/O:
//   * the per-step ASCII dumps `partloc` / `ex1d` (example.cpp:274-277, a full device->host copy and
//     an fprintf per particle every step) are opt-in: CPIC_DUMP=1 (CPIC_DUMP_FIELDS=0 keeps partloc only);
//   * energies.txt is started afresh instead of appended to;
//   * CPIC_STEPS overrides deck.num_steps, CPIC_SORT_INTERVAL switches the periodic sort on,
//     CPIC_ENERGY_INTERVAL thins the energy dumps; a wall-clock summary is printed at the end.
#include <Cabana_Core.hpp>

#include <chrono>
#include <cstdlib>
#include <iostream>

#include "types.h"
#include "helpers.h"
#include "fields.h"
#include "accumulator.h"
#include "interpolator.h"
#include "uncenter_p.h"
#include "push.h"
#include "input/deck.h"

Input_Deck deck;

static int env_int(const char* name, int fallback) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : fallback;
}

int main(int argc, char* argv[]) {
    Kokkos::ScopeGuard scope_guard(argc, argv);
    {
        deck.derive_params();
        deck.print_run_details();
        const int nx = deck.nx, ny = deck.ny, nz = deck.nz, ng = deck.num_ghosts;
#ifdef ES_FIELD_SOLVER
        if (ny > 1 || nz > 1) {
            std::cerr << "Error: ES Field solver supports 1D only.\n";
            return -1;
        }
        std::cout << "Created ES Solver (1D only)" << std::endl;
#else
        std::cout << "Created EM Solver" << std::endl;
#endif
        // step constants, all in real_t (example.cpp:77-113)
        const real_t dxp = 2.f / deck.nppc;
        const real_t dx = deck.dx, dy = deck.dy, dz = deck.dz, dt = deck.dt, c = deck.c;
        const real_t qsp = deck.qsp, me = deck.me, eps0 = deck.eps;
        const real_t Npe = deck.Npe;
        const size_t Ne = deck.Ne;
        const real_t qdt_2mc = qsp * dt / (2 * me * c);
        const real_t cdt_dx = c * dt / dx, cdt_dy = c * dt / dy, cdt_dz = c * dt / dz;
        const real_t dt_eps0 = dt / eps0;
        const real_t frac = 1.0f;
        const real_t we = (real_t)Npe / (real_t)Ne;
        const real_t px = (nx > 1) ? frac * c * dt / dx : 0;
        const real_t py = (ny > 1) ? frac * c * dt / dy : 0;
        const real_t pz = (nz > 1) ? frac * c * dt / dz : 0;
        const size_t num_particles = deck.num_particles;
        printf("#nppc %d nx %d ny %d nz %d  Ne %ld Npe %e we %e\n", (int)deck.nppc, nx, ny, nz, (long)Ne, Npe, we);
        printf("#c %e dt %e dx %e cdt_dx %e qdt_2mc %e\n", c, dt, dx, cdt_dx, qdt_2mc);

        particle_list_t particles("particles", num_particles);
        deck.initialize_particles(particles, nx, ny, nz, ng, dxp, deck.nppc, we, deck.v0);
        grid_t* grid = new grid_t();

        interpolator_array_t interpolators("interpolator", deck.num_cells);
        accumulator_array_t accumulators("accumulator", deck.num_cells);
        auto scatter_add = Kokkos::Experimental::create_scatter_view(accumulators);
        field_array_t fields("fields", deck.num_cells);
        initialize_interpolator(interpolators);
#ifdef ES_FIELD_SOLVER
        Field_Solver<ES_Field_Solver_1D> field_solver(fields);
#else
        Field_Solver<EM_Field_Solver> field_solver(fields);
#endif
        deck.initialize_fields(fields, nx, ny, nz, ng, deck.len_x, deck.len_y, deck.len_z, dx, dy, dz);
        const Boundary boundary = deck.BOUNDARY_TYPE;

        const int num_steps = env_int("CPIC_STEPS", deck.num_steps);
        const int energy_interval = env_int("CPIC_ENERGY_INTERVAL", 1);
        const bool dump = env_int("CPIC_DUMP", 0) != 0;
        FILE* fptr = dump ? fopen("partloc", "w") : nullptr;
        const bool dump_fields = dump && env_int("CPIC_DUMP_FIELDS", 1) != 0;      // (ex1d is nx lines per step)
        FILE* fpfd = dump_fields ? fopen("ex1d", "w") : nullptr;
        std::remove("energies.txt");
        if (dump) {
            fprintf(fptr, "#step=0\n0 ");
            dump_particles(fptr, particles, 0, 0, 0, dx, dy, dz, nx, ny, nz, ng);
        }
        printf("#num_step = %d\n", num_steps);

        if (deck.perform_uncenter) {
            load_interpolator_array(fields, interpolators, nx, ny, nz, ng);
            uncenter_particles(particles, interpolators, qdt_2mc);
        }

        const auto t0 = std::chrono::steady_clock::now();
        auto t1 = t0;      // after the first step (which carries the one-time upload of the host-initialised state)
        for (int step = 1; step <= num_steps; step++) {
            if (step == 2) { cpic_sync(cabanapic::Runtime::get().ctx()); t1 = std::chrono::steady_clock::now(); }
            load_interpolator_array(fields, interpolators, nx, ny, nz, ng);
            clear_accumulator_array(fields, accumulators, nx, ny, nz);
            push(particles, interpolators, qdt_2mc, cdt_dx, cdt_dy, cdt_dz, qsp, scatter_add, grid, nx, ny, nz, ng, boundary);
            Kokkos::Experimental::contribute(accumulators, scatter_add);
            scatter_add.reset_except(accumulators);
            unload_accumulator_array(fields, accumulators, nx, ny, nz, ng, dx, dy, dz, dt);
            field_solver.advance_b(fields, real_t(0.5) * px, real_t(0.5) * py, real_t(0.5) * pz, nx, ny, nz, ng);
            field_solver.advance_e(fields, px, py, pz, nx, ny, nz, ng, dt_eps0);
            field_solver.advance_b(fields, real_t(0.5) * px, real_t(0.5) * py, real_t(0.5) * pz, nx, ny, nz, ng);
            if (energy_interval > 0 && step % energy_interval == 0)
                dump_energies(field_solver, fields, step, step * dt, px, py, pz, nx, ny, nz, ng);
            if (dump_fields) {
                fprintf(fpfd, "#step=%d\n", step);
                field_solver.dump_fields(fpfd, fields, 0, 0, 0, dx, dy, dz, nx, ny, nz, ng);
            }
            if (dump) {
                fprintf(fptr, "#step=%d\n%e ", step, step * dt);
                dump_particles(fptr, particles, 0, 0, 0, dx, dy, dz, nx, ny, nz, ng);
            }
        }
        cpic_sync(cabanapic::Runtime::get().ctx());
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("#%d steps of %ld particles in %.3f s: %.3e particle-steps/s (incl. energy dumps)\n", num_steps,
               (long)num_particles, sec, num_steps * (double)num_particles / sec);
        if (num_steps > 1) {
            const double s1 = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
            printf("#steady: %d steps after the first (which uploads the initial state) in %.3f s: %.3e particle-steps/s\n",
                   num_steps - 1, s1, (num_steps - 1) * (double)num_particles / s1);
        }
        if (dump) fclose(fptr);
        if (dump_fields) fclose(fpfd);
        delete grid;
    }
    deck.finalize();
    cabanapic::Runtime::get().destroy();
    return 0;
}
