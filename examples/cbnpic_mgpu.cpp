// cbnpic_mgpu -- the deck-driven CabanaPIC run on SEVERAL GPUs, host side in C++ over the multi-GPU C ABI
// (include/cabanapic_b200_mgpu.h: z-slabs with NVLink ghost-plane exchange and particle migration for large 3-D
// grids, replicated grid + ncclAllReduce of the accumulator for small ones).  One process per GPU, no mpirun:
//
//     for r in 0 1 2 3; do CPIC_WORLD=4 CPIC_RANK=$r CPIC_MGPU_ID_FILE=/tmp/id.$$ ./cbnpic_mgpu_weibel_3d & done; wait
//
// (torchrun's RANK / WORLD_SIZE are honoured too).  Start-up is the reference's (example/example.cpp:44-213): derive
// the deck parameters and step constants in real_t, run the deck's particle and field initialisers on the host for
// the WHOLE box; then every rank keeps its share (slab mode: the particles and field planes of the z-planes it owns,
// cells re-based; replicated mode: a contiguous slice of the particle list) and the time loop of :216-271 is
// cpic_mgpu_step.  Rank 0 writes energies.txt in the reference's format (src/fields.h:752-761) from the
// all-reduced energies.  CPIC_STEPS, CPIC_ENERGY_INTERVAL, CPIC_MGPU_MODE (slab | replicated | auto), CPIC_GRAPH=0.
#include <Cabana_Core.hpp>

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>

#include "types.h"
#include "input/deck.h"
#include "cabanapic_b200_mgpu.h"

Input_Deck deck;

static int env_int(const char* a, const char* b, int fallback) {
    const char* e = std::getenv(a);
    if (!e && b) e = std::getenv(b);
    return e ? std::atoi(e) : fallback;
}
#define CK(call)                                                                                      \
    do {                                                                                              \
        const int rc_ = (call);                                                                       \
        if (rc_ != CPIC_OK) {                                                                         \
            std::fprintf(stderr, "[rank %d] %s failed (%d): %s | %s\n", rank, #call, rc_, cpic_mgpu_last_error(mg), \
                         mg ? cpic_last_error(cpic_mgpu_context(mg)) : "");                          \
            return 1;                                                                                 \
        }                                                                                             \
    } while (0)

template <std::size_t... I>
static void member_ptrs(const particle_list_t& p, const void** m, std::index_sequence<I...>) {
    const void* t[] = {p.host().template member_data<I>()...};
    for (int k = 0; k < 8; ++k) m[k] = t[k];
}
template <std::size_t... I>
static void field_ptrs(const field_array_t& f, const real_t** m, std::index_sequence<I...>) {
    const real_t* t[] = {f.host().template member_data<I>()...};
    for (int k = 0; k < 9; ++k) m[k] = t[k];
}

int main(int argc, char* argv[]) {
    Kokkos::ScopeGuard scope_guard(argc, argv);
    const int world = env_int("CPIC_WORLD", "WORLD_SIZE", 1), rank = env_int("CPIC_RANK", "RANK", 0);
    cpic_mgpu* mg = nullptr;
    deck.derive_params();
    if (rank == 0) deck.print_run_details();
    const int nx = deck.nx, ny = deck.ny, nz = deck.nz, ng = deck.num_ghosts;
    // step constants, all in real_t (example.cpp:77-113)
    const real_t dxp = 2.f / deck.nppc;
    const real_t dx = deck.dx, dy = deck.dy, dz = deck.dz, dt = deck.dt, c = deck.c;
    const real_t qsp = deck.qsp, me = deck.me, eps0 = deck.eps;
    const real_t we = (real_t)deck.Npe / (real_t)deck.Ne;
    cpic_consts k{};
    k.qdt_2mc = qsp * dt / (2 * me * c);
    k.cdt_dx = c * dt / dx; k.cdt_dy = c * dt / dy; k.cdt_dz = c * dt / dz; k.qsp = qsp;
    k.dx = dx; k.dy = dy; k.dz = dz; k.dt = dt;
    k.px = (nx > 1) ? (real_t)(c * dt / dx) : 0; k.py = (ny > 1) ? (real_t)(c * dt / dy) : 0; k.pz = (nz > 1) ? (real_t)(c * dt / dz) : 0;
    k.dt_eps0 = dt / eps0;

    // the deck's initialisers, on the host, for the whole box (dioctron's rand() stream depends on the serial order)
    const size_t np = deck.num_particles;
    particle_list_t particles("particles", np);
    deck.initialize_particles(particles, nx, ny, nz, ng, dxp, deck.nppc, we, deck.v0);
    field_array_t fields("fields", deck.num_cells);
    {
        const real_t* f[9];
        field_ptrs(fields, f, std::make_index_sequence<9>{});
        for (int m = 0; m < 9; ++m) std::memset(const_cast<real_t*>(f[m]), 0, deck.num_cells * sizeof(real_t));      // Field_Solver ctor, src/fields.h:279-315
    }
    deck.initialize_fields(fields, nx, ny, nz, ng, deck.len_x, deck.len_y, deck.len_z, dx, dy, dz);

    cpic_params gp{};
    gp.nx = nx; gp.ny = ny; gp.nz = nz; gp.ng = ng;
    gp.real_bytes = (int32_t)sizeof(real_t);
    gp.solver = CPIC_SOLVER_EM;
    gp.boundary = CPIC_BOUNDARY_PERIODIC;
    gp.device = env_int("CPIC_DEVICE", "LOCAL_RANK", rank);
    gp.fp_mode = CPIC_FP_STRICT;
    gp.deposit_mode = CPIC_DEPOSIT_AUTO;
    gp.enable_sort = 1;
    gp.max_particles = (int64_t)(np / world + np / (4 * world) + 4096);      // room for the migration imbalance
    int mode = CPIC_MGPU_AUTO;
    if (const char* e = std::getenv("CPIC_MGPU_MODE")) mode = !std::strcmp(e, "slab") ? CPIC_MGPU_SLAB : (!std::strcmp(e, "replicated") ? CPIC_MGPU_REPLICATED : CPIC_MGPU_AUTO);
    unsigned char id[CPIC_MGPU_ID_BYTES] = {0};
    if (world > 1) {
        const char* idf = std::getenv("CPIC_MGPU_ID_FILE");
        CK(cpic_mgpu_bootstrap_file(idf ? idf : "/tmp/cabanapic_b200.nccl_id", rank, world, 120.0, id));
    }
    CK(cpic_mgpu_create(&gp, rank, world, id, mode, 0, &mg));
    int32_t z0 = 0, nzl = nz;
    CK(cpic_mgpu_layout(mg, &mode, &z0, &nzl));
    cpic_ctx* ctx = cpic_mgpu_context(mg);
    if (rank == 0) std::printf("#multi-GPU: %d ranks, %s mode\n", world, mode == CPIC_MGPU_SLAB ? "z-slab" : "replicated");

    // ---- this rank's share
    const long long plane = (long long)(nx + 2 * ng) * (ny + 2 * ng);
    const void* pm[8];
    member_ptrs(particles, pm, std::make_index_sequence<8>{});
    const real_t* pr[7];
    for (int m = 0; m < 7; ++m) pr[m] = static_cast<const real_t*>(pm[m]);
    const int* pcell = static_cast<const int*>(pm[7]);
    std::vector<real_t> lp[7];
    std::vector<int32_t> lcell;
    if (mode == CPIC_MGPU_SLAB) {
        for (size_t n = 0; n < np; ++n) {
            const long long iz = pcell[n] / plane;
            if (iz < z0 + 1 || iz > z0 + nzl) continue;
            for (int m = 0; m < 7; ++m) lp[m].push_back(pr[m][n]);
            lcell.push_back((int32_t)(pcell[n] - (long long)z0 * plane));
        }
    } else {
        const size_t lo = np * (size_t)rank / world, hi = np * (size_t)(rank + 1) / world;
        for (int m = 0; m < 7; ++m) lp[m].assign(pr[m] + lo, pr[m] + hi);
        lcell.assign(pcell + lo, pcell + hi);
    }
    CK(cpic_upload_particles(ctx, lp[0].data(), lp[1].data(), lp[2].data(), lp[3].data(), lp[4].data(), lp[5].data(), lp[6].data(),
                             lcell.data(), (int64_t)lcell.size()));
    {
        const real_t* gf[9];
        field_ptrs(fields, gf, std::make_index_sequence<9>{});
        std::vector<real_t> lf[9];
        const void* up[9];
        for (int m = 0; m < 9; ++m) {
            if (mode == CPIC_MGPU_SLAB) {
                lf[m].resize((size_t)plane * (nzl + 2));
                for (int j = 0; j < nzl + 2; ++j) {      // local plane j <- global plane (ghost planes hold the neighbours' data)
                    long long zg = (z0 + j - 1 + nz) % nz + 1;
                    if (j == 0 && z0 == 0) zg = 0;
                    if (j == nzl + 1) zg = z0 + nzl + 1;
                    std::memcpy(lf[m].data() + (size_t)j * plane, gf[m] + (size_t)zg * plane, plane * sizeof(real_t));
                }
                up[m] = lf[m].data();
            } else {
                up[m] = gf[m];
            }
        }
        CK(cpic_upload_fields(ctx, up));
    }
    if (deck.perform_uncenter) {
        CK(cpic_load_interpolator_array(ctx));
        CK(cpic_uncenter_particles(ctx, k.qdt_2mc));
    }

    // ---- time loop (example.cpp:216-271)
    const int num_steps = env_int("CPIC_STEPS", nullptr, deck.num_steps);
    const int energy_interval = env_int("CPIC_ENERGY_INTERVAL", nullptr, 1);
    const int use_graph = env_int("CPIC_GRAPH", nullptr, 1);
    FILE* efile = (rank == 0 && energy_interval > 0) ? std::fopen("energies.txt", "w") : nullptr;
    if (rank == 0) std::printf("#num_step = %d\n", num_steps);
    const int chunk = energy_interval > 0 ? energy_interval : num_steps;
    CK(cpic_mgpu_sync(mg));
    const auto t0 = std::chrono::steady_clock::now();
    for (int step = 0; step < num_steps;) {
        const int n = std::min(chunk, num_steps - step);
        CK(cpic_mgpu_step(mg, &k, n, CPIC_SORT_FUSED, use_graph));
        step += n;
        if (energy_interval > 0) {
            double e = 0, b = 0;
            CK(cpic_mgpu_energies(mg, &e, &b));
            if (efile) std::fprintf(efile, "%d %g %g %g\n", step, (double)(real_t)(step * dt), (double)(real_t)e, (double)(real_t)b);
        }
    }
    CK(cpic_mgpu_sync(mg));
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (efile) std::fclose(efile);
    double dg[8];
    CK(cpic_mgpu_state_digest(mg, dg));
    int64_t mig[2] = {0, 0};
    CK(cpic_mgpu_migration_counts(mg, mig));
    std::printf("#rank %d: sent %lld particles down, %lld up\n", rank, (long long)mig[0], (long long)mig[1]);
    if (rank == 0) {
        std::printf("#%d steps of %ld particles on %d GPUs in %.3f s: %.3e particle-steps/s%s\n", num_steps, (long)np, world, sec,
                    num_steps * (double)np / sec, cpic_mgpu_used_graph(mg) ? " (CUDA-graph replay)" : "");
        std::printf("#digest: particles %.0f weight %.9g not-interior %.0f offsets-out %.0f kinetic %.9g E %.9g B %.9g migrated %.0f\n",
                    dg[0], dg[1], dg[2], dg[3], dg[4], dg[5], dg[6], dg[7]);
    }
    cpic_mgpu_destroy(mg);
    return 0;
}
