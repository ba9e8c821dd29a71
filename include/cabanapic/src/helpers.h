// Facade of the reference's src/helpers.h: the ASCII diagnostics a driver may call.
#ifndef CABANAPIC_B200_HELPERS_H
#define CABANAPIC_B200_HELPERS_H
#include "logger.h"
#include "Cabana_ExecutionPolicy.hpp"
#include "Cabana_Parallel.hpp"
#include "Cabana_DeepCopy.hpp"
#include "input/deck.h"

// ghost-free cell number -> voxel index incl. ghosts (reference :13-28)
inline int allow_for_ghosts(int pre_ghost) {
    size_t ix, iy, iz;
    RANK_TO_INDEX(pre_ghost, ix, iy, iz, deck.nx, deck.ny);
    return VOXEL(ix, iy, iz, deck.nx, deck.ny, deck.nz, deck.num_ghosts);
}

// one line `x v x v ...` over all particles: the `partloc` format (reference :31-94).  Downloads the
// particle store if the device copy is newer -- keep it out of timed loops.
inline void dump_particles(FILE* fp, const particle_list_t particles, const real_t xmin, const real_t, const real_t, const real_t dx,
                           const real_t, const real_t, size_t nx, size_t ny, size_t, size_t ng) {
    const bool device_was_current = particles.residency().device_valid;
    auto position_x = Cabana::slice<PositionX>(particles);
    auto velocity_x = Cabana::slice<VelocityX>(particles);
    auto cell = Cabana::slice<Cell_Index>(particles);
    particles.residency().device_valid = device_was_current;      // read-only use of the mirror
    for (size_t i = 0; i < particles.size(); i++) {
        size_t ix, iy, iz;
        const int ii = cell(i);
        RANK_TO_INDEX(ii, ix, iy, iz, nx + 2 * ng, ny + 2 * ng);
        (void)iy; (void)iz;
        const real_t x = xmin + (ix - 1 + (position_x(i) + 1.0) * 0.5) * dx;
        fprintf(fp, "%e  %e ", x, velocity_x(i));
    }
    fprintf(fp, "\n");
}

inline void print_fields(const field_array_t& fields) {
    const bool device_was_current = fields.residency().device_valid;
    auto ex = Cabana::slice<FIELD_EX>(fields);   auto ey = Cabana::slice<FIELD_EY>(fields);   auto ez = Cabana::slice<FIELD_EZ>(fields);
    auto jfx = Cabana::slice<FIELD_JFX>(fields); auto jfy = Cabana::slice<FIELD_JFY>(fields); auto jfz = Cabana::slice<FIELD_JFZ>(fields);
    fields.residency().device_valid = device_was_current;
    for (size_t i = 0; i < fields.size(); ++i)
        printf("%d e x %e y %e z %e jfx %e jfy %e jfz %e \n", (int)i, ex(i), ey(i), ez(i), jfx(i), jfy(i), jfz(i));
    std::cout << std::endl;
}
#endif
