// The reference allocates a VPIC-style grid_t (src/grid.h:59-117) and never fills it; push() and
// move_p() ignore the pointer (src/move_p.h:68).  The facade keeps the name so drivers compile;
// the geometry the kernels need travels in the C ABI's cpic_params instead.
#ifndef CABANAPIC_B200_GRID_H
#define CABANAPIC_B200_GRID_H
struct grid_t {
    int nx = 0, ny = 0, nz = 0;
};
#endif
