// Facade of the reference's src/push.h (:7-23): push<_accumulator>(...) with the reference's argument
// list, executed by the sm_100a kernel k_push2 / k_push (Boris push + move_p + current deposit).
#ifndef CABANAPIC_B200_PUSH_H
#define CABANAPIC_B200_PUSH_H
#include "types.h"
#include "move_p.h"
#include "input/deck.h"

template <class _accumulator>
void push(particle_list_t& particles, interpolator_array_t& f0, real_t qdt_2mc, real_t cdt_dx, real_t cdt_dy, real_t cdt_dz,
          real_t qsp, _accumulator& a0, grid_t*, const size_t, const size_t, const size_t, const size_t, Boundary boundary) {
    cabanapic::Runtime& rt = cabanapic::Runtime::get();
    (void)boundary;      // the context was created with deck.BOUNDARY_TYPE (device.h): Periodic, or Reflect = reflecting walls
    rt.need_on_device(particles);
    rt.need_on_device(f0);
    rt.need_on_device(a0.target);
    // the step the reference left commented out (Cabana::sortByKey, example/example.cpp:224-228), opt-in
    if (rt.sort_interval > 0 && rt.pushes % rt.sort_interval == 0) rt.check(cpic_sort_particles(rt.ctx()), "cpic_sort_particles");
    ++rt.pushes;
    cpic_consts k{};
    k.qdt_2mc = qdt_2mc; k.cdt_dx = cdt_dx; k.cdt_dy = cdt_dy; k.cdt_dz = cdt_dz; k.qsp = qsp;
    // large problems: push + cell ordering in one pass (the path bench.py measures); small decks keep the particle order
    if (rt.reorder(particles.size())) rt.check(cpic_push_reorder(rt.ctx(), &k), "cpic_push_reorder");
    else rt.check(cpic_push(rt.ctx(), &k), "cpic_push");
    rt.device_wrote(particles);
    rt.device_wrote(a0.target);
}
#endif
