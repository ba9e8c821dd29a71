// Facade of the reference's src/uncenter_p.h (:4-8).
#ifndef CABANAPIC_B200_UNCENTER_P_H
#define CABANAPIC_B200_UNCENTER_P_H
#include "types.h"
#include "input/deck.h"

inline void uncenter_particles(particle_list_t particles, interpolator_array_t& f0, real_t qdt_2mc) {
    cabanapic::Runtime& rt = cabanapic::Runtime::get();
    rt.need_on_device(particles);
    rt.need_on_device(f0);
    rt.check(cpic_uncenter_particles(rt.ctx(), qdt_2mc), "cpic_uncenter_particles");
    rt.device_wrote(particles);
}
#endif
