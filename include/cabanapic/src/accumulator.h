// Facade of the reference's src/accumulator.h (:15-34): same free functions, executed on the GPU.
#ifndef CABANAPIC_B200_ACCUMULATOR_H
#define CABANAPIC_B200_ACCUMULATOR_H
#include "types.h"
#include "grid.h"
#include "fields.h"

// reference src/accumulator.cpp:5-43
inline void clear_accumulator_array(field_array_t&, accumulator_array_t& accumulators, size_t, size_t, size_t) {
    cabanapic::Runtime& rt = cabanapic::Runtime::get();
    rt.check(cpic_clear_accumulator_array(rt.ctx()), "cpic_clear_accumulator_array");
    rt.device_wrote(accumulators);
}

// reference src/accumulator.cpp:45-122 -> k_unload_accumulator
inline void unload_accumulator_array(field_array_t& fields, accumulator_array_t& accumulators, size_t, size_t, size_t, size_t,
                                     real_t dx, real_t dy, real_t dz, real_t dt) {
    cabanapic::Runtime& rt = cabanapic::Runtime::get();
    rt.need_on_device(fields);
    rt.need_on_device(accumulators);
    cpic_consts k{};
    k.dx = dx; k.dy = dy; k.dz = dz; k.dt = dt;
    rt.check(cpic_unload_accumulator_array(rt.ctx(), &k), "cpic_unload_accumulator_array");
    rt.device_wrote(fields);
}
#endif
