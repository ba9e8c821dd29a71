// Facade of the reference's src/interpolator.h (:14-23): same free functions, executed on the GPU.
#ifndef CABANAPIC_B200_INTERPOLATOR_H
#define CABANAPIC_B200_INTERPOLATOR_H
#include "types.h"
#include "fields.h"

// reference src/interpolator.cpp:4-124 -> k_load_interpolator (cabanapic_b200/csrc/cpic_fields.cuh)
inline void load_interpolator_array(field_array_t fields, interpolator_array_t interpolators, size_t, size_t, size_t, size_t) {
    cabanapic::Runtime& rt = cabanapic::Runtime::get();
    rt.need_on_device(fields);
    rt.need_on_device(interpolators);       // ghost-cell records keep their previous contents
    rt.check(cpic_load_interpolator_array(rt.ctx()), "cpic_load_interpolator_array");
    rt.device_wrote(interpolators);
}

// reference src/interpolator.cpp:125-172
inline void initialize_interpolator(interpolator_array_t& f0) {
    cabanapic::Runtime& rt = cabanapic::Runtime::get();
    rt.check(cpic_initialize_interpolator(rt.ctx()), "cpic_initialize_interpolator");
    rt.device_wrote(f0);
}
#endif
