// The reference's move_p<> (src/move_p.h:59-374) is a device-inline function called from inside
// push<> (src/push.h:271).  Here the cell-crossing mover is part of the push kernel itself
// (mover_streak / cross_face / drain_movers in cabanapic_b200/csrc/cpic_particles.cuh), so this
// header only keeps the include that src/push.h expects.
#ifndef CABANAPIC_B200_MOVE_P_H
#define CABANAPIC_B200_MOVE_P_H
#include "types.h"
#endif
