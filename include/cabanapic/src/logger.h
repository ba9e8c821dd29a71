#ifndef CABANAPIC_B200_LOGGER_H
#define CABANAPIC_B200_LOGGER_H
#include <iostream>
// reference src/logger.h:1-6: a stream that is compiled out unless ENABLE_DEBUG is set
#ifdef ENABLE_DEBUG
#define logger std::cout << "LOG:" << __FILE__ << ":" << __LINE__ << " \t :: \t "
#else
#define logger if (0) std::cout
#endif
#endif
