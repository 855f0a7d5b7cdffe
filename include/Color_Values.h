/* Forwarding header: code written for the reference's "Color_Values.h" keeps compiling.
 * Everything lives in CelioRayTracer.hpp. */
#include "CelioRayTracer.hpp"
