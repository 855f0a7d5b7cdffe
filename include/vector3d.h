/* Forwarding header: code written for the reference's "vector3d.h" keeps compiling.
 * Everything lives in CelioRayTracer.hpp. */
#include "CelioRayTracer.hpp"
