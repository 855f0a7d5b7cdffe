/* Forwarding header: code written for the reference's "Texture_CheckerBoard.h" keeps compiling.
 * Everything lives in CelioRayTracer.hpp. */
#include "CelioRayTracer.hpp"
