/* Forwarding header: code written for the reference's "ObjTexture.h" keeps compiling.
 * Everything lives in CelioRayTracer.hpp. */
#include "CelioRayTracer.hpp"
