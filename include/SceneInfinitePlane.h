/* Forwarding header: code written for the reference's "SceneInfinitePlane.h" keeps compiling.
 * Everything lives in CelioRayTracer.hpp. */
#include "CelioRayTracer.hpp"
