/* Forwarding header: code written for the reference's "SceneFinitePlane.h" keeps compiling.
 * Everything lives in CelioRayTracer.hpp. */
#include "CelioRayTracer.hpp"
