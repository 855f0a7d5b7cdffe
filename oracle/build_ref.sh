#!/usr/bin/env bash
# oracle/build_ref.sh — TEST INFRASTRUCTURE.  Compiles the reference's own sources (read
# where they lie, /root/reference/src) into oracle/_ref/ (git-ignored).  Nothing from the
# reference is written into this repository: a scratch copy lives in a mktemp dir for the
# duration of the build and only the two binaries come back.
#
#   oracle/_ref/rt_asis      the reference program exactly as checked in (500x504, depth 50,
#                            SCENE 1, its own main / pixel loop / .txt writer).  Writes
#                            raytracer_screen.txt into the cwd.  Built with the three
#                            compile-only shims (1-3 below), no behavioural change.
#   oracle/_ref/rt_fixed     same + shim 4 (hitALightSource_Var initialised).
#   oracle/_ref/ref_render   oracle/ref_harness.cpp around the same sources: run-time W, H,
#                            recursion cap, scene; raw float32 output.  Shims 1-4.
#   oracle/_ref/ref_count    ref_render + two counter increments (rays entering calculatePixel past its
#                            recursion check, RayTracer.cpp:454-457; shadow rays entering inShade, :750).
#                            Never timed: bench.py --impl reference uses it once to count the rays of the
#                            pixel sample that the uninstrumented ref_render is then timed on.
#
# Shims — none changes the arithmetic of the path:
#   1. fixed_class.h / fixed_func.h: an un-vendored third-party fixed-point library
#      (vector3d.h:9-10, rt_project_parameters.h:4, SceneObject.h:10); only live use with
#      USING_FIXED_POINT false is a smoke printf in main (RayTracer.cpp:1129-1132).
#   2. SceneSphere.cpp:108-109,114: diagnostics that reference `.intValue` of a float and an
#      undeclared `root2_float` outside `#if USING_FIXED_POINT` (do not compile; the
#      branches are unreachable for finite inputs).
#   3. strcpy/strcat without <string.h>, string literals as char* (RayTracer.cpp:2075-2106):
#      -include string.h -fpermissive -w.
#   4. CollisionObject::hitALightSource_Var is never initialised (SceneObject.h:47-105,150);
#      set to false in the ctors.  Output-neutral for the stand-alone program (verified by
#      comparing rt_asis with rt_fixed, see tests), mandatory inside any other process.
# For ref_render only: MAX_RECURSION_LEVEL (rt_project_parameters.h:73) becomes the
# run-time global g_ref_max_depth.
set -euo pipefail
REF_ROOT="${TCRT_REFERENCE_DIR:-/root/reference}"
REF="$REF_ROOT/src"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
    if [ -x "$OUT/ref_render" ]; then
        echo "build_ref: $REF not present; keeping prebuilt $OUT" >&2
        exit 0
    fi
    echo "build_ref: $REF not present and no prebuilt oracle/_ref" >&2
    exit 3
fi
TMP="$(mktemp -d /tmp/tcrt_ref.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$OUT"
cp "$REF"/*.h "$REF"/*.cpp "$TMP"/

# shim 1
cat > "$TMP/fixed_class.h" <<'EOF'
#ifndef TCRT_STUB_FIXED_CLASS_H_
#define TCRT_STUB_FIXED_CLASS_H_
struct fixed { int intValue; fixed(double d = 0.0) : intValue((int)(d * 65536.0)) {} };
template <int P> inline float fix2float(int v) { return (float)v / (float)(1 << P); }
#endif
EOF
cat > "$TMP/fixed_func.h" <<'EOF'
#include "fixed_class.h"
EOF
# shim 2
sed -i -e 's/fix2float<16>(root1\.intValue)/root1/' \
       -e 's/fix2float<16>(root2\.intValue), root2_float/root2, root2/' "$TMP/SceneSphere.cpp"

CXXFLAGS="-std=gnu++17 -O2 -ffp-contract=off -fpermissive -w -include string.h"

# (a) the program as checked in
g++ $CXXFLAGS "$TMP"/RayTracer.cpp "$TMP"/Camera.cpp "$TMP"/Scene.cpp "$TMP"/SceneObject.cpp \
    "$TMP"/SceneSphere.cpp "$TMP"/SceneInfinitePlane.cpp "$TMP"/SceneFinitePlane.cpp \
    "$TMP"/PixelQueue.cpp "$TMP"/RChanSupport.cpp -o "$OUT/rt_asis"

# shim 4
sed -i -e 's/normal_ray = Ray(point, normal);/hitALightSource_Var = false; normal_ray = Ray(point, normal);/' \
       -e 's|//distance = 0;|hitALightSource_Var = false;|' "$TMP/SceneObject.h"
grep -c 'hitALightSource_Var = false' "$TMP/SceneObject.h" | grep -qx 2

g++ $CXXFLAGS "$TMP"/RayTracer.cpp "$TMP"/Camera.cpp "$TMP"/Scene.cpp "$TMP"/SceneObject.cpp \
    "$TMP"/SceneSphere.cpp "$TMP"/SceneInfinitePlane.cpp "$TMP"/SceneFinitePlane.cpp \
    "$TMP"/PixelQueue.cpp "$TMP"/RChanSupport.cpp -o "$OUT/rt_fixed"

# (b) harness with a run-time recursion cap
sed -i -e 's/^#define MAX_RECURSION_LEVEL 50/extern int g_ref_max_depth;\n#define MAX_RECURSION_LEVEL g_ref_max_depth/' \
    "$TMP/rt_project_parameters.h"
grep -q 'g_ref_max_depth' "$TMP/rt_project_parameters.h"
g++ $CXXFLAGS -I"$TMP" -I"$HERE/../scenes" "$HERE/ref_harness.cpp" -o "$OUT/ref_render"

# (c) the counting twin (identical pixels; tests compare its counters with the oracle's)
sed -i -e 's|/\* Lots of Magic happens here\.|g_ref_rays[recursion_level == 0 ? 0 : 2]++; /* Lots of Magic happens here.|' \
       -e 's|sdecimal32 dist_to_light = dir.length();|g_ref_rays[1]++; sdecimal32 dist_to_light = dir.length();|' \
    "$TMP/RayTracer.cpp"
grep -c 'g_ref_rays\[' "$TMP/RayTracer.cpp" | grep -qx 2
g++ $CXXFLAGS -DTCRT_REF_COUNT -I"$TMP" -I"$HERE/../scenes" "$HERE/ref_harness.cpp" -o "$OUT/ref_count"

echo "build_ref: built $(ls "$OUT" | tr '\n' ' ')" >&2
