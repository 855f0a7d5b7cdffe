// host_scene.cpp — implementation of include/CelioRayTracer.hpp (host construction API).
// Compile with -ffp-contract=off: the constructor arithmetic below is part of the parity
// contract with the reference's constructors (cited per function).
#include "CelioRayTracer.hpp"

#include <string.h>

namespace CelioRayTracer {

// Color_Values.h:7-17
const Color COLOR_WHITE(1.f, 1.f, 1.f);
const Color COLOR_RED(1.f, 0.f, 0.f);
const Color COLOR_YELLOW(1.f, 1.f, 0.f);
const Color COLOR_GREEN(0.f, 1.f, 0.f);
const Color COLOR_CYAN(0.f, 1.f, 1.f);
const Color COLOR_BLUE(0.f, 0.f, 1.f);
const Color COLOR_BLACK(0.f, 0.f, 0.f);
const Color COLOR_DARK_GREY(0.33f, 0.33f, 0.33f);
const Color COLOR_LIGHT_GREY(2 / 3.f, 2 / 3.f, 2 / 3.f);
const Color COLOR_BROWN(0.2f, 0.2f, 0.0f);

// ---- textures ---------------------------------------------------------------------------
Texture_CheckerBoard::Texture_CheckerBoard() : ObjTexture(2.0f, 2.0f), light_color(COLOR_WHITE), dark_color(COLOR_BLACK) {}
Texture_CheckerBoard::Texture_CheckerBoard(Color l, Color d) : ObjTexture(2.0f, 2.0f), light_color(l), dark_color(d) {}
bool Texture_CheckerBoard::describeChecker(float o[8]) const {
    o[0] = light_color.r; o[1] = light_color.g; o[2] = light_color.b; o[3] = width;
    o[4] = dark_color.r;  o[5] = dark_color.g;  o[6] = dark_color.b;  o[7] = height;
    return true;
}

// ---- SceneObject (SceneObject.cpp:9-27) ---------------------------------------------------
SceneObject::SceneObject() : origin(), isaLightSource(false), intensity(1.0f), my_object_index(-1) {
    myMaterial.setDiffuseFactor(0.25f);
}
SceneObject::SceneObject(vector3d o) : origin(o), isaLightSource(false), intensity(1.0f), my_object_index(-1) {}

// ---- SceneSphere (SceneSphere.cpp:38-48) --------------------------------------------------
SceneSphere::SceneSphere() : SceneObject(), radius(1.0f), radius_squared(1.0f) {}
SceneSphere::SceneSphere(vector3d _origin, sdecimal32 _radius) : SceneObject(_origin), radius(_radius) {
    radius_squared = _radius * _radius;
}
void SceneSphere::flattenGeometry(SceneFlattener& out, int object_index) const {
    out.addSphere(object_index, origin, radius_squared);
}

// ---- SceneInfinitePlane (SceneInfinitePlane.cpp:11-26) ---------------------------------------
SceneInfinitePlane::SceneInfinitePlane() : SceneObject(), distance_to_origin(0.0f) {}
SceneInfinitePlane::SceneInfinitePlane(vector3d o, vector3d n, vector3d h) : SceneObject(o) {
    normal = n;
    horizontal = h;
    normal.normalize();
    horizontal.normalize();
    vertical.cross(normal, horizontal);
    vertical.normalize();
    reverseNormal = -normal;                 // not re-normalised for this type
    distance_to_origin = -origin.dot(normal);
}
void SceneInfinitePlane::flattenGeometry(SceneFlattener& out, int object_index) const {
    out.addInfinitePlane(object_index, normal, -distance_to_origin, horizontal, vertical, origin, reverseNormal);
}

// ---- SceneFinitePlane (SceneFinitePlane.cpp:13-80) --------------------------------------------
SceneFinitePlane::SceneFinitePlane() : SceneObject(), v_distance(0.0f), h_distance(0.0f), distance_to_origin(0.0f) {}

SceneFinitePlane::SceneFinitePlane(vector3d _origin, vector3d _normal, vector3d _horizontal, float v_dist, float h_dist)
    : SceneObject(_origin) {
    plane_origin = _origin;
    normal = _normal;
    horizontal = _horizontal;
    vertical.cross(normal, horizontal);      // from the un-normalised inputs
    normal.normalize();
    horizontal.normalize();
    vertical.normalize();
    reverseNormal = -normal;
    reverseNormal.normalize();               // re-normalised for this type
    v_distance = v_dist;
    h_distance = h_dist;
    distance_to_origin = -_origin.dot(normal);
}

SceneFinitePlane::SceneFinitePlane(vector3d _origin, vector3d _vertical_corner, vector3d _horizontal_corner)
    : SceneObject(_origin) {
    plane_origin = _origin;
    horizontal = _horizontal_corner - _origin;
    vertical = _vertical_corner - _origin;
    normal.cross(horizontal, vertical);
    v_distance = vertical.length();
    h_distance = horizontal.length();
    vertical.normalize();
    horizontal.normalize();
    normal.normalize();
    reverseNormal = -normal;
    reverseNormal.normalize();
    distance_to_origin = -_origin.dot(normal);
    // The SceneObject origin (= light position if this plane is a light) moves to the far
    // corner: v*vertical + plane_origin + h*horizontal (SceneFinitePlane.cpp:74-79).
    vector3d h = h_distance * horizontal;
    vector3d far_corner = v_distance * vertical + plane_origin + h;
    changeOrigin(far_corner);
}
void SceneFinitePlane::flattenGeometry(SceneFlattener& out, int object_index) const {
    out.addFinitePlane(object_index, normal, -distance_to_origin, horizontal, h_distance, vertical, v_distance,
                       plane_origin, reverseNormal);
}

// ---- Camera (Camera.cpp) ---------------------------------------------------------------------
void Camera::aim(vector3d so, vector3d horiz, vector3d outward) {
    screen_origin = so;
    vector_horizontal = horiz;
    vector_outwards = outward;
    vector_vertical = vector3d();
    vector_vertical.cross(vector_horizontal, vector_outwards);   // before normalising either
    vector_outwards.normalize();
    vector_horizontal.normalize();
    vector_vertical.normalize();
    eye_distance = 1;   // the eye sits one unit behind the screen
    eye_origin = (-eye_distance) * vector_outwards + screen_origin;
}

Camera::Camera() {
    screen_width = 1;
    screen_height = 1;
    screen_halfwidth = screen_width / (sdecimal32)2.0;
    screen_halfheight = screen_height / (sdecimal32)2.0;
    aim(vector3d(-4.f, -4.f, 1.5f), vector3d(.1f, -.08f, 0.f), vector3d(.08f, .1f, .01f));   // Camera.cpp:22-38
}

void Camera::setSceneTwoMirrors() {
    // Camera.cpp:45-51; the literals -.00 there are negative zeros
    aim(vector3d(0.f, 0.f, 2.5f), vector3d(1.f, -0.0f, 0.f), vector3d(0.f, 1.f, -0.0f));
}

Ray* Camera::createEyeRay(sdecimal32 dx_percent, sdecimal32 dy_percent) const {
    sdecimal32 sx = dx_percent * screen_width - screen_halfwidth;
    sdecimal32 sy = dy_percent * screen_height - screen_halfheight;
    vector3d pixel = screen_origin + sx * vector_horizontal;
    pixel = pixel + sy * vector_vertical;
    return new Ray(eye_origin, pixel, eye_origin);
}

void Camera::exportTo(tcrt_camera* c) const {
    c->eye[0] = eye_origin.x; c->eye[1] = eye_origin.y; c->eye[2] = eye_origin.z;
    c->screen_origin[0] = screen_origin.x; c->screen_origin[1] = screen_origin.y; c->screen_origin[2] = screen_origin.z;
    c->horizontal[0] = vector_horizontal.x; c->horizontal[1] = vector_horizontal.y; c->horizontal[2] = vector_horizontal.z;
    c->vertical[0] = vector_vertical.x; c->vertical[1] = vector_vertical.y; c->vertical[2] = vector_vertical.z;
    c->screen_width = screen_width; c->screen_height = screen_height;
    c->screen_halfwidth = screen_halfwidth; c->screen_halfheight = screen_halfheight;
}

// ---- SceneFlattener -----------------------------------------------------------------------------
static void push3w(std::vector<float>& v, vector3d a, float w) {
    v.push_back(a.x); v.push_back(a.y); v.push_back(a.z); v.push_back(w);
}
void SceneFlattener::setInfo(int obj, int type, int slot, vector3d n, vector3d rn) {
    obj_info[4 * obj + 0] = type;
    obj_info[4 * obj + 1] = slot;
    float* q = &obj_normals[8 * obj];
    q[0] = n.x; q[1] = n.y; q[2] = n.z; q[3] = 0.f;
    q[4] = rn.x; q[5] = rn.y; q[6] = rn.z; q[7] = 0.f;
}
int SceneFlattener::addSphere(int obj, vector3d c, float r2) {
    int slot = (int)sphere_obj.size();
    push3w(sphere_geom, c, r2);
    sphere_obj.push_back(obj);
    setInfo(obj, TCRT_SPHERE, slot, vector3d(), vector3d());
    return slot;
}
int SceneFlattener::addFinitePlane(int obj, vector3d n, float neg_dto, vector3d h, float h_dist, vector3d v,
                                   float v_dist, vector3d plane_origin, vector3d reverse_n) {
    int slot = (int)fin_obj.size();
    push3w(fin_geom, n, neg_dto);
    push3w(fin_geom, h, h_dist);
    push3w(fin_geom, v, v_dist);
    push3w(fin_geom, plane_origin, 0.f);
    fin_obj.push_back(obj);
    setInfo(obj, TCRT_FINITE_PLANE, slot, n, reverse_n);
    return slot;
}
int SceneFlattener::addInfinitePlane(int obj, vector3d n, float neg_dto, vector3d h, vector3d v, vector3d o,
                                     vector3d reverse_n) {
    int slot = (int)inf_obj.size();
    push3w(inf_geom, n, neg_dto);
    push3w(inf_geom, h, 0.f);
    push3w(inf_geom, v, 0.f);
    push3w(inf_geom, o, 0.f);
    inf_obj.push_back(obj);
    setInfo(obj, TCRT_INFINITE_PLANE, slot, n, reverse_n);
    return slot;
}
tcrt_scene SceneFlattener::view() const {
    tcrt_scene s;
    memset(&s, 0, sizeof(s));
    s.n_objects = n_objects;
    s.n_spheres = (int)sphere_obj.size();
    s.n_fin_planes = (int)fin_obj.size();
    s.n_inf_planes = (int)inf_obj.size();
    s.n_lights = (int)light_obj.size();
    s.n_textures = (int)(textures.size() / 8);
    s.sphere_geom = sphere_geom.data(); s.sphere_obj = sphere_obj.data();
    s.fin_geom = fin_geom.data();       s.fin_obj = fin_obj.data();
    s.inf_geom = inf_geom.data();       s.inf_obj = inf_obj.data();
    s.obj_surface = obj_surface.data(); s.obj_material = obj_material.data();
    s.obj_origin = obj_origin.data();   s.obj_normals = obj_normals.data();
    s.obj_info = obj_info.data();       s.light_obj = light_obj.data();
    s.textures = textures.data();
    return s;
}

// ---- Scene ------------------------------------------------------------------------------------------
Scene::Scene() : objects(new SceneObject*[MAX_OBJECT_COUNT]), object_count(0) {}
Scene::~Scene() { delete[] objects; }   // like the reference, the objects themselves are not owned

void Scene::addObject(SceneObject* p) {
    if (object_count + 1 >= MAX_OBJECT_COUNT) {
        printf("***ERROR. Added too many objects to scene.\n");   // Scene.cpp:473-474: message, no failure code
        return;
    }
    objects[object_count++] = p;
}

SceneFinitePlane** Scene::makeSceneBox(vector3d o, vector3d d) {
    // Corner k has bit0 -> +dx, bit1 -> +dy, bit2 -> +dz applied to the origin; each
    // coordinate is a single float add (Scene.cpp:397-404).
    vector3d c[8];
    for (int k = 0; k < 8; k++)
        c[k] = vector3d((k & 1) ? o.x + d.x : o.x, (k & 2) ? o.y + d.y : o.y, (k & 4) ? o.z + d.z : o.z);
    // reference naming: c0=000 c1=x c2=y c3=z c4=xy c5=xz c6=yz c7=xyz
    const vector3d &c0 = c[0], &c1 = c[1], &c2 = c[2], &c3 = c[4], &c4 = c[3], &c5 = c[5], &c6 = c[6], &c7 = c[7];
    SceneFinitePlane** planes = new SceneFinitePlane*[6];
    // (origin, vertical corner, horizontal corner), face order of Scene.cpp:408-413
    planes[0] = new SceneFinitePlane(c0, c3, c2);
    planes[1] = new SceneFinitePlane(c0, c3, c1);
    planes[2] = new SceneFinitePlane(c0, c1, c2);
    planes[3] = new SceneFinitePlane(c7, c4, c6);
    planes[4] = new SceneFinitePlane(c7, c4, c5);
    planes[5] = new SceneFinitePlane(c7, c5, c6);
    return planes;
}

namespace {
struct BoxLook {
    Color color;
    float reflective, diffuse, specular;   // negative = leave at the material default
};
void add_box(Scene& s, vector3d o, vector3d d, const BoxLook& look) {
    SceneFinitePlane** faces = s.makeSceneBox(o, d);
    for (int f = 0; f < 6; f++) {
        ObjMaterial* m = faces[f]->getMaterial();
        m->setColor(look.color);
        if (look.reflective >= 0.f) m->setReflectiveFactor(look.reflective);
        if (look.diffuse >= 0.f) m->setDiffuseFactor(look.diffuse);
        if (look.specular >= 0.f) m->setSpecularFactor(look.specular);
        faces[f]->setIndex(s.getObjectCount());
        s.addObject(faces[f]);
    }
    delete[] faces;
}
SceneObject* add_sphere(Scene& s, vector3d c, float r) {
    SceneObject* o = new SceneSphere(c, r);
    o->setIndex(s.getObjectCount());
    s.addObject(o);
    return o;
}
SceneObject* add_checker_ground(Scene& s) {
    SceneObject* g = new SceneInfinitePlane(vector3d(0, 0, 0), vector3d(0, 0, 1), vector3d(1, 0, 0));
    ObjTexture* t = new Texture_CheckerBoard(COLOR_WHITE, COLOR_BLACK);
    t->setHeight(3.0f);
    t->setWidth(3.0f);
    g->getMaterial()->setTexture(t);
    g->setIndex(s.getObjectCount());
    s.addObject(g);
    return g;
}
}  // namespace

// The default "museum" scene: object list of Scene.cpp:209-387.
// Source literals there are doubles narrowed to float at the constructor boundary.
int Scene::initialize() {
    SceneObject* o;
    o = add_sphere(*this, vector3d((float)6.99, (float)6.99, 5.5f), (float).15);   // light 0
    o->setAsLightSource();
    o->setIntensity((float).75);
    o = add_sphere(*this, vector3d(0, 0, (float)4.8), (float).15);                   // light 1
    o->setAsLightSource();
    o->setIntensity(1.0f);

    o = add_sphere(*this, vector3d(0, 0, 2), 1);                                      // red mirror ball
    o->getMaterial()->setColor(COLOR_RED);
    o->getMaterial()->setReflectiveFactor(1.00f);
    add_sphere(*this, vector3d(0, 0, 0), (float)0.01);                                // hidden speck
    o = add_sphere(*this, vector3d((float)(-2.5 + 0 * 2.5), 3, 1), 1);                // red, half specular
    o->getMaterial()->setColor(COLOR_RED);
    o->getMaterial()->setSpecularFactor((float).5);
    o = add_sphere(*this, vector3d((float)(-2.5 + 1 * 2.5), 3, 1), 1);                // pure mirror (diffuse 0)
    o->getMaterial()->setColor(COLOR_WHITE);
    o->getMaterial()->setReflectiveFactor(1.00f);
    o->getMaterial()->setDiffuseFactor(0.00f);
    o = add_sphere(*this, vector3d(), (float).10);                                    // cyan origin ball
    o->getMaterial()->setColor(COLOR_CYAN);

    o = add_checker_ground(*this);
    o->getMaterial()->setColor(COLOR_GREEN);
    o->getMaterial()->setReflectiveFactor((float).5);
    o->getMaterial()->setDiffuseFactor((float).5);

    const BoxLook pedestal = {COLOR_BROWN, 0.00f, 1.00f, -1.f};
    const BoxLook pedestal_base = {COLOR_BROWN, -1.f, -1.f, 0.20f};
    const BoxLook room = {COLOR_DARK_GREY, 0.00f, 1.00f, 0.0f};
    const BoxLook ceiling = {COLOR_LIGHT_GREY, 0.5f, -1.f, 0.5f};
    add_box(*this, vector3d(-.5f, -.5f, 0), vector3d(1.f, 1.f, 1.f), pedestal);
    add_box(*this, vector3d(-.70f, -.70f, 0), vector3d(1.4f, 1.4f, 0.25f), pedestal_base);
    add_box(*this, vector3d(-7, -7, -1), vector3d(14, 14, 7), room);
    add_box(*this, vector3d(-6, -6, 5), vector3d(12, 12, 1), ceiling);
    return 0;
}

namespace {
// One stepped pyramid of spheres of Scene::initializeTwoMirrors (Scene.cpp:133-197): float
// loop counters, bounds shrinking by `offset` per layer, sphere centre (i+dx, j+dy, k).
// dx is added in double and narrowed, as `i+2.65` is in the source; for the integer
// offsets of the other two pyramids the double sum rounds to the same float as a float add.
void add_pyramid(Scene& s, float base_x, float base_y, float base_z, float offset, double dx, double dy, double radius,
                 Color col) {
    float i_start = 0, j_start = 0;
    for (float k = 0; k < base_z; k += offset) {
        i_start += offset;
        j_start += offset;
        for (float i = i_start; i < base_x - i_start; i += offset)
            for (float j = j_start; j < base_y - j_start; j += offset) {
                SceneObject* o = new SceneSphere(vector3d((float)(i + dx), (float)(j + dy), k), (float)radius);
                o->getMaterial()->setColor(col);
                s.addObject(o);
            }
    }
}
SceneObject* add_mirror_panel(Scene& s, vector3d o, vector3d n, vector3d h, float v, float hd) {
    SceneObject* p = new SceneFinitePlane(o, n, h, v, hd);
    s.addObject(p);
    return p;
}
}  // namespace

// "Two mirrors & pyramids": object list of Scene.cpp:23-206 (3920 objects).
int Scene::initializeTwoMirrors(Camera* myCamera) {
    SceneObject* o;
    o = new SceneSphere(vector3d(5, 10, 10), (float).15);
    o->setAsLightSource();
    o->setIntensity(.75f);
    addObject(o);
    o = new SceneSphere(vector3d(), (float).10);
    o->getMaterial()->setColor(COLOR_CYAN);
    addObject(o);
    o = new SceneSphere(vector3d(-40, 100, 40), 10);   // "the sun"
    o->getMaterial()->setColor(COLOR_YELLOW);
    o->getMaterial()->setSpecularFactor((float)0.25);
    addObject(o);
    o = new SceneSphere(vector3d(), (float).05);
    o->getMaterial()->setColor(COLOR_CYAN);
    addObject(o);
    o = new SceneSphere(vector3d(), (float).02);
    o->getMaterial()->setColor(COLOR_CYAN);
    addObject(o);

    o = add_checker_ground(*this);
    o->getMaterial()->setReflectiveFactor((float).05);
    o->getMaterial()->setDiffuseFactor((float).5);

    // mirror 1 and its brown backing, then mirror 2 and its backing
    o = add_mirror_panel(*this, vector3d((float)-1.75, 7, 0), vector3d(0, -1, 0), vector3d(1, 0, 0), 5, 3.5f);
    o->getMaterial()->setColor(COLOR_WHITE);
    o->getMaterial()->setReflectiveFactor(1.0f);
    o->getMaterial()->setDiffuseFactor(.0f);
    o = add_mirror_panel(*this, vector3d(-2, 7, 0), vector3d(0, -1, 0), vector3d(1, 0, 0), (float)5.25, 4.0f);
    o->getMaterial()->setColor(COLOR_BROWN);
    o->getMaterial()->setDiffuseFactor((float).5);
    o = add_mirror_panel(*this, vector3d((float)1.75, -7, 0), vector3d(0, 1, 0), vector3d(-1, 0, 0), 5, 3.5f);
    o->getMaterial()->setColor(COLOR_WHITE);
    o->getMaterial()->setDiffuseFactor(.0f);
    o->getMaterial()->setReflectiveFactor(1.0f);
    o = add_mirror_panel(*this, vector3d(2, -7, 0), vector3d(0, 1, 0), vector3d(-1, 0, 0), (float)5.25, 4.0f);
    o->getMaterial()->setColor(COLOR_BROWN);
    o->getMaterial()->setDiffuseFactor((float).5);

    add_pyramid(*this, 14.50f, 15.0f, 16.5f, 0.5f, 2.65, 15.0, 0.33, COLOR_GREEN);
    add_pyramid(*this, 5.f, 5.0f, 5.f, 0.65f, -6.0, 10.0, .5, COLOR_RED);
    add_pyramid(*this, 1.f, 1.0f, 1.f, 0.33f, 0.0, 20.0, 0.33, COLOR_RED);

    printf("ObjectCount: %d\n", object_count);
    myCamera->setSceneTwoMirrors();
    return 0;
}

void Scene::flatten(SceneFlattener& out) const {
    out = SceneFlattener();
    const int n = object_count;
    out.n_objects = n;
    out.obj_surface.assign(4 * (size_t)n, 0.f);
    out.obj_material.assign(4 * (size_t)n, 0.f);
    out.obj_origin.assign(4 * (size_t)n, 0.f);
    out.obj_normals.assign(8 * (size_t)n, 0.f);
    out.obj_info.assign(4 * (size_t)n, 0);
    std::vector<const ObjTexture*> seen;
    for (int i = 0; i < n; i++) {
        SceneObject* o = objects[i];
        o->flattenGeometry(out, i);
        ObjMaterial* m = o->getMaterial();
        Color c = m->getColor();
        float* s = &out.obj_surface[4 * i];
        s[0] = c.r; s[1] = c.g; s[2] = c.b; s[3] = m->getDiffuseFactor();
        float* q = &out.obj_material[4 * i];
        q[0] = m->getSpecularFactor(); q[1] = m->getReflectiveFactor(); q[2] = o->getIntensity(); q[3] = 0.f;
        vector3d og = o->getOrigin();
        float* g = &out.obj_origin[4 * i];
        g[0] = og.x; g[1] = og.y; g[2] = og.z; g[3] = 0.f;
        out.obj_info[4 * i + 2] = o->checkIsaLightSource() ? 1 : 0;
        if (o->checkIsaLightSource()) out.light_obj.push_back(i);
        int tex_id = -1;
        const ObjTexture* t = m->getTexture();
        float desc[8];
        if (t != NULL && t->describeChecker(desc)) {
            // textures can be edited after being attached, so they are described at flatten
            // time; identical pointers share one slot
            for (size_t k = 0; k < seen.size(); k++)
                if (seen[k] == t) tex_id = (int)k;
            if (tex_id < 0) {
                tex_id = (int)seen.size();
                seen.push_back(t);
                out.textures.insert(out.textures.end(), desc, desc + 8);
            }
        }
        out.obj_info[4 * i + 3] = tex_id;
    }
}

}  // namespace CelioRayTracer
