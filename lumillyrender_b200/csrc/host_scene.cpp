// host_scene.cpp — scene assembly on the host: TOML -> flat arrays -> LrSceneDesc.
// C++ restatement of scene_loader.rs (schema, transform composition, light binding) and
// description.rs (instances in TOML order, world-space triangles, camera selection).  Runs once
// per scene; nothing here is on the per-sample path.  fp32 throughout, -ffp-contract=off.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "host_scene.h"

namespace lr {

// ============================================================================ matrices (matrix4.rs)
Mat4 mat4_unit() { return Mat4{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}}; }                      // :9-18
Mat4 mat4_translate(Vec3 v) { return Mat4{{1, 0, 0, v[0], 0, 1, 0, v[1], 0, 0, 1, v[2], 0, 0, 0, 1}}; }  // :20-29
Mat4 mat4_scale(Vec3 v) { return Mat4{{v[0], 0, 0, 0, 0, v[1], 0, 0, 0, 0, v[2], 0, 0, 0, 0, 1}}; }      // :31-40
Mat4 mat4_axis_angle(Vec3 a, float t) {                                                                  // :42-54 (Rodrigues)
  const float c = std::cos(t), s = std::sin(t), k = 1.0f - c;
  Mat4 r;
  r.m[0] = c + a[0] * a[0] * k;        r.m[1] = a[0] * a[1] * k - a[2] * s; r.m[2] = a[0] * a[2] * k + a[1] * s;  r.m[3] = 0.0f;
  r.m[4] = a[1] * a[0] * k + a[2] * s; r.m[5] = c + a[1] * a[1] * k;        r.m[6] = a[1] * a[2] * k - a[0] * s;  r.m[7] = 0.0f;
  r.m[8] = a[2] * a[0] * k - a[1] * s; r.m[9] = a[2] * a[1] * k + a[0] * s; r.m[10] = c + a[2] * a[2] * k;        r.m[11] = 0.0f;
  r.m[12] = 0.0f; r.m[13] = 0.0f; r.m[14] = 0.0f; r.m[15] = 1.0f;
  return r;
}
// :56-68 — basis vectors are stored as ROWS and the origin in row 3 (quirk Q9)
Mat4 mat4_look_at(Vec3 origin, Vec3 target, Vec3 up) {
  const Vec3 za = vnormalize(vsub(origin, target));
  const Vec3 xa = vnormalize(vcross(up, za));
  const Vec3 ya = vcross(za, xa);
  return Mat4{{xa[0], xa[1], xa[2], 0.0f, ya[0], ya[1], ya[2], 0.0f, za[0], za[1], za[2], 0.0f, origin[0], origin[1], origin[2], 1.0f}};
}
static inline float dot4(const float* a, float b0, float b1, float b2, float b3) { return a[0] * b0 + a[1] * b1 + a[2] * b2 + a[3] * b3; }
Mat4 mat4_mul(const Mat4& a, const Mat4& b) {                                                            // :201-211
  Mat4 r;
  for (int y = 0; y < 4; y++)
    for (int x = 0; x < 4; x++) r.m[4 * y + x] = dot4(&a.m[4 * y], b.m[x], b.m[x + 4], b.m[x + 8], b.m[x + 12]);
  return r;
}
Vec3 mat4_apply(const Mat4& m, Vec3 p) {                                                                 // :185-199, w = 1 (vector4.rs:40-44)
  return vec3(dot4(&m.m[0], p[0], p[1], p[2], 1.0f), dot4(&m.m[4], p[0], p[1], p[2], 1.0f), dot4(&m.m[8], p[0], p[1], p[2], 1.0f));
}

// ============================================================================ cameras (camera.rs constructors)
static void put3(float* dst, Vec3 v) { dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2]; }

void camera_ideal_pinhole(const Mat4& m, float xfov, int w, int h, LrCamera& c) {                        // camera.rs:34-62
  std::memset(&c, 0, sizeof(c));
  const Vec3 aperture = vec3(m.m[12], m.m[13], m.m[14]);                 // matrix.row(3)
  const Vec3 forward = mat4_apply(m, vec3(0.0f, 0.0f, -1.0f));
  const Vec3 right = mat4_apply(m, vec3(1.0f, 0.0f, 0.0f));
  const Vec3 up = mat4_apply(m, vec3(0.0f, 1.0f, 0.0f));
  const Vec3 direction = vscale(forward, 50.0f);
  const Vec3 position = vsub(aperture, direction);
  const float dist = vnorm(direction);
  const float sx = 2.0f * dist * std::tan(xfov * kHostPI / 180.0f / 2.0f);
  const float sy = sx * (float)h / (float)w;
  c.type = LR_CAM_IDEAL_PINHOLE; c.width = w; c.height = h;
  put3(c.forward, forward); put3(c.right, right); put3(c.up, up);
  put3(c.position, position); put3(c.aperture_position, aperture);
  c.sensor_size[0] = sx; c.sensor_size[1] = sy;
  c.aperture_sensor_distance = dist;
  c.sensor_sensitivity = 1.0f;                                           // camera.rs:117-119
}
void camera_thin_lens(const Mat4& m, float xfov, float focus_distance, float f_number, int w, int h, LrCamera& c) {   // camera.rs:366-409
  camera_ideal_pinhole(m, xfov, w, h, c);
  c.type = LR_CAM_THIN_LENS;
  const float dist = c.aperture_sensor_distance;
  const float focal_length = 1.0f / (1.0f / dist + 1.0f / focus_distance);
  c.aperture_radius = focal_length / f_number / 2.0f;
  c.sensor_pixel_area = (c.sensor_size[0] * c.sensor_size[1]) / (float)((size_t)w * (size_t)h);
  c.sensor_sensitivity = dist * dist / (c.sensor_pixel_area * kHostPI * c.aperture_radius * c.aperture_radius);
  c.focus_distance = focus_distance;
}
void camera_omnidirectional(const Mat4& m, int w, int h, LrCamera& c) {                                  // camera.rs:149-166
  std::memset(&c, 0, sizeof(c));
  c.type = LR_CAM_OMNIDIRECTIONAL; c.width = w; c.height = h;
  put3(c.forward, mat4_apply(m, vec3(0.0f, 0.0f, -1.0f)));
  put3(c.right, mat4_apply(m, vec3(1.0f, 0.0f, 0.0f)));
  put3(c.up, mat4_apply(m, vec3(0.0f, 1.0f, 0.0f)));
  put3(c.aperture_position, vec3(m.m[12], m.m[13], m.m[14]));
  c.sensor_sensitivity = 1.0f;
}
void camera_pinhole(Vec3 position, Vec3 aperture, const float* sensor_size, int w, int h, float aperture_radius, LrCamera& c) {   // camera.rs:224-264
  std::memset(&c, 0, sizeof(c));
  const Vec3 direction = vsub(aperture, position);
  const float dist = vnorm(direction);
  const Vec3 forward = vnormalize(direction);
  const Vec3 helper = std::fabs(forward[1]) < 1.0f - 1e-3f ? vec3(0.0f, 1.0f, 0.0f) : vec3(1.0f, 0.0f, 0.0f);
  const Vec3 right = vnormalize(vcross(forward, helper));
  const Vec3 up = vcross(right, forward);
  c.type = LR_CAM_PINHOLE; c.width = w; c.height = h;
  put3(c.forward, forward); put3(c.right, right); put3(c.up, up);
  put3(c.position, position); put3(c.aperture_position, aperture);
  c.sensor_size[0] = sensor_size[0]; c.sensor_size[1] = sensor_size[1];
  c.aperture_radius = aperture_radius;
  c.aperture_sensor_distance = dist;
  c.sensor_pixel_area = (sensor_size[0] * sensor_size[1]) / (float)((size_t)w * (size_t)h);
  c.sensor_sensitivity = dist * dist / (c.sensor_pixel_area * kHostPI * aperture_radius * aperture_radius);
}

// ============================================================================ TOML -> scene
namespace {

struct LoadError { int code; std::string msg; };

const TomlValue& need(const TomlValue& t, const std::string& key, const std::string& where) {
  const TomlValue* v = t.get(key);
  if (!v) throw LoadError{LR_ERR_PARSE, "missing field `" + key + "` in " + where};
  return *v;
}
float as_f32(const TomlValue& v, const std::string& what) {
  if (!v.is_number()) throw LoadError{LR_ERR_PARSE, "`" + what + "` must be a number"};
  return (float)v.number();
}
long long as_int(const TomlValue& v, const std::string& what) {
  if (v.kind != TomlValue::INT || v.i < 0) throw LoadError{LR_ERR_PARSE, "`" + what + "` must be a non-negative integer"};
  return v.i;
}
Vec3 as_vec3(const TomlValue& v, const std::string& what) {
  if (v.kind != TomlValue::ARRAY || v.arr.size() != 3) throw LoadError{LR_ERR_PARSE, "`" + what + "` must be an array of 3 numbers"};
  return vec3(as_f32(v.arr[0], what), as_f32(v.arr[1], what), as_f32(v.arr[2], what));
}
const std::string& as_str(const TomlValue& v, const std::string& what) {
  if (v.kind != TomlValue::STRING) throw LoadError{LR_ERR_PARSE, "`" + what + "` must be a string"};
  return v.s;
}
// accepts kebab-case and, for the two keys welcome-2018.toml spells with underscores, snake_case
const TomlValue* get_either(const TomlValue& t, const std::string& kebab) {
  if (const TomlValue* v = t.get(kebab)) return v;
  std::string snake = kebab;
  for (char& c : snake) if (c == '-') c = '_';
  return t.get(snake);
}

// scene_loader.rs:88-104 — M = T_n ... T_2 T_1 (fold(unit, |p, c| c * p))
Mat4 transform_matrix(const TomlValue* list, const std::string& where) {
  Mat4 m = mat4_unit();
  if (!list) return m;
  if (list->kind != TomlValue::ARRAY) throw LoadError{LR_ERR_PARSE, "`transform` of " + where + " must be an array of tables"};
  for (const TomlValue& t : list->arr) {
    const std::string& type = as_str(need(t, "type", "transform"), "type");
    Mat4 c;
    if (type == "translate") c = mat4_translate(as_vec3(need(t, "vector", type), "vector"));
    else if (type == "scale") c = mat4_scale(as_vec3(need(t, "vector", type), "vector"));
    else if (type == "axis-angle") c = mat4_axis_angle(as_vec3(need(t, "axis", type), "axis"), as_f32(need(t, "angle", type), "angle") * kHostPI / 180.0f);
    else if (type == "look-at") c = mat4_look_at(as_vec3(need(t, "origin", type), "origin"), as_vec3(need(t, "target", type), "target"), as_vec3(need(t, "up", type), "up"));
    else throw LoadError{LR_ERR_PARSE, "unknown transform type `" + type + "`"};
    m = mat4_mul(c, m);
  }
  return m;
}

LrMaterial make_material(const TomlValue& m, Vec3 emission) {        // description.rs:95-130
  LrMaterial out;
  std::memset(&out, 0, sizeof(out));
  const std::string& type = as_str(need(m, "type", "material"), "type");
  if (type == "lambert") {
    out.type = LR_MAT_LAMBERT;
    const Vec3 a = as_vec3(need(m, "albedo", type), "albedo");
    out.color[0] = a[0]; out.color[1] = a[1]; out.color[2] = a[2];
    out.emission[0] = emission[0]; out.emission[1] = emission[1]; out.emission[2] = emission[2];
    return out;
  }
  const Vec3 r = as_vec3(need(m, "reflectance", type), "reflectance");
  out.color[0] = r[0]; out.color[1] = r[1]; out.color[2] = r[2];
  if (type == "phong") { out.type = LR_MAT_PHONG; out.param0 = as_f32(need(m, "alpha", type), "alpha"); }
  else if (type == "blinn-phong") { out.type = LR_MAT_BLINN_PHONG; out.param0 = as_f32(need(m, "alpha", type), "alpha"); }
  else if (type == "ggx") { out.type = LR_MAT_GGX; out.param0 = as_f32(need(m, "roughness", type), "roughness"); out.param1 = as_f32(need(m, "ior", type), "ior"); }
  else if (type == "ideal-refraction") {
    out.type = LR_MAT_IDEAL_REFRACTION;
    const TomlValue* ab = m.get("absorbtance");
    out.param0 = ab ? as_f32(*ab, "absorbtance") : 0.0f;
    out.param1 = as_f32(need(m, "ior", type), "ior");
  } else throw LoadError{LR_ERR_PARSE, "unknown material type `" + type + "`"};
  return out;
}

std::string join_path(const std::string& root, const std::string& rel) {
  if (root.empty() || (!rel.empty() && rel[0] == '/')) return rel;
  return root.back() == '/' ? root + rel : root + "/" + rel;
}

const TomlValue* find_named(const TomlValue* list, const std::string& name) {
  if (!list || list->kind != TomlValue::ARRAY) return nullptr;
  for (const TomlValue& t : list->arr) {
    const TomlValue* n = t.get("name");
    if (n && n->kind == TomlValue::STRING && n->s == name) return &t;
  }
  return nullptr;
}

void load_into(LrHostScene& hs, const std::string& toml_path, const std::string& asset_root, int ow, int oh) {
  std::ifstream f(toml_path, std::ios::binary);
  if (!f) throw LoadError{LR_ERR_IO, "File `" + toml_path + "` is not found."};            // description.rs:34
  std::ostringstream ss;
  ss << f.rdbuf();
  TomlValue root;
  std::string err;
  if (toml_parse(ss.str(), root, err) != LR_OK) throw LoadError{LR_ERR_PARSE, err};

  // ---- [renderer] scene_loader.rs:10-17, defaults description.rs:75-79 / main.rs:62-66
  LrSceneConfig& cfg = hs.config;
  std::memset(&cfg, 0, sizeof(cfg));
  const TomlValue& renderer = need(root, "renderer", "scene");
  cfg.samples = (int)as_int(need(renderer, "samples", "[renderer]"), "samples");
  cfg.depth = 5; cfg.depth_limit = 64; cfg.no_direct_emitter = 0; cfg.threads = 0; cfg.integrator = LR_INTEGRATOR_PT_DIRECT;
  if (const TomlValue* v = renderer.get("depth")) cfg.depth = (int)as_int(*v, "depth");
  if (const TomlValue* v = renderer.get("depth-limit")) cfg.depth_limit = (int)as_int(*v, "depth-limit");
  if (const TomlValue* v = renderer.get("no-direct-emitter")) { if (v->kind != TomlValue::BOOL) throw LoadError{LR_ERR_PARSE, "`no-direct-emitter` must be a boolean"}; cfg.no_direct_emitter = v->b; }
  if (const TomlValue* v = renderer.get("threads")) cfg.threads = (int)as_int(*v, "threads");
  if (const TomlValue* v = renderer.get("integrator")) {
    const std::string& s = as_str(*v, "integrator");
    if (s == "pt") cfg.integrator = LR_INTEGRATOR_PT;
    else if (s == "pt-direct") cfg.integrator = LR_INTEGRATOR_PT_DIRECT;
    else throw LoadError{LR_ERR_INVALID, "Unknown integrator type `" + s + "`"};           // main.rs:124
  }
  // ---- [film] scene_loader.rs:21-26
  const TomlValue& film = need(root, "film", "scene");
  const TomlValue& res = need(film, "resolution", "[film]");
  if (res.kind != TomlValue::ARRAY || res.arr.size() != 2) throw LoadError{LR_ERR_PARSE, "`resolution` must be [width, height]"};
  cfg.width = (int)as_int(res.arr[0], "resolution"); cfg.height = (int)as_int(res.arr[1], "resolution");
  if (ow > 0 && oh > 0) { cfg.width = ow; cfg.height = oh; }
  if (cfg.width <= 0 || cfg.height <= 0) throw LoadError{LR_ERR_INVALID, "resolution must be positive"};
  const std::string& output = as_str(need(film, "output", "[film]"), "output");
  if (output == "png") cfg.output = 0; else if (output == "hdr") cfg.output = 1;
  else throw LoadError{LR_ERR_INVALID, "Unsupported output type `" + output + "`"};        // main.rs:166
  cfg.gamma = 2.2f;                                                                        // main.rs:136
  if (const TomlValue* v = film.get("gamma")) cfg.gamma = as_f32(*v, "gamma");
  // film.sensitivity is parsed and ignored by the reference (scene_loader.rs:25)

  // ---- [sky] scene_loader.rs:30-40, description.rs:58-65
  LrSky sky;
  std::memset(&sky, 0, sizeof(sky));
  sky.type = LR_SKY_UNIFORM;
  if (const TomlValue* s = root.get("sky")) {
    const std::string& type = as_str(need(*s, "type", "[sky]"), "type");
    if (type == "uniform") {
      const Vec3 c = as_vec3(need(*s, "color", "[sky]"), "color");
      sky.color[0] = c[0]; sky.color[1] = c[1]; sky.color[2] = c[2];
    } else if (type == "ibl") {
      const std::string path = join_path(asset_root, as_str(need(*s, "path", "[sky]"), "path"));
      int w = 0, h = 0;
      if (int rc = load_hdr_file(path, hs.sky_pixels, w, h)) throw LoadError{rc, lr_last_error()};
      if ((long long)w * h < 2LL * h * h) throw LoadError{LR_ERR_UNSUPPORTED, "IBL image narrower than 2*height: the reference indexes it as 2H x H (sky.rs:64-72)"};
      sky.type = LR_SKY_IBL; sky.height = h; sky.n_pixels = (int64_t)w * h;
      if (const TomlValue* v = s->get("longitude-offset")) sky.longitude_offset = as_f32(*v, "longitude-offset");
    } else throw LoadError{LR_ERR_PARSE, "unknown sky type `" + type + "`"};
  }

  // ---- [camera] scene_loader.rs:108-125, description.rs:46-55
  const TomlValue& cam = need(root, "camera", "scene");
  const std::string& cam_type = as_str(need(cam, "type", "[camera]"), "type");
  LrCamera camera;
  if (cam_type == "pinhole") {
    // additive extension: the realistic pinhole of camera.rs:200-337 is unreachable from the reference's TOML
    float ss[2];
    const TomlValue& sz = need(cam, "sensor-size", "[camera]");
    if (sz.kind != TomlValue::ARRAY || sz.arr.size() != 2) throw LoadError{LR_ERR_PARSE, "`sensor-size` must be [x, y]"};
    ss[0] = as_f32(sz.arr[0], "sensor-size"); ss[1] = as_f32(sz.arr[1], "sensor-size");
    camera_pinhole(as_vec3(need(cam, "position", "[camera]"), "position"), as_vec3(need(cam, "aperture-position", "[camera]"), "aperture-position"),
                   ss, cfg.width, cfg.height, as_f32(need(cam, "aperture-radius", "[camera]"), "aperture-radius"), camera);
  } else {
    const Mat4 m = transform_matrix(cam.get("transform"), "[camera]");
    if (cam_type == "ideal-pinhole") camera_ideal_pinhole(m, as_f32(need(cam, "fov", "[camera]"), "fov"), cfg.width, cfg.height, camera);
    else if (cam_type == "thin-lens") {
      const TomlValue* fd = get_either(cam, "focus-distance");
      const TomlValue* fn = get_either(cam, "f-number");
      if (!fd) throw LoadError{LR_ERR_PARSE, "missing field `focus-distance` in [camera]"};
      if (!fn) throw LoadError{LR_ERR_PARSE, "missing field `f-number` in [camera]"};
      camera_thin_lens(m, as_f32(need(cam, "fov", "[camera]"), "fov"), as_f32(*fd, "focus-distance"), as_f32(*fn, "f-number"), cfg.width, cfg.height, camera);
    } else if (cam_type == "omnidirectional") camera_omnidirectional(m, cfg.width, cfg.height, camera);
    else throw LoadError{LR_ERR_PARSE, "unknown camera type `" + cam_type + "`"};
  }

  // ---- objects: scene_loader.rs:248-270 + description.rs:89-148
  const TomlValue* objects = root.get("object");
  const TomlValue* meshes = root.get("mesh");
  const TomlValue* materials = root.get("material");
  const TomlValue* lights = root.get("light");
  std::map<std::string, ObjFile> obj_cache;
  int prim_id = 0;
  if (objects && objects->kind == TomlValue::ARRAY) {
    for (const TomlValue& o : objects->arr) {
      const std::string& mesh_name = as_str(need(o, "mesh", "[[object]]"), "mesh");
      const TomlValue* mesh = find_named(meshes, mesh_name);
      if (!mesh) throw LoadError{LR_ERR_INVALID, "Mesh named `" + mesh_name + "` is not found."};          // scene_loader.rs:238
      const TomlValue* mat = nullptr;
      if (const TomlValue* mn = o.get("material")) {
        mat = find_named(materials, as_str(*mn, "material"));
        if (!mat) throw LoadError{LR_ERR_INVALID, "Material named `" + mn->s + "` is not found."};        // scene_loader.rs:243
      }
      // first [[light]] whose `object` equals this object's name (scene_loader.rs:254-262)
      Vec3 emission = vec3(0.0f, 0.0f, 0.0f);
      const TomlValue* oname = o.get("name");
      if (lights && lights->kind == TomlValue::ARRAY && oname && oname->kind == TomlValue::STRING) {
        for (const TomlValue& l : lights->arr) {
          const std::string& lt = as_str(need(l, "type", "[[light]]"), "type");
          if (lt != "area") throw LoadError{LR_ERR_PARSE, "unknown light type `" + lt + "`"};
          if (as_str(need(l, "object", "[[light]]"), "object") != oname->s) continue;
          const Vec3 e = as_vec3(need(l, "emission", "[[light]]"), "emission");
          const TomlValue* in = l.get("intensity");
          emission = vscale(e, in ? as_f32(*in, "intensity") : 1.0f);
          break;
        }
      }
      const Mat4 transform = transform_matrix(o.get("transform"), "[[object]]");
      const std::string& mesh_type = as_str(need(*mesh, "type", "[[mesh]]"), "type");
      int default_material = -1;
      if (mat) { hs.materials.push_back(make_material(*mat, emission)); default_material = (int)hs.materials.size() - 1; }
      if (mesh_type == "sphere") {                                                                         // description.rs:137-142
        if (default_material < 0) throw LoadError{LR_ERR_INVALID, "Material must be specified for object `" + mesh_name + "`"};
        const Vec3 c = mat4_apply(transform, vec3(0.0f, 0.0f, 0.0f));
        LrSphere s;
        s.center[0] = c[0]; s.center[1] = c[1]; s.center[2] = c[2];
        s.radius = as_f32(need(*mesh, "radius", "[[mesh]]"), "radius");       // radius is NOT scaled by the transform (quirk Q11)
        s.material = default_material; s.prim_id = prim_id++;
        hs.spheres.push_back(s);
      } else if (mesh_type == "obj") {                                                                     // description.rs:164-197
        const std::string path = join_path(asset_root, as_str(need(*mesh, "path", "[[mesh]]"), "path"));
        auto it = obj_cache.find(path);
        if (it == obj_cache.end()) {
          ObjFile of;
          if (int rc = load_obj(path, of)) throw LoadError{rc, lr_last_error()};
          it = obj_cache.emplace(path, std::move(of)).first;
        }
        const ObjFile& of = it->second;
        std::vector<int> mtl_material(of.materials.size(), -1);
        for (const ObjModel& m : of.models) {
          int material = default_material;
          if (material < 0) {
            if (m.material_id < 0) throw LoadError{LR_ERR_INVALID, "Specified material is not found in mlt file."};   // description.rs:178
            if (mtl_material[m.material_id] < 0) {
              LrMaterial lm;
              std::memset(&lm, 0, sizeof(lm));
              lm.type = LR_MAT_LAMBERT;
              for (int k = 0; k < 3; k++) { lm.color[k] = of.materials[m.material_id].diffuse[k]; lm.emission[k] = emission[k]; }
              hs.materials.push_back(lm);
              mtl_material[m.material_id] = (int)hs.materials.size() - 1;
            }
            material = mtl_material[m.material_id];
          }
          const size_t faces = m.positions.size() / 9;
          for (size_t fi = 0; fi < faces; fi++) {
            LrTriangle t;
            const float* p = &m.positions[9 * fi];
            const Vec3 a = mat4_apply(transform, vec3(p)), b = mat4_apply(transform, vec3(p + 3)), c = mat4_apply(transform, vec3(p + 6));
            for (int k = 0; k < 3; k++) { t.p0[k] = a[k]; t.p1[k] = b[k]; t.p2[k] = c[k]; }
            t.material = material; t.prim_id = prim_id++;
            hs.triangles.push_back(t);
          }
        }
      } else throw LoadError{LR_ERR_PARSE, "unknown mesh type `" + mesh_type + "`"};
    }
  }
  hs.desc.camera = camera;
  hs.desc.sky = sky;
}

}  // namespace
}  // namespace lr

using namespace lr;

int LrHostScene::finalize() {
  float seconds = 0.0f;
  int depth = 0;
  int n_flat = 0;
  // where rays start besides the triangles: the camera and the spheres (the pad of the node boxes scales with it, bvh_build.cpp)
  float tri_extent = 0.0f, origin_extent = 0.0f;
  for (const LrTriangle& t : triangles)
    for (int k = 0; k < 3; k++) tri_extent = std::fmax(tri_extent, std::fmax(std::fabs(t.p0[k]), std::fmax(std::fabs(t.p1[k]), std::fabs(t.p2[k]))));
  for (int k = 0; k < 3; k++)
    origin_extent = std::fmax(origin_extent, std::fmax(std::fabs(desc.camera.position[k]), std::fabs(desc.camera.aperture_position[k])));
  for (const LrSphere& sp : spheres)
    for (int k = 0; k < 3; k++) origin_extent = std::fmax(origin_extent, std::fabs(sp.center[k]) + std::fmin(std::fabs(sp.radius), 16.0f * tri_extent));
  if (int rc = build_bvh(triangles, nodes, depth, seconds, n_flat, origin_extent, bvh_builder, &bvh_builder_used, &bvh_device_kernel_ms)) return rc;
  desc.n_flat_triangles = n_flat;
  desc.materials = materials.data(); desc.n_materials = (int)materials.size();
  desc.triangles = triangles.data(); desc.n_triangles = (int)triangles.size();
  desc.spheres = spheres.data(); desc.n_spheres = (int)spheres.size();
  desc.nodes = nodes.data(); desc.n_nodes = (int)nodes.size();
  desc.bvh_depth = depth;
  if (desc.sky.type == LR_SKY_IBL) desc.sky.pixels = sky_pixels.data();
  config.n_prims = desc.n_triangles + desc.n_spheres;
  config.bvh_build_seconds = seconds;
  config.bvh_builder = bvh_builder_used;
  config.bvh_device_kernel_ms = bvh_device_kernel_ms;
  int n_em = 0;
  auto emissive = [&](int m) {
    const LrMaterial& mm = materials[m];
    return mm.type == LR_MAT_LAMBERT && (mm.emission[0] * mm.emission[0] + mm.emission[1] * mm.emission[1] + mm.emission[2] * mm.emission[2]) > 0.0f;
  };
  for (const LrTriangle& t : triangles) n_em += emissive(t.material);
  for (const LrSphere& s : spheres) n_em += emissive(s.material);
  config.n_emitters = n_em;
  return validate_desc(desc);
}

extern "C" {

int lr_host_scene_load(const char* toml_path, const char* asset_root, int32_t ow, int32_t oh, LrHostScene** out) {
  if (!toml_path || !out) return fail(LR_ERR_INVALID, "null argument");
  *out = nullptr;
  auto hs = std::make_unique<LrHostScene>();
  if (const char* b = std::getenv("LR_BVH_BUILDER")) hs->bvh_builder = std::strcmp(b, "device") == 0 ? LR_BVH_DEVICE : LR_BVH_HOST;
  try {
    load_into(*hs, toml_path, asset_root ? asset_root : "", ow, oh);
  } catch (const LoadError& e) {
    return fail(e.code, e.msg);
  } catch (const std::exception& e) {
    return fail(LR_ERR_INVALID, e.what());
  }
  if (int rc = hs->finalize()) return rc;
  *out = hs.release();
  return LR_OK;
}

static int lr_host_scene_rebuild_bvh_body(LrHostScene* hs, int32_t builder) {
  if (!hs) return fail(LR_ERR_INVALID, "null argument");
  if (builder != LR_BVH_HOST && builder != LR_BVH_DEVICE) return fail(LR_ERR_INVALID, "unknown BVH builder");
  const int prev = hs->bvh_builder;
  hs->bvh_builder = builder;
  const int rc = hs->finalize();
  if (rc == LR_OK) return rc;
  // the scene keeps a valid tree: rebuild with the builder it had (the failed attempt cleared the node array)
  const std::string msg = lr_last_error();
  hs->bvh_builder = prev;
  if (hs->finalize() != LR_OK) return rc;
  return fail(rc, msg);
}
int lr_host_scene_rebuild_bvh(LrHostScene* hs, int32_t builder) { LR_GUARDED(lr_host_scene_rebuild_bvh_body(hs, builder)); }

static int lr_host_scene_from_arrays_body(const LrMaterial* materials, int32_t n_materials, const LrTriangle* triangles, int32_t n_triangles,
                              const LrSphere* spheres, int32_t n_spheres, const LrCamera* camera, const LrSky* sky, LrHostScene** out) {
  if (!out || !camera || n_materials < 0 || n_triangles < 0 || n_spheres < 0) return fail(LR_ERR_INVALID, "bad argument");
  if ((n_materials && !materials) || (n_triangles && !triangles) || (n_spheres && !spheres)) return fail(LR_ERR_INVALID, "null array with non-zero count");
  *out = nullptr;
  auto hs = std::make_unique<LrHostScene>();
  hs->materials.assign(materials, materials + n_materials);
  hs->triangles.assign(triangles, triangles + n_triangles);
  hs->spheres.assign(spheres, spheres + n_spheres);
  hs->desc.camera = *camera;
  std::memset(&hs->desc.sky, 0, sizeof(LrSky));
  if (sky) {
    hs->desc.sky = *sky;
    if (sky->type == LR_SKY_IBL) {
      if (!sky->pixels || sky->n_pixels <= 0) return fail(LR_ERR_INVALID, "IBL sky without pixels");
      hs->sky_pixels.assign(sky->pixels, sky->pixels + 3 * sky->n_pixels);
    }
  }
  std::memset(&hs->config, 0, sizeof(LrSceneConfig));
  hs->config.width = camera->width; hs->config.height = camera->height;
  hs->config.depth = 5; hs->config.depth_limit = 64; hs->config.gamma = 2.2f; hs->config.integrator = LR_INTEGRATOR_PT_DIRECT;
  if (int rc = hs->finalize()) return rc;
  *out = hs.release();
  return LR_OK;
}
int lr_host_scene_from_arrays(const LrMaterial* materials, int32_t n_materials, const LrTriangle* triangles, int32_t n_triangles,
                              const LrSphere* spheres, int32_t n_spheres, const LrCamera* camera, const LrSky* sky, LrHostScene** out) {
  LR_GUARDED(lr_host_scene_from_arrays_body(materials, n_materials, triangles, n_triangles, spheres, n_spheres, camera, sky, out));
}

const LrSceneDesc* lr_host_scene_desc(const LrHostScene* hs) { return hs ? &hs->desc : nullptr; }
int lr_host_scene_config(const LrHostScene* hs, LrSceneConfig* cfg) {
  if (!hs || !cfg) return fail(LR_ERR_INVALID, "null argument");
  *cfg = hs->config;
  return LR_OK;
}
void lr_host_scene_free(LrHostScene* hs) { delete hs; }

static Mat4 m_from(const float* m) { Mat4 r; std::memcpy(r.m, m, sizeof(r.m)); return r; }
int lr_camera_ideal_pinhole(const float* matrix, float xfov, int32_t w, int32_t h, LrCamera* out) {
  if (!matrix || !out || w <= 0 || h <= 0) return fail(LR_ERR_INVALID, "bad argument");
  camera_ideal_pinhole(m_from(matrix), xfov, w, h, *out);
  return LR_OK;
}
int lr_camera_thin_lens(const float* matrix, float xfov, float focus_distance, float f_number, int32_t w, int32_t h, LrCamera* out) {
  if (!matrix || !out || w <= 0 || h <= 0) return fail(LR_ERR_INVALID, "bad argument");
  camera_thin_lens(m_from(matrix), xfov, focus_distance, f_number, w, h, *out);
  return LR_OK;
}
int lr_camera_omnidirectional(const float* matrix, int32_t w, int32_t h, LrCamera* out) {
  if (!matrix || !out || w <= 0 || h <= 0) return fail(LR_ERR_INVALID, "bad argument");
  camera_omnidirectional(m_from(matrix), w, h, *out);
  return LR_OK;
}
int lr_camera_pinhole(const float* position, const float* aperture_position, const float* sensor_size, int32_t w, int32_t h,
                      float aperture_radius, LrCamera* out) {
  if (!position || !aperture_position || !sensor_size || !out || w <= 0 || h <= 0) return fail(LR_ERR_INVALID, "bad argument");
  camera_pinhole(vec3(position), vec3(aperture_position), sensor_size, w, h, aperture_radius, *out);
  return LR_OK;
}
void lr_matrix_unit(float* m) { const Mat4 r = mat4_unit(); std::memcpy(m, r.m, sizeof(r.m)); }
void lr_matrix_translate(const float* v, float* m) { const Mat4 r = mat4_translate(vec3(v)); std::memcpy(m, r.m, sizeof(r.m)); }
void lr_matrix_scale(const float* v, float* m) { const Mat4 r = mat4_scale(vec3(v)); std::memcpy(m, r.m, sizeof(r.m)); }
void lr_matrix_axis_angle(const float* axis, float angle_deg, float* m) {
  const Mat4 r = mat4_axis_angle(vec3(axis), angle_deg * kHostPI / 180.0f);      // scene_loader.rs:93
  std::memcpy(m, r.m, sizeof(r.m));
}
void lr_matrix_look_at(const float* origin, const float* target, const float* up, float* m) {
  const Mat4 r = mat4_look_at(vec3(origin), vec3(target), vec3(up));
  std::memcpy(m, r.m, sizeof(r.m));
}
void lr_matrix_mul(const float* a, const float* b, float* out) { const Mat4 r = mat4_mul(m_from(a), m_from(b)); std::memcpy(out, r.m, sizeof(r.m)); }
void lr_matrix_apply(const float* m, const float* v3, float* out3) { const Vec3 r = mat4_apply(m_from(m), vec3(v3)); out3[0] = r[0]; out3[1] = r[1]; out3[2] = r[2]; }

}  // extern "C"
