// common.h — host-side helpers shared by the C-ABI implementation and the host front end.
// Host fp32 code that feeds the device (areas, camera blocks, transforms) is compiled with
// -ffp-contract=off so that it rounds like the reference's Rust (SURVEY.md §7 hard part 1).
#pragma once
#include <cmath>
#include <string>

#include "../../include/lumilly.h"

namespace lr {

int fail(int code, const std::string& msg);      // records the thread-local message, returns code
void set_error(const std::string& msg);

constexpr float kHostPI = 3.14159265358979323846264338327950288f;   // constant.rs:1

// ---- fp32 vector helpers with the reference's operation order (math/vector3.rs:76-146) ----
struct Vec3 {
  float v[3];
  float& operator[](int i) { return v[i]; }
  float operator[](int i) const { return v[i]; }
};
inline Vec3 vec3(float x, float y, float z) { return Vec3{{x, y, z}}; }
inline Vec3 vec3(const float* p) { return Vec3{{p[0], p[1], p[2]}}; }
inline Vec3 vsub(Vec3 a, Vec3 b) { return vec3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline Vec3 vadd(Vec3 a, Vec3 b) { return vec3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline Vec3 vscale(Vec3 a, float s) { return vec3(a[0] * s, a[1] * s, a[2] * s); }
inline Vec3 vdiv(Vec3 a, float s) { return vec3(a[0] / s, a[1] / s, a[2] / s); }
inline float vdot(Vec3 a, Vec3 b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline Vec3 vcross(Vec3 a, Vec3 b) {
  return vec3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
inline float vnorm(Vec3 a) { return std::sqrt(vdot(a, a)); }
inline Vec3 vnormalize(Vec3 a) { return vdiv(a, vnorm(a)); }           // traits.rs:38-42

// triangle.rs:37 — area = |(p1-p0) x (p2-p0)| * 0.5
inline float triangle_area(const float* p0, const float* p1, const float* p2) {
  const Vec3 a = vec3(p0), b = vec3(p1), c = vec3(p2);
  return vnorm(vcross(vsub(b, a), vsub(c, a))) * 0.5f;
}
// sphere.rs:25 — area = 4.0 * PI * radius.powi(2)
inline float sphere_area(float r) { return 4.0f * kHostPI * (r * r); }

int validate_desc(const LrSceneDesc& d);          // scene_validate.cpp

}  // namespace lr

// Nothing may throw across the C ABI: entry points that allocate on the host run their body through this.
#define LR_GUARDED(call)                                                                                   \
  try {                                                                                                    \
    return (call);                                                                                         \
  } catch (const std::bad_alloc&) {                                                                        \
    return lr::fail(LR_ERR_UNSUPPORTED, "out of host memory");                                             \
  } catch (const std::exception& e__) {                                                                    \
    return lr::fail(LR_ERR_INVALID, std::string("internal error: ") + e__.what());                         \
  }
