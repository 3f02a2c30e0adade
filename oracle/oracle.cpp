// oracle.cpp — CPU restatement of LumillyRender's path-tracing hot path.
//
// TEST INFRASTRUCTURE ONLY (see oracle.h).  Every function cites the reference file:line it
// follows.  The structure mirrors the reference on purpose: recursive radiance estimators,
// a boxed-tree-like SAH BVH with one primitive per leaf, unordered/unpruned traversal that
// collects candidates and then takes the minimum distance.  Compile with
// `-ffp-contract=off` so that fp32 expression order equals what rustc/LLVM emits
// (SURVEY.md §7 hard part 1).  All arithmetic is fp32 like the reference's Vector3.
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <random>
#include <thread>
#include <vector>

namespace {

// src/constant.rs:1-3
constexpr float PI = 3.14159265358979323846264338327950288f;
constexpr float EPS = 1e-3f;
constexpr float INF = 1e5f;

// ---------------------------------------------------------------- math (src/math/vector3.rs)
struct V3 {
  float x, y, z;
  float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 from3(const float* p) { return V3{p[0], p[1], p[2]}; }
inline void to3(V3 v, float* p) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }                         // vector3.rs:92-98
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }    // :100-106
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }    // :108-114
inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }       // :116-122
inline V3 operator*(float s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }       // :124-130
inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }    // :132-138
inline V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }       // :140-146
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }         // :76-80
inline V3 cross(V3 a, V3 b) {                                                      // :82-90
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float sqr_norm(V3 a) { return dot(a, a); }                                  // traits.rs:30-32
inline float norm(V3 a) { return std::sqrt(sqr_norm(a)); }                         // traits.rs:26-28
inline V3 normalize(V3 a) { return a / norm(a); }                                  // traits.rs:38-42
// Rust f32::max/min = IEEE maxNum/minNum (NaN-ignoring) == fmaxf/fminf
inline float rmax(float a, float b) { return std::fmax(a, b); }
inline float rmin(float a, float b) { return std::fmin(a, b); }
// llvm.powi expansion for a constant/non-negative exponent (repeated squaring, compiler-rt __powisf2)
inline float powi(float a, int b) {
  const bool recip = b < 0;
  float r = 1.0f;
  while (true) {
    if (b & 1) r *= a;
    b /= 2;
    if (b == 0) break;
    a *= a;
  }
  return recip ? 1.0f / r : r;
}

// ---------------------------------------------------------------- sin/cos
// The reference calls f32::sin / f32::cos = the platform libm (not bit-defined across platforms).
// math mode 0 uses libm like the reference; math mode 1 uses the fp32 algorithm specified for the device
// (csrc/device_path.cuh: spec_sincos, restated here operation for operation) so that replay comparisons
// are not perturbed by 1-ulp libm differences.  tests/ check that the two modes differ by <= 2 ulp.
thread_local int g_math_mode = 0;
inline void spec_sincos(float x, float* sn, float* cs) {
  const float kf = std::nearbyint(x * 0.636619772367581343f);
  const int k = (int)kf;
  float r = x - kf * 1.5703125f;
  r = r - kf * 4.837512969970703125e-4f;
  r = r - kf * 7.54978995489188216e-8f;
  const float z = r * r;
  const float s = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * r + r;
  const float c = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
  switch (k & 3) {
    case 0: *sn = s; *cs = c; break;
    case 1: *sn = c; *cs = -s; break;
    case 2: *sn = -s; *cs = -c; break;
    default: *sn = -c; *cs = s; break;
  }
}
inline float osin(float x) { if (g_math_mode == 0) return std::sin(x); float s, c; spec_sincos(x, &s, &c); return s; }
inline float ocos(float x) { if (g_math_mode == 0) return std::cos(x); float s, c; spec_sincos(x, &s, &c); return c; }

// ---------------------------------------------------------------- RNG
// The reference draws rand::random::<f32>() (rand 0.3: 24 random mantissa bits, [0,1)) from an
// OS-seeded thread-local generator, so its renders are not reproducible and
// parity is statistical by construction.  The oracle offers two U[0,1) sources:
//   mode 0: the counter-based PCG32 stream specified for the device (state seeded from
//           (seed, pixel, sample)), enabling per-sample replay comparisons;
//   mode 1: std::mt19937 per pixel, an unrelated stream for the statistical tests.
inline uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
struct Rng {
  int mode = 0;                 // 0 counter-based PCG, 1 mt19937, 2 scripted (unit tests)
  uint64_t state = 0;
  std::mt19937 mt;
  const float* script = nullptr;
  int script_pos = 0;
  void seed_counter(uint64_t seed, uint32_t pixel, uint32_t sample) {
    mode = 0;
    state = splitmix64(splitmix64(seed + 0x632BE59BD9B4E019ULL * (uint64_t)pixel) ^
                       ((uint64_t)sample * 0xD1B54A32D192ED03ULL));
    next_u32();
  }
  void seed_mt(uint64_t seed, uint32_t pixel) {
    mode = 1;
    std::seed_seq sq{(uint32_t)seed, (uint32_t)(seed >> 32), pixel, 0x5bd1e995u};
    mt.seed(sq);
  }
  uint32_t next_u32() {
    if (mode == 1) return (uint32_t)mt();
    const uint64_t old = state;
    state = old * 6364136223846793005ULL + 1442695040888963407ULL;
    const uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    const uint32_t rot = (uint32_t)(old >> 59u);
    return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31u));
  }
  float next() {
    if (mode == 2) return script[script_pos++];
    return (float)(next_u32() >> 8) * (1.0f / 16777216.0f);
  }
};

// ---------------------------------------------------------------- ray / intersection
struct Ray { V3 origin, direction; };                                   // src/ray.rs:3-6
struct Intersection {                                                   // src/intersection.rs:5-10
  V3 position; float distance; V3 normal; int material; int prim;
};

// ---------------------------------------------------------------- AABB (src/aabb.rs)
struct AABB { V3 min, max, center; };

inline float aabb_surface_area(const AABB& b) {                          // aabb.rs:17-28
  const V3 side = v3(std::fabs(b.max.x - b.min.x), std::fabs(b.max.y - b.min.y), std::fabs(b.max.z - b.min.z));
  return 2.0f * (side.x * side.y + side.y * side.z + side.z * side.x);
}
inline AABB aabb_merge_with(const AABB& a, const AABB& v) {              // aabb.rs:48-64
  const V3 mn = v3(rmin(a.min.x, v.min.x), rmin(a.min.y, v.min.y), rmin(a.min.z, v.min.z));
  const V3 mx = v3(rmax(a.max.x, v.max.x), rmax(a.max.y, v.max.y), rmax(a.max.z, v.max.z));
  return AABB{mn, mx, (mn + mx) / 2.0f};
}
inline AABB aabb_empty() {                                               // aabb.rs:66-72
  const float inf = std::numeric_limits<float>::infinity();
  return AABB{v3(inf, inf, inf), v3(-inf, -inf, -inf), v3(0, 0, 0)};
}
// aabb.rs:75-92 — slab test of the infinite line clipped to t in [-INF, INF]
inline bool aabb_is_intersect(const AABB& b, const Ray& ray) {
  float mn = -INF;
  float mx = INF;
  for (int i = 0; i < 3; i++) {
    const float inv_d = 1.0f / ray.direction[i];
    const float t1 = (b.min[i] - ray.origin[i]) * inv_d;
    const float t2 = (b.max[i] - ray.origin[i]) * inv_d;
    float t_min, t_max;
    if (t1 > t2) { t_min = t2; t_max = t1; } else { t_min = t1; t_max = t2; }
    if (mn < t_min) mn = t_min;
    if (mx > t_max) mx = t_max;
    if (mn > mx) return false;
  }
  return true;
}

// ---------------------------------------------------------------- sampling utilities (src/util.rs)
inline void orthonormal_basis(V3 n, V3& tangent, V3& binormal) {         // util.rs:12-21
  const V3 a = std::fabs(n.x) > EPS ? v3(0.0f, 1.0f, 0.0f) : v3(1.0f, 0.0f, 0.0f);
  tangent = normalize(cross(a, n));
  binormal = cross(n, tangent);
}
inline V3 reflect(V3 v, V3 normal) { return -v + normal * (dot(v, normal) * 2.0f); }  // util.rs:30-32
inline bool refract(V3 v, V3 normal, float from_per_to_ior, V3& out) {                 // util.rs:34-42
  const float dn = dot(v, normal);
  const float cos2theta = 1.0f - powi(from_per_to_ior, 2) * (1.0f - powi(dn, 2));
  if (cos2theta > 0.0f) {
    out = -v * from_per_to_ior - normal * (from_per_to_ior * -dn + std::sqrt(cos2theta));
    return true;
  }
  return false;
}
inline V3 hemisphere_cos_importance(float xi1, float xi2) {              // util.rs:87-96
  const float r1 = 2.0f * PI * xi1;
  const float r2 = xi2;
  const float r2s = std::sqrt(r2);
  return v3(ocos(r1) * r2s, osin(r1) * r2s, std::sqrt(1.0f - r2));
}
inline V3 sphere_uniform(float xi1, float xi2) {                         // util.rs:108-116
  const float r1 = 2.0f * PI * xi1;
  const float r2 = xi2 * 2.0f - 1.0f;
  const float r2s = std::sqrt(1.0f - r2 * r2);
  return v3(ocos(r1) * r2s, osin(r1) * r2s, r2);
}

// ---------------------------------------------------------------- materials (src/material/*.rs)
inline V3 orienting_normal(V3 out_, V3 normal) {                         // lambert.rs:14-21 (same in all)
  if (dot(normal, out_) < 0.0f) return normal * -1.0f;
  return normal;
}
inline float signed_mod(float base, float module) {                      // lambert.rs:58-64
  if (base > 0.0f) return std::fmod(base, module);
  return module - std::fmod(-base, module);
}
inline V3 checker(float u, float v) {                                    // lambert.rs:66-90
  const float lw = 2.0f, li = 150.0f, sw = 1.0f, si = 30.0f, cw = 150.0f, ci = 300.0f;
  const float lu = signed_mod(u, li), lv = signed_mod(v, li);
  const float su = signed_mod(u, si), sv = signed_mod(v, si);
  const float cu = signed_mod(u, ci), cv = signed_mod(v, ci);
  if (lu < lw || lv < lw) return v3(0.5f, 0.5f, 0.5f);
  if (su < sw || sv < sw) return v3(0.6f, 0.6f, 0.6f);
  if ((cu < cw || cv < cw) && !(cu < cw && cv < cw)) return v3(0.8f, 0.8f, 0.8f);
  return v3(1.0f, 1.0f, 1.0f);
}

struct MatSample { V3 value; float pdf; };

inline V3 mat_color(const LrMaterial& m) { return from3(m.color); }
inline V3 mat_emission(const LrMaterial& m) {                            // lambert.rs:23-25; others return zero
  return m.type == LR_MAT_LAMBERT ? from3(m.emission) : v3(0, 0, 0);
}
inline float mat_weight(const LrMaterial& m) {                           // lambert.rs:27-30, phong.rs:32-37, ...
  return rmax(rmax(m.color[0], m.color[1]), m.color[2]);
}

// ggx.rs:19-45
inline float ggx_alpha(const LrMaterial& m) { return m.param0 * m.param0; }
inline float g_ggx(const LrMaterial& m, V3 v, V3 n) {
  const float a2 = ggx_alpha(m) * ggx_alpha(m);
  const float c = dot(v, n);
  const float tan = 1.0f / (c * c) - 1.0f;
  return 2.0f / (1.0f + std::sqrt(1.0f + a2 * tan * tan));
}
inline float gaf_smith(const LrMaterial& m, V3 out_, V3 in_, V3 n) { return g_ggx(m, in_, n) * g_ggx(m, out_, n); }
inline float ggx_ndf(const LrMaterial& m, V3 mm, V3 n) {
  const float a2 = ggx_alpha(m) * ggx_alpha(m);
  const float mdn = dot(mm, n);
  const float x = (a2 - 1.0f) * mdn * mdn + 1.0f;
  return a2 / (PI * x * x);
}
inline float fresnel_schlick(const LrMaterial& m, V3 in_, V3 mm) {
  const float nnn = 1.0f - m.param1;
  const float nnp = 1.0f + m.param1;
  const float f_0 = (nnn * nnn) / (nnp * nnp);
  const float c = dot(in_, mm);
  return f_0 + (1.0f - f_0) * powi(1.0f - c, 5);
}

// ideal_refraction.rs:117-160
inline void ior_pair(const LrMaterial& m, V3 out_, V3 n, float& from_ior, float& to_ior) {
  const float ior_v = 1.0f;
  const float ior = m.param1;
  if (dot(out_, n) > 0.0f) { from_ior = ior_v; to_ior = ior; } else { from_ior = ior; to_ior = ior_v; }
}
inline float fresnel_exact(float from_ior, float to_ior, V3 out_, V3 in_, V3 on) {
  const float cos1 = dot(out_, on);
  const float cos2 = dot(in_, -on);
  const float n1 = from_ior, n2 = to_ior;
  const float rs = powi((n1 * cos1 - n2 * cos2) / (n1 * cos1 + n2 * cos2), 2);
  const float rp = powi((n1 * cos2 - n2 * cos1) / (n1 * cos2 + n2 * cos1), 2);
  return (rs + rp) / 2.0f;
}

// Material::brdf
inline V3 mat_brdf(const LrMaterial& m, V3 out_, V3 in_, V3 n, V3 pos) {
  switch (m.type) {
    case LR_MAT_LAMBERT:                                                 // lambert.rs:32-35
      return mat_color(m) * checker(pos.x, pos.z) / PI;
    case LR_MAT_PHONG: {                                                 // phong.rs:39-47
      const V3 on = orienting_normal(out_, n);
      if (dot(in_, on) <= 0.0f) return v3(0, 0, 0);
      const V3 r = reflect(out_, on);
      const float c = dot(r, in_);
      const float a = m.param0;
      return mat_color(m) * ((a + 2.0f) / (2.0f * PI) * std::pow(c, a));
    }
    case LR_MAT_BLINN_PHONG: {                                           // blinn_phong.rs:39-49
      const V3 on = orienting_normal(out_, n);
      if (dot(in_, on) <= 0.0f) return v3(0, 0, 0);
      const V3 h = normalize(in_ + out_);
      const float c = dot(h, on);
      const float a = m.param0;
      return mat_color(m) * ((a + 2.0f) * (a + 4.0f) / (8.0f * PI * (std::pow(2.0f, -a / 2.0f) + a)) * std::pow(c, a));
    }
    case LR_MAT_GGX: {                                                   // ggx.rs:71-85
      const V3 on = orienting_normal(out_, n);
      if (dot(in_, on) <= 0.0f) return v3(0, 0, 0);
      const V3 h = normalize(in_ + out_);
      const float f = fresnel_schlick(m, in_, h);
      const float g = gaf_smith(m, out_, in_, on);
      const float d = ggx_ndf(m, h, on);
      return mat_color(m) * f * g * d / (4.0f * dot(in_, on) * dot(out_, on));
    }
    case LR_MAT_IDEAL_REFRACTION: {                                      // ideal_refraction.rs:40-68
      const V3 on = orienting_normal(out_, n);
      float from_ior, to_ior;
      ior_pair(m, out_, n, from_ior, to_ior);
      const float from_per_to_ior = from_ior / to_ior;
      V3 r;
      if (refract(out_, on, from_per_to_ior, r)) {
        const float fr = fresnel_exact(from_ior, to_ior, out_, r, on);
        if (dot(in_, on) > 0.0f) {
          return mat_color(m) * 1.0f / dot(in_, n) * fr;
        } else {
          const float ft = (1.0f - fr) * powi(to_ior / from_ior, 2);
          return mat_color(m) * 1.0f / dot(in_, n) * ft;
        }
      }
      return mat_color(m) * 1.0f / dot(in_, n);
    }
  }
  return v3(0, 0, 0);
}

// Material::sample — draws in the reference's order
inline MatSample mat_sample(const LrMaterial& m, V3 out_, V3 n, Rng& rng) {
  switch (m.type) {
    case LR_MAT_LAMBERT: {                                               // lambert.rs:37-55
      const V3 on = orienting_normal(out_, n);
      const V3 w = on;
      V3 u, v;
      orthonormal_basis(w, u, v);
      const float xi1 = rng.next();
      const float xi2 = rng.next();
      const V3 s = hemisphere_cos_importance(xi1, xi2);
      const V3 in_ = u * s.x + v * s.y + w * s.z;
      const float cos_term = dot(in_, n);
      return MatSample{in_, cos_term / PI};
    }
    case LR_MAT_PHONG: {                                                 // phong.rs:49-69
      const V3 on = orienting_normal(out_, n);
      const float a = m.param0;
      const V3 r = reflect(out_, on);
      const V3 w = r;
      V3 u, v;
      orthonormal_basis(w, u, v);
      const float r1 = 2.0f * PI * rng.next();
      const float r2 = rng.next();
      const float t = std::pow(r2, 1.0f / (a + 2.0f));
      const float ts = std::sqrt(1.0f - t * t);
      const V3 in_ = u * ocos(r1) * ts + v * osin(r1) * ts + w * t;
      const float c = dot(r, in_);
      return MatSample{in_, (a + 2.0f) / (2.0f * PI) * std::pow(c, a)};
    }
    case LR_MAT_BLINN_PHONG: {                                           // blinn_phong.rs:51-73
      const V3 on = orienting_normal(out_, n);
      const float a = m.param0;
      const V3 w = on;
      V3 u, v;
      orthonormal_basis(w, u, v);
      const float r1 = 2.0f * PI * rng.next();
      const float r2 = rng.next();
      const float t = std::pow(r2, 1.0f / (a + 2.0f));
      const float ts = std::sqrt(1.0f - t * t);
      const V3 h = u * ocos(r1) * ts + v * osin(r1) * ts + w * t;
      const V3 in_ = h * (2.0f * dot(out_, h)) - out_;
      const float c = dot(on, h);
      return MatSample{in_, (a + 2.0f) / (2.0f * PI) * std::pow(c, a)};
    }
    case LR_MAT_GGX: {                                                   // ggx.rs:87-113
      const V3 on = orienting_normal(out_, n);
      const V3 w = on;
      V3 u, v;
      orthonormal_basis(w, u, v);
      const float r1 = 2.0f * PI * rng.next();
      const float r2 = rng.next();
      const float tan = ggx_alpha(m) * std::sqrt(r2 / (1.0f - r2));
      const float x = 1.0f + tan * tan;
      const float c = 1.0f / std::sqrt(x);
      const float s = tan / std::sqrt(x);
      const V3 h = u * ocos(r1) * s + v * osin(r1) * s + w * c;
      const float o_h = dot(out_, h);
      const V3 in_ = h * (2.0f * o_h) - out_;
      const float jacobian = 1.0f / (4.0f * o_h);
      return MatSample{in_, ggx_ndf(m, h, on) * dot(h, on) * jacobian};
    }
    case LR_MAT_IDEAL_REFRACTION: {                                      // ideal_refraction.rs:70-104
      float from_ior, to_ior;
      ior_pair(m, out_, n, from_ior, to_ior);
      const float from_per_to_ior = from_ior / to_ior;
      const V3 on = orienting_normal(out_, n);
      V3 r;
      if (refract(out_, on, from_per_to_ior, r)) {
        const float fr = fresnel_exact(from_ior, to_ior, out_, r, on);
        const float rr_prob = fr;
        if (rng.next() < rr_prob) return MatSample{reflect(out_, on), 1.0f * rr_prob};
        return MatSample{r, 1.0f * (1.0f - rr_prob)};
      }
      return MatSample{reflect(out_, on), 1.0f};
    }
  }
  return MatSample{v3(0, 0, 0), 0.0f};
}

// Material::coef — traits.rs:20-22 default, ideal_refraction.rs:106-113 override
inline V3 mat_coef(const LrMaterial& m, V3 out_, V3 n, float fly_distance) {
  if (m.type == LR_MAT_IDEAL_REFRACTION && dot(out_, n) < 0.0f) {
    const V3 v = -(v3(1.0f, 1.0f, 1.0f) - mat_color(m)) * m.param0 * fly_distance;
    return v3(std::exp(v.x), std::exp(v.y), std::exp(v.z));
  }
  return v3(1.0f, 1.0f, 1.0f);
}

// ---------------------------------------------------------------- shapes
struct Prim {
  int kind;            // 0 triangle, 1 sphere
  V3 p0, p1, p2;       // triangle.rs:15-23
  V3 normal;           // triangle.rs:36
  V3 center; float radius;  // sphere.rs:13-19
  float area;
  int material;
  AABB aabb;
};

inline AABB triangle_aabb(V3 p0, V3 p1, V3 p2) {                         // triangle.rs:102-119
  const V3 mn = v3(rmin(rmin(p0.x, p1.x), p2.x), rmin(rmin(p0.y, p1.y), p2.y), rmin(rmin(p0.z, p1.z), p2.z));
  const V3 mx = v3(rmax(rmax(p0.x, p1.x), p2.x), rmax(rmax(p0.y, p1.y), p2.y), rmax(rmax(p0.z, p1.z), p2.z));
  return AABB{mn, mx, (mx + mn) / 2.0f};
}
inline Prim make_triangle(V3 p0, V3 p1, V3 p2, int material) {           // triangle.rs:25-40
  Prim t{};
  t.kind = 0; t.p0 = p0; t.p1 = p1; t.p2 = p2;
  t.aabb = triangle_aabb(p0, p1, p2);
  t.normal = normalize(cross(p1 - p0, p2 - p0));
  t.area = norm(cross(p1 - p0, p2 - p0)) * 0.5f;
  t.material = material;
  return t;
}
inline Prim make_sphere(V3 position, float radius, int material) {       // sphere.rs:21-39
  Prim s{};
  s.kind = 1; s.center = position; s.radius = radius;
  s.area = 4.0f * PI * powi(radius, 2);
  s.material = material;
  const V3 r = v3(radius, radius, radius);
  s.aabb = AABB{position - r, position + r, position};
  return s;
}

// triangle.rs:69-100 — Möller–Trumbore
inline bool triangle_intersect_mt(const Prim& tr, const Ray& ray, Intersection& out) {
  const V3 e1 = tr.p1 - tr.p0;
  const V3 e2 = tr.p2 - tr.p0;
  const V3 pv = cross(ray.direction, e2);
  const float det = dot(e1, pv);
  if (std::fabs(det) < EPS) return false;
  const float invdet = 1.0f / det;
  const V3 tv = ray.origin - tr.p0;
  const float u = dot(tv, pv) * invdet;
  if (u < 0.0f || u > 1.0f) return false;
  const V3 qv = cross(tv, e1);
  const float v = dot(ray.direction, qv) * invdet;
  if (v < 0.0f || u + v > 1.0f) return false;
  const float t = dot(e2, qv) * invdet;
  if (t < EPS) return false;
  const V3 p = ray.origin + ray.direction * t;
  out.distance = t; out.normal = tr.normal; out.position = p; out.material = tr.material;
  return true;
}
// triangle.rs:42-67 — three cross products (test-only cross-check in the reference)
inline bool triangle_intersect_3c(const Prim& tr, const Ray& ray, Intersection& out) {
  const float dn = dot(ray.direction, tr.normal);
  const float t = dot(tr.p0 - ray.origin, tr.normal) / dn;
  if (t < EPS) return false;
  const V3 p = ray.origin + ray.direction * t;
  const V3 c0 = cross(tr.p1 - tr.p0, p - tr.p0);
  if (dot(c0, tr.normal) < 0.0f) return false;
  const V3 c1 = cross(tr.p2 - tr.p1, p - tr.p1);
  if (dot(c1, tr.normal) < 0.0f) return false;
  const V3 c2 = cross(tr.p0 - tr.p2, p - tr.p2);
  if (dot(c2, tr.normal) < 0.0f) return false;
  out.distance = t; out.normal = tr.normal; out.position = p; out.material = tr.material;
  return true;
}
// sphere.rs:42-63
inline bool sphere_intersect(const Prim& s, const Ray& ray, Intersection& out) {
  const V3 co = ray.origin - s.center;
  const float cod = dot(co, ray.direction);
  const float det = cod * cod - sqr_norm(co) + s.radius * s.radius;
  if (det <= 0.0f) return false;
  const float t1 = -cod - std::sqrt(det);
  const float t2 = -cod + std::sqrt(det);
  if (t1 < EPS && t2 < EPS) return false;
  const float distance = t1 > EPS ? t1 : t2;
  const V3 position = ray.origin + ray.direction * distance;
  const V3 outer_normal = normalize(position - s.center);
  out.distance = distance; out.position = position; out.normal = outer_normal; out.material = s.material;
  return true;
}
inline bool prim_intersect(const Prim& p, const Ray& ray, Intersection& out) {
  return p.kind == 0 ? triangle_intersect_mt(p, ray, out) : sphere_intersect(p, ray, out);
}

// ---------------------------------------------------------------- BVH (src/bvh.rs)
struct BvhNode {          // Node{aabb,left,right} / Leaf{aabb,index}  bvh.rs:10-36
  AABB aabb;
  int left = -1, right = -1;   // children (node indices); leaf iff index >= 0
  int index = -1;
};
struct LeafItem { AABB aabb; int index; };

struct Bvh {
  std::vector<BvhNode> nodes;
  int root = -1;

  // bvh.rs:69-127 — full-sweep SAH over 3 axes, centre sort, 1 primitive per leaf
  int construct(LeafItem* list, size_t n) {
    const float t_aabb = 1.0f;
    const float t_tri = 2.0f;
    if (n == 1) {
      BvhNode leaf; leaf.aabb = list[0].aabb; leaf.index = list[0].index;
      nodes.push_back(leaf);
      return (int)nodes.size() - 1;
    }
    AABB aabb = aabb_empty();
    int partition_axis = 0; size_t partition_index = 1; float best_t = 0.0f; bool have = false;
    std::vector<float> s1_a, s2_a;
    for (int axis = 0; axis < 3; axis++) {
      std::sort(list, list + n, [axis](const LeafItem& a, const LeafItem& b) { return a.aabb.center[axis] < b.aabb.center[axis]; });
      AABB s1_aabb = list[0].aabb;
      s1_a.clear();
      for (size_t i = 0; i < n; i++) { s1_aabb = aabb_merge_with(s1_aabb, list[i].aabb); s1_a.push_back(aabb_surface_area(s1_aabb)); }
      AABB s2_aabb = list[n - 1].aabb;
      s2_a.clear();
      for (size_t i = n - 1; i >= 1; i--) { s2_aabb = aabb_merge_with(s2_aabb, list[i].aabb); s2_a.push_back(aabb_surface_area(s2_aabb)); }
      aabb = aabb_merge_with(s1_aabb, list[n - 1].aabb);
      const float s_a = aabb_surface_area(aabb);
      size_t axis_best_i = 0; float axis_best_t = 0.0f; bool axis_have = false;
      for (size_t i = 0; i + 1 < n; i++) {
        const float s1_n = (float)(i + 1);
        const float s2_n = (float)(n - i - 1);
        const float t = 2.0f * t_aabb + (s1_a[i] * s1_n + s2_a[n - i - 2] * s2_n) * t_tri / s_a;
        // OrderedFloat ordering: NaN sorts greatest; min_by_key keeps the first minimum
        const bool less = !axis_have || (std::isnan(axis_best_t) ? !std::isnan(t) : t < axis_best_t);
        if (less) { axis_best_t = t; axis_best_i = i; axis_have = true; }
      }
      const bool less = !have || (std::isnan(best_t) ? !std::isnan(axis_best_t) : axis_best_t < best_t);
      if (less) { best_t = axis_best_t; partition_axis = axis; partition_index = axis_best_i + 1; have = true; }
    }
    std::sort(list, list + n, [partition_axis](const LeafItem& a, const LeafItem& b) { return a.aabb.center[partition_axis] < b.aabb.center[partition_axis]; });
    const int left = construct(list, partition_index);
    const int right = construct(list + partition_index, n - partition_index);
    BvhNode node; node.aabb = aabb; node.left = left; node.right = right;
    nodes.push_back(node);
    return (int)nodes.size() - 1;
  }
};

// ---------------------------------------------------------------- scene container
struct SceneImpl {
  std::vector<LrMaterial> materials;
  std::vector<Prim> prims;              // index = prim_id (Loader.instances order)
  Bvh bvh;
  std::vector<int> emission;            // objects.rs:19-23 (instance order)
  float emission_area = 0.0f;           // objects.rs:24
  LrCamera camera{};
  LrSky sky{};
  std::vector<float> sky_pixels;
  double build_seconds = 0.0;
};

struct Counters { uint64_t rays = 0, nodes = 0, prims = 0; };

// bvh.rs:21-25,39-44 — recursive, unordered, unpruned candidate collection
static void may_intersect(const SceneImpl& sc, int node, const Ray& ray, std::vector<int>& candidate, Counters& c) {
  const BvhNode& n = sc.bvh.nodes[node];
  c.nodes++;
  if (n.index >= 0) {
    if (aabb_is_intersect(n.aabb, ray)) candidate.push_back(n.index);
    return;
  }
  if (aabb_is_intersect(n.aabb, ray)) {
    may_intersect(sc, n.left, ray, candidate, c);
    may_intersect(sc, n.right, ray, candidate, c);
  }
}

// bvh.rs:131-141 — candidates in DFS order, first minimum distance wins (Iterator::min_by)
// tie_by_index = false: the reference's rule — among candidates at the same distance the FIRST in the depth-first order of
// its tree wins (min_by keeps the first minimum, bvh.rs:136-140).  That order is a property of the reference's own SAH
// build.  tie_by_index = true: the lowest primitive index wins instead — the rule of the device path, whose trees differ
// from the reference's (and from each other: host SAH / device radix tree); it makes the nearest hit a function of the
// scene alone.  The two rules differ only when two primitives answer with bit-identical distances (rays through shared
// vertices / edges of a mesh).
static bool bvh_intersect_faithful(const SceneImpl& sc, const Ray& ray, Intersection& best, Counters& c, bool tie_by_index = false) {
  if (sc.bvh.root < 0) return false;
  std::vector<int> candidate;
  may_intersect(sc, sc.bvh.root, ray, candidate, c);
  bool have = false;
  for (int i : candidate) {
    Intersection it;
    c.prims++;
    if (!prim_intersect(sc.prims[i], ray, it)) continue;
    it.prim = i;
    if (std::isnan(it.distance)) continue;   // the reference panics here (partial_cmp().unwrap()); treated as a miss
    if (!have || it.distance < best.distance || (tie_by_index && it.distance == best.distance && i < best.prim)) { best = it; have = true; }
  }
  return have;
}

// Same acceptance rule (leaf AABB line test AND primitive test, minimum distance), but the tree is
// walked near-to-far with conservative distance pruning.  A parent's slab interval contains its
// children's (monotone fp rounding), so pruning by the reference's own node test loses nothing.
static bool bvh_intersect_fast(const SceneImpl& sc, const Ray& ray, Intersection& best, Counters& c) {
  if (sc.bvh.root < 0) return false;
  const V3 inv = v3(1.0f / ray.direction.x, 1.0f / ray.direction.y, 1.0f / ray.direction.z);
  auto entry = [&](const AABB& b, float& t_entry) {
    float mn = -INF, mx = INF;
    for (int i = 0; i < 3; i++) {
      const float t1 = (b.min[i] - ray.origin[i]) * inv[i];
      const float t2 = (b.max[i] - ray.origin[i]) * inv[i];
      float t_min, t_max;
      if (t1 > t2) { t_min = t2; t_max = t1; } else { t_min = t1; t_max = t2; }
      if (mn < t_min) mn = t_min;
      if (mx > t_max) mx = t_max;
      if (mn > mx) return false;
    }
    t_entry = mn;
    return true;
  };
  bool have = false;
  int stack[256];
  int sp = 0;
  stack[sp++] = sc.bvh.root;
  float cull = std::numeric_limits<float>::infinity();
  while (sp > 0) {
    const int ni = stack[--sp];
    const BvhNode& n = sc.bvh.nodes[ni];
    c.nodes++;
    float t_entry;
    if (!entry(n.aabb, t_entry)) continue;
    if (t_entry > cull) continue;
    if (n.index >= 0) {
      Intersection it;
      c.prims++;
      if (!prim_intersect(sc.prims[n.index], ray, it)) continue;
      it.prim = n.index;
      if (std::isnan(it.distance)) continue;
      if (!have || it.distance < best.distance) {
        best = it; have = true;
        // slack: a primitive's computed t may be slightly before its box's computed entry
        cull = best.distance * 1.001f + 1.0f + 0.01f * std::fabs(best.distance);
      }
      continue;
    }
    if (sp + 2 > 256) return bvh_intersect_faithful(sc, ray, best, c);
    stack[sp++] = n.right;
    stack[sp++] = n.left;
  }
  return have;
}

static bool brute_intersect(const SceneImpl& sc, const Ray& ray, Intersection& best) {
  bool have = false;
  for (size_t i = 0; i < sc.prims.size(); i++) {
    if (!aabb_is_intersect(sc.prims[i].aabb, ray)) continue;   // a leaf's own AABB gate (bvh.rs:21-25)
    Intersection it;
    if (!prim_intersect(sc.prims[i], ray, it)) continue;
    it.prim = (int)i;
    if (std::isnan(it.distance)) continue;
    if (!have || it.distance < best.distance) { best = it; have = true; }
  }
  return have;
}

struct RayRecord { Ray ray; int prim; float distance; };
struct Tracer {
  const SceneImpl& sc;
  int traversal;
  Counters c;
  std::vector<RayRecord>* record = nullptr;      // orc_trace_path: every closest-hit query of one sample, in order
  // Objects::intersect objects.rs:63-65
  bool intersect(const Ray& ray, Intersection& out) {
    c.rays++;
    // traversal 0: the reference's algorithm and tie rule; 1: pruned near-to-far walk (same hits; fast);
    // 2: the reference's algorithm with the device's tie rule (lowest primitive index among equal distances)
    const bool hit = traversal == 1 ? bvh_intersect_fast(sc, ray, out, c) : bvh_intersect_faithful(sc, ray, out, c, traversal == 2);
    if (record) record->push_back(RayRecord{ray, hit ? out.prim : -1, hit ? out.distance : 0.0f});
    return hit;
  }
};

// ---------------------------------------------------------------- sky (src/sky.rs)
inline size_t f32_to_usize(float v) {   // Rust `as usize`: saturating, NaN -> 0
  if (!(v > 0.0f)) return 0;
  if (v >= 1.8446744e19f) return std::numeric_limits<size_t>::max();
  return (size_t)v;
}
static V3 sky_radiance(const LrSky& sky, const float* pixels, const Ray& ray) {
  if (sky.type == LR_SKY_UNIFORM) return from3(sky.color);                // sky.rs:17-21
  // sky.rs:57-79
  const float theta = std::acos(ray.direction.y);
  const float phi = std::atan2(ray.direction.z, ray.direction.x);
  const float u = std::fmod((phi + PI + sky.longitude_offset) / (2.0f * PI), 1.0f);
  const float v = std::fmod(theta / PI, 1.0f);
  const size_t height = (size_t)sky.height;
  const size_t width = height * 2;
  const size_t all = width * height;
  const size_t x = f32_to_usize(std::floor((float)width * u));
  const size_t y = f32_to_usize(std::floor((float)height * v));
  const size_t index = y * width + x;
  const float* c = pixels + 3 * (index % all);
  return v3(c[0], c[1], c[2]);
}

// ---------------------------------------------------------------- emitters (src/objects.rs)
struct PointSample { V3 value; float pdf; };

static PointSample prim_sample(const Prim& p, Rng& rng) {
  if (p.kind == 0) {                                                     // triangle.rs:140-149
    const float u = rng.next();
    const float v = rng.next();
    const float mn = rmin(u, v);
    const float mx = rmax(u, v);
    return PointSample{p.p0 * mn + p.p1 * (1.0f - mx) + p.p2 * (mx - mn), 1.0f / p.area};
  }
  const float xi1 = rng.next();                                          // sphere.rs:79-84
  const float xi2 = rng.next();
  return PointSample{p.center + p.radius * sphere_uniform(xi1, xi2), 1.0f / p.area};
}
static PointSample sample_emission(const SceneImpl& sc, Rng& rng) {      // objects.rs:37-51
  const float roulette = sc.emission_area * rng.next();
  float area = 0.0f;
  for (int idx : sc.emission) {
    const Prim& obj = sc.prims[idx];
    area += obj.area;
    if (roulette <= area) {
      const PointSample s = prim_sample(obj, rng);
      return PointSample{s.value, s.pdf * obj.area / sc.emission_area};
    }
  }
  // unreachable!() in the reference
  const Prim& obj = sc.prims[sc.emission.back()];
  const PointSample s = prim_sample(obj, rng);
  return PointSample{s.value, s.pdf * obj.area / sc.emission_area};
}

// ---------------------------------------------------------------- integrators (src/scene.rs)
struct Integrator {
  const SceneImpl& sc;
  Tracer tr;
  Rng& rng;
  int depth_cfg, depth_limit;
  bool no_direct_emitter;

  V3 sky(const Ray& ray) { return sky_radiance(sc.sky, sc.sky_pixels.data(), ray); }

  // scene.rs:64-76
  float russian_roulette(float init, int depth) {
    float continue_rr_prob = init;
    if (depth > depth_limit) continue_rr_prob *= powi(0.5f, depth - depth_limit);
    if (depth <= depth_cfg && continue_rr_prob > 0.0f) continue_rr_prob = 1.0f;
    return continue_rr_prob;
  }

  // scene.rs:78-102
  template <class F>
  V3 material_interaction_radiance(const Intersection& i, const Ray& ray, F f) {
    const LrMaterial& m = sc.materials[i.material];
    const V3 out_ = -ray.direction;
    const MatSample sample = mat_sample(m, out_, i.normal, rng);
    const V3 in_ = sample.value;
    const float pdf = sample.pdf;
    const V3 brdf = mat_brdf(m, out_, in_, i.normal, i.position);
    const V3 coef = mat_coef(m, out_, i.normal, i.distance);
    const float c = dot(in_, i.normal);
    const Ray new_ray{i.position, in_};
    const V3 l_i = f(new_ray);
    return brdf * coef * l_i * c / pdf;
  }

  // scene.rs:104-151
  V3 direct_light_radiance(const Intersection& i, const Ray& ray) {
    const LrMaterial& m = sc.materials[i.material];
    if (sqr_norm(mat_emission(m)) > 0.0f || !(sc.emission_area > 0.0f)) return v3(0, 0, 0);
    const PointSample direct_sample = sample_emission(sc, rng);
    const V3 direct_path = direct_sample.value - i.position;
    const Ray direct_ray{i.position, normalize(direct_path)};
    const V3 point_in = direct_ray.direction;
    const V3 point_out = -ray.direction;
    const V3 point_normal = orienting_normal(point_out, i.normal);
    if (dot(point_in, point_normal) <= 0.0f) return v3(0, 0, 0);
    Intersection direct_i;
    if (tr.intersect(direct_ray, direct_i)) {
      if (std::fabs(direct_i.distance - norm(direct_path)) > EPS) return v3(0, 0, 0);
      const V3 light_out = -direct_ray.direction;
      const V3 light_normal = direct_i.normal;
      const float light_cos = dot(light_out, light_normal);
      if (light_cos <= 0.0f) return v3(0, 0, 0);
      const float point_cos = dot(point_in, point_normal);
      const float g_term = point_cos * light_cos / sqr_norm(direct_path);
      const V3 brdf = mat_brdf(m, point_out, point_in, point_normal, i.position);
      const V3 l_i = mat_emission(sc.materials[direct_i.material]);
      const float pdf = direct_sample.pdf;
      return brdf * l_i * g_term / pdf;
    }
    return v3(0, 0, 0);
  }

  // scene.rs:24-32
  V3 radiance_recursive(const Ray& ray, int depth) {
    Intersection i;
    if (!tr.intersect(ray, i)) return sky(ray);
    return intersect_radiance(i, ray, depth);
  }
  // scene.rs:153-171
  V3 intersect_radiance(const Intersection& i, const Ray& ray, int depth) {
    const LrMaterial& m = sc.materials[i.material];
    const V3 l_e = (!(no_direct_emitter && depth == 0) && dot(-ray.direction, i.normal) > 0.0f) ? mat_emission(m) : v3(0, 0, 0);
    const float continue_rr_prob = russian_roulette(mat_weight(m), depth);
    if (continue_rr_prob != 1.0f && rng.next() >= continue_rr_prob) return l_e;
    const V3 material_radiance = material_interaction_radiance(i, ray, [&](const Ray& new_ray) { return radiance_recursive(new_ray, depth + 1); });
    return l_e + material_radiance / continue_rr_prob;
  }
  // scene.rs:38-46
  V3 radiance_nee_recursive(const Ray& ray, int depth, bool no_emission) {
    Intersection i;
    if (!tr.intersect(ray, i)) return sky(ray);
    return intersect_radiance_nee(i, ray, depth, no_emission);
  }
  // scene.rs:173-193
  V3 intersect_radiance_nee(const Intersection& i, const Ray& ray, int depth, bool no_emission) {
    const LrMaterial& m = sc.materials[i.material];
    const V3 l_e = (!(no_direct_emitter && depth == 0) && !no_emission && dot(-ray.direction, i.normal) > 0.0f) ? mat_emission(m) : v3(0, 0, 0);
    const float continue_rr_prob = russian_roulette(mat_weight(m), depth);
    if (continue_rr_prob != 1.0f && rng.next() >= continue_rr_prob) return l_e;
    const V3 direct = direct_light_radiance(i, ray);
    const V3 material_radiance = material_interaction_radiance(i, ray, [&](const Ray& new_ray) { return radiance_nee_recursive(new_ray, depth + 1, true); });
    return l_e + (direct + material_radiance) / continue_rr_prob;
  }
};

// ---------------------------------------------------------------- cameras (src/camera.rs)
struct CamSample { Ray ray; float pdf; float g_term; };

inline V3 cam_sample_sensor(const LrCamera& c, int left, int top, float u, float v) {   // camera.rs:64-81, 266-283, 411-428
  const float px = ((((float)left + u) / (float)c.width) - 0.5f) * c.sensor_size[0];
  const float py = ((((float)top + v) / (float)c.height) - 0.5f) * c.sensor_size[1];
  return from3(c.position) - from3(c.right) * px + from3(c.up) * py;
}
inline V3 cam_sample_aperture(const LrCamera& c, float xi1, float xi2) {                // camera.rs:285-300, 430-445
  const float u = 2.0f * PI * xi1;
  const float v = std::sqrt(xi2) * c.aperture_radius;
  const float px = ocos(u) * v;
  const float py = osin(u) * v;
  return from3(c.aperture_position) + from3(c.right) * px + from3(c.up) * py;
}
inline float cam_geometry_term(const LrCamera& c, V3 direction) {                       // camera.rs:302-309, 447-454
  const float cos_term = dot(direction, from3(c.forward));
  const float d = c.aperture_sensor_distance / cos_term;
  return cos_term * cos_term / (d * d);
}
template <class Draw>
static CamSample camera_sample(const LrCamera& c, int x, int y, Draw draw) {
  switch (c.type) {
    case LR_CAM_IDEAL_PINHOLE: {                                                        // camera.rs:100-115
      const float u = draw(); const float v = draw();
      const V3 sensor = cam_sample_sensor(c, x, y, u, v);
      const V3 ap = from3(c.aperture_position);
      const Ray ray{ap, normalize(ap - sensor)};
      return CamSample{ray, 1.0f * 1.0f, 1.0f};
    }
    case LR_CAM_OMNIDIRECTIONAL: {                                                      // camera.rs:168-188
      const float u = draw(); const float v = draw();
      const float p = ((float)x + u) / (float)c.width * PI * 2.0f;
      const float t = ((float)y + v) / (float)c.height * PI;
      const V3 direction = v3(osin(t) * ocos(p), osin(t) * osin(p), ocos(t));
      return CamSample{Ray{from3(c.aperture_position), direction}, 1.0f, 1.0f};
    }
    case LR_CAM_PINHOLE: {                                                              // camera.rs:313-328
      const float u = draw(); const float v = draw();
      const V3 sensor = cam_sample_sensor(c, x, y, u, v);
      const float sensor_pdf = 1.0f / c.sensor_pixel_area;
      const float a1 = draw(); const float a2 = draw();
      const V3 ap = cam_sample_aperture(c, a1, a2);
      const float ap_pdf = 1.0f / (PI * c.aperture_radius * c.aperture_radius);
      const Ray ray{ap, normalize(ap - sensor)};
      return CamSample{ray, sensor_pdf * ap_pdf, cam_geometry_term(c, ray.direction)};
    }
    case LR_CAM_THIN_LENS: {                                                            // camera.rs:458-476
      const float u = draw(); const float v = draw();
      const V3 sensor = cam_sample_sensor(c, x, y, u, v);
      const float sensor_pdf = 1.0f / c.sensor_pixel_area;
      const float a1 = draw(); const float a2 = draw();
      const V3 ap = cam_sample_aperture(c, a1, a2);
      const float ap_pdf = 1.0f / (PI * c.aperture_radius * c.aperture_radius);
      const V3 apc = from3(c.aperture_position);
      const V3 sensor_center = apc - sensor;
      const V3 object_plane = sensor_center * (c.focus_distance / dot(sensor_center, from3(c.forward)));
      const Ray ray{ap, normalize(apc + object_plane - ap)};
      return CamSample{ray, sensor_pdf * ap_pdf, cam_geometry_term(c, normalize(ap - sensor))};
    }
  }
  return CamSample{};
}

// ---------------------------------------------------------------- matrices (src/math/matrix4.rs)
struct M4 { float v[16]; };
inline M4 m4_unit() { return M4{{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}}; }                   // matrix4.rs:9-18
inline M4 m4_translate(V3 t) { return M4{{1, 0, 0, t.x, 0, 1, 0, t.y, 0, 0, 1, t.z, 0, 0, 0, 1}}; }    // :20-29
inline M4 m4_scale(V3 s) { return M4{{s.x, 0, 0, 0, 0, s.y, 0, 0, 0, 0, s.z, 0, 0, 0, 0, 1}}; }        // :31-40
inline M4 m4_axis_angle(V3 a, float t) {                                                                // :42-54
  const float c = std::cos(t), s = std::sin(t);
  return M4{{c + a.x * a.x * (1.0f - c), a.x * a.y * (1.0f - c) - a.z * s, a.x * a.z * (1.0f - c) + a.y * s, 0.0f,
             a.y * a.x * (1.0f - c) + a.z * s, c + a.y * a.y * (1.0f - c), a.y * a.z * (1.0f - c) - a.x * s, 0.0f,
             a.z * a.x * (1.0f - c) - a.y * s, a.z * a.y * (1.0f - c) + a.x * s, c + a.z * a.z * (1.0f - c), 0.0f,
             0.0f, 0.0f, 0.0f, 1.0f}};
}
inline M4 m4_look_at(V3 origin, V3 target, V3 up) {                                                     // :56-68
  const V3 za = normalize(origin - target);
  const V3 xa = normalize(cross(up, za));
  const V3 ya = cross(za, xa);
  return M4{{xa.x, xa.y, xa.z, 0.0f, ya.x, ya.y, ya.z, 0.0f, za.x, za.y, za.z, 0.0f, origin.x, origin.y, origin.z, 1.0f}};
}
inline float dot4(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]; }   // vector4.rs:71-75
inline M4 m4_mul(const M4& a, const M4& b) {                                                            // :201-211
  M4 o;
  for (int y = 0; y < 4; y++)
    for (int x = 0; x < 4; x++) {
      const float col[4] = {b.v[x], b.v[x + 4], b.v[x + 8], b.v[x + 12]};
      o.v[y * 4 + x] = dot4(&a.v[y * 4], col);
    }
  return o;
}
inline V3 m4_apply(const M4& m, V3 p) {                                                                 // :185-199 with vector4.rs:40-44 (w = 1)
  const float v4[4] = {p.x, p.y, p.z, 1.0f};
  return v3(dot4(&m.v[0], v4), dot4(&m.v[4], v4), dot4(&m.v[8], v4));
}

}  // namespace

// ================================================================= C entry points
struct OrcScene { SceneImpl impl; };

extern "C" {

int orc_scene_create(const LrSceneDesc* d, OrcScene** out) {
  if (!d || !out) return LR_ERR_INVALID;
  auto s = std::make_unique<OrcScene>();
  SceneImpl& sc = s->impl;
  sc.materials.assign(d->materials, d->materials + d->n_materials);
  const int n = d->n_triangles + d->n_spheres;
  sc.prims.resize(n);
  std::vector<char> seen(n, 0);
  for (int i = 0; i < d->n_triangles; i++) {
    const LrTriangle& t = d->triangles[i];
    if (t.prim_id < 0 || t.prim_id >= n || seen[t.prim_id]) return LR_ERR_INVALID;
    seen[t.prim_id] = 1;
    sc.prims[t.prim_id] = make_triangle(from3(t.p0), from3(t.p1), from3(t.p2), t.material);
  }
  for (int i = 0; i < d->n_spheres; i++) {
    const LrSphere& sp = d->spheres[i];
    if (sp.prim_id < 0 || sp.prim_id >= n || seen[sp.prim_id]) return LR_ERR_INVALID;
    seen[sp.prim_id] = 1;
    sc.prims[sp.prim_id] = make_sphere(from3(sp.center), sp.radius, sp.material);
  }
  // objects.rs:18-29
  for (int i = 0; i < n; i++)
    if (sqr_norm(mat_emission(sc.materials[sc.prims[i].material])) > 0.0f) sc.emission.push_back(i);
  float area = 0.0f;
  for (int i : sc.emission) area += sc.prims[i].area;
  sc.emission_area = area;
  // bvh.rs:57-67
  const auto t0 = std::chrono::steady_clock::now();
  if (n > 0) {
    std::vector<LeafItem> leaf(n);
    for (int i = 0; i < n; i++) leaf[i] = LeafItem{sc.prims[i].aabb, i};
    sc.bvh.nodes.reserve(2 * (size_t)n);
    sc.bvh.root = sc.bvh.construct(leaf.data(), (size_t)n);
  }
  sc.build_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  sc.camera = d->camera;
  sc.sky = d->sky;
  if (d->sky.type == LR_SKY_IBL) {
    if (!d->sky.pixels || d->sky.n_pixels < 2LL * d->sky.height * d->sky.height) return LR_ERR_INVALID;
    sc.sky_pixels.assign(d->sky.pixels, d->sky.pixels + 3 * d->sky.n_pixels);
    sc.sky.pixels = nullptr;
  }
  *out = s.release();
  return LR_OK;
}

void orc_scene_destroy(OrcScene* s) { delete s; }
int orc_scene_nodes(const OrcScene* s) { return (int)s->impl.bvh.nodes.size(); }

// main.rs:92-121 — one job per pixel, all samples of the pixel summed in order
int orc_render(const OrcScene* s, const LrRenderParams* p, int traversal, int rng_mode, int math_mode, int n_threads,
               int pixel_stride, float* out_sum, float* out_sumsq, OrcStats* stats) {
  if (!s || !p || !out_sum) return LR_ERR_INVALID;
  const SceneImpl& sc = s->impl;
  const int W = sc.camera.width, H = sc.camera.height;
  const int cx = p->crop_w > 0 ? p->crop_x : 0, cy = p->crop_w > 0 ? p->crop_y : 0;
  const int cw = p->crop_w > 0 ? p->crop_w : W, ch = p->crop_w > 0 ? p->crop_h : H;
  if (cx < 0 || cy < 0 || cx + cw > W || cy + ch > H) return LR_ERR_INVALID;
  if (pixel_stride < 1) pixel_stride = 1;
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  const int64_t n_pixels = (int64_t)cw * ch;
  std::atomic<int64_t> next{0};
  std::vector<Counters> counters(n_threads);
  std::vector<uint64_t> nonfinite(n_threads, 0), nsamples(n_threads, 0);
  const auto t0 = std::chrono::steady_clock::now();
  auto worker = [&](int tid) {
    g_math_mode = math_mode;
    Rng rng;
    Integrator in{sc, Tracer{sc, traversal, Counters{}}, rng, p->depth, p->depth_limit, p->no_direct_emitter != 0};
    const int64_t chunk = 64;
    while (true) {
      const int64_t begin = next.fetch_add(chunk);
      if (begin >= n_pixels) break;
      const int64_t end = std::min(n_pixels, begin + chunk);
      for (int64_t pi = begin; pi < end; pi++) {
        const int lx = (int)(pi % cw), ly = (int)(pi / cw);
        if (pixel_stride > 1 && ((lx % pixel_stride) != 0 || (ly % pixel_stride) != 0)) {
          out_sum[3 * pi + 0] = out_sum[3 * pi + 1] = out_sum[3 * pi + 2] = 0.0f;
          if (out_sumsq) out_sumsq[3 * pi + 0] = out_sumsq[3 * pi + 1] = out_sumsq[3 * pi + 2] = 0.0f;
          continue;
        }
        const int x = cx + lx, y = cy + ly;
        const uint32_t pixel = (uint32_t)(y * W + x);
        if (rng_mode == 1) rng.seed_mt(p->seed, pixel);
        V3 sum = v3(0, 0, 0), sumsq = v3(0, 0, 0);
        for (int si = 0; si < p->spp_count; si++) {
          const int sidx = p->spp_begin + si;
          if (rng_mode == 0) rng.seed_counter(p->seed, pixel, (uint32_t)sidx);
          const CamSample cs = camera_sample(sc.camera, x, y, [&]() { return rng.next(); });
          const V3 l = p->integrator == LR_INTEGRATOR_PT ? in.radiance_recursive(cs.ray, 0)
                                                          : in.radiance_nee_recursive(cs.ray, 0, false);
          const V3 e = l * cs.g_term;                                       // main.rs:99
          const V3 delta = e * (sc.camera.sensor_sensitivity / cs.pdf);    // main.rs:101
          if (!(std::isfinite(delta.x) && std::isfinite(delta.y) && std::isfinite(delta.z))) nonfinite[tid]++;
          sum = sum + delta;                                               // main.rs:102
          sumsq = sumsq + delta * delta;
          nsamples[tid]++;
        }
        to3(sum, out_sum + 3 * pi);
        if (out_sumsq) to3(sumsq, out_sumsq + 3 * pi);
      }
    }
    counters[tid] = in.tr.c;
  };
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; t++) th.emplace_back(worker, t);
  for (auto& t : th) t.join();
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    for (int t = 0; t < n_threads; t++) {
      stats->rays += counters[t].rays; stats->nodes_visited += counters[t].nodes; stats->prims_tested += counters[t].prims;
      stats->nonfinite_samples += nonfinite[t]; stats->samples += nsamples[t];
    }
    stats->render_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    stats->build_seconds = sc.build_seconds;
    stats->threads = n_threads;
    stats->bvh_nodes = (int)sc.bvh.nodes.size();
  }
  return LR_OK;
}

// Scene::normal / Scene::depth (scene.rs:48-62) over the camera rays of the sample range, averaged per pixel in sample
// order.  The reference defines the two functions but main.rs never calls them; the driver loop here is main.rs:92-104
// with the AOV in place of the radiance (no g_term / pdf weight: an AOV is not a radiance).  Streams: the counter-based
// one shared with the device (math_mode 1: the lens cameras' aperture sample uses the specified sincos).
int orc_render_aov(const OrcScene* s, const LrRenderParams* p, int kind, int traversal, int n_threads, float* out) {
  if (!s || !p || !out || p->spp_count <= 0) return LR_ERR_INVALID;
  const SceneImpl& sc = s->impl;
  const int W = sc.camera.width, H = sc.camera.height;
  const int cx = p->crop_w > 0 ? p->crop_x : 0, cy = p->crop_w > 0 ? p->crop_y : 0;
  const int cw = p->crop_w > 0 ? p->crop_w : W, ch = p->crop_w > 0 ? p->crop_h : H;
  if (cx < 0 || cy < 0 || cx + cw > W || cy + ch > H) return LR_ERR_INVALID;
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  std::atomic<int> next_row{0};
  auto worker = [&]() {
    g_math_mode = 1;
    Rng rng;
    Tracer tr{sc, traversal, Counters{}};
    while (true) {
      const int ly = next_row.fetch_add(1);
      if (ly >= ch) break;
      for (int lx = 0; lx < cw; lx++) {
        const int x = cx + lx, y = cy + ly;
        const uint32_t pixel = (uint32_t)(y * W + x);
        V3 sum = v3(0, 0, 0);
        for (int si = 0; si < p->spp_count; si++) {
          rng.seed_counter(p->seed, pixel, (uint32_t)(p->spp_begin + si));
          const CamSample cs = camera_sample(sc.camera, x, y, [&]() { return rng.next(); });
          Intersection it;
          V3 v = v3(0, 0, 0);                                                     // None => Vector3::zero() / 0.0
          if (tr.intersect(cs.ray, it)) v = kind == 0 ? it.normal / 2.0f + v3(0.5f, 0.5f, 0.5f)   // scene.rs:52
                                                      : v3(it.distance, 0, 0);                    // scene.rs:60
          sum = sum + v;
        }
        const size_t i = (size_t)ly * cw + lx;
        const float n = (float)p->spp_count;
        if (kind == 0) { out[3 * i] = sum.x / n; out[3 * i + 1] = sum.y / n; out[3 * i + 2] = sum.z / n; }
        else out[i] = sum.x / n;
      }
    }
  };
  std::vector<std::thread> th;
  for (int k = 0; k < n_threads; k++) th.emplace_back(worker);
  for (auto& k : th) k.join();
  return LR_OK;
}

// Every closest-hit query (Objects::intersect, objects.rs:63-65) that ONE sample of one pixel issues, in order: camera ray,
// then per vertex the shadow ray (pt-direct) and the extension ray — with what each hit.  A debugging probe for replay
// divergences: the rays can be handed to the device's nearest-hit probes one by one.  Shared counter-based stream.
int orc_trace_path(const OrcScene* s, const LrRenderParams* p, int x, int y, int sample, int traversal, int max_rays,
                   float* origins, float* directions, int32_t* prim, float* t, int32_t* n_rays) {
  if (!s || !p || !origins || !directions || !prim || !t || !n_rays) return LR_ERR_INVALID;
  const SceneImpl& sc = s->impl;
  if (x < 0 || y < 0 || x >= sc.camera.width || y >= sc.camera.height) return LR_ERR_INVALID;
  g_math_mode = 1;
  Rng rng;
  std::vector<RayRecord> rec;
  Integrator in{sc, Tracer{sc, traversal, Counters{}, &rec}, rng, p->depth, p->depth_limit, p->no_direct_emitter != 0};
  rng.seed_counter(p->seed, (uint32_t)(y * sc.camera.width + x), (uint32_t)sample);
  const CamSample cs = camera_sample(sc.camera, x, y, [&]() { return rng.next(); });
  if (p->integrator == LR_INTEGRATOR_PT) in.radiance_recursive(cs.ray, 0); else in.radiance_nee_recursive(cs.ray, 0, false);
  *n_rays = (int32_t)rec.size();
  for (int i = 0; i < (int)rec.size() && i < max_rays; i++) {
    to3(rec[i].ray.origin, origins + 3 * i); to3(rec[i].ray.direction, directions + 3 * i);
    prim[i] = rec[i].prim; t[i] = rec[i].distance;
  }
  return LR_OK;
}

int orc_trace_primary(const OrcScene* s, float u, float v, float ua, float va, int traversal, int n_threads,
                      int32_t* prim, float* t) {
  if (!s || !prim || !t) return LR_ERR_INVALID;
  const SceneImpl& sc = s->impl;
  const int W = sc.camera.width, H = sc.camera.height;
  if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
  std::atomic<int> next_row{0};
  auto worker = [&]() {
    g_math_mode = 1;   // the probe is a GPU-parity tool: aperture sampling uses the specified sincos
    Tracer tr{sc, traversal, Counters{}};
    while (true) {
      const int y = next_row.fetch_add(1);
      if (y >= H) break;
      for (int x = 0; x < W; x++) {
        int k = 0;
        const float draws[4] = {u, v, ua, va};
        const CamSample cs = camera_sample(sc.camera, x, y, [&]() { return draws[k++]; });
        Intersection it;
        const size_t i = (size_t)y * W + x;
        if (tr.intersect(cs.ray, it)) { prim[i] = it.prim; t[i] = it.distance; } else { prim[i] = -1; t[i] = 0.0f; }
      }
    }
  };
  std::vector<std::thread> th;
  for (int k = 0; k < n_threads; k++) th.emplace_back(worker);
  for (auto& k : th) k.join();
  return LR_OK;
}

int orc_trace_rays(const OrcScene* s, int64_t n, const float* origins, const float* directions, int traversal,
                   int brute_force, int32_t* prim, float* t, float* normal) {
  if (!s) return LR_ERR_INVALID;
  const SceneImpl& sc = s->impl;
  Tracer tr{sc, traversal, Counters{}};
  for (int64_t i = 0; i < n; i++) {
    const Ray ray{from3(origins + 3 * i), from3(directions + 3 * i)};
    Intersection it;
    const bool hit = brute_force ? brute_intersect(sc, ray, it) : tr.intersect(ray, it);
    prim[i] = hit ? it.prim : -1;
    t[i] = hit ? it.distance : 0.0f;
    if (normal) to3(hit ? it.normal : v3(0, 0, 0), normal + 3 * i);
  }
  return LR_OK;
}

void orc_set_math_mode(int mode) { g_math_mode = mode; }
void orc_spec_sincos(float x, float* s, float* c) { spec_sincos(x, s, c); }

int orc_camera_sample(const LrCamera* cam, int x, int y, float u, float v, float ua, float va, float* out9) {
  int k = 0;
  const float draws[4] = {u, v, ua, va};
  const CamSample cs = camera_sample(*cam, x, y, [&]() { return draws[k++]; });
  to3(cs.ray.origin, out9); to3(cs.ray.direction, out9 + 3);
  out9[6] = cs.pdf; out9[7] = cs.g_term; out9[8] = cam->sensor_sensitivity;
  return LR_OK;
}

static Prim tri_from9(const float* p9) { return make_triangle(from3(p9), from3(p9 + 3), from3(p9 + 6), 0); }
int orc_triangle_intersect_mt(const float* p9, const float* o, const float* d, float* t, float* pos, float* n) {
  Intersection it;
  if (!triangle_intersect_mt(tri_from9(p9), Ray{from3(o), from3(d)}, it)) return 0;
  *t = it.distance; to3(it.position, pos); to3(it.normal, n);
  return 1;
}
int orc_triangle_intersect_3c(const float* p9, const float* o, const float* d, float* t, float* pos, float* n) {
  Intersection it;
  if (!triangle_intersect_3c(tri_from9(p9), Ray{from3(o), from3(d)}, it)) return 0;
  *t = it.distance; to3(it.position, pos); to3(it.normal, n);
  return 1;
}
int orc_sphere_intersect(const float* center, float radius, const float* o, const float* d, float* t, float* pos, float* n) {
  Intersection it;
  if (!sphere_intersect(make_sphere(from3(center), radius, 0), Ray{from3(o), from3(d)}, it)) return 0;
  *t = it.distance; to3(it.position, pos); to3(it.normal, n);
  return 1;
}
int orc_aabb_is_intersect(const float* lo, const float* hi, const float* o, const float* d) {
  return aabb_is_intersect(AABB{from3(lo), from3(hi), v3(0, 0, 0)}, Ray{from3(o), from3(d)}) ? 1 : 0;
}
void orc_reflect(const float* v, const float* n, float* out) { to3(reflect(from3(v), from3(n)), out); }
int orc_refract(const float* v, const float* n, float eta, float* out) {
  V3 r;
  if (!refract(from3(v), from3(n), eta, r)) return 0;
  to3(r, out);
  return 1;
}
void orc_orthonormal_basis(const float* n, float* tangent, float* binormal) {
  V3 t, b;
  orthonormal_basis(from3(n), t, b);
  to3(t, tangent); to3(b, binormal);
}
float orc_checker(float u, float v) { return checker(u, v).x; }
int orc_material_brdf(const LrMaterial* m, const float* out_, const float* in_, const float* n, const float* pos, float* brdf3) {
  to3(mat_brdf(*m, from3(out_), from3(in_), from3(n), from3(pos)), brdf3);
  return LR_OK;
}
int orc_material_sample(const LrMaterial* m, const float* out_, const float* n, float r1, float r2, float* in3, float* pdf) {
  const float draws[2] = {r1, r2};
  Rng rng;
  rng.mode = 2; rng.script = draws; rng.script_pos = 0;
  const MatSample s = mat_sample(*m, from3(out_), from3(n), rng);
  to3(s.value, in3); *pdf = s.pdf;
  return LR_OK;
}
float orc_material_weight(const LrMaterial* m) { return mat_weight(*m); }
void orc_material_coef(const LrMaterial* m, const float* out_, const float* n, float dist, float* coef3) {
  to3(mat_coef(*m, from3(out_), from3(n), dist), coef3);
}
float orc_fresnel(float from_ior, float to_ior, const float* out_, const float* in_, const float* on) {
  return fresnel_exact(from_ior, to_ior, from3(out_), from3(in_), from3(on));
}
void orc_ior_pair(const LrMaterial* m, const float* out_, const float* n, float* from_ior, float* to_ior) {
  ior_pair(*m, from3(out_), from3(n), *from_ior, *to_ior);
}
void orc_sky_radiance(const LrSky* sky, const float* d, float* rgb) {
  to3(sky_radiance(*sky, sky->pixels, Ray{v3(0, 0, 0), from3(d)}), rgb);
}
float orc_rng_float(uint64_t seed, uint32_t pixel, uint32_t sample, int index) {
  Rng r;
  r.seed_counter(seed, pixel, sample);
  float f = 0.0f;
  for (int i = 0; i <= index; i++) f = r.next();
  return f;
}

static M4 m4_from(const float* m) { M4 o; std::memcpy(o.v, m, sizeof(o.v)); return o; }
void orc_matrix_unit(float* m) { const M4 o = m4_unit(); std::memcpy(m, o.v, sizeof(o.v)); }
void orc_matrix_translate(const float* v, float* m) { const M4 o = m4_translate(from3(v)); std::memcpy(m, o.v, sizeof(o.v)); }
void orc_matrix_scale(const float* v, float* m) { const M4 o = m4_scale(from3(v)); std::memcpy(m, o.v, sizeof(o.v)); }
void orc_matrix_axis_angle(const float* axis, float angle_deg, float* m) {   // scene_loader.rs:93
  const M4 o = m4_axis_angle(from3(axis), angle_deg * PI / 180.0f);
  std::memcpy(m, o.v, sizeof(o.v));
}
void orc_matrix_look_at(const float* origin, const float* target, const float* up, float* m) {
  const M4 o = m4_look_at(from3(origin), from3(target), from3(up));
  std::memcpy(m, o.v, sizeof(o.v));
}
void orc_matrix_mul(const float* a, const float* b, float* out) { const M4 o = m4_mul(m4_from(a), m4_from(b)); std::memcpy(out, o.v, sizeof(o.v)); }
void orc_matrix_apply(const float* m, const float* v, float* out) { to3(m4_apply(m4_from(m), from3(v)), out); }

// camera.rs:34-62
void orc_camera_ideal_pinhole(const float* matrix, float xfov, int w, int h, LrCamera* out) {
  const M4 m = m4_from(matrix);
  std::memset(out, 0, sizeof(*out));
  const V3 aperture_position = v3(m.v[12], m.v[13], m.v[14]);
  const V3 forward = m4_apply(m, v3(0.0f, 0.0f, -1.0f));
  const V3 right = m4_apply(m, v3(1.0f, 0.0f, 0.0f));
  const V3 up = m4_apply(m, v3(0.0f, 1.0f, 0.0f));
  const V3 direction = forward * 50.0f;
  const V3 position = aperture_position - direction;
  const float asd = norm(direction);
  const float sx = 2.0f * asd * std::tan(xfov * PI / 180.0f / 2.0f);
  const float sy = sx * (float)h / (float)w;
  out->type = LR_CAM_IDEAL_PINHOLE; out->width = w; out->height = h;
  to3(forward, out->forward); to3(right, out->right); to3(up, out->up);
  to3(position, out->position); to3(aperture_position, out->aperture_position);
  out->sensor_size[0] = sx; out->sensor_size[1] = sy;
  out->aperture_sensor_distance = asd;
  out->sensor_sensitivity = 1.0f;
}
// camera.rs:366-409
void orc_camera_thin_lens(const float* matrix, float xfov, float focus_distance, float f_number, int w, int h, LrCamera* out) {
  orc_camera_ideal_pinhole(matrix, xfov, w, h, out);
  out->type = LR_CAM_THIN_LENS;
  const float asd = out->aperture_sensor_distance;
  const float focal_length = 1.0f / (1.0f / asd + 1.0f / focus_distance);
  const float aperture_radius = focal_length / f_number / 2.0f;
  const float spa = (out->sensor_size[0] * out->sensor_size[1]) / (float)((size_t)w * (size_t)h);
  const float sens = asd * asd / (spa * PI * aperture_radius * aperture_radius);
  out->aperture_radius = aperture_radius;
  out->sensor_pixel_area = spa;
  out->sensor_sensitivity = sens;
  out->focus_distance = focus_distance;
}
// camera.rs:149-166
void orc_camera_omnidirectional(const float* matrix, int w, int h, LrCamera* out) {
  const M4 m = m4_from(matrix);
  std::memset(out, 0, sizeof(*out));
  out->type = LR_CAM_OMNIDIRECTIONAL; out->width = w; out->height = h;
  to3(m4_apply(m, v3(0.0f, 0.0f, -1.0f)), out->forward);
  to3(m4_apply(m, v3(1.0f, 0.0f, 0.0f)), out->right);
  to3(m4_apply(m, v3(0.0f, 1.0f, 0.0f)), out->up);
  to3(v3(m.v[12], m.v[13], m.v[14]), out->aperture_position);
  out->sensor_sensitivity = 1.0f;
}
// camera.rs:224-264
void orc_camera_pinhole(const float* position, const float* aperture_position, const float* sensor_size, int w, int h,
                        float aperture_radius, LrCamera* out) {
  std::memset(out, 0, sizeof(*out));
  const V3 pos = from3(position), ap = from3(aperture_position);
  const V3 direction = ap - pos;
  const float asd = norm(direction);
  const V3 forward = normalize(direction);
  const V3 right = normalize(cross(forward, std::fabs(forward.y) < 1.0f - EPS ? v3(0.0f, 1.0f, 0.0f) : v3(1.0f, 0.0f, 0.0f)));
  const V3 up = cross(right, forward);
  const float spa = (sensor_size[0] * sensor_size[1]) / (float)((size_t)w * (size_t)h);
  const float sens = asd * asd / (spa * PI * aperture_radius * aperture_radius);
  out->type = LR_CAM_PINHOLE; out->width = w; out->height = h;
  to3(forward, out->forward); to3(right, out->right); to3(up, out->up);
  to3(pos, out->position); to3(ap, out->aperture_position);
  out->sensor_size[0] = sensor_size[0]; out->sensor_size[1] = sensor_size[1];
  out->aperture_radius = aperture_radius; out->aperture_sensor_distance = asd;
  out->sensor_pixel_area = spa; out->sensor_sensitivity = sens;
}
float orc_triangle_area(const float* p9) { return tri_from9(p9).area; }
float orc_sphere_area(float r) { return make_sphere(v3(0, 0, 0), r, 0).area; }

}  // extern "C"
