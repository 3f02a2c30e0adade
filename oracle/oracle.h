/* oracle.h — C entry points of the CPU oracle (TEST INFRASTRUCTURE, not product code).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  Nothing under lumillyrender_b200/ includes, links or
 * calls it; the product fails loudly when its CUDA library is missing.
 *
 * PARITY PIN STATUS: the reference (Rust nightly, 11 un-vendored crates) cannot be built
 * or run in this environment, and it ships no golden files.  The oracle is pinned against
 * every known-answer test the reference holds for this path (src/triangle.rs:157-235,
 * src/util.rs:49-81, src/material/ideal_refraction.rs:167-312) and against the camera
 * set-up known answers of SURVEY.md Appendix C.  Everything else on the path is
 * "parity unpinned" by the reference itself and is pinned here by
 * line-by-line restatement plus internal cross-checks (brute force vs BVH, furnace tests).
 */
#ifndef LUMILLY_ORACLE_H
#define LUMILLY_ORACLE_H
#include "../include/lumilly.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcStats {
  uint64_t rays, samples, nodes_visited, prims_tested, nonfinite_samples;
  double build_seconds, render_seconds;
  int32_t threads, bvh_nodes;
} OrcStats;

typedef struct OrcScene OrcScene;

/* traversal: 0 = faithful reference algorithm (bvh.rs:131-141: unordered, unpruned, collect
 *               candidates then min_by), 1 = ordered + pruned traversal of the same tree with
 *               the same per-primitive acceptance rule (identical nearest hit, faster).
 * rng_mode:  0 = the counter-based PCG stream shared with the device (replay),
 *            1 = an independent std::mt19937 stream per pixel (statistical tests).        */
int orc_scene_create(const LrSceneDesc* desc, OrcScene** out);
void orc_scene_destroy(OrcScene* s);
int orc_scene_nodes(const OrcScene* s);

/* out_sum: crop_w*crop_h*3 per-pixel SUM over the sample range (not the mean),
 * out_sumsq: nullable, per-pixel sum of squares.  Pixel jobs are distributed over n_threads
 * (0 = hardware_concurrency), each pixel runs all its samples in order (main.rs:92-104). */
/* math_mode: 0 = libm sin/cos like the reference, 1 = the fp32 sincos specified for the device
 *            (bit-identical sampled directions on CPU and GPU). */
int orc_render(const OrcScene* s, const LrRenderParams* p, int traversal, int rng_mode, int math_mode,
               int n_threads, int pixel_stride, float* out_sum, float* out_sumsq, OrcStats* stats);
void orc_set_math_mode(int mode);     /* for the calling thread (unit-level entry points) */
void orc_spec_sincos(float x, float* s, float* c);

/* Scene::normal (kind 0, 3 floats per pixel) / Scene::depth (kind 1, 1 float per pixel), scene.rs:48-62, of the
 * camera rays of the sample range, averaged per pixel in sample order (shared counter-based stream). */
int orc_render_aov(const OrcScene* s, const LrRenderParams* p, int kind, int traversal, int n_threads, float* out);

/* every closest-hit query one sample of pixel (x, y) issues, in order, with its result (debugging probe) */
int orc_trace_path(const OrcScene* s, const LrRenderParams* p, int x, int y, int sample, int traversal, int max_rays,
                   float* origins, float* directions, int32_t* prim, float* t, int32_t* n_rays);

int orc_trace_primary(const OrcScene* s, float u, float v, float ua, float va, int traversal,
                      int n_threads, int32_t* prim, float* t);
int orc_trace_rays(const OrcScene* s, int64_t n, const float* origins, const float* directions,
                   int traversal, int brute_force, int32_t* prim, float* t, float* normal);

/* one camera sample with explicit random numbers (u,v sensor; ua,va aperture):
 * out = origin(3) direction(3) pdf g_term sensitivity                                     */
int orc_camera_sample(const LrCamera* cam, int x, int y, float u, float v, float ua, float va, float* out9);

/* ---- unit-level restatements used by the known-answer tests ---- */
int orc_triangle_intersect_mt(const float* p9, const float* o, const float* d, float* t, float* pos, float* n);
int orc_triangle_intersect_3c(const float* p9, const float* o, const float* d, float* t, float* pos, float* n);
int orc_sphere_intersect(const float* center, float radius, const float* o, const float* d, float* t, float* pos, float* n);
int orc_aabb_is_intersect(const float* lo, const float* hi, const float* o, const float* d);
void orc_reflect(const float* v, const float* n, float* out);
int orc_refract(const float* v, const float* n, float from_per_to_ior, float* out);
void orc_orthonormal_basis(const float* n, float* tangent, float* binormal);
float orc_checker(float u, float v);
/* material eval with explicit random numbers: brdf(3), sample dir(3), pdf, weight */
int orc_material_brdf(const LrMaterial* m, const float* out_, const float* in_, const float* n, const float* pos, float* brdf3);
int orc_material_sample(const LrMaterial* m, const float* out_, const float* n, float r1, float r2, float* in3, float* pdf);
float orc_material_weight(const LrMaterial* m);
void orc_material_coef(const LrMaterial* m, const float* out_, const float* n, float dist, float* coef3);
float orc_fresnel(float from_ior, float to_ior, const float* out_, const float* in_, const float* on);
void orc_ior_pair(const LrMaterial* m, const float* out_, const float* n, float* from_ior, float* to_ior);
void orc_sky_radiance(const LrSky* sky, const float* d, float* rgb);
float orc_rng_float(uint64_t seed, uint32_t pixel, uint32_t sample, int index);

/* matrix / camera constructors (matrix4.rs:9-77,185-223; camera.rs:34-62,149-166,224-264,366-409) */
void orc_matrix_unit(float* m);
void orc_matrix_translate(const float* v, float* m);
void orc_matrix_scale(const float* v, float* m);
void orc_matrix_axis_angle(const float* axis, float angle_deg, float* m);
void orc_matrix_look_at(const float* origin, const float* target, const float* up, float* m);
void orc_matrix_mul(const float* a, const float* b, float* out);
void orc_matrix_apply(const float* m, const float* v3, float* out3);
void orc_camera_ideal_pinhole(const float* matrix, float xfov, int w, int h, LrCamera* out);
void orc_camera_thin_lens(const float* matrix, float xfov, float focus, float fnum, int w, int h, LrCamera* out);
void orc_camera_omnidirectional(const float* matrix, int w, int h, LrCamera* out);
void orc_camera_pinhole(const float* position, const float* aperture_position, const float* sensor_size,
                        int w, int h, float aperture_radius, LrCamera* out);
float orc_triangle_area(const float* p9);
float orc_sphere_area(float r);

#ifdef __cplusplus
}
#endif
#endif
