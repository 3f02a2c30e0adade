// device_scene.h — layout of the scene in HBM and the kernel parameter blocks.
//
// Everything the kernels read is a flat, 16-byte aligned array so that every fetch on the
// traversal path is a 128-bit load:
//   nodes    : 4 x float4 per BVH node   (64 B)   — two child boxes + child codes
//   tris     : 3 x float4 per triangle   (48 B)   — p0|prim_id, e1 = p1-p0|material, e2 = p2-p0|0   (BVH leaf order,
//                                                   then the flat list of large triangles every ray tests)
//   tri_box  : 2 x float4 per triangle   (32 B)   — box of the vertices (lo | same-box-as-predecessor flag, hi | 0): the reference's
//                                                   leaf-AABB gate — once per ray for the nearest tree hit, once per flat box
//   tri_n    : 1 x float4 per triangle   (16 B)   — unit geometric normal | 0 (read once per path vertex)
//   spheres  : 1 x float4 per sphere     (16 B)   — centre|radius     (+ int2 material/prim_id)
//   mats     : 3 x float4 per material   (48 B)   — colour|type, emission|param0, param1|weight|emissive|0
//   emitters : 3 x float4 per emitter    (48 B)   — tri: p0|kind, p1|area, p2|0 ; sphere: centre|kind, radius,0,0|area
//   sky      : 1 x float4 per texel (rgb|0), nearest-texel lookup
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/lumilly.h"

namespace lr {

constexpr int kStackDepth = 64;          // per-thread traversal stack entries
constexpr int kBlockThreads = 128;       // 4 warps, each warp owns an 8x4 pixel tile

struct DevScene {
  const float4* nodes;
  const float4* tris;
  const float4* tri_box;                 // 2 x float4 per triangle: box lo | hi of the vertices (the reference's leaf-AABB gate)
  const float4* tri_n;                   // geometric normal per triangle, precomputed like Triangle::new does (triangle.rs:36)
  const float4* spheres;
  const int2* sphere_meta;
  const float4* mats;
  const float* emitter_cdf;
  const float4* emitters;
  const float4* sky_pixels;
  int n_nodes, n_tris, n_spheres, n_emitters;
  int n_bvh_tris;                        // tris[0, n_bvh_tris) are in BVH leaf order; tris[n_bvh_tris, n_tris) is the flat list
  float bvh_lo[3], bvh_hi[3];            // bounds of the tree (union of the root's child boxes, padded like them)
  float emission_area;
  int sky_type;
  float sky_color[3];
  int sky_height;
  float sky_longitude_offset;
  LrCamera cam;
};

struct DevParams {
  int integrator;
  int spp_begin, spp_count;
  int depth, depth_limit, no_direct_emitter;
  unsigned long long seed;
  int crop_x, crop_y, crop_w, crop_h;
  int splits;
  int tiles_x, tiles_y;                  // 8x4 pixel tiles over the crop window
  int defer_iters, defer_thresh;         // path vertices per phase A; suspended lanes that end it early (persistent.cuh)
  int organisation;                      // 0 = the build's measured best, 1 = one path per lane (persistent.cuh), 2 = pool (pool.cuh); BVH scenes only
};

// device counter block (unsigned long long each)
enum CounterSlot { C_RAYS = 0, C_NONFINITE = 1, C_NODES = 2, C_TRIS = 3, C_SPHERES = 4, C_RETRACE = 5, C_FLAT_TRIS = 6, C_NEXT_UNIT = 7, C_FLAT_BOXES = 8, C_COUNT = 9 };

}  // namespace lr
