// host_scene.h — host front end: TOML scene files -> flat arrays (LrSceneDesc).
// C++ restatement of scene_loader.rs (schema) + description.rs (assembly); runs once, outside the hot path.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "common.h"

namespace lr {

// ---- minimal TOML document model (the subset scenes/*.toml use) ----
struct TomlValue {
  enum Kind { NIL, BOOL, INT, FLOAT, STRING, ARRAY, TABLE } kind = NIL;
  bool b = false;
  long long i = 0;
  double f = 0.0;
  std::string s;
  std::vector<TomlValue> arr;
  std::vector<std::pair<std::string, TomlValue>> tab;   // insertion-ordered

  const TomlValue* get(const std::string& key) const;
  TomlValue* get(const std::string& key);
  TomlValue& insert(const std::string& key);
  bool is_number() const { return kind == INT || kind == FLOAT; }
  double number() const { return kind == INT ? (double)i : f; }
};
int toml_parse(const std::string& text, TomlValue& root, std::string& err);

// ---- Wavefront OBJ/MTL (what tobj::load_obj returns, description.rs:150-162) ----
struct ObjModel {
  std::string name;
  int material_id = -1;                 // index into ObjFile::materials, -1 = none
  std::vector<float> positions;         // xyz per corner, already de-indexed (3 corners per face)
};
struct ObjMaterial { std::string name; float diffuse[3] = {0.0f, 0.0f, 0.0f}; };
struct ObjFile { std::vector<ObjModel> models; std::vector<ObjMaterial> materials; };
int load_obj(const std::string& path, ObjFile& out);

// ---- 4x4 matrices with the reference's conventions (matrix4.rs) ----
struct Mat4 { float m[16]; };
Mat4 mat4_unit();
Mat4 mat4_translate(Vec3 v);
Mat4 mat4_scale(Vec3 v);
Mat4 mat4_axis_angle(Vec3 axis, float radians);
Mat4 mat4_look_at(Vec3 origin, Vec3 target, Vec3 up);
Mat4 mat4_mul(const Mat4& a, const Mat4& b);
Vec3 mat4_apply(const Mat4& m, Vec3 p);

void camera_ideal_pinhole(const Mat4& m, float xfov, int w, int h, LrCamera& out);
void camera_thin_lens(const Mat4& m, float xfov, float focus_distance, float f_number, int w, int h, LrCamera& out);
void camera_omnidirectional(const Mat4& m, int w, int h, LrCamera& out);
void camera_pinhole(Vec3 position, Vec3 aperture_position, const float* sensor_size, int w, int h, float aperture_radius, LrCamera& out);

// permutes `tris`: BVH leaf order first, then the `n_flat_out` large triangles kept outside the BVH
// builder: LR_BVH_HOST = binned SAH on the host's cores (bvh_build.cpp), LR_BVH_DEVICE = Morton-order radix tree on the GPU
// (bvh_build_gpu.cu; needs a CUDA device; falls back to the host builder for meshes of fewer than 1024 triangles or if the
// radix tree comes out deeper than the device traversal stack)
int build_bvh(std::vector<LrTriangle>& tris, std::vector<LrBvhNode>& nodes_out, int& depth_out, float& seconds_out, int& n_flat_out,
              float origin_extent = 0.0f, int builder = 0, int* builder_used = nullptr, float* device_kernel_ms = nullptr);
int build_bvh_device(std::vector<LrTriangle>& tris, int n_tree, float pad, int leaf_target, int max_depth, std::vector<LrBvhNode>& nodes_out,
                     int& depth_out, float& seconds_out, float& kernel_ms_out);

int load_hdr_file(const std::string& path, std::vector<float>& rgb, int& w, int& h);

}  // namespace lr

// The opaque handle of the C ABI: owns every array the LrSceneDesc points into.
struct LrHostScene {
  std::vector<LrMaterial> materials;
  std::vector<LrTriangle> triangles;
  std::vector<LrSphere> spheres;
  std::vector<LrBvhNode> nodes;
  std::vector<float> sky_pixels;
  LrSceneDesc desc{};
  LrSceneConfig config{};
  int bvh_builder = 0;        // LrBvhBuilder asked for (lr_host_scene_load: LR_BVH_BUILDER=device|host; lr_host_scene_rebuild_bvh)
  int bvh_builder_used = 0;   // ... and the one that built the tree
  float bvh_device_kernel_ms = 0.0f;
  int finalize();    // builds the BVH, counts emitters, fills desc
};
