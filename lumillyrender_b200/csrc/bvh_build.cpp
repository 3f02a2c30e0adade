// bvh_build.cpp — host SAH BVH build + flattening to the 64-byte two-child node the kernels read.
//
// Replaces BVH::new / BVH::construct (src/bvh.rs:57-127).  The reference's full-sweep SAH with one
// primitive per leaf is O(n log^2 n) and produces 2N-1 boxed nodes; the nearest hit does not depend
// on the topology (device_path.cuh), so this build is free to differ: binned SAH (16 bins x 3 axes,
// same cost model T_aabb = 1, T_tri = 2 as bvh.rs:71-72), leaves of up to 2 triangles (8 at most),
// child boxes stored in the parent and padded outward so the device's node test is conservative,
// nodes emitted in depth-first order (a node's near child is usually the next node in memory).  The top levels are
// forked over the host's cores (Builder::build_forked): same arrays as the sequential build, 1 M triangles in 1.1 s
// instead of 4.2 s on 8 cores.
//
// Large triangles stay OUTSIDE the tree ("flat list", tested by every ray like the spheres): a wall or floor
// whose box spans the scene is reached by every ray anyway, and in a tree it only forces all rays through a
// few near-root leaves in ray-dependent order — divergent work on a GPU — while a fixed flat loop runs in
// lock-step.  With the walls out, the tree bounds only the meshes, so most rays of a path tracer in a room
// never enter it.  The nearest hit is the minimum over all candidates either way.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <system_error>
#include <thread>
#include <vector>

#include "common.h"
#include "host_scene.h"

namespace lr {

namespace {

struct Box {
  float lo[3], hi[3];
  void reset() { for (int i = 0; i < 3; i++) { lo[i] = INFINITY; hi[i] = -INFINITY; } }
  void grow(const float* p) { for (int i = 0; i < 3; i++) { lo[i] = std::fmin(lo[i], p[i]); hi[i] = std::fmax(hi[i], p[i]); } }
  void grow(const Box& b) { for (int i = 0; i < 3; i++) { lo[i] = std::fmin(lo[i], b.lo[i]); hi[i] = std::fmax(hi[i], b.hi[i]); } }
  float area() const {
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return 2.0f * (dx * dy + dy * dz + dz * dx);
  }
};

constexpr int kBins = 16;
static int kLeafTarget = 2;        // stop splitting at <= 2 triangles (measured best of 1..8, profiles/r01_c_ab_s17.txt;
                                   // LR_LEAF_TARGET: development knob)
constexpr int kDeviceStackDepth = 64;   // device_scene.h: kStackDepth
constexpr int kLeafMax = 8;        // leaf code holds count-1 in 3 bits
constexpr int kSahDepthLimit = 32; // deeper than this: median splits only, so depth <= 32 + log2(n) < 64
constexpr int kFlatAllBelow = 16;  // scenes with this many triangles or fewer: no tree at all
constexpr int kFlatMax = 24;       // at most this many large triangles are kept outside the tree
constexpr float kFlatAreaFraction = 0.01f;

struct Builder {
  const std::vector<Box>& tri_box;
  const std::vector<Vec3>& centroid;
  std::vector<int>& order;           // permutation being partitioned in place
  std::vector<LrBvhNode> nodes;
  int max_depth = 0;
  float pad = 0.0f;
  int leaf_target = kLeafTarget;     // items per leaf (1 when the items are subtrees: rebuild_top_sah)
  const std::vector<int>* weight = nullptr;   // triangles an item stands for in the SAH cost (null: 1 each)

  Box bounds(int begin, int end) const {
    Box b; b.reset();
    for (int i = begin; i < end; i++) b.grow(tri_box[order[i]]);
    return b;
  }

  static int leaf_code(int first, int count) { return ~((first << 3) | (count - 1)); }

  void store_child(LrBvhNode& n, int slot, const Box& b, int code, int count) {
    for (int i = 0; i < 3; i++) {
      n.f[slot * 6 + i] = b.lo[i] - pad;
      n.f[slot * 6 + 3 + i] = b.hi[i] + pad;
    }
    n.c[slot] = code;
    n.n[slot] = count;
  }

  // picks a partition of [begin,end); returns mid (begin < mid < end)
  int partition(int begin, int end, const Box& node_box, int depth) {
    const int n = end - begin;
    Box cb; cb.reset();
    for (int i = begin; i < end; i++) cb.grow(centroid[order[i]].v);
    int best_axis = -1, best_bin = -1;
    float best_cost = INFINITY;
    if (depth < kSahDepthLimit) {
      const float parent_area = node_box.area();
      for (int axis = 0; axis < 3; axis++) {
        const float ext = cb.hi[axis] - cb.lo[axis];
        if (!(ext > 0.0f)) continue;
        Box bin_box[kBins]; int bin_n[kBins];
        for (int b = 0; b < kBins; b++) { bin_box[b].reset(); bin_n[b] = 0; }
        const float scale = (float)kBins / ext;
        for (int i = begin; i < end; i++) {
          int b = (int)((centroid[order[i]][axis] - cb.lo[axis]) * scale);
          b = std::min(std::max(b, 0), kBins - 1);
          bin_box[b].grow(tri_box[order[i]]); bin_n[b] += weight ? (*weight)[order[i]] : 1;
        }
        float right_area[kBins]; int right_n[kBins];
        Box acc; acc.reset(); int cnt = 0;
        for (int b = kBins - 1; b > 0; b--) { acc.grow(bin_box[b]); cnt += bin_n[b]; right_area[b] = acc.area(); right_n[b] = cnt; }
        acc.reset(); cnt = 0;
        for (int b = 0; b + 1 < kBins; b++) {
          acc.grow(bin_box[b]); cnt += bin_n[b];
          if (cnt == 0 || right_n[b + 1] == 0) continue;
          // bvh.rs:107: T = 2*T_aabb + (A(S1)*N(S1) + A(S2)*N(S2)) * T_tri / A(S)
          const float cost = 2.0f * 1.0f + (acc.area() * (float)cnt + right_area[b + 1] * (float)right_n[b + 1]) * 2.0f / parent_area;
          if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = b; }
        }
      }
    }
    if (best_axis >= 0) {
      const float ext = cb.hi[best_axis] - cb.lo[best_axis];
      const float scale = (float)kBins / ext;
      const float lo = cb.lo[best_axis];
      const int axis = best_axis, bin = best_bin;
      auto it = std::partition(order.begin() + begin, order.begin() + end, [&](int t) {
        int b = (int)((centroid[t][axis] - lo) * scale);
        b = std::min(std::max(b, 0), kBins - 1);
        return b <= bin;
      });
      const int mid = (int)(it - order.begin());
      if (mid > begin && mid < end) return mid;
    }
    // median split along the widest centroid axis (also the bounded-depth fallback)
    int axis = 0;
    for (int a = 1; a < 3; a++) if (cb.hi[a] - cb.lo[a] > cb.hi[axis] - cb.lo[axis]) axis = a;
    const int mid = begin + n / 2;
    std::nth_element(order.begin() + begin, order.begin() + mid, order.begin() + end,
                     [&](int a, int b) { return centroid[a][axis] < centroid[b][axis]; });
    return mid;
  }

  // builds the inner node covering [begin,end) (n >= 2) and returns its index
  int build_inner(int begin, int end, const Box& box, int depth) {
    max_depth = std::max(max_depth, depth + 1);
    const int me = (int)nodes.size();
    nodes.push_back(LrBvhNode{});
    const int mid = partition(begin, end, box, depth);
    const int range[2][2] = {{begin, mid}, {mid, end}};
    for (int slot = 0; slot < 2; slot++) {
      const int b = range[slot][0], e = range[slot][1], cnt = e - b;
      const Box cbx = bounds(b, e);
      bool leaf = cnt <= leaf_target;
      if (!leaf && cnt <= kLeafMax && depth + 1 >= kStackGuardDepth) leaf = true;
      if (leaf) {
        LrBvhNode tmp = nodes[me]; store_child(tmp, slot, cbx, leaf_code(b, cnt), cnt); nodes[me] = tmp;
      } else {
        const int child = build_inner(b, e, cbx, depth + 1);
        LrBvhNode tmp = nodes[me]; store_child(tmp, slot, cbx, child, 0); nodes[me] = tmp;
      }
    }
    return me;
  }

  static constexpr int kStackGuardDepth = 60;
  static constexpr int kParallelGrain = 16384;   // ranges smaller than this are not forked

  // Fork-join over the top levels of the tree.  The two halves of a range are independent — partition() works in place
  // on disjoint parts of `order`, everything else is read-only — so one half is built by another thread into its own
  // node array (root at index 0) and the arrays are concatenated in depth-first order with the inner child indices
  // shifted: byte for byte the array the sequential build_inner emits (tests/test_host_frontend.py compares them).
  std::vector<LrBvhNode> build_forked(int begin, int end, const Box& box, int depth, int forks, int& depth_max) const {
    if (forks <= 0 || end - begin < kParallelGrain) {
      Builder local{tri_box, centroid, order, {}, 0, pad, leaf_target, weight};
      local.nodes.reserve((size_t)(end - begin) / 2 + 16);
      local.build_inner(begin, end, box, depth);
      depth_max = local.max_depth;
      return std::move(local.nodes);
    }
    depth_max = depth + 1;
    Builder self{tri_box, centroid, order, {}, 0, pad, leaf_target, weight};            // partition / bounds / store_child on the shared arrays
    const int mid = self.partition(begin, end, box, depth);
    const int range[2][2] = {{begin, mid}, {mid, end}};
    Box cbx[2];
    bool leaf[2];
    std::vector<LrBvhNode> sub[2];
    int sub_depth[2] = {0, 0};
    for (int slot = 0; slot < 2; slot++) {
      const int cnt = range[slot][1] - range[slot][0];
      cbx[slot] = self.bounds(range[slot][0], range[slot][1]);
      leaf[slot] = cnt <= leaf_target || (cnt <= kLeafMax && depth + 1 >= kStackGuardDepth);
    }
    auto run = [&](int slot) { sub[slot] = build_forked(range[slot][0], range[slot][1], cbx[slot], depth + 1, forks - 1, sub_depth[slot]); };
    if (!leaf[0] && !leaf[1]) {
      std::thread other;
      bool forked = true;
      try { other = std::thread(run, 0); } catch (const std::system_error&) { forked = false; }   // no thread to be had: build both here
      run(1);
      if (forked) other.join(); else run(0);
    } else {
      for (int slot = 0; slot < 2; slot++) if (!leaf[slot]) run(slot);
    }
    std::vector<LrBvhNode> out;
    out.reserve(1 + sub[0].size() + sub[1].size());
    out.push_back(LrBvhNode{});
    for (int slot = 0; slot < 2; slot++) {
      const int b = range[slot][0], cnt = range[slot][1] - b;
      if (leaf[slot]) {
        self.store_child(out[0], slot, cbx[slot], leaf_code(b, cnt), cnt);
      } else {
        const int offset = (int)out.size();
        self.store_child(out[0], slot, cbx[slot], offset, 0);
        for (LrBvhNode nd : sub[slot]) {
          for (int k = 0; k < 2; k++) if (nd.c[k] >= 0) nd.c[k] += offset;
          out.push_back(nd);
        }
        depth_max = std::max(depth_max, sub_depth[slot]);
      }
    }
    return out;
  }
};

}  // namespace

// Outliers leave the tree.  A small primitive far from the rest — the area light of scenes/welcome-2018.toml hangs 2000 units
// above the mesh — makes the tree's bounds span the scene, so that nearly every ray (and every shadow ray, which is AIMED at
// the light) passes the reach test, visits the root and leaves again: one-step traversals that fill the BVH phase and idle
// its lanes (ncu: 2.8 of 32 lanes in the node loop of welcome-2018).  Both builders put such an outlier where SAH / Morton
// order puts it: in a leaf directly under the root.  While the root has a LEAF child whose removal at least halves the surface
// area of the tree's bounds, and the flat list has room, the leaf's triangles join the flat list (which every ray gates with
// one box test, device_path.cuh: flat_hits) and the other child becomes the root.  O(n) per peel; the nearest hit does not
// depend on which list a triangle is in.
static void peel_outliers(std::vector<LrTriangle>& tris, std::vector<LrBvhNode>& nodes, int& depth, int& n_flat) {
  auto box_area = [](const float* f) {
    const float dx = f[3] - f[0], dy = f[4] - f[1], dz = f[5] - f[2];
    return 2.0f * (dx * dy + dy * dz + dz * dx);
  };
  while (nodes.size() >= 2) {
    const LrBvhNode root = nodes[0];
    int leaf = -1;
    for (int k = 0; k < 2; k++) if (root.c[k] < 0 && root.c[1 - k] >= 0) leaf = k;       // a leaf beside an inner node
    if (leaf < 0) break;
    const int code = ~root.c[leaf], first = code >> 3, count = (code & 7) + 1;
    if (n_flat + count > kFlatMax) break;
    if (root.c[1 - leaf] != 1) break;                         // not depth-first order (cannot happen with our builders): leave the tree alone
    float uni[6];
    for (int a = 0; a < 3; a++) { uni[a] = std::fmin(root.f[a], root.f[6 + a]); uni[3 + a] = std::fmax(root.f[3 + a], root.f[9 + a]); }
    if (!(box_area(root.f + 6 * (1 - leaf)) <= 0.5f * box_area(uni))) break;
    const int n_tree = (int)tris.size() - n_flat;
    // the leaf's triangles go to the end of the tree range, i.e. to the front of the flat tail
    std::rotate(tris.begin() + first, tris.begin() + first + count, tris.begin() + n_tree);
    n_flat += count;
    // the inner child (node 1 in depth-first order) becomes the root: drop node 0, shift indices and leaf ranges
    nodes.erase(nodes.begin());
    for (LrBvhNode& nd : nodes) {
      for (int k = 0; k < 2; k++) {
        if (nd.c[k] >= 0) nd.c[k] -= 1;
        else {
          const int c = ~nd.c[k], f0 = c >> 3, cnt = c & 7;
          if (f0 > first) nd.c[k] = ~(((f0 - count) << 3) | cnt);
        }
      }
    }
    depth -= 1;
  }
}

// The top of a device-built tree, rebuilt with SAH on the host (the hierarchical-LBVH idea).  A radix tree splits space at
// Morton-cell boundaries whatever the geometry; most of what that costs a traversal is decided in the upper levels, where every
// ray passes.  The subtrees of at most `cut` triangles stay as the device built them; the nodes above them — a few thousand —
// are replaced by a binned-SAH tree over the subtrees' boxes, written into the same node slots (a binary tree over M items has
// M - 1 inner nodes either way; the root stays node 0).  Node order is no longer depth-first, which nothing depends on after
// peel_outliers has run.  Returns the new depth of the tree.
static int rebuild_top_sah(std::vector<LrBvhNode>& nodes, int cut, int max_depth) {
  const int n_nodes = (int)nodes.size();
  // triangles below every node: children follow their parent in the builders' depth-first layout
  std::vector<int> cnt(n_nodes, 0), below(n_nodes, 1);       // below: depth of the subtree rooted at the node, in nodes
  for (int i = n_nodes - 1; i >= 0; i--)
    for (int k = 0; k < 2; k++) {
      const int c = nodes[i].c[k];
      cnt[i] += c >= 0 ? cnt[c] : nodes[i].n[k];
      if (c >= 0) below[i] = std::max(below[i], below[c] + 1);
    }
  const int old_depth = n_nodes > 0 ? below[0] : 0;
  if (n_nodes < 8 || cnt[0] <= cut) return old_depth;
  struct Item { float box[6]; int code; int count; int depth; };
  std::vector<Item> items;
  std::vector<int> slots;                                    // the nodes above the cut, root first
  {
    std::vector<int> todo{0};
    while (!todo.empty()) {
      const int i = todo.back();
      todo.pop_back();
      slots.push_back(i);
      for (int k = 1; k >= 0; k--) {
        const int c = nodes[i].c[k];
        if (c >= 0 && cnt[c] > cut) { todo.push_back(c); continue; }
        Item it;
        std::memcpy(it.box, nodes[i].f + 6 * k, sizeof(it.box));
        it.code = c; it.count = nodes[i].n[k]; it.depth = c >= 0 ? below[c] : 0;
        items.push_back(it);
      }
    }
  }
  const int m = (int)items.size();
  if (m < 4 || (int)slots.size() != m - 1) return old_depth;
  std::vector<Box> box(m);
  std::vector<Vec3> centroid(m);
  std::vector<int> order(m);
  Box all; all.reset();
  for (int i = 0; i < m; i++) {
    for (int a = 0; a < 3; a++) { box[i].lo[a] = items[i].box[a]; box[i].hi[a] = items[i].box[3 + a]; }
    centroid[i] = vec3(0.5f * (box[i].lo[0] + box[i].hi[0]), 0.5f * (box[i].lo[1] + box[i].hi[1]), 0.5f * (box[i].lo[2] + box[i].hi[2]));
    order[i] = i;
    all.grow(box[i]);
  }
  std::vector<int> weight(m);
  for (int i = 0; i < m; i++) weight[i] = items[i].code >= 0 ? cnt[items[i].code] : items[i].count;
  Builder top{box, centroid, order, {}, 0, 0.0f, 1, &weight};   // pad 0: the items' boxes are padded already; one item per leaf; SAH weighs an item by its triangles
  top.nodes.reserve(m);
  top.build_inner(0, m, all, 0);
  if ((int)top.nodes.size() != m - 1) return old_depth;
  // depth of the stitched tree, and a check that every leaf of the top tree is ONE item; nothing is written before both hold
  int new_depth = 0;
  {
    std::vector<std::pair<int, int>> todo{{0, 1}};
    while (!todo.empty()) {
      const std::pair<int, int> cur = todo.back();
      todo.pop_back();
      for (int k = 0; k < 2; k++) {
        const int c = top.nodes[cur.first].c[k];
        if (c >= 0) { todo.push_back({c, cur.second + 1}); continue; }
        if (((~c) & 7) != 0) return old_depth;               // a leaf of several items (only beyond depth 60): keep the old top
        new_depth = std::max(new_depth, cur.second + items[order[(~c) >> 3]].depth);
      }
    }
  }
  if (new_depth >= max_depth) return old_depth;
  for (int j = 0; j < m - 1; j++) {
    LrBvhNode nd = top.nodes[j];
    for (int k = 0; k < 2; k++) {
      if (nd.c[k] >= 0) { nd.c[k] = slots[nd.c[k]]; nd.n[k] = 0; }
      else {
        const Item& it = items[order[(~nd.c[k]) >> 3]];        // the subtree (or device leaf) the one-item leaf stands for
        nd.c[k] = it.code; nd.n[k] = it.count;
        std::memcpy(nd.f + 6 * k, it.box, sizeof(it.box));
      }
    }
    nodes[slots[j]] = nd;
  }
  return new_depth;
}

int build_bvh(std::vector<LrTriangle>& tris, std::vector<LrBvhNode>& nodes_out, int& depth_out, float& seconds_out, int& n_flat_out,
              float origin_extent, int builder, int* builder_used, float* device_kernel_ms) {
  if (builder_used) *builder_used = LR_BVH_HOST;
  if (device_kernel_ms) *device_kernel_ms = 0.0f;
  const auto t0 = std::chrono::steady_clock::now();
  if (const char* e = std::getenv("LR_LEAF_TARGET")) kLeafTarget = std::max(1, std::min(8, std::atoi(e)));
  nodes_out.clear();
  depth_out = 0;
  seconds_out = 0.0f;
  n_flat_out = 0;
  const int n_all = (int)tris.size();
  if (n_all == 0) return LR_OK;
  if (n_all >= (1 << 28)) return fail(LR_ERR_UNSUPPORTED, "more than 2^28 triangles");
  // ---- flat list selection: all triangles of a tiny scene, else the (at most kFlatMax) largest triangles whose
  // box area is >= kFlatAreaFraction of the scene's box area
  Box scene; scene.reset();
  {
    // box area of every triangle + bounds of the scene: independent per triangle, so large meshes are cut over a few host
    // threads (a million triangles: 95 -> 15 ms; this pass is most of the wall time of a DEVICE build, whose kernels take 1 ms)
    std::vector<float> area(n_all);
    const int hw = (int)std::thread::hardware_concurrency();
    const int workers = n_all >= 65536 ? std::max(1, std::min(8, hw)) : 1;
    std::vector<Box> part(workers);
    std::vector<char> bad(workers, 0);
    auto scan = [&](int w, int i0, int i1) {
      Box sc; sc.reset();
      bool finite = true;
      for (int i = i0; i < i1; i++) {
        const LrTriangle& t = tris[i];
        Box b;
        for (int k = 0; k < 3; k++) {
          b.lo[k] = std::fmin(std::fmin(t.p0[k], t.p1[k]), t.p2[k]);
          b.hi[k] = std::fmax(std::fmax(t.p0[k], t.p1[k]), t.p2[k]);
          finite = finite && std::isfinite(t.p0[k]) && std::isfinite(t.p1[k]) && std::isfinite(t.p2[k]);
        }
        area[i] = b.area();
        sc.grow(b);
      }
      part[w] = sc;
      bad[w] = finite ? 0 : 1;
    };
    {
      struct Joiner { std::vector<std::thread> v; ~Joiner() { for (std::thread& th : v) if (th.joinable()) th.join(); } } pool;
      pool.v.reserve(workers);
      const int per = (n_all + workers - 1) / workers;
      int started = 1;
      try {
        for (int w = 1; w < workers; w++) {
          pool.v.emplace_back(scan, w, std::min(n_all, w * per), std::min(n_all, (w + 1) * per));
          started = w + 1;
        }
      } catch (const std::system_error&) {}                  // no thread to be had: this thread scans the rest
      scan(0, 0, std::min(n_all, per));
      for (int w = started; w < workers; w++) scan(w, std::min(n_all, w * per), std::min(n_all, (w + 1) * per));
    }
    for (int w = 0; w < workers; w++) {
      if (bad[w]) return fail(LR_ERR_INVALID, "non-finite triangle vertex");
      scene.grow(part[w]);
    }
    if (std::getenv("LR_BVH_TRACE")) std::fprintf(stderr, "build_bvh area scan      %8.2f ms (%d threads)\n", std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(), workers);
    std::vector<int> flat;
    if (n_all <= kFlatAllBelow) {
      for (int i = 0; i < n_all; i++) flat.push_back(i);
    } else {
      const float thresh = kFlatAreaFraction * scene.area();
      for (int i = 0; i < n_all; i++) if (area[i] >= thresh && area[i] > 0.0f) flat.push_back(i);
      if ((int)flat.size() > kFlatMax) {
        std::stable_sort(flat.begin(), flat.end(), [&](int a, int b) { return area[a] > area[b]; });
        flat.resize(kFlatMax);
        std::sort(flat.begin(), flat.end());
      }
    }
    // (a rebuild finds the flat triangles already at the tail, in order: nothing to move)
    bool in_place = !flat.empty();
    for (size_t i = 0; i < flat.size() && in_place; i++) in_place = flat[i] == n_all - (int)flat.size() + (int)i;
    if (in_place) {
      n_flat_out = (int)flat.size();
    } else if (!flat.empty()) {
      std::vector<char> is_flat(n_all, 0);
      for (int i : flat) is_flat[i] = 1;
      std::vector<LrTriangle> re; re.reserve(n_all);
      for (int i = 0; i < n_all; i++) if (!is_flat[i]) re.push_back(tris[i]);
      for (int i : flat) re.push_back(tris[i]);            // instance order is kept inside the flat list
      tris.swap(re);
      n_flat_out = (int)flat.size();
    }
  }
  if (std::getenv("LR_BVH_TRACE")) std::fprintf(stderr, "build_bvh flat selection %8.2f ms\n", std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count());
  const int n = n_all - n_flat_out;                        // triangles that go into the tree: tris[0, n)
  if (n == 0) { seconds_out = std::chrono::duration<float>(std::chrono::steady_clock::now() - t0).count(); return LR_OK; }
  if (builder == LR_BVH_DEVICE && n >= 1024) {
    // the device builder: same flat list, same pad, same leaf target; a radix tree deeper than the traversal stack (only
    // with tens of thousands of coincident centroids) goes to the host builder below instead
    float extent = 0.0f;
    for (int k = 0; k < 3; k++) extent = std::fmax(extent, std::fmax(std::fabs(scene.lo[k]), std::fabs(scene.hi[k])));
    extent = std::fmax(extent, std::fmin(origin_extent, 1e30f));
    float kernel_ms = 0.0f, sec = 0.0f;
    if (int rc = build_bvh_device(tris, n, 4e-6f * extent + 1e-30f, kLeafTarget, Builder::kStackGuardDepth, nodes_out, depth_out, sec, kernel_ms)) return rc;
    if (depth_out < Builder::kStackGuardDepth) {
      peel_outliers(tris, nodes_out, depth_out, n_flat_out);
      int cut = 512;                                           // LR_BVH_TOP_CUT: development knob (0 = keep the radix tree's top)
      if (const char* e = std::getenv("LR_BVH_TOP_CUT")) cut = std::atoi(e);
      if (cut > 0) depth_out = rebuild_top_sah(nodes_out, cut, Builder::kStackGuardDepth);
      if (builder_used) *builder_used = LR_BVH_DEVICE;
      if (device_kernel_ms) *device_kernel_ms = kernel_ms;
      seconds_out = std::chrono::duration<float>(std::chrono::steady_clock::now() - t0).count();
      return LR_OK;
    }
    nodes_out.clear();
    depth_out = 0;
  }
  std::vector<Box> tri_box(n);
  std::vector<Vec3> centroid(n);
  std::vector<int> order(n);
  Box all; all.reset();
  for (int i = 0; i < n; i++) {
    Box b; b.reset();
    b.grow(tris[i].p0); b.grow(tris[i].p1); b.grow(tris[i].p2);
    for (int k = 0; k < 3; k++)
      if (!std::isfinite(b.lo[k]) || !std::isfinite(b.hi[k])) return fail(LR_ERR_INVALID, "non-finite triangle vertex");
    tri_box[i] = b;
    centroid[i] = vec3(0.5f * (b.lo[0] + b.hi[0]), 0.5f * (b.lo[1] + b.hi[1]), 0.5f * (b.lo[2] + b.hi[2]));
    order[i] = i;
    all.grow(b);
  }
  Builder bld{tri_box, centroid, order, {}, 0, 0.0f};
  float extent = 0.0f;
  for (int k = 0; k < 3; k++) extent = std::fmax(extent, std::fmax(std::fabs(scene.lo[k]), std::fabs(scene.hi[k])));   // whole scene: ray origins lie on any surface
  // ... and at the camera and on the spheres (origin_extent, from the caller: max |coordinate| of the aperture / sensor
  // positions and of the spheres' boxes, a sphere's radius clamped to 16 x the triangles' extent — the far side of a
  // ground sphere of radius 1e5 is not a place rays reach the mesh from)
  extent = std::fmax(extent, std::fmin(origin_extent, 1e30f));
  bld.pad = 4e-6f * extent + 1e-30f;     // > the rounding error of (box - origin) * inv, see device_path.cuh
  bld.nodes.reserve((size_t)n / 2 + 16);
  if (n == 1) {
    // a single leaf: both children reference it (testing a triangle twice cannot change the minimum)
    LrBvhNode root{};
    bld.store_child(root, 0, all, Builder::leaf_code(0, 1), 1);
    bld.store_child(root, 1, all, Builder::leaf_code(0, 1), 1);
    bld.nodes.push_back(root);
    bld.max_depth = 1;
  } else {
    // threads: LR_BVH_THREADS (1 = the sequential build), default the host's cores; 2^forks subtrees are built concurrently
    int threads = (int)std::thread::hardware_concurrency();
    if (const char* e = std::getenv("LR_BVH_THREADS")) threads = std::atoi(e);
    int forks = 0;
    while ((1 << forks) < std::max(1, std::min(threads, 64))) forks++;
    if (forks == 0 || n < 2 * Builder::kParallelGrain) {
      bld.build_inner(0, n, all, 0);
    } else {
      bld.nodes = bld.build_forked(0, n, all, 0, forks + 1, bld.max_depth);      // one extra level: better balance
    }
  }
  std::vector<LrTriangle> permuted(tris);                  // the flat tail stays where it is
  for (int i = 0; i < n; i++) permuted[i] = tris[order[i]];
  tris.swap(permuted);
  nodes_out.swap(bld.nodes);
  depth_out = bld.max_depth;
  peel_outliers(tris, nodes_out, depth_out, n_flat_out);
  // test hook (tests/test_host_frontend.py): the stitching code of the device path, run over a host-built tree
  if (const char* e = std::getenv("LR_BVH_TOP_CUT_HOST")) if (std::atoi(e) > 0) depth_out = rebuild_top_sah(nodes_out, std::atoi(e), Builder::kStackGuardDepth);
  seconds_out = std::chrono::duration<float>(std::chrono::steady_clock::now() - t0).count();
  return LR_OK;
}

// Structural validation of a description handed over the C ABI (the reference panics instead).
int validate_desc(const LrSceneDesc& d) {
  if (d.n_materials < 0 || d.n_triangles < 0 || d.n_spheres < 0 || d.n_nodes < 0) return fail(LR_ERR_INVALID, "negative count");
  if ((d.n_materials > 0 && !d.materials) || (d.n_triangles > 0 && !d.triangles) || (d.n_spheres > 0 && !d.spheres) || (d.n_nodes > 0 && !d.nodes))
    return fail(LR_ERR_INVALID, "null array with non-zero count");
  if (d.n_flat_triangles < 0 || d.n_flat_triangles > d.n_triangles) return fail(LR_ERR_INVALID, "n_flat_triangles out of range");
  if (d.n_flat_triangles > 32) return fail(LR_ERR_UNSUPPORTED, "more than 32 triangles outside the BVH (the kernels keep a lane's flat-list candidates in a 32-bit mask and the list in 2.5 KB of shared memory)");
  const int n_bvh_tris = d.n_triangles - d.n_flat_triangles;
  if ((n_bvh_tris > 0) != (d.n_nodes > 0)) return fail(LR_ERR_INVALID, "BVH nodes must be present iff triangles are in the BVH (use lr_host_scene_from_arrays)");

  const int n_prims = d.n_triangles + d.n_spheres;
  for (int i = 0; i < d.n_materials; i++)
    if (d.materials[i].type < LR_MAT_LAMBERT || d.materials[i].type > LR_MAT_IDEAL_REFRACTION) return fail(LR_ERR_INVALID, "unknown material type");
  for (int i = 0; i < d.n_triangles; i++) {
    const LrTriangle& t = d.triangles[i];
    if (t.material < 0 || t.material >= d.n_materials) return fail(LR_ERR_INVALID, "triangle material index out of range");
    if (t.prim_id < 0 || t.prim_id >= n_prims) return fail(LR_ERR_INVALID, "triangle prim_id out of range");
  }
  for (int i = 0; i < d.n_spheres; i++) {
    const LrSphere& s = d.spheres[i];
    if (s.material < 0 || s.material >= d.n_materials) return fail(LR_ERR_INVALID, "sphere material index out of range");
    if (s.prim_id < 0 || s.prim_id >= n_prims) return fail(LR_ERR_INVALID, "sphere prim_id out of range");
  }
  for (int i = 0; i < d.n_nodes; i++) {
    for (int k = 0; k < 2; k++) {
      const int c = d.nodes[i].c[k];
      if (c >= 0) { if (c >= d.n_nodes) return fail(LR_ERR_INVALID, "BVH child index out of range"); }
      else {
        const int code = ~c, first = code >> 3, count = (code & 7) + 1;
        if (first < 0 || first + count > n_bvh_tris) return fail(LR_ERR_INVALID, "BVH leaf range out of bounds");
      }
    }
  }
  // The device traversal has a fixed stack and no cycle check (device_path.cuh), and the node array comes over a public
  // ABI: walk the tree from node 0 and demand that it IS a tree — every inner node reached exactly once — whose real
  // depth fits the stack.  The caller's bvh_depth is not trusted.
  if (d.n_nodes > 0) {
    std::vector<int> depth_of(d.n_nodes, 0);                 // 0 = not reached yet
    std::vector<int> todo;
    todo.push_back(0);
    depth_of[0] = 1;
    int reached = 0, max_depth = 0;
    while (!todo.empty()) {
      const int i = todo.back();
      todo.pop_back();
      reached++;
      max_depth = std::max(max_depth, depth_of[i]);
      for (int k = 0; k < 2; k++) {
        const int c = d.nodes[i].c[k];
        if (c < 0) continue;
        if (depth_of[c] != 0) return fail(LR_ERR_INVALID, "BVH node reached twice (cycle or shared subtree)");
        depth_of[c] = depth_of[i] + 1;
        todo.push_back(c);
      }
    }
    if (reached != d.n_nodes) return fail(LR_ERR_INVALID, "BVH has nodes that cannot be reached from the root");
    if (max_depth >= kDeviceStackDepth) return fail(LR_ERR_UNSUPPORTED, "BVH deeper than the device traversal stack (64)");
  }
  const LrCamera& c = d.camera;
  if (c.type < LR_CAM_IDEAL_PINHOLE || c.type > LR_CAM_OMNIDIRECTIONAL) return fail(LR_ERR_INVALID, "unknown camera type");
  if (c.width <= 0 || c.height <= 0) return fail(LR_ERR_INVALID, "camera resolution must be positive");
  if ((int64_t)c.width * c.height >= (1LL << 31)) return fail(LR_ERR_UNSUPPORTED, "film larger than 2^31 pixels");
  if (d.sky.type == LR_SKY_IBL) {
    if (!d.sky.pixels || d.sky.height <= 0) return fail(LR_ERR_INVALID, "IBL sky without pixels");
    if (d.sky.n_pixels < 2LL * d.sky.height * d.sky.height) return fail(LR_ERR_INVALID, "IBL sky: n_pixels < 2*height*height (sky.rs:64-72 indexes a 2H x H image)");
  } else if (d.sky.type != LR_SKY_UNIFORM) return fail(LR_ERR_INVALID, "unknown sky type");
  return LR_OK;
}

}  // namespace lr
