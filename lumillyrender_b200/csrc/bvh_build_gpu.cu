// bvh_build_gpu.cu — BVH construction on the device (sm_100a): replaces BVH::new / BVH::construct (src/bvh.rs:57-127) for
// large meshes, where the host's binned-SAH build (bvh_build.cpp) takes seconds (1 M triangles: 2.5 s) and a file-to-image
// run at low spp is dominated by it.
//
// The nearest hit does not depend on the topology of the tree (device_path.cuh: a hit is decided by the reference's primitive
// test and the reference's slab test on the primitive's OWN box; node boxes are only a conservative cull), so the device is
// free to build a different tree than the host and than the reference.  What is built here:
//   1. per-triangle box + centroid, scene bounds of the centroids (warp-reduced, then atomicMin / atomicMax on
//      order-preserving integer images of the floats);
//   2. 30-bit Morton code of the centroid, ties broken by the triangle index -> unique 62-bit keys;
//   3. a stable LSD radix sort of (code, index) pairs, 4 passes of 8 bits, hand-written (histogram per block -> one scan
//      over (digit, block) -> stable scatter ranked with __match_any_sync);
//   4. the binary radix tree of Karras 2012 over the sorted keys (every inner node finds its key range and split from
//      common-prefix lengths, fully parallel);
//   5. one bottom-up pass (a thread per leaf climbs; the second arrival at a node merges): exact min / max boxes, triangle
//      counts and the number of nodes each subtree will EMIT — subtrees of <= 2 triangles collapse into leaves, the
//      2-triangle leaves the render kernels were tuned for (bvh_build.cpp: kLeafTarget);
//   6. layout without a top-down sweep: a node's position in depth-first order is its emitted depth plus the emitted sizes
//      of the left siblings along its path to the root, and likewise the first triangle of its leaf range — every node walks
//      up its <= 62 ancestors independently;
//   7. emission of the 64-byte two-child node the kernels read (child boxes padded outward like the host does) and of the
//      triangle array permuted into leaf order.
// All kernels are HBM/L2-bound integer and min/max work; grids are sized in multiples of the SM count where the work is large.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

#include "common.h"
#include "host_scene.h"

namespace lr {

namespace {

constexpr int kThreads = 256;
constexpr unsigned kFullMask = 0xffffffffu;

struct Box6 { float lo[3], hi[3]; };

__device__ __forceinline__ int float_order(float f) {            // monotone float -> int image (for atomicMin / atomicMax)
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float order_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__device__ __forceinline__ unsigned expand10(unsigned v) {       // 10 bits -> every third bit
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}

// ---- 1. boxes, centroids, centroid bounds
__global__ void __launch_bounds__(kThreads) tri_bounds_kernel(const LrTriangle* __restrict__ tris, int n, Box6* __restrict__ box, int* __restrict__ bounds) {
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const LrTriangle t = tris[i];
    Box6 b;
    for (int k = 0; k < 3; k++) {
      b.lo[k] = fminf(fminf(t.p0[k], t.p1[k]), t.p2[k]);         // triangle.rs:102-119: min / max of the vertices
      b.hi[k] = fmaxf(fmaxf(t.p0[k], t.p1[k]), t.p2[k]);
      const float c = 0.5f * (b.lo[k] + b.hi[k]);
      lo[k] = fminf(lo[k], c); hi[k] = fmaxf(hi[k], c);
    }
    box[i] = b;
  }
  for (int k = 0; k < 3; k++) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(kFullMask, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(kFullMask, hi[k], o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(bounds + k, float_order(lo[k]));
      atomicMax(bounds + 3 + k, float_order(hi[k]));
    }
  }
}

// ---- 2. Morton codes
__global__ void __launch_bounds__(kThreads) morton_kernel(const Box6* __restrict__ box, int n, const int* __restrict__ bounds,
                                                          unsigned* __restrict__ code, unsigned* __restrict__ index) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  unsigned q[3];
  for (int k = 0; k < 3; k++) {
    const float lo = order_float(bounds[k]), hi = order_float(bounds[3 + k]);
    const float c = 0.5f * (box[i].lo[k] + box[i].hi[k]);
    const float ext = hi - lo;
    const float u = ext > 0.0f ? (c - lo) / ext : 0.0f;
    q[k] = (unsigned)fminf(fmaxf(u * 1024.0f, 0.0f), 1023.0f);
  }
  code[i] = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
  index[i] = (unsigned)i;
}

// ---- 3. stable LSD radix sort of (code, index), 8 bits per pass
constexpr int kSortItems = 8;                                    // keys per thread
constexpr int kSortTile = kThreads * kSortItems;                 // keys per block
constexpr int kWarps = kThreads / 32;

__global__ void __launch_bounds__(kThreads) sort_histogram_kernel(const unsigned* __restrict__ key, int n, int shift, unsigned* __restrict__ hist, int n_blocks) {
  __shared__ unsigned h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * kSortTile;
  for (int k = 0; k < kSortItems; k++) {
    const int i = base + k * kThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(key[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];    // digit-major: one exclusive scan gives every (digit, block) its base
}

// exclusive scan of `count` words in place, one block (count = 256 * n_blocks: a few hundred thousand at most)
__global__ void __launch_bounds__(1024) scan_kernel(unsigned* __restrict__ data, int count) {
  __shared__ unsigned warp_sum[32];
  __shared__ unsigned carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < count; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned v = i < count ? data[i] : 0u;
    unsigned x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(kFullMask, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned w = warp_sum[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(kFullMask, w, o);
        if (threadIdx.x >= o) w += y;
      }
      warp_sum[threadIdx.x] = w;
    }
    __syncthreads();
    const unsigned before = carry + (threadIdx.x >= 32 ? warp_sum[(threadIdx.x >> 5) - 1] : 0u) + x - v;
    if (i < count) data[i] = before;
    __syncthreads();
    if (threadIdx.x == 1023) carry = before + v;
    __syncthreads();
  }
}

// Stable scatter.  A block owns kSortTile consecutive keys, warp w of it the w-th run of kSortTile / kWarps of them; a key's
// destination is  base(digit, block) + keys of that digit in earlier warps of the block + keys of that digit earlier in the warp's run.
__global__ void __launch_bounds__(kThreads) sort_scatter_kernel(const unsigned* __restrict__ key_in, const unsigned* __restrict__ val_in,
                                                                unsigned* __restrict__ key_out, unsigned* __restrict__ val_out, int n, int shift,
                                                                const unsigned* __restrict__ base, int n_blocks) {
  __shared__ unsigned cnt[kWarps][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = lane; d < 256; d += 32) cnt[warp][d] = 0;
  __syncwarp();
  constexpr int kRun = kSortTile / kWarps;                       // 256 keys per warp, 8 groups of 32
  const int run0 = blockIdx.x * kSortTile + warp * kRun;
  for (int g = 0; g < kRun / 32; g++) {
    const int i = run0 + g * 32 + lane;
    if (i < n) atomicAdd(&cnt[warp][(key_in[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  // per digit: exclusive scan over the warps, plus the block's global base
  {
    const int d = threadIdx.x;                                   // kThreads == 256 digits
    unsigned run = base[d * n_blocks + blockIdx.x];
    for (int w = 0; w < kWarps; w++) { const unsigned c = cnt[w][d]; cnt[w][d] = run; run += c; }
  }
  __syncthreads();
  for (int g = 0; g < kRun / 32; g++) {
    const int i = run0 + g * 32 + lane;
    const bool live = i < n;
    const unsigned k = live ? key_in[i] : 0u;
    const unsigned d = live ? ((k >> shift) & 255u) : 256u + lane;          // dead lanes match nobody
    const unsigned peers = __match_any_sync(kFullMask, d);
    const unsigned before = __popc(peers & ((1u << lane) - 1u));
    unsigned dst = 0;
    if (live) dst = cnt[warp][d] + before;
    __syncwarp();
    if (live && before == 0) cnt[warp][d] += __popc(peers);                 // the first lane of each digit group advances the cursor
    __syncwarp();
    if (live) { key_out[dst] = k; val_out[dst] = val_in[i]; }
  }
}

// ---- 4. Karras 2012: binary radix tree over sorted unique keys (code << 32 | index)
__device__ __forceinline__ int delta(const unsigned* __restrict__ code, const unsigned* __restrict__ index, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const unsigned long long a = ((unsigned long long)code[i] << 32) | index[i], b = ((unsigned long long)code[j] << 32) | index[j];
  return __clzll((long long)(a ^ b));
}
// node ids: inner nodes 0 .. n-2 (0 = root), leaf of sorted position k = (n - 1) + k
__global__ void __launch_bounds__(kThreads) radix_tree_kernel(const unsigned* __restrict__ code, const unsigned* __restrict__ index, int n,
                                                              int2* __restrict__ child, int* __restrict__ parent) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n - 1) return;
  const int d = delta(code, index, n, i, i + 1) - delta(code, index, n, i, i - 1) >= 0 ? 1 : -1;
  const int dmin = delta(code, index, n, i, i - d);
  int lmax = 2;
  while (delta(code, index, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta(code, index, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = delta(code, index, n, i, j);
  int s = 0;
  int t = l;
  do {
    t = (t + 1) >> 1;
    if (delta(code, index, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  const int split = i + s * d + min(d, 0);
  const int left = min(i, j) == split ? (n - 1) + split : split;
  const int right = max(i, j) == split + 1 ? (n - 1) + split + 1 : split + 1;
  child[i] = make_int2(left, right);
  parent[left] = i;
  parent[right] = i;
  if (i == 0) parent[0] = -1;
}

// ---- 5. bottom-up: boxes, triangle counts, emitted-node counts
struct NodeInfo { int tris; int emits; };                        // triangles below the node; nodes the subtree emits (0: it collapses into a leaf)

__global__ void __launch_bounds__(kThreads) fit_kernel(const unsigned* __restrict__ index, const Box6* __restrict__ tri_box, int n,
                                                       const int2* __restrict__ child, const int* __restrict__ parent, Box6* __restrict__ node_box,
                                                       NodeInfo* __restrict__ info, int* __restrict__ arrived, int leaf_target) {
  const int k = blockIdx.x * kThreads + threadIdx.x;
  if (k >= n) return;
  const int leaf = (n - 1) + k;
  node_box[leaf] = tri_box[index[k]];
  info[leaf] = NodeInfo{1, 0};
  __threadfence();
  int node = parent[leaf];
  while (node >= 0) {
    if (atomicAdd(arrived + node, 1) == 0) return;               // the first arrival leaves; the second one sees both children
    __threadfence();
    const int2 c = child[node];
    // the children's records were written by other SMs: read them from L2 (ld.global.cg), not through this SM's L1
    const float* pa = (const float*)(node_box + c.x);
    const float* pb = (const float*)(node_box + c.y);
    Box6 m;
    for (int q = 0; q < 3; q++) { m.lo[q] = fminf(__ldcg(pa + q), __ldcg(pb + q)); m.hi[q] = fmaxf(__ldcg(pa + 3 + q), __ldcg(pb + 3 + q)); }
    NodeInfo ia, ib;
    ia.tris = __ldcg(&info[c.x].tris); ia.emits = __ldcg(&info[c.x].emits);
    ib.tris = __ldcg(&info[c.y].tris); ib.emits = __ldcg(&info[c.y].emits);
    NodeInfo mi;
    mi.tris = ia.tris + ib.tris;
    mi.emits = (mi.tris > leaf_target || node == 0) ? 1 + ia.emits + ib.emits : 0;   // the root is always emitted
    node_box[node] = m;
    info[node] = mi;
    __threadfence();
    node = parent[node];
  }
}

// ---- 6 + 7. layout by climbing, emission
// Position of an emitted node in depth-first order (parent, left subtree, right subtree) and first triangle of its range:
// climbing from the node to the root, every step adds 1 (the ancestor itself precedes it) and, where the node lies in the
// ancestor's RIGHT subtree, everything the left sibling emits / holds.
__device__ __forceinline__ void place(int node, const int2* __restrict__ child, const int* __restrict__ parent, const NodeInfo* __restrict__ info,
                                      int& out_index, int& first_tri, int& depth) {
  out_index = 0; first_tri = 0; depth = 0;
  int c = node;
  for (int a = parent[c]; a >= 0; c = a, a = parent[a]) {
    out_index += 1;
    depth += 1;
    const int2 ch = child[a];
    if (ch.y == c) { out_index += info[ch.x].emits; first_tri += info[ch.x].tris; }
  }
}

// leaf triangles of a collapsed subtree (<= leaf_target <= 8 triangles), left to right
__device__ int gather_leaf(int node, int n, const int2* __restrict__ child, const unsigned* __restrict__ index, int* out) {
  int stack[16];
  int sp = 0, count = 0;
  stack[sp++] = node;
  while (sp > 0) {
    const int c = stack[--sp];
    if (c >= n - 1) out[count++] = (int)index[c - (n - 1)];
    else { const int2 ch = child[c]; stack[sp++] = ch.y; stack[sp++] = ch.x; }
  }
  return count;
}

__global__ void __launch_bounds__(kThreads) emit_kernel(int n, const int2* __restrict__ child, const int* __restrict__ parent, const NodeInfo* __restrict__ info,
                                                        const Box6* __restrict__ node_box, const unsigned* __restrict__ index,
                                                        const LrTriangle* __restrict__ tris_in, LrTriangle* __restrict__ tris_out,
                                                        LrBvhNode* __restrict__ nodes_out, float pad, int* __restrict__ max_depth) {
  const int node = blockIdx.x * kThreads + threadIdx.x;
  if (node >= n - 1 || info[node].emits == 0) return;
  int me, first, depth;
  place(node, child, parent, info, me, first, depth);
  atomicMax(max_depth, depth + 1);
  const int2 ch = child[node];
  LrBvhNode out;
  int tri_at = first, node_at = me + 1;
  for (int slot = 0; slot < 2; slot++) {
    const int c = slot == 0 ? ch.x : ch.y;
    const Box6 b = node_box[c];
    for (int q = 0; q < 3; q++) { out.f[slot * 6 + q] = b.lo[q] - pad; out.f[slot * 6 + 3 + q] = b.hi[q] + pad; }
    const NodeInfo ci = info[c];
    if (ci.emits > 0) {
      out.c[slot] = node_at;
      out.n[slot] = 0;
    } else {
      int ids[8];
      const int count = gather_leaf(c, n, child, index, ids);
      for (int k = 0; k < count; k++) tris_out[tri_at + k] = tris_in[ids[k]];
      out.c[slot] = ~((tri_at << 3) | (count - 1));
      out.n[slot] = count;
    }
    tri_at += ci.tris;
    node_at += ci.emits;
  }
  nodes_out[me] = out;
}

// scratch from the stream-ordered allocator (the device's default pool keeps freed blocks, lr_init: a second build reuses
// them in microseconds where cudaMalloc / cudaFree cost milliseconds each and synchronise the device)
struct DeviceBuffers {
  std::vector<void*> ptrs;
  template <class T> cudaError_t alloc(T** p, size_t count) {
    cudaError_t e = cudaMallocAsync((void**)p, std::max<size_t>(count, 1) * sizeof(T), 0);
    if (e == cudaSuccess) ptrs.push_back(*p);
    return e;
  }
  ~DeviceBuffers() { for (void* p : ptrs) cudaFreeAsync(p, 0); }
};

}  // namespace

// Builds the tree over tris[0, n_tree) (the flat tail behind it is left alone) on the current device; on success `tris` is
// permuted into leaf order and nodes_out holds the flattened tree.  leaf_target in 1..8.  seconds_out covers everything
// from the host triangles to the host node array (H2D, kernels, D2H).
// If the tree comes out deeper than max_depth, `tris` and nodes_out are left untouched (depth_out says why).
int build_bvh_device(std::vector<LrTriangle>& tris, int n_tree, float pad, int leaf_target, int max_depth, std::vector<LrBvhNode>& nodes_out,
                     int& depth_out, float& seconds_out, float& kernel_ms_out) {
  const auto t0 = std::chrono::steady_clock::now();
  const int n = n_tree;
  if (n < 3) return fail(LR_ERR_INVALID, "build_bvh_device needs at least 3 triangles");
  int device_count = 0;
  if (cudaGetDeviceCount(&device_count) != cudaSuccess || device_count <= 0)
    return fail(LR_ERR_NO_DEVICE, "the device BVH builder needs a CUDA device (there is no CPU fallback for it; use the host builder)");
  int sm_count = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);

  DeviceBuffers buf;
  LrTriangle *d_in = nullptr, *d_out = nullptr;
  Box6 *d_tri_box = nullptr, *d_node_box = nullptr;
  unsigned *d_code[2] = {nullptr, nullptr}, *d_index[2] = {nullptr, nullptr}, *d_hist = nullptr;
  int *d_bounds = nullptr, *d_parent = nullptr, *d_arrived = nullptr, *d_depth = nullptr;
  int2* d_child = nullptr;
  NodeInfo* d_info = nullptr;
  LrBvhNode* d_nodes = nullptr;
  const int n_sort_blocks = (n + kSortTile - 1) / kSortTile;
  const int n_all = 2 * n - 1;
  cudaError_t e = buf.alloc(&d_in, n);
  if (e == cudaSuccess) e = buf.alloc(&d_out, n);
  if (e == cudaSuccess) e = buf.alloc(&d_tri_box, n);
  if (e == cudaSuccess) e = buf.alloc(&d_node_box, n_all);
  for (int k = 0; k < 2 && e == cudaSuccess; k++) { e = buf.alloc(&d_code[k], n); if (e == cudaSuccess) e = buf.alloc(&d_index[k], n); }
  if (e == cudaSuccess) e = buf.alloc(&d_hist, (size_t)256 * n_sort_blocks);
  if (e == cudaSuccess) e = buf.alloc(&d_bounds, 6);
  if (e == cudaSuccess) e = buf.alloc(&d_parent, n_all);
  if (e == cudaSuccess) e = buf.alloc(&d_arrived, n);
  if (e == cudaSuccess) e = buf.alloc(&d_depth, 1);
  if (e == cudaSuccess) e = buf.alloc(&d_child, n);
  if (e == cudaSuccess) e = buf.alloc(&d_info, n_all);
  if (e == cudaSuccess) e = buf.alloc(&d_nodes, n);
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (e == cudaSuccess) e = cudaEventCreate(&ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&ev1);
  int n_emitted = 0;
  if (e == cudaSuccess) {
    e = cudaMemcpyAsync(d_in, tris.data(), (size_t)n * sizeof(LrTriangle), cudaMemcpyHostToDevice, 0);
    const int init_bounds[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_bounds, init_bounds, sizeof(init_bounds), cudaMemcpyHostToDevice, 0);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_arrived, 0, (size_t)n * sizeof(int), 0);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_depth, 0, sizeof(int), 0);
    if (e == cudaSuccess) e = cudaEventRecord(ev0, 0);
    const int blocks = (n + kThreads - 1) / kThreads;
    if (e == cudaSuccess) {
      tri_bounds_kernel<<<std::min(blocks, sm_count * 8), kThreads>>>(d_in, n, d_tri_box, d_bounds);
      morton_kernel<<<blocks, kThreads>>>(d_tri_box, n, d_bounds, d_code[0], d_index[0]);
      int cur = 0;
      for (int pass = 0; pass < 4; pass++) {                     // 30 code bits: 4 passes of 8
        sort_histogram_kernel<<<n_sort_blocks, kThreads>>>(d_code[cur], n, pass * 8, d_hist, n_sort_blocks);
        scan_kernel<<<1, 1024>>>(d_hist, 256 * n_sort_blocks);
        sort_scatter_kernel<<<n_sort_blocks, kThreads>>>(d_code[cur], d_index[cur], d_code[cur ^ 1], d_index[cur ^ 1], n, pass * 8, d_hist, n_sort_blocks);
        cur ^= 1;
      }
      radix_tree_kernel<<<(n - 1 + kThreads - 1) / kThreads, kThreads>>>(d_code[cur], d_index[cur], n, d_child, d_parent);
      fit_kernel<<<blocks, kThreads>>>(d_index[cur], d_tri_box, n, d_child, d_parent, d_node_box, d_info, d_arrived, leaf_target);
      if (e == cudaSuccess) emit_kernel<<<(n - 1 + kThreads - 1) / kThreads, kThreads>>>(n, d_child, d_parent, d_info, d_node_box, d_index[cur], d_in, d_out, d_nodes, pad, d_depth);
      if (e == cudaSuccess) e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaEventRecord(ev1, 0);
    NodeInfo root{};
    if (e == cudaSuccess) e = cudaMemcpy(&root, d_info, sizeof(root), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(&depth_out, d_depth, sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&kernel_ms_out, ev0, ev1);
    if (e == cudaSuccess && depth_out < max_depth) {
      n_emitted = root.emits;
      if (root.tris != n || n_emitted < 1 || n_emitted > n - 1) {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        return fail(LR_ERR_CUDA, "device BVH build produced an inconsistent tree");
      }
      nodes_out.resize(n_emitted);
      e = cudaMemcpy(nodes_out.data(), d_nodes, (size_t)n_emitted * sizeof(LrBvhNode), cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) e = cudaMemcpy(tris.data(), d_out, (size_t)n * sizeof(LrTriangle), cudaMemcpyDeviceToHost);
    }
  }
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  if (e != cudaSuccess) return fail(LR_ERR_CUDA, std::string("device BVH build: ") + cudaGetErrorString(e));
  seconds_out = std::chrono::duration<float>(std::chrono::steady_clock::now() - t0).count();
  return LR_OK;
}

}  // namespace lr
