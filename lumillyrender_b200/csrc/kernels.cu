// kernels.cu — hand-written sm_100a kernels of the path-tracing hot path.
//
// K1-K7 of SURVEY.md §2 live in ONE persistent-per-pixel megakernel (render_kernel):
//   camera ray generation -> BVH traversal + triangle/sphere tests -> emission / Russian
//   roulette -> next-event estimation (shadow ray through the same trace call site) ->
//   BSDF sample + eval -> sky lookup -> fp32 accumulation.
// Organisation: one thread per (pixel, spp split); a warp owns an 8x4 pixel tile so primary rays
// are coherent.  Each thread runs a flattened state machine with PATH REGENERATION: a lane whose
// path ended immediately starts the next sample of its pixel instead of idling until the warp's
// longest path finishes, and every iteration funnels through a single trace() call site
// (extension rays and NEE shadow rays alike) so the warp reconverges at the memory-heavy part.
// Per-pixel sums are accumulated in registers in sample order (main.rs:92-104) and written once.
//
// Compiled with -fmad=false (see device_path.cuh for the numerics contract).
#include "device_path.cuh"
#include "kernels.h"

namespace lr {

// one thread = (pixel, split).  tile = warp = 8x4 pixels.
LR_DEV bool thread_pixel(const DevParams& p, long long gid, int& lx, int& ly, int& split) {
  const long long warp = gid >> 5;
  const int lane = (int)(gid & 31);
  const long long tiles = (long long)p.tiles_x * p.tiles_y;
  split = (int)(warp / tiles);
  const long long tile = warp % tiles;
  const int tx = (int)(tile % p.tiles_x), ty = (int)(tile / p.tiles_x);
  lx = tx * 8 + (lane & 7);
  ly = ty * 4 + (lane >> 3);
  return split < p.splits && lx < p.crop_w && ly < p.crop_h;
}

LR_DEV unsigned int warp_sum(unsigned int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int INTEGRATOR, bool COUNT, bool SUMSQ>
__global__ void __launch_bounds__(kBlockThreads)
render_kernel(const __grid_constant__ DevScene sc, const __grid_constant__ DevParams p,
              float* __restrict__ out_sum, float* __restrict__ out_sumsq, unsigned long long* __restrict__ counters) {
  const long long gid = (long long)blockIdx.x * kBlockThreads + threadIdx.x;
  int lx, ly, split;
  const bool active = thread_pixel(p, gid, lx, ly, split);

  unsigned int n_rays = 0, n_nonfinite = 0;
  TraceCounters tc;
  tc.nodes = tc.tris = tc.spheres = 0;

  // Warp-synchronous loop: every iteration all lanes of the warp meet at __any_sync, lanes with work
  // trace ONE ray together (extension or shadow), then shade.  Lanes whose pixel is finished stay in the
  // loop (idle) until the whole warp is done, so the reconvergence points are well defined.
  const unsigned kFull = 0xffffffffu;
  const int x = p.crop_x + lx, y = p.crop_y + ly;
  const unsigned int pixel = (unsigned int)(y * sc.cam.width + x);
  // sample sub-range of this split
  const int per = p.spp_count / p.splits, rem = p.spp_count % p.splits;
  int s = p.spp_begin + split * per + min(split, rem);
  const int s_end = active ? s + per + (split < rem ? 1 : 0) : s;

  F3 sum = f3(0.0f, 0.0f, 0.0f), sumsq = f3(0.0f, 0.0f, 0.0f);
  Pcg rng;
  rng.state = 0;
  // path state
  F3 o = f3(0, 0, 0), d = f3(0, 0, 1);
  F3 T = f3(1, 1, 1), L = f3(0, 0, 0);
  float cam_g = 1.0f, cam_w = 1.0f;
  int depth = 0;
  bool allow_emission = true;
  bool need_new = true;
  bool live = active;
  // vertex state (kept across the NEE shadow ray)
  bool shadow = false;
  F3 v_pos = f3(0, 0, 0), v_n = f3(0, 0, 1), v_wo = f3(0, 0, 1);
  int v_mat = 0;
  float v_prr = 1.0f, v_dist = 0.0f;
  float nee_dist = 0.0f, nee_sqr = 1.0f, nee_pdf = 1.0f;

  while (true) {
    if (live && need_new) {
      if (s >= s_end) {
        live = false;
      } else {
        rng.seed(p.seed, pixel, (unsigned int)s);
        camera_sample(sc.cam, x, y, [&]() { return rng.next(); }, o, d, cam_g, cam_w);
        T = f3(1.0f, 1.0f, 1.0f);
        L = f3(0.0f, 0.0f, 0.0f);
        depth = 0;
        allow_emission = true;
        need_new = false;
        shadow = false;
      }
    }
    if (!__any_sync(kFull, live)) break;

    float t = 0.0f;
    int id = -1;
    if (live) {
      trace<COUNT>(sc, o, d, t, id, tc);          // Objects::intersect (objects.rs:63-65)
      n_rays++;
    }
    __syncwarp(kFull);

    if (live) {
      bool finish = false;
      bool sample_bsdf = false;
      if (!shadow) {
        if (id == -1) {
          L = L + T * sky_radiance(sc, d);        // scene.rs:29 / 43
          finish = true;
        } else {
          const Surface sf = surface_at(sc, o, d, t, id);
          const Mat m = load_mat(sc, sf.mat);
          const F3 wo = -d;
          // emission: scene.rs:155-159 / 175-179
          if (!(p.no_direct_emitter && depth == 0) && allow_emission && dot(wo, sf.n) > 0.0f && m.emissive)
            L = L + T * m.emission;
          // Russian roulette: scene.rs:64-76, 161-164
          float prr = m.weight;
          if (depth > p.depth_limit) prr *= scalbnf(1.0f, -(depth - p.depth_limit));
          if (depth <= p.depth && prr > 0.0f) prr = 1.0f;
          if (prr != 1.0f && rng.next() >= prr) {
            finish = true;
          } else {
            v_pos = sf.pos; v_n = sf.n; v_wo = wo; v_mat = sf.mat; v_prr = prr; v_dist = t;
            sample_bsdf = true;
            if (INTEGRATOR == LR_INTEGRATOR_PT_DIRECT) {
              allow_emission = false;             // every deeper vertex: no_emission = true (scene.rs:189)
              // direct_light_radiance: scene.rs:104-125
              if (!m.emissive && sc.n_emitters > 0) {
                // Objects::sample_emission objects.rs:37-51 (prefix-sum CDF, first i with roulette <= cdf[i])
                const float roulette = sc.emission_area * rng.next();
                int lo = 0, hi = sc.n_emitters - 1;
                while (lo < hi) {
                  const int mid = (lo + hi) >> 1;
                  if (roulette <= __ldg(sc.emitter_cdf + mid)) hi = mid; else lo = mid + 1;
                }
                const float4 e0 = ldg4(sc.emitters + 3 * lo), e1 = ldg4(sc.emitters + 3 * lo + 1), e2 = ldg4(sc.emitters + 3 * lo + 2);
                const float u1 = rng.next();
                const float u2 = rng.next();
                F3 q;
                const float area = e1.w;
                if (__float_as_int(e0.w) == 0) {  // Triangle::sample triangle.rs:140-149
                  const float mn = fminf(u1, u2), mx = fmaxf(u1, u2);
                  q = f3(e0) * mn + f3(e1) * (1.0f - mx) + f3(e2) * (mx - mn);
                } else {                          // Sphere::sample sphere.rs:79-84 + util.rs:108-116
                  const float r1 = 2.0f * kPI * u1;
                  const float r2 = u2 * 2.0f - 1.0f;
                  const float r2s = sqrtf(1.0f - r2 * r2);
                  float sn, cs;
                  spec_sincos(r1, &sn, &cs);
                  q = f3(e0) + e1.x * f3(cs * r2s, sn * r2s, r2);
                }
                nee_pdf = (1.0f / area) * area / sc.emission_area;
                const F3 direct_path = q - sf.pos;
                const F3 dir = normalize(direct_path);
                const F3 pn = orienting_normal(wo, sf.n);
                if (dot(dir, pn) > 0.0f) {
                  nee_dist = norm(direct_path);
                  nee_sqr = sqr_norm(direct_path);
                  o = sf.pos;
                  d = dir;
                  shadow = true;                  // the shadow ray is traced at the common call site next iteration
                  sample_bsdf = false;
                }
              }
            }
          }
        }
      } else {
        // visibility + contribution: scene.rs:127-150
        shadow = false;
        sample_bsdf = true;
        if (id != -1 && fabsf(t - nee_dist) <= kEPS) {
          const Surface lf = surface_at(sc, o, d, t, id);
          const float light_cos = dot(-d, lf.n);
          if (light_cos > 0.0f) {
            const Mat lm = load_mat(sc, lf.mat);
            const Mat m = load_mat(sc, v_mat);
            const F3 pn = orienting_normal(v_wo, v_n);
            const float point_cos = dot(d, pn);
            const float g_term = point_cos * light_cos / nee_sqr;
            const F3 brdf = mat_brdf(m, v_wo, d, pn, v_pos);
            const F3 l_i = lm.emissive ? lm.emission : f3(0.0f, 0.0f, 0.0f);
            const F3 direct = brdf * l_i * g_term / nee_pdf;
            L = L + T * (direct / v_prr);         // scene.rs:192
          }
        }
      }

      if (sample_bsdf) {
        // material_interaction_radiance: scene.rs:78-102
        const Mat m = load_mat(sc, v_mat);
        F3 wi;
        float pdf;
        mat_sample(m, v_wo, v_n, rng, wi, pdf);
        const F3 brdf = mat_brdf(m, v_wo, wi, v_n, v_pos);
        const F3 coef = mat_coef(m, v_wo, v_n, v_dist);
        const float c = dot(wi, v_n);             // UNoriented normal (scene.rs:91)
        T = T * (brdf * coef * c / pdf) / v_prr;
        o = v_pos;                                // no origin offset (scene.rs:94-97)
        d = wi;
        depth++;
      }

      if (finish) {
        // main.rs:99-102
        const F3 e = (L * cam_g) * cam_w;
        if (!(isfinite(e.x) && isfinite(e.y) && isfinite(e.z))) n_nonfinite++;
        sum = sum + e;
        if (SUMSQ) sumsq = sumsq + e * e;
        s++;
        need_new = true;
      }
    }
  }

  if (active) {
    const size_t pi = (size_t)ly * p.crop_w + lx;
    const size_t n_px = (size_t)p.crop_w * p.crop_h;
    if (p.splits == 1) {
      out_sum[3 * pi + 0] += sum.x; out_sum[3 * pi + 1] += sum.y; out_sum[3 * pi + 2] += sum.z;
      if (SUMSQ) { out_sumsq[3 * pi + 0] += sumsq.x; out_sumsq[3 * pi + 1] += sumsq.y; out_sumsq[3 * pi + 2] += sumsq.z; }
    } else {
      // per-split partial buffers, reduced in split order by reduce_splits_kernel (deterministic)
      float* ps = out_sum + 3 * (n_px * split + pi);
      ps[0] = sum.x; ps[1] = sum.y; ps[2] = sum.z;
      if (SUMSQ) { float* pq = out_sumsq + 3 * (n_px * split + pi); pq[0] = sumsq.x; pq[1] = sumsq.y; pq[2] = sumsq.z; }
    }
  }

  // counters: warp reduce, one atomic per warp
  n_rays = warp_sum(n_rays);
  n_nonfinite = warp_sum(n_nonfinite);
  if (COUNT) { tc.nodes = warp_sum(tc.nodes); tc.tris = warp_sum(tc.tris); tc.spheres = warp_sum(tc.spheres); }
  if ((threadIdx.x & 31) == 0) {
    if (n_rays) atomicAdd(counters + C_RAYS, (unsigned long long)n_rays);
    if (n_nonfinite) atomicAdd(counters + C_NONFINITE, (unsigned long long)n_nonfinite);
    if (COUNT) {
      atomicAdd(counters + C_NODES, (unsigned long long)tc.nodes);
      atomicAdd(counters + C_TRIS, (unsigned long long)tc.tris);
      atomicAdd(counters + C_SPHERES, (unsigned long long)tc.spheres);
    }
  }
}

// dst[i] += sum_k partial[k][i] in k order
__global__ void reduce_splits_kernel(float* __restrict__ dst, const float* __restrict__ partial, size_t n, int splits) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.0f;
  for (int k = 0; k < splits; k++) acc += partial[(size_t)k * n + i];
  dst[i] += acc;
}

__global__ void scale_kernel(float* __restrict__ dst, size_t n, float divisor) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = dst[i] / divisor;   // `estimated_sum / spp as f32` (main.rs:104): a true division
}

// parity probe: nearest hit of the primary ray of every film pixel with fixed random numbers
__global__ void __launch_bounds__(kBlockThreads)
primary_kernel(const __grid_constant__ DevScene sc, int tiles_x, int tiles_y, float u, float v, float ua, float va,
               int* __restrict__ prim, float* __restrict__ tout) {
  const long long gid = (long long)blockIdx.x * kBlockThreads + threadIdx.x;
  const long long tile = gid >> 5;
  const int lane = (int)(gid & 31);
  if (tile >= (long long)tiles_x * tiles_y) return;
  const int x = (int)(tile % tiles_x) * 8 + (lane & 7), y = (int)(tile / tiles_x) * 4 + (lane >> 3);
  if (x >= sc.cam.width || y >= sc.cam.height) return;
  const float draws[4] = {u, v, ua, va};
  int k = 0;
  F3 o, d;
  float g, w;
  camera_sample(sc.cam, x, y, [&]() { return draws[k++]; }, o, d, g, w);
  float t;
  int id;
  TraceCounters tc;
  trace<false>(sc, o, d, t, id, tc);
  const size_t i = (size_t)y * sc.cam.width + x;
  if (id == -1) { prim[i] = -1; tout[i] = 0.0f; }
  else { const Surface s = surface_at(sc, o, d, t, id); prim[i] = s.prim; tout[i] = t; }
}

__global__ void __launch_bounds__(kBlockThreads)
rays_kernel(const __grid_constant__ DevScene sc, long long n, const float* __restrict__ org, const float* __restrict__ dir,
            int* __restrict__ prim, float* __restrict__ tout, float* __restrict__ nout) {
  const long long i = (long long)blockIdx.x * kBlockThreads + threadIdx.x;
  if (i >= n) return;
  const F3 o = f3(org + 3 * i), d = f3(dir + 3 * i);
  float t;
  int id;
  TraceCounters tc;
  trace<false>(sc, o, d, t, id, tc);
  if (id == -1) {
    prim[i] = -1; tout[i] = 0.0f;
    if (nout) { nout[3 * i] = 0.0f; nout[3 * i + 1] = 0.0f; nout[3 * i + 2] = 0.0f; }
  } else {
    const Surface s = surface_at(sc, o, d, t, id);
    prim[i] = s.prim; tout[i] = t;
    if (nout) { nout[3 * i] = s.n.x; nout[3 * i + 1] = s.n.y; nout[3 * i + 2] = s.n.z; }
  }
}

// bandwidth microbenchmark: every thread streams 128-bit loads over a working set `n4` float4s,
// `iters` passes.  With a working set << 126 MB it measures the L2 read peak, >> 126 MB the HBM read peak.
__global__ void __launch_bounds__(256) read_bw_kernel(const float4* __restrict__ buf, size_t n4, int iters, float* __restrict__ sink) {
  float acc = 0.0f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; it++) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n4; i += 4 * stride) {
      float4 a, b, c, d;
      asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(buf + i));
      asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(buf + i + stride));
      asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "l"(buf + i + 2 * stride));
      asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "l"(buf + i + 3 * stride));
      acc += a.x + b.y + c.z + d.w;
    }
    for (; i < n4; i += stride) {
      float4 a;
      asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(buf + i));
      acc += a.x;
    }
  }
  if (acc == 1.2345e-30f) *sink = acc;
}

// ------------------------------------------------------------------ launchers
template <int INTEGRATOR>
static cudaError_t launch_render_t(const DevScene& sc, const DevParams& p, bool count, bool sumsq, float* out_sum,
                                   float* out_sumsq, unsigned long long* counters, cudaStream_t stream) {
  const long long threads = (long long)p.tiles_x * p.tiles_y * p.splits * 32;
  const unsigned int blocks = (unsigned int)((threads + kBlockThreads - 1) / kBlockThreads);
  if (count) {
    if (sumsq) render_kernel<INTEGRATOR, true, true><<<blocks, kBlockThreads, 0, stream>>>(sc, p, out_sum, out_sumsq, counters);
    else render_kernel<INTEGRATOR, true, false><<<blocks, kBlockThreads, 0, stream>>>(sc, p, out_sum, out_sumsq, counters);
  } else {
    if (sumsq) render_kernel<INTEGRATOR, false, true><<<blocks, kBlockThreads, 0, stream>>>(sc, p, out_sum, out_sumsq, counters);
    else render_kernel<INTEGRATOR, false, false><<<blocks, kBlockThreads, 0, stream>>>(sc, p, out_sum, out_sumsq, counters);
  }
  return cudaGetLastError();
}

cudaError_t launch_render(const DevScene& sc, const DevParams& p, bool count, bool sumsq, float* out_sum, float* out_sumsq,
                          unsigned long long* counters, cudaStream_t stream) {
  if (p.integrator == LR_INTEGRATOR_PT) return launch_render_t<LR_INTEGRATOR_PT>(sc, p, count, sumsq, out_sum, out_sumsq, counters, stream);
  return launch_render_t<LR_INTEGRATOR_PT_DIRECT>(sc, p, count, sumsq, out_sum, out_sumsq, counters, stream);
}

cudaError_t launch_reduce_splits(float* dst, const float* partial, size_t n, int splits, cudaStream_t stream) {
  reduce_splits_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, stream>>>(dst, partial, n, splits);
  return cudaGetLastError();
}

cudaError_t launch_scale(float* dst, size_t n, float divisor, cudaStream_t stream) {
  scale_kernel<<<(unsigned int)((n + 255) / 256), 256, 0, stream>>>(dst, n, divisor);
  return cudaGetLastError();
}

cudaError_t launch_primary(const DevScene& sc, float u, float v, float ua, float va, int* prim, float* t, cudaStream_t stream) {
  const int tiles_x = (sc.cam.width + 7) / 8, tiles_y = (sc.cam.height + 3) / 4;
  const long long threads = (long long)tiles_x * tiles_y * 32;
  primary_kernel<<<(unsigned int)((threads + kBlockThreads - 1) / kBlockThreads), kBlockThreads, 0, stream>>>(sc, tiles_x, tiles_y, u, v, ua, va, prim, t);
  return cudaGetLastError();
}

cudaError_t launch_rays(const DevScene& sc, long long n, const float* org, const float* dir, int* prim, float* t, float* nout, cudaStream_t stream) {
  rays_kernel<<<(unsigned int)((n + kBlockThreads - 1) / kBlockThreads), kBlockThreads, 0, stream>>>(sc, n, org, dir, prim, t, nout);
  return cudaGetLastError();
}

cudaError_t launch_read_bw(const float4* buf, size_t n4, int iters, int blocks, float* sink, cudaStream_t stream) {
  read_bw_kernel<<<blocks, 256, 0, stream>>>(buf, n4, iters, sink);
  return cudaGetLastError();
}

}  // namespace lr
