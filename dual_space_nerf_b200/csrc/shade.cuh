// Shading (normal_local2world + LightingMLP) and compositing (raw2outputs).
#pragma once
#include "geom.cuh"

namespace dsn {

struct LightWeights {
  const float* w1t;  // [9][128]   lights_encoding.0 transposed
  const float* b1;   // [128]
  const float* w2t;  // [128][128] lights_encoding.2 transposed
  const float* b2;   // [128]
  const float* w3;   // [128]      lights_encoding.4
  float b3;
};

struct ShadeArgs {
  const float4* active;   // (xyz_cano, bits(sample))
  const int* active_tri;  // posed-space nearest triangle per active sample (search hint), may be NULL
  const float4* mlp_a;    // (sigma, essence)
  const float4* mlp_g;    // (d sigma / d xyz_cano, -)
  const unsigned long long* n_active;
  int64_t n_active_host;  // used when n_active == NULL
  const float* ray_o; const float* ray_d; const float* near; const float* far; const float* z_in; const float* tvals;
  int N;
  const float* posed; const float* canon; const int* faces; const float* cent_canon; int F;
  const float4* normal_m;  // (F,3) rows of the per-triangle canonical -> world normal map of the current frame
  const int* active_cidx;  // nearest canonical triangle per active sample (canon_nearest_kernel); NULL => searched in place
  // explicit-point mode (dsnerf_eval_points): world position / view direction per point instead of rays
  const float* xyz_world; const float* view_dir;
  float light_shift[3]; int has_shift;
  float rot[4]; float rot_center[2]; int has_rot;
  float4* raw;            // (R*N) rgb + sigma, indexed by sample id
};

constexpr int SHADE_THREADS = 256;
// shared memory: first layer as [128 units][12] (9 weights, bias, 2 pad), second layer transposed [128][128], b2, w3
constexpr size_t SHADE_SMEM = (size_t)(128 * 12 + 128 * 128 + 128 + 128) * sizeof(float);

// model/spacenet.py:278-298 normal_local2world.  The reference maps xyz_cano and xyz_cano+g onto the posed triangle and
// normalises the difference; the map is affine in the point, so the difference is the linear part applied to g.
// normal_local2world is linear in the gradient: r = M_f g with M_f = e2w a^T + e1w b^T + nw nc^T per triangle f, where
// (a, b) = rows of the inverse Gram matrix applied to the canonical edges (u = a.g, v = b.g; the edges are orthogonal to nc)
// and nc / nw are the unit normals of the canonical / posed triangle.  M_f depends on the frame only: one thread per
// triangle fills it in dsnerf_set_frame, the shading kernels then need 3 loads and 9 FMAs per sample instead of 21 gathers
// and ~150 operations.
__global__ void normal_matrix_kernel(const float* __restrict__ canon, const float* __restrict__ posed, const int* __restrict__ faces, int F,
                                     float4* __restrict__ M) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
  const V3 c0 = ldv3(canon, i0), c1 = ldv3(canon, i1), c2 = ldv3(canon, i2);
  const V3 m0 = ldv3(posed, i0), m1 = ldv3(posed, i1), m2 = ldv3(posed, i2);
  V3 e1c = v3(c1.x - c0.x, c1.y - c0.y, c1.z - c0.z), e2c = v3(c2.x - c0.x, c2.y - c0.y, c2.z - c0.z);
  V3 nc = v3(e1c.y * e2c.z - e1c.z * e2c.y, e1c.z * e2c.x - e1c.x * e2c.z, e1c.x * e2c.y - e1c.y * e2c.x);
  float inc = rsqrtf(nc.x * nc.x + nc.y * nc.y + nc.z * nc.z);
  nc = v3(nc.x * inc, nc.y * inc, nc.z * inc);
  float d00 = e2c.x * e2c.x + e2c.y * e2c.y + e2c.z * e2c.z;
  float d01 = e2c.x * e1c.x + e2c.y * e1c.y + e2c.z * e1c.z;
  float d11 = e1c.x * e1c.x + e1c.y * e1c.y + e1c.z * e1c.z;
  float inv = 1.0f / (d00 * d11 - d01 * d01);
  V3 a = v3((d11 * e2c.x - d01 * e1c.x) * inv, (d11 * e2c.y - d01 * e1c.y) * inv, (d11 * e2c.z - d01 * e1c.z) * inv);
  V3 b = v3((d00 * e1c.x - d01 * e2c.x) * inv, (d00 * e1c.y - d01 * e2c.y) * inv, (d00 * e1c.z - d01 * e2c.z) * inv);
  V3 e1w = v3(m1.x - m0.x, m1.y - m0.y, m1.z - m0.z), e2w = v3(m2.x - m0.x, m2.y - m0.y, m2.z - m0.z);
  V3 nw = v3(e1w.y * e2w.z - e1w.z * e2w.y, e1w.z * e2w.x - e1w.x * e2w.z, e1w.x * e2w.y - e1w.y * e2w.x);
  float inw = rsqrtf(nw.x * nw.x + nw.y * nw.y + nw.z * nw.z);
  nw = v3(nw.x * inw, nw.y * inw, nw.z * inw);
  M[3 * f] = make_float4(e2w.x * a.x + e1w.x * b.x + nw.x * nc.x, e2w.x * a.y + e1w.x * b.y + nw.x * nc.y, e2w.x * a.z + e1w.x * b.z + nw.x * nc.z, 0.f);
  M[3 * f + 1] = make_float4(e2w.y * a.x + e1w.y * b.x + nw.y * nc.x, e2w.y * a.y + e1w.y * b.y + nw.y * nc.y, e2w.y * a.z + e1w.y * b.z + nw.y * nc.z, 0.f);
  M[3 * f + 2] = make_float4(e2w.z * a.x + e1w.z * b.x + nw.z * nc.x, e2w.z * a.y + e1w.z * b.y + nw.z * nc.y, e2w.z * a.z + e1w.z * b.z + nw.z * nc.z, 0.f);
}

// Inputs of the lighting MLP for active sample t: world normal (normal_local2world, model/spacenet.py:278-298, incl. the
// exact nearest canonical centroid), world position (with the optional rot / light_center shift, :254-263) and the
// normalised view direction (:179-181).
// exact nearest canonical centroid through the canonical mesh's lookup table (cells requested by mark_points_kernel)
__device__ __forceinline__ int canon_nearest(const Grid& gc, const float* __restrict__ cent, int F, float x, float y, float z) {
  int idx = table_nearest(gc, live_cell(gc, x, y, z), x, y, z);
  if (idx < 0) idx = brute_nearest(cent, F, x, y, z);  // outside the table / far from the canonical mesh (rare for warped points)
  return idx;
}
// The search is a chain of dependent gathers (cell byte -> record -> candidate list, ~45 candidates per lookup on the
// benchmark scene): it runs here at full occupancy instead of inside the register- and TMEM-limited lighting kernel
// (16 warps per SM), which then only reads the result (lighting 1.42 ms -> 0.47 + 0.36 ms).  Consecutive lookups are
// consecutive samples of a ray and mostly share a cell, so the lanes of a warp walk the same list with broadcast loads;
// bucketing the lookups of a block by list length (as sample_warp_kernel does) breaks that and was slower (0.82 ms).
//
// Two passes.  0.5 % of the lookups fall into cells whose candidates did not fit the builder's buffers (cells that see a dense
// cluster or a whole ring of centroids); they used to cost a ball scan through the grid each -- several thousand instructions,
// scattered one or two per warp, so that almost every fifth warp walked through a scan with one or two lanes busy: about
// half of the kernel's 82 M warp instructions (ncu: 16 of 32 lanes busy, 44 % warps active).  Now the builder writes those
// cells' lists straight into the pool (long_list, geom.cuh), pass 1 (canon_nearest_kernel) answers every lookup whose list has
// at most COOP_LIST entries and queues the others, and pass 2 (canon_long_kernel) gives every queued lookup to a whole warp:
// 32 candidates per step and a butterfly argmin (or warp_scan_ball_flat where there is no list: pool full); same arithmetic,
// strict '<', lowest index on ties, so the result is the one list_nearest / scan_nearest return.  gc.pool_used[19] counts the
// queue, [18] is pass 1's work counter (both zeroed by ensure_cells in front of every launch).
__global__ void __launch_bounds__(256) canon_nearest_kernel(const float4* __restrict__ active, const unsigned long long* __restrict__ n_ptr,
                                                            int64_t n_host, Grid gc, const float* __restrict__ cent, int F, int* __restrict__ out,
                                                            int* __restrict__ queue) {
  const int64_t n = n_ptr ? (int64_t)*n_ptr : n_host;
  const int lane = threadIdx.x & 31;
  // lists hold 1..COOP_LIST candidates and long ones come in runs: warps take chunks of 64 consecutive lookups from a work
  // counter (gc.pool_used[18]) instead of striding (132 vs 162-184 us; chunks of 256 are as slow as striding)
  unsigned int* work = reinterpret_cast<unsigned int*>(gc.pool_used + 18);
  for (;;) {
    unsigned int c = 0;
    if (lane == 0) c = atomicAdd(work, 1u);
    c = __shfl_sync(0xffffffffu, c, 0);
    if ((int64_t)c * 64 >= n) break;
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    const int64_t t = (int64_t)c * 64 + h * 32 + lane;
    bool deferred = false;
    if (t < n) {
      const float4 p = active[t];
      const int cell = live_cell(gc, p.x, p.y, p.z);
      int idx = -1;
      if (cell >= 0) {
        const int2 rec = gc.trec[cell];
        deferred = rec.y == -2 || rec.y > COOP_LIST;
        if (rec.y >= 0 && !deferred) idx = list_nearest(gc, rec.x, rec.y, p.x, p.y, p.z);
      }
      if (idx < 0 && !deferred) idx = brute_nearest(cent, F, p.x, p.y, p.z);  // outside the table / far from the canonical mesh (rare for warped points)
      out[t] = idx;
      if (gc.debug) {  // profile bit 2: which path the lookups take (dsnerf_debug_table [13] list, [14] scan, [15] exhaustive)
        const int c2 = cell < 0 ? -1 : gc.trec[cell].y;
        atomicAdd(gc.pool_used + (c2 >= 0 ? 13 : (c2 == -2 ? 14 : 15)), 1);
        if (c2 > 0) atomicAdd(gc.pool_used + 10, c2);
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, deferred);
    if (m) {
      int base = 0;
      if (lane == 0) base = atomicAdd(gc.pool_used + 19, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (deferred) queue[base + __popc(m & ((1u << lane) - 1))] = (int)t;
    }
  }
  }
}

__global__ void __launch_bounds__(256, 4) canon_long_kernel(const float4* __restrict__ active, Grid gc, const int* __restrict__ queue, int* __restrict__ out) {
  const int nq = gc.pool_used[19];
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nq; i += warps) {
    const int t = queue[i];
    const float4 p = active[t];
    const float px = p.x, py = p.y, pz = p.z;
    const int2 rec = gc.trec[live_cell(gc, px, py, pz)];
    const int besti = warp_nearest(gc, rec, px, py, pz);
    if (lane == 0) out[t] = besti;
  }
}

__device__ __forceinline__ void shade_inputs(const ShadeArgs& a, const Grid& gc, float4 ac, float4 mg, float (&in)[9], int& sample, int cidx = -1) {
  sample = __float_as_int(ac.w);
  const int idx = cidx >= 0 ? cidx : canon_nearest(gc, a.cent_canon, a.F, ac.x, ac.y, ac.z);
  V3 nw = v3(0.f, 0.f, 0.f);
  {
    // normalize(M_idx g) with F.normalize's eps; the scale of g is removed first (range), g = 0 stays 0
    const float sc = fmaxf(fmaxf(fabsf(mg.x), fabsf(mg.y)), fabsf(mg.z));
    if (sc > 0.f) {
      const float is = 1.0f / sc;
      const float gx = mg.x * is, gy = mg.y * is, gz = mg.z * is;
      const float4 r0 = __ldg(a.normal_m + 3 * idx), r1 = __ldg(a.normal_m + 3 * idx + 1), r2 = __ldg(a.normal_m + 3 * idx + 2);
      const float rx = r0.x * gx + r0.y * gy + r0.z * gz, ry = r1.x * gx + r1.y * gy + r1.z * gz, rz = r2.x * gx + r2.y * gy + r2.z * gz;
      const float n = sqrtf(rx * rx + ry * ry + rz * rz);
      const float in_ = 1.0f / fmaxf(n, 1e-12f);
      nw = v3(rx * in_, ry * in_, rz * in_);
    }
  }
  float px, py, pz, dx, dy, dz;
  if (a.xyz_world) {
    px = a.xyz_world[3 * (int64_t)sample]; py = a.xyz_world[3 * (int64_t)sample + 1]; pz = a.xyz_world[3 * (int64_t)sample + 2];
    dx = a.view_dir[3 * (int64_t)sample]; dy = a.view_dir[3 * (int64_t)sample + 1]; dz = a.view_dir[3 * (int64_t)sample + 2];
  } else {
    int64_t r = sample / a.N;
    int i = sample - (int)(r * a.N);
    float z = a.z_in ? a.z_in[sample] : sample_z(a.near[r], a.far[r], a.tvals[i]);
    dx = a.ray_d[3 * r]; dy = a.ray_d[3 * r + 1]; dz = a.ray_d[3 * r + 2];
    px = xadd(a.ray_o[3 * r], xmul(dx, z)); py = xadd(a.ray_o[3 * r + 1], xmul(dy, z)); pz = xadd(a.ray_o[3 * r + 2], xmul(dz, z));
  }
  if (a.has_rot) {  // model/spacenet.py:254-258: xy <- (xy - c) @ rot + c
    float qx = px - a.rot_center[0], qy = py - a.rot_center[1];
    px = qx * a.rot[0] + qy * a.rot[2] + a.rot_center[0];
    py = qx * a.rot[1] + qy * a.rot[3] + a.rot_center[1];
  }
  if (a.has_shift) { px += a.light_shift[0]; py += a.light_shift[1]; pz += a.light_shift[2]; }  // :260-263
  float dn = xnorm3(v3(dx, dy, dz));
  in[0] = nw.x; in[1] = nw.y; in[2] = nw.z; in[3] = px; in[4] = py; in[5] = pz;
  in[6] = xdiv(dx, dn); in[7] = xdiv(dy, dn); in[8] = xdiv(dz, dn);
}

// fp32 SIMT shading (verification path, flag DSNERF_MLP_FP32_SIMT).  One thread per active sample; the lighting
// layers run out of shared memory; the 128 hidden units of the first layer are recomputed per 32-wide output chunk
// instead of being staged, which keeps the block at ~71 KB of shared memory (3 blocks = 24 warps per SM).
__global__ void __launch_bounds__(SHADE_THREADS, 3) shade_kernel(ShadeArgs a, LightWeights L, Grid gc) {
  extern __shared__ __align__(16) float sm[];
  float* w1p = sm;                 // [128][12]
  float* w2t = w1p + 128 * 12;     // [128][128]
  float* b2 = w2t + 128 * 128;     // 128
  float* w3 = b2 + 128;            // 128
  for (int i = threadIdx.x; i < 128 * 12; i += SHADE_THREADS) {
    int k = i / 12, j = i - k * 12;
    w1p[i] = j < 9 ? L.w1t[j * 128 + k] : (j == 9 ? L.b1[k] : 0.f);
  }
  for (int i = threadIdx.x; i < 128 * 128; i += SHADE_THREADS) w2t[i] = L.w2t[i];
  for (int i = threadIdx.x; i < 128; i += SHADE_THREADS) { b2[i] = L.b2[i]; w3[i] = L.w3[i]; }
  __syncthreads();
  int64_t n_active = a.n_active ? (int64_t)*a.n_active : a.n_active_host;
  for (int64_t t = (int64_t)blockIdx.x * SHADE_THREADS + threadIdx.x; t < n_active; t += (int64_t)gridDim.x * SHADE_THREADS) {
    float in[9];
    const float4 ma = a.mlp_a[t];
    int sample;
    shade_inputs(a, gc, a.active[t], a.mlp_g[t], in, sample);
    // LightingMLP (model/spacenet.py:165-188): 9 -> 128 -> 128 -> 1, ReLU, ReLU, ELU; color = (out+1)*essence
    float out = L.b3;
#pragma unroll 1
    for (int jb = 0; jb < 128; jb += 32) {
      float acc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) acc[j] = b2[jb + j];
#pragma unroll 2
      for (int k = 0; k < 128; ++k) {
        const float4* w1 = reinterpret_cast<const float4*>(w1p + k * 12);
        float4 wa = w1[0], wb = w1[1], wc = w1[2];
        float h = wc.y;
        h = fmaf(in[0], wa.x, h); h = fmaf(in[1], wa.y, h); h = fmaf(in[2], wa.z, h); h = fmaf(in[3], wa.w, h);
        h = fmaf(in[4], wb.x, h); h = fmaf(in[5], wb.y, h); h = fmaf(in[6], wb.z, h); h = fmaf(in[7], wb.w, h);
        h = fmaf(in[8], wc.x, h);
        h = fmaxf(h, 0.f);
        const float4* wr = reinterpret_cast<const float4*>(w2t + k * 128 + jb);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float4 w = wr[j4];
          acc[4 * j4] = fmaf(h, w.x, acc[4 * j4]); acc[4 * j4 + 1] = fmaf(h, w.y, acc[4 * j4 + 1]);
          acc[4 * j4 + 2] = fmaf(h, w.z, acc[4 * j4 + 2]); acc[4 * j4 + 3] = fmaf(h, w.w, acc[4 * j4 + 3]);
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) out = fmaf(fmaxf(acc[j], 0.f), w3[jb + j], out);
    }
    float light = (out > 0.f ? out : expm1f(out)) + 1.0f;
    a.raw[sample] = make_float4(light * ma.y, light * ma.z, light * ma.w, ma.x);
  }
}

// density-only variant (Renderer.query_volume): raw.w = sigma
__global__ void scatter_density_kernel(const float4* __restrict__ active, const float4* __restrict__ mlp_a, int64_t n, float* __restrict__ density) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) density[__float_as_int(active[t].w)] = mlp_a[t].x;
}

// utils/nerf_net_utils.py:5-56 raw2outputs (raw_noise_std = 0, white_bkgd = False).
// One warp per ray; 32 samples per step with an in-register exclusive product scan for the
// transmittance and a running carry across steps.
struct CompositeArgs {
  const float4* raw;       // (R,N) rgb+sigma
  const float* ray_d;      // (R,3)
  const float* near; const float* far; const float* tvals;  // z = near(1-t)+far t   (z_in == NULL)
  const float* z_in;       // explicit (R,N) z
  const unsigned* sample_mask; // bit (s & 31) of word s >> 5 set <=> raw[s] was written (s = r*N + i); NULL => all
  int64_t R; int N;
  float* rgb; float* depth; float* acc; float* disp; float* weights; float* z_out;
  // training mode (utils/nerf_net_utils.py:29-33): noise (R,N) = randn * raw_noise_std is added to the density before the
  // ReLU.  all_raw: raw holds EVERY sample and a clear mask bit only zeroes the density (can_render.py:118-120).
  const float* noise; int all_raw;
  // fused all-gather (dsnerf_render_gather): the per-ray outputs are ALSO stored into this rank's block
  // [rgb (R,3) | depth (R) | acc (R) | disp (R)] of every peer GPU's frame buffer (peer-mapped pointers over NVLink), or once
  // through the NVSwitch multicast address `mc` (multimem.st: the switch replicates the store to every GPU of the group).
  int n_peers; float* peer[7]; float* mc;
};

constexpr int COMP_RAYS = 8;   // rays per block, one warp each (32-ray blocks measured 0.158 ms per frame against 0.085 ms: three of four rays are empty and
                               // their warps then sit behind the block barrier)

// one ray by one warp; the six outputs are valid in lane 0
__device__ __forceinline__ void composite_ray(const CompositeArgs& a, int64_t r, int lane, float (&o)[6]) {
  float nd = xnorm3(v3(a.ray_d[3 * r], a.ray_d[3 * r + 1], a.ray_d[3 * r + 2]));
  float near = a.z_in ? 0.f : a.near[r], far = a.z_in ? 0.f : a.far[r];
  if (a.sample_mask && !a.z_in && !a.noise && !a.all_raw) {
    // A ray without a single evaluated sample (3 of 4 rays of a frame): every density is 0, so alpha = 1 - exp(-0 * dist) = 0,
    // T = 1 and all sums are exactly 0 (disp = 0/0 = NaN) whenever the distances are finite -- written without the scan.
    const int64_t b0 = r * a.N, b1 = b0 + a.N;  // bit range of the ray in the mask
    bool any = false;
    for (int64_t wd = (b0 >> 5) + lane; wd <= ((b1 - 1) >> 5); wd += 32) {
      unsigned m = a.sample_mask[wd];
      if (wd == (b0 >> 5)) m &= ~0u << (b0 & 31);
      if (wd == ((b1 - 1) >> 5) && (b1 & 31)) m &= ~0u >> (32 - (b1 & 31));
      any |= m != 0u;
    }
    const float span = fabsf(near) + fabsf(far) + nd;  // finite <=> near, far and |d| are
    if (!__any_sync(0xffffffffu, any) && span < 3.0e38f) {
      for (int i = lane; i < a.N; i += 32) {
        if (a.weights) a.weights[b0 + i] = 0.f;
        if (a.z_out) a.z_out[b0 + i] = sample_z(near, far, a.tvals[i]);
      }
      o[0] = o[1] = o[2] = o[3] = o[4] = 0.f;
      o[5] = __int_as_float(0x7fc00000);
      return;
    }
  }
  float T = 1.0f;
  float sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
  for (int base = 0; base < a.N; base += 32) {
    int i = base + lane;
    bool live = i < a.N;
    float z = 0.f, zn = 0.f;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      z = a.z_in ? a.z_in[r * a.N + i] : sample_z(near, far, a.tvals[i]);
      if (i + 1 < a.N) zn = a.z_in ? a.z_in[r * a.N + i + 1] : sample_z(near, far, a.tvals[i + 1]);
      const int64_t sidx = r * a.N + i;
      bool has = a.sample_mask ? ((a.sample_mask[sidx >> 5] >> (sidx & 31)) & 1u) : true;
      if (has || a.all_raw) c = a.raw[r * a.N + i];
      if (!has) c.w = 0.f;
      if (a.noise) c.w = xadd(c.w, a.noise[sidx]);
    }
    float dist = (i + 1 < a.N) ? xsub(zn, z) : 1e10f;
    dist = xmul(dist, nd);
    float alpha = live ? xsub(1.0f, expf(-xmul(fmaxf(c.w, 0.f), dist))) : 0.f;
    float t = xadd(xsub(1.0f, alpha), 1e-10f);
    // inclusive product scan
    float p = t;
#pragma unroll
    for (int o2 = 1; o2 < 32; o2 <<= 1) {
      float q = __shfl_up_sync(0xffffffffu, p, o2);
      if (lane >= o2) p *= q;
    }
    float excl = __shfl_up_sync(0xffffffffu, p, 1);
    if (lane == 0) excl = 1.0f;
    float w = alpha * (T * excl);
    T *= __shfl_sync(0xffffffffu, p, 31);
    if (live) {
      sr = fmaf(w, c.x, sr); sg = fmaf(w, c.y, sg); sb = fmaf(w, c.z, sb);
      sd = fmaf(w, z, sd); sa += w;
      if (a.weights) a.weights[r * a.N + i] = w;
      if (a.z_out) a.z_out[r * a.N + i] = z;
    }
  }
#pragma unroll
  for (int o2 = 16; o2 > 0; o2 >>= 1) {
    sr += __shfl_xor_sync(0xffffffffu, sr, o2); sg += __shfl_xor_sync(0xffffffffu, sg, o2); sb += __shfl_xor_sync(0xffffffffu, sb, o2);
    sd += __shfl_xor_sync(0xffffffffu, sd, o2); sa += __shfl_xor_sync(0xffffffffu, sa, o2);
  }
  o[0] = sr; o[1] = sg; o[2] = sb; o[3] = sd; o[4] = sa;
  float q = xdiv(sd, sa);  // 0/0 = NaN when nothing was hit, as in the reference
  o[5] = (q != q) ? q : xdiv(1.0f, fmaxf(1e-10f, q));
}

// A block composites COMP_RAYS consecutive rays (one warp each), stages their six outputs in shared memory and writes them
// with coalesced stores: 3 C contiguous floats of rgb and C each of depth / acc / disp per destination (C = COMP_RAYS).  With
// gather targets every destination receives 96- and 32-byte segments over NVLink instead of 4-byte scattered stores.
__global__ void __launch_bounds__(COMP_RAYS * 32) composite_kernel(CompositeArgs a) {
  __shared__ float s_out[6 * COMP_RAYS];  // [0, 3C) rgb as (ray,3); then C floats each of depth, acc, disp (C = COMP_RAYS)
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r0 = (int64_t)blockIdx.x * COMP_RAYS;
  const int64_t r = r0 + w;
  if (r < a.R) {
    float o[6];
    composite_ray(a, r, lane, o);
    if (lane == 0) {
      s_out[3 * w] = o[0]; s_out[3 * w + 1] = o[1]; s_out[3 * w + 2] = o[2];
      s_out[3 * COMP_RAYS + w] = o[3]; s_out[4 * COMP_RAYS + w] = o[4]; s_out[5 * COMP_RAYS + w] = o[5];
    }
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t >= 6 * COMP_RAYS) return;
  const int nr = (int)min((int64_t)COMP_RAYS, a.R - r0);
  const int ch = t < 3 * COMP_RAYS ? 0 : t / COMP_RAYS - 2;       // 0 rgb, 1 depth, 2 acc, 3 disp
  const int k = t < 3 * COMP_RAYS ? t : t % COMP_RAYS;            // element inside the block's run of that channel
  if (k >= (ch == 0 ? 3 * nr : nr)) return;
  const float v = s_out[t];
  float* const loc = ch == 0 ? a.rgb : (ch == 1 ? a.depth : (ch == 2 ? a.acc : a.disp));
  const int64_t e = (ch == 0 ? 3 * r0 : r0) + k;    // element inside the channel
  loc[e] = v;
  const int64_t eb = (ch == 0 ? 0 : (int64_t)(2 + ch) * a.R) + e;  // element inside a [rgb | depth | acc | disp] block
  if (a.mc) {
    asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(a.mc + eb), "f"(v) : "memory");
  } else {
    for (int p = 0; p < a.n_peers; ++p) a.peer[p][eb] = v;
  }
}

// ---- early ray termination (optional, flag DSNERF_EARLY_STOP) ---------------------------------------------------------
// The samples of a ray are processed in `waves` index ranges, front to back.  After wave k the transmittance T of every ray is
// known; since sum_{i >= cut} w_i <= T_cut, a ray with T <= tau can gain at most tau in acc, tau*max|c| in colour and
// tau*z_far in depth from all its remaining samples, so they are not evaluated at all (the reference evaluates them and adds
// those < tau amounts).  tau = 1e-6 keeps every output within 4e-6 of the exhaustive result, 25x inside the 1e-4 tolerance.
struct RayState { float T, r, g, b, d, a; };  // running transmittance and sums of one ray

// wave k > 0: keep the entries of its region whose ray is still alive (compaction into `out`)
__global__ void __launch_bounds__(256) filter_wave_kernel(const float4* __restrict__ in, const int* __restrict__ in_tri,
                                                          const unsigned long long* __restrict__ n_in, const RayState* __restrict__ state, int N,
                                                          float tau, float4* __restrict__ out, int* __restrict__ out_tri,
                                                          unsigned long long* __restrict__ n_out) {
  const int64_t n = (int64_t)*n_in;
  const int lane = threadIdx.x & 31;
  for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x; t0 < n; t0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = t0 + threadIdx.x;
    bool keep = false;
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < n) {
      e = in[t];
      keep = state[__float_as_int(e.w) / N].T > tau;
    }
    unsigned m = __ballot_sync(0xffffffffu, keep);
    if (!m) continue;
    int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(n_out, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (keep) {
      unsigned long long slot = base + __popc(m & ((1u << lane) - 1));
      out[slot] = e;
      out_tri[slot] = in_tri[t];
    }
  }
}

// raw2outputs restricted to samples [i0, i1) of every ray (one warp per ray, same arithmetic as composite_kernel), carrying
// the ray state across waves.  A ray that was dead when the wave started (state.T <= tau) is skipped: its samples of this wave
// were not evaluated.  The last wave writes the outputs.
__global__ void __launch_bounds__(256) composite_wave_kernel(CompositeArgs a, RayState* __restrict__ state, int i0, int i1, float tau, int last) {
  int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= a.R) return;
  RayState st;
  if (i0 == 0) { st.T = 1.f; st.r = st.g = st.b = st.d = st.a = 0.f; }
  else st = state[r];
  const bool alive = st.T > tau;
  float nd = xnorm3(v3(a.ray_d[3 * r], a.ray_d[3 * r + 1], a.ray_d[3 * r + 2]));
  float near = a.z_in ? 0.f : a.near[r], far = a.z_in ? 0.f : a.far[r];
  float T = st.T;
  float sr = 0.f, sg = 0.f, sb = 0.f, sd = 0.f, sa = 0.f;
  for (int base = i0; base < i1; base += 32) {
    int i = base + lane;
    bool live = i < i1;
    float z = 0.f, zn = 0.f;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      z = a.z_in ? a.z_in[r * a.N + i] : sample_z(near, far, a.tvals[i]);
      if (i + 1 < a.N) zn = a.z_in ? a.z_in[r * a.N + i + 1] : sample_z(near, far, a.tvals[i + 1]);
      const int64_t sidx = r * a.N + i;
      bool has = alive && ((a.sample_mask[sidx >> 5] >> (sidx & 31)) & 1u);
      if (has) c = a.raw[sidx];
    }
    float dist = (i + 1 < a.N) ? xsub(zn, z) : 1e10f;
    dist = xmul(dist, nd);
    float alpha = live ? xsub(1.0f, expf(-xmul(fmaxf(c.w, 0.f), dist))) : 0.f;
    float t = xadd(xsub(1.0f, alpha), 1e-10f);
    float p = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      float q = __shfl_up_sync(0xffffffffu, p, o);
      if (lane >= o) p *= q;
    }
    float excl = __shfl_up_sync(0xffffffffu, p, 1);
    if (lane == 0) excl = 1.0f;
    float w = alpha * (T * excl);
    T *= __shfl_sync(0xffffffffu, p, 31);
    if (live) {
      sr = fmaf(w, c.x, sr); sg = fmaf(w, c.y, sg); sb = fmaf(w, c.z, sb);
      sd = fmaf(w, z, sd); sa += w;
      if (a.weights) a.weights[r * a.N + i] = w;
      if (a.z_out) a.z_out[r * a.N + i] = z;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sr += __shfl_xor_sync(0xffffffffu, sr, o); sg += __shfl_xor_sync(0xffffffffu, sg, o); sb += __shfl_xor_sync(0xffffffffu, sb, o);
    sd += __shfl_xor_sync(0xffffffffu, sd, o); sa += __shfl_xor_sync(0xffffffffu, sa, o);
  }
  if (lane == 0) {
    st.T = alive ? T : st.T;
    st.r += sr; st.g += sg; st.b += sb; st.d += sd; st.a += sa;
    if (!last) {
      state[r] = st;
    } else {
      a.rgb[3 * r] = st.r; a.rgb[3 * r + 1] = st.g; a.rgb[3 * r + 2] = st.b;
      a.depth[r] = st.d; a.acc[r] = st.a;
      float q = xdiv(st.d, st.a);  // 0/0 = NaN when nothing was hit, as in the reference
      a.disp[r] = (q != q) ? q : xdiv(1.0f, fmaxf(1e-10f, q));
    }
  }
}

// Hierarchical resampling (config 3).  The reference calls an undefined Renderer.resampling
// (can_render.py:213); this implements the written spec in DESIGN.md "Config 3" =
// oracle/oracle.py:sample_pdf: deterministic inverse-CDF sampling of the coarse weights
// (NeRF sample_pdf, det=True) merged with the coarse z and sorted.  One warp per ray.
__global__ void __launch_bounds__(128) resample_kernel(const float* __restrict__ z_in, const float* __restrict__ weights, int64_t R, int N,
                                                       int n_imp, float* __restrict__ z_out) {
  extern __shared__ float rs[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t r = (int64_t)blockIdx.x * 4 + warp;
  if (r >= R) return;
  int per = 2 * N + n_imp;
  float* bins = rs + warp * per;   // N-1
  float* cdf = bins + N;           // N-1
  float* zn = cdf + N;             // n_imp
  const float* z = z_in + r * N;
  const float* w = weights + r * N;
  float part = 0.f;
  for (int j = lane; j < N - 1; j += 32) bins[j] = 0.5f * (z[j + 1] + z[j]);
  for (int j = lane; j < N - 2; j += 32) part += w[j + 1] + 1e-5f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  __syncwarp();
  if (lane == 0) {
    float c = 0.f;
    cdf[0] = 0.f;
    for (int j = 0; j < N - 2; ++j) { c += (w[j + 1] + 1e-5f) / part; cdf[j + 1] = c; }
  }
  __syncwarp();
  int len = N - 1;
  for (int k = lane; k < n_imp; k += 32) {
    float u = linspace01(k, n_imp);
    int lo = 0, hi = len;  // first index with cdf > u
    while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
    int below = max(lo - 1, 0), above = min(lo, len - 1);
    float c0 = cdf[below], c1 = cdf[above], b0 = bins[below], b1 = bins[above];
    float den = c1 - c0;
    if (den < 1e-5f) den = 1.0f;
    zn[k] = b0 + (u - c0) / den * (b1 - b0);
  }
  __syncwarp();
  float* out = z_out + r * (N + n_imp);
  for (int i = lane; i < N; i += 32) {  // coarse z: rank = i + #(zn < z[i])
    float v = z[i];
    int lo = 0, hi = n_imp;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (zn[mid] < v) lo = mid + 1; else hi = mid; }
    out[i + lo] = v;
  }
  for (int k = lane; k < n_imp; k += 32) {  // new z: rank = k + #(z <= zn[k])
    float v = zn[k];
    int lo = 0, hi = N;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (z[mid] <= v) lo = mid + 1; else hi = mid; }
    out[k + lo] = v;
  }
}

// utils/blend_utils.py:72-81 ppts_to_pts: x_c = R^-1 (x_p - t) with [R|t] = sum_j bw[j] A_j (no caller in the reference;
// SURVEY.md 8a #23).  One thread per point; HBM bound: 27 floats in (3 + 24 weights, coalesced along p), 3 out = 120 B/point;
// the 24 joint transforms (1.5 KB) are staged in shared memory.  The blended 3x3 is inverted through its cofactors.
__global__ void __launch_bounds__(256) lbs_inverse_kernel(const float* __restrict__ ppts, const float* __restrict__ bw,
                                                          const float* __restrict__ A, int64_t P, float* __restrict__ out) {
  __shared__ float sA[24 * 12];  // rows 0..2 of each 4x4
  for (int i = threadIdx.x; i < 24 * 12; i += blockDim.x) sA[i] = A[(i / 12) * 16 + (i % 12)];
  __syncthreads();
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float m[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) m[k] = 0.f;
#pragma unroll 4
  for (int j = 0; j < 24; ++j) {
    const float w = __ldg(bw + (int64_t)j * P + p);
#pragma unroll
    for (int k = 0; k < 12; ++k) m[k] = fmaf(w, sA[j * 12 + k], m[k]);
  }
  const float dx = ppts[3 * p] - m[3], dy = ppts[3 * p + 1] - m[7], dz = ppts[3 * p + 2] - m[11];
  // cofactors of R = [m0 m1 m2; m4 m5 m6; m8 m9 m10]
  const float c00 = m[5] * m[10] - m[6] * m[9], c01 = m[6] * m[8] - m[4] * m[10], c02 = m[4] * m[9] - m[5] * m[8];
  const float c10 = m[2] * m[9] - m[1] * m[10], c11 = m[0] * m[10] - m[2] * m[8], c12 = m[1] * m[8] - m[0] * m[9];
  const float c20 = m[1] * m[6] - m[2] * m[5], c21 = m[2] * m[4] - m[0] * m[6], c22 = m[0] * m[5] - m[1] * m[4];
  const float inv = 1.0f / (m[0] * c00 + m[1] * c01 + m[2] * c02);  // singular blend -> inf / NaN (torch.inverse raises)
  out[3 * p] = (c00 * dx + c10 * dy + c20 * dz) * inv;
  out[3 * p + 1] = (c01 * dx + c11 * dy + c21 * dz) * inv;
  out[3 * p + 2] = (c02 * dx + c12 * dy + c22 * dz) * inv;
}

// Camera rays on the device: utils/rays_utils.py:16-30 get_rays and :63-97 get_near_far as the
// inference branch of my_sample_ray (:173-189) chains them.  The reference works in float64 (numpy) and casts to float32
// at the end; so does this kernel, in the reference's operation order, which makes directions / near / far agree to the
// last float32 bit except for double-rounding ties.  One thread per pixel; HBM bound: 33 B written per pixel.
struct CameraArgs {
  double kinv[9], rot[9], t[3], origin[3];  // K^-1, R, T, -R^T T
  double lo[3], hi[3];                      // bounds already padded by -/+ 0.01
  int H, W;
};
__global__ void __launch_bounds__(256) camera_rays_kernel(CameraArgs c, float* __restrict__ ray_o, float* __restrict__ ray_d,
                                                          float* __restrict__ near, float* __restrict__ far, uint8_t* __restrict__ mask) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (int64_t)c.H * c.W) return;
  const double i = (double)(int)(p % c.W), j = (double)(int)(p / c.W);
  double cam[3], dir[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) cam[k] = __dadd_rn(__dadd_rn(__dmul_rn(i, c.kinv[3 * k]), __dmul_rn(j, c.kinv[3 * k + 1])), c.kinv[3 * k + 2]) - c.t[k];
#pragma unroll
  for (int k = 0; k < 3; ++k)
    dir[k] = __dadd_rn(__dadd_rn(__dmul_rn(cam[0], c.rot[k]), __dmul_rn(cam[1], c.rot[3 + k])), __dmul_rn(cam[2], c.rot[6 + k])) - c.origin[k];
  const float of[3] = {(float)c.origin[0], (float)c.origin[1], (float)c.origin[2]};
  const float df[3] = {(float)dir[0], (float)dir[1], (float)dir[2]};
#pragma unroll
  for (int k = 0; k < 3; ++k) { ray_o[3 * p + k] = of[k]; ray_d[3 * p + k] = df[k]; }
  // get_near_far on the float32 rays, in float64
  const double o[3] = {of[0], of[1], of[2]}, d[3] = {df[0], df[1], df[2]};
  const double eps = 1e-6;
  int hits = 0;
  double h0[3] = {0, 0, 0}, h1[3] = {0, 0, 0};
#pragma unroll
  for (int m = 0; m < 6; ++m) {
    const double plane = m < 3 ? c.lo[m] : c.hi[m - 3];
    const double t = (plane - o[m % 3]) / d[m % 3];
    double q[3];
    bool in = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      q[k] = __dadd_rn(__dmul_rn(t, d[k]), o[k]);
      in = in && (q[k] >= c.lo[k] - eps) && (q[k] <= c.hi[k] + eps);
    }
    if (in) {
      if (hits == 0) { h0[0] = q[0]; h0[1] = q[1]; h0[2] = q[2]; }
      else if (hits == 1) { h1[0] = q[0]; h1[1] = q[1]; h1[2] = q[2]; }
      ++hits;
    }
  }
  const bool ok = hits == 2;
  mask[p] = ok ? 1 : 0;
  float n = 0.f, f = 0.f;
  if (ok) {
    auto norm3 = [](double x, double y, double z) { return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z))); };
    // np.linalg.norm of the float32 directions stays in float32 (rays_utils.py:90)
    const double nd = (double)__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(df[0], df[0]), __fmul_rn(df[1], df[1])), __fmul_rn(df[2], df[2])));
    const double d0 = norm3(h0[0] - o[0], h0[1] - o[1], h0[2] - o[2]) / nd, d1 = norm3(h1[0] - o[0], h1[1] - o[1], h1[2] - o[2]) / nd;
    n = (float)fmin(d0, d1);
    f = (float)fmax(d0, d1);
  }
  near[p] = n;
  far[p] = f;
}

}  // namespace dsn
