// Geometry kernels: GG ray bounds, sample placement, exact nearest-centroid
// search on a uniform grid, barycentric warp posed -> canonical.
//
// Every value that feeds a DISCRETE decision of the reference (nearest-triangle
// argmin, transparent mask, GG hit test) is computed with explicitly rounded
// IEEE fp32 operations in the reference's own order, so the decisions are
// bit-identical to the reference's torch ops (see DESIGN.md "Exact geometry").
// nvcc never contracts the __f*_rn intrinsics into FMAs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dsn {

__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float xsqrt(float a) { return __fsqrt_rn(a); }

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 xsub3(V3 a, V3 b) { return v3(xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z)); }
// torch.cross: a_i*b_j - a_j*b_i evaluated as fma(a_i, b_j, -(a_j*b_i))
__device__ __forceinline__ V3 xcross(V3 a, V3 b) {
  return v3(xfma(a.y, b.z, -xmul(a.z, b.y)), xfma(a.z, b.x, -xmul(a.x, b.z)), xfma(a.x, b.y, -xmul(a.y, b.x)));
}
// torch.norm over 3 elements: fma-accumulated sum of squares
__device__ __forceinline__ float xnorm3(V3 a) { return xsqrt(xfma(a.z, a.z, xfma(a.y, a.y, xmul(a.x, a.x)))); }
// (a*b).sum(-1) / einsum('ij,ij->i'): products rounded, summed left to right
__device__ __forceinline__ float xdot3(V3 a, V3 b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }

__device__ __forceinline__ V3 ldv3(const float* __restrict__ p, int i) { return v3(__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2)); }

// utils/geo_utils.py:181-200 project_point2mesh + :96-113 get_barycentric_coordinates
__device__ __forceinline__ void project_point(V3 p, V3 m0, V3 m1, V3 m2, float& u, float& v, float& h) {
  V3 v10 = xsub3(m1, m0), v20 = xsub3(m2, m0);
  V3 n = xcross(v10, v20);
  float nn = xnorm3(n);
  n = v3(xdiv(n.x, nn), xdiv(n.y, nn), xdiv(n.z, nn));
  V3 t = xsub3(p, m0);
  float sd = xdot3(t, n);
  V3 proj = v3(xsub(p.x, xmul(n.x, sd)), xsub(p.y, xmul(n.y, sd)), xsub(p.z, xmul(n.z, sd)));
  V3 w = xsub3(proj, m0);
  float d00 = xdot3(v20, v20), d01 = xdot3(v20, v10), d02 = xdot3(v20, w);
  float d11 = xdot3(v10, v10), d12 = xdot3(v10, w);
  float inv = xdiv(1.0f, xsub(xmul(d00, d11), xmul(d01, d01)));
  u = xmul(xsub(xmul(d11, d02), xmul(d01, d12)), inv);
  v = xmul(xsub(xmul(d00, d12), xmul(d01, d02)), inv);
  h = sd;
}

// utils/render_utils.py:103-109 get_transparent_mask (NaN compares false, as in torch)
__device__ __forceinline__ bool is_transparent(float u, float v, float h) {
  return (u > 5.0f) || (u < -4.0f) || (v > 5.0f) || (v < -4.0f) || (fabsf(h) > 0.1f);
}

// utils/geo_utils.py:138-156 barycentric_map2can
__device__ __forceinline__ V3 map_to_triangle(float u, float v, float h, V3 c0, V3 c1, V3 c2) {
  V3 e2 = xsub3(c2, c0), e1 = xsub3(c1, c0);
  V3 n = xcross(e1, e2);
  float nn = xnorm3(n);
  V3 r;
  r.x = xadd(xadd(xadd(c0.x, xmul(u, e2.x)), xmul(v, e1.x)), xmul(h, xdiv(n.x, nn)));
  r.y = xadd(xadd(xadd(c0.y, xmul(u, e2.y)), xmul(v, e1.y)), xmul(h, xdiv(n.y, nn)));
  r.z = xadd(xadd(xadd(c0.z, xmul(u, e2.z)), xmul(v, e1.z)), xmul(h, xdiv(n.z, nn)));
  return r;
}

// ---------------------------------------------------------------------------------------------
// Uniform grid over triangle centroids (one per mesh: posed = per frame, canonical = static).
// Cells are x-fastest, so a run of cells along x is one contiguous run of sorted centroids.
struct Grid {
  float ox, oy, oz;     // origin of cell (0,0,0)
  float cell, inv_cell; // edge length
  int nx, ny, nz;
  float half_diag;      // cell * sqrt(3)/2 (rounded up)
  float r_cap;          // beyond this distance to the nearest centroid a point is provably transparent
  const int* __restrict__ cell_start;    // ncell+1
  const float4* __restrict__ sorted;     // (x,y,z,bits(idx)) sorted by cell
  const float* __restrict__ cent;        // (F,3) centroids by index
  const float* __restrict__ center_dist; // per cell: distance from the cell centre to its nearest centroid (huge = provably transparent)
  const int* __restrict__ center_idx;    // per cell: index of that centroid (search seed)
};

__global__ void centroid_kernel(const float* __restrict__ verts, const int* __restrict__ faces, int F, float* __restrict__ cent,
                                float4* __restrict__ tri_n) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  V3 a = ldv3(verts, faces[3 * f]), b = ldv3(verts, faces[3 * f + 1]), c = ldv3(verts, faces[3 * f + 2]);
  // meshes.mean(dim=-2): ((a+b)+c)/3  (utils/render_utils.py:94)
  cent[3 * f] = xdiv(xadd(xadd(a.x, b.x), c.x), 3.0f);
  cent[3 * f + 1] = xdiv(xadd(xadd(a.y, b.y), c.y), 3.0f);
  cent[3 * f + 2] = xdiv(xadd(xadd(a.z, b.z), c.z), 3.0f);
  // unit normal, only used by the conservative cell classification below (NaN for degenerate triangles)
  V3 n = xcross(xsub3(b, a), xsub3(c, a));
  float nn = xnorm3(n);
  tri_n[f] = make_float4(n.x / nn, n.y / nn, n.z / nn, 0.f);
}

__device__ __forceinline__ int grid_coord(float p, float o, float inv, int n) {
  int c = (int)floorf((p - o) * inv);
  return c < 0 ? 0 : (c >= n ? n - 1 : c);
}

__global__ void grid_count_kernel(Grid g, const float* __restrict__ cent, int F, int* __restrict__ counts) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  int cx = grid_coord(cent[3 * f], g.ox, g.inv_cell, g.nx);
  int cy = grid_coord(cent[3 * f + 1], g.oy, g.inv_cell, g.ny);
  int cz = grid_coord(cent[3 * f + 2], g.oz, g.inv_cell, g.nz);
  atomicAdd(&counts[(cz * g.ny + cy) * g.nx + cx], 1);
}

// single-block exclusive scan: counts[ncell] -> start[ncell+1]; also copies start into cursor
__global__ void grid_scan_kernel(const int* __restrict__ counts, int ncell, int* __restrict__ start, int* __restrict__ cursor) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < ncell; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = i < ncell ? counts[i] : 0;
    int s = v;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
    if (lane == 31) warp_sums[wid] = s;
    __syncthreads();
    if (wid == 0) {
      int w = lane < (blockDim.x >> 5) ? warp_sums[lane] : 0;
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      warp_sums[lane] = w;
    }
    __syncthreads();
    int excl = carry + (wid ? warp_sums[wid - 1] : 0) + s - v;
    if (i < ncell) { start[i] = excl; cursor[i] = excl; }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) start[ncell] = carry;
}

__global__ void grid_fill_kernel(Grid g, const float* __restrict__ cent, int F, int* __restrict__ cursor, float4* __restrict__ sorted) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  float x = cent[3 * f], y = cent[3 * f + 1], z = cent[3 * f + 2];
  int cx = grid_coord(x, g.ox, g.inv_cell, g.nx), cy = grid_coord(y, g.oy, g.inv_cell, g.ny), cz = grid_coord(z, g.oz, g.inv_cell, g.nz);
  int slot = atomicAdd(&cursor[(cz * g.ny + cy) * g.nx + cx], 1);
  sorted[slot] = make_float4(x, y, z, __int_as_float(f));
}

// Per cell of the lookup table: distance dc from the cell centre to its nearest centroid, or
// +huge when every point of the cell is PROVABLY transparent.  Proof: for p in the cell, its
// nearest centroid c* satisfies |centre - c*| <= dc + 2*half_diag (candidate set), the signed
// plane distance h* is 1-Lipschitz, so |h*(centre)| > 0.1 + half_diag for every candidate implies
// |h*(p)| > 0.1 = max_dist of get_transparent_mask (utils/render_utils.py:103).  Brute force over
// all centroids, tiled through shared memory: two sweeps (distance, then classification).
__global__ void grid_center_dist_kernel(Grid g, const float* __restrict__ cent, const float4* __restrict__ tri_n, int F, int classify,
                                        float* __restrict__ out, int* __restrict__ out_idx) {
  __shared__ float sx[1024], sy[1024], sz[1024], snx[1024], sny[1024], snz[1024];
  int ncell = g.nx * g.ny * g.nz;
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  int cx = c % g.nx, cy = (c / g.nx) % g.ny, cz = c / (g.nx * g.ny);
  float px = g.ox + (cx + 0.5f) * g.cell, py = g.oy + (cy + 0.5f) * g.cell, pz = g.oz + (cz + 0.5f) * g.cell;
  float best = 3.0e38f;
  int besti = 0;
  for (int base = 0; base < F; base += 1024) {
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
      int f = base + i;
      bool ok = f < F;
      sx[i] = ok ? cent[3 * f] : 1.0e18f;
      sy[i] = ok ? cent[3 * f + 1] : 1.0e18f;
      sz[i] = ok ? cent[3 * f + 2] : 1.0e18f;
    }
    __syncthreads();
#pragma unroll 8
    for (int i = 0; i < 1024; ++i) {
      float dx = px - sx[i], dy = py - sy[i], dz = pz - sz[i];
      float d2 = dx * dx + dy * dy + dz * dz;
      if (d2 < best) { best = d2; besti = base + i; }
    }
  }
  float dc = sqrtf(best) * 1.00001f + 1e-7f;
  const float h_thr = 0.1f + g.half_diag * 1.001f + 1e-4f;
  // the centre's own nearest centroid is a candidate with |h| <= dc: only the band needs the second sweep
  bool undecided = classify && (c < ncell) && (dc > h_thr) && (dc - g.half_diag <= g.r_cap);
  bool search = (c < ncell) && (!classify || dc <= h_thr);
  if (__syncthreads_or(undecided)) {
    float thr = dc + 2.0f * g.half_diag * 1.001f + 1e-5f;
    float thr2 = undecided ? thr * thr : -1.0f;
    for (int base = 0; base < F; base += 1024) {
      __syncthreads();
      for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        int f = base + i;
        bool ok = f < F;
        float4 n = ok ? tri_n[f] : make_float4(0.f, 0.f, 0.f, 0.f);
        sx[i] = ok ? cent[3 * f] : 1.0e18f;
        sy[i] = ok ? cent[3 * f + 1] : 1.0e18f;
        sz[i] = ok ? cent[3 * f + 2] : 1.0e18f;
        snx[i] = n.x; sny[i] = n.y; snz[i] = n.z;
      }
      __syncthreads();
#pragma unroll 4
      for (int i = 0; i < 1024; ++i) {
        float dx = px - sx[i], dy = py - sy[i], dz = pz - sz[i];
        float d2 = dx * dx + dy * dy + dz * dz;
        if (d2 <= thr2) {
          float h = dx * snx[i] + dy * sny[i] + dz * snz[i];
          if (!(fabsf(h) > h_thr)) search = true;  // NaN normal (degenerate triangle) keeps the cell searchable
        }
      }
    }
  }
  if (c < ncell) { out[c] = search ? dc : 3.0e30f; out_idx[c] = besti; }
}

// Exact nearest centroid: squared L2 accumulated as d0*d0, fma(d1,d1,.), fma(d2,d2,.) and
// strict '<' with lowest index on ties -- the arithmetic of pytorch3d 0.4.0 knn_points(K=1)
// as called at utils/render_utils.py:95.  Returns -1 when the point is provably farther than
// g.r_cap from every centroid (then it is transparent whatever its nearest triangle is).
// The search visits only grid rows that intersect the ball of the current best radius, which
// starts from the cell-centre distance table, so it returns the same index as a full scan.
// `hint` (>= 0) is any centroid index expected to be close (the previous sample's answer along a
// ray, or the posed-space triangle for the canonical search): its exact distance seeds the search
// radius, which only prunes -- the result is still the exact argmin.
__device__ __forceinline__ int nearest_centroid(const Grid& g, float px, float py, float pz, unsigned long long* cand_counter,
                                                int hint = -1, const float* __restrict__ cent = nullptr) {
  float fx = (px - g.ox) * g.inv_cell, fy = (py - g.oy) * g.inv_cell, fz = (pz - g.oz) * g.inv_cell;
  // points outside the table region are farther than r_cap from the mesh by construction
  if (!(fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)g.nx && fy < (float)g.ny && fz < (float)g.nz)) return -1;
  int hx = (int)fx, hy = (int)fy, hz = (int)fz;
  const int cell = (hz * g.ny + hy) * g.nx + hx;
  float dc = __ldg(g.center_dist + cell);
  if (hint < 0) {
    if (dc - g.half_diag > g.r_cap) return -1;
    // seed: the centroid nearest to the cell centre.  Its exact distance to p is at most dc + half_diag
    // (the table bound) and usually much less, which shrinks the ball that has to be scanned.
    hint = __ldg(g.center_idx + cell);
    cent = g.cent;
  } else if (dc > 1.0e29f) {
    dc = g.r_cap;  // classified cell but the caller vouches for a nearby centroid: let the hint set the radius
  }
  float rho = fminf(dc + g.half_diag, g.r_cap * 1.0001f + g.half_diag);
  float rho2 = rho * rho * 1.0001f;
  const float rho2_init = rho2;
  float best = 3.0e38f;
  int besti = -1;
  if (hint >= 0) {
    float dx = xsub(px, __ldg(cent + 3 * hint)), dy = xsub(py, __ldg(cent + 3 * hint + 1)), dz = xsub(pz, __ldg(cent + 3 * hint + 2));
    float d = xfma(dz, dz, xfma(dy, dy, xmul(dx, dx)));
    best = d;
    besti = hint;
    rho2 = fminf(rho2, best * 1.0001f + 1e-12f);
    rho = fminf(rho, sqrtf(rho2) * 1.0001f);
  }
  unsigned long long ncand = 0;
  int z0 = max(0, (int)floorf((pz - rho - g.oz) * g.inv_cell)), z1 = min(g.nz - 1, (int)floorf((pz + rho - g.oz) * g.inv_cell));
  int y0 = max(0, (int)floorf((py - rho - g.oy) * g.inv_cell)), y1 = min(g.ny - 1, (int)floorf((py + rho - g.oy) * g.inv_cell));
  for (int cz = z0; cz <= z1; ++cz) {
    float zl = g.oz + cz * g.cell;
    float dz = fmaxf(0.f, fmaxf(zl - pz, pz - (zl + g.cell)));
    float dz2 = dz * dz * 0.9999f;
    if (dz2 > rho2) continue;
    for (int cy = y0; cy <= y1; ++cy) {
      float yl = g.oy + cy * g.cell;
      float dy = fmaxf(0.f, fmaxf(yl - py, py - (yl + g.cell)));
      float rem = rho2 - dz2 - dy * dy * 0.9999f;
      if (rem < 0.f) continue;
      float rx = sqrtf(rem) * 1.0001f + 1e-6f;
      int x0 = max(0, (int)floorf((px - rx - g.ox) * g.inv_cell)), x1 = min(g.nx - 1, (int)floorf((px + rx - g.ox) * g.inv_cell));
      if (x0 > x1) continue;
      int row = (cz * g.ny + cy) * g.nx;
      int b = __ldg(g.cell_start + row + x0), e = __ldg(g.cell_start + row + x1 + 1);
      ncand += (unsigned)(e - b);
      for (int j = b; j < e; ++j) {
        float4 c = __ldg(g.sorted + j);
        float dx = xsub(px, c.x), dy2 = xsub(py, c.y), dzz = xsub(pz, c.z);
        float d = xmul(dx, dx);
        d = xfma(dy2, dy2, d);
        d = xfma(dzz, dzz, d);
        int id = __float_as_int(c.w);
        if (d < best || (d == best && id < besti)) {
          best = d;
          besti = id;
          rho2 = fminf(rho2, best * 1.0001f + 1e-12f);
        }
      }
    }
  }
  if (cand_counter && ncand) atomicAdd(cand_counter, ncand);
  // a candidate beyond the initial radius came from a partially covered cell: the true nearest
  // may sit in an unvisited one, but it is farther than r_cap either way
  if (best > rho2_init) besti = -1;
  return besti;
}

// ---------------------------------------------------------------------------------------------
// geometry_guided_ray_marching, utils/pts_utils.py:18-53 (near/far only).
// vq = (vertex - o0, |vertex - o0|^2) per vertex, prepared by gg_prep_kernel.
__global__ void gg_prep_kernel(const float* __restrict__ xyz, int V, const float* __restrict__ ray_o, float4* __restrict__ vq) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  // the reference uses ray_o[:, 0:1] -- the FIRST ray's origin -- for every ray (pts_utils.py:31,33)
  float qx = xsub(xyz[3 * v], ray_o[0]), qy = xsub(xyz[3 * v + 1], ray_o[1]), qz = xsub(xyz[3 * v + 2], ray_o[2]);
  float qq = xadd(xadd(xmul(qx, qx), xmul(qy, qy)), xmul(qz, qz));
  vq[v] = make_float4(qx, qy, qz, qq);
}

constexpr int GG_THREADS = 256;
constexpr int GG_VCHUNK = 2048;  // vertices staged per smem tile (32 KB)

__global__ void __launch_bounds__(GG_THREADS) gg_bounds_kernel(const float4* __restrict__ vq, int V, const float* __restrict__ ray_d,
                                                               const float* __restrict__ near_in, const float* __restrict__ far_in,
                                                               int64_t R, float gamma2, float* __restrict__ near_out, float* __restrict__ far_out) {
  __shared__ float4 sv[GG_VCHUNK];
  int64_t r = (int64_t)blockIdx.x * GG_THREADS + threadIdx.x;
  bool live = r < R;
  float dx = 0.f, dy = 0.f, dz = 1.f;
  if (live) { dx = ray_d[3 * r]; dy = ray_d[3 * r + 1]; dz = ray_d[3 * r + 2]; }
  float norm = xnorm3(v3(dx, dy, dz));
  float ux = xdiv(dx, norm), uy = xdiv(dy, norm), uz = xdiv(dz, norm);
  float zmin = 99999.0f, zmax = -99999.0f;
  bool any = false;
  for (int base = 0; base < V; base += GG_VCHUNK) {
    int n = min(GG_VCHUNK, V - base);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += GG_THREADS) sv[i] = vq[base + i];
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
      float4 q = sv[i];
      float z0 = xadd(xadd(xmul(q.x, ux), xmul(q.y, uy)), xmul(q.z, uz));
      float tmp = xsub(q.w, xmul(z0, z0));
      if (tmp < gamma2) {
        float del = xsqrt(xsub(gamma2, tmp));
        zmin = fminf(zmin, xsub(z0, del));
        zmax = fmaxf(zmax, xadd(z0, del));
        any = true;
      }
    }
  }
  if (!live) return;
  zmin = xdiv(zmin, norm);
  zmax = xdiv(zmax, norm);
  bool use = any && (zmin < zmax);
  near_out[r] = use ? zmin : near_in[r];
  far_out[r] = use ? zmax : far_in[r];
}

// torch.linspace(0,1,N) exactly (second half uses a fused multiply-add), see oracle/geom.c
__host__ __device__ inline float linspace01(int i, int n) {
  if (n == 1) return 0.f;
  float step = 1.0f / (float)(n - 1);
#ifdef __CUDA_ARCH__
  return (i < n / 2) ? __fmul_rn(step, (float)i) : __fmaf_rn(-step, (float)(n - i - 1), 1.0f);
#else
  return (i < n / 2) ? step * (float)i : fmaf(-step, (float)(n - i - 1), 1.0f);
#endif
}

// utils/pts_utils.py:3-16 (eval): z = near*(1-t) + far*t
__device__ __forceinline__ float sample_z(float near, float far, float t) { return xadd(xmul(near, xsub(1.0f, t)), xmul(far, t)); }

// Renderer.w2l_without_lbs over all samples (can_render.py:333-379), one thread per sample.
// Phase 1 places the sample and looks its cell up: ~80 % of the samples sit in cells that are far
// from the mesh or provably transparent and stop there.  The rest are compacted into a per-block
// queue so that phase 2 (exact nearest centroid -> project -> mask -> re-emit on the canonical
// triangle) runs with full warps.  Non-transparent samples are appended to the active list; one
// bit per sample tells the compositor which samples carry a raw value, so nothing is written for
// transparent samples (their weight is exactly 0, can_render.py:118-120).
struct WarpArgs {
  const float* ray_o; const float* ray_d; const float* near; const float* far;  // near/far after GG
  const float* z_in;       // optional explicit z (R,N); NULL => linspace
  const float* tvals;      // (N)
  const float* posed; const float* canon; const int* faces;
  int64_t R; int N;
  float4* active;          // (x_c, y_c, z_c, bits(sample id))
  int* active_tri;         // posed-space nearest triangle of each active sample
  unsigned* sample_mask;   // ceil(R*N/32) words, bit s&31 of word s>>5
  unsigned long long* counters;
  int count_candidates;
};

constexpr int WARP_THREADS = 256;

__device__ __forceinline__ void sample_position(const WarpArgs& a, int64_t s, float& px, float& py, float& pz) {
  int64_t r = s / a.N;
  int i = (int)(s - r * a.N);
  float z = a.z_in ? a.z_in[s] : sample_z(a.near[r], a.far[r], __ldg(a.tvals + i));
  px = xadd(a.ray_o[3 * r], xmul(a.ray_d[3 * r], z));
  py = xadd(a.ray_o[3 * r + 1], xmul(a.ray_d[3 * r + 1], z));
  pz = xadd(a.ray_o[3 * r + 2], xmul(a.ray_d[3 * r + 2], z));
}

__global__ void __launch_bounds__(WARP_THREADS) sample_warp_kernel(WarpArgs a, Grid g) {
  __shared__ int queue[WARP_THREADS];
  __shared__ int qn;
  __shared__ unsigned char flag[WARP_THREADS];
  const int64_t P = a.R * a.N;
  const int64_t s0 = (int64_t)blockIdx.x * WARP_THREADS;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) qn = 0;
  flag[threadIdx.x] = 0;
  __syncthreads();
  {
    const int64_t s = s0 + threadIdx.x;
    bool need = false;
    if (s < P) {
      float px, py, pz;
      sample_position(a, s, px, py, pz);
      float fx = (px - g.ox) * g.inv_cell, fy = (py - g.oy) * g.inv_cell, fz = (pz - g.oz) * g.inv_cell;
      if (fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)g.nx && fy < (float)g.ny && fz < (float)g.nz) {
        float dc = __ldg(g.center_dist + ((int)fz * g.ny + (int)fy) * g.nx + (int)fx);
        need = !(dc - g.half_diag > g.r_cap);
      }
    }
    unsigned m = __ballot_sync(0xffffffffu, need);
    int base = 0;
    if (m) {
      int leader = __ffs(m) - 1;
      if (lane == leader) base = atomicAdd(&qn, __popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (need) queue[base + __popc(m & ((1u << lane) - 1))] = threadIdx.x;
    }
  }
  __syncthreads();
  {
    const int n = qn;
    if (a.count_candidates && threadIdx.x == 0 && n) atomicAdd(a.counters + 2, (unsigned long long)n);
    bool act = false;
    V3 xc = v3(0, 0, 0);
    int idx = -1, t = 0;
    if (threadIdx.x < n) {
      t = queue[threadIdx.x];
      float px, py, pz;
      sample_position(a, s0 + t, px, py, pz);
      idx = nearest_centroid(g, px, py, pz, a.count_candidates ? a.counters + 1 : nullptr);
      if (idx >= 0) {
        int i0 = a.faces[3 * idx], i1 = a.faces[3 * idx + 1], i2 = a.faces[3 * idx + 2];
        float u, v, h;
        project_point(v3(px, py, pz), ldv3(a.posed, i0), ldv3(a.posed, i1), ldv3(a.posed, i2), u, v, h);
        if (!is_transparent(u, v, h)) {
          xc = map_to_triangle(u, v, h, ldv3(a.canon, i0), ldv3(a.canon, i1), ldv3(a.canon, i2));
          act = true;
        }
      }
    }
    if ((threadIdx.x & ~31) < n) {  // warp-uniform: this warp holds queue entries
      unsigned m = __ballot_sync(0xffffffffu, act);
      if (m) {
        int leader = __ffs(m) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(a.counters, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (act) {
          unsigned long long slot = base + __popc(m & ((1u << lane) - 1));
          a.active[slot] = make_float4(xc.x, xc.y, xc.z, __int_as_float((int)(s0 + t)));
          a.active_tri[slot] = idx;
          flag[t] = 1;
        }
      }
    }
  }
  __syncthreads();
  {
    unsigned m = __ballot_sync(0xffffffffu, flag[threadIdx.x] != 0);
    const int64_t s = s0 + threadIdx.x;
    if (lane == 0 && s < P) a.sample_mask[s >> 5] = m;
  }
}

// stand-alone warp op (dsnerf_warp_points)
__global__ void warp_points_kernel(const float* __restrict__ pts, int64_t P, const float* __restrict__ posed, const float* __restrict__ canon,
                                   const int* __restrict__ faces, Grid g, int F, const float* __restrict__ cent,
                                   float* __restrict__ xyz_cano, uint8_t* __restrict__ transparent, int* __restrict__ idx_out) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= P) return;
  float px = pts[3 * s], py = pts[3 * s + 1], pz = pts[3 * s + 2];
  int idx = nearest_centroid(g, px, py, pz, nullptr);
  if (idx < 0) {
    // provably transparent, but the stand-alone op still reports the reference's values:
    // fall back to the exhaustive scan for this point (rare: far from the mesh)
    float best = 3.0e38f;
    for (int f = 0; f < F; ++f) {
      float dx = xsub(px, cent[3 * f]), dy = xsub(py, cent[3 * f + 1]), dz = xsub(pz, cent[3 * f + 2]);
      float d = xfma(dz, dz, xfma(dy, dy, xmul(dx, dx)));
      if (d < best) { best = d; idx = f; }
    }
  }
  int i0 = faces[3 * idx], i1 = faces[3 * idx + 1], i2 = faces[3 * idx + 2];
  float u, v, h;
  project_point(v3(px, py, pz), ldv3(posed, i0), ldv3(posed, i1), ldv3(posed, i2), u, v, h);
  V3 xc = map_to_triangle(u, v, h, ldv3(canon, i0), ldv3(canon, i1), ldv3(canon, i2));
  xyz_cano[3 * s] = xc.x; xyz_cano[3 * s + 1] = xc.y; xyz_cano[3 * s + 2] = xc.z;
  if (transparent) transparent[s] = is_transparent(u, v, h) ? 1 : 0;
  if (idx_out) idx_out[s] = idx;
}

}  // namespace dsn
