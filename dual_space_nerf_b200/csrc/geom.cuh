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
// Enumeration grid: cells are x-fastest, so a run of cells along x is one contiguous run of sorted
// centroids; one 64-bit occupancy word per (z,y) row lets the scan skip empty rows and trim ranges.
// Lookup table: twice as fine as the enumeration grid; per cell the distance from the cell centre
// to its nearest centroid (or "provably transparent") and that centroid's index (search seed).
struct Grid {
  float ox, oy, oz;     // origin of cell (0,0,0) of both lattices
  float cell, inv_cell; // enumeration cell edge
  int nx, ny, nz;       // enumeration cells (nx <= 64)
  float tinv;           // 1 / table cell edge (= 2 / cell)
  int tnx, tny, tnz;    // table cells (2nx, 2ny, 2nz)
  float thalf_diag;     // table cell half diagonal (rounded up)
  float r_cap;          // beyond this distance to the nearest centroid a point is provably transparent
  const int* __restrict__ cell_start;              // ncell+1
  const float4* __restrict__ sorted;               // (x,y,z,bits(idx)) sorted by cell
  const unsigned long long* __restrict__ row_mask; // (nz*ny) occupancy bits along x
  const float* __restrict__ cent;                  // (F,3) centroids by index
  const float* __restrict__ center_dist;           // table: distance centre -> nearest centroid (huge = provably transparent)
  const int* __restrict__ center_idx;              // table: index of that centroid
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

__global__ void grid_count_kernel(Grid g, const float* __restrict__ cent, int F, int* __restrict__ counts, unsigned long long* __restrict__ row_mask) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  int cx = grid_coord(cent[3 * f], g.ox, g.inv_cell, g.nx);
  int cy = grid_coord(cent[3 * f + 1], g.oy, g.inv_cell, g.ny);
  int cz = grid_coord(cent[3 * f + 2], g.oz, g.inv_cell, g.nz);
  atomicAdd(&counts[(cz * g.ny + cy) * g.nx + cx], 1);
  atomicOr(&row_mask[cz * g.ny + cy], 1ull << cx);
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

// Visit every centroid stored in a grid cell that intersects the ball (p, sqrt(rho2)).  rho2 may shrink while
// the scan runs (the visitor holds a reference); the initial extent comes from rho.  All bounds are rounded
// outwards, so no cell touching the ball is skipped.
template <class Visit>
__device__ __forceinline__ void scan_ball(const Grid& g, float px, float py, float pz, float rho, const float& rho2, Visit&& visit) {
  int z0 = max(0, (int)floorf((pz - rho - g.oz) * g.inv_cell)), z1 = min(g.nz - 1, (int)floorf((pz + rho - g.oz) * g.inv_cell));
  int y0 = max(0, (int)floorf((py - rho - g.oy) * g.inv_cell)), y1 = min(g.ny - 1, (int)floorf((py + rho - g.oy) * g.inv_cell));
  for (int cz = z0; cz <= z1; ++cz) {
    float zl = g.oz + cz * g.cell;
    float dz = fmaxf(0.f, fmaxf(zl - pz, pz - (zl + g.cell)));
    float dz2 = dz * dz * 0.9999f;
    if (dz2 > rho2) continue;
    for (int cy = y0; cy <= y1; ++cy) {
      unsigned long long m = __ldg(g.row_mask + cz * g.ny + cy);
      if (!m) continue;
      float yl = g.oy + cy * g.cell;
      float dy = fmaxf(0.f, fmaxf(yl - py, py - (yl + g.cell)));
      float rem = rho2 - dz2 - dy * dy * 0.9999f;
      if (rem < 0.f) continue;
      float rx = sqrtf(rem) * 1.0001f + 1e-6f;
      int x0 = max(0, (int)floorf((px - rx - g.ox) * g.inv_cell)), x1 = min(g.nx - 1, (int)floorf((px + rx - g.ox) * g.inv_cell));
      if (x0 > x1) continue;
      m &= (x1 - x0 >= 63 ? ~0ull : ((1ull << (x1 - x0 + 1)) - 1ull)) << x0;
      if (!m) continue;
      x0 = __ffsll((long long)m) - 1;  // trim to the occupied cells of the range
      x1 = 63 - __clzll((long long)m);
      int row = (cz * g.ny + cy) * g.nx;
      int b = __ldg(g.cell_start + row + x0), e = __ldg(g.cell_start + row + x1 + 1);
      for (int j = b; j < e; ++j) visit(__ldg(g.sorted + j));
    }
  }
}

// ---- lookup-table construction ---------------------------------------------------------------
// Level 0: brute force (tiled through shared memory) on a lattice 4x coarser than the table: nearest centroid of
// every coarse cell centre.  Only seeds level 1.
__global__ void table_coarse_kernel(float ox, float oy, float oz, float cell0, int n0x, int n0y, int n0z, const float* __restrict__ cent,
                                    const float4* __restrict__ tri_n, int F, int classify, float r_cap, float* __restrict__ dc0,
                                    int* __restrict__ idx0) {
  __shared__ float sx[1024], sy[1024], sz[1024], snx[1024], sny[1024], snz[1024];
  int n = n0x * n0y * n0z;
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  int cx = c % n0x, cy = (c / n0x) % n0y, cz = c / (n0x * n0y);
  float px = ox + (cx + 0.5f) * cell0, py = oy + (cy + 0.5f) * cell0, pz = oz + (cz + 0.5f) * cell0;
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
  // Same transparency proof as for the table cells, one level up: a certified coarse cell certifies every table cell inside
  // it, which spares the far band (the most expensive scans) in table_fine_kernel.
  const float hd0 = cell0 * 0.8660254f * 1.001f;
  const float h_thr = 0.1f + hd0 + 1e-4f;
  bool undecided = classify && (c < n) && (dc > h_thr) && (dc - hd0 <= r_cap);
  bool search = (c < n) && (!classify || dc <= h_thr);
  if (__syncthreads_or(undecided)) {
    float thr = dc + 2.0f * hd0 + 1e-5f;
    float thr2 = undecided ? thr * thr : -1.0f;
    for (int base = 0; base < F; base += 1024) {
      __syncthreads();
      for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        int f = base + i;
        bool ok = f < F;
        float4 nq = ok ? tri_n[f] : make_float4(0.f, 0.f, 0.f, 0.f);
        sx[i] = ok ? cent[3 * f] : 1.0e18f;
        sy[i] = ok ? cent[3 * f + 1] : 1.0e18f;
        sz[i] = ok ? cent[3 * f + 2] : 1.0e18f;
        snx[i] = nq.x; sny[i] = nq.y; snz[i] = nq.z;
      }
      __syncthreads();
#pragma unroll 4
      for (int i = 0; i < 1024; ++i) {
        float dx = px - sx[i], dy = py - sy[i], dz = pz - sz[i];
        if (dx * dx + dy * dy + dz * dz <= thr2) {
          float h = dx * snx[i] + dy * sny[i] + dz * snz[i];
          if (!(fabsf(h) > h_thr)) search = true;
        }
      }
    }
  }
  // +huge marks a far or certified (provably transparent) coarse cell
  if (c < n) { dc0[c] = (search && dc - hd0 <= r_cap) ? dc : 3.0e30f; idx0[c] = besti; }
}

// Finer levels, each seeded by its parent (cells twice as large): per cell the distance dc from the centre to its
// nearest centroid (found through the enumeration grid) and that centroid's index; +huge when every point of the cell is
// farther than r_cap from all centroids or PROVABLY transparent.  Proof: for p in the cell, its nearest centroid c*
// satisfies |centre - c*| <= dc + 2*half_diag (candidate set), the signed plane distance h is 1-Lipschitz, so
// |h*(centre)| > 0.1 + half_diag for every candidate implies |h*(p)| > 0.1 = max_dist of get_transparent_mask
// (utils/render_utils.py:103).  A far/certified parent settles all its children.  The last level is the lookup table.
struct TableLevel {
  float cell, half_diag;
  int nx, ny, nz;
};

constexpr int TABLE_THREADS = 256;

__global__ void __launch_bounds__(TABLE_THREADS) table_level_kernel(Grid g, TableLevel lv, const float4* __restrict__ tri_n, int pnx, int pny, int pnz,
                                                                    const float* __restrict__ p_dc, const int* __restrict__ p_idx, int classify,
                                                                    float* __restrict__ out, int* __restrict__ out_idx) {
  // Most cells are settled by their parent (far / certified); the others are compacted into a per-block queue so that the
  // scans below run with full warps.
  __shared__ int queue[TABLE_THREADS];
  __shared__ int qn;
  const int n = lv.nx * lv.ny * lv.nz;
  const int c0 = blockIdx.x * TABLE_THREADS;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) qn = 0;
  __syncthreads();
  {
    int c = c0 + threadIdx.x;
    bool need = false;
    if (c < n) {
      int cx = c % lv.nx, cy = (c / lv.nx) % lv.ny, cz = c / (lv.nx * lv.ny);
      int pc = (min(cz / 2, pnz - 1) * pny + min(cy / 2, pny - 1)) * pnx + min(cx / 2, pnx - 1);
      if (p_dc[pc] > 1.0e29f) { out[c] = 3.0e30f; out_idx[c] = p_idx[pc]; }
      else need = true;
    }
    unsigned m = __ballot_sync(0xffffffffu, need);
    if (m) {
      int leader = __ffs(m) - 1, base = 0;
      if (lane == leader) base = atomicAdd(&qn, __popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (need) queue[base + __popc(m & ((1u << lane) - 1))] = threadIdx.x;
    }
  }
  __syncthreads();
  if (threadIdx.x >= qn) return;
  const int c = c0 + queue[threadIdx.x];
  int cx = c % lv.nx, cy = (c / lv.nx) % lv.ny, cz = c / (lv.nx * lv.ny);
  float px = g.ox + (cx + 0.5f) * lv.cell, py = g.oy + (cy + 0.5f) * lv.cell, pz = g.oz + (cz + 0.5f) * lv.cell;
  int pc = (min(cz / 2, pnz - 1) * pny + min(cy / 2, pny - 1)) * pnx + min(cx / 2, pnx - 1);
  int seed = p_idx[pc];
  float sx = px - g.cent[3 * seed], sy = py - g.cent[3 * seed + 1], sz = pz - g.cent[3 * seed + 2];
  float best = sx * sx + sy * sy + sz * sz;
  int besti = seed;
  float rho2 = best * 1.0001f + 1e-12f;
  scan_ball(g, px, py, pz, sqrtf(rho2) * 1.0001f, rho2, [&](float4 q) {
    float dx = px - q.x, dy = py - q.y, dz = pz - q.z;
    float d2 = dx * dx + dy * dy + dz * dz;
    if (d2 < best) { best = d2; besti = __float_as_int(q.w); rho2 = best * 1.0001f + 1e-12f; }
  });
  float dc = sqrtf(best) * 1.00001f + 1e-7f;
  out_idx[c] = besti;
  if (dc - lv.half_diag > g.r_cap) { out[c] = 3.0e30f; return; }
  const float h_thr = 0.1f + lv.half_diag * 1.001f + 1e-4f;
  bool search = !classify || dc <= h_thr;  // the nearest centroid itself is a candidate with |h| <= dc
  if (!search) {
    float thr = dc + 2.0f * lv.half_diag * 1.001f + 1e-5f;
    float thr2 = thr * thr;
    float live2 = thr2;  // set negative to cut the scan short once the cell is known to need searching
    scan_ball(g, px, py, pz, thr * 1.0001f, live2, [&](float4 q) {
      float dx = px - q.x, dy = py - q.y, dz = pz - q.z;
      if (dx * dx + dy * dy + dz * dz <= thr2) {
        float4 nq = __ldg(tri_n + __float_as_int(q.w));
        float h = dx * nq.x + dy * nq.y + dz * nq.z;
        if (!(fabsf(h) > h_thr)) { search = true; live2 = -1.0f; }  // NaN normal (degenerate triangle) keeps the cell searchable
      }
    });
  }
  out[c] = search ? dc : 3.0e30f;
}

// Exact nearest centroid: squared L2 accumulated as d0*d0, fma(d1,d1,.), fma(d2,d2,.) and
// strict '<' with lowest index on ties -- the arithmetic of pytorch3d 0.4.0 knn_points(K=1)
// as called at utils/render_utils.py:95.  Returns -1 when the point is provably farther than
// g.r_cap from every centroid or sits in a provably transparent cell.  The scan visits only
// grid rows that intersect the ball of the current best radius, which starts at the exact
// distance to a seed centroid, so it returns the same index as a full scan.
// `hint` (>= 0) is a caller-supplied seed (the posed-space triangle for the canonical search);
// without it the seed is the table's centroid nearest to the cell centre.
__device__ __forceinline__ int nearest_centroid(const Grid& g, float px, float py, float pz, unsigned long long* cand_counter,
                                                int hint = -1) {
  float fx = (px - g.ox) * g.tinv, fy = (py - g.oy) * g.tinv, fz = (pz - g.oz) * g.tinv;
  // points outside the table region are farther than r_cap from the mesh by construction
  if (!(fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)g.tnx && fy < (float)g.tny && fz < (float)g.tnz)) return -1;
  const int cell = ((int)fz * g.tny + (int)fy) * g.tnx + (int)fx;
  float dc = __ldg(g.center_dist + cell);
  if (hint < 0) {
    if (dc - g.thalf_diag > g.r_cap) return -1;
    hint = __ldg(g.center_idx + cell);
  } else if (dc > 1.0e29f) {
    dc = g.r_cap;  // classified cell, but the caller vouches for a nearby centroid: let the hint set the radius
  }
  // true nearest distance <= dc + half_diag; a candidate beyond that bound means the point is farther than r_cap anyway
  const float bound = fminf(dc + g.thalf_diag, g.r_cap * 1.0001f + g.thalf_diag);
  const float rho2_init = bound * bound * 1.0001f;
  float best, rho2;
  int besti = hint;
  {
    float dx = xsub(px, __ldg(g.cent + 3 * hint)), dy = xsub(py, __ldg(g.cent + 3 * hint + 1)), dz = xsub(pz, __ldg(g.cent + 3 * hint + 2));
    best = xfma(dz, dz, xfma(dy, dy, xmul(dx, dx)));
    rho2 = fminf(rho2_init, best * 1.0001f + 1e-12f);
  }
  unsigned ncand = 0;
  scan_ball(g, px, py, pz, sqrtf(rho2) * 1.0001f, rho2, [&](float4 c) {
    float dx = xsub(px, c.x), dy = xsub(py, c.y), dz = xsub(pz, c.z);
    float d = xmul(dx, dx);
    d = xfma(dy, dy, d);
    d = xfma(dz, dz, d);
    int id = __float_as_int(c.w);
    ++ncand;
    if (d < best || (d == best && id < besti)) {
      best = d;
      besti = id;
      rho2 = fminf(rho2, best * 1.0001f + 1e-12f);
    }
  });
  if (cand_counter && ncand) atomicAdd(cand_counter, (unsigned long long)ncand);
  if (best > rho2_init) besti = -1;
  return besti;
}

// ---------------------------------------------------------------------------------------------
// geometry_guided_ray_marching, utils/pts_utils.py:18-53 (near/far only).
// vq = (vertex - o0, |vertex - o0|^2) per vertex, prepared by gg_prep_kernel.
__device__ __forceinline__ unsigned f2key(float f) { unsigned b = __float_as_uint(f); return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u); }
__device__ __forceinline__ float key2f(unsigned k) { return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu)); }

// qbox: 6 order-preserving keys, [0..2] = min (initialised to 0xff..), [3..5] = max (initialised to 0) of q
__global__ void gg_prep_kernel(const float* __restrict__ xyz, int V, const float* __restrict__ ray_o, float4* __restrict__ vq,
                               unsigned* __restrict__ qbox) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  // the reference uses ray_o[:, 0:1] -- the FIRST ray's origin -- for every ray (pts_utils.py:31,33)
  float qx = 0.f, qy = 0.f, qz = 0.f;
  bool live = v < V;
  if (live) {
    qx = xsub(xyz[3 * v], ray_o[0]); qy = xsub(xyz[3 * v + 1], ray_o[1]); qz = xsub(xyz[3 * v + 2], ray_o[2]);
    float qq = xadd(xadd(xmul(qx, qx), xmul(qy, qy)), xmul(qz, qz));
    vq[v] = make_float4(qx, qy, qz, qq);
  }
  float lo[3] = {live ? qx : 3e38f, live ? qy : 3e38f, live ? qz : 3e38f};
  float hi[3] = {live ? qx : -3e38f, live ? qy : -3e38f, live ? qz : -3e38f};
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int k = 0; k < 3; ++k) { atomicMin(qbox + k, f2key(lo[k])); atomicMax(qbox + 3 + k, f2key(hi[k])); }
}

constexpr int GG_THREADS = 256;
constexpr int GG_VCHUNK = 2048;  // vertices staged per smem tile (32 KB)

__global__ void __launch_bounds__(GG_THREADS) gg_bounds_kernel(const float4* __restrict__ vq, int V, const unsigned* __restrict__ qbox,
                                                               const float* __restrict__ ray_d, const float* __restrict__ near_in,
                                                               const float* __restrict__ far_in, int64_t R, float gamma2, float gamma,
                                                               float* __restrict__ near_out, float* __restrict__ far_out) {
  __shared__ float4 sv[GG_VCHUNK];
  int64_t r = (int64_t)blockIdx.x * GG_THREADS + threadIdx.x;
  bool live = r < R;
  float dx = 0.f, dy = 0.f, dz = 1.f;
  if (live) { dx = ray_d[3 * r]; dy = ray_d[3 * r + 1]; dz = ray_d[3 * r + 2]; }
  float norm = xnorm3(v3(dx, dy, dz));
  float ux = xdiv(dx, norm), uy = xdiv(dy, norm), uz = xdiv(dz, norm);
  // cull: a vertex can be within gamma of the ray's LINE (the test below has no t >= 0 restriction) only if the line
  // meets the vertex box inflated by gamma (+ rounding slack); rays of whole blocks far from the body skip the scan
  bool maybe = false;
  if (live) {
    const float pad = gamma * 1.001f + 1e-4f;
    float t0 = -3e38f, t1 = 3e38f;
    const float u[3] = {ux, uy, uz};
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float lo = key2f(qbox[k]) - pad, hi = key2f(qbox[3 + k]) + pad;
      if (fabsf(u[k]) < 1e-12f) {
        ok = ok && (lo <= 0.f && 0.f <= hi);
      } else {
        float a = lo / u[k], b = hi / u[k];
        t0 = fmaxf(t0, fminf(a, b));
        t1 = fminf(t1, fmaxf(a, b));
      }
    }
    maybe = ok && (t0 <= t1 * (1.0f + 1e-5f) + 1e-5f);
  }
  float zmin = 99999.0f, zmax = -99999.0f;
  bool any = false;
  if (__syncthreads_or(maybe)) {
    for (int base = 0; base < V; base += GG_VCHUNK) {
      int n = min(GG_VCHUNK, V - base);
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += GG_THREADS) sv[i] = vq[base + i];
      __syncthreads();
      if (maybe) {
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
  __shared__ int bin_count[8], bin_start[8];
  __shared__ unsigned char flag[WARP_THREADS];
  const int64_t P = a.R * a.N;
  const int64_t s0 = (int64_t)blockIdx.x * WARP_THREADS;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 8) bin_count[threadIdx.x] = 0;
  flag[threadIdx.x] = 0;
  __syncthreads();
  // phase 1: place the sample, look its table cell up.  Samples that need the exact search are queued, bucketed by the
  // table's distance (= expected search radius) so that the lanes of a warp in phase 2 do similar amounts of work.
  int my_bin = -1;
  {
    const int64_t s = s0 + threadIdx.x;
    if (s < P) {
      float px, py, pz;
      sample_position(a, s, px, py, pz);
      float fx = (px - g.ox) * g.tinv, fy = (py - g.oy) * g.tinv, fz = (pz - g.oz) * g.tinv;
      if (fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)g.tnx && fy < (float)g.tny && fz < (float)g.tnz) {
        float dc = __ldg(g.center_dist + ((int)fz * g.tny + (int)fy) * g.tnx + (int)fx);
        if (!(dc - g.thalf_diag > g.r_cap)) my_bin = min(7, (int)(dc * 40.0f));  // 2.5 cm classes
      }
    }
    if (my_bin >= 0) atomicAdd(&bin_count[my_bin], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = 7; b >= 0; --b) { bin_start[b] = acc; acc += bin_count[b]; }  // big radii first
    qn = acc;
  }
  __syncthreads();
  if (my_bin >= 0) queue[atomicAdd(&bin_start[my_bin], 1)] = threadIdx.x;
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
