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
// centroids; one 64-bit occupancy word per (z,y) row lets a scan skip empty rows and trim ranges.
// Every enumeration cell also carries an approximate nearest centroid (jump flooding), the seed of
// all exact searches.
//
// Lookup table: twice as fine as the enumeration grid and built LAZILY, only for the cells that the
// points of a call actually fall into (mark_*_kernel -> build_cells_kernel).  A built cell holds
//   * cnt == -1: every point of the cell is farther than r_cap from all centroids or PROVABLY
//     transparent (posed mesh only) -> no search at all;
//   * cnt >= 0: the CANDIDATE LIST of the cell, i.e. every centroid that can be the nearest one of
//     some point of the cell (off = first entry in the pool, (x, y, z, bits(index)) each);
//   * cnt == -2: list too long / pool exhausted -> exact ball scan seeded with centroid `off`.
// Candidate filter: with c0 the nearest centroid of the cell centre x and a the cell's half edge,
// f(p) = |p-c|^2 - |p-c0|^2 is linear in p, so its minimum over the cube is f(x) - 2a|c-c0|_1: c
// can beat c0 (and hence be nearest) somewhere in the cell only if f(x) <= 2a|c-c0|_1.  All such c
// lie within dmin + 2*half_diag of x.  The list is a superset of the possible exact winners (slack
// covers fp32 rounding), so searching it with the reference arithmetic returns the same index as a
// brute-force scan, ties included.
// Transparency proof (posed): the signed plane distance h of a triangle is 1-Lipschitz and the
// centroid lies in the plane, so |(x-c).n_c| > 0.1 + half_diag for every listed c implies |h| > 0.1
// = max_dist of get_transparent_mask (utils/render_utils.py:103) for every point of the cell.
struct Grid {
  float ox, oy, oz;     // origin of cell (0,0,0) of both lattices
  float cell, inv_cell; // enumeration cell edge
  int nx, ny, nz;       // enumeration cells (nx <= 64)
  float tinv;           // 1 / table cell edge (= 2 / cell)
  int tnx, tny, tnz;    // table cells (2nx, 2ny, 2nz)
  float thalf_diag;     // table cell half diagonal (rounded up)
  float r_cap;          // beyond this distance to the nearest centroid a point is provably transparent
  int classify;         // 1: posed mesh (transparency proof), 0: canonical
  int F;
  const int* __restrict__ cell_start;              // ncell+1
  const float4* __restrict__ sorted;               // (x,y,z,bits(idx)) sorted by cell
  const unsigned long long* __restrict__ row_mask; // (nz*ny) occupancy bits along x
  const float* __restrict__ cent;                  // (F,3) centroids by index
  const float4* __restrict__ tri_n;                // (F) unit normals (classification only)
  const int* __restrict__ enum_seed;               // per enumeration cell: approximately nearest centroid
  const unsigned char* __restrict__ enum_far;      // per enumeration cell: 1 = every point of it is farther than r_cap from all centroids
  unsigned char* __restrict__ estate;              // per enumeration cell (posed): 0 untouched, 1 requested, 2 certified transparent, 3 not certified
  int* __restrict__ ereq;                          // requested enumeration cells of the current call (count in pool_used[2])
  int* __restrict__ req2;                          // table cells left to LEVEL 2 (count in pool_used[3])
  unsigned char* __restrict__ tstate;              // per table cell: 0 untouched, 1 requested, 2 built
  int2* __restrict__ trec;                         // per table cell: (off, cnt), see above
  float4* __restrict__ pool;                       // candidate lists
  int pool_cap;
  int* __restrict__ pool_used;                     // [0] entries used, [1] table-cell requests, [2] enumeration-cell requests, [3] LEVEL-2 left-overs, [4..10] and [13..15] debug counters, [16], [17] work counters of the two build levels, [19] lookups canon_nearest_kernel queued for canon_long_kernel
  int debug;                                       // count build outcomes in pool_used[4..10]
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

// single-block exclusive scan: counts[ncell] -> start[ncell+1]; also copies start into cursor.
// 8192 elements per round: coalesced load into shared memory, 8 contiguous elements per thread, one block scan of the partials.
__global__ void __launch_bounds__(1024) grid_scan_kernel(const int* __restrict__ counts, int ncell, int* __restrict__ start, int* __restrict__ cursor) {
  __shared__ int sh[8192];
  __shared__ int warp_sums[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < ncell; base += 8192) {
    __syncthreads();
    for (int i = threadIdx.x; i < 8192; i += 1024) sh[i] = base + i < ncell ? counts[base + i] : 0;
    __syncthreads();
    int v[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { v[k] = sh[threadIdx.x * 8 + k]; sum += v[k]; }
    int s = sum;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
    if (lane == 31) warp_sums[wid] = s;
    __syncthreads();
    if (wid == 0) {
      int w = warp_sums[lane];
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
      warp_sums[lane] = w;
    }
    __syncthreads();
    int excl = carry + (wid ? warp_sums[wid - 1] : 0) + s - sum;
#pragma unroll
    for (int k = 0; k < 8; ++k) { sh[threadIdx.x * 8 + k] = excl; excl += v[k]; }
    __syncthreads();
    for (int i = threadIdx.x; i < 8192; i += 1024)
      if (base + i < ncell) { start[base + i] = sh[i]; cursor[base + i] = sh[i]; }
    if (threadIdx.x == 1023) carry = excl;
  }
  __syncthreads();
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

// ---- seeds: approximately nearest centroid per enumeration cell (jump flooding) ------------------
__global__ void jfa_init_kernel(Grid g, int* __restrict__ seed) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g.nx * g.ny * g.nz) return;
  int b = g.cell_start[c], e = g.cell_start[c + 1];
  seed[c] = b < e ? __float_as_int(g.sorted[b].w) : -1;
}
__global__ void jfa_pass_kernel(Grid g, int step, const int* __restrict__ in, int* __restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g.nx * g.ny * g.nz) return;
  int cx = c % g.nx, cy = (c / g.nx) % g.ny, cz = c / (g.nx * g.ny);
  float px = g.ox + (cx + 0.5f) * g.cell, py = g.oy + (cy + 0.5f) * g.cell, pz = g.oz + (cz + 0.5f) * g.cell;
  int best = in[c];
  float bd = 3.0e38f;
  if (best >= 0) { float dx = px - g.cent[3 * best], dy = py - g.cent[3 * best + 1], dz = pz - g.cent[3 * best + 2]; bd = dx * dx + dy * dy + dz * dz; }
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy)
      for (int dx = -1; dx <= 1; ++dx) {
        int x = cx + dx * step, y = cy + dy * step, z = cz + dz * step;
        if ((dx | dy | dz) == 0 || x < 0 || y < 0 || z < 0 || x >= g.nx || y >= g.ny || z >= g.nz) continue;
        int cand = in[(z * g.ny + y) * g.nx + x];
        if (cand < 0 || cand == best) continue;
        float ex = px - g.cent[3 * cand], ey = py - g.cent[3 * cand + 1], ez = pz - g.cent[3 * cand + 2];
        float d = ex * ex + ey * ey + ez * ez;
        if (d < bd) { bd = d; best = cand; }
      }
  out[c] = best;
}

// Exact far test per enumeration cell: for a point p of cell A and a centroid q of occupied cell B the per-axis gap is at
// least max(0, |A_k - B_k| - 1) cells, so cell^2 * min_B sum_k max(0, |A_k - B_k| - 1)^2 > r_cap^2 proves that every point of
// A is farther than r_cap from all centroids (=> transparent, see transparency_radius).  Rows are searched through the
// occupancy words (nearest set bit on either side of the cell's x).
// The minimum separates: pass 1 takes, per cell, the minimum over the rows y' of its z slab of gx^2 + gy^2 (window rows), pass
// 2 the minimum over the slabs z' of that plus gz^2 -- 2 x 13 steps per cell instead of 13 x 13, same integers.
__global__ void enum_far_rows_kernel(Grid g, int window, int* __restrict__ tmp) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g.nx * g.ny * g.nz) return;
  int cx = c % g.nx, cy = (c / g.nx) % g.ny, cz = c / (g.nx * g.ny);
  int best = 1 << 28;
  for (int y = max(0, cy - window); y <= min(g.ny - 1, cy + window); ++y) {
    unsigned long long m = __ldg(g.row_mask + cz * g.ny + y);
    if (!m) continue;
    int gy = max(0, abs(y - cy) - 1);
    // nearest set bit at or below cx, and at or above cx
    unsigned long long lo = m & (cx >= 63 ? ~0ull : ((2ull << cx) - 1ull)), hi = m & (~0ull << cx);
    int dx = 1 << 12;
    if (lo) dx = min(dx, cx - (63 - __clzll((long long)lo)));
    if (hi) dx = min(dx, (__ffsll((long long)hi) - 1) - cx);
    int gx = max(0, dx - 1);
    best = min(best, gx * gx + gy * gy);
  }
  tmp[c] = best;
}
__global__ void enum_far_kernel(Grid g, int window, const int* __restrict__ tmp, unsigned char* __restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= g.nx * g.ny * g.nz) return;
  int cz = c / (g.nx * g.ny);
  const int plane = g.nx * g.ny;
  const float need = (g.r_cap * 1.0002f + 1e-5f) * g.inv_cell;
  const float need2 = need * need;
  int best = 1 << 28;
  for (int z = max(0, cz - window); z <= min(g.nz - 1, cz + window); ++z) {
    int gz = max(0, abs(z - cz) - 1);
    best = min(best, __ldg(tmp + c + (z - cz) * plane) + gz * gz);
  }
  out[c] = ((float)best <= need2) ? 0 : 1;
}

// ---- lazy lookup table -----------------------------------------------------------------------
__device__ __forceinline__ int table_cell(const Grid& g, float px, float py, float pz) {
  float fx = (px - g.ox) * g.tinv, fy = (py - g.oy) * g.tinv, fz = (pz - g.oz) * g.tinv;
  // points outside the table region are farther than r_cap from the mesh by construction
  if (!(fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)g.tnx && fy < (float)g.tny && fz < (float)g.tnz)) return -1;
  return ((int)fz * g.tny + (int)fy) * g.tnx + (int)fx;
}
__device__ __forceinline__ int parent_cell(const Grid& g, int cell) {  // enumeration cell that contains a table cell
  int tx = cell % g.tnx, ty = (cell / g.tnx) % g.tny, tz = cell / (g.tnx * g.tny);
  return ((tz >> 1) * g.ny + (ty >> 1)) * g.nx + (tx >> 1);
}
// table cell of a point, or -1 when the point is outside the table or in a provably far enumeration cell
__device__ __forceinline__ int live_cell(const Grid& g, float px, float py, float pz) {
  float fx = (px - g.ox) * g.tinv, fy = (py - g.oy) * g.tinv, fz = (pz - g.oz) * g.tinv;
  if (!(fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)g.tnx && fy < (float)g.tny && fz < (float)g.tnz)) return -1;
  const int tx = (int)fx, ty = (int)fy, tz = (int)fz;
  if (__ldg(g.enum_far + ((tz >> 1) * g.ny + (ty >> 1)) * g.nx + (tx >> 1))) return -1;
  return (tz * g.tny + ty) * g.tnx + tx;
}
// byte-wide "0 -> 1" compare-and-swap through the containing word; true for the one thread that made the transition
__device__ __forceinline__ bool claim_byte(unsigned char* bytes, int i) {
  unsigned int* word = reinterpret_cast<unsigned int*>(bytes + (i & ~3));
  const unsigned sh = (unsigned)(i & 3) * 8u;
  unsigned old = *word;
  for (;;) {
    if ((old >> sh) & 0xffu) return false;  // somebody else requested / built it
    unsigned seen = atomicCAS(word, old, old | (1u << sh));
    if (seen == old) return true;
    old = seen;
  }
}
// request the table cell of a point (and, for the posed mesh, its enumeration cell): one atomic per distinct cell of a
// warp, and only while the cell is untouched
__device__ __forceinline__ void request_cell(const Grid& g, int cell) {
  bool want = cell >= 0 && g.tstate[cell] == 0;
  unsigned act = __ballot_sync(0xffffffffu, want);
  if (!want) return;
  unsigned peers = __match_any_sync(act, cell);
  if ((__ffs(peers) - 1) != (int)(threadIdx.x & 31)) return;
  if (!claim_byte(g.tstate, cell)) return;
  atomicAdd(g.pool_used + 1, 1);
  const int par = parent_cell(g, cell);
  const unsigned char es = g.estate[par];
  if (es == 0) {
    if (claim_byte(g.estate, par)) g.ereq[atomicAdd(g.pool_used + 2, 1)] = par;  // built (with its requested children) by LEVEL 1
  } else if (es >= 2) {
    g.req2[atomicAdd(g.pool_used + 3, 1)] = cell;  // parent settled by an earlier call of this frame: LEVEL 2
  }
}

// Warp-cooperative visit of every centroid stored in a grid cell that intersects the ball (p, rho): the (z, y) rows of the
// ball's bounding box are dealt to the lanes (a row costs three dependent loads before its centroids can be read, so 32
// rows in flight hide that latency), each lane walks its row's run of centroids serially.  visit(q) runs per lane.
// visit(q, j): q = g.sorted[j]
template <class Visit>
__device__ __forceinline__ void warp_scan_ball(const Grid& g, float px, float py, float pz, float rho, Visit&& visit) {
  const int lane = threadIdx.x & 31;
  const float rho2 = rho * rho;
  const int z0 = max(0, (int)floorf((pz - rho - g.oz) * g.inv_cell)), z1 = min(g.nz - 1, (int)floorf((pz + rho - g.oz) * g.inv_cell));
  const int y0 = max(0, (int)floorf((py - rho - g.oy) * g.inv_cell)), y1 = min(g.ny - 1, (int)floorf((py + rho - g.oy) * g.inv_cell));
  const int ny = y1 - y0 + 1, nrows = ny * (z1 - z0 + 1);
  for (int r = lane; r < nrows; r += 32) {
    const int cz = z0 + r / ny, cy = y0 + r % ny;
    unsigned long long m = __ldg(g.row_mask + cz * g.ny + cy);
    if (!m) continue;
    const float zl = g.oz + cz * g.cell, yl = g.oy + cy * g.cell;
    const float dz = fmaxf(0.f, fmaxf(zl - pz, pz - (zl + g.cell)));
    const float dy = fmaxf(0.f, fmaxf(yl - py, py - (yl + g.cell)));
    const float rem = rho2 - (dz * dz + dy * dy) * 0.9999f;
    if (rem < 0.f) continue;
    const float rx = sqrtf(rem) * 1.0001f + 1e-6f;
    int x0 = max(0, (int)floorf((px - rx - g.ox) * g.inv_cell)), x1 = min(g.nx - 1, (int)floorf((px + rx - g.ox) * g.inv_cell));
    if (x0 > x1) continue;
    m &= (x1 - x0 >= 63 ? ~0ull : ((1ull << (x1 - x0 + 1)) - 1ull)) << x0;
    if (!m) continue;
    x0 = __ffsll((long long)m) - 1;
    x1 = 63 - __clzll((long long)m);
    const int row = (cz * g.ny + cy) * g.nx;
    const int b = __ldg(g.cell_start + row + x0), e = __ldg(g.cell_start + row + x1 + 1);
    if (b < e) {
      float4 q = __ldg(g.sorted + b);
      for (int j = b + 1; j < e; ++j) {
        const float4 qn = __ldg(g.sorted + j);  // next load in flight while q is visited
        visit(q, j - 1);
        q = qn;
      }
      visit(q, e - 1);
    }
  }
  __syncwarp();
}

// Lane-uniform form of warp_scan_ball (the table builder's scans): the row-per-lane walk above leaves 8 of 32 lanes busy on
// average (rows hold 0..30 centroids; ncu: 8.3 active threads per instruction in build_cells_kernel<1>).  Here the (z, y)
// rows of the ball are taken in batches of 32, one per lane; a warp prefix sum over the rows' centroid counts turns the
// batch into one flat index space, and the lanes walk that space 32 items at a time (the row of an item is found by a 5-step
// binary search over the lanes' prefix values through shuffles).  The warp stays converged, so the visitor may use ballots:
// visit(q, j, valid) is called by ALL lanes, q = g.sorted[j] where valid.
// BATCH > 1: the records of BATCH consecutive steps are loaded before the first of them is visited.  One step is one dependent
// L2 access of the warp; the second build level runs one warp per cell through three scans of a few thousand centroids each,
// and that chain IS the kernel's duration (8 % of the warp slots in use), so there the loads are worth the registers.
template <int BATCH = 1, class Visit>
__device__ __forceinline__ void warp_scan_ball_flat(const Grid& g, float px, float py, float pz, float rho, Visit&& visit) {
  const int lane = threadIdx.x & 31;
  const float rho2 = rho * rho;
  const int z0 = max(0, (int)floorf((pz - rho - g.oz) * g.inv_cell)), z1 = min(g.nz - 1, (int)floorf((pz + rho - g.oz) * g.inv_cell));
  const int y0 = max(0, (int)floorf((py - rho - g.oy) * g.inv_cell)), y1 = min(g.ny - 1, (int)floorf((py + rho - g.oy) * g.inv_cell));
  const int ny = y1 - y0 + 1, nrows = ny * (z1 - z0 + 1);
  for (int r0 = 0; r0 < nrows; r0 += 32) {
    const int r = r0 + lane;
    int b = 0, cnt = 0;
    if (r < nrows) {
      const int cz = z0 + r / ny, cy = y0 + r % ny;
      unsigned long long m = __ldg(g.row_mask + cz * g.ny + cy);
      if (m) {
        const float zl = g.oz + cz * g.cell, yl = g.oy + cy * g.cell;
        const float dz = fmaxf(0.f, fmaxf(zl - pz, pz - (zl + g.cell)));
        const float dy = fmaxf(0.f, fmaxf(yl - py, py - (yl + g.cell)));
        const float rem = rho2 - (dz * dz + dy * dy) * 0.9999f;
        if (rem >= 0.f) {
          const float rx = sqrtf(rem) * 1.0001f + 1e-6f;
          int x0 = max(0, (int)floorf((px - rx - g.ox) * g.inv_cell)), x1 = min(g.nx - 1, (int)floorf((px + rx - g.ox) * g.inv_cell));
          if (x0 <= x1) {
            m &= (x1 - x0 >= 63 ? ~0ull : ((1ull << (x1 - x0 + 1)) - 1ull)) << x0;
            if (m) {
              x0 = __ffsll((long long)m) - 1;
              x1 = 63 - __clzll((long long)m);
              const int row = (cz * g.ny + cy) * g.nx;
              b = __ldg(g.cell_start + row + x0);
              cnt = __ldg(g.cell_start + row + x1 + 1) - b;
            }
          }
        }
      }
    }
    // inclusive prefix of the counts over the lanes
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - cnt;
    for (int t0 = 0; t0 < total; t0 += 32 * BATCH) {
      float4 q[BATCH];
      int j[BATCH];
#pragma unroll
      for (int u = 0; u < BATCH; ++u) {
        const int t = t0 + 32 * u + lane;
        // the row of item t: the first lane whose inclusive prefix exceeds t
        int lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
          const int probe = __shfl_sync(0xffffffffu, incl, lo + step - 1);
          if (probe <= t) lo += step;
        }
        lo = min(lo, 31);
        const int rb = __shfl_sync(0xffffffffu, b, lo), re = __shfl_sync(0xffffffffu, excl, lo);
        j[u] = rb + (t - re);
        q[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < total) q[u] = __ldg(g.sorted + j[u]);
      }
#pragma unroll
      for (int u = 0; u < BATCH; ++u)
        if (u == 0 || t0 + 32 * u < total) visit(q[u], j[u], t0 + 32 * u + lane < total);  // (warp uniform)
    }
  }
  __syncwarp();
}

#ifndef DSN_LIST_CAP
#define DSN_LIST_CAP 256
#endif
constexpr int LIST_CAP = DSN_LIST_CAP;      // longest candidate list settle_cell assembles in shared memory (64 -> 256: the cells deep
                                         // inside the body see a whole ring of centroids; 7 % -> 0.5 % of the lookups scanned); longer
                                         // ones are filtered a second time straight into the pool (settle_cell / long_list)
#ifndef DSN_BUF_CAP
#define DSN_BUF_CAP 448
#endif
constexpr int BUF_CAP = DSN_BUF_CAP;      // candidates buffered per enumeration cell (superset shared by its 8 table cells)
constexpr int BUILD_WARPS = 4;

// Candidate filter of a cube (centre x, half edge a) against a reference centroid r with |x - r|^2 = dr2: a centroid q
// (|x - q|^2 = d) can be nearer than r somewhere in the cube only if d - dr2 <= 2a|q - r|_1 (slack for fp32 rounding).
__device__ __forceinline__ bool can_beat(float d, float dr2, float a, float l1) { return d - dr2 <= 2.0f * a * l1 * 1.0002f + 1e-7f + 2e-6f * d; }

// Settle one table cell from a buffered candidate superset S[0..n) (all lanes of the warp call this): exact nearest c0 of the
// cell centre, filter against c0, transparency proof (posed mesh), list -> pool.  `list` is a LIST_CAP scratch buffer.
// S and list hold POSITIONS in g.sorted (4 bytes per candidate instead of the 16-byte record: the records are re-read
// through L1, and the shared memory saved raises the builder's occupancy from 20 to 28 warps per SM).
__device__ __forceinline__ int2 settle_cell(const Grid& g, const int* S, int n, int* list, float px, float py, float pz, float a, float rho,
                                            int& kind) {
  const int lane = threadIdx.x & 31;
  float mb = 3.0e38f;
  int mi = 0x7fffffff;
  float c0x = 0.f, c0y = 0.f, c0z = 0.f;
  for (int i = lane; i < n; i += 32) {
    float4 q = __ldg(g.sorted + S[i]);
    float dx = px - q.x, dy = py - q.y, dz = pz - q.z;
    float d = dx * dx + dy * dy + dz * dz;
    int id = __float_as_int(q.w);
    if (d < mb || (d == mb && id < mi)) { mb = d; mi = id; c0x = q.x; c0y = q.y; c0z = q.z; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ob = __shfl_xor_sync(0xffffffffu, mb, o);
    int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    float ox = __shfl_xor_sync(0xffffffffu, c0x, o), oy = __shfl_xor_sync(0xffffffffu, c0y, o), oz = __shfl_xor_sync(0xffffffffu, c0z, o);
    if (ob < mb || (ob == mb && oi < mi)) { mb = ob; mi = oi; c0x = ox; c0y = oy; c0z = oz; }
  }
  const float best = mb, dmin = sqrtf(mb);
  kind = 4;
  if (dmin * 0.9999f - rho > g.r_cap) return make_int2(mi, -1);  // far
  const float h_thr = 0.1f + rho * 1.001f + 1e-4f;
  bool search = !g.classify;
  int keep = 0;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    bool cand = false;
    int qi = 0;
    if (i < n) {
      qi = S[i];
      const float4 q = __ldg(g.sorted + qi);
      float dx = px - q.x, dy = py - q.y, dz = pz - q.z;
      float d = dx * dx + dy * dy + dz * dz;
      cand = can_beat(d, best, a, fabsf(q.x - c0x) + fabsf(q.y - c0y) + fabsf(q.z - c0z));
      if (cand && g.classify) {
        float4 nq = __ldg(g.tri_n + __float_as_int(q.w));
        float h = dx * nq.x + dy * nq.y + dz * nq.z;
        if (!(fabsf(h) > h_thr)) search = true;  // NaN normal (degenerate triangle) keeps the cell searchable
      }
    }
    unsigned m = __ballot_sync(0xffffffffu, cand);
    if (cand) {
      int slot = keep + __popc(m & ((1u << lane) - 1));
      if (list != nullptr && slot < LIST_CAP) list[slot] = qi;
    }
    keep += __popc(m);
  }
  search = __any_sync(0xffffffffu, search);
  __syncwarp();
  kind = 5;
  if (!search) return make_int2(mi, -1);  // certified transparent
  kind = 7;
  int2 rec = make_int2(mi, -2);
  if (list != nullptr) {
    int off = 0;
    if (lane == 0) off = atomicAdd(g.pool_used, keep);
    off = __shfl_sync(0xffffffffu, off, 0);
    if (off + keep <= g.pool_cap) {
      if (keep <= LIST_CAP) {
        for (int i = lane; i < keep; i += 32) g.pool[off + i] = __ldg(g.sorted + list[i]);
      } else {  // longer than the scratch list: the same filter once more, straight into the pool
        int k = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
          const int i = i0 + lane;
          bool cand = false;
          float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < n) {
            q = __ldg(g.sorted + S[i]);
            float dx = px - q.x, dy = py - q.y, dz = pz - q.z;
            cand = can_beat(dx * dx + dy * dy + dz * dz, best, a, fabsf(q.x - c0x) + fabsf(q.y - c0y) + fabsf(q.z - c0z));
          }
          const unsigned m = __ballot_sync(0xffffffffu, cand);
          if (cand) g.pool[off + k + __popc(m & ((1u << lane) - 1))] = q;
          k += __popc(m);
        }
      }
      rec = make_int2(off, keep);
      kind = 6;
    }
  }
  __syncwarp();
  return rec;
}

#ifndef DSN_LONG_CAP
#define DSN_LONG_CAP 8192
#endif
constexpr int LONG_CAP = DSN_LONG_CAP;    // longest list the second build level writes straight into the pool

// A table cell whose candidates do not fit the builder's shared-memory buffers (more than BUF_CAP in the superset or more than
// LIST_CAP in the list: cells that see a dense cluster or a whole ring of centroids).  A ball scan per LOOKUP through such a
// cell visits every centroid of every grid cell the ball touches, thousands of them, and 0.5 % of the lookups cost half of the
// search kernels' instructions that way; the list -- the centroids that can win somewhere in the cell -- is several times
// shorter, so it is worth one or two more scans at build time: count the centroids that pass the filter against c0 (unless the
// caller's last scan was that very count), then the same scan again writing them straight into the pool.  All lanes call this; `seed` = c0, the exact nearest centroid of the cell
// centre (both callers have it: settle_cell's, or the reduced one of the builder's second scan).  No transparency proof: the
// cell stays searchable.
__device__ __forceinline__ int2 long_list(const Grid& g, float px, float py, float pz, float a, float rho, int seed, int known_total, float known_best, int& kind) {
  const int lane = threadIdx.x & 31;
  const int mi = seed;
  const float c0x = __ldg(g.cent + 3 * seed), c0y = __ldg(g.cent + 3 * seed + 1), c0z = __ldg(g.cent + 3 * seed + 2);
  const float mb = known_total >= 0 ? known_best : (px - c0x) * (px - c0x) + (py - c0y) * (py - c0y) + (pz - c0z) * (pz - c0z);
  const float radius = (sqrtf(mb) + 2.0f * rho) * 1.0002f + 1e-5f;
  const float best = mb;
  int total = known_total;  // >= 0: the caller has just run this very scan (same ball, same filter) and counted
  if (total < 0) {
    total = 0;
    warp_scan_ball_flat<4>(g, px, py, pz, radius, [&](float4 q, int, bool valid) {
      const float dx = px - q.x, dy = py - q.y, dz = pz - q.z;
      const bool keep = valid && can_beat(dx * dx + dy * dy + dz * dz, best, a, fabsf(q.x - c0x) + fabsf(q.y - c0y) + fabsf(q.z - c0z));
      total += __popc(__ballot_sync(0xffffffffu, keep));
    });
  }
  kind = 7;
  if (total > LONG_CAP) return make_int2(mi, -2);
  int off = 0;
  if (lane == 0) off = atomicAdd(g.pool_used, total);
  off = __shfl_sync(0xffffffffu, off, 0);
  if (off + total > g.pool_cap) return make_int2(mi, -2);
  int n = 0;
  warp_scan_ball_flat<4>(g, px, py, pz, radius, [&](float4 q, int, bool valid) {
    const float dx = px - q.x, dy = py - q.y, dz = pz - q.z;
    const bool keep = valid && can_beat(dx * dx + dy * dy + dz * dz, best, a, fabsf(q.x - c0x) + fabsf(q.y - c0y) + fabsf(q.z - c0z));
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    const int slot = n + __popc(m & ((1u << lane) - 1));
    if (keep && slot < total) g.pool[off + slot] = q;
    n += __popc(m);
  });
  __syncwarp();
  // the count came from another instance of the same scan: should the two ever disagree (a differently contracted fma), a
  // longer list is not trusted (scan fallback) and a shorter one is padded with c0 (a duplicate candidate is harmless)
  if (n > total) return make_int2(mi, -2);
  for (int i = n + lane; i < total; i += 32) g.pool[off + i] = make_float4(c0x, c0y, c0z, __int_as_float(mi));
  __syncwarp();
  kind = 6;
  return make_int2(off, total);
}

// Build requested cells, one warp per REQUESTED ENUMERATION CELL (LEVEL 1) or per left-over table cell (LEVEL 2).
// LEVEL 1: one ball scan collects the candidate superset of the whole 4 cm cell (reference = jump-flooded seed; any reference
// is valid because the nearest centroid of a point beats every other one).  Posed mesh: if the proof certifies the whole cell,
// its requested table cells are settled at once; otherwise (and for the canonical mesh) each requested table cell of it is
// settled from the buffered superset (a table cell's candidates are a subset of its parent's) without another scan.
// LEVEL 2: table cells still unsettled (parent built by an earlier call, or its superset overflowed): own scan.
template <int LEVEL>
__global__ void __launch_bounds__(BUILD_WARPS * 32) build_cells_kernel(Grid g) {
  __shared__ int buf[BUILD_WARPS][BUF_CAP];
  __shared__ int lst[BUILD_WARPS][LIST_CAP];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n_req = g.pool_used[LEVEL == 1 ? 2 : 3];
  const float tcell = 1.0f / g.tinv;
  const float ecell = LEVEL == 1 ? g.cell : tcell;
  const float a = 0.5f * ecell * 1.0002f + 2e-6f;  // half edge of the cell, rounded up (covers the rounding of the cell lookup)
  const float at = 0.5f * tcell * 1.0002f + 2e-6f;
  const float rho = LEVEL == 1 ? 2.0f * g.thalf_diag : g.thalf_diag;
  const int lnx = LEVEL == 1 ? g.nx : g.tnx, lny = LEVEL == 1 ? g.ny : g.tny;
  for (;;) {
    // cells differ in cost by two orders of magnitude (deep inside the body a cell sees a whole ring of centroids): warps
    // fetch their next cell from a work counter instead of striding
    int r = 0;
    if (lane == 0) r = atomicAdd(g.pool_used + 15 + LEVEL, 1);
    r = __shfl_sync(0xffffffffu, r, 0);
    if (r >= n_req) break;
    const int cell = LEVEL == 1 ? g.ereq[r] : g.req2[r];
    const int tx = cell % lnx, ty = (cell / lnx) % lny, tz = cell / (lnx * lny);
    const int par = LEVEL == 1 ? cell : ((tz >> 1) * g.ny + (ty >> 1)) * g.nx + (tx >> 1);
    int kind = 4, visits = 0;
    if (LEVEL == 2 && g.classify && g.estate[par] == 2) {  // the whole enumeration cell is certified transparent
      if (lane == 0) { g.trec[cell] = make_int2(0, -1); g.tstate[cell] = 2; }
      if (g.debug && lane == 0) atomicAdd(g.pool_used + 4, 1);
      continue;
    }
    const float px = g.ox + (tx + 0.5f) * ecell, py = g.oy + (ty + 0.5f) * ecell, pz = g.oz + (tz + 0.5f) * ecell;
    const int cref = __ldg(g.enum_seed + par);
    const float rx = __ldg(g.cent + 3 * cref), ry = __ldg(g.cent + 3 * cref + 1), rz = __ldg(g.cent + 3 * cref + 2);
    const float dref2 = (px - rx) * (px - rx) + (py - ry) * (py - ry) + (pz - rz) * (pz - rz);
    float mb = dref2;
    int mi = cref;
    int nbuf = 0;  // warp uniform: the candidates are appended through ballots (the flat scan keeps the warp converged)
    warp_scan_ball_flat<LEVEL == 2 ? 4 : 2>(g, px, py, pz, (sqrtf(dref2) + 2.0f * rho) * 1.0002f + 1e-5f, [&](float4 q, int j, bool valid) {
      float dx = px - q.x, dy = py - q.y, dz = pz - q.z;
      float d = dx * dx + dy * dy + dz * dz;
      bool keep = false;
      if (valid) {
        ++visits;
        if (d < mb) { mb = d; mi = __float_as_int(q.w); }
        keep = can_beat(d, dref2, a, fabsf(q.x - rx) + fabsf(q.y - ry) + fabsf(q.z - rz));
      }
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (keep) {
        const int slot = nbuf + __popc(m & ((1u << lane) - 1));
        if (slot < BUF_CAP) buf[w][slot] = j;
      }
      nbuf += __popc(m);
    });
    __syncwarp();
    if (nbuf > BUF_CAP) {
      // superset too long against the seed: take the centre's true nearest centroid (found by the scan above) as the
      // reference and scan once more (smaller ball, tighter filter)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, mb, o);
        int oi = __shfl_xor_sync(0xffffffffu, mi, o);
        if (ob < mb || (ob == mb && oi < mi)) { mb = ob; mi = oi; }
      }
      const float c0x = __ldg(g.cent + 3 * mi), c0y = __ldg(g.cent + 3 * mi + 1), c0z = __ldg(g.cent + 3 * mi + 2);
      __syncwarp();
      nbuf = 0;
      warp_scan_ball_flat<LEVEL == 2 ? 4 : 2>(g, px, py, pz, (sqrtf(mb) + 2.0f * rho) * 1.0002f + 1e-5f, [&](float4 q, int j, bool valid) {
        float dx = px - q.x, dy = py - q.y, dz = pz - q.z;
        bool keep = false;
        if (valid) {
          ++visits;
          keep = can_beat(dx * dx + dy * dy + dz * dz, mb, a, fabsf(q.x - c0x) + fabsf(q.y - c0y) + fabsf(q.z - c0z));
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
          const int slot = nbuf + __popc(m & ((1u << lane) - 1));
          if (slot < BUF_CAP) buf[w][slot] = j;
        }
        nbuf += __popc(m);
      });
      __syncwarp();
    }
    const bool overflow = nbuf > BUF_CAP;
    if (LEVEL == 1) {
      // the enumeration cell itself: proof only (never a list)
      int2 prec = make_int2(0, -2);
      if (!overflow && g.classify) prec = settle_cell(g, buf[w], nbuf, nullptr, px, py, pz, a, rho, kind);
      const bool certified = prec.y == -1;
      if (lane == 0) g.estate[cell] = certified ? 2 : 3;
      // its requested table cells
      for (int k = 0; k < 8; ++k) {
        const int cx = 2 * tx + (k & 1), cy = 2 * ty + ((k >> 1) & 1), cz = 2 * tz + (k >> 2);
        const int child = (cz * g.tny + cy) * g.tnx + cx;
        if (g.tstate[child] != 1) continue;  // warp uniform
        int2 rec = make_int2(0, -1);
        int ck = 4;
        if (!certified) {
          if (overflow) {  // left to LEVEL 2
            if (lane == 0) g.req2[atomicAdd(g.pool_used + 3, 1)] = child;
            continue;
          }
          rec = settle_cell(g, buf[w], nbuf, lst[w], g.ox + (cx + 0.5f) * tcell, g.oy + (cy + 0.5f) * tcell, g.oz + (cz + 0.5f) * tcell, at,
                            g.thalf_diag, ck);
          if (rec.y == -2) {  // pool full: LEVEL 2 may still find room (long_list) or leaves the cell to the scans
            if (lane == 0) g.req2[atomicAdd(g.pool_used + 3, 1)] = child;
            continue;
          }
        }
        __syncwarp();
        if (lane == 0) {
          g.trec[child] = rec;
          g.tstate[child] = 2;
          if (g.debug) { atomicAdd(g.pool_used + ck, 1); if (rec.y > 0) atomicAdd(g.pool_used + 8, rec.y); }
        }
      }
    } else {
      int2 rec;
      if (overflow) { rec = make_int2(__shfl_sync(0xffffffffu, mi, 0), -2); kind = 7; }  // (mi: a centroid seen by lane 0's rows; any centroid seeds a scan)
      else rec = settle_cell(g, buf[w], nbuf, lst[w], px, py, pz, a, rho, kind);
      if (rec.y == -2) rec = long_list(g, px, py, pz, a, rho, rec.x, overflow ? nbuf : -1, mb, kind);
      __syncwarp();
      if (lane == 0) {
        g.trec[cell] = rec;
        g.tstate[cell] = 2;
        if (g.debug) { atomicAdd(g.pool_used + kind, 1); if (rec.y > 0) atomicAdd(g.pool_used + 8, rec.y); }
      }
    }
    if (g.debug) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) visits += __shfl_xor_sync(0xffffffffu, visits, o);
      if (lane == 0) atomicAdd(g.pool_used + (LEVEL == 1 ? 9 : 10), visits);
    }
  }
}

// Exact nearest centroid of a point: squared L2 accumulated as d0*d0, fma(d1,d1,.), fma(d2,d2,.) and strict '<' with
// lowest index on ties -- the arithmetic of pytorch3d 0.4.0 knn_points(K=1) as called at utils/render_utils.py:95.
// (a) over a candidate list of a built table cell
__device__ __forceinline__ int list_nearest(const Grid& g, int off, int cnt, float px, float py, float pz) {
  float best = 3.0e38f;
  int besti = 0x7fffffff;
  const float4* __restrict__ L = g.pool + off;
#pragma unroll 4
  for (int i = 0; i < cnt; ++i) {
    float4 c = __ldg(L + i);
    float dx = xsub(px, c.x), dy = xsub(py, c.y), dz = xsub(pz, c.z);
    float d = xfma(dz, dz, xfma(dy, dy, xmul(dx, dx)));
    int id = __float_as_int(c.w);
    if (d < best || (d == best && id < besti)) { best = d; besti = id; }
  }
  return besti;
}
// (b) ball scan through the enumeration grid, seeded with any centroid `hint`: visits every grid row that intersects the
// ball of the current best radius, so it returns the same index as a full scan.
__device__ __forceinline__ int scan_nearest(const Grid& g, float px, float py, float pz, int hint) {
  float best, rho2;
  int besti = hint;
  {
    float dx = xsub(px, __ldg(g.cent + 3 * hint)), dy = xsub(py, __ldg(g.cent + 3 * hint + 1)), dz = xsub(pz, __ldg(g.cent + 3 * hint + 2));
    best = xfma(dz, dz, xfma(dy, dy, xmul(dx, dx)));
    rho2 = best * 1.0001f + 1e-12f;
  }
  scan_ball(g, px, py, pz, sqrtf(rho2) * 1.0001f, rho2, [&](float4 c) {
    float dx = xsub(px, c.x), dy = xsub(py, c.y), dz = xsub(pz, c.z);
    float d = xmul(dx, dx);
    d = xfma(dy, dy, d);
    d = xfma(dz, dz, d);
    int id = __float_as_int(c.w);
    if (d < best || (d == best && id < besti)) {
      best = d;
      besti = id;
      rho2 = fminf(rho2, best * 1.0001f + 1e-12f);
    }
  });
  return besti;
}
// (c) exhaustive
__device__ __forceinline__ int brute_nearest(const float* __restrict__ cent, int F, float px, float py, float pz) {
  float best = 3.0e38f;
  int idx = 0;
  for (int f = 0; f < F; ++f) {
    float dx = xsub(px, cent[3 * f]), dy = xsub(py, cent[3 * f + 1]), dz = xsub(pz, cent[3 * f + 2]);
    float d = xfma(dz, dz, xfma(dy, dy, xmul(dx, dx)));
    if (d < best) { best = d; idx = f; }
  }
  return idx;
}
// Nearest centroid of a point whose table cell has been built; -1 = no search needed (provably transparent / far).
__device__ __forceinline__ int table_nearest(const Grid& g, int cell, float px, float py, float pz) {
  if (cell < 0) return -1;
  const int2 rec = g.trec[cell];
  if (rec.y == -1) return -1;
  if (rec.y >= 0) return list_nearest(g, rec.x, rec.y, px, py, pz);
  return scan_nearest(g, px, py, pz, rec.x);
}

// The same search by a whole warp (all lanes pass the same point and record, all lanes get the result): 32 candidates per
// step and a butterfly argmin for a candidate list, warp_scan_ball_flat where the cell has no list (pool full).  Same
// arithmetic, strict '<', lowest index on ties: the result is the one list_nearest / scan_nearest return.  For lookups
// through lists longer than COOP_LIST (cells that see a dense cluster or a whole ring of centroids): a lane walking such a
// list alone keeps its whole warp (and, in sample_warp_kernel, its block) waiting.
#ifndef DSN_COOP_LIST
#define DSN_COOP_LIST 256
#endif
constexpr int COOP_LIST = DSN_COOP_LIST;
__device__ __forceinline__ int warp_nearest(const Grid& g, int2 rec, float px, float py, float pz) {
  const int lane = threadIdx.x & 31;
  float best = 3.0e38f;
  int besti = 0x7fffffff;
  auto take = [&](float4 q) {
    const float dx = xsub(px, q.x), dy = xsub(py, q.y), dz = xsub(pz, q.z);
    float d = xmul(dx, dx);
    d = xfma(dy, dy, d);
    d = xfma(dz, dz, d);
    const int id = __float_as_int(q.w);
    if (d < best || (d == best && id < besti)) { best = d; besti = id; }
  };
  if (rec.y >= 0) {
    const float4* __restrict__ L = g.pool + rec.x;
    for (int k = lane; k < rec.y; k += 32) take(__ldg(L + k));
  } else {
    const float dx = xsub(px, __ldg(g.cent + 3 * rec.x)), dy = xsub(py, __ldg(g.cent + 3 * rec.x + 1)), dz = xsub(pz, __ldg(g.cent + 3 * rec.x + 2));
    const float d = xfma(dz, dz, xfma(dy, dy, xmul(dx, dx)));
    warp_scan_ball_flat(g, px, py, pz, sqrtf(d * 1.0001f + 1e-12f) * 1.0001f, [&](float4 q, int, bool valid) { if (valid) take(q); });
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob < best || (ob == best && oi < besti)) { best = ob; besti = oi; }
  }
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
constexpr int GG_TILES = 48;     // direction tiles per axis
constexpr int GG_CAP = 384;      // vertices per tile (a gamma = 5 cm cone holds ~75 vertices of an SMPL-sized mesh); fuller tiles fall back to the exhaustive scan

// Every ray's LINE passes through the same point o0 (the reference uses the first ray's origin for all rays, and its test
// |q|^2 - (q.u)^2 < gamma^2 has no t >= 0 restriction), so "vertex v is within gamma of the line" only depends on the line's
// direction: it must lie inside the cone of half angle asin(gamma / |q_v|) around q_v.  Directions are binned in a gnomonic
// chart (a, b) = (u.e1, u.e2) / (u.c) around c = direction of the centre of the vertex box; each vertex is entered into
// every tile its (conservatively bounded) cone touches, a ray tests the vertices of its own tile only -- with the reference's
// arithmetic, so near/far stay bit-identical.  ok = 0 (camera inside / very close to the mesh) selects the exhaustive scan.
struct GgFrame {
  float c[3], e1[3], e2[3];
  float a0, b0, inv_da, inv_db;
  int ok;
};

__global__ void gg_frame_kernel(const unsigned* __restrict__ qbox, float gamma_pad, GgFrame* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float lo[3], hi[3];
  for (int k = 0; k < 3; ++k) { lo[k] = key2f(qbox[k]) - gamma_pad; hi[k] = key2f(qbox[3 + k]) + gamma_pad; }
  GgFrame f;
  float cx = 0.5f * (lo[0] + hi[0]), cy = 0.5f * (lo[1] + hi[1]), cz = 0.5f * (lo[2] + hi[2]);
  float n = sqrtf(cx * cx + cy * cy + cz * cz);
  f.ok = n > 1e-3f;
  if (!f.ok) n = 1.f;
  f.c[0] = cx / n; f.c[1] = cy / n; f.c[2] = cz / n;
  // any unit vector not parallel to c
  float hx = fabsf(f.c[0]) < 0.6f ? 1.f : 0.f, hy = hx == 0.f ? 1.f : 0.f;
  float e1x = f.c[1] * 0.f - f.c[2] * hy, e1y = f.c[2] * hx - f.c[0] * 0.f, e1z = f.c[0] * hy - f.c[1] * hx;
  float en = sqrtf(e1x * e1x + e1y * e1y + e1z * e1z);
  f.e1[0] = e1x / en; f.e1[1] = e1y / en; f.e1[2] = e1z / en;
  f.e2[0] = f.c[1] * f.e1[2] - f.c[2] * f.e1[1]; f.e2[1] = f.c[2] * f.e1[0] - f.c[0] * f.e1[2]; f.e2[2] = f.c[0] * f.e1[1] - f.c[1] * f.e1[0];
  // chart extent = box of the 8 projected corners of the padded vertex box (a projective map keeps convex hulls)
  float amin = 3e38f, amax = -3e38f, bmin = 3e38f, bmax = -3e38f;
  for (int k = 0; k < 8; ++k) {
    float x = (k & 1) ? hi[0] : lo[0], y = (k & 2) ? hi[1] : lo[1], z = (k & 4) ? hi[2] : lo[2];
    float s = x * f.c[0] + y * f.c[1] + z * f.c[2];
    if (!(s > 0.05f * n)) f.ok = 0;  // a corner is beside / behind the chart plane
    float a = (x * f.e1[0] + y * f.e1[1] + z * f.e1[2]) / s, b = (x * f.e2[0] + y * f.e2[1] + z * f.e2[2]) / s;
    amin = fminf(amin, a); amax = fmaxf(amax, a); bmin = fminf(bmin, b); bmax = fmaxf(bmax, b);
  }
  float da = (amax - amin) * 1.001f + 1e-6f, db = (bmax - bmin) * 1.001f + 1e-6f;
  f.a0 = amin - 0.0005f * da; f.b0 = bmin - 0.0005f * db;
  f.inv_da = (float)GG_TILES / da; f.inv_db = (float)GG_TILES / db;
  if (!(da < 8.f && db < 8.f)) f.ok = 0;
  *out = f;
}

__global__ void gg_bin_kernel(const float4* __restrict__ vq, int V, float gamma_pad, const GgFrame* __restrict__ fr, int* __restrict__ counts,
                              int* __restrict__ lists, int* __restrict__ bad) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const GgFrame f = *fr;
  if (!f.ok) return;
  float4 q = vq[v];
  float qn = sqrtf(q.w);
  float s = q.x * f.c[0] + q.y * f.c[1] + q.z * f.c[2];
  if (!(qn > 2.f * gamma_pad) || !(s > 0.f)) { atomicExch(bad, 1); return; }
  float cth = fminf(1.f, s / qn);
  float theta = acosf(cth), alpha = asinf(fminf(1.f, gamma_pad / qn));
  if (!(theta + alpha < 1.2f)) { atomicExch(bad, 1); return; }
  // the gnomonic chart stretches angles by at most 1 / cos^2 within the cone
  float cc = cosf(theta + alpha);
  float r = alpha / (cc * cc) * 1.02f + 1e-6f;
  float a = (q.x * f.e1[0] + q.y * f.e1[1] + q.z * f.e1[2]) / s, b = (q.x * f.e2[0] + q.y * f.e2[1] + q.z * f.e2[2]) / s;
  int i0 = max(0, (int)floorf((a - r - f.a0) * f.inv_da)), i1 = min(GG_TILES - 1, (int)floorf((a + r - f.a0) * f.inv_da));
  int j0 = max(0, (int)floorf((b - r - f.b0) * f.inv_db)), j1 = min(GG_TILES - 1, (int)floorf((b + r - f.b0) * f.inv_db));
  for (int j = j0; j <= j1; ++j)
    for (int i = i0; i <= i1; ++i) {
      int t = j * GG_TILES + i;
      int slot = atomicAdd(counts + t, 1);
      if (slot < GG_CAP) lists[t * GG_CAP + slot] = v;
    }
}

// One warp per ray: the lanes share the vertices of the ray's direction tile (or all vertices when the tile is overfull /
// the chart is unusable), then reduce min / max.  min and max are order independent, so the result equals the reference's.
__global__ void __launch_bounds__(GG_THREADS) gg_bounds_kernel(const float4* __restrict__ vq, int V, const unsigned* __restrict__ qbox,
                                                               const GgFrame* __restrict__ fr, const int* __restrict__ counts,
                                                               const int* __restrict__ lists, const int* __restrict__ bad,
                                                               const float* __restrict__ ray_d, const float* __restrict__ near_in,
                                                               const float* __restrict__ far_in, int64_t R, float gamma2, float gamma,
                                                               float* __restrict__ near_out, float* __restrict__ far_out) {
  const int64_t r = ((int64_t)blockIdx.x * GG_THREADS + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float dx = ray_d[3 * r], dy = ray_d[3 * r + 1], dz = ray_d[3 * r + 2];
  const float norm = xnorm3(v3(dx, dy, dz));
  const float ux = xdiv(dx, norm), uy = xdiv(dy, norm), uz = xdiv(dz, norm);
  float zmin = 99999.0f, zmax = -99999.0f;
  bool any = false;
  auto test = [&](float4 q) {
    float z0 = xadd(xadd(xmul(q.x, ux), xmul(q.y, uy)), xmul(q.z, uz));
    float tmp = xsub(q.w, xmul(z0, z0));
    if (tmp < gamma2) {
      float del = xsqrt(xsub(gamma2, tmp));
      zmin = fminf(zmin, xsub(z0, del));
      zmax = fmaxf(zmax, xadd(z0, del));
      any = true;
    }
  };
  const GgFrame f = *fr;
  bool full = !f.ok || *bad != 0;
  int tile = -1;
  if (!full) {
    float s = ux * f.c[0] + uy * f.c[1] + uz * f.c[2];
    if (fabsf(s) > 1e-6f) {
      float a = (ux * f.e1[0] + uy * f.e1[1] + uz * f.e1[2]) / s, b = (ux * f.e2[0] + uy * f.e2[1] + uz * f.e2[2]) / s;
      float fa = (a - f.a0) * f.inv_da, fb = (b - f.b0) * f.inv_db;
      if (fa >= 0.f && fb >= 0.f && fa < (float)GG_TILES && fb < (float)GG_TILES) tile = (int)fb * GG_TILES + (int)fa;
    }
  }
  int n = 0;
  if (tile >= 0) {
    n = __ldg(counts + tile);
    if (n > GG_CAP) full = true;  // overfull tile
  }
  if (full) {
    // exhaustive: cull by the padded vertex box first (the line can be within gamma of a vertex only if it meets the box)
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
    if (ok && (t0 <= t1 * (1.0f + 1e-5f) + 1e-5f))
      for (int i = lane; i < V; i += 32) test(__ldg(vq + i));
  } else {
    const int* __restrict__ L = lists + tile * GG_CAP;
    for (int i = lane; i < n; i += 32) test(__ldg(vq + __ldg(L + i)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    zmin = fminf(zmin, __shfl_xor_sync(0xffffffffu, zmin, o));
    zmax = fmaxf(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
  }
  any = __any_sync(0xffffffffu, any);
  if (lane != 0) return;
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
  // early-ray-termination mode: the active list is split into `waves` regions of `region` entries by sample index / wave_size
  int waves; int wave_size; int64_t region; unsigned long long* wave_counters;
};

#ifndef DSN_WARP_THREADS
#define DSN_WARP_THREADS 256
#endif
constexpr int WARP_THREADS = DSN_WARP_THREADS;

__device__ __forceinline__ void sample_position(const WarpArgs& a, int64_t s, float& px, float& py, float& pz) {
  int64_t r;
  int i;
  if (s < 0x7fffffffLL) { unsigned q = (unsigned)s / (unsigned)a.N; r = q; i = (int)((unsigned)s - q * (unsigned)a.N); }  // (64-bit division is ~5x dearer)
  else { r = s / a.N; i = (int)(s - r * a.N); }
  float z = a.z_in ? a.z_in[s] : sample_z(a.near[r], a.far[r], __ldg(a.tvals + i));
  px = xadd(a.ray_o[3 * r], xmul(a.ray_d[3 * r], z));
  py = xadd(a.ray_o[3 * r + 1], xmul(a.ray_d[3 * r + 1], z));
  pz = xadd(a.ray_o[3 * r + 2], xmul(a.ray_d[3 * r + 2], z));
}

// pass 0: request the table cell of every sample (the cells are then built by build_cells_kernel).  One thread walks
// MARK_SPT consecutive samples of one ray: the ray is loaded once, no per-sample division, and a run of samples in the same
// cell (8 mm steps through 2 cm cells) is requested once.
constexpr int MARK_SPT = 8;
__global__ void __launch_bounds__(WARP_THREADS) mark_samples_kernel(WarpArgs a, Grid g) {
  const int chunks = (a.N + MARK_SPT - 1) / MARK_SPT;  // per ray
  const int64_t t = (int64_t)blockIdx.x * WARP_THREADS + threadIdx.x;
  const int64_t r = t < 0x7fffffffLL ? (int64_t)((unsigned)t / (unsigned)chunks) : t / chunks;  // (64-bit division is ~5x dearer)
  const int i0 = (int)(t - r * chunks) * MARK_SPT;
  const bool live = r < a.R;
  float ox = 0.f, oy = 0.f, oz = 0.f, dx = 0.f, dy = 0.f, dz = 0.f, near = 0.f, far = 0.f;
  if (live) {
    ox = a.ray_o[3 * r]; oy = a.ray_o[3 * r + 1]; oz = a.ray_o[3 * r + 2];
    dx = a.ray_d[3 * r]; dy = a.ray_d[3 * r + 1]; dz = a.ray_d[3 * r + 2];
    if (!a.z_in) { near = a.near[r]; far = a.far[r]; }
  }
  int prev = -1;
#pragma unroll 1
  for (int k = 0; k < MARK_SPT; ++k) {
    const int i = i0 + k;
    int cell = -1;
    if (live && i < a.N) {
      const float z = a.z_in ? a.z_in[r * a.N + i] : sample_z(near, far, __ldg(a.tvals + i));
      cell = live_cell(g, xadd(ox, xmul(dx, z)), xadd(oy, xmul(dy, z)), xadd(oz, xmul(dz, z)));
      if (cell == prev) cell = -1; else prev = cell;
    }
    request_cell(g, cell);
  }
}
// same for explicit points (x, y, z, *) records, n given on the device or by the host
__global__ void __launch_bounds__(256) mark_points_kernel(const float4* __restrict__ pts, const unsigned long long* __restrict__ n_ptr, int64_t n_host, Grid g) {
  const int64_t n = n_ptr ? (int64_t)*n_ptr : n_host;
  for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x; t0 < n; t0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = t0 + threadIdx.x;
    int cell = -1;
    if (t < n) { float4 p = pts[t]; cell = live_cell(g, p.x, p.y, p.z); }
    request_cell(g, cell);
  }
}

// A block places WARP_SPT x WARP_THREADS consecutive samples (WARP_SPT per thread, strided by the block size so that every
// sub-round is coalesced) and then works its queue off with all threads: with one sample per thread only the ~16 % of the
// threads that hold a queue entry work in phase 2 while the other warps of the block wait at the barrier (ncu: 41 % of the
// stall samples) and hold their warp slots.
constexpr int WARP_SPT = 4;
constexpr int WARP_SPB = WARP_THREADS * WARP_SPT;
__global__ void __launch_bounds__(WARP_THREADS) sample_warp_kernel(WarpArgs a, Grid g) {
  __shared__ unsigned short queue[WARP_SPB];
  __shared__ int qn;
  __shared__ int bin_count[16], bin_start[16];
  __shared__ unsigned char flag[WARP_SPB];
  constexpr int LONG_SLOTS = 256;     // cooperative searches per block (further bin-15 entries are searched per lane)
  __shared__ int long_idx[LONG_SLOTS];
  const int64_t P = a.R * a.N;
  const int64_t s0 = (int64_t)blockIdx.x * WARP_SPB;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 16) bin_count[threadIdx.x] = 0;
#pragma unroll
  for (int k = 0; k < WARP_SPT; ++k) flag[k * WARP_THREADS + threadIdx.x] = 0;
  __syncthreads();
  // phase 1: place the samples, look their table cells up.  Samples that need the exact search are queued, bucketed by the
  // length of their cell's candidate list (= work) so that the lanes of a warp in phase 2 do similar amounts of work.
  int my_bin[WARP_SPT];
#pragma unroll
  for (int k = 0; k < WARP_SPT; ++k) {
    my_bin[k] = -1;
    const int64_t s = s0 + k * WARP_THREADS + threadIdx.x;
    if (s < P) {
      float px, py, pz;
      sample_position(a, s, px, py, pz);
      const int cell = live_cell(g, px, py, pz);
      if (cell >= 0) {
        const int cnt = g.trec[cell].y;
        if (cnt != -1) my_bin[k] = (cnt < 0 || cnt > COOP_LIST) ? 15 : min(14, cnt >> 3);
      }
    }
    if (my_bin[k] >= 0) atomicAdd(&bin_count[my_bin[k]], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = 15; b >= 0; --b) { bin_start[b] = acc; acc += bin_count[b]; }  // long lists first
    qn = acc;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < WARP_SPT; ++k)
    if (my_bin[k] >= 0) queue[atomicAdd(&bin_start[my_bin[k]], 1)] = (unsigned short)(k * WARP_THREADS + threadIdx.x);
  __syncthreads();
  const int n = qn;
  if (a.count_candidates && threadIdx.x == 0 && n) atomicAdd(a.counters + 2, (unsigned long long)n);
  // bin 15 (the head of the queue) = lookups through a list longer than COOP_LIST or through a cell without a list: one warp
  // each (warp_nearest) before the per-lane searches, so that no lane walks hundreds of candidates while its block waits
  const int n_long = min(bin_count[15], LONG_SLOTS);
  for (int i = threadIdx.x >> 5; i < n_long; i += WARP_THREADS / 32) {
    float px, py, pz;
    sample_position(a, s0 + queue[i], px, py, pz);
    const int idx = warp_nearest(g, g.trec[table_cell(g, px, py, pz)], px, py, pz);
    if (lane == 0) long_idx[i] = idx;
  }
  if (n_long) __syncthreads();  // (block uniform)
  for (int q0 = 0; q0 < n; q0 += WARP_THREADS) {
    const int qi = q0 + threadIdx.x;
    bool act = false;
    V3 xc = v3(0, 0, 0);
    int idx = -1, t = 0;
    if (qi < n) {
      t = queue[qi];
      float px, py, pz;
      sample_position(a, s0 + t, px, py, pz);
      idx = qi < n_long ? long_idx[qi] : table_nearest(g, table_cell(g, px, py, pz), px, py, pz);
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
    if (q0 + (int)(threadIdx.x & ~31u) < n) {  // warp-uniform: this warp holds queue entries
      const int nw = a.waves > 1 ? a.waves : 1;
      const int my_wave = (a.waves > 1 && act) ? (int)((s0 + t) % a.N) / a.wave_size : 0;
      for (int w = 0; w < nw; ++w) {
        const bool mine = act && my_wave == w;
        unsigned m = __ballot_sync(0xffffffffu, mine);
        if (!m) continue;
        int leader = __ffs(m) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(a.waves > 1 ? a.wave_counters + w : a.counters, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (mine) {
          unsigned long long slot = (unsigned long long)w * (unsigned long long)a.region + base + __popc(m & ((1u << lane) - 1));
          a.active[slot] = make_float4(xc.x, xc.y, xc.z, __int_as_float((int)(s0 + t)));
          a.active_tri[slot] = idx;
          flag[t] = 1;
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < WARP_SPT; ++k) {
    unsigned m = __ballot_sync(0xffffffffu, flag[k * WARP_THREADS + threadIdx.x] != 0);
    const int64_t s = s0 + k * WARP_THREADS + threadIdx.x;
    if (lane == 0 && s < P) a.sample_mask[s >> 5] = m;
  }
}

// exact nearest centroid of an arbitrary point without the lookup table: ball scan seeded by the enumeration cell's
// approximate nearest centroid; points outside the grid fall back to the exhaustive scan (rare: far from the mesh).
__device__ __forceinline__ int nearest_any(const Grid& g, const float* __restrict__ cent, int F, float px, float py, float pz) {
  float fx = (px - g.ox) * g.inv_cell, fy = (py - g.oy) * g.inv_cell, fz = (pz - g.oz) * g.inv_cell;
  if (fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)g.nx && fy < (float)g.ny && fz < (float)g.nz)
    return scan_nearest(g, px, py, pz, g.enum_seed[((int)fz * g.ny + (int)fy) * g.nx + (int)fx]);
  return brute_nearest(cent, F, px, py, pz);
}

// stand-alone warp op (dsnerf_warp_points): reports the reference's values for every point, transparent or not, so it
// does not use the lookup table (whose cells may hold "provably transparent" instead of a triangle).
__global__ void warp_points_kernel(const float* __restrict__ pts, int64_t P, const float* __restrict__ posed, const float* __restrict__ canon,
                                   const int* __restrict__ faces, Grid g, int F, const float* __restrict__ cent,
                                   float* __restrict__ xyz_cano, uint8_t* __restrict__ transparent, int* __restrict__ idx_out) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= P) return;
  float px = pts[3 * s], py = pts[3 * s + 1], pz = pts[3 * s + 2];
  const int idx = nearest_any(g, cent, F, px, py, pz);
  int i0 = faces[3 * idx], i1 = faces[3 * idx + 1], i2 = faces[3 * idx + 2];
  float u, v, h;
  project_point(v3(px, py, pz), ldv3(posed, i0), ldv3(posed, i1), ldv3(posed, i2), u, v, h);
  V3 xc = map_to_triangle(u, v, h, ldv3(canon, i0), ldv3(canon, i1), ldv3(canon, i2));
  xyz_cano[3 * s] = xc.x; xyz_cano[3 * s + 1] = xc.y; xyz_cano[3 * s + 2] = xc.z;
  if (transparent) transparent[s] = is_transparent(u, v, h) ? 1 : 0;
  if (idx_out) idx_out[s] = idx;
}

// ---- training-mode forward (Renderer.render with net.training) -------------------------------------
// Stratified jitter of uniform_sampling (utils/pts_utils.py:6-13): z' = lower + (upper - lower) * t_rand with
// mids = .5 (z[1:] + z[:-1]); t_rand is the caller's torch.rand draw.  One thread per sample.
__global__ void __launch_bounds__(256) jitter_z_kernel(const float* __restrict__ near, const float* __restrict__ far, const float* __restrict__ tvals,
                                                       const float* __restrict__ t_rand, int64_t R, int N, float* __restrict__ z_out) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= R * N) return;
  const int64_t r = s / N;
  const int i = (int)(s - r * N);
  const float n = near[r], f = far[r];
  const float z = sample_z(n, f, tvals[i]);
  const float lower = i > 0 ? xmul(0.5f, xadd(z, sample_z(n, f, tvals[i - 1]))) : z;
  const float upper = i + 1 < N ? xmul(0.5f, xadd(sample_z(n, f, tvals[i + 1]), z)) : z;
  z_out[s] = xadd(lower, xmul(xsub(upper, lower), t_rand[s]));
}

// With density noise a transparent sample has alpha = 1 - exp(-relu(0 + noise) dist) > 0 and its colour counts, so the
// reference's "network on every sample" (can_render.py:113-120 zeroes only the density) is kept literally: every sample
// goes to the active list (entry s = sample s) with its exact nearest triangle, and the mask bit says whether the
// compositor keeps the sample's density (bit set) or replaces it by 0 (transparent).
__global__ void __launch_bounds__(256) sample_warp_all_kernel(WarpArgs a, Grid g, int F, const float* __restrict__ cent) {
  const int64_t P = a.R * a.N;
  const int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x;
  bool opaque = false;
  if (s < P) {
    float px, py, pz;
    sample_position(a, s, px, py, pz);
    const int idx = nearest_any(g, cent, F, px, py, pz);
    int i0 = a.faces[3 * idx], i1 = a.faces[3 * idx + 1], i2 = a.faces[3 * idx + 2];
    float u, v, h;
    project_point(v3(px, py, pz), ldv3(a.posed, i0), ldv3(a.posed, i1), ldv3(a.posed, i2), u, v, h);
    const V3 xc = map_to_triangle(u, v, h, ldv3(a.canon, i0), ldv3(a.canon, i1), ldv3(a.canon, i2));
    a.active[s] = make_float4(xc.x, xc.y, xc.z, __int_as_float((int)s));
    a.active_tri[s] = idx;
    opaque = !is_transparent(u, v, h);
  }
  const unsigned m = __ballot_sync(0xffffffffu, opaque);
  if ((threadIdx.x & 31) == 0 && s < P) a.sample_mask[s >> 5] = m;
  if (threadIdx.x == 0) atomicAdd(a.counters, (unsigned long long)min((int64_t)256, P - s));
}

}  // namespace dsn
