// libdsnerf.so -- host side of the C ABI declared in include/dsnerf.h.
// Owns device buffers, stages weights/meshes, launches the kernels of the render path.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/dsnerf.h"
#include "geom.cuh"
#include "mlp_simt.cuh"
#include "mlp_tc2.cuh"
#include "mlp_tc.cuh"
#include "shade.cuh"
#include "light_tc.cuh"

using namespace dsn;

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct MeshGrid {
  Grid g{};
  int ncell = 0, ntab = 0;
  int launches = 0;  // kernels of the last build_grid
  DevBuf cent, tri_n, counts, start, cursor, sorted, rowmask, seed_a, seed_b, tstate, trec, pool, pool_used, req, efar, estate, ereq;
  void release() {
    DevBuf* b[] = {&cent, &tri_n, &counts, &start, &cursor, &sorted, &rowmask, &seed_a, &seed_b, &tstate, &trec, &pool, &pool_used, &req, &efar, &estate, &ereq};
    for (DevBuf* x : b) x->release();
  }
};

// state_dict order
enum {
  T_EMB = 0, T_S1_0_W, T_S1_0_B, T_S1_2_W, T_S1_2_B, T_S1_4_W, T_S1_4_B, T_S1_6_W, T_S1_6_B,
  T_S2_0_W, T_S2_0_B, T_S2_2_W, T_S2_2_B, T_S2_4_W, T_S2_4_B, T_DENS_W, T_DENS_B,
  T_RGB1_W, T_RGB1_B, T_RGB3_W, T_RGB3_B, T_L0_W, T_L0_B, T_L2_W, T_L2_B, T_L4_W, T_L4_B,
  T_P0_W, T_P0_B, T_P2_W, T_P2_B, T_P4_W, T_P4_B
};
const int kTensorSize[DSNERF_NUM_WEIGHT_TENSORS] = {
    500 * 8, 256 * 87, 256, 256 * 256, 256, 256 * 256, 256, 256 * 256, 256,
    256 * 319, 256, 256 * 256, 256, 256 * 256, 256, 256, 1,
    128 * 256, 128, 3 * 128, 3, 128 * 9, 128, 128 * 128, 128, 128, 1,
    64 * 92, 64, 64 * 64, 64, 16 * 64, 16};

}  // namespace

struct dsnerf_ctx {
  int device = 0;
  std::string err;
  int sm_count = 148;
  // ---- weights
  bool have_weights = false;
  std::vector<float> hw[DSNERF_NUM_WEIGHT_TENSORS];
  DevBuf wblob;      // fp32 SIMT/light layouts
  DevBuf bias0;      // per-frame folded first-layer bias (256)
  SimtWeights sw{};
  LightWeights lw{};
  TcWeights tw;      // fp16 hi/lo tensor-core layouts
  float rgb_probe_err = 0.f;  // mean |d essence| of a single-pass rgb head on the probe points (decides tw.rgb3)
  DevBuf light_w2;   // lighting layer 2 as a packed fp16 B operand
  // ---- mesh
  bool have_mesh = false;
  int F = 0, V = 0;
  std::vector<int32_t> h_faces;
  std::vector<float> h_canon;
  DevBuf faces, canon, posed, vq, gg_bins, normal_m;
  MeshGrid g_canon, g_posed;
  // ---- frame
  bool have_frame = false;
  float light_shift[3] = {0, 0, 0};
  int has_shift = 0;
  float rot[4] = {1, 0, 0, 1}, rot_center[2] = {0, 0};
  int has_rot = 0;
  // ---- workspace
  DevBuf tc_timing;
  DevBuf relu_scratch;                // mlp_tc2_kernel: ReLU bits of the tiles in flight
  unsigned int* h_tc_dbg = nullptr;   // mlp_tc2_kernel: watchdog record, mapped host memory (DSNERF_TC_WATCHDOG)
  int mlp_variant = 2;           // 2: two tiles in flight per CTA (mlp_tc2.cuh), 1: one tile (mlp_tc.cuh); DSNERF_MLP_VARIANT
  bool tc_watchdog = false;
  DevBuf near2, far2, raw, active, active_tri, active_cidx, canon_queue, ray_mask, mlp_a, mlp_g, tvals, counters, io;
  DevBuf ert_active, ert_tri, ert_state, ert_cnt;  // early-ray-termination mode only
  unsigned long long* h_ert = nullptr;              // pinned, 8 entries
  int tvals_n = 0;
  void* pin = nullptr;
  size_t pin_cap = 0;
  cudaEvent_t pin_free = nullptr;
  bool pin_busy = false;
  unsigned long long* h_counters = nullptr;  // pinned, 4 entries
  cudaEvent_t stats_ready = nullptr;
  cudaStream_t side = nullptr;   // dsnerf_set_frame builds the posed-mesh grid here while the caller's stream already runs the GG ray bounds
  cudaEvent_t ev_fork = nullptr, ev_grid = nullptr;
  bool grid_on_side = false;
  cudaStream_t copy = nullptr;   // dsnerf_render_host: ray upload next to the grid build that dsnerf_set_frame left on the caller's stream
  cudaEvent_t copy_done = nullptr;
  cudaStream_t down = nullptr;   // dsnerf_render_host_async: read-back of frame k next to the kernels of frame k + 1
  cudaEvent_t render_done = nullptr, host_done[2] = {nullptr, nullptr};
  bool host_pending[2] = {false, false};
  unsigned long long host_seq = 0;
  DevBuf io2[2];
  dsnerf_stats_t stats{};
  // ---- profiling
  int profile = 0;
  int ert_mode = 0;   // last render used early ray termination (stats come from h_ert)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
  double mlp_ms = 0;
  int64_t mlp_launches = 0;
};

namespace {

int fail(dsnerf_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return fail(ctx, DSNERF_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));             \
  } while (0)

#define CKL(what)                                                                                        \
  do {                                                                                                   \
    cudaError_t e_ = cudaGetLastError();                                                                 \
    if (e_ != cudaSuccess)                                                                               \
      return fail(ctx, DSNERF_ERR_CUDA, std::string("launch ") + what + ": " + cudaGetErrorString(e_));  \
  } while (0)

int ensure_pin(dsnerf_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->pin_cap) return 0;
  if (ctx->pin) cudaFreeHost(ctx->pin);
  ctx->pin = nullptr;
  ctx->pin_cap = 0;
  CK(cudaMallocHost(&ctx->pin, bytes + 4096));
  ctx->pin_cap = bytes + 4096;
  return 0;
}

// wait until the previous async upload out of the pinned staging buffer has been consumed
int pin_acquire(dsnerf_ctx* ctx, size_t bytes) {
  if (ctx->pin_busy) { CK(cudaEventSynchronize(ctx->pin_free)); ctx->pin_busy = false; }
  return ensure_pin(ctx, bytes);
}
int pin_release(dsnerf_ctx* ctx, cudaStream_t st) {
  CK(cudaEventRecord(ctx->pin_free, st));
  ctx->pin_busy = true;
  return 0;
}

// Radius beyond which a point cannot be non-transparent w.r.t. triangle T when T's centroid is
// its nearest centroid: |h| <= 0.1 and uv in [-4,5]^2 (utils/render_utils.py:103) confine the
// point to a box around T whose farthest corner from the centroid (uv = 1/3,1/3) is R_T.
float transparency_radius(const float* verts, const int32_t* faces, int F) {
  double best = 0;
  const double uvs[2] = {-4.0 - 1.0 / 3.0, 5.0 - 1.0 / 3.0};
  for (int f = 0; f < F; ++f) {
    const float* m0 = verts + 3 * faces[3 * f];
    const float* m1 = verts + 3 * faces[3 * f + 1];
    const float* m2 = verts + 3 * faces[3 * f + 2];
    double e0[3], e1[3];
    for (int k = 0; k < 3; ++k) { e0[k] = (double)m2[k] - m0[k]; e1[k] = (double)m1[k] - m0[k]; }
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        double s = 0;
        for (int k = 0; k < 3; ++k) { double d = uvs[a] * e0[k] + uvs[b] * e1[k]; s += d * d; }
        best = std::max(best, s);
      }
  }
  return (float)(sqrt(0.1 * 0.1 + best) * 1.002 + 1e-4);
}

// (Re)build the nearest-centroid grid of a mesh: centroids, counting sort by cell, jump-flooded seeds; the lookup table
// is only reset here (its cells are built on demand, see ensure_cells).  h_verts is the host copy (bbox and r_cap are
// computed on the host).
constexpr int kPoolEntries = 4 << 20;  // candidate-list pool per mesh (float4 entries, 64 MB)

int build_grid(dsnerf_ctx* ctx, MeshGrid& mg, const float* d_verts, const float* h_verts, int classify, cudaStream_t st) {
  const int F = ctx->F, V = ctx->V;
  float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
  for (int v = 0; v < V; ++v)
    for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], h_verts[3 * v + k]); hi[k] = std::max(hi[k], h_verts[3 * v + k]); }
  for (int k = 0; k < 3; ++k)
    if (!(lo[k] <= hi[k]) || !std::isfinite(lo[k]) || !std::isfinite(hi[k])) return fail(ctx, DSNERF_ERR_INVALID, "mesh vertices are not finite");
  float r_cap = transparency_radius(h_verts, ctx->h_faces.data(), F);
  // enumeration cell: 4 cm, grown until the lattice has at most 64 cells along x (row occupancy words) and ~3M cells
  float cell = 0.04f;
  double ext[3];
  for (;;) {
    double n = 1;
    for (int k = 0; k < 3; ++k) { ext[k] = (double)hi[k] - lo[k] + 2.0 * (r_cap + 2.0 * cell); n *= ceil(ext[k] / cell); }
    if (n <= 3.0e6 && ceil(ext[0] / cell) <= 64) break;
    cell *= 1.26f;
  }
  Grid& g = mg.g;
  g.cell = cell;
  g.inv_cell = 1.0f / cell;
  g.ox = lo[0] - (r_cap + 2 * cell); g.oy = lo[1] - (r_cap + 2 * cell); g.oz = lo[2] - (r_cap + 2 * cell);
  g.nx = (int)ceil(ext[0] / cell); g.ny = (int)ceil(ext[1] / cell); g.nz = (int)ceil(ext[2] / cell);
  g.tinv = 2.0f / cell;
  g.tnx = 2 * g.nx; g.tny = 2 * g.ny; g.tnz = 2 * g.nz;
  g.thalf_diag = 0.5f * cell * 0.8660254f * 1.0001f;
  g.r_cap = r_cap;
  g.classify = classify;
  g.F = F;
  mg.ncell = g.nx * g.ny * g.nz;
  mg.ntab = 8 * mg.ncell;
  const int nrow = g.ny * g.nz;
  CK(mg.cent.ensure(sizeof(float) * 3 * F));
  CK(mg.tri_n.ensure(sizeof(float4) * F));
  CK(mg.counts.ensure(sizeof(int) * mg.ncell));
  CK(mg.start.ensure(sizeof(int) * (mg.ncell + 1)));
  CK(mg.cursor.ensure(sizeof(int) * mg.ncell));
  CK(mg.sorted.ensure(sizeof(float4) * F));
  CK(mg.rowmask.ensure(sizeof(unsigned long long) * nrow));
  CK(mg.seed_a.ensure(sizeof(int) * mg.ncell));
  CK(mg.seed_b.ensure(sizeof(int) * mg.ncell));
  CK(mg.tstate.ensure((size_t)mg.ntab + 4));
  CK(mg.trec.ensure(sizeof(int2) * (size_t)mg.ntab));
  CK(mg.req.ensure(sizeof(int) * (size_t)mg.ntab));  // LEVEL-2 left-overs
  CK(mg.efar.ensure((size_t)mg.ncell + 4));
  CK(mg.estate.ensure((size_t)mg.ncell + 4));
  CK(mg.ereq.ensure(sizeof(int) * (size_t)mg.ncell));
  CK(mg.pool.ensure(sizeof(float4) * (size_t)kPoolEntries));
  CK(mg.pool_used.ensure(sizeof(int) * 32));
  g.cell_start = mg.start.as<int>();
  g.sorted = mg.sorted.as<float4>();
  g.row_mask = mg.rowmask.as<unsigned long long>();
  g.cent = mg.cent.as<float>();
  g.tri_n = mg.tri_n.as<float4>();
  g.tstate = mg.tstate.as<unsigned char>();
  g.trec = mg.trec.as<int2>();
  g.pool = mg.pool.as<float4>();
  g.pool_cap = (ctx->profile & 32) ? 4096 : kPoolEntries;  // profile bit 32 (tests): tiny pool, most cells fall back to the ball scan
  g.pool_used = mg.pool_used.as<int>();
  g.req2 = mg.req.as<int>();
  g.enum_far = mg.efar.as<unsigned char>();
  g.estate = mg.estate.as<unsigned char>();
  g.ereq = mg.ereq.as<int>();
  g.debug = (ctx->profile & 2) ? 1 : 0;
  int fb = (F + 255) / 256;
  centroid_kernel<<<fb, 256, 0, st>>>(d_verts, ctx->faces.as<int>(), F, mg.cent.as<float>(), mg.tri_n.as<float4>());
  CKL("centroid");
  CK(cudaMemsetAsync(mg.counts.p, 0, sizeof(int) * mg.ncell, st));
  CK(cudaMemsetAsync(mg.rowmask.p, 0, sizeof(unsigned long long) * nrow, st));
  CK(cudaMemsetAsync(mg.tstate.p, 0, (size_t)mg.ntab + 4, st));
  CK(cudaMemsetAsync(mg.estate.p, 0, (size_t)mg.ncell + 4, st));
  CK(cudaMemsetAsync(mg.pool_used.p, 0, sizeof(int) * 32, st));
  grid_count_kernel<<<fb, 256, 0, st>>>(g, mg.cent.as<float>(), F, mg.counts.as<int>(), mg.rowmask.as<unsigned long long>());
  CKL("grid_count");
  grid_scan_kernel<<<1, 1024, 0, st>>>(mg.counts.as<int>(), mg.ncell, mg.start.as<int>(), mg.cursor.as<int>());
  CKL("grid_scan");
  grid_fill_kernel<<<fb, 256, 0, st>>>(g, mg.cent.as<float>(), F, mg.cursor.as<int>(), mg.sorted.as<float4>());
  CKL("grid_fill");
  // seeds: jump flooding over the enumeration grid (approximate nearest centroid per cell; exactness comes from the scans)
  const int cb = (mg.ncell + 255) / 256;
  int* sa = mg.seed_a.as<int>();
  int* sb = mg.seed_b.as<int>();
  jfa_init_kernel<<<cb, 256, 0, st>>>(g, sa);
  CKL("jfa_init");
  int top = 1;
  while (top * 2 < std::max(g.nx, std::max(g.ny, g.nz))) top *= 2;
  for (int step = top; step >= 1; step >>= 1) {
    jfa_pass_kernel<<<cb, 256, 0, st>>>(g, step, sa, sb);
    CKL("jfa_pass");
    std::swap(sa, sb);
  }
  jfa_pass_kernel<<<cb, 256, 0, st>>>(g, 1, sa, sb);
  CKL("jfa_pass");
  std::swap(sa, sb);
  g.enum_seed = sa;
  mg.launches = 4 + 1 + 1 + 2 + 1;  // centroid/count/scan/fill, jfa_init, final jfa pass, 2 x enum_far (+ normal_matrix for the posed mesh)
  for (int step = top; step >= 1; step >>= 1) ++mg.launches;
  // enumeration cells provably farther than r_cap from every centroid (never requested, never searched)
  enum_far_rows_kernel<<<cb, 256, 0, st>>>(g, (int)ceil(r_cap / cell) + 2, sb);  // (sb: the jump-flooding buffer that does not hold the seeds)
  CKL("enum_far_rows");
  enum_far_kernel<<<cb, 256, 0, st>>>(g, (int)ceil(r_cap / cell) + 2, sb, mg.efar.as<unsigned char>());
  CKL("enum_far");
  return 0;
}

// Build the lookup-table cells requested by a mark kernel (launched by the caller through `mark`).
template <class Mark>
int ensure_cells(dsnerf_ctx* ctx, MeshGrid& mg, cudaStream_t st, Mark&& mark) {
  CK(cudaMemsetAsync(mg.pool_used.as<int>() + 1, 0, 3 * sizeof(int), st));
  CK(cudaMemsetAsync(mg.pool_used.as<int>() + 16, 0, 4 * sizeof(int), st));  // work counters of the two build levels, [18] / [19] work counter and queue length of canon_nearest_kernel
  mark();
  CKL("mark");
  build_cells_kernel<1><<<ctx->sm_count * 8, BUILD_WARPS * 32, 0, st>>>(mg.g);  // per requested enumeration cell
  CKL("build_cells<1>");
  build_cells_kernel<2><<<ctx->sm_count * 8, BUILD_WARPS * 32, 0, st>>>(mg.g);
  CKL("build_cells<2>");
  return 0;
}

struct BlobBuilder {
  std::vector<float> data;
  size_t add(size_t n) {
    size_t off = (data.size() + 63) / 64 * 64;
    data.resize(off + n, 0.f);
    return off;
  }
};

void rod2quat_host(const float* poses /*24x3*/, float* q /*92*/) {
  // model/spacenet.py:314-331 on joints 1..23
  for (int j = 1; j < 24; ++j) {
    const float* r = poses + 3 * j;
    float a0 = r[0] + 1e-16f, a1 = r[1] + 1e-16f, a2 = r[2] + 1e-16f;
    float angle = sqrtf(fmaf(a2, a2, fmaf(a1, a1, a0 * a0)));
    float c = cosf(angle / 2.f), s = sinf(angle / 2.f);
    float* o = q + 4 * (j - 1);
    o[0] = r[0] / angle * s; o[1] = r[1] / angle * s; o[2] = r[2] / angle * s; o[3] = c - 1.0f;
  }
}

void linear_host(const std::vector<float>& w, const std::vector<float>& b, int out, int in, const float* x, float* y, bool relu) {
  for (int o = 0; o < out; ++o) {
    float acc = 0.f;
    for (int i = 0; i < in; ++i) acc += w[(size_t)o * in + i] * x[i];
    acc += b[o];
    y[o] = relu ? std::max(acc, 0.f) : acc;
  }
}

// Does the single-pass fp16 rgb head (256 -> 128 -> 3) keep the colour inside the parity budget for THESE weights?
// Probe at weight-staging time: 64 fixed canonical points through the fp32 network on the host (code row 0, rest pose), then the
// 256 -> 128 layer once exactly and once with both operands rounded to fp16.  Returns the mean |d essence|.  With default-init
// weights (|h6| ~ 0.2) it is 8e-6 (rendered colour error 2.4e-5); with per-layer gains > 1, as trained checkpoints have
// (|h6| ~ 4), it reaches 1e-4 and the rendered error 3e-4 -- those weights get the 3-pass rgb head (+4 % kernel time).
float rgb_single_pass_probe(const std::vector<float>* hw) {
  const int sw_idx[7] = {T_S1_0_W, T_S1_2_W, T_S1_4_W, T_S1_6_W, T_S2_0_W, T_S2_2_W, T_S2_4_W};
  float q[92], h1[64], h2[64], pf[16], zero_pose[72] = {0};
  rod2quat_host(zero_pose, q);
  linear_host(hw[T_P0_W], hw[T_P0_B], 64, 92, q, h1, true);
  linear_host(hw[T_P2_W], hw[T_P2_B], 64, 64, h1, h2, true);
  linear_host(hw[T_P4_W], hw[T_P4_B], 16, 64, h2, pf, false);
  auto h16 = [](float v) { return __half2float(__float2half_rn(v)); };
  std::vector<float> w1h(hw[T_RGB1_W].size());
  for (size_t i = 0; i < w1h.size(); ++i) w1h[i] = h16(hw[T_RGB1_W][i]);
  uint32_t lcg = 12345u;
  auto rnd = [&]() { lcg = lcg * 1664525u + 1013904223u; return (float)(lcg >> 8) * (2.0f / 16777216.0f) - 1.0f; };
  double err = 0.0;
  const int n_probe = 64;
  std::vector<float> x(320), y(256), xh(256);
  for (int p = 0; p < n_probe; ++p) {
    float in0[87], pe[63];
    const float pt[3] = {rnd(), rnd(), rnd()};
    for (int c = 0; c < 3; ++c) pe[c] = pt[c];
    for (int k = 0; k < 10; ++k)
      for (int c = 0; c < 3; ++c) { pe[3 + 6 * k + c] = sinf(ldexpf(pt[c], k)); pe[6 + 6 * k + c] = cosf(ldexpf(pt[c], k)); }
    for (int j = 0; j < 8; ++j) in0[j] = hw[T_EMB][j];
    for (int j = 0; j < 63; ++j) in0[8 + j] = pe[j];
    for (int j = 0; j < 16; ++j) in0[71 + j] = pf[j];
    const float* cur = in0;
    int cur_n = 87;
    for (int l = 0; l < 7; ++l) {
      if (l == 4) {  // [h | PE]
        for (int j = 0; j < 256; ++j) x[j] = y[j];
        for (int j = 0; j < 63; ++j) x[256 + j] = pe[j];
        cur = x.data();
        cur_n = 319;
      }
      const std::vector<float>& w = hw[sw_idx[l]];
      const std::vector<float>& b = hw[sw_idx[l] + 1];
      float out[256];
      for (int o = 0; o < 256; ++o) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float* wr = w.data() + (size_t)o * cur_n;
        int i = 0;
        for (; i + 3 < cur_n; i += 4) { a0 += wr[i] * cur[i]; a1 += wr[i + 1] * cur[i + 1]; a2 += wr[i + 2] * cur[i + 2]; a3 += wr[i + 3] * cur[i + 3]; }
        for (; i < cur_n; ++i) a0 += wr[i] * cur[i];
        out[o] = std::max((a0 + a1) + (a2 + a3) + b[o], 0.f);
      }
      for (int o = 0; o < 256; ++o) y[o] = out[o];
      cur = y.data();
      cur_n = 256;
    }
    for (int j = 0; j < 256; ++j) xh[j] = h16(y[j]);
    float e_ex[3] = {0, 0, 0}, e_1[3] = {0, 0, 0};
    for (int o = 0; o < 128; ++o) {
      double a = 0.0, ah = 0.0;
      for (int i = 0; i < 256; ++i) { a += (double)hw[T_RGB1_W][(size_t)o * 256 + i] * y[i]; ah += (double)w1h[(size_t)o * 256 + i] * xh[i]; }
      const float r = std::max((float)a + hw[T_RGB1_B][o], 0.f), rh = std::max((float)ah + hw[T_RGB1_B][o], 0.f);
      for (int c = 0; c < 3; ++c) { e_ex[c] += hw[T_RGB3_W][(size_t)c * 128 + o] * r; e_1[c] += hw[T_RGB3_W][(size_t)c * 128 + o] * rh; }
    }
    for (int c = 0; c < 3; ++c) err += fabs((double)e_1[c] - e_ex[c]);
  }
  return (float)(err / (3.0 * n_probe));
}

int ensure_workspace(dsnerf_ctx* ctx, int64_t R, int N) {
  int64_t P = R * (int64_t)N;
  CK(ctx->near2.ensure(sizeof(float) * R));
  CK(ctx->far2.ensure(sizeof(float) * R));
  CK(ctx->raw.ensure(sizeof(float4) * P));
  CK(ctx->active.ensure(sizeof(float4) * (P + 128)));
  CK(ctx->active_tri.ensure(sizeof(int) * (P + 128)));
  CK(ctx->active_cidx.ensure(sizeof(int) * (P + 128)));
  CK(ctx->canon_queue.ensure(sizeof(int) * (P + 128)));
  CK(ctx->ray_mask.ensure(sizeof(unsigned) * (size_t)((P + 31) / 32 + 8)));
  CK(ctx->mlp_a.ensure(sizeof(float4) * (P + 128)));
  CK(ctx->mlp_g.ensure(sizeof(float4) * (P + 128)));
  CK(ctx->counters.ensure(sizeof(unsigned long long) * 4));
  return 0;
}

int ensure_tvals(dsnerf_ctx* ctx, int N, cudaStream_t st) {
  if (ctx->tvals_n == N) return 0;
  CK(ctx->tvals.ensure(sizeof(float) * N));
  if (int e = pin_acquire(ctx, sizeof(float) * N)) return e;
  float* t = reinterpret_cast<float*>(ctx->pin);
  for (int i = 0; i < N; ++i) t[i] = linspace01(i, N);
  CK(cudaMemcpyAsync(ctx->tvals.p, t, sizeof(float) * N, cudaMemcpyHostToDevice, st));
  if (int e = pin_release(ctx, st)) return e;
  ctx->tvals_n = N;
  return 0;
}

void profile_begin(dsnerf_ctx* ctx, cudaStream_t st, cudaEvent_t* a, cudaEvent_t* b) {
  *a = *b = nullptr;
  if (!(ctx->profile & 1)) return;
  cudaEventCreate(a);
  cudaEventCreate(b);
  cudaEventRecord(*a, st);
}
void profile_end(dsnerf_ctx* ctx, cudaStream_t st, cudaEvent_t a, cudaEvent_t b) {
  if (!(ctx->profile & 1) || !a) return;
  cudaEventRecord(b, st);
  ctx->pending.emplace_back(a, b);
}

// SpaceNet + gradient on the active list (count on the device or given by the host)
int launch_mlp(dsnerf_ctx* ctx, const unsigned long long* d_count, int64_t host_count, unsigned flags, int density_only, cudaStream_t st,
               const float4* list = nullptr) {
  const float4* active = list ? list : ctx->active.as<float4>();
  cudaEvent_t a, b;
  profile_begin(ctx, st, &a, &b);
  if ((flags & DSNERF_MLP_FP32_SIMT) || !ctx->tw.fp16_ok) {  // weights outside fp16 range: the fp32 kernel is the product path
    mlp_simt_kernel<<<ctx->sm_count, SIMT_THREADS, SIMT_SMEM, st>>>(ctx->sw, active, d_count, host_count,
                                                                    ctx->mlp_a.as<float4>(), ctx->mlp_g.as<float4>(), density_only);
    CKL("mlp_simt");
  } else {
    long long* timing = nullptr;
    if (ctx->profile & 4) {
      if (ctx->tc_timing.ensure(sizeof(long long) * 128) != cudaSuccess) return fail(ctx, DSNERF_ERR_CUDA, "timing buffer");
      timing = ctx->tc_timing.as<long long>();
    }
    if (ctx->mlp_variant == 2 && !density_only) {
      if (ctx->relu_scratch.ensure(tc2_scratch_bytes(ctx->sm_count)) != cudaSuccess) return fail(ctx, DSNERF_ERR_CUDA, "ReLU scratch");
      unsigned int* dbg = nullptr;
      if (ctx->tc_watchdog) {  // debug: bounded barrier waits; the record lives in mapped host memory so that it survives the trap
        if (!ctx->h_tc_dbg && cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_tc_dbg), 64, cudaHostAllocMapped) != cudaSuccess)
          return fail(ctx, DSNERF_ERR_CUDA, "watchdog buffer");
        memset(ctx->h_tc_dbg, 0, 64);
        CK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&dbg), ctx->h_tc_dbg, 0));
      }
      if (int e = tc2_launch(ctx->tw, timing, dbg, (ctx->profile >> 3) & 3, ctx->relu_scratch.as<uint32_t>(), active, d_count, host_count, ctx->mlp_a.as<float4>(),
                             ctx->mlp_g.as<float4>(), ctx->sm_count, st))
        return fail(ctx, DSNERF_ERR_CUDA, std::string("launch mlp_tc2: ") + cudaGetErrorString((cudaError_t)e));
      if (ctx->tc_watchdog) {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
          char buf[256];
          snprintf(buf, sizeof buf, "mlp_tc2 watchdog: %s; barrier id 0x%x, block %u, thread %u, parity %u", cudaGetErrorString(e),
                   ctx->h_tc_dbg[1], ctx->h_tc_dbg[2], ctx->h_tc_dbg[3], ctx->h_tc_dbg[4]);
          return fail(ctx, DSNERF_ERR_CUDA, buf);
        }
      }
    } else if (int e = tc_launch(ctx->tw, timing, (ctx->profile >> 3) & 3, (ctx->profile >> 6) & 3, active, d_count, host_count, ctx->mlp_a.as<float4>(),
                          ctx->mlp_g.as<float4>(), density_only, ctx->sm_count, st))
      return fail(ctx, DSNERF_ERR_CUDA, std::string("launch mlp_tc: ") + cudaGetErrorString((cudaError_t)e));
  }
  profile_end(ctx, st, a, b);
  return 0;
}

// make `st` wait for the posed-mesh grid that dsnerf_set_frame is building on the side stream
int join_grid(dsnerf_ctx* ctx, cudaStream_t st) {
  if (ctx->grid_on_side) CK(cudaStreamWaitEvent(st, ctx->ev_grid, 0));
  return 0;
}

int check_ready(dsnerf_ctx* ctx, bool need_frame) {
  if (!ctx) return DSNERF_ERR_INVALID;
  if (!ctx->have_weights) return fail(ctx, DSNERF_ERR_STATE, "dsnerf_set_weights has not been called");
  if (!ctx->have_mesh) return fail(ctx, DSNERF_ERR_STATE, "dsnerf_set_mesh has not been called");
  if (need_frame && !ctx->have_frame) return fail(ctx, DSNERF_ERR_STATE, "dsnerf_set_frame has not been called");
  return 0;
}

int launch_shade(dsnerf_ctx* ctx, const ShadeArgs& sa, unsigned flags, cudaStream_t st, int* launches = nullptr) {
  if (!ctx->tw.fp16_ok) flags |= DSNERF_MLP_FP32_SIMT;
  if (launches) *launches += (flags & DSNERF_MLP_FP32_SIMT) ? 4 : 6;  // mark_points, build<1>, build<2>, (canon_nearest, canon_long,) lighting
  // canonical-space lookup cells of the active points (static mesh: cells stay built across frames)
  if (int e = ensure_cells(ctx, ctx->g_canon, st, [&] {
        mark_points_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(sa.active, sa.n_active, sa.n_active_host, ctx->g_canon.g);
      }))
    return e;
  if (flags & DSNERF_MLP_FP32_SIMT) {
    shade_kernel<<<ctx->sm_count * 3, SHADE_THREADS, SHADE_SMEM, st>>>(sa, ctx->lw, ctx->g_canon.g);
    CKL("shade");
  } else {
    ShadeArgs sb = sa;
    canon_nearest_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(sa.active, sa.n_active, sa.n_active_host, ctx->g_canon.g, ctx->g_canon.cent.as<float>(),
                                                            ctx->F, ctx->active_cidx.as<int>(), ctx->canon_queue.as<int>());
    CKL("canon_nearest");
    canon_long_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(sa.active, ctx->g_canon.g, ctx->canon_queue.as<int>(), ctx->active_cidx.as<int>());
    CKL("canon_long");
    sb.active_cidx = ctx->active_cidx.as<int>();
    if (ctx->tw.rgb3)  // precise weight mode: the lighting layer runs the 3-pass split as well (light_tc.cuh)
      light_tc_kernel<true><<<ctx->sm_count, LT_THREADS, LT_SMEM3, st>>>(sb, ctx->lw, ctx->light_w2.as<uint8_t>(), ctx->g_canon.g);
    else
      light_tc_kernel<false><<<ctx->sm_count * 2, LT_THREADS, LT_SMEM, st>>>(sb, ctx->lw, ctx->light_w2.as<uint8_t>(), ctx->g_canon.g);
    CKL("light_tc");
  }
  return 0;
}

ShadeArgs base_shade_args(dsnerf_ctx* ctx) {
  ShadeArgs s{};
  s.active = ctx->active.as<float4>();
  s.mlp_a = ctx->mlp_a.as<float4>();
  s.mlp_g = ctx->mlp_g.as<float4>();
  s.posed = ctx->posed.as<float>();
  s.canon = ctx->canon.as<float>();
  s.faces = ctx->faces.as<int>();
  s.cent_canon = ctx->g_canon.cent.as<float>();
  s.normal_m = ctx->normal_m.as<float4>();
  s.F = ctx->F;
  for (int k = 0; k < 3; ++k) s.light_shift[k] = ctx->light_shift[k];
  s.has_shift = ctx->has_shift;
  for (int k = 0; k < 4; ++k) s.rot[k] = ctx->rot[k];
  s.rot_center[0] = ctx->rot_center[0]; s.rot_center[1] = ctx->rot_center[1];
  s.has_rot = ctx->has_rot;
  s.raw = ctx->raw.as<float4>();
  return s;
}

// peer destinations of the fused all-gather (dsnerf_render_gather)
struct GatherTargets {
  int n_peers = 0;
  float* peer[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float* mc = nullptr;
};

// shared body of dsnerf_render / dsnerf_render_z / dsnerf_render_train (jitter, raw_noise: training-mode draws or NULL)
int render_impl(dsnerf_ctx* ctx, const float* ray_o, const float* ray_d, const float* near, const float* far, const float* z_in,
                int64_t R, int N, unsigned flags, float* rgb, float* depth, float* acc, float* disp, float* weights, float* z_out,
                cudaStream_t st, const float* jitter = nullptr, const float* raw_noise = nullptr, const GatherTargets* gt = nullptr) {
  if (int e = check_ready(ctx, true)) return e;
  if (gt && (flags & DSNERF_EARLY_STOP)) return fail(ctx, DSNERF_ERR_INVALID, "dsnerf_render_gather does not combine with DSNERF_EARLY_STOP");
  if (R < 0 || N < 1 || N > 4096) return fail(ctx, DSNERF_ERR_INVALID, "n_rays must be >= 0 and 1 <= n_samples <= 4096");
  if (R * (int64_t)N > 0x7fffffffLL) return fail(ctx, DSNERF_ERR_INVALID, "n_rays * n_samples must fit in 31 bits (sample ids are 32-bit); split the batch");
  if (!ray_o || !ray_d || (!z_in && (!near || !far)) || !rgb || !depth || !acc || !disp) {
    if (R > 0) return fail(ctx, DSNERF_ERR_INVALID, "null input/output pointer");
  }
  if (jitter && !z_out && R > 0) return fail(ctx, DSNERF_ERR_INVALID, "jittered sampling needs the z_vals output buffer");
  ctx->stats = dsnerf_stats_t{};
  if (R == 0) return 0;
  CK(cudaSetDevice(ctx->device));
  if (int e = ensure_workspace(ctx, R, N)) return e;
  if (int e = ensure_tvals(ctx, N, st)) return e;
  int launches = 0;
  int64_t P = R * (int64_t)N;
  unsigned long long* cnt = ctx->counters.as<unsigned long long>();
  CK(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * 4, st));
  const float* near_use = near;
  const float* far_use = far;
  if (!z_in && (flags & DSNERF_SAMPLE_GG)) {
    CK(ctx->vq.ensure(sizeof(float4) * ctx->V + 32));
    unsigned* qbox = reinterpret_cast<unsigned*>(ctx->vq.as<float4>() + ctx->V);
    CK(cudaMemsetAsync(qbox, 0xff, 12, st));
    CK(cudaMemsetAsync(qbox + 3, 0, 12, st));
    gg_prep_kernel<<<(ctx->V + 255) / 256, 256, 0, st>>>(ctx->posed.as<float>(), ctx->V, ray_o, ctx->vq.as<float4>(), qbox);
    CKL("gg_prep");
    float gamma2 = (float)(0.05 * 0.05);  // python double 0.05**2 rounded to fp32 (pts_utils.py:36)
    const float gamma_pad = 0.05f * 1.002f + 1e-4f;
    // direction tiles: [GgFrame (64 B)][bad flag][counts][lists]
    const size_t gg_bytes = 256 + sizeof(int) * (size_t)GG_TILES * GG_TILES * (1 + GG_CAP);
    CK(ctx->gg_bins.ensure(gg_bytes));
    GgFrame* frame = ctx->gg_bins.as<GgFrame>();
    int* bad = reinterpret_cast<int*>(ctx->gg_bins.as<char>() + 128);
    int* counts = reinterpret_cast<int*>(ctx->gg_bins.as<char>() + 256);
    int* lists = counts + GG_TILES * GG_TILES;
    CK(cudaMemsetAsync(bad, 0, 128 + sizeof(int) * GG_TILES * GG_TILES, st));
    gg_frame_kernel<<<1, 32, 0, st>>>(qbox, gamma_pad, frame);
    CKL("gg_frame");
    gg_bin_kernel<<<(ctx->V + 255) / 256, 256, 0, st>>>(ctx->vq.as<float4>(), ctx->V, gamma_pad, frame, counts, lists, bad);
    CKL("gg_bin");
    gg_bounds_kernel<<<(unsigned)((R * 32 + GG_THREADS - 1) / GG_THREADS), GG_THREADS, 0, st>>>(
        ctx->vq.as<float4>(), ctx->V, qbox, frame, counts, lists, bad, ray_d, near, far, R, gamma2, 0.05f, ctx->near2.as<float>(),
        ctx->far2.as<float>());
    CKL("gg_bounds");
    launches += 4;
    near_use = ctx->near2.as<float>();
    far_use = ctx->far2.as<float>();
  }
  if (int e = join_grid(ctx, st)) return e;
  if (jitter) {  // training mode: stratified jitter (pts_utils.py:6-13) written straight into the caller's z_vals
    jitter_z_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(near_use, far_use, ctx->tvals.as<float>(), jitter, R, N, z_out);
    CKL("jitter_z");
    ++launches;
    z_in = z_out;
    z_out = nullptr;
  }
  WarpArgs wa{};
  wa.ray_o = ray_o; wa.ray_d = ray_d; wa.near = near_use; wa.far = far_use; wa.z_in = z_in; wa.tvals = ctx->tvals.as<float>();
  wa.posed = ctx->posed.as<float>(); wa.canon = ctx->canon.as<float>(); wa.faces = ctx->faces.as<int>();
  wa.R = R; wa.N = N;
  wa.active = ctx->active.as<float4>(); wa.active_tri = ctx->active_tri.as<int>(); wa.sample_mask = ctx->ray_mask.as<unsigned>();
  wa.counters = cnt;
  wa.count_candidates = (ctx->profile & 2) ? 1 : 0;
  if (!raw_noise)
  if (int e = ensure_cells(ctx, ctx->g_posed, st, [&] {
        const int64_t mark_threads = R * ((N + MARK_SPT - 1) / MARK_SPT);
        mark_samples_kernel<<<(unsigned)((mark_threads + WARP_THREADS - 1) / WARP_THREADS), WARP_THREADS, 0, st>>>(wa, ctx->g_posed.g);
      }))
    return e;
  ShadeArgs sa = base_shade_args(ctx);
  sa.ray_o = ray_o; sa.ray_d = ray_d; sa.near = near_use; sa.far = far_use; sa.z_in = z_in; sa.tvals = ctx->tvals.as<float>();
  sa.N = N;
  CompositeArgs ca{};
  ca.sample_mask = ctx->ray_mask.as<unsigned>();
  ca.raw = ctx->raw.as<float4>(); ca.ray_d = ray_d; ca.near = near_use; ca.far = far_use; ca.tvals = ctx->tvals.as<float>(); ca.z_in = z_in;
  ca.R = R; ca.N = N; ca.rgb = rgb; ca.depth = depth; ca.acc = acc; ca.disp = disp; ca.weights = weights; ca.z_out = z_out;
  if (gt) {
    ca.n_peers = gt->n_peers;
    for (int i = 0; i < gt->n_peers; ++i) ca.peer[i] = gt->peer[i];
    ca.mc = gt->mc;
  }
  ctx->ert_mode = 0;
  if (raw_noise) {
    // ---- training mode with density noise: the network runs on every sample (see sample_warp_all_kernel)
    sample_warp_all_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(wa, ctx->g_posed.g, ctx->F, ctx->g_posed.cent.as<float>());
    CKL("sample_warp_all");
    if (int e = launch_mlp(ctx, cnt, 0, flags, 0, st)) return e;
    sa.n_active = cnt;
    sa.active_tri = ctx->active_tri.as<int>();
    if (int e = launch_shade(ctx, sa, flags, st, &launches)) return e;
    ca.noise = raw_noise;
    ca.all_raw = 1;
    composite_kernel<<<(unsigned)((R + COMP_RAYS - 1) / COMP_RAYS), COMP_RAYS * 32, 0, st>>>(ca);
    CKL("composite");
    launches += 3;  // sample_warp_all, MLP, composite
  } else if ((flags & DSNERF_EARLY_STOP) && N >= 8) {
    // ---- early ray termination: front-to-back waves of samples; see shade.cuh.  Measured on the 512x512x64 benchmark frame
    // (DSNERF_ERT_WAVES = 2 / 3 / 4): 10 / 18 / 22 % of the samples skipped, frame 8.60 / 8.31 / 8.34 ms against 8.87 ms
    // exhaustive -- every wave costs 8 more launches, so three waves are the default
    int kWaves = 3;
    if (const char* ev = getenv("DSNERF_ERT_WAVES")) kWaves = std::max(2, std::min(4, atoi(ev)));
    const float tau = 1e-6f;
    const int wsize = (N + kWaves - 1) / kWaves;
    const int64_t region = R * (int64_t)wsize;  // a wave holds at most wsize samples per ray
    CK(ctx->ert_active.ensure(sizeof(float4) * (size_t)(region + 128)));
    CK(ctx->ert_tri.ensure(sizeof(int) * (size_t)(region + 128)));
    CK(ctx->ert_state.ensure(sizeof(RayState) * (size_t)R));
    CK(ctx->ert_cnt.ensure(sizeof(unsigned long long) * 8));
    unsigned long long* ec = ctx->ert_cnt.as<unsigned long long>();  // [0..3] wave sizes, [4..7] evaluated per wave
    CK(cudaMemsetAsync(ec, 0, sizeof(unsigned long long) * 8, st));
    wa.waves = kWaves; wa.wave_size = wsize; wa.region = region; wa.wave_counters = ec;
    sample_warp_kernel<<<(unsigned)((P + WARP_SPB - 1) / WARP_SPB), WARP_THREADS, 0, st>>>(wa, ctx->g_posed.g);
    CKL("sample_warp");
    launches += 4;
    for (int k = 0; k < kWaves; ++k) {
      const float4* list = ctx->active.as<float4>() + k * region;
      const int* tri = ctx->active_tri.as<int>() + k * region;
      const unsigned long long* count = ec + k;
      if (k > 0) {
        filter_wave_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(list, tri, ec + k, ctx->ert_state.as<RayState>(), N, tau,
                                                              ctx->ert_active.as<float4>(), ctx->ert_tri.as<int>(), ec + 4 + k);
        CKL("filter_wave");
        ++launches;
        list = ctx->ert_active.as<float4>();
        tri = ctx->ert_tri.as<int>();
        count = ec + 4 + k;
      }
      if (int e = launch_mlp(ctx, count, 0, flags, 0, st, list)) return e;
      sa.active = list;
      sa.active_tri = tri;
      sa.n_active = count;
      if (int e = launch_shade(ctx, sa, flags, st, &launches)) return e;
      const int i0 = k * wsize, i1 = std::min(N, (k + 1) * wsize);
      composite_wave_kernel<<<(unsigned)((R * 32 + 255) / 256), 256, 0, st>>>(ca, ctx->ert_state.as<RayState>(), i0, i1, tau, k == kWaves - 1);
      CKL("composite_wave");
      launches += 2;  // MLP, composite_wave
    }
    CK(cudaMemcpyAsync(ctx->h_ert, ec, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, st));
    ctx->ert_mode = 1;
  } else {
    sample_warp_kernel<<<(unsigned)((P + WARP_SPB - 1) / WARP_SPB), WARP_THREADS, 0, st>>>(wa, ctx->g_posed.g);
    CKL("sample_warp");
    launches += 4;
    if (int e = launch_mlp(ctx, cnt, 0, flags, 0, st)) return e;
    ++launches;
    sa.n_active = cnt;
    sa.active_tri = ctx->active_tri.as<int>();
    if (int e = launch_shade(ctx, sa, flags, st, &launches)) return e;
    composite_kernel<<<(unsigned)((R + COMP_RAYS - 1) / COMP_RAYS), COMP_RAYS * 32, 0, st>>>(ca);
    CKL("composite");
    ++launches;
  }
  CK(cudaMemcpyAsync(ctx->h_counters, cnt, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaEventRecord(ctx->stats_ready, st));
  ctx->stats.rays = R;
  ctx->stats.samples = P;
  ctx->stats.kernel_launches = launches + ctx->g_posed.launches;  // incl. the per-frame grid build of dsnerf_set_frame
  return 0;
}

}  // namespace

extern "C" {

int dsnerf_abi_version(void) { return DSNERF_ABI_VERSION; }

int dsnerf_create(dsnerf_ctx** out, int device) {
  if (!out) return DSNERF_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return DSNERF_ERR_NO_DEVICE;
  if (device < 0 || device >= n) return DSNERF_ERR_INVALID;
  dsnerf_ctx* ctx = new dsnerf_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return DSNERF_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return DSNERF_ERR_CUDA; }
  if (prop.major != 10) { delete ctx; return DSNERF_ERR_NO_DEVICE; }  // built for sm_100a only
  ctx->sm_count = prop.multiProcessorCount;
  cudaEventCreateWithFlags(&ctx->pin_free, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->stats_ready, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming);
  if (!getenv("DSNERF_NO_COPY_STREAM")) cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking);
  if (!getenv("DSNERF_NO_COPY_STREAM")) cudaStreamCreateWithFlags(&ctx->down, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&ctx->render_done, cudaEventDisableTiming);
  for (int i = 0; i < 2; ++i) cudaEventCreateWithFlags(&ctx->host_done[i], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_grid, cudaEventDisableTiming);
  if (!getenv("DSNERF_NO_SIDE_STREAM")) cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking);
  cudaMallocHost(reinterpret_cast<void**>(&ctx->h_counters), sizeof(unsigned long long) * 4);
  memset(ctx->h_counters, 0, sizeof(unsigned long long) * 4);
  cudaMallocHost(reinterpret_cast<void**>(&ctx->h_ert), sizeof(unsigned long long) * 8);
  memset(ctx->h_ert, 0, sizeof(unsigned long long) * 8);
  cudaFuncSetAttribute(mlp_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SIMT_SMEM);
  cudaFuncSetAttribute(shade_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SHADE_SMEM);
  cudaFuncSetAttribute(light_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LT_SMEM);
  cudaFuncSetAttribute(light_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LT_SMEM3);
  tc_configure();
  tc2_configure();
  if (const char* ev = getenv("DSNERF_MLP_VARIANT")) ctx->mlp_variant = atoi(ev) == 1 ? 1 : 2;
  ctx->tc_watchdog = getenv("DSNERF_TC_WATCHDOG") != nullptr;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { delete ctx; return DSNERF_ERR_CUDA; }
  *out = ctx;
  return 0;
}

void dsnerf_destroy(dsnerf_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  DevBuf* bufs[] = {&ctx->light_w2, &ctx->wblob, &ctx->bias0, &ctx->faces, &ctx->canon, &ctx->posed, &ctx->vq, &ctx->gg_bins, &ctx->normal_m, &ctx->near2, &ctx->far2, &ctx->raw,
                    &ctx->active, &ctx->active_tri, &ctx->active_cidx, &ctx->ray_mask, &ctx->mlp_a, &ctx->mlp_g, &ctx->tvals, &ctx->counters, &ctx->io,
                    &ctx->ert_active, &ctx->ert_tri, &ctx->ert_state, &ctx->ert_cnt};
  for (DevBuf* b : bufs) b->release();
  ctx->g_canon.release();
  ctx->g_posed.release();
  ctx->tw.release();
  if (ctx->pin) cudaFreeHost(ctx->pin);
  if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
  if (ctx->h_ert) cudaFreeHost(ctx->h_ert);
  if (ctx->pin_free) cudaEventDestroy(ctx->pin_free);
  if (ctx->stats_ready) cudaEventDestroy(ctx->stats_ready);
  if (ctx->copy_done) cudaEventDestroy(ctx->copy_done);
  if (ctx->copy) cudaStreamDestroy(ctx->copy);
  if (ctx->down) cudaStreamDestroy(ctx->down);
  if (ctx->render_done) cudaEventDestroy(ctx->render_done);
  for (int i = 0; i < 2; ++i) { if (ctx->host_done[i]) cudaEventDestroy(ctx->host_done[i]); ctx->io2[i].release(); }
  if (ctx->side) cudaStreamDestroy(ctx->side);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_grid) cudaEventDestroy(ctx->ev_grid);
  for (auto& p : ctx->pending) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
  delete ctx;
}

const char* dsnerf_last_error(const dsnerf_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int dsnerf_set_weights(dsnerf_ctx* ctx, const float* const* t, int n_tensors) {
  if (!ctx) return DSNERF_ERR_INVALID;
  if (!t || n_tensors != DSNERF_NUM_WEIGHT_TENSORS) return fail(ctx, DSNERF_ERR_INVALID, "expected 33 tensors in state_dict order");
  // validate everything before touching the context: a rejected call leaves the previously staged weights in use
  for (int i = 0; i < n_tensors; ++i) {
    if (!t[i]) return fail(ctx, DSNERF_ERR_INVALID, "null weight tensor");
    for (int j = 0; j < kTensorSize[i]; ++j)
      if (!std::isfinite(t[i][j])) return fail(ctx, DSNERF_ERR_INVALID, "non-finite weight");
  }
  for (int i = 0; i < n_tensors; ++i) ctx->hw[i].assign(t[i], t[i] + kTensorSize[i]);
  ctx->have_weights = false;  // until the staging below has succeeded
  ctx->have_frame = false;
  CK(cudaSetDevice(ctx->device));
  CK(cudaDeviceSynchronize());  // weights may be in use by work in flight
  BlobBuilder bb;
  size_t off_wt[7], off_w[7], off_b[7];
  const int sw_idx[7] = {T_S1_0_W, T_S1_2_W, T_S1_4_W, T_S1_6_W, T_S2_0_W, T_S2_2_W, T_S2_4_W};
  const int fwdK[7] = {64, 256, 256, 256, 320, 256, 256};
  for (int l = 0; l < 7; ++l) {
    const std::vector<float>& w = ctx->hw[sw_idx[l]];
    int in_dim = (l == 0) ? 87 : (l == 4 ? 319 : 256);
    int col0 = (l == 0) ? 8 : 0;                     // layer 0: PE columns 8..70 only (code/pose folded into the bias)
    int ncol = (l == 0) ? 63 : in_dim;
    off_wt[l] = bb.add((size_t)fwdK[l] * 256);
    off_w[l] = bb.add((size_t)256 * fwdK[l]);
    for (int o = 0; o < 256; ++o)
      for (int k = 0; k < ncol; ++k) {
        float v = w[(size_t)o * in_dim + col0 + k];
        bb.data[off_wt[l] + (size_t)k * 256 + o] = v;
        bb.data[off_w[l] + (size_t)o * fwdK[l] + k] = v;
      }
    off_b[l] = bb.add(256);
    for (int o = 0; o < 256; ++o) bb.data[off_b[l] + o] = ctx->hw[sw_idx[l] + 1][o];
  }
  size_t off_wd = bb.add(256);
  for (int o = 0; o < 256; ++o) bb.data[off_wd + o] = ctx->hw[T_DENS_W][o];
  size_t off_r1 = bb.add(256 * 128), off_r1b = bb.add(128), off_r2 = bb.add(3 * 128), off_r2b = bb.add(4);
  for (int o = 0; o < 128; ++o)
    for (int k = 0; k < 256; ++k) bb.data[off_r1 + (size_t)k * 128 + o] = ctx->hw[T_RGB1_W][(size_t)o * 256 + k];
  for (int o = 0; o < 128; ++o) bb.data[off_r1b + o] = ctx->hw[T_RGB1_B][o];
  for (int i = 0; i < 3 * 128; ++i) bb.data[off_r2 + i] = ctx->hw[T_RGB3_W][i];
  for (int i = 0; i < 3; ++i) bb.data[off_r2b + i] = ctx->hw[T_RGB3_B][i];
  size_t off_l1 = bb.add(9 * 128), off_l1b = bb.add(128), off_l2 = bb.add(128 * 128), off_l2b = bb.add(128), off_l3 = bb.add(128);
  for (int o = 0; o < 128; ++o)
    for (int k = 0; k < 9; ++k) bb.data[off_l1 + (size_t)k * 128 + o] = ctx->hw[T_L0_W][(size_t)o * 9 + k];
  for (int o = 0; o < 128; ++o)
    for (int k = 0; k < 128; ++k) bb.data[off_l2 + (size_t)k * 128 + o] = ctx->hw[T_L2_W][(size_t)o * 128 + k];
  for (int o = 0; o < 128; ++o) {
    bb.data[off_l1b + o] = ctx->hw[T_L0_B][o];
    bb.data[off_l2b + o] = ctx->hw[T_L2_B][o];
    bb.data[off_l3 + o] = ctx->hw[T_L4_W][o];
  }
  CK(ctx->wblob.ensure(bb.data.size() * sizeof(float)));
  CK(cudaMemcpy(ctx->wblob.p, bb.data.data(), bb.data.size() * sizeof(float), cudaMemcpyHostToDevice));
  CK(ctx->bias0.ensure(256 * sizeof(float)));
  const float* d = ctx->wblob.as<float>();
  for (int l = 0; l < 7; ++l) { ctx->sw.wt[l] = d + off_wt[l]; ctx->sw.w[l] = d + off_w[l]; ctx->sw.bias[l] = d + off_b[l]; }
  ctx->sw.bias[0] = ctx->bias0.as<float>();
  ctx->sw.w_dens = d + off_wd;
  ctx->sw.b_dens = ctx->hw[T_DENS_B][0];
  ctx->sw.wt_rgb1 = d + off_r1; ctx->sw.b_rgb1 = d + off_r1b; ctx->sw.w_rgb2 = d + off_r2; ctx->sw.b_rgb2 = d + off_r2b;
  ctx->lw.w1t = d + off_l1; ctx->lw.b1 = d + off_l1b; ctx->lw.w2t = d + off_l2; ctx->lw.b2 = d + off_l2b; ctx->lw.w3 = d + off_l3;
  ctx->lw.b3 = ctx->hw[T_L4_B][0];
  // rgb head precision: 1 pass where the probe says it is enough, else 3 (DSNERF_RGB_PASSES=1|3 overrides, for A/B runs)
  ctx->rgb_probe_err = rgb_single_pass_probe(ctx->hw);
  bool rgb3 = ctx->rgb_probe_err > 1.2e-5f;
  if (const char* ev = getenv("DSNERF_RGB_PASSES")) rgb3 = atoi(ev) >= 3;
  if (int e = ctx->tw.stage(ctx->hw[T_S1_0_W], ctx->hw[T_S1_2_W], ctx->hw[T_S1_4_W], ctx->hw[T_S1_6_W], ctx->hw[T_S2_0_W], ctx->hw[T_S2_2_W],
                            ctx->hw[T_S2_4_W], ctx->hw[T_S1_2_B], ctx->hw[T_S1_4_B], ctx->hw[T_S1_6_B], ctx->hw[T_S2_0_B], ctx->hw[T_S2_2_B],
                            ctx->hw[T_S2_4_B], ctx->hw[T_DENS_W], ctx->hw[T_DENS_B][0], ctx->hw[T_RGB1_W], ctx->hw[T_RGB1_B],
                            ctx->hw[T_RGB3_W], ctx->hw[T_RGB3_B], rgb3))
    return fail(ctx, DSNERF_ERR_CUDA, std::string("staging tensor-core weights: ") + cudaGetErrorString((cudaError_t)e));
  {
    std::vector<__half> w2p;
    light_pack_w2(ctx->hw[T_L2_W], ctx->hw[T_L0_W], ctx->hw[T_L0_B], w2p);
    CK(ctx->light_w2.ensure(w2p.size() * sizeof(__half)));
    CK(cudaMemcpy(ctx->light_w2.p, w2p.data(), w2p.size() * sizeof(__half), cudaMemcpyHostToDevice));
  }
  ctx->have_weights = true;
  ctx->have_frame = false;  // the folded bias depends on the weights
  return 0;
}

int dsnerf_set_mesh(dsnerf_ctx* ctx, const int32_t* faces, int n_faces, const float* canonical_verts, int n_verts) {
  if (!ctx) return DSNERF_ERR_INVALID;
  if (!faces || !canonical_verts || n_faces < 1 || n_verts < 3) return fail(ctx, DSNERF_ERR_INVALID, "bad mesh");
  for (int i = 0; i < 3 * n_faces; ++i)
    if (faces[i] < 0 || faces[i] >= n_verts) return fail(ctx, DSNERF_ERR_INVALID, "face index out of range");
  CK(cudaSetDevice(ctx->device));
  CK(cudaDeviceSynchronize());
  ctx->F = n_faces;
  ctx->V = n_verts;
  ctx->h_faces.assign(faces, faces + 3 * (size_t)n_faces);
  ctx->h_canon.assign(canonical_verts, canonical_verts + 3 * (size_t)n_verts);
  CK(ctx->faces.ensure(sizeof(int) * 3 * n_faces));
  CK(ctx->canon.ensure(sizeof(float) * 3 * n_verts));
  CK(ctx->posed.ensure(sizeof(float) * 3 * n_verts));
  CK(cudaMemcpy(ctx->faces.p, faces, sizeof(int) * 3 * n_faces, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ctx->canon.p, canonical_verts, sizeof(float) * 3 * n_verts, cudaMemcpyHostToDevice));
  if (int e = build_grid(ctx, ctx->g_canon, ctx->canon.as<float>(), ctx->h_canon.data(), 0, 0)) return e;
  CK(cudaDeviceSynchronize());
  ctx->have_mesh = true;
  ctx->have_frame = false;
  return 0;
}

int dsnerf_set_frame(dsnerf_ctx* ctx, const float* posed_verts, const float* poses, int frame, int zero_code, const float* light_shift,
                     const float* rot, const float* rot_center, void* stream) {
  if (int e = check_ready(ctx, false)) return e;
  if (!posed_verts || !poses) return fail(ctx, DSNERF_ERR_INVALID, "null posed_verts/poses");
  if (frame < 0 || frame >= 500) return fail(ctx, DSNERF_ERR_INVALID, "frame index outside the 500-row code table");
  if ((rot == nullptr) != (rot_center == nullptr)) return fail(ctx, DSNERF_ERR_INVALID, "rot and rot_center must be given together");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  if (int e = join_grid(ctx, st)) return e;  // a previous frame's grid build may still read the vertex buffer overwritten below
  // pose feature: batch_rod2quat -> pose_mlp (model/spacenet.py:223-236); identical for every sample of the frame
  float q[92], h1[64], h2[64], pf[16];
  rod2quat_host(poses, q);
  linear_host(ctx->hw[T_P0_W], ctx->hw[T_P0_B], 64, 92, q, h1, true);
  linear_host(ctx->hw[T_P2_W], ctx->hw[T_P2_B], 64, 64, h1, h2, true);
  linear_host(ctx->hw[T_P4_W], ctx->hw[T_P4_B], 16, 64, h2, pf, false);
  size_t vbytes = sizeof(float) * 3 * ctx->V;
  if (int e = pin_acquire(ctx, vbytes + 256 * sizeof(float))) return e;
  float* pv = reinterpret_cast<float*>(ctx->pin);
  float* pb = pv + 3 * (size_t)ctx->V;
  memcpy(pv, posed_verts, vbytes);
  // fold code (cols 0..7) and pose feature (cols 71..86) of stage1.0 into its bias (spacenet.py:125-130)
  const std::vector<float>& w0 = ctx->hw[T_S1_0_W];
  const float* code = ctx->hw[T_EMB].data() + 8 * (size_t)frame;
  for (int o = 0; o < 256; ++o) {
    double acc = ctx->hw[T_S1_0_B][o];
    if (!zero_code)
      for (int j = 0; j < 8; ++j) acc += (double)w0[(size_t)o * 87 + j] * code[j];
    for (int j = 0; j < 16; ++j) acc += (double)w0[(size_t)o * 87 + 71 + j] * pf[j];
    pb[o] = (float)acc;
  }
  CK(cudaMemcpyAsync(ctx->posed.p, pv, vbytes, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->bias0.p, pb, 256 * sizeof(float), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(ctx->tw.bias0_slot(), pb, 256 * sizeof(float), cudaMemcpyHostToDevice, st));
  // The posed-mesh grid (16 small kernels, ~0.25 ms, latency bound) only depends on the vertex upload: it is built on the
  // context's side stream, forked here, so that the render call that follows can already run its GG ray bounds (which need
  // the vertices, not the grid) on the caller's stream; every consumer of the grid joins through join_grid().
  cudaStream_t gs = ctx->side ? ctx->side : st;
  if (gs != st) {
    CK(cudaEventRecord(ctx->ev_fork, st));   // after the uploads, and after everything the previous frame left on `st`
    CK(cudaStreamWaitEvent(gs, ctx->ev_fork, 0));
  }
  int e = build_grid(ctx, ctx->g_posed, ctx->posed.as<float>(), pv, 1, gs);
  if (!e) {
    if (ctx->normal_m.ensure(sizeof(float4) * 3 * (size_t)ctx->F) != cudaSuccess) e = fail(ctx, DSNERF_ERR_CUDA, "normal matrices");
    else normal_matrix_kernel<<<(ctx->F + 255) / 256, 256, 0, gs>>>(ctx->canon.as<float>(), ctx->posed.as<float>(), ctx->faces.as<int>(), ctx->F,
                                                                  ctx->normal_m.as<float4>());
  }
  ctx->grid_on_side = gs != st;
  if (gs != st) {
    CK(cudaEventRecord(ctx->ev_grid, gs));
    if (e) CK(cudaStreamWaitEvent(st, ctx->ev_grid, 0));
  }
  if (int e2 = pin_release(ctx, st)) return e2;
  if (e) return e;
  ctx->has_shift = light_shift != nullptr;
  for (int k = 0; k < 3; ++k) ctx->light_shift[k] = light_shift ? light_shift[k] : 0.f;
  ctx->has_rot = rot != nullptr;
  if (rot) { for (int k = 0; k < 4; ++k) ctx->rot[k] = rot[k]; ctx->rot_center[0] = rot_center[0]; ctx->rot_center[1] = rot_center[1]; }
  ctx->have_frame = true;
  return 0;
}

int dsnerf_render(dsnerf_ctx* ctx, const float* ray_o, const float* ray_d, const float* near, const float* far, int64_t n_rays,
                  int n_samples, unsigned flags, float* rgb, float* depth, float* acc, float* disp, float* weights, float* z_vals,
                  void* stream) {
  if (!ctx) return DSNERF_ERR_INVALID;
  return render_impl(ctx, ray_o, ray_d, near, far, nullptr, n_rays, n_samples, flags, rgb, depth, acc, disp, weights, z_vals,
                     reinterpret_cast<cudaStream_t>(stream));
}

int dsnerf_render_train(dsnerf_ctx* ctx, const float* ray_o, const float* ray_d, const float* near, const float* far, int64_t n_rays,
                        int n_samples, unsigned flags, const float* jitter, const float* raw_noise, float* rgb, float* depth, float* acc,
                        float* disp, float* weights, float* z_vals, void* stream) {
  if (!ctx) return DSNERF_ERR_INVALID;
  return render_impl(ctx, ray_o, ray_d, near, far, nullptr, n_rays, n_samples, flags, rgb, depth, acc, disp, weights, z_vals,
                     reinterpret_cast<cudaStream_t>(stream), jitter, raw_noise);
}

int dsnerf_render_z(dsnerf_ctx* ctx, const float* ray_o, const float* ray_d, const float* z_vals, int64_t n_rays, int n_samples,
                    unsigned flags, float* rgb, float* depth, float* acc, float* disp, float* weights, void* stream) {
  if (!ctx) return DSNERF_ERR_INVALID;
  if (!z_vals && n_rays > 0) return fail(ctx, DSNERF_ERR_INVALID, "null z_vals");
  return render_impl(ctx, ray_o, ray_d, nullptr, nullptr, z_vals, n_rays, n_samples, flags, rgb, depth, acc, disp, weights, nullptr,
                     reinterpret_cast<cudaStream_t>(stream));
}

int dsnerf_render_gather(dsnerf_ctx* ctx, const float* ray_o, const float* ray_d, const float* near, const float* far, int64_t n_rays,
                         int n_samples, unsigned flags, float* own_block, float* const* peer_blocks, int n_peers, float* multicast_block,
                         void* stream) {
  if (!ctx) return DSNERF_ERR_INVALID;
  if (n_peers < 0 || n_peers > 7) return fail(ctx, DSNERF_ERR_INVALID, "at most 7 peer blocks (8 GPUs of one NVSwitch domain)");
  if (n_rays > 0 && (!own_block || (n_peers > 0 && !peer_blocks && !multicast_block)))
    return fail(ctx, DSNERF_ERR_INVALID, "null output block");
  GatherTargets gt;
  if (multicast_block) {
    gt.mc = multicast_block;
  } else {
    gt.n_peers = n_peers;
    for (int i = 0; i < n_peers; ++i) {
      if (!peer_blocks[i]) return fail(ctx, DSNERF_ERR_INVALID, "null peer block");
      gt.peer[i] = peer_blocks[i];
    }
  }
  float* b = own_block;
  return render_impl(ctx, ray_o, ray_d, near, far, nullptr, n_rays, n_samples, flags, b, b + 3 * n_rays, b + 4 * n_rays, b + 5 * n_rays, nullptr,
                     nullptr, reinterpret_cast<cudaStream_t>(stream), nullptr, nullptr, &gt);
}

int dsnerf_render_host_async(dsnerf_ctx* ctx, const float* ray_o, const float* ray_d, const float* near, const float* far, int64_t R,
                             int N, unsigned flags, float* rgb, float* depth, float* acc, float* disp, float* weights, float* z_vals,
                             void* stream, int* ticket) {
  if (int e = check_ready(ctx, true)) return e;
  if (R < 0 || N < 1) return fail(ctx, DSNERF_ERR_INVALID, "bad sizes");
  if (!ticket) return fail(ctx, DSNERF_ERR_INVALID, "null ticket");
  const int slot = (int)(ctx->host_seq & 1);
  *ticket = (int)(ctx->host_seq & 0x7fffffff);
  ++ctx->host_seq;
  if (R == 0) { ctx->host_pending[slot] = false; return 0; }
  if (!ray_o || !ray_d || !near || !far || !rgb || !depth || !acc || !disp) return fail(ctx, DSNERF_ERR_INVALID, "null input/output pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  // two staging slots on the device: the frame submitted two calls ago must have left this one
  if (ctx->host_pending[slot]) { CK(cudaEventSynchronize(ctx->host_done[slot])); ctx->host_pending[slot] = false; }
  size_t in_f = (size_t)R * 8, out_f = (size_t)R * 6, opt_f = (size_t)R * N;
  size_t total = in_f + out_f + (weights ? opt_f : 0) + (z_vals ? opt_f : 0);
  CK(ctx->io2[slot].ensure(total * sizeof(float)));
  float* d = ctx->io2[slot].as<float>();
  float *d_o = d, *d_d = d + 3 * R, *d_n = d + 6 * R, *d_f = d + 7 * R;
  float *d_rgb = d + 8 * R, *d_dep = d_rgb + 3 * R, *d_acc = d_dep + R, *d_dsp = d_acc + R;
  float* d_w = weights ? d_dsp + R : nullptr;
  float* d_z = z_vals ? (d_dsp + R + (weights ? opt_f : 0)) : nullptr;
  // The upload does not depend on anything queued on `st` (typically the per-frame work of dsnerf_set_frame): it runs on the
  // context's copy stream and `st` joins it.
  cudaStream_t cs = ctx->copy ? ctx->copy : st;
  CK(cudaMemcpyAsync(d_o, ray_o, sizeof(float) * 3 * R, cudaMemcpyHostToDevice, cs));
  CK(cudaMemcpyAsync(d_d, ray_d, sizeof(float) * 3 * R, cudaMemcpyHostToDevice, cs));
  CK(cudaMemcpyAsync(d_n, near, sizeof(float) * R, cudaMemcpyHostToDevice, cs));
  CK(cudaMemcpyAsync(d_f, far, sizeof(float) * R, cudaMemcpyHostToDevice, cs));
  if (cs != st) {
    CK(cudaEventRecord(ctx->copy_done, cs));
    CK(cudaStreamWaitEvent(st, ctx->copy_done, 0));
  }
  if (int e = render_impl(ctx, d_o, d_d, d_n, d_f, nullptr, R, N, flags, d_rgb, d_dep, d_acc, d_dsp, d_w, d_z, st)) return e;
  // The read-back runs on the download stream behind the render, so `st` is free for the next frame's kernels at once.
  cudaStream_t ds = ctx->down ? ctx->down : st;
  if (ds != st) {
    CK(cudaEventRecord(ctx->render_done, st));
    CK(cudaStreamWaitEvent(ds, ctx->render_done, 0));
  }
  CK(cudaMemcpyAsync(rgb, d_rgb, sizeof(float) * 3 * R, cudaMemcpyDeviceToHost, ds));
  CK(cudaMemcpyAsync(depth, d_dep, sizeof(float) * R, cudaMemcpyDeviceToHost, ds));
  CK(cudaMemcpyAsync(acc, d_acc, sizeof(float) * R, cudaMemcpyDeviceToHost, ds));
  CK(cudaMemcpyAsync(disp, d_dsp, sizeof(float) * R, cudaMemcpyDeviceToHost, ds));
  if (weights) CK(cudaMemcpyAsync(weights, d_w, sizeof(float) * opt_f, cudaMemcpyDeviceToHost, ds));
  if (z_vals) CK(cudaMemcpyAsync(z_vals, d_z, sizeof(float) * opt_f, cudaMemcpyDeviceToHost, ds));
  CK(cudaEventRecord(ctx->host_done[slot], ds));
  ctx->host_pending[slot] = true;
  return 0;
}

int dsnerf_wait(dsnerf_ctx* ctx, int ticket) {
  if (!ctx) return DSNERF_ERR_INVALID;
  const int slot = ticket & 1;
  if (ctx->host_pending[slot]) {
    CK(cudaEventSynchronize(ctx->host_done[slot]));
    ctx->host_pending[slot] = false;
  }
  return 0;
}

int dsnerf_render_host(dsnerf_ctx* ctx, const float* ray_o, const float* ray_d, const float* near, const float* far, int64_t R,
                       int N, unsigned flags, float* rgb, float* depth, float* acc, float* disp, float* weights, float* z_vals,
                       void* stream) {
  int ticket = 0;
  if (int e = dsnerf_render_host_async(ctx, ray_o, ray_d, near, far, R, N, flags, rgb, depth, acc, disp, weights, z_vals, stream, &ticket)) return e;
  return dsnerf_wait(ctx, ticket);
}

int dsnerf_composite(dsnerf_ctx* ctx, const float* raw, const float* z_vals, const float* ray_d, int64_t R, int N, float* rgb,
                     float* depth, float* acc, float* disp, float* weights, void* stream) {
  return dsnerf_composite_noise(ctx, raw, z_vals, ray_d, nullptr, R, N, rgb, depth, acc, disp, weights, stream);
}

int dsnerf_composite_noise(dsnerf_ctx* ctx, const float* raw, const float* z_vals, const float* ray_d, const float* raw_noise, int64_t R,
                           int N, float* rgb, float* depth, float* acc, float* disp, float* weights, void* stream) {
  if (!ctx) return DSNERF_ERR_INVALID;
  if (R < 0 || N < 1) return fail(ctx, DSNERF_ERR_INVALID, "bad sizes");
  if (R == 0) return 0;
  if (!raw || !z_vals || !ray_d || !rgb || !depth || !acc || !disp) return fail(ctx, DSNERF_ERR_INVALID, "null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  CompositeArgs ca{};
  ca.raw = reinterpret_cast<const float4*>(raw); ca.ray_d = ray_d; ca.z_in = z_vals; ca.R = R; ca.N = N;
  ca.rgb = rgb; ca.depth = depth; ca.acc = acc; ca.disp = disp; ca.weights = weights; ca.z_out = nullptr;
  ca.noise = raw_noise;
  composite_kernel<<<(unsigned)((R + COMP_RAYS - 1) / COMP_RAYS), COMP_RAYS * 32, 0, st>>>(ca);
  CKL("composite");
  return 0;
}

int dsnerf_warp_points(dsnerf_ctx* ctx, const float* pts, int64_t P, float* xyz_cano, uint8_t* transparent, int32_t* idx, void* stream) {
  if (int e = check_ready(ctx, true)) return e;
  if (P < 0) return fail(ctx, DSNERF_ERR_INVALID, "bad size");
  if (P == 0) return 0;
  if (!pts || !xyz_cano) return fail(ctx, DSNERF_ERR_INVALID, "null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  if (int e = join_grid(ctx, st)) return e;
  warp_points_kernel<<<(unsigned)((P + 127) / 128), 128, 0, st>>>(pts, P, ctx->posed.as<float>(), ctx->canon.as<float>(), ctx->faces.as<int>(),
                                                                  ctx->g_posed.g, ctx->F, ctx->g_posed.cent.as<float>(), xyz_cano, transparent, idx);
  CKL("warp_points");
  return 0;
}

namespace {
__global__ void pack_active_kernel(const float* __restrict__ xyz, const uint8_t* __restrict__ skip, int64_t P, float4* __restrict__ active,
                                   unsigned long long* __restrict__ counter, float* __restrict__ zero_a, float* __restrict__ zero_b, int nb) {
  int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool act = s < P && !(skip && skip[s]);
  if (s < P && !act) {
    if (zero_a) zero_a[s] = 0.f;
    if (zero_b) for (int k = 0; k < nb; ++k) zero_b[nb * s + k] = 0.f;
  }
  unsigned m = __ballot_sync(0xffffffffu, act);
  if (m) {
    int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (act) active[base + __popc(m & ((1u << lane) - 1))] = make_float4(xyz[3 * s], xyz[3 * s + 1], xyz[3 * s + 2], __int_as_float((int)s));
  }
}
__global__ void scatter_points_kernel(const float4* __restrict__ active, const float4* __restrict__ raw, const unsigned long long* __restrict__ n,
                                      float* __restrict__ color, float* __restrict__ density) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)*n) return;
  int s = __float_as_int(active[t].w);
  float4 r = raw[s];
  if (color) { color[3 * (int64_t)s] = r.x; color[3 * (int64_t)s + 1] = r.y; color[3 * (int64_t)s + 2] = r.z; }
  density[s] = r.w;
}
__global__ void scatter_density2_kernel(const float4* __restrict__ active, const float4* __restrict__ mlp_a, const unsigned long long* __restrict__ n,
                                        float* __restrict__ density) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)*n) return;
  density[__float_as_int(active[t].w)] = mlp_a[t].x;
}
}  // namespace

int dsnerf_query_density(dsnerf_ctx* ctx, const float* xyz_cano, const uint8_t* transparent, int64_t P, float* density, unsigned flags,
                         void* stream) {
  if (int e = check_ready(ctx, true)) return e;
  if (P < 0 || P > 0x7fffffff) return fail(ctx, DSNERF_ERR_INVALID, "bad size");
  if (P == 0) return 0;
  if (!xyz_cano || !density) return fail(ctx, DSNERF_ERR_INVALID, "null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  if (int e = ensure_workspace(ctx, P, 1)) return e;
  unsigned long long* cnt = ctx->counters.as<unsigned long long>();
  CK(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * 4, st));
  unsigned blocks = (unsigned)((P + 255) / 256);
  pack_active_kernel<<<blocks, 256, 0, st>>>(xyz_cano, transparent, P, ctx->active.as<float4>(), cnt, density, nullptr, 0);
  CKL("pack_active");
  if (int e = launch_mlp(ctx, cnt, 0, flags, 1, st)) return e;
  scatter_density2_kernel<<<blocks, 256, 0, st>>>(ctx->active.as<float4>(), ctx->mlp_a.as<float4>(), cnt, density);
  CKL("scatter_density");
  return 0;
}

int dsnerf_eval_points(dsnerf_ctx* ctx, const float* xyz_world, const float* xyz_cano, const float* view_dir, int64_t P, float* color,
                       float* density, unsigned flags, void* stream) {
  if (int e = check_ready(ctx, true)) return e;
  if (P < 0 || P > 0x7fffffff) return fail(ctx, DSNERF_ERR_INVALID, "bad size");
  if (P == 0) return 0;
  if (!xyz_world || !xyz_cano || !view_dir || !color || !density) return fail(ctx, DSNERF_ERR_INVALID, "null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  if (int e = ensure_workspace(ctx, P, 1)) return e;
  if (int e = join_grid(ctx, st)) return e;  // the normal matrices are built next to the grid
  unsigned long long* cnt = ctx->counters.as<unsigned long long>();
  CK(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * 4, st));
  unsigned blocks = (unsigned)((P + 255) / 256);
  pack_active_kernel<<<blocks, 256, 0, st>>>(xyz_cano, nullptr, P, ctx->active.as<float4>(), cnt, nullptr, nullptr, 0);
  CKL("pack_active");
  if (int e = launch_mlp(ctx, cnt, 0, flags, 0, st)) return e;
  ShadeArgs sa = base_shade_args(ctx);
  sa.n_active = cnt;
  sa.xyz_world = xyz_world;
  sa.view_dir = view_dir;
  sa.N = 1;
  if (int e = launch_shade(ctx, sa, flags, st)) return e;
  scatter_points_kernel<<<blocks, 256, 0, st>>>(ctx->active.as<float4>(), ctx->raw.as<float4>(), cnt, color, density);
  CKL("scatter_points");
  return 0;
}

int dsnerf_resample(dsnerf_ctx* ctx, const float* z_in, const float* weights, int64_t R, int N, int n_importance, float* z_out, void* stream) {
  if (!ctx) return DSNERF_ERR_INVALID;
  if (R < 0 || N < 3 || N > 1024 || n_importance < 1 || n_importance > 1024) return fail(ctx, DSNERF_ERR_INVALID, "bad sizes");
  if (R == 0) return 0;
  if (!z_in || !weights || !z_out) return fail(ctx, DSNERF_ERR_INVALID, "null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  size_t smem = sizeof(float) * (size_t)(2 * N + n_importance) * 4;
  resample_kernel<<<(unsigned)((R + 3) / 4), 128, smem, st>>>(z_in, weights, R, N, n_importance, z_out);
  CKL("resample");
  return 0;
}

int dsnerf_ppts_to_pts(dsnerf_ctx* ctx, const float* ppts, const float* bw, const float* A, int64_t P, float* out, void* stream) {
  if (!ctx) return DSNERF_ERR_INVALID;
  if (P < 0) return fail(ctx, DSNERF_ERR_INVALID, "bad size");
  if (P == 0) return 0;
  if (!ppts || !bw || !A || !out) return fail(ctx, DSNERF_ERR_INVALID, "null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  lbs_inverse_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(ppts, bw, A, P, out);
  CKL("lbs_inverse");
  return 0;
}

int dsnerf_camera_rays(dsnerf_ctx* ctx, int H, int W, const double* K, const double* R, const double* T, const float* bounds, float* ray_o,
                       float* ray_d, float* near, float* far, uint8_t* mask_at_box, void* stream) {
  if (!ctx) return DSNERF_ERR_INVALID;
  if (H < 0 || W < 0) return fail(ctx, DSNERF_ERR_INVALID, "bad size");
  if (H == 0 || W == 0) return 0;
  if (!K || !R || !T || !bounds || !ray_o || !ray_d || !near || !far || !mask_at_box) return fail(ctx, DSNERF_ERR_INVALID, "null pointer");
  CameraArgs c{};
  // K^-1 through the cofactors, in double (np.linalg.inv(K), rays_utils.py:24)
  const double a = K[0], b = K[1], cc = K[2], d = K[3], e = K[4], f = K[5], g = K[6], h = K[7], i = K[8];
  const double det = a * (e * i - f * h) - b * (d * i - f * g) + cc * (d * h - e * g);
  if (!(fabs(det) > 0.0)) return fail(ctx, DSNERF_ERR_INVALID, "singular K");
  const double inv[9] = {(e * i - f * h) / det, (cc * h - b * i) / det, (b * f - cc * e) / det, (f * g - d * i) / det, (a * i - cc * g) / det,
                         (cc * d - a * f) / det, (d * h - e * g) / det, (b * g - a * h) / det, (a * e - b * d) / det};
  for (int k = 0; k < 9; ++k) { c.kinv[k] = inv[k]; c.rot[k] = R[k]; }
  for (int k = 0; k < 3; ++k) {
    c.t[k] = T[k];
    c.origin[k] = -(R[k] * T[0] + R[3 + k] * T[1] + R[6 + k] * T[2]);  // -R^T T
    c.lo[k] = (double)bounds[k] - 0.01;
    c.hi[k] = (double)bounds[3 + k] + 0.01;
  }
  c.H = H;
  c.W = W;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  const int64_t P = (int64_t)H * W;
  camera_rays_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(c, ray_o, ray_d, near, far, mask_at_box);
  CKL("camera_rays");
  return 0;
}

namespace {
__global__ void expand_mask_kernel(const unsigned* __restrict__ bits, int64_t P, uint8_t* __restrict__ transparent) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < P) transparent[s] = ((bits[s >> 5] >> (s & 31)) & 1u) ? 0 : 1;
}
}  // namespace

int dsnerf_last_transparent_mask(dsnerf_ctx* ctx, int64_t R, int N, uint8_t* transparent, void* stream) {
  if (!ctx) return DSNERF_ERR_INVALID;
  if (R < 0 || N < 1) return fail(ctx, DSNERF_ERR_INVALID, "bad sizes");
  if (R == 0) return 0;
  if (!transparent) return fail(ctx, DSNERF_ERR_INVALID, "null pointer");
  if (ctx->stats.rays != R || ctx->stats.samples != R * (int64_t)N || !ctx->ray_mask.p)
    return fail(ctx, DSNERF_ERR_STATE, "no render of that shape precedes this call on the context");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  const int64_t P = R * (int64_t)N;
  expand_mask_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(ctx->ray_mask.as<unsigned>(), P, transparent);
  CKL("expand_mask");
  return 0;
}

int dsnerf_tensor_path_active(const dsnerf_ctx* ctx) {
  if (!ctx || !ctx->have_weights) return DSNERF_ERR_STATE;
  return ctx->tw.fp16_ok ? (ctx->tw.rgb3 ? 3 : 1) : 0;
}

int dsnerf_mlp_kernel_variant(const dsnerf_ctx* ctx) { return ctx ? ctx->mlp_variant : DSNERF_ERR_INVALID; }

int dsnerf_get_stats(dsnerf_ctx* ctx, dsnerf_stats_t* out) {
  if (!ctx || !out) return DSNERF_ERR_INVALID;
  if (ctx->stats.rays > 0) {
    CK(cudaEventSynchronize(ctx->stats_ready));
    ctx->stats.evaluated_samples = ctx->ert_mode ? (int64_t)(ctx->h_ert[0] + ctx->h_ert[5] + ctx->h_ert[6] + ctx->h_ert[7]) : (int64_t)ctx->h_counters[0];
    ctx->stats.nn_candidates = (int64_t)ctx->h_counters[1];
    ctx->stats.reserved = (int32_t)std::min<unsigned long long>(ctx->h_counters[2], 0x7fffffffull);
    ctx->stats.algorithmic_flop = 1804544.0 * (double)ctx->stats.evaluated_samples;
  }
  *out = ctx->stats;
  return 0;
}

int dsnerf_profile(dsnerf_ctx* ctx, int enable) {
  if (!ctx) return DSNERF_ERR_INVALID;
  ctx->profile = enable & 255;
  return 0;
}

namespace {
// SM clock as the SM sees it: cycles of clock64 per nanosecond of the global timer over a ~30 us spin (one warp)
__global__ void sm_clock_probe_kernel(float* __restrict__ out, unsigned spin_ns) {
  if (threadIdx.x != 0) return;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  long long c0 = clock64();
  do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); } while (t1 - t0 < spin_ns);
  long long c1 = clock64();
  out[0] = (float)((double)(c1 - c0) * 1.0e3 / (double)(t1 - t0));
  out[1] = (float)(t1 - t0) * 1e-3f;
}
}  // namespace

int dsnerf_debug_sm_clock(dsnerf_ctx* ctx, float* d_out2, void* stream) {
  if (!ctx || !d_out2) return DSNERF_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  sm_clock_probe_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_out2, 30000u);
  CKL("sm_clock_probe");
  return 0;
}

int dsnerf_debug_active(dsnerf_ctx* ctx, int64_t capacity, float* active_xyz_id, int32_t* active_tri, int64_t* n_out) {
  if (!ctx || !n_out) return DSNERF_ERR_INVALID;
  CK(cudaSetDevice(ctx->device));
  CK(cudaDeviceSynchronize());
  unsigned long long n = 0;
  if (!ctx->counters.p) { *n_out = 0; return 0; }
  CK(cudaMemcpy(&n, ctx->counters.p, sizeof(n), cudaMemcpyDeviceToHost));
  *n_out = (int64_t)n;
  int64_t m = std::min<int64_t>((int64_t)n, capacity);
  if (m > 0 && active_xyz_id) CK(cudaMemcpy(active_xyz_id, ctx->active.p, sizeof(float4) * m, cudaMemcpyDeviceToHost));
  if (m > 0 && active_tri) CK(cudaMemcpy(active_tri, ctx->active_tri.p, sizeof(int32_t) * m, cudaMemcpyDeviceToHost));
  return 0;
}

int dsnerf_debug_table(dsnerf_ctx* ctx, int which, int* out16) {
  if (!ctx || !out16) return DSNERF_ERR_INVALID;
  MeshGrid& mg = which ? ctx->g_canon : ctx->g_posed;
  if (!mg.pool_used.p) return DSNERF_ERR_STATE;
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out16, mg.pool_used.p, sizeof(int) * 16, cudaMemcpyDeviceToHost));
  out16[11] = mg.ntab;
  out16[12] = mg.ncell;
  return 0;
}

int dsnerf_debug_tc_timing(dsnerf_ctx* ctx, long long* out64) {
  if (!ctx || !out64 || !ctx->tc_timing.p) return DSNERF_ERR_INVALID;
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out64, ctx->tc_timing.p, sizeof(long long) * 128, cudaMemcpyDeviceToHost));
  return 0;
}

int dsnerf_profile_read(dsnerf_ctx* ctx, double* mlp_ms, int64_t* mlp_launches, int reset) {
  if (!ctx) return DSNERF_ERR_INVALID;
  for (auto& p : ctx->pending) {
    CK(cudaEventSynchronize(p.second));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, p.first, p.second));
    ctx->mlp_ms += ms;
    ctx->mlp_launches += 1;
    cudaEventDestroy(p.first);
    cudaEventDestroy(p.second);
  }
  ctx->pending.clear();
  if (mlp_ms) *mlp_ms = ctx->mlp_ms;
  if (mlp_launches) *mlp_launches = ctx->mlp_launches;
  if (reset) { ctx->mlp_ms = 0; ctx->mlp_launches = 0; }
  return 0;
}

}  // extern "C"
