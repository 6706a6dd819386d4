// SpaceNet forward + analytic density gradient on the 5th-gen tensor cores (tcgen05, sm_100a).
//
// Restates model/spacenet.py:93-148 (SpaceNet.forward: PE -> 4x256 -> [h|PE] -> 3x256 -> density,
// rgb head) and :301-311 (gradient = d density / d xyz_cano, which the reference gets from
// autograd) for a tile of 128 canonical points per CTA (UMMA M = 128, cta_group::1).
//
// Numerics
//   * forward GEMMs run 3 MMAs per k-step (x_hi*w_hi + x_hi*w_lo + x_lo*w_hi, fp16 operands, fp32
//     accumulate in TMEM): single-pass fp16/bf16/tf32 misses the 1e-4 parity bound by >10x because a
//     rounding-sized change of a pre-activation flips ReLUs of the gradient path;
//   * the backward-data chain for the normal runs single-pass fp16 through W^T using the ReLU bits
//     recorded in the forward pass; the seed is w_dens / max|w_dens| so every gradient stays inside
//     fp16 range (the normal is scale invariant).
//
// Pipeline (one CTA = one SM, 10 warps)
//   * measured on B200 (profiles/): an MMA whose A operand comes from shared memory occupies the tensor pipe for
//     >= 128 cycles whatever N is (the 128 x 16 A tile is read at 32 B/cycle), so every MMA here is N = 256
//     (128 cycles = the ideal rate); an earlier version with two N = 128 halves ran 1.7-2x slower per layer;
//   * the A operand (fp16 hi + lo, K-major no-swizzle core matrices, 128 KB) is rewritten in place by the
//     epilogue; two 256-column TMEM accumulators alternate between layers, so the epilogue of layer l (8 warps:
//     tcgen05.ld -> bias/ReLU/mask bits/hi-lo split -> st.shared) overlaps the MMAs of layer l+1: the epilogue
//     publishes its first 128 columns early and the next layer's first 8 k-steps start on them;
//   * weights are pre-packed on the host into the exact shared-memory image, in consumption order, and streamed
//     by one elected lane with cp.async.bulk (TMA engine) into a 4x16 KB mbarrier ring (cluster-ready: with
//     TC_CLUSTER = 2 each CTA fetches half of every slab and multicasts it; the ring is latency- not
//     bandwidth-bound on B200, so the default is 1);
//   * one elected lane of a converged warp issues all tcgen05.mma (slab-templated: descriptors are
//     "base + immediate"); tcgen05.commit releases ring slots and publishes finished accumulators;
//   * ReLU bits live in registers (28 words per thread), d sigma / d PE of layer 4 is parked in the idle A-lo
//     region during the backward chain, cross-thread partial sums go through 8 TMEM cells per row.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace dsn {

constexpr int TC_TILE = 128;
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = (TC_EPI_WARPS + 2) * 32;
constexpr int TC_STAGES = 4;
constexpr int TC_CLUSTER = 1;
constexpr uint32_t TC_STAGE_BYTES = 16384;
// shared memory map (bytes)
constexpr uint32_t SM_A_HI = 0;          // A operand, fp16 hi: [K/8 = 32 chunks][128 rows][8]
constexpr uint32_t SM_A_LO = 65536;      // A operand, fp16 lo (forward only); backward: fp32 stash of layer 4's d sigma / d PE
constexpr uint32_t SM_PE_HI = 131072;    // positional encoding hi, 8 chunks
constexpr uint32_t SM_PE_LO = 147456;    // positional encoding lo
constexpr uint32_t SM_RING = 163840;
constexpr uint32_t SM_BAR = SM_RING + TC_STAGES * TC_STAGE_BYTES;  // 229376
constexpr uint32_t TC_SMEM = SM_BAR + 128;
constexpr uint32_t A_CHUNK = TC_TILE * 16;  // bytes between consecutive 8-wide K chunks of an A operand
// tensor memory map (32-bit columns)
constexpr uint32_t TM_ACC = 256;    // accumulator b (= op & 1) starts at column b * TM_ACC
constexpr uint32_t TM_XCH = 256;    // 8 columns of cross-thread partial sums at the very end of a tile (accumulator 1 is idle then)
constexpr uint32_t TM_COLS = 512;

constexpr int TC_NUM_OPS = 15;
enum { A_ACT = 0, A_PE = 1, A_ACT_PE = 2 };

enum { K_FWD = 0, K_RGB = 1, K_BWD = 2, K_BW4 = 3, K_BW0 = 4 };

struct TcOp {
  uint32_t src_off;     // byte offset of the op's first slab in the packed weight blob
  uint32_t slab_bytes;  // bytes per slab (<= TC_STAGE_BYTES, multiple of 32)
  uint16_t n_slabs;
  uint16_t ksteps;      // k-steps (of 16) per slab
  uint8_t kind;         // K_FWD (N=256, 3-pass), K_RGB (N=128, 3-pass), K_BWD (N=256), K_BW4 (N=256 + 64 extra), K_BW0 (N=64)
  uint8_t a_src;        // A_ACT, A_PE, A_ACT_PE (k-steps >= 16 come from the PE region)
  uint8_t pad[2];
};

__constant__ TcOp c_tc_ops[TC_NUM_OPS];

struct TcParams {
  const uint8_t* wpack;    // packed fp16 weights
  const float* bias;       // [7][256]; row 0 = per-frame folded bias (code + pose feature)
  const float* b_rgb1;     // [128]
  const float* w_rgb2;     // [3][128]
  const float* w_dens;     // [256]
  const float* seed;       // [256] w_dens / seed_scale
  float b_rgb2[3];
  float b_dens;
  float seed_scale;
  const float4* active;
  const unsigned long long* n_active_ptr;
  int64_t n_active_host;
  float4* out_a;
  float4* out_g;
  int density_only;
  long long* timing;       // debug: clock64 stamps of CTA 0 / first tile, NULL in production
};

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar, uint16_t cta_mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t mbar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(mbar), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]   (A: lane = row, 8 columns of packed fp16 pairs per k-step)
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 B contiguous; SBO = 128 B between 8-row groups,
// LBO = byte distance between the two 8-wide K chunks of one k-step.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc(uint32_t n) {  // kind::f16, A=B=F16, D=F32, K-major, M=128
  return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
#define DSN_R32(r) \
  "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), \
  "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),  \
  "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),  \
  "=r"(r[31])
#define DSN_RW32(r) \
  "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), \
  "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]),  \
  "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]),  \
  "+r"(r[31])
#define DSN_IN32(r) \
  "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),   \
  "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),     \
  "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : DSN_R32(r) : "r"(taddr));
}
// wait for outstanding tcgen05.ld; the registers are in/out operands so the compiler cannot hoist a read above the wait
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&r)[32]) { asm volatile("tcgen05.wait::ld.sync.aligned;" : DSN_RW32(r)::"memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  tmem_ld32_nowait(taddr, r);
  tmem_wait_ld(r);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), DSN_IN32(r) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])::"memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// one lane of a fully converged warp (the warp stays converged, so descriptors live in uniform registers)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// split two fp32 values into fp16 hi and lo pairs (x = hi + lo up to 2^-22 relative)
__device__ __forceinline__ void split_h2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __half2 h = __floats2half2_rn(a, b);
  float2 f = __half22float2(h);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = pack_h2(a - f.x, b - f.y);
}

// ReLU bits of the 7 forward layers for this thread's 128 columns (4 words per layer), kept in registers.
struct ReluBits {
  uint32_t w[7][4];
  __device__ __forceinline__ void put(int layer, int i, uint32_t v) {
#pragma unroll
    for (int l = 0; l < 7; ++l)
      if (l == layer) w[l][i] = v;
  }
  __device__ __forceinline__ uint32_t get(int layer, int i) const {
    uint32_t v = 0;
#pragma unroll
    for (int l = 0; l < 7; ++l)
      if (l == layer) v = w[l][i];
    return v;
  }
};

// Issue all MMAs of one weight slab (KSTEPS k-steps) and release its ring slot.  Shapes are template
// parameters so that every descriptor is "slab base + immediate": the issuing lane spends a couple of
// uniform-datapath adds per tcgen05.mma instead of rebuilding 64-bit descriptors.
//   a_word / a_lo_word / b_word: low 32 bits of the A-hi / A-lo / B-hi shared-memory descriptors at the slab's first k-step
template <int ROWS, int KSTEPS, bool THREE, int NMMA, bool EXTRA>
__device__ __forceinline__ void issue_slab(uint32_t d_main, uint32_t d_extra, uint32_t a_word, uint32_t a_lo_word, uint32_t b_word,
                                           uint32_t first_acc, uint32_t empty_bar, uint16_t mc_mask) {
  constexpr uint32_t DHI = (128u >> 4) | (1u << 14);             // descriptor bits 32..63: SBO = 128 B, version 1
  constexpr uint32_t A_STEP = (2 * A_CHUNK) >> 4;                // one k-step along K in the A operand
  constexpr uint32_t B_STEP = (2 * ROWS * 16) >> 4;              // one k-step in the slab
  constexpr uint32_t B_LO = (KSTEPS * 2 * ROWS * 16) >> 4;       // hi part -> lo part of the slab
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(NMMA >> 3) << 17) | ((128u >> 4) << 24);
  constexpr uint32_t IDESC_X = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
  if (elect_one()) {
#pragma unroll
    for (int j = 0; j < KSTEPS; ++j) {
      const uint64_t da = ((uint64_t)DHI << 32) | (a_word + j * A_STEP);
      const uint64_t db = ((uint64_t)DHI << 32) | (b_word + j * B_STEP);
      tc_mma_ss(d_main, da, db, IDESC, j == 0 ? first_acc : 1u);
      if (THREE) {
        const uint64_t dbl = ((uint64_t)DHI << 32) | (b_word + j * B_STEP + B_LO);
        const uint64_t dal = ((uint64_t)DHI << 32) | (a_lo_word + j * A_STEP);
        tc_mma_ss(d_main, da, dbl, IDESC, 1u);
        tc_mma_ss(d_main, dal, db, IDESC, 1u);
      }
      if (EXTRA) {
        const uint64_t dbx = ((uint64_t)DHI << 32) | (b_word + j * B_STEP + ((256 * 16) >> 4));
        tc_mma_ss(d_extra, da, dbx, IDESC_X, j == 0 ? first_acc : 1u);
      }
    }
    tc_commit_mc(empty_bar, mc_mask);  // frees the ring slot (in every CTA of the cluster) once these MMAs have read it
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------ kernel
__global__ void __cluster_dims__(TC_CLUSTER, 1, 1) __launch_bounds__(TC_THREADS, 1) mlp_tc_kernel(TcParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + SM_BAR;             // [TC_STAGES]   weights landed
  const uint32_t bar_empty = bar_full + 8 * TC_STAGES;  // [TC_STAGES]   ring slot consumed by every CTA of the cluster
  const uint32_t bar_acc = bar_empty + 8 * TC_STAGES;   //               accumulator complete (MMA -> epilogue)
  const uint32_t bar_a = bar_acc + 8;                   // [4]           epilogue done with column quarter q (epilogue -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 8 * (2 * TC_STAGES + 6));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, TC_CLUSTER); }
    mbar_init(bar_acc, 1);
    for (int q4 = 0; q4 < 4; ++q4) mbar_init(bar_a + 8 * q4, TC_EPI_WARPS * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_EPI_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peers' barriers are initialised before anyone multicasts into / arrives on them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t cta_rank = cluster_ctarank();
  const uint16_t mc_mask = (uint16_t)((1u << TC_CLUSTER) - 1);

  const int64_t n_active = P.n_active_ptr ? (int64_t)*P.n_active_ptr : P.n_active_host;
  const int64_t n_tiles = (n_active + TC_TILE - 1) / TC_TILE;
  // every CTA runs the same number of iterations (the ring of a cluster advances in lockstep); tiles past the end are dummies
  const int64_t n_iter = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int n_ops = P.density_only ? 7 : TC_NUM_OPS;

  if (warp == TC_EPI_WARPS + 1) {
    // =============================== weight loader (one elected lane of a converged warp) =======
    {
      uint32_t stage = 0, phase = 0;
      for (int64_t it = 0; it < n_iter; ++it) {
        for (int op = 0; op < n_ops; ++op) {
          const TcOp o = c_tc_ops[op];
          const uint32_t part = o.slab_bytes / TC_CLUSTER;
          const uint8_t* src = P.wpack + o.src_off + cta_rank * part;
          for (int s = 0; s < o.n_slabs; ++s) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            if (elect_one()) {
              mbar_expect_tx(bar_full + 8 * stage, o.slab_bytes);
              bulk_g2s_mc(sbase + SM_RING + stage * TC_STAGE_BYTES + cta_rank * part, src + (size_t)s * o.slab_bytes, part,
                          bar_full + 8 * stage, mc_mask);
            }
            __syncwarp();
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == TC_EPI_WARPS) {
    // =============================== MMA issuer (one elected lane of a converged warp) =========
    {
      uint32_t stage = 0, phase = 0, a_phase = 0;
      constexpr uint32_t A_LBO = (A_CHUNK >> 4) << 16;  // LBO field of every A descriptor
      for (int64_t it = 0; it < n_iter; ++it) {
        for (int op = 0; op < n_ops; ++op) {
          const TcOp o = c_tc_ops[op];
          const bool mstamp = P.timing && blockIdx.x == 0 && it == 0 && lane == 0;
          long long w_full = 0, w_a = 0, t_op0 = mstamp ? clock64() : 0;
          const uint32_t d_main = tmem + (uint32_t)(op & 1) * TM_ACC;
          const uint32_t d_extra = tmem + (uint32_t)((op & 1) ^ 1) * TM_ACC;
          // the producer's epilogue publishes its 256 output columns in four quarters of 64 (= 4 k-steps of this op's A operand)
          int waited = 0;
          auto need_quarters = [&](int nq) {
            if (waited >= nq) return;
            long long t0 = mstamp ? clock64() : 0;
            while (waited < nq) { mbar_wait(bar_a + 8 * waited, a_phase); ++waited; }
            tc_fence_after();
            if (mstamp) w_a += clock64() - t0;
          };
          need_quarters(o.a_src == A_PE ? 4 : 1);
          const uint32_t rows = o.kind == K_FWD || o.kind == K_BWD ? 256u : (o.kind == K_RGB ? 128u : (o.kind == K_BW4 ? 320u : 64u));
          const uint32_t b_lbo = ((rows * 16) >> 4) << 16;
          uint32_t kk = 0;
          for (int s = 0; s < o.n_slabs; ++s, kk += o.ksteps) {
            if (kk < 16) need_quarters(min(4, (int)((kk + o.ksteps - 1) >> 2) + 1));
            long long t1 = mstamp ? clock64() : 0;
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            if (mstamp) w_full += clock64() - t1;
            const uint32_t b_word = b_lbo | ((sbase + SM_RING + stage * TC_STAGE_BYTES) >> 4);
            const uint32_t ebar = bar_empty + 8 * stage;
            const bool from_pe = (o.a_src == A_PE) || (o.a_src == A_ACT_PE && kk >= 16);
            const uint32_t pc = (o.a_src == A_PE ? kk : kk - 16) * 2;
            const uint32_t a_word = A_LBO | ((from_pe ? sbase + SM_PE_HI + pc * A_CHUNK : sbase + SM_A_HI + kk * 2 * A_CHUNK) >> 4);
            const uint32_t a_lo_word = A_LBO | ((from_pe ? sbase + SM_PE_LO + pc * A_CHUNK : sbase + SM_A_LO + kk * 2 * A_CHUNK) >> 4);
            const uint32_t first_acc = (uint32_t)(kk > 0);
            switch (o.kind) {
              case K_FWD: issue_slab<256, 1, true, 256, false>(d_main, 0u, a_word, a_lo_word, b_word, first_acc, ebar, mc_mask); break;
              case K_RGB: issue_slab<128, 2, true, 128, false>(d_main, 0u, a_word, a_lo_word, b_word, first_acc, ebar, mc_mask); break;
              case K_BWD: issue_slab<256, 2, false, 256, false>(d_main, 0u, a_word, 0u, b_word, first_acc, ebar, mc_mask); break;
              case K_BW4: issue_slab<320, 1, false, 256, true>(d_main, d_extra, a_word, 0u, b_word, first_acc, ebar, mc_mask); break;
              default: issue_slab<64, 8, false, 64, false>(d_main, 0u, a_word, 0u, b_word, first_acc, ebar, mc_mask); break;
            }
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
          need_quarters(4);
          if (elect_one()) tc_commit(bar_acc);  // accumulator (and the extra columns of K_BW4) complete
          __syncwarp();
          a_phase ^= 1;
          if (mstamp) { P.timing[64 + 3 * op] = w_full; P.timing[65 + 3 * op] = w_a; P.timing[66 + 3 * op] = clock64() - t_op0; }
        }
      }
    }
  } else {
    // =============================== epilogue warps ===========================================
    // warp w: TMEM lane quarter q = w % 4 (rows 32q..32q+31), column sub-block sub = w / 4: of every 64-column
    // accumulator quarter this thread handles columns [32*sub, 32*sub + 32) of its row.
    const int q = warp & 3, sub = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t acc_phase = 0;
    ReluBits relu;
#pragma unroll
    for (int l = 0; l < 7; ++l)
#pragma unroll
      for (int i = 0; i < 4; ++i) relu.w[l][i] = 0;
    float4 pt_next = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((int64_t)blockIdx.x * TC_TILE + row < n_active) pt_next = P.active[(int64_t)blockIdx.x * TC_TILE + row];
    for (int64_t it = 0; it < n_iter; ++it) {
      const int64_t tile = blockIdx.x + it * gridDim.x;
      const int64_t base = tile * TC_TILE;
      const bool live = base + row < n_active;
      const float4 pt = pt_next;
      {  // fetch the next tile's point now: its DRAM latency hides behind this tile
        const int64_t nb = (tile + gridDim.x) * TC_TILE + row;
        pt_next = nb < n_active ? P.active[nb] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const float xs[3] = {pt.x, pt.y, pt.z};
      // ---- positional encoding (model/dimension_kernel.py:5-35) = A operand of layer 0 and tail of layer 4's
      {
        auto put = [&](int c, float v) {
          __half hh = __float2half_rn(v);
          __half ll = __float2half_rn(v - __half2float(hh));
          uint32_t off = (uint32_t)(c >> 3) * A_CHUNK + row * 16 + (c & 7) * 2;
          *reinterpret_cast<__half*>(smem + SM_PE_HI + off) = hh;
          *reinterpret_cast<__half*>(smem + SM_PE_LO + off) = ll;
        };
        if (sub == 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) put(c, xs[c]);
        } else {
          put(63, 0.f);
        }
        for (int k = sub * 5; k < sub * 5 + 5; ++k) {
          float f = (float)(1 << k);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float sn, cs;
            sincosf(xs[c] * f, &sn, &cs);
            put(3 + 6 * k + c, sn);
            put(6 + 6 * k + c, cs);
          }
        }
      }
      const bool stamp = P.timing && blockIdx.x == 0 && it == 0 && threadIdx.x == 0;
      if (stamp) P.timing[0] = clock64();
      fence_proxy_async();
      tc_fence_before();
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) mbar_arrive(bar_a + 8 * q4);

      float sigma_part = 0.f, e0 = 0.f, e1 = 0.f, e2 = 0.f;
      float* stash = reinterpret_cast<float*>(smem + SM_A_LO);  // [64 PE columns][128 rows] fp32, backward chain only
      for (int op = 0; op < n_ops; ++op) {
        mbar_wait(bar_acc, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        const uint32_t t_accb = t_lane + (uint32_t)(op & 1) * TM_ACC;
        if (op == 10) {
          // layer 4 backward also produced d sigma / d PE (64 columns) in the idle accumulator: park this thread's 32 of them
          uint32_t v[32];
          tmem_ld32(t_lane + (uint32_t)((op & 1) ^ 1) * TM_ACC + sub * 32, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) stash[(sub * 32 + i) * TC_TILE + row] = __uint_as_float(v[i]);
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          if (stamp && (q4 & 1) == 0) P.timing[1 + 4 * op + q4] = clock64();
          const int col0 = q4 * 64 + sub * 32;  // this thread's 32 output columns of the quarter
          if (op <= 6) {
            // ---------- forward layer: bias + ReLU, record ReLU bits, split to fp16 hi / lo = next A operand (in place)
            const float* __restrict__ bias = P.bias + op * 256;
            uint32_t v[32];
            tmem_ld32(t_accb + col0, v);
            uint32_t m = 0;
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0 + i));
              float h0 = fmaxf(__uint_as_float(v[i]) + b.x, 0.f), h1 = fmaxf(__uint_as_float(v[i + 1]) + b.y, 0.f);
              float h2 = fmaxf(__uint_as_float(v[i + 2]) + b.z, 0.f), h3 = fmaxf(__uint_as_float(v[i + 3]) + b.w, 0.f);
              // ReLU bit = (h != 0): h >= +0, so bits(h) + 0x7fffffff carries into bit 31 iff h > 0; shifted in MSB-first
              m = __funnelshift_l(__float_as_uint(h0) + 0x7fffffffu, m, 1);
              m = __funnelshift_l(__float_as_uint(h1) + 0x7fffffffu, m, 1);
              m = __funnelshift_l(__float_as_uint(h2) + 0x7fffffffu, m, 1);
              m = __funnelshift_l(__float_as_uint(h3) + 0x7fffffffu, m, 1);
              if (op == 6) {
                const float4 wd = __ldg(reinterpret_cast<const float4*>(P.w_dens + col0 + i));
                sigma_part = fmaf(wd.x, h0, sigma_part); sigma_part = fmaf(wd.y, h1, sigma_part);
                sigma_part = fmaf(wd.z, h2, sigma_part); sigma_part = fmaf(wd.w, h3, sigma_part);
              }
              split_h2(h0, h1, hi[i / 2], lo[i / 2]);
              split_h2(h2, h3, hi[i / 2 + 1], lo[i / 2 + 1]);
            }
            relu.put(op, q4, __brev(m));  // element i of the chunk -> bit i
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const uint32_t off = (uint32_t)(col0 / 8 + t) * A_CHUNK + row * 16;
              *reinterpret_cast<uint4*>(smem + SM_A_HI + off) = make_uint4(hi[4 * t], hi[4 * t + 1], hi[4 * t + 2], hi[4 * t + 3]);
              *reinterpret_cast<uint4*>(smem + SM_A_LO + off) = make_uint4(lo[4 * t], lo[4 * t + 1], lo[4 * t + 2], lo[4 * t + 3]);
            }
          } else if (op == 7) {
            // ---------- rgb head: relu(acc[0:128] + b) -> Linear(128,3) partials (model/spacenet.py:75-80) ...
            if (q4 < 2) {
              uint32_t vr[32];
              tmem_ld32(t_accb + col0, vr);
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(P.b_rgb1 + col0 + i));
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(P.w_rgb2 + col0 + i));
                const float4 w1 = __ldg(reinterpret_cast<const float4*>(P.w_rgb2 + 128 + col0 + i));
                const float4 w2 = __ldg(reinterpret_cast<const float4*>(P.w_rgb2 + 256 + col0 + i));
                const float r0 = fmaxf(__uint_as_float(vr[i]) + b.x, 0.f), r1 = fmaxf(__uint_as_float(vr[i + 1]) + b.y, 0.f);
                const float r2 = fmaxf(__uint_as_float(vr[i + 2]) + b.z, 0.f), r3 = fmaxf(__uint_as_float(vr[i + 3]) + b.w, 0.f);
                e0 = fmaf(w0.x, r0, e0); e0 = fmaf(w0.y, r1, e0); e0 = fmaf(w0.z, r2, e0); e0 = fmaf(w0.w, r3, e0);
                e1 = fmaf(w1.x, r0, e1); e1 = fmaf(w1.y, r1, e1); e1 = fmaf(w1.z, r2, e1); e1 = fmaf(w1.w, r3, e1);
                e2 = fmaf(w2.x, r0, e2); e2 = fmaf(w2.y, r1, e2); e2 = fmaf(w2.z, r2, e2); e2 = fmaf(w2.w, r3, e2);
              }
            }
            // ... and the seed of the backward chain (the rgb head's MMAs have read the A operand): G6 = (w_dens / scale) * relu'(a6)
            {
              const uint32_t mb = relu.w[6][q4];
              uint32_t hi[16];
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 sd = __ldg(reinterpret_cast<const float4*>(P.seed + col0 + i));
                float g0 = ((mb >> i) & 1u) ? sd.x : 0.f, g1 = ((mb >> (i + 1)) & 1u) ? sd.y : 0.f;
                float g2 = ((mb >> (i + 2)) & 1u) ? sd.z : 0.f, g3 = ((mb >> (i + 3)) & 1u) ? sd.w : 0.f;
                hi[i / 2] = pack_h2(g0, g1);
                hi[i / 2 + 1] = pack_h2(g2, g3);
              }
#pragma unroll
              for (int t = 0; t < 4; ++t)
                *reinterpret_cast<uint4*>(smem + SM_A_HI + (uint32_t)(col0 / 8 + t) * A_CHUNK + row * 16) =
                    make_uint4(hi[4 * t], hi[4 * t + 1], hi[4 * t + 2], hi[4 * t + 3]);
            }
          } else if (op <= 13) {
            // ---------- backward layer: G_{l-1} = (G_l W_l) * relu'(a_{l-1}), fp16 single pass, hi only
            const uint32_t mb = relu.get(13 - op, q4);  // op 8 -> layer 5 ... op 13 -> layer 0
            uint32_t v[32];
            tmem_ld32(t_accb + col0, v);
            uint32_t hi[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float g0 = ((mb >> i) & 1u) ? __uint_as_float(v[i]) : 0.f;
              float g1 = ((mb >> (i + 1)) & 1u) ? __uint_as_float(v[i + 1]) : 0.f;
              hi[i / 2] = pack_h2(g0, g1);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t)
              *reinterpret_cast<uint4*>(smem + SM_A_HI + (uint32_t)(col0 / 8 + t) * A_CHUNK + row * 16) =
                  make_uint4(hi[4 * t], hi[4 * t + 1], hi[4 * t + 2], hi[4 * t + 3]);
          }
          if (stamp && (q4 & 1) == 1) P.timing[1 + 4 * op + q4] = clock64();
          if (op != n_ops - 1) {
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(bar_a + 8 * q4);
          }
        }
      }
      // ---------- tile outputs: both accumulator halves of the last op are complete
      {
        float gx[3] = {0.f, 0.f, 0.f};
        if (!P.density_only) {
          // d sigma / d PE (64 columns) -> chain rule through the encoding; this thread owns octaves 5*sub .. 5*sub+4
          // layer-0 part: accumulator 0 (op 14), columns 0..63; layer-4 part: the stash written at op 10
          uint32_t g0[32], g1[32];
          tmem_ld32(t_lane, g0);
          tmem_ld32(t_lane + 32, g1);
          auto gpe = [&](int c) -> float { return __uint_as_float(c < 32 ? g0[c] : g1[c - 32]) + stash[c * TC_TILE + row]; };
          // sin/cos are still in the PE region (written by this very thread)
          auto pe_val = [&](int col) -> float {
            const uint32_t off = (uint32_t)(col >> 3) * A_CHUNK + row * 16 + (col & 7) * 2;
            return __half2float(*reinterpret_cast<const __half*>(smem + SM_PE_HI + off)) +
                   __half2float(*reinterpret_cast<const __half*>(smem + SM_PE_LO + off));
          };
          if (sub == 0) { gx[0] = gpe(0); gx[1] = gpe(1); gx[2] = gpe(2); }
#pragma unroll
          for (int kq = 0; kq < 5; ++kq) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const int k0 = kq, k1 = kq + 5;
              const float gs = sub ? gpe(3 + 6 * k1 + c) : gpe(3 + 6 * k0 + c);
              const float gc = sub ? gpe(6 + 6 * k1 + c) : gpe(6 + 6 * k0 + c);
              const float f = sub ? (float)(1 << k1) : (float)(1 << k0);
              const int ks = sub ? k1 : k0;
              const float sn = pe_val(3 + 6 * ks + c), cs = pe_val(6 + 6 * ks + c);
              gx[c] = fmaf((gs * cs - gc * sn), f, gx[c]);
            }
          }
        }
        // cross-thread sums (the two threads of a row live in warps w and w+4): through 8 TMEM cells of the row
        if (sub == 1) {
          uint32_t x[8] = {__float_as_uint(sigma_part), __float_as_uint(e0), __float_as_uint(e1), __float_as_uint(e2),
                           __float_as_uint(gx[0]), __float_as_uint(gx[1]), __float_as_uint(gx[2]), 0u};
          tmem_st8(t_lane + TM_XCH, x);
          tmem_wait_st();
          tc_fence_before();
        }
        epi_bar();
        if (sub == 0) {
          tc_fence_after();
          uint32_t x[8];
          tmem_ld8(t_lane + TM_XCH, x);
          if (live) {
            const float sigma = sigma_part + __uint_as_float(x[0]) + P.b_dens;
            if (P.density_only) {
              P.out_a[base + row] = make_float4(sigma, 0.f, 0.f, 0.f);
            } else {
              P.out_a[base + row] = make_float4(sigma, e0 + __uint_as_float(x[1]) + P.b_rgb2[0], e1 + __uint_as_float(x[2]) + P.b_rgb2[1],
                                                e2 + __uint_as_float(x[3]) + P.b_rgb2[2]);
              P.out_g[base + row] = make_float4((gx[0] + __uint_as_float(x[4])) * P.seed_scale, (gx[1] + __uint_as_float(x[5])) * P.seed_scale,
                                                (gx[2] + __uint_as_float(x[6])) * P.seed_scale, 0.f);
            }
          }
          tc_fence_before();
        }
        if (stamp) P.timing[62] = clock64();
        epi_bar();  // nobody overwrites the PE region / stash / TM_XCH of this tile before everyone is done with them
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA leaves while a peer may still multicast into its ring or arrive on its barriers
  if (warp == TC_EPI_WARPS) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host
struct TcWeights {
  void* d_pack = nullptr;
  float* d_f32 = nullptr;  // biases 0..6 (7x256; row 0 per frame), b_rgb1 (128), w_rgb2 (384), w_dens (256), seed (256)
  float b_rgb2[3] = {0, 0, 0};
  float b_dens = 0.f, seed_scale = 1.f;
  TcOp ops[TC_NUM_OPS];
  static constexpr int F32_BRGB1 = 7 * 256, F32_WRGB2 = F32_BRGB1 + 128, F32_WDENS = F32_WRGB2 + 384, F32_SEED = F32_WDENS + 256,
                       F32_TOTAL = F32_SEED + 256;
  float* bias0_slot() const { return d_f32; }

  void release() {
    if (d_pack) cudaFree(d_pack);
    if (d_f32) cudaFree(d_f32);
    d_pack = nullptr;
    d_f32 = nullptr;
  }

  // B[n][k] (rows x K) packed as slabs of `ksteps` k-steps:
  //   [hi: (2*ksteps chunks) x rows x 8 halves][lo: same]   -- exactly the image the UMMA descriptors address
  static void pack_op(std::vector<__half>& blob, TcOp& op, int kind, int a_src, int rows, int K, int ksteps, bool with_lo,
                      const std::vector<float>& B) {
    const int n_slabs = K / (16 * ksteps);
    const size_t part = (size_t)ksteps * 2 * rows * 8;  // halves per hi (or lo) part
    const size_t slab = part * (with_lo ? 2 : 1);
    while (blob.size() % 64) blob.push_back(__float2half_rn(0.f));  // 128-byte aligned slabs
    op.src_off = (uint32_t)(blob.size() * sizeof(__half));
    op.slab_bytes = (uint32_t)(slab * sizeof(__half));
    op.n_slabs = (uint16_t)n_slabs;
    op.ksteps = (uint16_t)ksteps;
    op.kind = (uint8_t)kind;
    op.a_src = (uint8_t)a_src;
    op.pad[0] = op.pad[1] = 0;
    const size_t base = blob.size();
    blob.resize(base + slab * n_slabs);
    for (int n = 0; n < rows; ++n)
      for (int k = 0; k < K; ++k) {
        const int s = k / (16 * ksteps), j = (k / 16) % ksteps, cc = (k % 16) / 8, e = k % 8;
        const size_t off = base + (size_t)s * slab + ((size_t)(j * 2 + cc) * rows + n) * 8 + e;
        const float w = B[(size_t)n * K + k];
        const __half hh = __float2half_rn(w);
        blob[off] = hh;
        if (with_lo) blob[off + part] = __float2half_rn(w - __half2float(hh));
      }
  }

  int stage(const std::vector<float>& w0, const std::vector<float>& w1, const std::vector<float>& w2, const std::vector<float>& w3,
            const std::vector<float>& w4, const std::vector<float>& w5, const std::vector<float>& w6, const std::vector<float>& b1,
            const std::vector<float>& b2, const std::vector<float>& b3, const std::vector<float>& b4, const std::vector<float>& b5,
            const std::vector<float>& b6, const std::vector<float>& wd, float bd, const std::vector<float>& wr1,
            const std::vector<float>& br1, const std::vector<float>& wr2, const std::vector<float>& br2) {
    const std::vector<float>* W[7] = {&w0, &w1, &w2, &w3, &w4, &w5, &w6};
    std::vector<__half> blob;
    std::vector<float> B;
    int oi = 0;
    // forward layers 0..6: N = 256, 3-pass (hi + lo), one k-step per 16 KB slab
    for (int l = 0; l < 7; ++l) {
      const int in_dim = l == 0 ? 87 : (l == 4 ? 319 : 256);
      const int K = l == 0 ? 64 : (l == 4 ? 320 : 256);
      B.assign((size_t)256 * K, 0.f);
      for (int n = 0; n < 256; ++n)
        for (int k = 0; k < (l == 0 ? 63 : in_dim); ++k) B[(size_t)n * K + k] = (*W[l])[(size_t)n * in_dim + (l == 0 ? 8 + k : k)];
      pack_op(blob, ops[oi++], K_FWD, l == 0 ? A_PE : (l == 4 ? A_ACT_PE : A_ACT), 256, K, 1, true, B);
    }
    // rgb head first layer: 256 -> 128
    B.assign((size_t)128 * 256, 0.f);
    for (int n = 0; n < 128; ++n)
      for (int k = 0; k < 256; ++k) B[(size_t)n * 256 + k] = wr1[(size_t)n * 256 + k];
    pack_op(blob, ops[oi++], K_RGB, A_ACT, 128, 256, 2, true, B);
    // backward through layers 6..1: B[n][k] = W[k][n] (n = input index, k = output index), 1-pass
    for (int l = 6; l >= 1; --l) {
      if (l == 4) {  // 256 hidden inputs + 63 PE inputs (+1 pad): rows 256..319 feed the extra N = 64 MMA
        B.assign((size_t)320 * 256, 0.f);
        for (int n = 0; n < 319; ++n)
          for (int k = 0; k < 256; ++k) B[(size_t)n * 256 + k] = w4[(size_t)k * 319 + n];
        pack_op(blob, ops[oi++], K_BW4, A_ACT, 320, 256, 1, false, B);
      } else {
        B.assign((size_t)256 * 256, 0.f);
        for (int n = 0; n < 256; ++n)
          for (int k = 0; k < 256; ++k) B[(size_t)n * 256 + k] = (*W[l])[(size_t)k * 256 + n];
        pack_op(blob, ops[oi++], K_BWD, A_ACT, 256, 256, 2, false, B);
      }
    }
    // layer 0 backward, PE columns only (N = 64); added to the stashed layer-4 PE gradient in the tile's last stage
    B.assign((size_t)64 * 256, 0.f);
    for (int n = 0; n < 63; ++n)
      for (int k = 0; k < 256; ++k) B[(size_t)n * 256 + k] = w0[(size_t)k * 87 + 8 + n];
    pack_op(blob, ops[oi++], K_BW0, A_ACT, 64, 256, 8, false, B);
    if (oi != TC_NUM_OPS) return (int)cudaErrorUnknown;
    for (int i = 0; i < TC_NUM_OPS; ++i)
      if (ops[i].slab_bytes > TC_STAGE_BYTES || (ops[i].slab_bytes & 31) || (ops[i].src_off & 15)) return (int)cudaErrorInvalidValue;
    release();
    cudaError_t e = cudaMalloc(&d_pack, blob.size() * sizeof(__half));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpy(d_pack, blob.data(), blob.size() * sizeof(__half), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return (int)e;
    std::vector<float> f(F32_TOTAL, 0.f);
    const std::vector<float>* bs[6] = {&b1, &b2, &b3, &b4, &b5, &b6};
    for (int l = 0; l < 6; ++l)
      for (int i = 0; i < 256; ++i) f[(l + 1) * 256 + i] = (*bs[l])[i];
    for (int i = 0; i < 128; ++i) f[F32_BRGB1 + i] = br1[i];
    for (int i = 0; i < 384; ++i) f[F32_WRGB2 + i] = wr2[i];
    float mx = 0.f;
    for (int i = 0; i < 256; ++i) mx = fmaxf(mx, fabsf(wd[i]));
    seed_scale = mx > 0.f ? mx : 1.f;
    for (int i = 0; i < 256; ++i) { f[F32_WDENS + i] = wd[i]; f[F32_SEED + i] = wd[i] / seed_scale; }
    e = cudaMalloc(reinterpret_cast<void**>(&d_f32), f.size() * sizeof(float));
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpy(d_f32, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return (int)e;
    for (int i = 0; i < 3; ++i) b_rgb2[i] = br2[i];
    b_dens = bd;
    e = cudaMemcpyToSymbol(c_tc_ops, ops, sizeof(ops));
    return (int)e;
  }
};

inline void tc_configure() { cudaFuncSetAttribute(mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM); }

inline int tc_launch(TcWeights& w, long long* timing, const float4* active, const unsigned long long* n_active_ptr, int64_t n_active_host,
                     float4* out_a, float4* out_g, int density_only, int sm_count, cudaStream_t st) {
  TcParams p{};
  p.wpack = reinterpret_cast<const uint8_t*>(w.d_pack);
  p.bias = w.d_f32;
  p.b_rgb1 = w.d_f32 + TcWeights::F32_BRGB1;
  p.w_rgb2 = w.d_f32 + TcWeights::F32_WRGB2;
  p.w_dens = w.d_f32 + TcWeights::F32_WDENS;
  p.seed = w.d_f32 + TcWeights::F32_SEED;
  for (int i = 0; i < 3; ++i) p.b_rgb2[i] = w.b_rgb2[i];
  p.b_dens = w.b_dens;
  p.seed_scale = w.seed_scale;
  p.active = active;
  p.n_active_ptr = n_active_ptr;
  p.n_active_host = n_active_host;
  p.out_a = out_a;
  p.out_g = out_g;
  p.density_only = density_only;
  p.timing = timing;
  mlp_tc_kernel<<<sm_count / TC_CLUSTER * TC_CLUSTER, TC_THREADS, TC_SMEM, st>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace dsn
