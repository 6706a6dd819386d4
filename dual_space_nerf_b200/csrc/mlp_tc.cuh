// SpaceNet forward + analytic density gradient on the 5th-gen tensor cores (tcgen05, sm_100a).
//
// Restates model/spacenet.py:93-148 (SpaceNet.forward: PE -> 4x256 -> [h|PE] -> 3x256 -> density,
// rgb head) and :301-311 (gradient = d density / d xyz_cano, which the reference gets from
// autograd) for a tile of 128 canonical points per CTA:
//
//   * activations are the A operand (K-major, no-swizzle core-matrix layout) in shared memory,
//     kept as an fp16 hi/lo pair; weights are the B operand, pre-packed on the host into the
//     exact shared-memory image and streamed slab by slab with cp.async.bulk (TMA engine) into a
//     4-stage mbarrier ring; accumulators (128 x 256 fp32) live in TMEM;
//   * forward GEMMs run 3 MMAs per k-step (x_hi*w_hi + x_hi*w_lo + x_lo*w_hi, fp32 accumulate):
//     single-pass fp16/bf16/tf32 misses the 1e-4 parity bound by >10x because a rounding-sized
//     change of a pre-activation flips ReLUs of the gradient path;
//   * the backward-data chain for the normal runs single-pass fp16 through W^T with the ReLU
//     bit masks parked in TMEM (tcgen05.st / tcgen05.ld); the seed is normalised by max|w_dens|
//     so every gradient stays inside fp16 range (the normal is scale invariant);
//   * warp roles: 8 epilogue warps (TMEM -> registers -> bias/ReLU/mask/split -> smem), one MMA
//     issuer thread, one weight-loader thread.
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
constexpr int TC_CLUSTER = 2;
constexpr uint32_t TC_STAGE_BYTES = 16384;
// shared memory map (bytes)
constexpr uint32_t SM_ACT_HI = 0;
constexpr uint32_t SM_ACT_LO = 65536;
constexpr uint32_t SM_PE_HI = 131072;
constexpr uint32_t SM_PE_LO = 147456;
constexpr uint32_t SM_RING = 163840;
constexpr uint32_t SM_BAR = SM_RING + TC_STAGES * TC_STAGE_BYTES;  // 229376
constexpr uint32_t TC_SMEM = SM_BAR + 128;
constexpr uint32_t A_CHUNK = TC_TILE * 16;  // bytes between consecutive 8-wide K chunks of the A operand
// tensor memory map (columns)
constexpr uint32_t TM_ACC = 0;      // 256 columns: layer accumulator
constexpr uint32_t TM_GPE = 256;    // 64 columns: d sigma / d PE
constexpr uint32_t TM_MASK = 320;   // 7 layers x 8 columns: ReLU bits
constexpr uint32_t TM_COLS = 512;

constexpr int TC_NUM_OPS = 15;
enum { A_ACT = 0, A_PE = 1, A_ACT_PE = 2 };

struct TcOp {
  uint32_t src_off;     // byte offset of the op's first slab in the packed weight blob
  uint32_t slab_bytes;  // bytes per slab (<= TC_STAGE_BYTES)
  uint16_t n_slabs;
  uint16_t ksteps;      // k-steps (of 16) per slab
  uint16_t n_main;      // N of the main MMA (rows 0..n_main-1 of the slab)
  uint16_t n_extra;     // N of the extra MMA into TM_GPE (rows n_main..), 0 if none
  uint8_t passes;       // 3 = hi/lo split, 1 = hi only
  uint8_t a_src;        // A_ACT, A_PE, A_ACT_PE (k-steps >= 16 come from the PE region)
  uint8_t main_to_gpe;  // main MMA accumulates into TM_GPE instead of TM_ACC
  uint8_t gpe_accum;    // first k-step of the TM_GPE MMA accumulates (1) or overwrites (0)
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
  long long* timing;       // debug: clock64 stamps of CTA 0 / tile 0 (2 per op + 2), NULL in production
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
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
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 B contiguous; SBO = 128 B between 8-row groups,
// LBO = byte distance between the two 8-wide K chunks of one k-step.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc(uint32_t n) {  // kind::f16, A=B=F16, D=F32, K-major, M=128
  return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
// wait for outstanding tcgen05.ld; the registers are listed as in/out operands so that the
// compiler cannot hoist a read of them above the wait
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                 "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                 "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :: "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
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

// ------------------------------------------------------------------------------------------ kernel
// Launched as clusters of TC_CLUSTER CTAs: every weight slab is fetched from L2 once per cluster (each CTA
// loads 1/TC_CLUSTER of it and multicasts it into all CTAs' rings).  Measured: without multicast the kernel is
// bound by L2->SM weight streaming (~30 B/cycle/SM), see profiles/.
__global__ void __cluster_dims__(TC_CLUSTER, 1, 1) __launch_bounds__(TC_THREADS, 1) mlp_tc_kernel(TcParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + SM_BAR;             // [TC_STAGES]
  const uint32_t bar_empty = bar_full + 8 * TC_STAGES;  // [TC_STAGES]
  const uint32_t bar_acc = bar_empty + 8 * TC_STAGES;   // accumulator ready (MMA -> epilogue)
  const uint32_t bar_a = bar_acc + 8;                   // A operand ready / accumulator free (epilogue -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 8 * (2 * TC_STAGES + 2));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, TC_CLUSTER); }
    mbar_init(bar_acc, 1);
    mbar_init(bar_a, TC_EPI_WARPS * 32);
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
  // every CTA runs the same number of iterations (the ring of a cluster advances in lockstep); tiles past the
  // end are dummies (no live rows)
  const int64_t n_iter = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int n_ops = P.density_only ? 7 : TC_NUM_OPS;

  if (warp == TC_EPI_WARPS + 1) {
    // =============================== weight loader (one thread) ===============================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int64_t it = 0; it < n_iter; ++it) {
        for (int op = 0; op < n_ops; ++op) {
          const TcOp o = c_tc_ops[op];
          const uint32_t part = o.slab_bytes / TC_CLUSTER;
          const uint8_t* src = P.wpack + o.src_off + cta_rank * part;
          for (int s = 0; s < o.n_slabs; ++s) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);  // every CTA of the cluster has consumed this slot
            mbar_expect_tx(bar_full + 8 * stage, o.slab_bytes);
            bulk_g2s_mc(sbase + SM_RING + stage * TC_STAGE_BYTES + cta_rank * part, src + (size_t)s * o.slab_bytes, part,
                        bar_full + 8 * stage, mc_mask);
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == TC_EPI_WARPS) {
    // =============================== MMA issuer (one thread) ==================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, a_phase = 0;
      for (int64_t it = 0; it < n_iter; ++it) {
        for (int op = 0; op < n_ops; ++op) {
          const TcOp o = c_tc_ops[op];
          const uint32_t rows = o.n_main + o.n_extra;
          const uint32_t lbo_b = rows * 16;
          const uint32_t hi_bytes = o.ksteps * 2 * lbo_b;
          const uint32_t idesc_main = make_idesc(o.n_main);
          const uint32_t idesc_extra = make_idesc(64);
          const uint32_t d_main = tmem + (o.main_to_gpe ? TM_GPE : TM_ACC);
          mbar_wait(bar_a, a_phase);
          a_phase ^= 1;
          tc_fence_after();
          uint32_t kk = 0;
          for (int s = 0; s < o.n_slabs; ++s) {
            mbar_wait(bar_full + 8 * stage, phase);
            tc_fence_after();
            const uint32_t sb = sbase + SM_RING + stage * TC_STAGE_BYTES;
            for (int j = 0; j < o.ksteps; ++j, ++kk) {
              uint32_t a_hi, a_lo;
              if (o.a_src == A_PE || (o.a_src == A_ACT_PE && kk >= 16)) {
                uint32_t c = (o.a_src == A_PE ? kk : kk - 16) * 2;
                a_hi = sbase + SM_PE_HI + c * A_CHUNK;
                a_lo = sbase + SM_PE_LO + c * A_CHUNK;
              } else {
                a_hi = sbase + SM_ACT_HI + kk * 2 * A_CHUNK;
                a_lo = sbase + SM_ACT_LO + kk * 2 * A_CHUNK;
              }
              const uint64_t da_hi = smem_desc(a_hi, A_CHUNK);
              const uint64_t db_hi = smem_desc(sb + j * 2 * lbo_b, lbo_b);
              const uint32_t acc_first = o.main_to_gpe ? (uint32_t)(o.gpe_accum | (kk > 0)) : (uint32_t)(kk > 0);
              tc_mma(d_main, da_hi, db_hi, idesc_main, acc_first);
              if (o.passes == 3) {
                const uint64_t da_lo = smem_desc(a_lo, A_CHUNK);
                const uint64_t db_lo = smem_desc(sb + hi_bytes + j * 2 * lbo_b, lbo_b);
                tc_mma(d_main, da_hi, db_lo, idesc_main, 1u);
                tc_mma(d_main, da_lo, db_hi, idesc_main, 1u);
              }
              if (o.n_extra) {
                const uint64_t db_x = smem_desc(sb + j * 2 * lbo_b + o.n_main * 16, lbo_b);
                tc_mma(tmem + TM_GPE, da_hi, db_x, idesc_extra, (uint32_t)(o.gpe_accum | (kk > 0)));
              }
            }
            tc_commit_mc(bar_empty + 8 * stage, mc_mask);  // frees the slot in every CTA of the cluster once these MMAs have read it
            if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
          }
          tc_commit(bar_acc);  // accumulator complete
        }
      }
    }
  } else {
    // =============================== epilogue warps ===========================================
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t acc_phase = 0;
    // [128][2][8] floats of cross-half partial sums.  Lives in the A-lo region, which is idle once the last
    // 3-pass op (rgb head) has run; the density-only path ends earlier and uses the PE region (idle after op 4).
    float* xch = reinterpret_cast<float*>(smem + (P.density_only ? SM_PE_HI : SM_ACT_LO));
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
      // ---- positional encoding (model/dimension_kernel.py:5-35) as the A operand of op 0 / tail of op 4
      {
        auto put = [&](int c, float v) {
          __half h = __float2half_rn(v);
          __half l = __float2half_rn(v - __half2float(h));
          uint32_t off = (uint32_t)(c >> 3) * A_CHUNK + row * 16 + (c & 7) * 2;
          *reinterpret_cast<__half*>(smem + SM_PE_HI + off) = h;
          *reinterpret_cast<__half*>(smem + SM_PE_LO + off) = l;
        };
        if (half == 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) put(c, xs[c]);
        } else {
          put(63, 0.f);
        }
        for (int k = half * 5; k < half * 5 + 5; ++k) {
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
      mbar_arrive(bar_a);

      float sigma_part = 0.f;
      for (int op = 0; op < n_ops; ++op) {
        mbar_wait(bar_acc, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
        if (stamp) P.timing[1 + 2 * op] = clock64();
        if (op <= 6) {
          // ---------- forward layer: bias + ReLU, record mask bits, split to fp16 hi/lo -> next A operand
          const float* __restrict__ bias = P.bias + op * 256;
          uint32_t mbits[4];
          uint32_t vbuf[2][32];
          tmem_ld32_nowait(t_lane + TM_ACC + half * 128, vbuf[0]);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int col0 = half * 128 + c * 32;
            tmem_wait_ld(vbuf[c & 1]);
            if (c + 1 < 4) tmem_ld32_nowait(t_lane + TM_ACC + col0 + 32, vbuf[(c + 1) & 1]);  // overlaps the math below
            uint32_t (&v)[32] = vbuf[c & 1];
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
            mbits[c] = __brev(m);  // element i of the chunk -> bit i
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const uint32_t off = (uint32_t)(col0 / 8 + t) * A_CHUNK + row * 16;
              *reinterpret_cast<uint4*>(smem + SM_ACT_HI + off) = make_uint4(hi[4 * t], hi[4 * t + 1], hi[4 * t + 2], hi[4 * t + 3]);
              *reinterpret_cast<uint4*>(smem + SM_ACT_LO + off) = make_uint4(lo[4 * t], lo[4 * t + 1], lo[4 * t + 2], lo[4 * t + 3]);
            }
          }
          if (!P.density_only) tmem_st4(t_lane + TM_MASK + op * 8 + half * 4, mbits);
        } else if (op == 7) {
          // ---------- rgb head: relu(acc[0:128] + b) -> Linear(128,3) partials (model/spacenet.py:75-80);
          //            then seed the backward chain: G6 = (w_dens / scale) * relu'(a6)
          float e0 = 0.f, e1 = 0.f, e2 = 0.f;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            const int col0 = half * 64 + c * 32;
            uint32_t v[32];
            tmem_ld32(t_lane + TM_ACC + col0, v);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(P.b_rgb1 + col0 + i));
              const float4 w0 = __ldg(reinterpret_cast<const float4*>(P.w_rgb2 + col0 + i));
              const float4 w1 = __ldg(reinterpret_cast<const float4*>(P.w_rgb2 + 128 + col0 + i));
              const float4 w2 = __ldg(reinterpret_cast<const float4*>(P.w_rgb2 + 256 + col0 + i));
              const float r0 = fmaxf(__uint_as_float(v[i]) + b.x, 0.f), r1 = fmaxf(__uint_as_float(v[i + 1]) + b.y, 0.f);
              const float r2 = fmaxf(__uint_as_float(v[i + 2]) + b.z, 0.f), r3 = fmaxf(__uint_as_float(v[i + 3]) + b.w, 0.f);
              e0 = fmaf(w0.x, r0, e0); e0 = fmaf(w0.y, r1, e0); e0 = fmaf(w0.z, r2, e0); e0 = fmaf(w0.w, r3, e0);
              e1 = fmaf(w1.x, r0, e1); e1 = fmaf(w1.y, r1, e1); e1 = fmaf(w1.z, r2, e1); e1 = fmaf(w1.w, r3, e1);
              e2 = fmaf(w2.x, r0, e2); e2 = fmaf(w2.y, r1, e2); e2 = fmaf(w2.z, r2, e2); e2 = fmaf(w2.w, r3, e2);
            }
          }
          float* x = xch + (row * 2 + half) * 8;
          x[0] = sigma_part; x[1] = e0; x[2] = e1; x[3] = e2;
          uint32_t mb[4];
          tmem_ld4(t_lane + TM_MASK + 6 * 8 + half * 4, mb);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int col0 = half * 128 + c * 32;
            uint32_t hi[16];
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              const float4 sd = __ldg(reinterpret_cast<const float4*>(P.seed + col0 + i));
              float g0 = ((mb[c] >> i) & 1u) ? sd.x : 0.f, g1 = ((mb[c] >> (i + 1)) & 1u) ? sd.y : 0.f;
              float g2 = ((mb[c] >> (i + 2)) & 1u) ? sd.z : 0.f, g3 = ((mb[c] >> (i + 3)) & 1u) ? sd.w : 0.f;
              hi[i / 2] = pack_h2(g0, g1);
              hi[i / 2 + 1] = pack_h2(g2, g3);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t)
              *reinterpret_cast<uint4*>(smem + SM_ACT_HI + (uint32_t)(col0 / 8 + t) * A_CHUNK + row * 16) =
                  make_uint4(hi[4 * t], hi[4 * t + 1], hi[4 * t + 2], hi[4 * t + 3]);
          }
        } else if (op <= 13) {
          // ---------- backward layer: G_{l-1} = (G_l W_l) * relu'(a_{l-1}), fp16 single pass
          const int mask_layer = 13 - op;  // op 8 -> layer 5 ... op 13 -> layer 0
          uint32_t mb[4];
          tmem_ld4(t_lane + TM_MASK + mask_layer * 8 + half * 4, mb);
          uint32_t vbuf[2][32];
          tmem_ld32_nowait(t_lane + TM_ACC + half * 128, vbuf[0]);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int col0 = half * 128 + c * 32;
            tmem_wait_ld(vbuf[c & 1]);
            if (c + 1 < 4) tmem_ld32_nowait(t_lane + TM_ACC + col0 + 32, vbuf[(c + 1) & 1]);
            uint32_t (&v)[32] = vbuf[c & 1];
            uint32_t hi[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float g0 = ((mb[c] >> i) & 1u) ? __uint_as_float(v[i]) : 0.f;
              float g1 = ((mb[c] >> (i + 1)) & 1u) ? __uint_as_float(v[i + 1]) : 0.f;
              hi[i / 2] = pack_h2(g0, g1);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t)
              *reinterpret_cast<uint4*>(smem + SM_ACT_HI + (uint32_t)(col0 / 8 + t) * A_CHUNK + row * 16) =
                  make_uint4(hi[4 * t], hi[4 * t + 1], hi[4 * t + 2], hi[4 * t + 3]);
          }
        } else {
          // ---------- op 14: d sigma / d PE complete in TM_GPE -> chain rule through the encoding
          uint32_t g0[32], g1[32];
          tmem_ld32(t_lane + TM_GPE, g0);
          tmem_ld32(t_lane + TM_GPE + 32, g1);
          auto gpe = [&](int c) -> float { return __uint_as_float(c < 32 ? g0[c] : g1[c - 32]); };
          float gx[3] = {0.f, 0.f, 0.f};
          if (half == 0) { gx[0] = gpe(0); gx[1] = gpe(1); gx[2] = gpe(2); }
          // sin/cos of this thread's five octaves are still in the PE region (written by this very thread)
          auto pe_val = [&](int col) -> float {
            const uint32_t off = (uint32_t)(col >> 3) * A_CHUNK + row * 16 + (col & 7) * 2;
            return __half2float(*reinterpret_cast<const __half*>(smem + SM_PE_HI + off)) +
                   __half2float(*reinterpret_cast<const __half*>(smem + SM_PE_LO + off));
          };
#pragma unroll
          for (int kk = 0; kk < 5; ++kk) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const int k0 = kk, k1 = kk + 5;
              const float gs = half ? gpe(3 + 6 * k1 + c) : gpe(3 + 6 * k0 + c);
              const float gc = half ? gpe(6 + 6 * k1 + c) : gpe(6 + 6 * k0 + c);
              const float f = half ? (float)(1 << k1) : (float)(1 << k0);
              const int ks = half ? k1 : k0;
              const float sn = pe_val(3 + 6 * ks + c), cs = pe_val(6 + 6 * ks + c);
              gx[c] = fmaf((gs * cs - gc * sn), f, gx[c]);
            }
          }
          float* x = xch + (row * 2 + half) * 8;
          x[4] = gx[0]; x[5] = gx[1]; x[6] = gx[2];
        }
        if (stamp) P.timing[2 + 2 * op] = clock64();
        if (op == n_ops - 1) {
          // ---------- tile outputs
          if (P.density_only) {
            float* x = xch + (row * 2 + half) * 8;
            x[0] = sigma_part;
          }
          epi_bar();
          if (half == 0 && live) {
            const float* a = xch + (row * 2) * 8;
            const float* b = a + 8;
            float sigma = a[0] + b[0] + P.b_dens;
            if (P.density_only) {
              P.out_a[base + row] = make_float4(sigma, 0.f, 0.f, 0.f);
            } else {
              P.out_a[base + row] = make_float4(sigma, a[1] + b[1] + P.b_rgb2[0], a[2] + b[2] + P.b_rgb2[1], a[3] + b[3] + P.b_rgb2[2]);
              P.out_g[base + row] = make_float4((a[4] + b[4]) * P.seed_scale, (a[5] + b[5]) * P.seed_scale, (a[6] + b[6]) * P.seed_scale, 0.f);
            }
          }
          epi_bar();  // xch aliases the PE region that the next tile overwrites
        } else {
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(bar_a);
        }
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

  // B[n][k] packed as slabs of `ksteps` k-steps: [hi: (2*ksteps chunks) x rows x 8 halves][lo: same]
  static void pack_op(std::vector<__half>& blob, TcOp& op, int rows, int K, int ksteps, bool with_lo, const std::vector<float>& B) {
    const int n_slabs = K / (16 * ksteps);
    const size_t part = (size_t)ksteps * 2 * rows * 8;  // halves per hi (or lo) part
    const size_t slab = part * (with_lo ? 2 : 1);
    op.src_off = (uint32_t)(blob.size() * sizeof(__half));
    op.slab_bytes = (uint32_t)(slab * sizeof(__half));
    op.n_slabs = (uint16_t)n_slabs;
    op.ksteps = (uint16_t)ksteps;
    size_t base = blob.size();
    blob.resize(base + slab * n_slabs);
    for (int n = 0; n < rows; ++n)
      for (int k = 0; k < K; ++k) {
        const int s = k / (16 * ksteps), j = (k / 16) % ksteps, cc = (k % 16) / 8, e = k % 8;
        const size_t off = base + (size_t)s * slab + ((size_t)(j * 2 + cc) * rows + n) * 8 + e;
        const float w = B[(size_t)n * K + k];
        const __half h = __float2half_rn(w);
        blob[off] = h;
        if (with_lo) blob[off + part] = __float2half_rn(w - __half2float(h));
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
    auto set = [&](TcOp& o, int n_main, int n_extra, int passes, int a_src, int to_gpe, int gpe_acc) {
      o.n_main = (uint16_t)n_main; o.n_extra = (uint16_t)n_extra; o.passes = (uint8_t)passes; o.a_src = (uint8_t)a_src;
      o.main_to_gpe = (uint8_t)to_gpe; o.gpe_accum = (uint8_t)gpe_acc;
    };
    // forward layers 0..6
    for (int l = 0; l < 7; ++l) {
      const int in_dim = l == 0 ? 87 : (l == 4 ? 319 : 256);
      const int K = l == 0 ? 64 : (l == 4 ? 320 : 256);
      B.assign((size_t)256 * K, 0.f);
      for (int n = 0; n < 256; ++n)
        for (int k = 0; k < (l == 0 ? 63 : in_dim); ++k) B[(size_t)n * K + k] = (*W[l])[(size_t)n * in_dim + (l == 0 ? 8 + k : k)];
      pack_op(blob, ops[oi], 256, K, 1, true, B);
      set(ops[oi], 256, 0, 3, l == 0 ? A_PE : (l == 4 ? A_ACT_PE : A_ACT), 0, 0);
      ++oi;
    }
    // rgb head first layer: 256 -> 128
    B.assign((size_t)128 * 256, 0.f);
    for (int n = 0; n < 128; ++n)
      for (int k = 0; k < 256; ++k) B[(size_t)n * 256 + k] = wr1[(size_t)n * 256 + k];
    pack_op(blob, ops[oi], 128, 256, 2, true, B);
    set(ops[oi], 128, 0, 3, A_ACT, 0, 0);
    ++oi;
    // backward: layers 6,5 (B[n][k] = W[k][n])
    for (int l = 6; l >= 1; --l) {
      if (l == 4) {
        B.assign((size_t)320 * 256, 0.f);
        for (int n = 0; n < 319; ++n)
          for (int k = 0; k < 256; ++k) B[(size_t)n * 256 + k] = w4[(size_t)k * 319 + n];
        pack_op(blob, ops[oi], 320, 256, 1, false, B);
        set(ops[oi], 256, 64, 1, A_ACT, 0, 0);
      } else {
        B.assign((size_t)256 * 256, 0.f);
        for (int n = 0; n < 256; ++n)
          for (int k = 0; k < 256; ++k) B[(size_t)n * 256 + k] = (*W[l])[(size_t)k * 256 + n];
        pack_op(blob, ops[oi], 256, 256, 2, false, B);
        set(ops[oi], 256, 0, 1, A_ACT, 0, 0);
      }
      ++oi;
    }
    // layer 0 backward, PE columns only, accumulated onto the layer-4 PE gradient
    B.assign((size_t)64 * 256, 0.f);
    for (int n = 0; n < 63; ++n)
      for (int k = 0; k < 256; ++k) B[(size_t)n * 256 + k] = w0[(size_t)k * 87 + 8 + n];
    pack_op(blob, ops[oi], 64, 256, 8, false, B);
    set(ops[oi], 64, 0, 1, A_ACT, 1, 1);
    ++oi;
    if (oi != TC_NUM_OPS) return (int)cudaErrorUnknown;
    for (int i = 0; i < TC_NUM_OPS; ++i)
      if (ops[i].slab_bytes > TC_STAGE_BYTES || (ops[i].slab_bytes & 15) || (ops[i].src_off & 15)) return (int)cudaErrorInvalidValue;
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
