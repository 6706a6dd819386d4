// fp32 SIMT evaluation of SpaceNet + analytic density gradient.
//
// This is the *verification* kernel (flag DSNERF_MLP_FP32_SIMT): plain fp32 FMAs,
// no tensor cores, used to cross-check the tcgen05 kernel at sizes the CPU oracle
// cannot reach.  Same inputs/outputs as the tensor-core kernel in mlp_tc.cuh.
//
// Restates model/spacenet.py:93-148 (SpaceNet.forward) and :301-311 (gradient):
// the reference obtains d(density)/d(xyz_cano) from autograd; here it is the
// explicit backward-data chain through the stored ReLU masks and the positional
// encoding.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dsn {

constexpr int PE_DIM = 63;
constexpr int PE_PAD = 64;
constexpr int WIDTH = 256;
constexpr int HEAD = 128;

// fp32 device weights for the SIMT path (all row-major)
struct SimtWeights {
  const float* wt[7];   // forward, transposed [K][256]: K = 64, 256, 256, 256, 320, 256, 256
  const float* w[7];    // backward, [256][Kb]: Kb = 64 (PE cols of layer 0), 256, 256, 256, 320, 256, 256
  const float* bias[7]; // [256]; bias[0] is the per-frame folded bias (code + pose feature)
  const float* w_dens;  // [256]
  float b_dens;
  const float* wt_rgb1; // [256][128]
  const float* b_rgb1;  // [128]
  const float* w_rgb2;  // [3][128]
  const float* b_rgb2;  // [3]
};

constexpr int ST = 64;  // samples per tile
constexpr int SIMT_THREADS = 256;
constexpr size_t SIMT_SMEM = (size_t)(320 * ST * 2 + 64 * ST) * sizeof(float) + 7 * 256 * 8;

enum { EP_FWD_RELU = 0, EP_BWD_MASK = 1, EP_BWD_PLAIN = 2, EP_FWD_RELU_NOMASK = 3 };

// out[o][s] = epilogue( sum_k wt[k*ldw + o] * in[k][s] ),  o in [0,n_out), s in [0,64)
template <int MODE>
__device__ __forceinline__ void simt_layer(const float* __restrict__ wt, int K, int ldw, int n_out, const float* in, float* out,
                                           const float* __restrict__ bias, uint8_t* mask) {
  int ty = threadIdx.x >> 3, tx = threadIdx.x & 7;
  for (int ob = ty * 8; ob < n_out; ob += 256) {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int k = 0; k < K; ++k) {
      float4 w0 = __ldg(reinterpret_cast<const float4*>(wt + (size_t)k * ldw + ob));
      float4 w1 = __ldg(reinterpret_cast<const float4*>(wt + (size_t)k * ldw + ob + 4));
      float4 x0 = *reinterpret_cast<const float4*>(in + k * ST + tx * 8);
      float4 x1 = *reinterpret_cast<const float4*>(in + k * ST + tx * 8 + 4);
      float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int o = ob + i;
      float r[8];
      if (MODE == EP_FWD_RELU || MODE == EP_FWD_RELU_NOMASK) {
        float b = __ldg(bias + o);
        unsigned bits = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = acc[i][j] + b;
          bits |= (a > 0.f ? 1u : 0u) << j;
          r[j] = fmaxf(a, 0.f);
        }
        if (MODE == EP_FWD_RELU) mask[o * 8 + tx] = (uint8_t)bits;
      } else if (MODE == EP_BWD_MASK) {
        unsigned bits = mask[o * 8 + tx];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = ((bits >> j) & 1u) ? acc[i][j] : 0.f;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = acc[i][j];
      }
      *reinterpret_cast<float4*>(out + o * ST + tx * 8) = make_float4(r[0], r[1], r[2], r[3]);
      *reinterpret_cast<float4*>(out + o * ST + tx * 8 + 4) = make_float4(r[4], r[5], r[6], r[7]);
    }
  }
}

// active: (x,y,z,bits(sample)) ; out_a: (sigma, e0, e1, e2) ; out_g: (gx, gy, gz, 0)
__global__ void __launch_bounds__(SIMT_THREADS, 1)
mlp_simt_kernel(SimtWeights W, const float4* __restrict__ active, const unsigned long long* __restrict__ n_active_ptr, int64_t n_active_host,
                float4* __restrict__ out_a, float4* __restrict__ out_g, int density_only) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* A = reinterpret_cast<float*>(smem_raw);  // [320][64]: rows 0..255 activations, 256..319 PE
  float* B = A + 320 * ST;                         // [320][64]
  float* GPE = B + 320 * ST;                       // [64][64]
  uint8_t* masks = reinterpret_cast<uint8_t*>(GPE + 64 * ST);  // [7][256][8]
  int64_t n_active = n_active_ptr ? (int64_t)*n_active_ptr : n_active_host;
  int64_t n_tiles = (n_active + ST - 1) / ST;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    int64_t base = tile * ST;
    __syncthreads();
    // positional encoding (model/dimension_kernel.py:5-35) into A rows 256..319, [feature][sample]
    if (threadIdx.x < ST) {
      int s = threadIdx.x;
      float4 p = base + s < n_active ? active[base + s] : make_float4(0.f, 0.f, 0.f, 0.f);
      float x[3] = {p.x, p.y, p.z};
      float* pe = A + 256 * ST;
#pragma unroll
      for (int c = 0; c < 3; ++c) pe[c * ST + s] = x[c];
      for (int k = 0; k < 10; ++k) {
        float f = (float)(1 << k);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float sn, cs;
          sincosf(x[c] * f, &sn, &cs);
          pe[(3 + 6 * k + c) * ST + s] = sn;
          pe[(6 + 6 * k + c) * ST + s] = cs;
        }
      }
      pe[63 * ST + s] = 0.f;
    }
    __syncthreads();
    // stage 1
    simt_layer<EP_FWD_RELU>(W.wt[0], 64, 256, 256, A + 256 * ST, B, W.bias[0], masks + 0 * 2048);
    __syncthreads();
    simt_layer<EP_FWD_RELU>(W.wt[1], 256, 256, 256, B, A, W.bias[1], masks + 1 * 2048);
    __syncthreads();
    simt_layer<EP_FWD_RELU>(W.wt[2], 256, 256, 256, A, B, W.bias[2], masks + 2 * 2048);
    __syncthreads();
    simt_layer<EP_FWD_RELU>(W.wt[3], 256, 256, 256, B, A, W.bias[3], masks + 3 * 2048);
    __syncthreads();
    // stage 2: input [h (A rows 0..255) | PE (A rows 256..319)] is contiguous in A
    simt_layer<EP_FWD_RELU>(W.wt[4], 320, 256, 256, A, B, W.bias[4], masks + 4 * 2048);
    __syncthreads();
    simt_layer<EP_FWD_RELU>(W.wt[5], 256, 256, 256, B, A, W.bias[5], masks + 5 * 2048);
    __syncthreads();
    simt_layer<EP_FWD_RELU>(W.wt[6], 256, 256, 256, A, B, W.bias[6], masks + 6 * 2048);
    __syncthreads();
    // B rows 0..255 = feature h6.  density head
    if (threadIdx.x < ST) {
      int s = threadIdx.x;
      float acc = 0.f;
      for (int o = 0; o < 256; ++o) acc = fmaf(__ldg(W.w_dens + o), B[o * ST + s], acc);
      GPE[s] = acc + W.b_dens;  // parked in GPE row 0 until the output write
    }
    float e0 = 0.f, e1 = 0.f, e2 = 0.f, sigma = 0.f;
    if (!density_only) {
      // rgb head: ReLU (no-op on h6) -> Linear(256,128) -> ReLU -> Linear(128,3)  (spacenet.py:75-80)
      simt_layer<EP_FWD_RELU_NOMASK>(W.wt_rgb1, 256, 128, 128, B, A, W.b_rgb1, nullptr);
      __syncthreads();
      if (threadIdx.x < ST) {
        int s = threadIdx.x;
        e0 = __ldg(W.b_rgb2 + 0); e1 = __ldg(W.b_rgb2 + 1); e2 = __ldg(W.b_rgb2 + 2);
        for (int o = 0; o < 128; ++o) {
          float h = A[o * ST + s];
          e0 = fmaf(__ldg(W.w_rgb2 + o), h, e0);
          e1 = fmaf(__ldg(W.w_rgb2 + 128 + o), h, e1);
          e2 = fmaf(__ldg(W.w_rgb2 + 256 + o), h, e2);
        }
        sigma = GPE[s];
      }
      __syncthreads();
      // backward: seed d sigma / d a6 = w_dens * mask6, into A rows 0..255
      for (int i = threadIdx.x; i < 256 * 8; i += SIMT_THREADS) {
        int o = i >> 3, tx = i & 7;
        unsigned bits = masks[6 * 2048 + o * 8 + tx];
        float w = __ldg(W.w_dens + o);
#pragma unroll
        for (int j = 0; j < 8; ++j) A[o * ST + tx * 8 + j] = ((bits >> j) & 1u) ? w : 0.f;
      }
      __syncthreads();
      simt_layer<EP_BWD_MASK>(W.w[6], 256, 256, 256, A, B, nullptr, masks + 5 * 2048);
      __syncthreads();
      simt_layer<EP_BWD_MASK>(W.w[5], 256, 256, 256, B, A, nullptr, masks + 4 * 2048);
      __syncthreads();
      // layer 4 backward: 320 outputs = [d/dh3 (masked by m3) | d/dPE]
      simt_layer<EP_BWD_MASK>(W.w[4], 256, 320, 256, A, B, nullptr, masks + 3 * 2048);
      simt_layer<EP_BWD_PLAIN>(W.w[4] + 256, 256, 320, 64, A, GPE, nullptr, nullptr);
      __syncthreads();
      simt_layer<EP_BWD_MASK>(W.w[3], 256, 256, 256, B, A, nullptr, masks + 2 * 2048);
      __syncthreads();
      simt_layer<EP_BWD_MASK>(W.w[2], 256, 256, 256, A, B, nullptr, masks + 1 * 2048);
      __syncthreads();
      simt_layer<EP_BWD_MASK>(W.w[1], 256, 256, 256, B, A, nullptr, masks + 0 * 2048);
      __syncthreads();
      // layer 0 backward (PE columns only) -> B rows 0..63, then add the layer-4 PE gradient
      simt_layer<EP_BWD_PLAIN>(W.w[0], 256, 64, 64, A, B, nullptr, nullptr);
      __syncthreads();
    } else if (threadIdx.x < ST) {
      sigma = GPE[threadIdx.x];
    }
    if (threadIdx.x < ST) {
      int s = threadIdx.x;
      if (base + s < n_active) {
        out_a[base + s] = make_float4(sigma, e0, e1, e2);
        if (!density_only) {
          // sin/cos are still in A rows 256..318 (nothing after layer 4 writes there)
          const float* pe = A + 256 * ST;
          float g[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) g[c] = GPE[c * ST + s] + B[c * ST + s];
          for (int k = 0; k < 10; ++k) {
            float f = (float)(1 << k);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              float sn = pe[(3 + 6 * k + c) * ST + s], cs = pe[(6 + 6 * k + c) * ST + s];
              float gs = GPE[(3 + 6 * k + c) * ST + s] + B[(3 + 6 * k + c) * ST + s];
              float gc = GPE[(6 + 6 * k + c) * ST + s] + B[(6 + 6 * k + c) * ST + s];
              g[c] += (gs * cs - gc * sn) * f;
            }
          }
          out_g[base + s] = make_float4(g[0], g[1], g[2], 0.f);
        }
      }
    }
  }
}

}  // namespace dsn
