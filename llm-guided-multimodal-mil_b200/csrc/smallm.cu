// smallm.cu — dense layers on a HANDFUL of rows (m <= 16): the text-token side of the two-way transformer
// (T = 1 or 10 tokens: q/k/v/out projections, the 512->2048->512 MLP; model/sam/transformer.py:283-300,
// model/sam/common.py:26), fc_CI2CT / fc_CI2Pth (model/aggregator.py:44,66) and the classification heads
// (aggregator.py:128-131, aggregator_clip.py:63-75, aggregator_wMask.py:67-70).
//
// These are weight-streaming problems (0.5-8 MB of weights against a few KB of activations): a 128-row tensor-core
// tile or a 64x64 FFMA tile leaves all but one or two SMs idle and serialises the K loop inside one CTA.  Here every
// weight element is read exactly once, by as many CTAs as the shape allows, with 128-bit loads; fp32 accumulation.
//   forward   one warp per output column: lanes stride the K dimension, activations staged in shared memory
//   dX        g = dY * act'(Y) staged in shared memory; a CTA owns 64 input columns, its 8 warps split the
//             output rows of W and fold their partials in shared memory (fixed order)
//   dW, db    one thread per 4 weights: an m-term outer-product sum, streamed out (optionally accumulated)
#include <algorithm>

#include "common.cuh"

namespace milb200 {

constexpr int SMALLM_MAX = 16;
constexpr int SMALLM_MAXDIM = 2048;  // k (forward) and n (dX) must fit the shared-memory staging: 16 x 2048 fp32 = 128 KB

template <typename T>
__device__ __forceinline__ float act_grad(float g, float y, int act) {
  if (act == MILB200_ACT_TANH) return g * (1.f - y * y);
  if (act == MILB200_ACT_RELU) return y > 0.f ? g : 0.f;
  if (act == MILB200_ACT_SIGMOID) return g * y * (1.f - y);
  return g;
}
__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == MILB200_ACT_TANH) return tanhf(v);
  if (act == MILB200_ACT_RELU) return fmaxf(v, 0.f);
  if (act == MILB200_ACT_SIGMOID) return sigmoid_precise(v);
  return v;
}

// Y[i, n] = act(sum_k (X[i,k] + add[i,k]) W[n,k] + bias[n]);  MT = padded row count (power of two >= m)
template <typename T, int MT>
__global__ void __launch_bounds__(256)
k_smallm_fwd(const T* __restrict__ X, const T* __restrict__ add, const T* __restrict__ W, const float* __restrict__ bias,
             T* __restrict__ Y, int m, int n, int k, int act, int cols_per_cta) {
  constexpr int VN = Vec16<T>::N;
  extern __shared__ __align__(16) float xs[];  // [MT][k]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int i = t; i < MT * k; i += 256) {
    const int r = i / k;
    float v = 0.f;
    if (r < m) v = to_f32<T>(X[i]) + (add ? to_f32<T>(add[i]) : 0.f);
    xs[i] = v;
  }
  __syncthreads();
  const int c0 = blockIdx.x * cols_per_cta;
  const int c1 = min(n, c0 + cols_per_cta);
  const int nvec = k / VN;
  for (int c = c0 + warp; c < c1; c += 8) {
    float acc[MT];
#pragma unroll
    for (int i = 0; i < MT; ++i) acc[i] = 0.f;
    const uint4* wrow = reinterpret_cast<const uint4*>(W + static_cast<int64_t>(c) * k);
    for (int v = lane; v < nvec; v += 32) {
      float w[VN];
      Vec16<T>::unpack(__ldg(wrow + v), w);
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const float4* xr = reinterpret_cast<const float4*>(xs + i * k + v * VN);
#pragma unroll
        for (int q = 0; q < VN / 4; ++q) {
          const float4 x4 = xr[q];
          acc[i] = fmaf(w[4 * q], x4.x, acc[i]);
          acc[i] = fmaf(w[4 * q + 1], x4.y, acc[i]);
          acc[i] = fmaf(w[4 * q + 2], x4.z, acc[i]);
          acc[i] = fmaf(w[4 * q + 3], x4.w, acc[i]);
        }
      }
    }
    const float b = bias ? __ldg(bias + c) : 0.f;
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      const float s = warp_sum(acc[i]);
      if (lane == 0 && i < m) Y[static_cast<int64_t>(i) * n + c] = from_f32<T>(act_apply(s + b, act));
    }
  }
}

// dX[i, kk] = sum_n g[i,n] W[n,kk], g = dY * act'(Y).  CTA (x, y) = 64 input columns (2 per lane) x the y-th slice of the
// output rows n; its 8 warps split the slice and fold in shared memory; slices are folded by k_smallm_dx_reduce in fixed
// order.  (One CTA per 64 columns alone leaves 140 SMs idle and serialises 0.25-1 MB of weight reads behind 16 KB in flight.)
constexpr int SMALLM_NSPLIT_MAX = 16;
template <typename T, int MT>
__global__ void __launch_bounds__(256)
k_smallm_dx(const T* __restrict__ W, const T* __restrict__ Y, const T* __restrict__ dY, float* __restrict__ part, int m, int n,
            int k, int act, int rows_per_split) {
  extern __shared__ __align__(16) float sm[];  // g[MT][rows_per_split] | red[8][MT][64]
  const int n0 = blockIdx.y * rows_per_split;
  const int n1 = min(n, n0 + rows_per_split);
  const int ns = n1 - n0;
  float* gs = sm;
  float* red = sm + MT * rows_per_split;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int i = t; i < MT * ns; i += 256) {
    const int r = i / ns, c = i % ns;
    float v = 0.f;
    if (r < m) {
      const int64_t yo = static_cast<int64_t>(r) * n + n0 + c;
      v = act_grad<T>(to_f32<T>(dY[yo]), act != MILB200_ACT_NONE ? to_f32<T>(Y[yo]) : 0.f, act);
    }
    gs[r * rows_per_split + c] = v;
  }
  __syncthreads();
  const int kk = blockIdx.x * 64 + lane * 2;
  float acc[MT][2];
#pragma unroll
  for (int i = 0; i < MT; ++i) acc[i][0] = acc[i][1] = 0.f;
  if (kk < k) {
    constexpr int UNR = 8;
    auto ldw = [&](int r, float& w0, float& w1) {
      if (sizeof(T) == 4) {
        const float2 w2 = __ldg(reinterpret_cast<const float2*>(W + static_cast<int64_t>(n0 + r) * k + kk));
        w0 = w2.x; w1 = w2.y;
      } else {
        const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(W + static_cast<int64_t>(n0 + r) * k + kk));
        w0 = bf16lo(u); w1 = bf16hi(u);
      }
    };
    int r = warp;
    for (; r + 8 * (UNR - 1) < ns; r += 8 * UNR) {
      float w0[UNR], w1[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) ldw(r + 8 * u, w0[u], w1[u]);
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          const float g = gs[i * rows_per_split + r + 8 * u];
          acc[i][0] = fmaf(g, w0[u], acc[i][0]);
          acc[i][1] = fmaf(g, w1[u], acc[i][1]);
        }
      }
    }
    for (; r < ns; r += 8) {
      float w0, w1;
      ldw(r, w0, w1);
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const float g = gs[i * rows_per_split + r];
        acc[i][0] = fmaf(g, w0, acc[i][0]);
        acc[i][1] = fmaf(g, w1, acc[i][1]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    red[(warp * MT + i) * 64 + lane * 2] = acc[i][0];
    red[(warp * MT + i) * 64 + lane * 2 + 1] = acc[i][1];
  }
  __syncthreads();
  float* out = part + static_cast<int64_t>(blockIdx.y) * m * k;
  for (int e = t; e < m * 64; e += 256) {
    const int i = e / 64, c = e % 64;
    const int col = blockIdx.x * 64 + c;
    if (col >= k) continue;
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[(w * MT + i) * 64 + c];
    out[static_cast<int64_t>(i) * k + col] = a;
  }
}
template <typename T>
__global__ void k_smallm_dx_reduce(const float* __restrict__ part, int splits, int64_t mk, T* __restrict__ dX) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= mk) return;
  float a = 0.f;
  for (int s = 0; s < splits; ++s) a += part[static_cast<int64_t>(s) * mk + i];
  dX[i] = from_f32<T>(a);
}

// dW[n, kk..kk+3] (+)= sum_i g[i,n] (X[i,kk..] + add[i,kk..]);  dbias[n] (+)= sum_i g[i,n]
template <typename T>
__global__ void __launch_bounds__(256)
k_smallm_dw(const T* __restrict__ X, const T* __restrict__ add, const T* __restrict__ Y, const T* __restrict__ dY,
            float* __restrict__ dW, float* __restrict__ dbias, int m, int n, int k, int act, int accumulate) {
  const int kv = k / 4;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(n) * kv) return;
  const int r = static_cast<int>(idx / kv), c = static_cast<int>(idx % kv) * 4;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, gb = 0.f;
  for (int i = 0; i < m; ++i) {
    const int64_t yo = static_cast<int64_t>(i) * n + r;
    const float g = act_grad<T>(to_f32<T>(dY[yo]), act != MILB200_ACT_NONE ? to_f32<T>(Y[yo]) : 0.f, act);
    const T* xr = X + static_cast<int64_t>(i) * k + c;
    float x0 = to_f32<T>(xr[0]), x1 = to_f32<T>(xr[1]), x2 = to_f32<T>(xr[2]), x3 = to_f32<T>(xr[3]);
    if (add) {
      const T* ar = add + static_cast<int64_t>(i) * k + c;
      x0 += to_f32<T>(ar[0]); x1 += to_f32<T>(ar[1]); x2 += to_f32<T>(ar[2]); x3 += to_f32<T>(ar[3]);
    }
    a0 = fmaf(g, x0, a0); a1 = fmaf(g, x1, a1); a2 = fmaf(g, x2, a2); a3 = fmaf(g, x3, a3);
    gb += g;
  }
  float4* dst = reinterpret_cast<float4*>(dW + static_cast<int64_t>(r) * k + c);
  float4 o = make_float4(a0, a1, a2, a3);
  if (accumulate) { const float4 p = *dst; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
  *dst = o;
  if (c == 0 && dbias) dbias[r] = accumulate ? dbias[r] + gb : gb;
}

static int pad_rows(int64_t m) { return m <= 1 ? 1 : (m <= 2 ? 2 : (m <= 4 ? 4 : (m <= 8 ? 8 : 16))); }

bool smallm_ok(int64_t m, int n, int k, int dtype) {
  const int vn = dtype == MILB200_BF16 ? 8 : 4;
  return m >= 1 && m <= SMALLM_MAX && k % vn == 0 && k <= SMALLM_MAXDIM && n <= SMALLM_MAXDIM && n >= 1 && k % 4 == 0;
}

template <typename T>
static int smallm_fwd_t(const T* X, const T* add, const T* W, const float* bias, T* Y, int m, int n, int k, int act,
                        cudaStream_t st) {
  const int mt = pad_rows(m);
  // enough CTAs to cover the SMs, at least one column per warp
  int cols = std::max(8, (n + 2 * sm_count() - 1) / (2 * sm_count()));
  cols = (cols + 7) / 8 * 8;
  const unsigned grid = static_cast<unsigned>((n + cols - 1) / cols);
  const size_t smem = sizeof(float) * static_cast<size_t>(mt) * k;
#define MIL_SMALLM_FWD(MT)                                                                                      \
  {                                                                                                             \
    auto kern = k_smallm_fwd<T, MT>;                                                                            \
    if (smem > 48 * 1024) MIL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, 256, smem, st>>>(X, add, W, bias, Y, m, n, k, act, cols);                                      \
  }
  switch (mt) {
    case 1: MIL_SMALLM_FWD(1) break;
    case 2: MIL_SMALLM_FWD(2) break;
    case 4: MIL_SMALLM_FWD(4) break;
    case 8: MIL_SMALLM_FWD(8) break;
    default: MIL_SMALLM_FWD(16) break;
  }
#undef MIL_SMALLM_FWD
  MIL_LAUNCH_CHECK();
  return MILB200_OK;
}

size_t smallm_ws_bytes(int64_t m, int k) { return sizeof(float) * SMALLM_NSPLIT_MAX * static_cast<size_t>(m) * k; }

template <typename T>
static int smallm_bwd_t(const T* X, const T* add, const T* W, const T* Y, const T* dY, T* dX, float* dW, float* dbias, int m,
                        int n, int k, int act, int accumulate, float* ws, cudaStream_t st) {
  const int mt = pad_rows(m);
  if (dX) {
    // enough (column block, row slice) CTAs to cover the SMs; at least 8 rows of W per warp
    const int colblocks = (k + 63) / 64;
    int splits = std::min(SMALLM_NSPLIT_MAX, std::max(1, (2 * sm_count() + colblocks - 1) / colblocks));
    splits = std::min(splits, std::max(1, n / 64));
    const int rps = (n + splits - 1) / splits;
    splits = (n + rps - 1) / rps;
    const dim3 grid(static_cast<unsigned>(colblocks), static_cast<unsigned>(splits));
    const size_t smem = sizeof(float) * (static_cast<size_t>(mt) * rps + 8 * static_cast<size_t>(mt) * 64);
#define MIL_SMALLM_DX(MT)                                                                                       \
  {                                                                                                             \
    auto kern = k_smallm_dx<T, MT>;                                                                             \
    if (smem > 48 * 1024) MIL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, 256, smem, st>>>(W, Y, dY, ws, m, n, k, act, rps);                                             \
  }
    switch (mt) {
      case 1: MIL_SMALLM_DX(1) break;
      case 2: MIL_SMALLM_DX(2) break;
      case 4: MIL_SMALLM_DX(4) break;
      case 8: MIL_SMALLM_DX(8) break;
      default: MIL_SMALLM_DX(16) break;
    }
#undef MIL_SMALLM_DX
    MIL_LAUNCH_CHECK();
    const int64_t mk = static_cast<int64_t>(m) * k;
    k_smallm_dx_reduce<T><<<static_cast<unsigned>((mk + 255) / 256), 256, 0, st>>>(ws, splits, mk, dX);
    MIL_LAUNCH_CHECK();
  }
  if (dW) {
    const int64_t threads = static_cast<int64_t>(n) * (k / 4);
    k_smallm_dw<T><<<static_cast<unsigned>((threads + 255) / 256), 256, 0, st>>>(X, add, Y, dY, dW, dbias, m, n, k, act,
                                                                                accumulate);
    MIL_LAUNCH_CHECK();
  }
  return MILB200_OK;
}

int smallm_fwd(const void* X, const void* add, const void* W, const float* bias, void* Y, int64_t m, int n, int k, int act,
               int dtype, cudaStream_t st) {
  if (dtype == MILB200_BF16)
    return smallm_fwd_t<__nv_bfloat16>((const __nv_bfloat16*)X, (const __nv_bfloat16*)add, (const __nv_bfloat16*)W, bias,
                                       (__nv_bfloat16*)Y, (int)m, n, k, act, st);
  return smallm_fwd_t<float>((const float*)X, (const float*)add, (const float*)W, bias, (float*)Y, (int)m, n, k, act, st);
}

int smallm_bwd(const void* X, const void* add, const void* W, const void* Y, const void* dY, void* dX, float* dW,
               float* dbias, int64_t m, int n, int k, int act, int dtype, int accumulate, float* ws, cudaStream_t st) {
  if (dtype == MILB200_BF16)
    return smallm_bwd_t<__nv_bfloat16>((const __nv_bfloat16*)X, (const __nv_bfloat16*)add, (const __nv_bfloat16*)W,
                                       (const __nv_bfloat16*)Y, (const __nv_bfloat16*)dY, (__nv_bfloat16*)dX, dW, dbias,
                                       (int)m, n, k, act, accumulate, ws, st);
  return smallm_bwd_t<float>((const float*)X, (const float*)add, (const float*)W, (const float*)Y, (const float*)dY,
                             (float*)dX, dW, dbias, (int)m, n, k, act, accumulate, ws, st);
}

}  // namespace milb200
