// simt_gemm.cuh — FFMA (fp32-accumulate, no tensor core) tiled GEMM used by the MILB200_F32 path
// (<=1e-5 parity needs full fp32 products) and for shapes the tcgen05 kernels do not cover.
// C[m,n] = sum_k A(m,k) * B(n,k) with either operand stored k-contiguous or m/n-contiguous.
#pragma once

#include "common.cuh"

namespace milb200 {
namespace simt {

constexpr int TM = 64, TN = 64, TK = 16, THREADS = 256;

// Epilogue functors receive (row, col, value, split) for every in-range element.
template <typename TO>
struct EpiStore {  // out = act(v + bias[col]) (+ rank-1 term attn[row] * dM[bag(row), col])
  TO* out;
  int64_t ldo;
  const float* bias;
  int act;
  const float* attn;       // optional: per-row weight
  const float* dM;         // optional: [B, N]
  const int32_t* offsets;  // CSR offsets for the row -> bag lookup
  int B;
  int64_t row_base;        // global row of local row 0 (attn / offsets are indexed globally)
  __device__ __forceinline__ void operator()(int64_t m, int n, float v, int /*split*/) const {
    if (bias) v += __ldg(bias + n);
    if (act == MILB200_ACT_TANH) v = tanhf(v);
    else if (act == MILB200_ACT_RELU) v = fmaxf(v, 0.f);
    else if (act == MILB200_ACT_SIGMOID) v = sigmoid_precise(v);
    if (attn) {
      int b = find_bag(offsets, B, m + row_base);
      v = fmaf(__ldg(attn + m + row_base), __ldg(dM + static_cast<int64_t>(b) * ldo + n), v);
    }
    out[m * ldo + n] = from_f32<TO>(v);
  }
};

struct EpiPartial {  // split-K partials: part[split][m][n]
  float* part;
  int64_t M;
  int N;
  __device__ __forceinline__ void operator()(int64_t m, int n, float v, int split) const {
    part[(static_cast<int64_t>(split) * M + m) * N + n] = v;
  }
};

template <typename TA, typename TB, bool A_KCONTIG, bool B_KCONTIG, class Epi>
__global__ void __launch_bounds__(THREADS)
k_gemm(const TA* __restrict__ A, int64_t lda, const TB* __restrict__ B, int64_t ldb, int64_t M, int N,
       int64_t K, int64_t k_per_split, Epi epi) {
  __shared__ __align__(16) float As[TK][TM + 4];
  __shared__ __align__(16) float Bs[TK][TN + 4];
  const int t = threadIdx.x;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * TM;
  const int n0 = blockIdx.y * TN;
  const int split = blockIdx.z;
  const int64_t kbeg = split * k_per_split;
  const int64_t kend = (kbeg + k_per_split < K) ? kbeg + k_per_split : K;
  const int ty = t / 16, tx = t % 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kbeg; k0 < kend; k0 += TK) {
    // ---- stage A tile ----
    if (A_KCONTIG) {
      int mi = t / 4, kk = (t % 4) * 4;
      int64_t m = m0 + mi;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int64_t k = k0 + kk + j;
        As[kk + j][mi] = (m < M && k < kend) ? to_f32<TA>(A[m * lda + k]) : 0.f;
      }
    } else {
      int kk = t / 16, mi = (t % 16) * 4;
      int64_t k = k0 + kk;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int64_t m = m0 + mi + j;
        As[kk][mi + j] = (m < M && k < kend) ? to_f32<TA>(A[k * lda + m]) : 0.f;
      }
    }
    // ---- stage B tile ----
    if (B_KCONTIG) {
      int ni = t / 4, kk = (t % 4) * 4;
      int n = n0 + ni;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int64_t k = k0 + kk + j;
        Bs[kk + j][ni] = (n < N && k < kend) ? to_f32<TB>(B[static_cast<int64_t>(n) * ldb + k]) : 0.f;
      }
    } else {
      int kk = t / 16, ni = (t % 16) * 4;
      int64_t k = k0 + kk;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int n = n0 + ni + j;
        Bs[kk][ni + j] = (n < N && k < kend) ? to_f32<TB>(B[k * ldb + n]) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < N) epi(m, n, acc[i][j], split);
    }
  }
}

template <typename TA, typename TB, bool A_KCONTIG, bool B_KCONTIG, class Epi>
inline int launch(const TA* A, int64_t lda, const TB* B, int64_t ldb, int64_t M, int N, int64_t K,
                  int splits, Epi epi, cudaStream_t st) {
  if (M <= 0 || N <= 0) return MILB200_OK;
  int64_t kps = (K + splits - 1) / splits;
  kps = (kps + TK - 1) / TK * TK;
  dim3 grid(static_cast<unsigned>((M + TM - 1) / TM), static_cast<unsigned>((N + TN - 1) / TN),
            static_cast<unsigned>(splits));
  k_gemm<TA, TB, A_KCONTIG, B_KCONTIG, Epi><<<grid, THREADS, 0, st>>>(A, lda, B, ldb, M, N, K, kps, epi);
  MIL_LAUNCH_CHECK();
  return MILB200_OK;
}

}  // namespace simt

// out[i] = (accumulate ? out[i] : 0) + sum_s part[s][i]   (fixed order => deterministic)
int splitk_reduce(const float* part, int splits, int64_t n, float* out, int accumulate, cudaStream_t st);

}  // namespace milb200
