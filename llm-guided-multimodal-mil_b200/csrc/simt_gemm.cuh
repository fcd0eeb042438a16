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

// ---- 128 x 128 x 8 tiles, 8 x 8 outputs per thread, register-prefetched double buffering ---------------------------
// Same contract as k_gemm; used when both output extents reach a full tile.  Per k step a thread issues four 128-bit
// shared loads for 64 FMAs (the 64 x 64 kernel: two for 16), and the global loads of tile i+1 are in flight while tile i
// is multiplied.
constexpr int BM2 = 128, BN2 = 128, BK2 = 8;

// four consecutive elements of `p` (guarded element-wise by `ok`), one 128-/64-bit load when the address allows it
template <typename T>
__device__ __forceinline__ void load4(const T* p, bool ok0, bool ok1, bool ok2, bool ok3, bool vec, float (&v)[4]) {
  if (vec && ok3) {   // ok3 implies ok0..ok2 (indices grow)
    if (sizeof(T) == 4) {
      float4 q = __ldg(reinterpret_cast<const float4*>(p));
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
      uint2 q = __ldg(reinterpret_cast<const uint2*>(p));
      v[0] = __uint_as_float(q.x << 16); v[1] = __uint_as_float(q.x & 0xffff0000u);
      v[2] = __uint_as_float(q.y << 16); v[3] = __uint_as_float(q.y & 0xffff0000u);
    }
    return;
  }
  v[0] = ok0 ? to_f32<T>(p[0]) : 0.f;
  v[1] = ok1 ? to_f32<T>(p[1]) : 0.f;
  v[2] = ok2 ? to_f32<T>(p[2]) : 0.f;
  v[3] = ok3 ? to_f32<T>(p[3]) : 0.f;
}

// stage registers of one operand tile [128 (mn) x 8 (k)]: KCONTIG: thread -> (mn = t / 2, k = 4 (t % 2) ..+3);
// otherwise thread -> (k = t / 32, mn = 4 (t % 32) ..+3)
template <typename T, bool KCONTIG>
__device__ __forceinline__ void fetch_tile(const T* __restrict__ P, int64_t ld, int64_t mn0, int64_t MN, int64_t k0,
                                           int64_t kend, int t, bool vec, float (&v)[4]) {
  if (KCONTIG) {
    const int64_t mn = mn0 + (t >> 1), k = k0 + ((t & 1) << 2);
    const bool in = mn < MN;
    load4<T>(P + mn * ld + k, in && k < kend, in && k + 1 < kend, in && k + 2 < kend, in && k + 3 < kend, vec, v);
  } else {
    const int64_t k = k0 + (t >> 5), mn = mn0 + ((t & 31) << 2);
    const bool in = k < kend;
    load4<T>(P + k * ld + mn, in && mn < MN, in && mn + 1 < MN, in && mn + 2 < MN, in && mn + 3 < MN, vec, v);
  }
}
template <bool KCONTIG>
__device__ __forceinline__ void stash_tile(float (*S)[BM2 + 4], int t, const float (&v)[4]) {
  if (KCONTIG) {
    const int mn = t >> 1, k = (t & 1) << 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) S[k + j][mn] = v[j];
  } else {
    *reinterpret_cast<float4*>(&S[t >> 5][(t & 31) << 2]) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

template <typename TA, typename TB, bool A_KCONTIG, bool B_KCONTIG, class Epi>
__global__ void __launch_bounds__(THREADS, 2)
k_gemm128(const TA* __restrict__ A, int64_t lda, const TB* __restrict__ B, int64_t ldb, int64_t M, int N, int64_t K,
          int64_t k_per_split, Epi epi) {
  __shared__ __align__(16) float As[2][BK2][BM2 + 4];
  __shared__ __align__(16) float Bs[2][BK2][BN2 + 4];
  const int t = threadIdx.x;
  const int64_t m0 = static_cast<int64_t>(blockIdx.x) * BM2;
  const int n0 = blockIdx.y * BN2;
  const int split = blockIdx.z;
  const int64_t kbeg = split * k_per_split;
  const int64_t kend = (kbeg + k_per_split < K) ? kbeg + k_per_split : K;
  const int ty = t >> 4, tx = t & 15;
  // vector loads need 16-byte (fp32) / 8-byte (bf16) aligned groups of four: base pointer, leading dimension, k origin
  const bool vecA = (reinterpret_cast<uintptr_t>(A) % (4 * sizeof(TA)) == 0) && lda % 4 == 0 && (A_KCONTIG ? kbeg % 4 == 0 : true);
  const bool vecB = (reinterpret_cast<uintptr_t>(B) % (4 * sizeof(TB)) == 0) && ldb % 4 == 0 && (B_KCONTIG ? kbeg % 4 == 0 : true);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float ra[4], rb[4];
  if (kbeg < kend) {
    fetch_tile<TA, A_KCONTIG>(A, lda, m0, M, kbeg, kend, t, vecA, ra);
    fetch_tile<TB, B_KCONTIG>(B, ldb, n0, N, kbeg, kend, t, vecB, rb);
    stash_tile<A_KCONTIG>(As[0], t, ra);
    stash_tile<B_KCONTIG>(Bs[0], t, rb);
  }
  __syncthreads();
  int buf = 0;
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK2) {
    const bool more = k0 + BK2 < kend;
    if (more) {
      fetch_tile<TA, A_KCONTIG>(A, lda, m0, M, k0 + BK2, kend, t, vecA, ra);
      fetch_tile<TB, B_KCONTIG>(B, ldb, n0, N, k0 + BK2, kend, t, vecB, rb);
    }
#pragma unroll
    for (int kk = 0; kk < BK2; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (more) {
      stash_tile<A_KCONTIG>(As[buf ^ 1], t, ra);
      stash_tile<B_KCONTIG>(Bs[buf ^ 1], t, rb);
    }
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
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
  if (M >= BM2 && N >= BN2) {
    dim3 grid2(static_cast<unsigned>((M + BM2 - 1) / BM2), static_cast<unsigned>((N + BN2 - 1) / BN2),
               static_cast<unsigned>(splits));
    k_gemm128<TA, TB, A_KCONTIG, B_KCONTIG, Epi><<<grid2, THREADS, 0, st>>>(A, lda, B, ldb, M, N, K, kps, epi);
    MIL_LAUNCH_CHECK();
    return MILB200_OK;
  }
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
