// xfusion.cu — the image side of the two-way cross-modal attention (model/sam/transformer.py:278-309,418-450 as called
// by model/aggregator.py:160,168) with the key / value projections FOLDED INTO THE TOKEN SIDE.
//
// The reference projects every image token: K = (keys + pe) Wk^T + bk, V = keys Wv^T + bv  ([N, 256] each, two
// N x 512 x 256 GEMMs per attention, six per TwoWayTransformer call), then lets T <= 16 text tokens attend to them.
// With so few queries the projections can move to the other side of the product exactly:
//     score_h[t, n] = q_h[t] . K_h[n] / sqrt(c) = (keys[n] + pe[n]) . (Wk_h^T q_h[t]) / sqrt(c) + (q_h[t] . bk_h)/sqrt(c)
//     out_h[t]      = sum_n a_h[t, n] V_h[n]    = Wv_h (sum_n a_h[t, n] keys[n]) + bv_h          (sum_n a = 1)
// The bias term of the score is constant over n and cancels in the softmax over n (the reference gives k_proj.bias a
// gradient that is zero up to float noise for the same reason).  So per attention the image side needs
//     U[(t, h), :] = Wk_h^T q_h[t]                       [T*8, 512]   token side, 32 x 512 per head  (headdiag_expand)
//     S = scale (keys + pe) U^T, a = softmax_n(S), Pool = a^T keys    one pass over keys              (t2i_fwd)
//     o[t, h*32 + c] = Wv[h*32 + c, :] . Pool[(t, h), :] + bv          token side                     (headdiag_contract)
// i.e. 2 x 8T x 512 flop per image token instead of 2 x 2 x 256 x 512, no [N, 256] K / V tensors, and the pass is
// bandwidth-bound on ONE read of keys (+ pe): the same class of kernel as the gated-attention pool (pool.cu) with
// 8T "heads" whose scores are linear in the row.  The backward mirrors it (t2i_bwd: one pass, dkeys and dU).
// With a single text token the image -> token attention is the broadcast of one row (softmax over one key = 1,
// SURVEY F10): ln_seg_* is LayerNorm(keys + row[segment]) over segments, each segment with its own row.
//
// Segments: the CT bag and the pathology bag of a patient (and of several patients) run through the SAME transformer
// weights (TwoWayTransformer_Both, aggregator.py:160,168), so they are processed as segments of one launch.
//
// All kernels: one warp per row of 512 values (16 per lane), fp32 math, `TK` storage (float or bf16) for the key
// stream, deterministic fixed-order reductions (per-item partials + a merge kernel), no atomics.
#include <cfloat>

#include "xfusion.cuh"

namespace milb200 {
namespace xf {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr float SCALE = 0.17677669529663687f;   // 1 / sqrt(CH), CH = 32 (transformer.py:441)

// exp for softmax weights: full-precision for the fp32 (<= 1e-5 parity) path, ex2.approx for bf16 storage
template <typename TK> __device__ __forceinline__ float xexp(float x);
template <> __device__ __forceinline__ float xexp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ float xexp<__nv_bfloat16>(float x) { return __expf(x); }

// Sums of 8 per-lane values over the warp, every lane ends with all 8 sums: a halving butterfly (4 + 2 + 1 exchanges), two
// plain steps and 8 broadcasts = 17 shuffles instead of 8 x 5.  After the butterfly lane l holds the sum of value
// j = 4 bit4(l) + 2 bit3(l) + bit2(l).  The association is fixed, so results are reproducible.
__device__ __forceinline__ void warp_sum8(float* v, int lane) {
  const unsigned full = 0xffffffffu;
  const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
  float a[4], b[2];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float send = h4 ? v[k] : v[k + 4], keep = h4 ? v[k + 4] : v[k];
    a[k] = keep + __shfl_xor_sync(full, send, 16);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float send = h3 ? a[k] : a[k + 2], keep = h3 ? a[k + 2] : a[k];
    b[k] = keep + __shfl_xor_sync(full, send, 8);
  }
  float c = (h2 ? b[1] : b[0]) + __shfl_xor_sync(full, h2 ? b[0] : b[1], 4);
  c += __shfl_xor_sync(full, c, 2);
  c += __shfl_xor_sync(full, c, 1);
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __shfl_sync(full, c, ((j & 4) ? 16 : 0) | ((j & 2) ? 8 : 0) | ((j & 1) ? 4 : 0));
}

// ---- a 512-wide row spread over a warp ------------------------------------------------------------------------------
// Lane l owns the 16-byte vectors l, l+32, ... of the row: value idx = i*VN + e  <->  column (l + 32 i) * VN + e.
template <typename TK> struct Row {
  static constexpr int VN = Vec16<TK>::N;   // 4 (fp32) or 8 (bf16)
  static constexpr int NV = 16 / VN;        // 4 or 2 vectors per lane
  static constexpr int HV = VN / 4;         // float4 pieces per vector
  __device__ static __forceinline__ void load(const TK* row, int lane, float* v) {
    const uint4* p = reinterpret_cast<const uint4*>(row);
#pragma unroll
    for (int i = 0; i < NV; ++i) Vec16<TK>::unpack(ldg_stream(p + lane + 32 * i), v + i * VN);
  }
  __device__ static __forceinline__ void load_cached(const TK* row, int lane, float* v) {
    const uint4* p = reinterpret_cast<const uint4*>(row);
#pragma unroll
    for (int i = 0; i < NV; ++i) Vec16<TK>::unpack(p[lane + 32 * i], v + i * VN);
  }
  __device__ static __forceinline__ void store(TK* row, int lane, const float* v) {
    uint4* p = reinterpret_cast<uint4*>(row);
#pragma unroll
    for (int i = 0; i < NV; ++i) p[lane + 32 * i] = Vec16<TK>::pack(v + i * VN);
  }
  // shared-memory image of an fp32 row in the lane order above, split into float4 planes so that a warp's 128-bit reads
  // are conflict-free for both storage types: column c -> index ((e/4) * (E/VN) + c/VN) * 4 + e%4, e = c % VN
  __device__ static __forceinline__ int sm_index(int c) {
    const int v = c / VN, e = c % VN;
    return ((e >> 2) * (E / VN) + v) * 4 + (e & 3);
  }
  __device__ static __forceinline__ void sm_read(const float* sm_row, int lane, float* v) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int h = 0; h < HV; ++h) {
        const float4 f = *reinterpret_cast<const float4*>(sm_row + (h * (E / VN) + lane + 32 * i) * 4);
        v[i * VN + h * 4 + 0] = f.x; v[i * VN + h * 4 + 1] = f.y; v[i * VN + h * 4 + 2] = f.z; v[i * VN + h * 4 + 3] = f.w;
      }
  }
  __device__ static __forceinline__ int col(int lane, int idx) { return (lane + 32 * (idx / VN)) * VN + idx % VN; }
};

// fp32 global row -> registers in the lane order of TK (for parameters / small fp32 operands)
template <typename TK>
__device__ __forceinline__ void load_f32_row(const float* row, int lane, float* v) {
  constexpr int VN = Row<TK>::VN, NV = Row<TK>::NV, HV = Row<TK>::HV;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int h = 0; h < HV; ++h) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(row + (lane + 32 * i) * VN + h * 4));
      v[i * VN + h * 4 + 0] = f.x; v[i * VN + h * 4 + 1] = f.y; v[i * VN + h * 4 + 2] = f.z; v[i * VN + h * 4 + 3] = f.w;
    }
}

template <typename TK>
__device__ __forceinline__ void store_f32_row(float* row, int lane, const float* v) {
  constexpr int VN = Row<TK>::VN, NV = Row<TK>::NV, HV = Row<TK>::HV;
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int h = 0; h < HV; ++h)
      *reinterpret_cast<float4*>(row + (lane + 32 * i) * VN + h * 4) =
          make_float4(v[i * VN + h * 4 + 0], v[i * VN + h * 4 + 1], v[i * VN + h * 4 + 2], v[i * VN + h * 4 + 3]);
}
// rows of storage type T handled in the lane order of the mapping type TM (TM = bf16 whenever either side of a kernel
// stores bf16, so that one lane owns the same 16 columns of every operand)
template <typename T, typename TM>
__device__ __forceinline__ void rload(const T* row, int lane, float* v) {
  if constexpr (sizeof(T) == sizeof(TM)) Row<T>::load(row, lane, v);
  else load_f32_row<TM>(reinterpret_cast<const float*>(row), lane, v);
}
template <typename T, typename TM>
__device__ __forceinline__ void rstore(T* row, int lane, const float* v) {
  if constexpr (sizeof(T) == sizeof(TM)) Row<T>::store(row, lane, v);
  else store_f32_row<TM>(reinterpret_cast<float*>(row), lane, v);
}

__device__ __forceinline__ int find_seg(const Segs& sg, int item) {
  int s = 0;
  while (s + 1 < sg.n && item >= sg.item0[s + 1]) ++s;
  return s;
}

// stage `rows` fp32 rows of E values (global, row stride E) into the plane layout
template <typename TK>
__device__ __forceinline__ void stage_rows(float* sm, const float* g, int rows) {
  for (int i = threadIdx.x; i < rows * (E / 4); i += THREADS) {
    const int r = i / (E / 4), c = (i % (E / 4)) * 4;
    const float4 f = __ldg(reinterpret_cast<const float4*>(g + static_cast<int64_t>(r) * E + c));
    *reinterpret_cast<float4*>(sm + r * E + Row<TK>::sm_index(c)) = f;     // c % 4 == 0: one float4 stays one float4
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// token -> image attention, forward
// ---------------------------------------------------------------------------------------------------------------------
// grid (items, T); CTA = one item (<= rows_per_item rows of one segment) x the 8 heads of token t.
// part_acc[(item*T + t)*8 + h][E], part_ml[(item*T + t)*8 + h] = (max, sum exp) of the item.
template <typename TK>
__global__ void __launch_bounds__(THREADS)
k_t2i_fwd(const TK* __restrict__ K, const float* __restrict__ PE, const float* __restrict__ U, const Segs sg, const int bag_layout,
          float* __restrict__ S, float* __restrict__ part_acc, float2* __restrict__ part_ml) {
  __shared__ __align__(16) float Us[H * E];
  __shared__ __align__(16) float red[WARPS * E];
  __shared__ float wm[WARPS][H], wl[WARPS][H];
  const int item = blockIdx.x, t = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = find_seg(sg, item);
  const int i0 = (item - sg.item0[s]) * sg.rows_per_item;
  const int i1 = min(sg.len[s], i0 + sg.rows_per_item);
  const int64_t base = bag_layout ? sg.out_start[s] : sg.k_start[s];
  const int J = sg.T * H;
  stage_rows<TK>(Us, U + static_cast<int64_t>((s * sg.T + t) * H) * E, H);
  __syncthreads();

  float m[H], l[H], acc[H][16];
#pragma unroll
  for (int j = 0; j < H; ++j) {
    m[j] = -FLT_MAX; l[j] = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[j][e] = 0.f;
  }
  // the next row's loads are issued before the current row's arithmetic (one CTA per SM at ~200 registers: the warp has
  // to cover its own memory latency)
  float kv[16], pe[16], kvn[16], pen[16];
  int i = i0 + warp;
  if (i < i1) {
    Row<TK>::load(K + (base + i) * E, lane, kv);
    load_f32_row<TK>(PE + static_cast<int64_t>(i) * E, lane, pe);
  }
  for (; i < i1; i += WARPS) {
    const int64_t n = base + i;
    const int inext = i + WARPS;
    if (inext < i1) {
      Row<TK>::load(K + (base + inext) * E, lane, kvn);
      load_f32_row<TK>(PE + static_cast<int64_t>(inext) * E, lane, pen);
    }
    float kp[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) kp[e] = pe[e] + kv[e];
    float sc[H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
      float u[16];
      Row<TK>::sm_read(Us + j * E, lane, u);
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 16; ++e) d = fmaf(kp[e], u[e], d);
      sc[j] = d;
    }
    warp_sum8(sc, lane);
#pragma unroll
    for (int j = 0; j < H; ++j) sc[j] *= SCALE;
    float mine = 0.f;
#pragma unroll
    for (int j = 0; j < H; ++j) mine = (lane == j) ? sc[j] : mine;
    if (lane < H) S[n * J + t * H + lane] = mine;
#pragma unroll
    for (int j = 0; j < H; ++j) {
      if (sc[j] > m[j]) {             // warp-uniform: rescale the running sums to the new maximum
        const float f = xexp<TK>(m[j] - sc[j]);
        l[j] *= f;
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[j][e] *= f;
        m[j] = sc[j];
      }
      const float w = xexp<TK>(sc[j] - m[j]);
      l[j] += w;
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[j][e] = fmaf(w, kv[e], acc[j][e]);
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) { kv[e] = kvn[e]; pe[e] = pen[e]; }
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < H; ++j) { wm[warp][j] = m[j]; wl[warp][j] = l[j]; }
  }
  // fold the 8 warps head by head, fixed order
  const int64_t pbase = static_cast<int64_t>(item * sg.T + t) * H;
#pragma unroll 1
  for (int j = 0; j < H; ++j) {
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      float v = 0.f;
#pragma unroll
      for (int jj = 0; jj < H; ++jj) v = (jj == j) ? acc[jj][e] : v;
      red[warp * E + Row<TK>::col(lane, e)] = v;
    }
    __syncthreads();
    float gm = wm[0][j];
#pragma unroll
    for (int w = 1; w < WARPS; ++w) gm = fmaxf(gm, wm[w][j]);
    float f[WARPS];
#pragma unroll
    for (int w = 0; w < WARPS; ++w) f[w] = xexp<TK>(wm[w][j] - gm);
    for (int c = threadIdx.x; c < E; c += THREADS) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) a = fmaf(f[w], red[w * E + c], a);
      part_acc[(pbase + j) * E + c] = a;
    }
    if (threadIdx.x == 0) {
      float gl = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) gl = fmaf(f[w], wl[w][j], gl);
      part_ml[pbase + j] = make_float2(gm, gl);
    }
  }
}

// Sum of per-item partial rows [E], optionally weighted: the block's 512 threads are 128 float4 columns x 4 item slices
// (slice q takes items p0 + q, p0 + q + 4, ...; four loads in flight per thread), the slices are folded in fixed order.
constexpr int MERGE_THREADS = 512;
constexpr int MERGE_MAX_ITEMS = 2048;      // per segment (weights staged in shared memory)
template <bool WEIGHTED>
__device__ __forceinline__ float4 merge_items(const float* __restrict__ part, int64_t row_stride, int64_t row0, int p0, int p1,
                                              const float* wsm, float4* red) {
  const int c = (threadIdx.x & 127) * 4, q = threadIdx.x >> 7;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  int p = p0 + q;
  for (; p + 12 < p1; p += 16) {
    float4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const float4*>(part + (row0 + (p + 4 * k) * row_stride) * E + c);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float f = WEIGHTED ? wsm[p + 4 * k - p0] : 1.f;
      a.x = fmaf(f, v[k].x, a.x); a.y = fmaf(f, v[k].y, a.y); a.z = fmaf(f, v[k].z, a.z); a.w = fmaf(f, v[k].w, a.w);
    }
  }
  for (; p < p1; p += 4) {
    const float4 v = *reinterpret_cast<const float4*>(part + (row0 + p * row_stride) * E + c);
    const float f = WEIGHTED ? wsm[p - p0] : 1.f;
    a.x = fmaf(f, v.x, a.x); a.y = fmaf(f, v.y, a.y); a.z = fmaf(f, v.z, a.z); a.w = fmaf(f, v.w, a.w);
  }
  red[threadIdx.x] = a;
  __syncthreads();
  if (q == 0) {
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const float4 o = red[threadIdx.x + 128 * k];
      a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
    }
  }
  return a;      // valid in slice 0
}

// grid (n_segs*T*H), 512 threads: merge the items of a segment (softmax statistics first, then the weighted rows)
__global__ void __launch_bounds__(MERGE_THREADS)
k_t2i_merge(const float* __restrict__ part_acc, const float2* __restrict__ part_ml, const Segs sg, float* __restrict__ Pool,
            float* __restrict__ lse) {
  __shared__ float wsm[MERGE_MAX_ITEMS];
  __shared__ float4 red[MERGE_THREADS];
  __shared__ float wred[MERGE_THREADS / 32];
  __shared__ float gm_s, gl_s;
  const int row = blockIdx.x;                    // (s*T + t)*H + h
  const int h = row % H, st = row / H, t = st % sg.T, s = st / sg.T;
  const int p0 = sg.item0[s], p1 = sg.item0[s + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t stride = static_cast<int64_t>(sg.T) * H, r0 = static_cast<int64_t>(t) * H + h;
  float mx = -FLT_MAX;
  for (int p = p0 + threadIdx.x; p < p1; p += MERGE_THREADS) mx = fmaxf(mx, part_ml[p * stride + r0].x);
  mx = warp_max(mx);
  if (lane == 0) wred[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float g = wred[0];
    for (int w = 1; w < MERGE_THREADS / 32; ++w) g = fmaxf(g, wred[w]);
    gm_s = g;
  }
  __syncthreads();
  const float gm = gm_s;
  float ls = 0.f;
  for (int p = p0 + threadIdx.x; p < p1; p += MERGE_THREADS) {
    const float2 ml = part_ml[p * stride + r0];
    const float f = expf(ml.x - gm);
    wsm[p - p0] = f;
    ls = fmaf(ml.y, f, ls);
  }
  ls = warp_sum(ls);
  __syncthreads();                 // wred is reused
  if (lane == 0) wred[warp] = ls;
  __syncthreads();
  if (threadIdx.x == 0) {
    float g = 0.f;
    for (int w = 0; w < MERGE_THREADS / 32; ++w) g += wred[w];
    gl_s = g;
  }
  __syncthreads();
  float4 a = merge_items<true>(part_acc, stride, r0, p0, p1, wsm, red);
  if (threadIdx.x < 128) {
    const float inv = 1.f / gl_s;
    a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
    *reinterpret_cast<float4*>(Pool + static_cast<int64_t>(row) * E + threadIdx.x * 4) = a;
  }
  if (threadIdx.x == 0) lse[row] = gm + logf(gl_s);
}

// ---------------------------------------------------------------------------------------------------------------------
// token -> image attention, backward
// ---------------------------------------------------------------------------------------------------------------------
// Per row n of segment s and column j = (t, h):   a = exp(S - lse),  g = keys[n] . dPool[j],  dS = a (g - dPool[j].Pool[j])
//   dkeys[n] (+)= sum_j a dPool[j] + scale dS U[j]          dU[j] += scale dS (keys[n] + pe[n])
// grid (items); the CTA walks the T tokens (U_t, dPool_t staged per token), dK is read-modify-written by the warp that
// owns the row, dU partials go to part_du[(item*T + t)*8 + h][E].
template <typename TK>
__global__ void __launch_bounds__(THREADS)
k_t2i_bwd(const TK* __restrict__ K, const float* __restrict__ PE, const float* __restrict__ U, const float* __restrict__ S,
          const float* __restrict__ lse, const float* __restrict__ Pool, const float* __restrict__ dPool, const Segs sg,
          const int bag_layout, TK* __restrict__ dK, const int accumulate, float* __restrict__ part_du) {
  __shared__ __align__(16) float Us[H * E];
  __shared__ __align__(16) float Ds[H * E];
  __shared__ float delta[H], lses[H];
  float* red = Us;      // the fold of a token's dU runs after its row loop (barrier first): U_t is dead by then
  const int item = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = find_seg(sg, item);
  const int i0 = (item - sg.item0[s]) * sg.rows_per_item;
  const int i1 = min(sg.len[s], i0 + sg.rows_per_item);
  const int64_t base = bag_layout ? sg.out_start[s] : sg.k_start[s];
  const int J = sg.T * H;
#pragma unroll 1
  for (int t = 0; t < sg.T; ++t) {
    const int64_t urow = static_cast<int64_t>((s * sg.T + t) * H);
    __syncthreads();
    stage_rows<TK>(Us, U + urow * E, H);
    stage_rows<TK>(Ds, dPool + urow * E, H);
    {   // delta[h] = dPool[h] . Pool[h]: warp h (8 warps, 8 heads)
      float a[16], b[16];
      load_f32_row<TK>(dPool + (urow + warp) * E, lane, a);
      load_f32_row<TK>(Pool + (urow + warp) * E, lane, b);
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 16; ++e) d = fmaf(a[e], b[e], d);
      d = warp_sum(d);
      if (lane == 0) { delta[warp] = d; lses[warp] = lse[urow + warp]; }
    }
    __syncthreads();
    float du[H][16];
#pragma unroll
    for (int j = 0; j < H; ++j)
#pragma unroll
      for (int e = 0; e < 16; ++e) du[j][e] = 0.f;
    // (no software prefetch here: 128 dU accumulators + the row leave no registers for a second row in flight)
    for (int i = i0 + warp; i < i1; i += WARPS) {
      const int64_t n = base + i;
      float kv[16], pe[16], dk[16];
      Row<TK>::load(K + n * E, lane, kv);
      load_f32_row<TK>(PE + static_cast<int64_t>(i) * E, lane, pe);
      const float srow = (lane < H) ? S[n * J + t * H + lane] : 0.f;
      if (t > 0 || accumulate) {
        Row<TK>::load_cached(dK + n * E, lane, dk);
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) dk[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) pe[e] += kv[e];          // keys + pe
#pragma unroll
      for (int j = 0; j < H; ++j) {
        float d[16], u[16];
        Row<TK>::sm_read(Ds + j * E, lane, d);
        float g = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) g = fmaf(kv[e], d[e], g);
        g = warp_sum(g);
        const float a = xexp<TK>(__shfl_sync(0xffffffffu, srow, j) - lses[j]);
        const float ds = a * (g - delta[j]);
        const float cs = ds * SCALE;
        Row<TK>::sm_read(Us + j * E, lane, u);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          dk[e] = fmaf(a, d[e], dk[e]);
          dk[e] = fmaf(cs, u[e], dk[e]);
          du[j][e] = fmaf(cs, pe[e], du[j][e]);
        }
      }
      Row<TK>::store(dK + n * E, lane, dk);
    }
    // fold dU over the 8 warps, head by head, fixed order
    const int64_t pbase = static_cast<int64_t>(item * sg.T + t) * H;
#pragma unroll 1
    for (int j = 0; j < H; ++j) {
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        float v = 0.f;
#pragma unroll
        for (int jj = 0; jj < H; ++jj) v = (jj == j) ? du[jj][e] : v;
        red[warp * E + Row<TK>::col(lane, e)] = v;
      }
      __syncthreads();
      for (int c = threadIdx.x; c < E; c += THREADS) {
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) a += red[w * E + c];
        part_du[(pbase + j) * E + c] = a;
      }
    }
  }
}

// out[(s*T + t)*H + h][:] = sum over the items of segment s of part[(item*T + t)*H + h][:]
__global__ void __launch_bounds__(MERGE_THREADS)
k_sum_items(const float* __restrict__ part, const Segs sg, float* __restrict__ out) {
  __shared__ float4 red[MERGE_THREADS];
  const int row = blockIdx.x;
  const int h = row % H, st = row / H, t = st % sg.T, s = st / sg.T;
  const float4 a = merge_items<false>(part, static_cast<int64_t>(sg.T) * H, static_cast<int64_t>(t) * H + h, sg.item0[s],
                                      sg.item0[s + 1], nullptr, red);
  if (threadIdx.x < 128) *reinterpret_cast<float4*>(out + static_cast<int64_t>(row) * E + threadIdx.x * 4) = a;
}

// ---------------------------------------------------------------------------------------------------------------------
// LayerNorm(keys + row[segment])   (eps 1e-5)
// ---------------------------------------------------------------------------------------------------------------------
// grid (items); rows are read at k_start, written at out_start when bag_layout_out.
// TI / TO: storage of the input rows and of the result (fp32 key stream -> bf16 packed bag on the last layer).
template <typename TI, typename TO>
__global__ void __launch_bounds__(THREADS)
k_ln_seg_fwd(const TI* __restrict__ K, const float* __restrict__ R, const float* __restrict__ gamma, const float* __restrict__ beta,
             const Segs sg, const int bag_layout_out, TO* __restrict__ Y, float* __restrict__ mean, float* __restrict__ rstd) {
  const int item = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = find_seg(sg, item);
  const int i0 = (item - sg.item0[s]) * sg.rows_per_item;
  const int i1 = min(sg.len[s], i0 + sg.rows_per_item);
  const int64_t in0 = sg.k_start[s], out0 = bag_layout_out ? sg.out_start[s] : sg.k_start[s];
  float r[16], ga[16], be[16];
  load_f32_row<TO>(R + static_cast<int64_t>(s) * E, lane, r);
  load_f32_row<TO>(gamma, lane, ga);
  load_f32_row<TO>(beta, lane, be);
  for (int i = i0 + warp; i < i1; i += WARPS) {
    float v[16];
    rload<TI, TO>(K + (in0 + i) * E, lane, v);
    float sum = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) { v[e] += r[e]; sum += v[e]; }
    const float mu = warp_sum(sum) * (1.f / E);
    float q = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) { const float d = v[e] - mu; q = fmaf(d, d, q); }
    const float rs = rsqrtf(warp_sum(q) * (1.f / E) + 1e-5f);
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = (v[e] - mu) * rs * ga[e] + be[e];
    Row<TO>::store(Y + (out0 + i) * E, lane, v);
    if (lane == 0) { mean[in0 + i] = mu; rstd[in0 + i] = rs; }
  }
}

// dXR = rstd (g - mean(g) - xhat mean(g xhat)), g = dY gamma; per item: part[item][0..2][E] = (dgamma, dbeta, sum dXR)
template <typename TI, typename TO>
__global__ void __launch_bounds__(THREADS)
k_ln_seg_bwd(const TI* __restrict__ K, const float* __restrict__ R, const float* __restrict__ gamma, const float* __restrict__ mean,
             const float* __restrict__ rstd, const TO* __restrict__ dY, const Segs sg, const int bag_layout_out,
             TI* __restrict__ dK, const int accumulate, float* __restrict__ part) {
  using TK = TO;      // lane order of every operand
  __shared__ __align__(16) float red[WARPS * E];
  const int item = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = find_seg(sg, item);
  const int i0 = (item - sg.item0[s]) * sg.rows_per_item;
  const int i1 = min(sg.len[s], i0 + sg.rows_per_item);
  const int64_t in0 = sg.k_start[s], out0 = bag_layout_out ? sg.out_start[s] : sg.k_start[s];
  float r[16], ga[16], dg[16], db[16], dr[16];
  load_f32_row<TK>(R + static_cast<int64_t>(s) * E, lane, r);
  load_f32_row<TK>(gamma, lane, ga);
#pragma unroll
  for (int e = 0; e < 16; ++e) dg[e] = db[e] = dr[e] = 0.f;
  for (int i = i0 + warp; i < i1; i += WARPS) {
    float x[16], dy[16], g[16];
    rload<TI, TO>(K + (in0 + i) * E, lane, x);
    Row<TO>::load(dY + (out0 + i) * E, lane, dy);
    const float mu = mean[in0 + i], rs = rstd[in0 + i];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      x[e] = (x[e] + r[e] - mu) * rs;          // xhat
      g[e] = dy[e] * ga[e];
      s1 += g[e];
      s2 = fmaf(g[e], x[e], s2);
      dg[e] = fmaf(dy[e], x[e], dg[e]);
      db[e] += dy[e];
    }
    s1 = warp_sum(s1) * (1.f / E);
    s2 = warp_sum(s2) * (1.f / E);
    float o[16];
    if (accumulate) {
      rload<TI, TO>(dK + (in0 + i) * E, lane, o);
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) o[e] = 0.f;
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const float d = rs * (g[e] - s1 - x[e] * s2);
      dr[e] += d;
      o[e] += d;
    }
    rstore<TI, TO>(dK + (in0 + i) * E, lane, o);
  }
#pragma unroll 1
  for (int k = 0; k < 3; ++k) {
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 16; ++e) red[warp * E + Row<TK>::col(lane, e)] = (k == 0) ? dg[e] : (k == 1 ? db[e] : dr[e]);
    __syncthreads();
    for (int c = threadIdx.x; c < E; c += THREADS) {
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) a += red[w * E + c];
      part[(static_cast<int64_t>(item) * 3 + k) * E + c] = a;
    }
  }
}

// grid (2 + n_segs), 512 threads: block 0/1: dgamma / dbeta = sum over all items;  block 2+s: dR[s] = sum over the items of s
__global__ void __launch_bounds__(MERGE_THREADS)
k_ln_seg_reduce(const float* __restrict__ part, const Segs sg, float* __restrict__ dgamma, float* __restrict__ dbeta,
                const int accumulate, float* __restrict__ dR) {
  __shared__ float4 red[MERGE_THREADS];
  const int b = blockIdx.x, c = (threadIdx.x & 127) * 4;
  const int k = b < 2 ? b : 2;
  const int p0 = b < 2 ? 0 : sg.item0[b - 2], p1 = b < 2 ? sg.n_items : sg.item0[b - 1];
  float4 a = merge_items<false>(part, 3, k, p0, p1, nullptr, red);
  if (threadIdx.x >= 128) return;
  float* dst = (b == 0 ? dgamma : (b == 1 ? dbeta : dR + static_cast<int64_t>(b - 2) * E)) + c;
  if (b < 2 && accumulate) {
    const float4 o = *reinterpret_cast<float4*>(dst);
    a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
  }
  *reinterpret_cast<float4*>(dst) = a;
}

// token rows <-> their rows of the packed bag: grid (n_segs*T), 128 threads x 4 columns
template <typename TK>
__global__ void __launch_bounds__(128) k_tok_scatter(const float* __restrict__ tokens, const Segs sg, TK* __restrict__ bag) {
  const int r = blockIdx.x, c = threadIdx.x * 4;
  const float4 v = *reinterpret_cast<const float4*>(tokens + static_cast<int64_t>(r) * E + c);
  TK* dst = bag + static_cast<int64_t>(sg.tok_row[r / sg.T] + r % sg.T) * E + c;
  dst[0] = from_f32<TK>(v.x); dst[1] = from_f32<TK>(v.y); dst[2] = from_f32<TK>(v.z); dst[3] = from_f32<TK>(v.w);
}
template <typename TK>
__global__ void __launch_bounds__(128) k_tok_gather(const TK* __restrict__ dbag, const Segs sg, float* __restrict__ dtokens) {
  const int r = blockIdx.x, c = threadIdx.x * 4;
  const TK* src = dbag + static_cast<int64_t>(sg.tok_row[r / sg.T] + r % sg.T) * E + c;
  *reinterpret_cast<float4*>(dtokens + static_cast<int64_t>(r) * E + c) =
      make_float4(to_f32<TK>(src[0]), to_f32<TK>(src[1]), to_f32<TK>(src[2]), to_f32<TK>(src[3]));
}

// ---------------------------------------------------------------------------------------------------------------------
// head-block-diagonal products on the token side
// ---------------------------------------------------------------------------------------------------------------------
// y[r*H + h, :] = sum_c x[r, h*CH + c] W[h*CH + c, :]       grid (R*H), 128 threads x float4
__global__ void __launch_bounds__(128)
k_hd_expand(const float* __restrict__ x, const float* __restrict__ W, float* __restrict__ y) {
  __shared__ float xs[CH];
  const int row = blockIdx.x, r = row / H, h = row % H, c4 = threadIdx.x * 4;
  if (threadIdx.x < CH) xs[threadIdx.x] = x[static_cast<int64_t>(r) * CI + h * CH + threadIdx.x];
  __syncthreads();
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
  for (int c = 0; c < CH; ++c) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(W + static_cast<int64_t>(h * CH + c) * E + c4));
    const float xv = xs[c];
    a.x = fmaf(xv, w.x, a.x); a.y = fmaf(xv, w.y, a.y); a.z = fmaf(xv, w.z, a.z); a.w = fmaf(xv, w.w, a.w);
  }
  *reinterpret_cast<float4*>(y + static_cast<int64_t>(row) * E + c4) = a;
}

// x[r, h*CH + c] = y[r*H + h, :] . W[h*CH + c, :] (+ bias)   grid (R*H), 256 threads: warp w owns channels w*4 .. w*4+3
__global__ void __launch_bounds__(256)
k_hd_contract(const float* __restrict__ y, const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ x) {
  const int row = blockIdx.x, r = row / H, h = row % H, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float yv[16];
  load_f32_row<float>(y + static_cast<int64_t>(row) * E, lane, yv);
#pragma unroll
  for (int k = 0; k < CH / 8; ++k) {
    const int ch = h * CH + warp * (CH / 8) + k;
    float wv[16];
    load_f32_row<float>(W + static_cast<int64_t>(ch) * E, lane, wv);
    float d = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) d = fmaf(yv[e], wv[e], d);
    d = warp_sum(d);
    if (lane == 0) x[static_cast<int64_t>(r) * CI + ch] = d + (bias ? __ldg(bias + ch) : 0.f);
  }
}

// dW[ch, c4..] (+)= sum_r x[r, ch] y[r*H + ch/CH, c4..];  db[ch] (+)= sum_r x[r, ch]     one thread per float4 of dW
__global__ void __launch_bounds__(256)
k_hd_dw(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ dW, float* __restrict__ db, const int R,
        const int accumulate) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= CI * (E / 4)) return;
  const int ch = idx / (E / 4), c4 = (idx % (E / 4)) * 4, h = ch / CH;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  float sb = 0.f;
  for (int r = 0; r < R; ++r) {
    const float xv = x[static_cast<int64_t>(r) * CI + ch];
    const float4 v = *reinterpret_cast<const float4*>(y + static_cast<int64_t>(r * H + h) * E + c4);
    a.x = fmaf(xv, v.x, a.x); a.y = fmaf(xv, v.y, a.y); a.z = fmaf(xv, v.z, a.z); a.w = fmaf(xv, v.w, a.w);
    sb += xv;
  }
  float4* dst = reinterpret_cast<float4*>(dW + static_cast<int64_t>(ch) * E + c4);
  if (accumulate) { const float4 o = *dst; a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w; }
  *dst = a;
  if (c4 == 0 && db) db[ch] = accumulate ? db[ch] + sb : sb;
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
int make_segs(const milb200_segment* segs, int n_segs, int T, Segs* out) {
  MIL_CHECK_ARG(segs && n_segs >= 1 && n_segs <= MAXSEG, MILB200_EINVAL, "segments: count %d not in [1, %d]", n_segs, MAXSEG);
  MIL_CHECK_ARG(T >= 1 && T <= 16, MILB200_EINVAL, "segments: %d tokens per segment not in [1, 16]", T);
  Segs& g = *out;
  g.n = n_segs;
  g.T = T;
  int64_t total = 0;
  g.max_len = 0;
  for (int s = 0; s < n_segs; ++s) {
    MIL_CHECK_ARG(segs[s].len >= 1 && segs[s].k_start >= 0 && segs[s].out_start >= 0 && segs[s].tok_row >= 0, MILB200_EINVAL,
                  "segments: segment %d is empty or has a negative row", s);
    g.k_start[s] = segs[s].k_start; g.len[s] = segs[s].len; g.out_start[s] = segs[s].out_start; g.tok_row[s] = segs[s].tok_row;
    total += segs[s].len;
    g.max_len = std::max(g.max_len, segs[s].len);
  }
  // ~1 item per SM (the attention kernels hold ~200 registers per thread: one CTA per SM), 32..512 rows each: every
  // item folds 8 head accumulators through shared memory and leaves a 16 KB partial for the merge kernels, so small
  // items pay mostly for that; very large ones leave SMs idle
  int64_t rpi = (total + sm_count() - 1) / sm_count();
  rpi = std::min<int64_t>(512, std::max<int64_t>(32, (rpi + 7) / 8 * 8));
  g.rows_per_item = static_cast<int>(rpi);
  int items = 0;
  for (int s = 0; s < n_segs; ++s) {
    g.item0[s] = items;
    items += (g.len[s] + g.rows_per_item - 1) / g.rows_per_item;
    MIL_CHECK_ARG((g.len[s] + g.rows_per_item - 1) / g.rows_per_item <= MERGE_MAX_ITEMS, MILB200_EUNSUPPORTED,
                  "segments: segment %d has %d rows: more than %d work items", s, g.len[s], MERGE_MAX_ITEMS);
  }
  g.item0[n_segs] = items;
  for (int s = n_segs + 1; s <= MAXSEG; ++s) g.item0[s] = items;
  g.n_items = items;
  return MILB200_OK;
}

Segs finer(const Segs& sg, int factor) {
  Segs g = sg;
  g.rows_per_item = std::max(16, (sg.rows_per_item / factor + 7) / 8 * 8);
  int items = 0;
  for (int s = 0; s < g.n; ++s) {
    g.item0[s] = items;
    items += (g.len[s] + g.rows_per_item - 1) / g.rows_per_item;
  }
  for (int s = g.n; s <= MAXSEG; ++s) g.item0[s] = items;
  g.n_items = items;
  return g;
}
constexpr int LN_FINER = 4;

size_t t2i_ws_bytes(const Segs& sg) {
  const size_t rows = static_cast<size_t>(sg.n_items) * sg.T * H;
  return align_up(rows * E * sizeof(float), 256) + align_up(rows * sizeof(float2), 256) + 256;
}
size_t ln_seg_ws_bytes(const Segs& sg) { return static_cast<size_t>(finer(sg, LN_FINER).n_items) * 3 * E * sizeof(float) + 256; }

int headdiag_expand(const float* x, const float* W, float* y, int R, cudaStream_t st) {
  MIL_CHECK_ARG(x && W && y && R > 0, MILB200_EINVAL, "headdiag_expand: bad argument");
  k_hd_expand<<<R * H, 128, 0, st>>>(x, W, y);
  MIL_LAUNCH_CHECK();
  return MILB200_OK;
}
int headdiag_contract(const float* y, const float* W, const float* bias, float* x, int R, cudaStream_t st) {
  MIL_CHECK_ARG(x && W && y && R > 0, MILB200_EINVAL, "headdiag_contract: bad argument");
  k_hd_contract<<<R * H, 256, 0, st>>>(y, W, bias, x);
  MIL_LAUNCH_CHECK();
  return MILB200_OK;
}
int headdiag_dw(const float* x, const float* y, float* dW, float* db, int R, int accumulate, cudaStream_t st) {
  MIL_CHECK_ARG(x && y && dW && R > 0, MILB200_EINVAL, "headdiag_dw: bad argument");
  k_hd_dw<<<(CI * (E / 4) + 255) / 256, 256, 0, st>>>(x, y, dW, db, R, accumulate);
  MIL_LAUNCH_CHECK();
  return MILB200_OK;
}

int tok_scatter(const float* tokens, const Segs& sg, void* bag, int dtype, cudaStream_t st) {
  MIL_CHECK_ARG(tokens && bag, MILB200_EINVAL, "tok_scatter: null pointer");
  if (dtype == MILB200_BF16) k_tok_scatter<__nv_bfloat16><<<sg.n * sg.T, 128, 0, st>>>(tokens, sg, (__nv_bfloat16*)bag);
  else k_tok_scatter<float><<<sg.n * sg.T, 128, 0, st>>>(tokens, sg, (float*)bag);
  MIL_LAUNCH_CHECK();
  return MILB200_OK;
}
int tok_gather(const void* dbag, const Segs& sg, float* dtokens, int dtype, cudaStream_t st) {
  MIL_CHECK_ARG(dbag && dtokens, MILB200_EINVAL, "tok_gather: null pointer");
  if (dtype == MILB200_BF16) k_tok_gather<__nv_bfloat16><<<sg.n * sg.T, 128, 0, st>>>((const __nv_bfloat16*)dbag, sg, dtokens);
  else k_tok_gather<float><<<sg.n * sg.T, 128, 0, st>>>((const float*)dbag, sg, dtokens);
  MIL_LAUNCH_CHECK();
  return MILB200_OK;
}

int t2i_fwd(const void* K, const float* PE, const float* U, const Segs& sg, int bag_layout, float* S, float* lse, float* Pool,
            int dtype, void* ws, size_t ws_bytes, cudaStream_t st) {
  MIL_CHECK_ARG(K && PE && U && S && lse && Pool, MILB200_EINVAL, "t2i_fwd: null pointer");
  MIL_CHECK_ARG(ws && ws_bytes >= t2i_ws_bytes(sg), MILB200_EWORKSPACE, "t2i_fwd: workspace %zu < %zu", ws_bytes, t2i_ws_bytes(sg));
  const size_t rows = static_cast<size_t>(sg.n_items) * sg.T * H;
  float* part_acc = static_cast<float*>(ws);
  float2* part_ml = reinterpret_cast<float2*>(static_cast<char*>(ws) + align_up(rows * E * sizeof(float), 256));
  const dim3 grid(sg.n_items, sg.T);
  if (dtype == MILB200_BF16)
    k_t2i_fwd<__nv_bfloat16><<<grid, THREADS, 0, st>>>((const __nv_bfloat16*)K, PE, U, sg, bag_layout, S,
                                                       part_acc, part_ml);
  else
    k_t2i_fwd<float><<<grid, THREADS, 0, st>>>((const float*)K, PE, U, sg, bag_layout, S, part_acc, part_ml);
  MIL_LAUNCH_CHECK();
  k_t2i_merge<<<sg.n * sg.T * H, MERGE_THREADS, 0, st>>>(part_acc, part_ml, sg, Pool, lse);
  MIL_LAUNCH_CHECK();
  return MILB200_OK;
}

int t2i_bwd(const void* K, const float* PE, const float* U, const float* S, const float* lse, const float* Pool,
            const float* dPool, const Segs& sg, int bag_layout, void* dK, int accumulate_dk, float* dU, int dtype, void* ws,
            size_t ws_bytes, cudaStream_t st) {
  MIL_CHECK_ARG(K && PE && U && S && lse && Pool && dPool && dK && dU, MILB200_EINVAL, "t2i_bwd: null pointer");
  MIL_CHECK_ARG(ws && ws_bytes >= t2i_ws_bytes(sg), MILB200_EWORKSPACE, "t2i_bwd: workspace %zu < %zu", ws_bytes, t2i_ws_bytes(sg));
  float* part = static_cast<float*>(ws);
  if (dtype == MILB200_BF16)
    k_t2i_bwd<__nv_bfloat16><<<sg.n_items, THREADS, 0, st>>>((const __nv_bfloat16*)K, PE, U, S, lse, Pool,
                                                             dPool, sg, bag_layout, (__nv_bfloat16*)dK, accumulate_dk, part);
  else
    k_t2i_bwd<float><<<sg.n_items, THREADS, 0, st>>>((const float*)K, PE, U, S, lse, Pool, dPool, sg, bag_layout,
                                                     (float*)dK, accumulate_dk, part);
  MIL_LAUNCH_CHECK();
  k_sum_items<<<sg.n * sg.T * H, MERGE_THREADS, 0, st>>>(part, sg, dU);
  MIL_LAUNCH_CHECK();
  return MILB200_OK;
}

int ln_seg_fwd(const void* K, const float* R, const float* gamma, const float* beta, const Segs& sg_in, int bag_layout_out,
               void* Y, float* mean, float* rstd, int in_dtype, int out_dtype, cudaStream_t st) {
  MIL_CHECK_ARG(K && R && gamma && beta && Y && mean && rstd, MILB200_EINVAL, "ln_seg_fwd: null pointer");
  const Segs sg = finer(sg_in, LN_FINER);
  MIL_CHECK_ARG(in_dtype == out_dtype || (in_dtype == MILB200_F32 && out_dtype == MILB200_BF16), MILB200_EUNSUPPORTED,
                "ln_seg_fwd: storage pair (in %d, out %d) is not built", in_dtype, out_dtype);
  const unsigned grid = sg.n_items;
  using bf = __nv_bfloat16;
  if (out_dtype == MILB200_F32)
    k_ln_seg_fwd<float, float><<<grid, THREADS, 0, st>>>((const float*)K, R, gamma, beta, sg, bag_layout_out, (float*)Y, mean, rstd);
  else if (in_dtype == MILB200_BF16)
    k_ln_seg_fwd<bf, bf><<<grid, THREADS, 0, st>>>((const bf*)K, R, gamma, beta, sg, bag_layout_out, (bf*)Y, mean, rstd);
  else
    k_ln_seg_fwd<float, bf><<<grid, THREADS, 0, st>>>((const float*)K, R, gamma, beta, sg, bag_layout_out, (bf*)Y, mean, rstd);
  MIL_LAUNCH_CHECK();
  return MILB200_OK;
}

int ln_seg_bwd(const void* K, const float* R, const float* gamma, const float* mean, const float* rstd, const void* dY,
               const Segs& sg_in, int bag_layout_out, void* dK, int accumulate_dk, float* dR, float* dgamma, float* dbeta,
               int accumulate_params, int in_dtype, int out_dtype, void* ws, size_t ws_bytes, cudaStream_t st) {
  MIL_CHECK_ARG(K && R && gamma && mean && rstd && dY && dK && dR && dgamma && dbeta, MILB200_EINVAL, "ln_seg_bwd: null pointer");
  const Segs sg = finer(sg_in, LN_FINER);
  MIL_CHECK_ARG(in_dtype == out_dtype || (in_dtype == MILB200_F32 && out_dtype == MILB200_BF16), MILB200_EUNSUPPORTED,
                "ln_seg_bwd: storage pair (in %d, out %d) is not built", in_dtype, out_dtype);
  MIL_CHECK_ARG(ws && ws_bytes >= ln_seg_ws_bytes(sg_in), MILB200_EWORKSPACE, "ln_seg_bwd: workspace %zu < %zu", ws_bytes,
                ln_seg_ws_bytes(sg_in));
  float* part = static_cast<float*>(ws);
  const unsigned rgrid = 2 + sg.n;
  using bf = __nv_bfloat16;
  if (out_dtype == MILB200_F32) {
    k_ln_seg_bwd<float, float><<<sg.n_items, THREADS, 0, st>>>((const float*)K, R, gamma, mean, rstd, (const float*)dY, sg,
                                                               bag_layout_out, (float*)dK, accumulate_dk, part);
  } else if (in_dtype == MILB200_BF16) {
    k_ln_seg_bwd<bf, bf><<<sg.n_items, THREADS, 0, st>>>((const bf*)K, R, gamma, mean, rstd, (const bf*)dY, sg, bag_layout_out,
                                                         (bf*)dK, accumulate_dk, part);
  } else {
    k_ln_seg_bwd<float, bf><<<sg.n_items, THREADS, 0, st>>>((const float*)K, R, gamma, mean, rstd, (const bf*)dY, sg,
                                                            bag_layout_out, (float*)dK, accumulate_dk, part);
  }
  MIL_LAUNCH_CHECK();
  k_ln_seg_reduce<<<rgrid, MERGE_THREADS, 0, st>>>(part, sg, dgamma, dbeta, accumulate_params, dR);
  MIL_LAUNCH_CHECK();
  return MILB200_OK;
}

}  // namespace xf
}  // namespace milb200
