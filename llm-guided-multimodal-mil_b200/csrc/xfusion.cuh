// xfusion.cuh — host entry points of the collapsed cross-modal attention kernels (xfusion.cu), used by the tape
// executor (tape.cu).  See xfusion.cu for the algebra.
#pragma once

#include "common.cuh"

namespace milb200 {
namespace xf {

constexpr int E = 512;          // joint embedding width of the fusion path (model/aggregator.py:13)
constexpr int H = 8;            // attention heads (aggregator.py:72-76)
constexpr int CI = 256;         // internal width of the cross attentions (attention_downsample_rate = 2)
constexpr int CH = CI / H;      // 32 channels per head
constexpr int MAXSEG = MILB200_MAX_SEGMENTS;

// Segment table passed BY VALUE to the kernels (<= 16 segments: the CT bag and the pathology bag of up to 8 patients).
// A segment is one image-side bag: `len` key rows starting at row k_start of the key stream; in the packed multi-modal
// bag of aggregator.py:173 the same rows start at out_start and the segment's T token rows at tok_row.  Work is cut
// into items of rows_per_item consecutive rows of ONE segment; item0[s] is the first item of segment s.
struct Segs {
  int n, T, rows_per_item, n_items, max_len;
  int k_start[MAXSEG], len[MAXSEG], out_start[MAXSEG], tok_row[MAXSEG], item0[MAXSEG + 1];
};
// The LayerNorm kernels are light (100-160 registers): they take a finer cut of the same segments, ~4 CTAs per SM.
Segs finer(const Segs& sg, int factor);

int make_segs(const milb200_segment* segs, int n_segs, int T, Segs* out);

// workspace (bytes) the kernels below need for `sg`
size_t t2i_ws_bytes(const Segs& sg);
size_t ln_seg_ws_bytes(const Segs& sg);

// y[r*H + h, :] = sum_c x[r, h*CH + c] W[h*CH + c, :]            x [R, CI] f32, W [CI, E] f32, y [R*H, E] f32
int headdiag_expand(const float* x, const float* W, float* y, int R, cudaStream_t st);
// x[r, h*CH + c] = sum_e y[r*H + h, e] W[h*CH + c, e] (+ bias[h*CH + c])
int headdiag_contract(const float* y, const float* W, const float* bias, float* x, int R, cudaStream_t st);
// dW[h*CH + c, :] (+)= sum_r x[r, h*CH + c] y[r*H + h, :];  db[h*CH + c] (+)= sum_r x[r, h*CH + c]  (db may be NULL)
int headdiag_dw(const float* x, const float* y, float* dW, float* db, int R, int accumulate, cudaStream_t st);

// token -> image attention with the key / value projections folded into the token side (see xfusion.cu):
//   S[n, j] = scale (K[n] + PE[pos(n)]) . U[seg(n), j],  a = softmax over the segment's rows,  Pool[seg, j] = sum_n a K[n]
// K in `dtype` (rows of E), PE fp32 (rows of E, row i = position i inside a segment), U / Pool [n_segs*T*H, E] f32, S [rows, T*H] f32, lse [n_segs*T*H] f32.
int t2i_fwd(const void* K, const float* PE, const float* U, const Segs& sg, int bag_layout, float* S, float* lse,
            float* Pool, int dtype, void* ws, size_t ws_bytes, cudaStream_t st);
// backward: dK (dtype; overwritten or accumulated in place) and dU [n_segs*T*H, E] f32 from dPool.
int t2i_bwd(const void* K, const float* PE, const float* U, const float* S, const float* lse, const float* Pool,
            const float* dPool, const Segs& sg, int bag_layout, void* dK, int accumulate_dk, float* dU, int dtype,
            void* ws, size_t ws_bytes, cudaStream_t st);

// Y[out(n)] = LayerNorm(K[n] + R[seg(n)]) * gamma + beta  (R: ONE row per segment — the image -> token attention with
// a single token).  bag_layout_out: rows are written at out_start (the packed bag) instead of k_start.
// Storage pairs (in, out): equal, or fp32 key stream -> bf16 packed bag (the last layer of a bf16 program).
int ln_seg_fwd(const void* K, const float* R, const float* gamma, const float* beta, const Segs& sg, int bag_layout_out,
               void* Y, float* mean, float* rstd, int in_dtype, int out_dtype, cudaStream_t st);
// dK (in place accumulate optional), dR [n_segs, E] f32, dgamma/dbeta (accumulate optional)
int ln_seg_bwd(const void* K, const float* R, const float* gamma, const float* mean, const float* rstd, const void* dY,
               const Segs& sg, int bag_layout_out, void* dK, int accumulate_dk, float* dR, float* dgamma, float* dbeta,
               int accumulate_params, int in_dtype, int out_dtype, void* ws, size_t ws_bytes, cudaStream_t st);

// bag[tok_row[s] + t] = tokens[s*T + t] (cast to dtype) / dtokens[s*T + t] = dbag[tok_row[s] + t] (fp32)
int tok_scatter(const float* tokens, const Segs& sg, void* bag, int dtype, cudaStream_t st);
int tok_gather(const void* dbag, const Segs& sg, float* dtokens, int dtype, cudaStream_t st);

}  // namespace xf
}  // namespace milb200
