// tape.cu — native executor for static operator programs ("tapes") over the library's own kernels.
//
// Why: the cross-modal fusion path (model/sam/transformer.py:58-120,278-309 inside model/aggregator.py:134-203) is
// ~130 kernel launches forward and ~370 backward per bag.  Driven op by op from Python (one autograd node, several
// allocations and a ctypes call per launch) it is host-bound at ~25 us per launch; the reference has the same shape
// of problem (eager PyTorch, ~120 launches per TwoWayTransformer call).  Here the host side of the whole
// program is ONE call: the Python module describes its forward once as a list of ops over numbered tensor slots,
// this file runs it, keeps the activations in a caller-provided arena, and runs the reverse program for the
// backward (a small reverse-mode engine: per-slot gradient buffers, first-write/accumulate tracking, parameter
// gradients accumulated into one flat fp32 buffer).  No allocation, no synchronisation, everything on `stream`.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "xfusion.cuh"

namespace milb200 {

struct TapePlan {
  std::vector<size_t> slot_off;   // arena offset of every internal slot (SIZE_MAX for external ones)
  std::vector<size_t> aux_off;    // arena offset of every op's saved statistics (lse / mean,rstd)
  size_t arena_bytes = 0;
  size_t max_slot_bytes = 0;
  size_t scratch_fwd = 0, scratch_bwd = 0;
  int lanes = 1;                  // 2 when any op asks for lane 1
  std::vector<int> alias_of;      // JOIN: slot whose storage (value and gradient) is the head of another slot's, else -1
  bool has_segs = false;
  xf::Segs sg;
};

static inline bool slot_ext(const milb200_tape_slot& s) { return (s.external & MILB200_SLOT_EXTERNAL) != 0; }
// an op's arithmetic/storage dtype is its slots' own: the program dtype unless the slot is forced to fp32
static inline int slot_dtype(const milb200_tape_slot& s, int dtype) { return (s.external & MILB200_SLOT_F32) ? MILB200_F32 : dtype; }
static inline size_t slot_bytes(const milb200_tape_slot& s, int dtype) {
  return static_cast<size_t>(s.rows) * static_cast<size_t>(s.cols) * elem_size(slot_dtype(s, dtype));
}
static inline bool is_seg_op(int kind) {
  return kind == MILB200_OP_T2I_POOL || kind == MILB200_OP_LN_SEG || kind == MILB200_OP_TOK_SCATTER;
}

static int tape_validate(const milb200_tape_op* ops, int n_ops, const milb200_tape_slot* slots, int n_slots,
                         const milb200_tape_param* params, int n_params, int dtype, const milb200_segments* segs) {
  auto dt = [&](int s) { return slot_dtype(slots[s], dtype); };
  const int n_segs = segs ? segs->n_segs : 0, T = segs ? segs->tokens : 0;
  MIL_CHECK_ARG(ops && slots && n_ops > 0 && n_slots > 0, MILB200_EINVAL, "tape: empty program");
  auto ok_slot = [&](int s, bool optional) { return (optional && s < 0) || (s >= 0 && s < n_slots); };
  for (int i = 0; i < n_slots; ++i)
    MIL_CHECK_ARG(slots[i].rows > 0 && slots[i].cols > 0, MILB200_EINVAL, "tape: slot %d has shape %lld x %d", i,
                  (long long)slots[i].rows, slots[i].cols);
  for (int i = 0; i < n_ops; ++i) {
    const milb200_tape_op& o = ops[i];
    MIL_CHECK_ARG(ok_slot(o.out, false) && ok_slot(o.in0, false), MILB200_EINVAL, "tape: op %d has a bad slot id", i);
    MIL_CHECK_ARG(o.lane == 0 || o.lane == 1, MILB200_EINVAL, "tape: op %d asks for lane %d (two lanes: 0, 1)", i, o.lane);
    const milb200_tape_slot& so = slots[o.out];
    const milb200_tape_slot& s0 = slots[o.in0];
    switch (o.kind) {
      case MILB200_OP_LINEAR: {
        MIL_CHECK_ARG(ok_slot(o.in1, true), MILB200_EINVAL, "tape: op %d has a bad `add` slot", i);
        MIL_CHECK_ARG(params && o.p0 >= 0 && o.p0 < n_params && o.p1 < n_params, MILB200_EINVAL, "tape: op %d bad param id", i);
        const milb200_tape_param& w = params[o.p0];
        MIL_CHECK_ARG(w.rows == so.cols && w.cols == s0.cols && so.rows == s0.rows, MILB200_EINVAL,
                      "tape: op %d linear shape mismatch: x %lld x %d, W %d x %d, y %lld x %d", i, (long long)s0.rows, s0.cols,
                      w.rows, w.cols, (long long)so.rows, so.cols);
        MIL_CHECK_ARG(o.in1 < 0 || (o.in1 != o.in0 && slots[o.in1].rows == s0.rows && slots[o.in1].cols == s0.cols),
                      MILB200_EINVAL, "tape: op %d `add` slot must differ from x and share its shape", i);
        MIL_CHECK_ARG(o.p1 < 0 || params[o.p1].rows * params[o.p1].cols == so.cols, MILB200_EINVAL, "tape: op %d bias size", i);
        MIL_CHECK_ARG((dt(o.out) == dt(o.in0) && (o.in1 < 0 || dt(o.in1) == dt(o.in0))) ||
                          (dt(o.in0) == MILB200_BF16 && dt(o.out) == MILB200_F32 && o.in1 < 0),
                      MILB200_EINVAL, "tape: op %d linear operands must share one dtype (or bf16 input -> fp32 result)", i);
        break;
      }
      case MILB200_OP_ATTENTION: {
        MIL_CHECK_ARG(ok_slot(o.in1, false) && ok_slot(o.in2, false), MILB200_EINVAL, "tape: op %d bad k/v slot", i);
        MIL_CHECK_ARG(o.in0 != o.in1 && o.in0 != o.in2 && o.in1 != o.in2, MILB200_EINVAL,
                      "tape: op %d attention operands must be distinct slots", i);
        const milb200_tape_slot& sk = slots[o.in1];
        const milb200_tape_slot& sv = slots[o.in2];
        MIL_CHECK_ARG(o.a0 > 0 && s0.cols % o.a0 == 0 && sk.cols == s0.cols && sv.cols == s0.cols && sk.rows == sv.rows &&
                          so.rows == s0.rows && so.cols == s0.cols,
                      MILB200_EINVAL, "tape: op %d attention shape mismatch", i);
        break;
      }
      case MILB200_OP_LAYERNORM: {
        MIL_CHECK_ARG(ok_slot(o.in1, true), MILB200_EINVAL, "tape: op %d bad residual slot", i);
        MIL_CHECK_ARG(params && o.p0 >= 0 && o.p0 < n_params && o.p1 >= 0 && o.p1 < n_params, MILB200_EINVAL,
                      "tape: op %d bad param id", i);
        // the residual may be a single row that is broadcast over all rows of in0 (T = 1 shortcut)
        MIL_CHECK_ARG(so.rows == s0.rows && so.cols == s0.cols &&
                          (o.in1 < 0 || (o.in1 != o.in0 && slots[o.in1].cols == s0.cols &&
                                         (slots[o.in1].rows == s0.rows || slots[o.in1].rows == 1))),
                      MILB200_EINVAL, "tape: op %d layernorm shape mismatch", i);
        break;
      }
      case MILB200_OP_ADD: {
        MIL_CHECK_ARG(ok_slot(o.in1, false) && o.in1 != o.in0, MILB200_EINVAL, "tape: op %d add needs two distinct slots", i);
        MIL_CHECK_ARG(so.rows == s0.rows && so.cols == s0.cols && slots[o.in1].rows == s0.rows && slots[o.in1].cols == s0.cols,
                      MILB200_EINVAL, "tape: op %d add shape mismatch", i);
        MIL_CHECK_ARG(dt(o.out) == dt(o.in0) && dt(o.in1) == dt(o.in0), MILB200_EINVAL, "tape: op %d add dtype mismatch", i);
        break;
      }
      case MILB200_OP_JOIN: {
        MIL_CHECK_ARG(ok_slot(o.in1, false) && o.in1 != o.in0 && o.out != o.in0 && o.out != o.in1, MILB200_EINVAL,
                      "tape: op %d join needs three distinct slots", i);
        MIL_CHECK_ARG(!slot_ext(so) && so.rows == s0.rows + slots[o.in1].rows && so.cols == s0.cols && slots[o.in1].cols == s0.cols &&
                          dt(o.out) == dt(o.in0) && dt(o.in1) == dt(o.in0),
                      MILB200_EINVAL, "tape: op %d join shape/dtype mismatch (the result must be an internal slot)", i);
        break;
      }
      case MILB200_OP_HEADDIAG_U:
      case MILB200_OP_HEADDIAG_O: {
        const bool up = o.kind == MILB200_OP_HEADDIAG_U;
        const milb200_tape_slot& small = up ? s0 : so;
        const milb200_tape_slot& big = up ? so : s0;
        MIL_CHECK_ARG(params && o.p0 >= 0 && o.p0 < n_params && o.p1 < n_params && params[o.p0].rows == xf::CI &&
                          params[o.p0].cols == xf::E && (o.p1 < 0 || (!up && params[o.p1].rows * params[o.p1].cols == xf::CI)),
                      MILB200_EINVAL, "tape: op %d head-diagonal product needs a [256, 512] weight", i);
        MIL_CHECK_ARG(small.cols == xf::CI && big.cols == xf::E && big.rows == small.rows * xf::H && dt(o.in0) == MILB200_F32 &&
                          dt(o.out) == MILB200_F32,
                      MILB200_EINVAL, "tape: op %d head-diagonal product: [R, 256] <-> [R*8, 512] fp32 slots", i);
        break;
      }
      case MILB200_OP_T2I_POOL: {
        MIL_CHECK_ARG(ok_slot(o.in1, false) && ok_slot(o.in2, false) && n_segs > 0, MILB200_EINVAL,
                      "tape: op %d t2i_pool needs keys, position table, U and a segment table", i);
        MIL_CHECK_ARG(s0.cols == xf::E && slots[o.in1].cols == xf::E && dt(o.in1) == MILB200_F32 && slots[o.in2].cols == xf::E &&
                          slots[o.in2].rows == static_cast<int64_t>(n_segs) * T * xf::H && so.rows == slots[o.in2].rows &&
                          so.cols == xf::E && dt(o.in2) == MILB200_F32 && dt(o.out) == MILB200_F32,
                      MILB200_EINVAL, "tape: op %d t2i_pool shape/dtype mismatch", i);
        break;
      }
      case MILB200_OP_LN_SEG: {
        MIL_CHECK_ARG(ok_slot(o.in1, false) && o.in2 < 0 && n_segs > 0 && T == 1, MILB200_EINVAL,
                      "tape: op %d ln_seg needs keys, one row per segment and a segment table with one token per segment", i);
        MIL_CHECK_ARG(params && o.p0 >= 0 && o.p0 < n_params && o.p1 >= 0 && o.p1 < n_params, MILB200_EINVAL,
                      "tape: op %d bad param id", i);
        MIL_CHECK_ARG(s0.cols == xf::E && so.cols == xf::E &&
                          (dt(o.out) == dt(o.in0) || (dt(o.in0) == MILB200_F32 && dt(o.out) == MILB200_BF16)) &&
                          slots[o.in1].cols == xf::E &&
                          slots[o.in1].rows == n_segs && dt(o.in1) == MILB200_F32 && ((o.a0 & 1) || so.rows == s0.rows),
                      MILB200_EINVAL, "tape: op %d ln_seg shape/dtype mismatch", i);
        break;
      }
      case MILB200_OP_TOK_SCATTER: {
        MIL_CHECK_ARG(ok_slot(o.in1, false) && n_segs > 0 && o.in1 != o.out, MILB200_EINVAL,
                      "tape: op %d tok_scatter needs token rows, the bag and a segment table", i);
        MIL_CHECK_ARG(s0.cols == xf::E && s0.rows == static_cast<int64_t>(n_segs) * T && dt(o.in0) == MILB200_F32 &&
                          slot_ext(so) && slot_ext(slots[o.in1]) && so.cols == xf::E && slots[o.in1].cols == xf::E &&
                          so.rows == slots[o.in1].rows && dt(o.out) == dt(o.in1),
                      MILB200_EINVAL, "tape: op %d tok_scatter shape/dtype mismatch (bag slots must be external)", i);
        break;
      }
      default:
        MIL_CHECK_ARG(false, MILB200_EINVAL, "tape: op %d has unknown kind %d", i, o.kind);
    }
  }
  return MILB200_OK;
}

static int tape_plan(const milb200_tape_op* ops, int n_ops, const milb200_tape_slot* slots, int n_slots, int dtype,
                     const milb200_segments* segs, TapePlan* out) {
  TapePlan& p = *out;
  p.slot_off.assign(n_slots, SIZE_MAX);
  p.aux_off.assign(n_ops, SIZE_MAX);
  p.alias_of.assign(n_slots, -1);
  if (segs && segs->n_segs > 0) {
    int rc = xf::make_segs(segs->seg, segs->n_segs, segs->tokens, &p.sg);
    if (rc) return rc;
    p.has_segs = true;
  }
  // JOIN: an internal first operand is produced in place at the head of the result
  for (int i = 0; i < n_ops; ++i)
    if (ops[i].kind == MILB200_OP_JOIN && !slot_ext(slots[ops[i].in0]) && p.alias_of[ops[i].in0] < 0) p.alias_of[ops[i].in0] = ops[i].out;
  size_t off = 0;
  for (int i = 0; i < n_slots; ++i) {
    const size_t b = slot_bytes(slots[i], dtype);
    p.max_slot_bytes = std::max(p.max_slot_bytes, b);
    if (slot_ext(slots[i]) || p.alias_of[i] >= 0) continue;
    off = align_up(off, 256);
    p.slot_off[i] = off;
    off += b;
  }
  for (int i = 0; i < n_slots; ++i)
    if (p.alias_of[i] >= 0) p.slot_off[i] = p.slot_off[p.alias_of[i]];
  for (int i = 0; i < n_ops; ++i) {
    const milb200_tape_op& o = ops[i];
    size_t aux = 0;
    if (o.kind == MILB200_OP_ATTENTION) aux = sizeof(float) * static_cast<size_t>(o.a0) * slots[o.in0].rows;
    if (o.kind == MILB200_OP_LAYERNORM) aux = sizeof(float) * 2 * static_cast<size_t>(slots[o.in0].rows);
    if (is_seg_op(o.kind)) MIL_CHECK_ARG(p.has_segs, MILB200_EINVAL, "tape: op %d needs a segment table", i);
    if (o.kind == MILB200_OP_T2I_POOL)      // scores [rows, T*H] + lse [n_segs*T*H]
      aux = sizeof(float) * (static_cast<size_t>(slots[o.in0].rows) * p.sg.T * xf::H + static_cast<size_t>(slots[o.out].rows));
    if (o.kind == MILB200_OP_LN_SEG) aux = sizeof(float) * 2 * static_cast<size_t>(slots[o.in0].rows);
    if (aux) {
      off = align_up(off, 256);
      p.aux_off[i] = off;
      off += aux;
    }
    size_t f = 0, b = 0;
    const int dt0 = slot_dtype(slots[o.in0], dtype);
    if (o.kind == MILB200_OP_LINEAR) {
      f = milb200_linear_workspace_bytes(slots[o.in0].rows, slots[o.out].cols, slots[o.in0].cols, dt0, 0);
      b = milb200_linear_workspace_bytes(slots[o.in0].rows, slots[o.out].cols, slots[o.in0].cols, dt0, 1);
    } else if (o.kind == MILB200_OP_ATTENTION) {
      const int c = slots[o.in0].cols / o.a0;
      f = milb200_attention_workspace_bytes(slots[o.in0].rows, slots[o.in1].rows, o.a0, c, 0);
      b = milb200_attention_workspace_bytes(slots[o.in0].rows, slots[o.in1].rows, o.a0, c, 1);
    } else if (o.kind == MILB200_OP_LAYERNORM) {
      b = milb200_layernorm_workspace_bytes(slots[o.in0].rows, slots[o.in0].cols) + 256 + sizeof(float) * slots[o.in0].cols;
    } else if (o.kind == MILB200_OP_T2I_POOL) {
      f = b = xf::t2i_ws_bytes(p.sg);
    } else if (o.kind == MILB200_OP_LN_SEG) {
      b = xf::ln_seg_ws_bytes(p.sg);
    }
    p.scratch_fwd = std::max(p.scratch_fwd, f);
    p.scratch_bwd = std::max(p.scratch_bwd, b);
    if (o.lane == 1) p.lanes = 2;
  }
  p.arena_bytes = align_up(off, 256) + 256;
  p.scratch_fwd = align_up(p.scratch_fwd, 256) + 256;
  p.scratch_bwd = align_up(p.scratch_bwd, 256) + 256;
  return MILB200_OK;
}

// backward workspace image: per lane [kernel scratch][3 temporaries of max_slot_bytes], then [gradient buffer of every
// internal slot]
static size_t tape_bwd_lane_bytes(const TapePlan& p) { return p.scratch_bwd + 3 * align_up(p.max_slot_bytes, 256); }
static size_t tape_bwd_ws_bytes(const TapePlan& p) { return p.lanes * tape_bwd_lane_bytes(p) + p.arena_bytes; }

// ---- lanes: two streams, ordered through events wherever they share a resource -----------------------------------
// Every op records an event on its lane's stream; a resource (slot value, slot gradient, parameter gradient) remembers
// its last toucher, and an op on the other lane waits for that toucher's event first.  Touches of one resource are
// thereby totally ordered (same lane: stream order; other lane: event), which is all the reverse-mode accumulation
// needs.  Inside a stream capture the same calls become the edges of two parallel graph branches.
struct LaneRun {
  cudaStream_t st[2] = {nullptr, nullptr};
  bool two = false;
  std::vector<cudaEvent_t>* pool = nullptr;
  size_t next_ev = 0;
  struct Touch { int lane = -1; int64_t seq = -1; cudaEvent_t ev = nullptr; };
  std::vector<Touch> touch;          // per resource
  int64_t seq = 0;
  int64_t waited[2] = {-1, -1};      // waited[l]: newest op (seq) of the OTHER lane that lane l has already waited for
  bool used1 = false;

  cudaEvent_t fresh() {
    if (next_ev == pool->size()) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
      pool->push_back(e);
    }
    return (*pool)[next_ev++];
  }
  int begin(cudaStream_t main, cudaStream_t aux, int n_resources, std::vector<cudaEvent_t>* p) {
    st[0] = main;
    st[1] = aux;
    two = aux != nullptr;
    pool = p;
    touch.assign(n_resources, Touch{});
    return MILB200_OK;
  }
  // call once the prologue (seeds, memsets) is enqueued on lane 0
  int fork() {
    if (!two) return MILB200_OK;
    cudaEvent_t e = fresh();
    MIL_CHECK_ARG(e != nullptr, MILB200_ECUDA, "tape: cannot create an event");
    MIL_CUDA(cudaEventRecord(e, st[0]));
    MIL_CUDA(cudaStreamWaitEvent(st[1], e, 0));
    return MILB200_OK;
  }
  cudaStream_t stream_of(int lane) const { return two ? st[lane] : st[0]; }
  int before(int lane, const int* res, int n) {
    if (!two) return MILB200_OK;
    if (lane == 1) used1 = true;
    for (int i = 0; i < n; ++i) {
      if (res[i] < 0) continue;
      const Touch& t = touch[res[i]];
      if (t.lane < 0 || t.lane == lane || t.seq <= waited[lane]) continue;
      MIL_CUDA(cudaStreamWaitEvent(st[lane], t.ev, 0));
      waited[lane] = t.seq;
    }
    return MILB200_OK;
  }
  int after(int lane, const int* res, int n) {
    if (!two) return MILB200_OK;
    cudaEvent_t e = fresh();
    MIL_CHECK_ARG(e != nullptr, MILB200_ECUDA, "tape: cannot create an event");
    MIL_CUDA(cudaEventRecord(e, st[lane]));
    const int64_t s = seq++;
    for (int i = 0; i < n; ++i)
      if (res[i] >= 0) touch[res[i]] = Touch{lane, s, e};
    return MILB200_OK;
  }
  int join() {
    if (!two) return MILB200_OK;
    cudaEvent_t e = fresh();      // always join: lane 1 was forked into the caller's stream / capture
    MIL_CHECK_ARG(e != nullptr, MILB200_ECUDA, "tape: cannot create an event");
    MIL_CUDA(cudaEventRecord(e, st[1]));
    MIL_CUDA(cudaStreamWaitEvent(st[0], e, 0));
    return MILB200_OK;
  }
};
static thread_local std::vector<cudaEvent_t> t_lane_events;

}  // namespace milb200

using namespace milb200;

extern "C" {

size_t milb200_tape_arena_bytes(const milb200_tape_op* ops, int n_ops, const milb200_tape_slot* slots, int n_slots, int dtype,
                                const milb200_segments* segs) {
  if (!ops || !slots || n_ops <= 0 || n_slots <= 0) return 256;
  TapePlan p;
  if (tape_plan(ops, n_ops, slots, n_slots, dtype, segs, &p)) return 256;
  return p.arena_bytes;
}

size_t milb200_tape_workspace_bytes(const milb200_tape_op* ops, int n_ops, const milb200_tape_slot* slots, int n_slots,
                                    int dtype, int backward, const milb200_segments* segs) {
  if (!ops || !slots || n_ops <= 0 || n_slots <= 0) return 256;
  TapePlan p;
  if (tape_plan(ops, n_ops, slots, n_slots, dtype, segs, &p)) return 256;
  return backward ? tape_bwd_ws_bytes(p) : p.lanes * p.scratch_fwd;
}

static int tape_forward_run(const milb200_tape_op* ops, int n_ops, const milb200_tape_slot* slots, int n_slots,
                            const milb200_tape_param* params, int n_params, void* const* ext_ptrs, const void* w_compute,
                            const float* p_f32, void* arena, size_t arena_bytes, void* workspace, size_t ws_bytes, int dtype,
                            const milb200_segments* segs, void* stream_main, void* stream_aux) {
  MIL_CHECK_ARG(dtype == MILB200_F32 || dtype == MILB200_BF16, MILB200_EINVAL, "tape: bad dtype %d", dtype);
  int rc = tape_validate(ops, n_ops, slots, n_slots, params, n_params, dtype, segs);
  if (rc) return rc;
  MIL_CHECK_ARG(ext_ptrs && w_compute && p_f32 && arena, MILB200_EINVAL, "tape_forward: null pointer");
  TapePlan pl;
  if ((rc = tape_plan(ops, n_ops, slots, n_slots, dtype, segs, &pl))) return rc;
  MIL_CHECK_ARG(arena_bytes >= pl.arena_bytes, MILB200_EWORKSPACE, "tape_forward: arena %zu < %zu", arena_bytes, pl.arena_bytes);
  MIL_CHECK_ARG(workspace && ws_bytes >= pl.lanes * pl.scratch_fwd, MILB200_EWORKSPACE, "tape_forward: workspace %zu < %zu",
                ws_bytes, pl.lanes * pl.scratch_fwd);
  LaneRun lr;
  lr.begin(static_cast<cudaStream_t>(stream_main), pl.lanes == 2 ? static_cast<cudaStream_t>(stream_aux) : nullptr, n_slots,
           &t_lane_events);
  if ((rc = lr.fork())) return rc;
  char* ar = static_cast<char*>(arena);
  std::vector<void*> ptr(n_slots);
  for (int i = 0; i < n_slots; ++i) {
    ptr[i] = slot_ext(slots[i]) ? ext_ptrs[i] : static_cast<void*>(ar + pl.slot_off[i]);
    MIL_CHECK_ARG(ptr[i] != nullptr && aligned16(ptr[i]), MILB200_EALIGN, "tape_forward: slot %d pointer is null or misaligned", i);
  }
  auto dt = [&](int s) { return slot_dtype(slots[s], dtype); };
  // weights of an op: the compute-dtype image for program-dtype operands, the fp32 master for fp32 slots
  auto weight = [&](int p, int d) -> const void* {
    return d == dtype ? static_cast<const void*>(static_cast<const char*>(w_compute) + params[p].offset * elem_size(dtype))
                      : static_cast<const void*>(p_f32 + params[p].offset);
  };
  for (int i = 0; i < n_ops; ++i) {
    const milb200_tape_op& o = ops[i];
    const milb200_tape_slot& s0 = slots[o.in0];
    const milb200_tape_slot& so = slots[o.out];
    const int res[4] = {o.in0, o.in1, o.in2, o.out};
    if ((rc = lr.before(o.lane, res, 4))) return rc;
    void* stream = lr.stream_of(o.lane);
    cudaStream_t cst = static_cast<cudaStream_t>(stream);
    void* workspace_l = static_cast<char*>(workspace) + (lr.two ? o.lane : 0) * pl.scratch_fwd;
    const size_t ws_l = pl.scratch_fwd;
    const int d0 = dt(o.in0);
    switch (o.kind) {
      case MILB200_OP_LINEAR:
        if (d0 == MILB200_BF16 && dt(o.out) == MILB200_F32) {
          rc = milb200_linear_f32out_fwd(ptr[o.in0], weight(o.p0, d0), o.p1 >= 0 ? p_f32 + params[o.p1].offset : nullptr,
                                         static_cast<float*>(ptr[o.out]), s0.rows, so.cols, s0.cols, o.a0, stream);
          break;
        }
        rc = milb200_linear_fwd(ptr[o.in0], o.in1 >= 0 ? ptr[o.in1] : nullptr, weight(o.p0, d0),
                                o.p1 >= 0 ? p_f32 + params[o.p1].offset : nullptr, ptr[o.out], s0.rows, so.cols, s0.cols, o.a0,
                                d0, workspace_l, ws_l, stream);
        break;
      case MILB200_OP_ATTENTION:
        rc = milb200_attention_fwd(ptr[o.in0], ptr[o.in1], ptr[o.in2], ptr[o.out], reinterpret_cast<float*>(ar + pl.aux_off[i]),
                                   s0.rows, slots[o.in1].rows, o.a0, s0.cols / o.a0, d0, workspace_l, ws_l, stream);
        break;
      case MILB200_OP_LAYERNORM: {
        float* mean = reinterpret_cast<float*>(ar + pl.aux_off[i]);
        const int bcast = (o.in1 >= 0 && slots[o.in1].rows == 1 && s0.rows > 1) ? 1 : 0;
        rc = milb200_layernorm_fwd(ptr[o.in0], o.in1 >= 0 ? ptr[o.in1] : nullptr, p_f32 + params[o.p0].offset,
                                   p_f32 + params[o.p1].offset, ptr[o.out], mean, mean + s0.rows, s0.rows, s0.cols, d0,
                                   bcast, stream);
        break;
      }
      case MILB200_OP_ADD:
        rc = milb200_add(ptr[o.in0], ptr[o.in1], ptr[o.out], s0.rows * s0.cols, d0, stream);
        break;
      case MILB200_OP_JOIN: {
        char* dst = static_cast<char*>(ptr[o.out]);
        const size_t b0 = slot_bytes(s0, dtype);
        if (pl.alias_of[o.in0] != o.out) {
          MIL_CUDA(cudaMemcpyAsync(dst, ptr[o.in0], b0, cudaMemcpyDeviceToDevice, cst));
          count_launch();
        }
        MIL_CUDA(cudaMemcpyAsync(dst + b0, ptr[o.in1], slot_bytes(slots[o.in1], dtype), cudaMemcpyDeviceToDevice, cst));
        count_launch();
        break;
      }
      case MILB200_OP_HEADDIAG_U:
        rc = xf::headdiag_expand(static_cast<const float*>(ptr[o.in0]), p_f32 + params[o.p0].offset,
                                 static_cast<float*>(ptr[o.out]), static_cast<int>(s0.rows), cst);
        break;
      case MILB200_OP_HEADDIAG_O:
        rc = xf::headdiag_contract(static_cast<const float*>(ptr[o.in0]), p_f32 + params[o.p0].offset,
                                   o.p1 >= 0 ? p_f32 + params[o.p1].offset : nullptr, static_cast<float*>(ptr[o.out]),
                                   static_cast<int>(so.rows), cst);
        break;
      case MILB200_OP_T2I_POOL: {
        MIL_CHECK_ARG(slots[o.in1].rows >= pl.sg.max_len, MILB200_EINVAL, "tape: op %d position table has %lld rows < longest segment %d",
                      i, (long long)slots[o.in1].rows, pl.sg.max_len);
        float* S = reinterpret_cast<float*>(ar + pl.aux_off[i]);
        float* lse = S + static_cast<size_t>(s0.rows) * pl.sg.T * xf::H;
        rc = xf::t2i_fwd(ptr[o.in0], static_cast<const float*>(ptr[o.in1]), static_cast<const float*>(ptr[o.in2]), pl.sg, o.a0 & 1, S, lse,
                         static_cast<float*>(ptr[o.out]), d0, workspace_l, ws_l, cst);
        break;
      }
      case MILB200_OP_LN_SEG: {
        float* mean = reinterpret_cast<float*>(ar + pl.aux_off[i]);
        rc = xf::ln_seg_fwd(ptr[o.in0], static_cast<const float*>(ptr[o.in1]), p_f32 + params[o.p0].offset,
                            p_f32 + params[o.p1].offset, pl.sg, o.a0 & 1, ptr[o.out], mean, mean + s0.rows, d0, dt(o.out), cst);
        break;
      }
      case MILB200_OP_TOK_SCATTER:
        MIL_CHECK_ARG(ptr[o.out] == ptr[o.in1], MILB200_EINVAL, "tape: op %d tok_scatter works in place: out and in1 must share one address", i);
        rc = xf::tok_scatter(static_cast<const float*>(ptr[o.in0]), pl.sg, ptr[o.out], dt(o.out), cst);
        break;
      default:
        rc = MILB200_EINVAL;
    }
    if (rc) return rc;
    if ((rc = lr.after(o.lane, res, 4))) return rc;
  }
  return lr.join();
}

static int tape_backward_run(const milb200_tape_op* ops, int n_ops, const milb200_tape_slot* slots, int n_slots,
                             const milb200_tape_param* params, int n_params, void* const* ext_ptrs, void* const* ext_grad_ptrs,
                             const void* const* seed_ptrs, const void* w_compute, const float* p_f32, float* g_f32,
                             const void* arena, size_t arena_bytes, void* workspace, size_t ws_bytes, int dtype,
                             const milb200_segments* segs, void* stream_main, void* stream_aux) {
  MIL_CHECK_ARG(dtype == MILB200_F32 || dtype == MILB200_BF16, MILB200_EINVAL, "tape: bad dtype %d", dtype);
  int rc = tape_validate(ops, n_ops, slots, n_slots, params, n_params, dtype, segs);
  if (rc) return rc;
  MIL_CHECK_ARG(ext_ptrs && ext_grad_ptrs && seed_ptrs && w_compute && p_f32 && g_f32 && arena, MILB200_EINVAL,
                "tape_backward: null pointer");
  TapePlan pl;
  if ((rc = tape_plan(ops, n_ops, slots, n_slots, dtype, segs, &pl))) return rc;
  MIL_CHECK_ARG(arena_bytes >= pl.arena_bytes, MILB200_EWORKSPACE, "tape_backward: arena %zu < %zu", arena_bytes, pl.arena_bytes);
  const size_t need = tape_bwd_ws_bytes(pl);
  MIL_CHECK_ARG(workspace && ws_bytes >= need, MILB200_EWORKSPACE, "tape_backward: workspace %zu < %zu", ws_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream_main);
  const char* ar = static_cast<const char*>(arena);
  char* ws = static_cast<char*>(workspace);
  const size_t scratch_bytes = pl.scratch_bwd;
  const size_t tmp_stride = align_up(pl.max_slot_bytes, 256);
  const size_t lane_bytes = tape_bwd_lane_bytes(pl);
  char* gbase = ws + pl.lanes * lane_bytes;
  // the lane of the op being processed selects the stream, kernel scratch and temporaries the lambdas below use
  void* stream = stream_main;
  void* scratch = ws;
  char* tmp0 = ws + pl.scratch_bwd;
  LaneRun lr;
  lr.begin(st, pl.lanes == 2 ? static_cast<cudaStream_t>(stream_aux) : nullptr, n_slots + std::max(n_params, 0), &t_lane_events);
  auto dt = [&](int s) { return slot_dtype(slots[s], dtype); };
  auto weight = [&](int p, int d) -> const void* {
    return d == dtype ? static_cast<const void*>(static_cast<const char*>(w_compute) + params[p].offset * elem_size(dtype))
                      : static_cast<const void*>(p_f32 + params[p].offset);
  };

  std::vector<const void*> val(n_slots);
  std::vector<void*> grad(n_slots);
  std::vector<char> has(n_slots, 0), needs(n_slots, 1);
  for (int i = 0; i < n_slots; ++i) {
    if (slot_ext(slots[i])) {
      val[i] = ext_ptrs[i];
      grad[i] = ext_grad_ptrs[i];
      needs[i] = grad[i] != nullptr;
    } else {
      val[i] = ar + pl.slot_off[i];
      grad[i] = gbase + pl.slot_off[i];
    }
    MIL_CHECK_ARG(val[i] != nullptr, MILB200_EINVAL, "tape_backward: slot %d has no value pointer", i);
  }
  // upstream gradients: the first contribution of their slots
  for (int i = 0; i < n_slots; ++i) {
    if (!seed_ptrs[i]) continue;
    MIL_CHECK_ARG(grad[i] != nullptr, MILB200_EINVAL, "tape_backward: seeded slot %d has no gradient buffer", i);
    if (grad[i] != seed_ptrs[i])     // equal pointers: the caller seeded its own gradient buffer in place
      MIL_CUDA(cudaMemcpyAsync(grad[i], seed_ptrs[i], slot_bytes(slots[i], dtype), cudaMemcpyDeviceToDevice, st));
    has[i] = 1;
  }
  std::vector<char> ptouched(n_params > 0 ? n_params : 1, 0);
  {  // parameters no gradient reaches must read as zero: clear the whole flat gradient range first
    int64_t total = 0;
    for (int i = 0; i < n_params; ++i)
      total = std::max<int64_t>(total, params[i].offset + static_cast<int64_t>(params[i].rows) * params[i].cols);
    if (total > 0) MIL_CUDA(cudaMemsetAsync(g_f32, 0, sizeof(float) * static_cast<size_t>(total), st));
  }
  if ((rc = lr.fork())) return rc;

  // where an op should write its gradient w.r.t. slot s: straight into the slot's buffer if nothing is there yet,
  // else into temporary `t` (folded in by settle())
  auto target = [&](int s, int t) -> void* {
    if (s < 0 || !needs[s]) return nullptr;
    return has[s] ? static_cast<void*>(tmp0 + t * tmp_stride) : grad[s];
  };
  auto settle = [&](int s, void* wrote) -> int {
    if (s < 0 || !needs[s] || !wrote) return MILB200_OK;
    if (wrote == grad[s]) { has[s] = 1; return MILB200_OK; }
    const int64_t n = slots[s].rows * slots[s].cols;
    if (!has[s]) {
      MIL_CUDA(cudaMemcpyAsync(grad[s], wrote, slot_bytes(slots[s], dtype), cudaMemcpyDeviceToDevice,
                               static_cast<cudaStream_t>(stream)));
      count_launch();
      has[s] = 1;
      return MILB200_OK;
    }
    return milb200_add(grad[s], wrote, grad[s], n, dt(s), stream);
  };

  for (int i = n_ops - 1; i >= 0; --i) {
    const milb200_tape_op& o = ops[i];
    if (!has[o.out]) continue;  // no gradient reaches this op
    const milb200_tape_slot& s0 = slots[o.in0];
    const milb200_tape_slot& so = slots[o.out];
    const void* dY = grad[o.out];
    const int res[6] = {o.out, o.in0, o.in1, o.in2, o.p0 >= 0 ? n_slots + o.p0 : -1, o.p1 >= 0 ? n_slots + o.p1 : -1};
    if ((rc = lr.before(o.lane, res, 6))) return rc;
    {
      const int ln = lr.two ? o.lane : 0;
      stream = lr.stream_of(o.lane);
      scratch = ws + ln * lane_bytes;
      tmp0 = ws + ln * lane_bytes + pl.scratch_bwd;
    }
    cudaStream_t cst = static_cast<cudaStream_t>(stream);
    const int d0 = dt(o.in0);
    switch (o.kind) {
      case MILB200_OP_LINEAR: {
        const bool need_dx = needs[o.in0] || (o.in1 >= 0 && needs[o.in1]);
        void* dX = nullptr;
        if (need_dx) dX = needs[o.in0] ? target(o.in0, 0) : target(o.in1, 0);
        const int acc = ptouched[o.p0] ? 1 : 0;
        if (d0 == MILB200_BF16 && dt(o.out) == MILB200_F32)
          rc = milb200_linear_f32out_bwd(val[o.in0], weight(o.p0, d0), static_cast<const float*>(val[o.out]),
                                         static_cast<const float*>(dY), dX, g_f32 + params[o.p0].offset,
                                         o.p1 >= 0 ? g_f32 + params[o.p1].offset : nullptr, s0.rows, so.cols, s0.cols, o.a0,
                                         acc, scratch, scratch_bytes, stream);
        else
          rc = milb200_linear_bwd(val[o.in0], o.in1 >= 0 ? val[o.in1] : nullptr, weight(o.p0, d0), val[o.out], dY, dX,
                                  g_f32 + params[o.p0].offset, o.p1 >= 0 ? g_f32 + params[o.p1].offset : nullptr, s0.rows,
                                  so.cols, s0.cols, o.a0, d0, acc, scratch, scratch_bytes, stream);
        if (rc) return rc;
        ptouched[o.p0] = 1;
        if (o.p1 >= 0) ptouched[o.p1] = 1;
        if (need_dx) {
          if (needs[o.in0]) {
            rc = settle(o.in0, dX);
            if (rc) return rc;
            // after settle, the complete contribution lives in dX (direct or temporary); the `add` operand gets the same
            if (o.in1 >= 0) rc = settle(o.in1, dX);
          } else {
            rc = settle(o.in1, dX);
          }
          if (rc) return rc;
        }
        break;
      }
      case MILB200_OP_ATTENTION: {
        void* dQ = target(o.in0, 0);
        void* dK = target(o.in1, 1);
        void* dV = target(o.in2, 2);
        // the kernels always produce all three; route unwanted ones to the temporaries
        void* dQw = dQ ? dQ : static_cast<void*>(tmp0);
        void* dKw = dK ? dK : static_cast<void*>(tmp0 + tmp_stride);
        void* dVw = dV ? dV : static_cast<void*>(tmp0 + 2 * tmp_stride);
        rc = milb200_attention_bwd(val[o.in0], val[o.in1], val[o.in2], val[o.out],
                                   reinterpret_cast<const float*>(ar + pl.aux_off[i]), dY, dQw, dKw, dVw, s0.rows,
                                   slots[o.in1].rows, o.a0, s0.cols / o.a0, d0, scratch, scratch_bytes, stream);
        if (rc) return rc;
        if ((rc = settle(o.in0, dQ))) return rc;
        if ((rc = settle(o.in1, dK))) return rc;
        if ((rc = settle(o.in2, dV))) return rc;
        break;
      }
      case MILB200_OP_LAYERNORM: {
        const float* mean = reinterpret_cast<const float*>(ar + pl.aux_off[i]);
        const int bcast = (o.in1 >= 0 && slots[o.in1].rows == 1 && s0.rows > 1) ? 1 : 0;
        const bool need_res = o.in1 >= 0 && needs[o.in1];
        const bool need_dx = needs[o.in0] || need_res;
        void* dX = needs[o.in0] ? target(o.in0, 0) : ((need_res && !bcast) ? target(o.in1, 0) : nullptr);
        if (!dX) dX = tmp0;  // the kernel always writes dXR
        const int acc = ptouched[o.p0] ? 1 : 0;
        const size_t ln_ws = milb200_layernorm_workspace_bytes(s0.rows, s0.cols);
        rc = milb200_layernorm_bwd(val[o.in0], o.in1 >= 0 ? val[o.in1] : nullptr, p_f32 + params[o.p0].offset, mean,
                                   mean + s0.rows, dY, dX, g_f32 + params[o.p0].offset, g_f32 + params[o.p1].offset, s0.rows,
                                   s0.cols, d0, acc, bcast, scratch, ln_ws, stream);
        if (rc) return rc;
        ptouched[o.p0] = ptouched[o.p1] = 1;
        if (need_dx) {
          if (bcast && need_res) {
            // gradient of the broadcast row = column sums of dXR (fp32), converted to the slot dtype; dX may live in tmp0,
            // so the converted row goes to temporary 1
            float* csum = reinterpret_cast<float*>(static_cast<char*>(scratch) + align_up(ln_ws, 256));
            if ((rc = milb200_colsum(dX, s0.rows, s0.cols, csum, d0, 0, stream))) return rc;
            void* row = csum;  // fp32 slots take the sums as they are (consumed in stream order before scratch is reused)
            if (d0 != MILB200_F32) {
              row = tmp0 + tmp_stride;
              if ((rc = milb200_cast(csum, MILB200_F32, row, d0, s0.cols, stream))) return rc;
            }
            if ((rc = settle(o.in1, row))) return rc;
          }
          if (needs[o.in0]) {
            if ((rc = settle(o.in0, dX))) return rc;
            if (!bcast && need_res && (rc = settle(o.in1, dX))) return rc;
          } else if (!bcast && need_res) {
            if ((rc = settle(o.in1, dX))) return rc;
          }
        }
        break;
      }
      case MILB200_OP_ADD: {
        // d(in0) = d(in1) = dY: fold the op's gradient buffer into both operands (no kernel of its own)
        if ((rc = settle(o.in0, grad[o.out]))) return rc;
        if ((rc = settle(o.in1, grad[o.out]))) return rc;
        break;
      }
      case MILB200_OP_JOIN: {
        char* g = static_cast<char*>(grad[o.out]);
        if (pl.alias_of[o.in0] == o.out) {
          // the head of the result's gradient buffer IS in0's gradient buffer
          MIL_CHECK_ARG(!has[o.in0], MILB200_EINVAL, "tape_backward: op %d: a JOIN operand produced in place has another consumer", i);
          has[o.in0] = 1;
        } else if ((rc = settle(o.in0, g))) {
          return rc;
        }
        if ((rc = settle(o.in1, g + slot_bytes(s0, dtype)))) return rc;
        break;
      }
      case MILB200_OP_HEADDIAG_U: {
        // out = expand(in0, W):  d in0 = contract(dY, W),  dW += in0 (x) dY
        float* dx = static_cast<float*>(target(o.in0, 0));
        if (dx && (rc = xf::headdiag_contract(static_cast<const float*>(dY), p_f32 + params[o.p0].offset, nullptr, dx,
                                              static_cast<int>(s0.rows), cst)))
          return rc;
        if ((rc = xf::headdiag_dw(static_cast<const float*>(val[o.in0]), static_cast<const float*>(dY),
                                  g_f32 + params[o.p0].offset, nullptr, static_cast<int>(s0.rows), ptouched[o.p0] ? 1 : 0, cst)))
          return rc;
        ptouched[o.p0] = 1;
        if ((rc = settle(o.in0, dx))) return rc;
        break;
      }
      case MILB200_OP_HEADDIAG_O: {
        // out = contract(in0, W) + b:  d in0 = expand(dY, W),  dW += dY (x) in0,  db += colsum(dY)
        float* dy_big = static_cast<float*>(target(o.in0, 0));
        if (dy_big && (rc = xf::headdiag_expand(static_cast<const float*>(dY), p_f32 + params[o.p0].offset, dy_big,
                                                static_cast<int>(so.rows), cst)))
          return rc;
        if ((rc = xf::headdiag_dw(static_cast<const float*>(dY), static_cast<const float*>(val[o.in0]),
                                  g_f32 + params[o.p0].offset, o.p1 >= 0 ? g_f32 + params[o.p1].offset : nullptr,
                                  static_cast<int>(so.rows), ptouched[o.p0] ? 1 : 0, cst)))
          return rc;
        ptouched[o.p0] = 1;
        if (o.p1 >= 0) ptouched[o.p1] = 1;
        if ((rc = settle(o.in0, dy_big))) return rc;
        break;
      }
      case MILB200_OP_T2I_POOL: {
        const float* S = reinterpret_cast<const float*>(ar + pl.aux_off[i]);
        const float* lse = S + static_cast<size_t>(s0.rows) * pl.sg.T * xf::H;
        // dkeys is accumulated in place by the kernel (every row belongs to one warp): no temporary, no add pass
        void* dK = needs[o.in0] ? grad[o.in0] : static_cast<void*>(tmp0);
        const int acc = (needs[o.in0] && has[o.in0]) ? 1 : 0;
        float* dU = static_cast<float*>(target(o.in2, 1));
        float* dUw = dU ? dU : reinterpret_cast<float*>(tmp0 + tmp_stride);
        rc = xf::t2i_bwd(val[o.in0], static_cast<const float*>(val[o.in1]), static_cast<const float*>(val[o.in2]), S, lse,
                         static_cast<const float*>(val[o.out]), static_cast<const float*>(dY), pl.sg, o.a0 & 1, dK, acc, dUw, d0,
                         scratch, scratch_bytes, cst);
        if (rc) return rc;
        if (needs[o.in0]) has[o.in0] = 1;
        if ((rc = settle(o.in2, dU))) return rc;
        break;
      }
      case MILB200_OP_LN_SEG: {
        const float* mean = reinterpret_cast<const float*>(ar + pl.aux_off[i]);
        void* dK = needs[o.in0] ? grad[o.in0] : static_cast<void*>(tmp0);
        const int acc = (needs[o.in0] && has[o.in0]) ? 1 : 0;
        float* dR = static_cast<float*>(target(o.in1, 1));
        float* dRw = dR ? dR : reinterpret_cast<float*>(tmp0 + tmp_stride);
        rc = xf::ln_seg_bwd(val[o.in0], static_cast<const float*>(val[o.in1]), p_f32 + params[o.p0].offset, mean,
                            mean + s0.rows, dY, pl.sg, o.a0 & 1, dK, acc, dRw, g_f32 + params[o.p0].offset,
                            g_f32 + params[o.p1].offset, ptouched[o.p0] ? 1 : 0, d0, dt(o.out), scratch, scratch_bytes, cst);
        if (rc) return rc;
        ptouched[o.p0] = ptouched[o.p1] = 1;
        if (needs[o.in0]) has[o.in0] = 1;
        if ((rc = settle(o.in1, dR))) return rc;
        break;
      }
      case MILB200_OP_TOK_SCATTER: {
        MIL_CHECK_ARG(grad[o.out] == grad[o.in1], MILB200_EINVAL,
                      "tape_backward: op %d tok_scatter works in place: out and in1 must share one gradient buffer", i);
        float* dtok = static_cast<float*>(target(o.in0, 0));
        if (dtok && (rc = xf::tok_gather(dY, pl.sg, dtok, dt(o.out), cst))) return rc;
        if ((rc = settle(o.in0, dtok))) return rc;
        has[o.in1] = 1;        // same buffer: the key rows of the bag gradient are already in place
        break;
      }
      default:
        return MILB200_EINVAL;
    }
    if ((rc = lr.after(o.lane, res, 6))) return rc;
  }
  return lr.join();
}


// ---- CUDA-graph replay --------------------------------------------------------------------------------------------
// A tape call enqueues 100-350 kernels; at ~3 us of launch cost each the HOST is the bottleneck for one-bag-per-step
// training (train_ddp.py:75).  The program is static, so a call whose shapes and pointers have been seen before is
// replayed as one CUDA graph: key = hash(program, slot shapes, every pointer, sizes, dtype).  First sight runs
// eagerly (a one-off bag must not pay for instantiation), second sight captures on a private stream (PyTorch's
// current stream is usually the legacy default stream, which cannot be captured) and launches on the caller's
// stream, later sights replay.  MILB200_TAPE_GRAPHS=0 switches it off.
}  // extern "C" (reopened below)

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_set>

namespace milb200 {
namespace {

struct GraphEntry {
  uint64_t key;
  cudaGraphExec_t exec;
  int nodes;
  uint64_t last_use;
};
std::mutex g_graph_mu;
std::vector<GraphEntry> g_graph_cache;
std::unordered_set<uint64_t> g_graph_seen;
cudaStream_t g_capture_stream = nullptr;
cudaStream_t g_capture_lane = nullptr;         // lane 1 while capturing
thread_local cudaStream_t t_eager_lane = nullptr;   // lane 1 when the program runs eagerly on the caller's stream

cudaStream_t eager_lane() {
  if (!t_eager_lane && cudaStreamCreateWithFlags(&t_eager_lane, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    t_eager_lane = nullptr;      // no second stream: LaneRun then keeps everything on the main one
  }
  return t_eager_lane;
}
uint64_t g_graph_tick = 0;
constexpr size_t GRAPH_CACHE_MAX = 32;

bool graphs_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MILB200_TAPE_GRAPHS");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

struct Hasher {
  uint64_t h = 1469598103934665603ull;
  void bytes(const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  }
  template <typename T> void val(const T& v) { bytes(&v, sizeof(T)); }
};

// run `body(stream, lane-1 stream)` either eagerly, or as a captured / replayed graph keyed by `key`
template <class Body>
int run_keyed(uint64_t key, cudaStream_t user_stream, Body body) {
  if (!graphs_enabled()) return body(user_stream, eager_lane());
  std::unique_lock<std::mutex> lock(g_graph_mu);
  ++g_graph_tick;
  for (auto& e : g_graph_cache) {
    if (e.key == key) {
      e.last_use = g_graph_tick;
      cudaGraphExec_t exec = e.exec;
      const int nodes = e.nodes;
      lock.unlock();
      MIL_CUDA(cudaGraphLaunch(exec, user_stream));
      count_launch(nodes);
      return MILB200_OK;
    }
  }
  if (!g_graph_seen.count(key)) {
    if (g_graph_seen.size() > 8192) g_graph_seen.clear();
    g_graph_seen.insert(key);
    lock.unlock();
    return body(user_stream, eager_lane());
  }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(user_stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    lock.unlock();
    return body(user_stream, eager_lane());  // the caller is capturing already: just enqueue into its capture
  }
  if (!g_capture_stream) MIL_CUDA(cudaStreamCreateWithFlags(&g_capture_stream, cudaStreamNonBlocking));
  if (!g_capture_lane) MIL_CUDA(cudaStreamCreateWithFlags(&g_capture_lane, cudaStreamNonBlocking));
  const int64_t before = milb200_launch_count();
  MIL_CUDA(cudaStreamBeginCapture(g_capture_stream, cudaStreamCaptureModeThreadLocal));
  const int rc = body(g_capture_stream, g_capture_lane);
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(g_capture_stream, &graph);
  count_launch(-static_cast<int>(milb200_launch_count() - before));  // nothing ran yet
  if (rc != MILB200_OK || ce != cudaSuccess || graph == nullptr) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    lock.unlock();
    if (rc != MILB200_OK) return rc;
    return body(user_stream, eager_lane());  // capture refused: run eagerly
  }
  size_t nodes = 0;
  cudaGraphGetNodes(graph, nullptr, &nodes);
  cudaGraphExec_t exec = nullptr;
  ce = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess || exec == nullptr) {
    cudaGetLastError();
    lock.unlock();
    return body(user_stream, eager_lane());
  }
  if (g_graph_cache.size() >= GRAPH_CACHE_MAX) {
    size_t victim = 0;
    for (size_t i = 1; i < g_graph_cache.size(); ++i)
      if (g_graph_cache[i].last_use < g_graph_cache[victim].last_use) victim = i;
    cudaGraphExecDestroy(g_graph_cache[victim].exec);  // freed once its last launch has drained
    g_graph_cache.erase(g_graph_cache.begin() + victim);
  }
  g_graph_cache.push_back(GraphEntry{key, exec, static_cast<int>(nodes), g_graph_tick});
  lock.unlock();
  MIL_CUDA(cudaGraphLaunch(exec, user_stream));
  count_launch(static_cast<int>(nodes));
  return MILB200_OK;
}

}  // namespace
}  // namespace milb200

extern "C" {

static void hash_segs(Hasher& h, const milb200_segments* segs) {
  const int n = segs ? segs->n_segs : 0;
  h.val(n);
  if (n > 0) {
    h.val(segs->tokens);
    h.bytes(segs->seg, sizeof(milb200_segment) * n);
  }
}

int milb200_tape_forward(const milb200_tape_op* ops, int n_ops, const milb200_tape_slot* slots, int n_slots,
                         const milb200_tape_param* params, int n_params, void* const* ext_ptrs, const void* w_compute,
                         const float* p_f32, void* arena, size_t arena_bytes, void* workspace, size_t ws_bytes, int dtype,
                         const milb200_segments* segs, void* stream) {
  MIL_CHECK_ARG(ops && slots && n_ops > 0 && n_slots > 0 && ext_ptrs, MILB200_EINVAL, "tape_forward: null pointer");
  MIL_CHECK_ARG(!segs || segs->n_segs == 0 || (segs->seg && segs->n_segs > 0 && segs->n_segs <= MILB200_MAX_SEGMENTS),
                MILB200_EINVAL, "tape_forward: bad segment table");
  Hasher h;
  h.val(static_cast<int>(1));
  h.bytes(ops, sizeof(milb200_tape_op) * n_ops);
  h.bytes(slots, sizeof(milb200_tape_slot) * n_slots);
  if (params && n_params > 0) h.bytes(params, sizeof(milb200_tape_param) * n_params);
  h.bytes(ext_ptrs, sizeof(void*) * n_slots);
  h.val(w_compute); h.val(p_f32); h.val(arena); h.val(arena_bytes); h.val(workspace); h.val(ws_bytes); h.val(dtype);
  hash_segs(h, segs);
  return run_keyed(h.h, static_cast<cudaStream_t>(stream), [&](cudaStream_t st, cudaStream_t lane1) {
    return tape_forward_run(ops, n_ops, slots, n_slots, params, n_params, ext_ptrs, w_compute, p_f32, arena, arena_bytes,
                            workspace, ws_bytes, dtype, segs, st, lane1);
  });
}

int milb200_tape_backward(const milb200_tape_op* ops, int n_ops, const milb200_tape_slot* slots, int n_slots,
                          const milb200_tape_param* params, int n_params, void* const* ext_ptrs, void* const* ext_grad_ptrs,
                          const void* const* seed_ptrs, const void* w_compute, const float* p_f32, float* g_f32,
                          const void* arena, size_t arena_bytes, void* workspace, size_t ws_bytes, int dtype,
                          const milb200_segments* segs, void* stream) {
  MIL_CHECK_ARG(ops && slots && n_ops > 0 && n_slots > 0 && ext_ptrs && ext_grad_ptrs && seed_ptrs, MILB200_EINVAL,
                "tape_backward: null pointer");
  MIL_CHECK_ARG(!segs || segs->n_segs == 0 || (segs->seg && segs->n_segs > 0 && segs->n_segs <= MILB200_MAX_SEGMENTS),
                MILB200_EINVAL, "tape_backward: bad segment table");
  Hasher h;
  h.val(static_cast<int>(2));
  h.bytes(ops, sizeof(milb200_tape_op) * n_ops);
  h.bytes(slots, sizeof(milb200_tape_slot) * n_slots);
  if (params && n_params > 0) h.bytes(params, sizeof(milb200_tape_param) * n_params);
  h.bytes(ext_ptrs, sizeof(void*) * n_slots);
  h.bytes(ext_grad_ptrs, sizeof(void*) * n_slots);
  h.bytes(seed_ptrs, sizeof(void*) * n_slots);
  h.val(w_compute); h.val(p_f32); h.val(g_f32); h.val(arena); h.val(arena_bytes); h.val(workspace); h.val(ws_bytes);
  h.val(dtype);
  hash_segs(h, segs);
  return run_keyed(h.h, static_cast<cudaStream_t>(stream), [&](cudaStream_t st, cudaStream_t lane1) {
    return tape_backward_run(ops, n_ops, slots, n_slots, params, n_params, ext_ptrs, ext_grad_ptrs, seed_ptrs, w_compute,
                             p_f32, g_f32, arena, arena_bytes, workspace, ws_bytes, dtype, segs, st, lane1);
  });
}

}  // extern "C"
