// Host side of the upstream feeder (reference: dataset.py:366-393).
//
// The reference loads one slide's patch-feature matrix per sample (`np.load`, [n, 768] float), optionally keeps a
// sorted random subset of its rows (augmentation, dataset.py:375-381), zero-pads it to 15 592 rows when the batch
// holds more than one bag (dataset.py:383-390) and converts to fp32 torch tensors.  Here the bags of a step are
// gathered ONCE into the packed-CSR layout the kernels read — rows back to back in pinned host memory, already in
// the compute dtype (bf16, round-to-nearest-even) — so the copy engine moves half the bytes and nothing is padded.
// Plain C++ threads; no CUDA calls (the caller owns the pinned buffer and the H2D copy).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

inline uint16_t f32_to_bf16_rne(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fc0u;  // NaN, as c10::BFloat16's rounding does
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

inline float f16_to_f32(uint16_t h) {
  uint32_t sign = static_cast<uint32_t>(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu, man = h & 0x3ffu;
  uint32_t u;
  if (exp == 0) {
    if (man == 0) {
      u = sign;
    } else {  // subnormal half: renormalise
      int e = -1;
      do {
        man <<= 1;
        ++e;
      } while (!(man & 0x400u));
      u = sign | (static_cast<uint32_t>(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
    }
  } else if (exp == 31) {
    u = sign | 0x7f800000u | (man << 13);
  } else {
    u = sign | ((exp + 127 - 15) << 23) | (man << 13);
  }
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

inline float bf16_to_f32(uint16_t h) {
  uint32_t u = static_cast<uint32_t>(h) << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

template <int SRC>
inline float load_src(const void* row, int j) {
  if (SRC == MILB200_HOST_F32) return static_cast<const float*>(row)[j];
  if (SRC == MILB200_HOST_F64) return static_cast<float>(static_cast<const double*>(row)[j]);
  if (SRC == MILB200_HOST_F16) return f16_to_f32(static_cast<const uint16_t*>(row)[j]);
  return bf16_to_f32(static_cast<const uint16_t*>(row)[j]);
}

constexpr int src_bytes(int src) {
  return src == MILB200_HOST_F64 ? 8 : (src == MILB200_HOST_F32 ? 4 : 2);
}

template <int SRC, bool TO_BF16>
void convert_row(const void* src, void* dst, int L) {
  if (SRC == MILB200_HOST_F32 && !TO_BF16) {
    std::memcpy(dst, src, static_cast<size_t>(L) * 4);
    return;
  }
  if (SRC == MILB200_HOST_BF16 && TO_BF16) {
    std::memcpy(dst, src, static_cast<size_t>(L) * 2);
    return;
  }
  if (TO_BF16 && SRC == MILB200_HOST_F32) {   // the common case (np.float32 slides): branch-free so the compiler vectorises it
    const uint32_t* __restrict__ u32 = static_cast<const uint32_t*>(src);
    uint16_t* __restrict__ d = static_cast<uint16_t*>(dst);
    for (int j = 0; j < L; ++j) {
      const uint32_t u = u32[j];
      const uint32_t r = (u + 0x7fffu + ((u >> 16) & 1u)) >> 16;
      d[j] = static_cast<uint16_t>((u & 0x7fffffffu) > 0x7f800000u ? 0x7fc0u : r);
    }
  } else if (TO_BF16) {
    uint16_t* d = static_cast<uint16_t*>(dst);
    for (int j = 0; j < L; ++j) d[j] = f32_to_bf16_rne(load_src<SRC>(src, j));
  } else {
    float* d = static_cast<float*>(dst);
    for (int j = 0; j < L; ++j) d[j] = load_src<SRC>(src, j);
  }
}

struct PackJob {
  const void* const* bag_ptrs;
  const int64_t* bag_rows;
  const int64_t* bag_pitch;  // bytes between rows, or null
  const int32_t* const* keep_rows;
  const int64_t* keep_counts;
  const int32_t* offsets;
  int n_bags, L;
  void* dst;
};

// rows [r0, r1) of the PACKED batch
template <int SRC, bool TO_BF16>
void pack_range(const PackJob& j, int64_t r0, int64_t r1) {
  const size_t dst_row = static_cast<size_t>(j.L) * (TO_BF16 ? 2 : 4);
  int b = static_cast<int>(std::upper_bound(j.offsets, j.offsets + j.n_bags + 1, static_cast<int32_t>(r0)) - j.offsets) - 1;
  for (int64_t r = r0; r < r1;) {
    while (r >= j.offsets[b + 1]) ++b;
    const int64_t end = std::min<int64_t>(r1, j.offsets[b + 1]);
    const size_t pitch = j.bag_pitch ? static_cast<size_t>(j.bag_pitch[b]) : static_cast<size_t>(j.L) * src_bytes(SRC);
    const uint8_t* base = static_cast<const uint8_t*>(j.bag_ptrs[b]);
    const int32_t* keep = (j.keep_rows && j.keep_rows[b]) ? j.keep_rows[b] : nullptr;
    for (; r < end; ++r) {
      const int64_t local = r - j.offsets[b];
      const int64_t srow = keep ? keep[local] : local;
      convert_row<SRC, TO_BF16>(base + static_cast<size_t>(srow) * pitch,
                                static_cast<uint8_t*>(j.dst) + static_cast<size_t>(r) * dst_row, j.L);
    }
  }
}

using RangeFn = void (*)(const PackJob&, int64_t, int64_t);

RangeFn pick(int src, bool to_bf16) {
  switch (src) {
    case MILB200_HOST_F32: return to_bf16 ? pack_range<MILB200_HOST_F32, true> : pack_range<MILB200_HOST_F32, false>;
    case MILB200_HOST_F64: return to_bf16 ? pack_range<MILB200_HOST_F64, true> : pack_range<MILB200_HOST_F64, false>;
    case MILB200_HOST_F16: return to_bf16 ? pack_range<MILB200_HOST_F16, true> : pack_range<MILB200_HOST_F16, false>;
    case MILB200_HOST_BF16: return to_bf16 ? pack_range<MILB200_HOST_BF16, true> : pack_range<MILB200_HOST_BF16, false>;
  }
  return nullptr;
}

}  // namespace

extern "C" {

int milb200_pack_bags_offsets(const int64_t* bag_rows, const int64_t* keep_counts, int n_bags, int32_t* offsets,
                              int64_t* total_rows) {
  MIL_CHECK_ARG(bag_rows && offsets && n_bags >= 0, MILB200_EINVAL, "pack_bags_offsets: bad arguments");
  int64_t acc = 0;
  offsets[0] = 0;
  for (int b = 0; b < n_bags; ++b) {
    const int64_t n = keep_counts ? keep_counts[b] : bag_rows[b];
    MIL_CHECK_ARG(n >= 1 && n <= bag_rows[b], MILB200_EINVAL, "pack_bags_offsets: every bag needs 1..rows instances");
    acc += n;
    MIL_CHECK_ARG(acc <= INT32_MAX, MILB200_EINVAL, "pack_bags_offsets: more than 2^31-1 instances in one batch");
    offsets[b + 1] = static_cast<int32_t>(acc);
  }
  if (total_rows) *total_rows = acc;
  return MILB200_OK;
}

int milb200_pack_bags_host(const void* const* bag_ptrs, const int64_t* bag_rows, const int64_t* bag_pitch_bytes,
                           const int32_t* const* keep_rows, const int64_t* keep_counts, int n_bags, int L,
                           int src_dtype, void* dst, int dst_dtype, int64_t dst_capacity_rows, int32_t* offsets,
                           int n_threads) {
  MIL_CHECK_ARG(bag_ptrs && bag_rows && dst && offsets && n_bags >= 1 && L >= 1, MILB200_EINVAL,
                "pack_bags_host: bad arguments");
  MIL_CHECK_ARG(dst_dtype == MILB200_F32 || dst_dtype == MILB200_BF16, MILB200_EUNSUPPORTED,
                "pack_bags_host: destination must be fp32 or bf16");
  MIL_CHECK_ARG((keep_rows == nullptr) == (keep_counts == nullptr), MILB200_EINVAL,
                "pack_bags_host: keep_rows and keep_counts go together");
  RangeFn fn = pick(src_dtype, dst_dtype == MILB200_BF16);
  MIL_CHECK_ARG(fn != nullptr, MILB200_EUNSUPPORTED, "pack_bags_host: source dtype must be a MILB200_HOST_* code");
  int64_t total = 0;
  int rc = milb200_pack_bags_offsets(bag_rows, keep_counts, n_bags, offsets, &total);
  if (rc != MILB200_OK) return rc;
  MIL_CHECK_ARG(total <= dst_capacity_rows, MILB200_EWORKSPACE, "pack_bags_host: destination holds fewer rows than the batch");
  for (int b = 0; b < n_bags; ++b) {
    MIL_CHECK_ARG(bag_ptrs[b] != nullptr, MILB200_EINVAL, "pack_bags_host: null bag pointer");
    if (keep_rows && keep_rows[b]) {   // sorted, in range (the reference samples `sorted(random.sample(range(n), k))`)
      const int32_t* k = keep_rows[b];
      for (int64_t i = 0; i < keep_counts[b]; ++i)
        MIL_CHECK_ARG(k[i] >= 0 && k[i] < bag_rows[b] && (i == 0 || k[i] > k[i - 1]), MILB200_EINVAL,
                      "pack_bags_host: keep_rows must be strictly increasing row indices of the bag");
    } else if (keep_counts) {
      MIL_CHECK_ARG(keep_counts[b] == bag_rows[b], MILB200_EINVAL, "pack_bags_host: keep_counts without keep_rows must equal rows");
    }
  }
  PackJob job{bag_ptrs, bag_rows, bag_pitch_bytes, keep_rows, keep_counts, offsets, n_bags, L, dst};
  int nt = n_threads > 0 ? n_threads : static_cast<int>(std::thread::hardware_concurrency());
  nt = std::max(1, std::min<int>(nt, static_cast<int>(std::min<int64_t>(total, 256))));
  if (nt == 1) {
    fn(job, 0, total);
    return MILB200_OK;
  }
  // contiguous row ranges per thread: every thread streams its own slice of the destination
  std::vector<std::thread> pool;
  pool.reserve(nt);
  const int64_t per = (total + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    const int64_t r0 = std::min<int64_t>(total, per * t), r1 = std::min<int64_t>(total, per * (t + 1));
    if (r0 < r1) pool.emplace_back([=, &job] { fn(job, r0, r1); });
  }
  for (auto& th : pool) th.join();
  return MILB200_OK;
}

}  // extern "C"
