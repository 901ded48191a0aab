// elementwise.cu — Map / Binary operator templates as coalesced 256-bit column sweeps.
//
// One launch covers every column of a block (Modin partition): a persistent grid of
// (SM count x resident CTAs) walks (column, row-tile) pairs.  Each thread moves 4 x 256-bit
// vectors per operand per tile, all loads issued before the first use (LDG.E.256, L1 bypass,
// L2 evict-first), results stored with STG.E.256.  The path is HBM-bound: algorithmic traffic
// is 8 B read per operand element + 8 B written (1 B for predicates).
//
// pandas semantics restated here (reference call sites):
//   abs/neg/isna/notna  Map.register(pandas.DataFrame.abs ...)   qc.py:2036, 2063-2106
//   fillna(scalar)      qc.fillna -> frame.map                    qc.py:2710-2813
//   a OP b, a OP s      Binary.register(pandas.DataFrame.add ...) qc.py:535-624
//   a*b+c               two Binary passes in the reference (alg/binary.py:420-430) -- fused here
//                       but with TWO IEEE roundings (__dmul_rn then __dadd_rn), never an FMA,
//                       so results are bit-identical to pandas.
#include <type_traits>

#include "common.cuh"

namespace mb200 {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;
constexpr int kVec = 4;
constexpr int kTile = kThreads * kUnroll * kVec;  // 4096 elements = 32 KiB of f64 per operand

struct MapParams {
  const void* in0[MB200_MAX_COLS];
  const void* in1[MB200_MAX_COLS];
  const void* in2[MB200_MAX_COLS];
  void* out[MB200_MAX_COLS];
  uint64_t s0[MB200_MAX_COLS];
  uint64_t s1[MB200_MAX_COLS];
  int ncols;
  long long nrows;
  long long tiles_per_col;
};

template <typename T>
__device__ __forceinline__ T from_bits(uint64_t b);
template <>
__device__ __forceinline__ double from_bits<double>(uint64_t b) {
  return __longlong_as_double((long long)b);
}
template <>
__device__ __forceinline__ long long from_bits<long long>(uint64_t b) {
  return (long long)b;
}
template <>
__device__ __forceinline__ uint8_t from_bits<uint8_t>(uint64_t b) {
  return (uint8_t)(b != 0);
}

constexpr __host__ __device__ int op_nin(int op) { return op >= 64 ? 3 : (op >= 32 ? 2 : 1); }
constexpr __host__ __device__ bool op_is_pred(int op) {
  return op == MB200_OP_ISNA || op == MB200_OP_NOTNA || (op >= MB200_OP_EQ_S && op <= MB200_OP_GE_S) ||
         (op >= MB200_OP_EQ && op <= MB200_OP_GE);
}

// ---- the arithmetic.  TI in {double, long long}; TO in {double, long long, uint8_t}.
template <int OP, typename TI, typename TO>
__device__ __forceinline__ TO apply(TI a, TI b, TI c, TI s0, TI s1) {
  constexpr bool F = std::is_same<TI, double>::value;  // floating input?
  if constexpr (OP == MB200_OP_ABS) {
    if constexpr (F) return (TO)fabs(a);
    else return (TO)(a < 0 ? (TI)(0ULL - (unsigned long long)a) : a);
  } else if constexpr (OP == MB200_OP_NEG) {
    // sign-bit flip (like numpy/x86 xorpd): `-a` as arithmetic would not flip the sign of a NaN
    if constexpr (F) return (TO)__longlong_as_double(__double_as_longlong(a) ^ (long long)0x8000000000000000ULL);
    else return (TO)(TI)(0ULL - (unsigned long long)a);
  } else if constexpr (OP == MB200_OP_ISNA) {
    return (TO)(a != a);
  } else if constexpr (OP == MB200_OP_NOTNA) {
    return (TO)(a == a);
  } else if constexpr (OP == MB200_OP_FILLNA_S) {
    return (TO)((a != a) ? s0 : a);
  } else if constexpr (OP == MB200_OP_AFFINE) {
    if constexpr (F) return (TO)__dadd_rn(__dmul_rn(a, s0), s1);
    else return (TO)(TI)((unsigned long long)a * (unsigned long long)s0 + (unsigned long long)s1);
  } else if constexpr (OP == MB200_OP_ADD_S || OP == MB200_OP_ADD) {
    TI r = (OP == MB200_OP_ADD) ? b : s0;
    if constexpr (F) return (TO)__dadd_rn(a, r);
    else return (TO)(TI)((unsigned long long)a + (unsigned long long)r);
  } else if constexpr (OP == MB200_OP_SUB_S || OP == MB200_OP_SUB) {
    TI r = (OP == MB200_OP_SUB) ? b : s0;
    if constexpr (F) return (TO)__dsub_rn(a, r);
    else return (TO)(TI)((unsigned long long)a - (unsigned long long)r);
  } else if constexpr (OP == MB200_OP_RSUB_S) {
    if constexpr (F) return (TO)__dsub_rn(s0, a);
    else return (TO)(TI)((unsigned long long)s0 - (unsigned long long)a);
  } else if constexpr (OP == MB200_OP_MUL_S || OP == MB200_OP_MUL) {
    TI r = (OP == MB200_OP_MUL) ? b : s0;
    if constexpr (F) return (TO)__dmul_rn(a, r);
    else return (TO)(TI)((unsigned long long)a * (unsigned long long)r);
  } else if constexpr (OP == MB200_OP_DIV_S || OP == MB200_OP_DIV) {
    TI r = (OP == MB200_OP_DIV) ? b : s0;
    return (TO)__ddiv_rn((double)a, (double)r);  // true division: int64 inputs promote to f64
  } else if constexpr (OP == MB200_OP_RDIV_S) {
    return (TO)__ddiv_rn((double)s0, (double)a);
  } else if constexpr (OP == MB200_OP_EQ_S || OP == MB200_OP_EQ) {
    return (TO)(a == ((OP == MB200_OP_EQ) ? b : s0));
  } else if constexpr (OP == MB200_OP_NE_S || OP == MB200_OP_NE) {
    return (TO)(a != ((OP == MB200_OP_NE) ? b : s0));
  } else if constexpr (OP == MB200_OP_LT_S || OP == MB200_OP_LT) {
    return (TO)(a < ((OP == MB200_OP_LT) ? b : s0));
  } else if constexpr (OP == MB200_OP_LE_S || OP == MB200_OP_LE) {
    return (TO)(a <= ((OP == MB200_OP_LE) ? b : s0));
  } else if constexpr (OP == MB200_OP_GT_S || OP == MB200_OP_GT) {
    return (TO)(a > ((OP == MB200_OP_GT) ? b : s0));
  } else if constexpr (OP == MB200_OP_GE_S || OP == MB200_OP_GE) {
    return (TO)(a >= ((OP == MB200_OP_GE) ? b : s0));
  } else if constexpr (OP == MB200_OP_CLIP_S) {
    // pandas clip = where(x >= lower, lower) / where(x <= upper, upper) with NaN kept: comparisons, not
    // fmin / fmax (which would turn a -0.0 at lower = 0.0 into +0.0); absent bounds are passed as -inf / +inf
    return (TO)(a < s0 ? s0 : (a > s1 ? s1 : a));
  } else if constexpr (OP == MB200_OP_ROUND_S) {
    // numpy.round(x, d) (pandas DataFrame.round): d >= 0: rint(x * 10^d) / 10^d, d < 0: rint(x / 10^-d) * 10^-d;
    // s0 = 10^|d| (exact in float64 for |d| <= 22), s1 = sign of d.  rint = round-half-to-even.  Like numpy,
    // no guard against the scaled value overflowing (round(1e308, 2) = inf).
    if constexpr (F) {
      if (s1 >= (TI)0) return (TO)__ddiv_rn(rint(__dmul_rn(a, s0)), s0);
      return (TO)__dmul_rn(rint(__ddiv_rn(a, s0)), s0);
    } else {
      return (TO)a;  // integers: decimals >= 0 is the identity (negative decimals are not on this path)
    }
  } else if constexpr (OP == MB200_OP_NOT) {
    return (TO)(a == (TI)0);
  } else if constexpr (OP == MB200_OP_AND) {
    return (TO)((a != (TI)0) && (b != (TI)0));
  } else if constexpr (OP == MB200_OP_OR) {
    return (TO)((a != (TI)0) || (b != (TI)0));
  } else if constexpr (OP == MB200_OP_XOR) {
    return (TO)((a != (TI)0) != (b != (TI)0));
  } else if constexpr (OP == MB200_OP_ORDERED_S) {
    // sort key: an int64 whose signed order is the order sort_values wants.  float64: flip the magnitude bits
    // of negatives (total order of IEEE doubles), NaN -> INT64_MAX (na_position="last" in either direction);
    // s0 != 0 = descending: bitwise NOT reverses the order without overflow and keeps ties stable.
    long long o;
    if constexpr (F) {
      if (a != a) return (TO)0x7fffffffffffffffLL;
      const long long b = __double_as_longlong(a);
      o = b ^ ((b >> 63) & 0x7fffffffffffffffLL);
    } else {
      o = (long long)a;
    }
    return (TO)(s0 != (TI)0 ? ~o : o);
  } else if constexpr (OP == MB200_OP_COPY) {
    return (TO)a;
  } else if constexpr (OP == MB200_OP_FILLNA) {
    return (TO)((a != a) ? b : a);
  } else if constexpr (OP == MB200_OP_FMA3) {
    if constexpr (F) return (TO)__dadd_rn(__dmul_rn(a, b), c);
    else return (TO)(TI)((unsigned long long)a * (unsigned long long)b + (unsigned long long)c);
  } else {
    return (TO)a;
  }
}

template <typename T>
struct Vec4;
template <>
struct Vec4<double> {
  using type = f64x4;
  static __device__ __forceinline__ f64x4 load(const double* p) { return ldg_stream_f64x4(p); }
  static __device__ __forceinline__ void store(double* p, const f64x4& v) { stg_stream_f64x4(p, v); }
};
template <>
struct Vec4<long long> {
  using type = i64x4;
  static __device__ __forceinline__ i64x4 load(const long long* p) { return ldg_stream_i64x4(p); }
  static __device__ __forceinline__ void store(long long* p, const i64x4& v) { stg_stream_i64x4(p, v); }
};

// bool columns (uint8 0 / 1): four elements = one 32-bit word
struct u8x4 {
  uint8_t x, y, z, w;
};
template <>
struct Vec4<uint8_t> {
  using type = u8x4;
  static __device__ __forceinline__ u8x4 load(const uint8_t* p) {
    const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(p));
    u8x4 r;
    r.x = (uint8_t)(v & 0xffu);
    r.y = (uint8_t)((v >> 8) & 0xffu);
    r.z = (uint8_t)((v >> 16) & 0xffu);
    r.w = (uint8_t)(v >> 24);
    return r;
  }
};

template <typename TO, typename VO>
__device__ __forceinline__ void store_out(TO* p, const VO& v) {
  if constexpr (sizeof(TO) == 1) {
    uint32_t packed = (uint32_t)v.x | ((uint32_t)v.y << 8) | ((uint32_t)v.z << 16) | ((uint32_t)v.w << 24);
    *reinterpret_cast<uint32_t*>(p) = packed;
  } else {
    Vec4<TO>::store(p, v);
  }
}
template <typename TO>
struct OutVec {
  TO x, y, z, w;
};
template <>
struct OutVec<double> : f64x4 {};
template <>
struct OutVec<long long> : i64x4 {};

// VEC=true: every operand pointer is 32-byte aligned -> 256-bit path; else scalar sweep.
template <int OP, typename TI, typename TO, bool VEC>
__global__ void __launch_bounds__(kThreads) map_kernel(const __grid_constant__ MapParams p) {
  constexpr int NIN = op_nin(OP);
  const long long ntiles = p.tiles_per_col * p.ncols;
  const int tid = threadIdx.x;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int col = (int)(t / p.tiles_per_col);
    const long long base = (t - (long long)col * p.tiles_per_col) * kTile;
    const TI* __restrict__ a = static_cast<const TI*>(p.in0[col]);
    const TI* __restrict__ b = NIN >= 2 ? static_cast<const TI*>(p.in1[col]) : nullptr;
    const TI* __restrict__ c = NIN >= 3 ? static_cast<const TI*>(p.in2[col]) : nullptr;
    TO* __restrict__ o = static_cast<TO*>(p.out[col]);
    const TI s0 = from_bits<TI>(p.s0[col]);
    const TI s1 = from_bits<TI>(p.s1[col]);
    if (VEC && base + kTile <= p.nrows) {
      using V = typename Vec4<TI>::type;
      V va[kUnroll], vb[kUnroll], vc[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long i = base + (long long)(u * kThreads + tid) * kVec;
        va[u] = Vec4<TI>::load(a + i);
        if constexpr (NIN >= 2) vb[u] = Vec4<TI>::load(b + i);
        if constexpr (NIN >= 3) vc[u] = Vec4<TI>::load(c + i);
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const long long i = base + (long long)(u * kThreads + tid) * kVec;
        OutVec<TO> r;
        if constexpr (NIN == 1) {
          r.x = apply<OP, TI, TO>(va[u].x, 0, 0, s0, s1);
          r.y = apply<OP, TI, TO>(va[u].y, 0, 0, s0, s1);
          r.z = apply<OP, TI, TO>(va[u].z, 0, 0, s0, s1);
          r.w = apply<OP, TI, TO>(va[u].w, 0, 0, s0, s1);
        } else if constexpr (NIN == 2) {
          r.x = apply<OP, TI, TO>(va[u].x, vb[u].x, 0, s0, s1);
          r.y = apply<OP, TI, TO>(va[u].y, vb[u].y, 0, s0, s1);
          r.z = apply<OP, TI, TO>(va[u].z, vb[u].z, 0, s0, s1);
          r.w = apply<OP, TI, TO>(va[u].w, vb[u].w, 0, s0, s1);
        } else {
          r.x = apply<OP, TI, TO>(va[u].x, vb[u].x, vc[u].x, s0, s1);
          r.y = apply<OP, TI, TO>(va[u].y, vb[u].y, vc[u].y, s0, s1);
          r.z = apply<OP, TI, TO>(va[u].z, vb[u].z, vc[u].z, s0, s1);
          r.w = apply<OP, TI, TO>(va[u].w, vb[u].w, vc[u].w, s0, s1);
        }
        store_out<TO>(o + i, r);
      }
    } else {
      const long long end = (base + kTile < p.nrows) ? base + kTile : p.nrows;
      for (long long i = base + tid; i < end; i += kThreads) {
        TI x = a[i];
        TI y = NIN >= 2 ? b[i] : (TI)0;
        TI z = NIN >= 3 ? c[i] : (TI)0;
        o[i] = apply<OP, TI, TO>(x, y, z, s0, s1);
      }
    }
  }
}

template <int OP, typename TI, typename TO>
static int launch_map(const MapParams& p, bool vec, int grid, cudaStream_t st) {
  if (vec)
    map_kernel<OP, TI, TO, true><<<grid, kThreads, 0, st>>>(p);
  else
    map_kernel<OP, TI, TO, false><<<grid, kThreads, 0, st>>>(p);
  MB_LAUNCH_CHECK("map_kernel");
  return 0;
}

template <int OP, typename TI, typename TO>
static int grid_for(int* grid, long long ntiles) {
  DevProps dp;
  if (int rc = dev_props(&dp)) return rc;
  int occ = 0;
  MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, map_kernel<OP, TI, TO, true>, kThreads, 0));
  if (occ < 1) occ = 1;
  long long g = (long long)dp.sm_count * occ;  // one full wave of resident CTAs, persistent loop
  if (g > ntiles) g = ntiles;
  if (g < 1) g = 1;
  *grid = (int)g;
  return 0;
}

#define MB_CASE(OPC, TI, TO)                                              \
  case OPC: {                                                             \
    int grid;                                                             \
    if (int rc = grid_for<OPC, TI, TO>(&grid, ntiles)) return rc;         \
    return launch_map<OPC, TI, TO>(p, vec, grid, st);                     \
  }

}  // namespace mb200

using namespace mb200;

extern "C" int mb200_map(int op, int dtype, int ncols, const void* const* in0, const void* const* in1,
                         const void* const* in2, void* const* out, int64_t nrows, const uint64_t* s0,
                         const uint64_t* s1, mb200_stream_t stream) {
  if (ncols < 0 || ncols > MB200_MAX_COLS) return fail("mb200_map", "ncols out of range (0..32)");
  if (nrows < 0) return fail("mb200_map", "negative nrows");
  if (ncols == 0 || nrows == 0) return 0;
  if (!in0 || !out) return fail("mb200_map", "null column array");
  const int nin = op_nin(op);
  if (nin >= 2 && !in1) return fail("mb200_map", "binary op needs in1");
  if (nin >= 3 && !in2) return fail("mb200_map", "ternary op needs in2");
  MapParams p;
  memset(&p, 0, sizeof(p));
  bool vec = true;
  for (int c = 0; c < ncols; ++c) {
    p.in0[c] = in0[c];
    p.in1[c] = nin >= 2 ? in1[c] : nullptr;
    p.in2[c] = nin >= 3 ? in2[c] : nullptr;
    p.out[c] = out[c];
    p.s0[c] = s0 ? s0[c] : 0;
    p.s1[c] = s1 ? s1[c] : 0;
    if (!in0[c] || !out[c]) return fail("mb200_map", "null column pointer");
    vec = vec && aligned32(in0[c]) && aligned32(out[c]);
    if (nin >= 2) vec = vec && in1[c] && aligned32(in1[c]);
    if (nin >= 3) vec = vec && in2[c] && aligned32(in2[c]);
    if ((nin >= 2 && !in1[c]) || (nin >= 3 && !in2[c])) return fail("mb200_map", "null column pointer");
  }
  p.ncols = ncols;
  p.nrows = nrows;
  p.tiles_per_col = (nrows + kTile - 1) / kTile;
  const long long ntiles = p.tiles_per_col * ncols;
  cudaStream_t st = (cudaStream_t)stream;

  if (dtype == MB200_F64) {
    switch (op) {
      MB_CASE(MB200_OP_ABS, double, double)
      MB_CASE(MB200_OP_NEG, double, double)
      MB_CASE(MB200_OP_ISNA, double, uint8_t)
      MB_CASE(MB200_OP_NOTNA, double, uint8_t)
      MB_CASE(MB200_OP_FILLNA_S, double, double)
      MB_CASE(MB200_OP_AFFINE, double, double)
      MB_CASE(MB200_OP_ADD_S, double, double)
      MB_CASE(MB200_OP_SUB_S, double, double)
      MB_CASE(MB200_OP_RSUB_S, double, double)
      MB_CASE(MB200_OP_MUL_S, double, double)
      MB_CASE(MB200_OP_DIV_S, double, double)
      MB_CASE(MB200_OP_RDIV_S, double, double)
      MB_CASE(MB200_OP_EQ_S, double, uint8_t)
      MB_CASE(MB200_OP_NE_S, double, uint8_t)
      MB_CASE(MB200_OP_LT_S, double, uint8_t)
      MB_CASE(MB200_OP_LE_S, double, uint8_t)
      MB_CASE(MB200_OP_GT_S, double, uint8_t)
      MB_CASE(MB200_OP_GE_S, double, uint8_t)
      MB_CASE(MB200_OP_CLIP_S, double, double)
      MB_CASE(MB200_OP_COPY, double, double)
      MB_CASE(MB200_OP_ROUND_S, double, double)
      MB_CASE(MB200_OP_ORDERED_S, double, long long)
      MB_CASE(MB200_OP_ADD, double, double)
      MB_CASE(MB200_OP_SUB, double, double)
      MB_CASE(MB200_OP_MUL, double, double)
      MB_CASE(MB200_OP_DIV, double, double)
      MB_CASE(MB200_OP_EQ, double, uint8_t)
      MB_CASE(MB200_OP_NE, double, uint8_t)
      MB_CASE(MB200_OP_LT, double, uint8_t)
      MB_CASE(MB200_OP_LE, double, uint8_t)
      MB_CASE(MB200_OP_GT, double, uint8_t)
      MB_CASE(MB200_OP_GE, double, uint8_t)
      MB_CASE(MB200_OP_FILLNA, double, double)
      MB_CASE(MB200_OP_FMA3, double, double)
      default:
        return fail("mb200_map", "unsupported op for float64");
    }
  } else if (dtype == MB200_I64) {
    switch (op) {
      MB_CASE(MB200_OP_ABS, long long, long long)
      MB_CASE(MB200_OP_NEG, long long, long long)
      MB_CASE(MB200_OP_AFFINE, long long, long long)
      MB_CASE(MB200_OP_ADD_S, long long, long long)
      MB_CASE(MB200_OP_SUB_S, long long, long long)
      MB_CASE(MB200_OP_RSUB_S, long long, long long)
      MB_CASE(MB200_OP_MUL_S, long long, long long)
      MB_CASE(MB200_OP_DIV_S, long long, double)
      MB_CASE(MB200_OP_RDIV_S, long long, double)
      MB_CASE(MB200_OP_EQ_S, long long, uint8_t)
      MB_CASE(MB200_OP_NE_S, long long, uint8_t)
      MB_CASE(MB200_OP_LT_S, long long, uint8_t)
      MB_CASE(MB200_OP_LE_S, long long, uint8_t)
      MB_CASE(MB200_OP_GT_S, long long, uint8_t)
      MB_CASE(MB200_OP_GE_S, long long, uint8_t)
      MB_CASE(MB200_OP_CLIP_S, long long, long long)
      MB_CASE(MB200_OP_COPY, long long, long long)
      MB_CASE(MB200_OP_ROUND_S, long long, long long)
      MB_CASE(MB200_OP_ORDERED_S, long long, long long)
      MB_CASE(MB200_OP_ADD, long long, long long)
      MB_CASE(MB200_OP_SUB, long long, long long)
      MB_CASE(MB200_OP_MUL, long long, long long)
      MB_CASE(MB200_OP_DIV, long long, double)
      MB_CASE(MB200_OP_EQ, long long, uint8_t)
      MB_CASE(MB200_OP_NE, long long, uint8_t)
      MB_CASE(MB200_OP_LT, long long, uint8_t)
      MB_CASE(MB200_OP_LE, long long, uint8_t)
      MB_CASE(MB200_OP_GT, long long, uint8_t)
      MB_CASE(MB200_OP_GE, long long, uint8_t)
      MB_CASE(MB200_OP_FMA3, long long, long long)
      default:
        return fail("mb200_map", "unsupported op for int64");
    }
  } else if (dtype == MB200_U8) {
    // bool columns: logical ops stay bool; COPY widens to int64 (what a reduction over booleans consumes)
    switch (op) {
      MB_CASE(MB200_OP_COPY, uint8_t, long long)
      MB_CASE(MB200_OP_NOT, uint8_t, uint8_t)
      MB_CASE(MB200_OP_AND, uint8_t, uint8_t)
      MB_CASE(MB200_OP_OR, uint8_t, uint8_t)
      MB_CASE(MB200_OP_XOR, uint8_t, uint8_t)
      default:
        return fail("mb200_map", "unsupported op for bool");
    }
  }
  return fail("mb200_map", "unsupported dtype");
}
