// synth.cu — counter-based synthetic column generators (from_map-style on-device ingest,
// cf. BaseIO.from_map modin/core/io/io.py:184-209) and the L2 flush used between timed runs.
//
// value(seed, col, row) is a pure function, so any row range of any partition on any GPU is
// reproducible; modin_b200/synth.py holds the bit-identical numpy twin used by the tests.
//   z1 = mix64(seed * K1 + col * K2 + row + 1),  z2 = mix64(z1 + K3)
//   f64: four exact uniforms u = (32-bit piece) * 2^-32 ; x = ((u0+u1)+(u2+u3) - 2) * sqrt(3)
//        (Irwin-Hall(4), mean 0, variance 1; every operation is a single IEEE rounding, so
//        numpy reproduces it bit for bit);  NaN when (z2 >> 48) < nan_per_64k... uses z1 low bits.
//   i64: ((z1 >> 32) * modulus) >> 32   (uniform in [0, modulus), modulus < 2^32)
#include "common.cuh"

namespace mb200 {

constexpr unsigned long long K1 = 0x9E3779B97F4A7C15ULL;
constexpr unsigned long long K2 = 0xD1B54A32D192ED03ULL;
constexpr unsigned long long K3 = 0x8CB92BA72F3D8DD7ULL;

__device__ __forceinline__ double synth_f64(unsigned long long seed, unsigned long long col, long long row,
                                            int nan_per_64k) {
  const unsigned long long z1 = mix64(seed * K1 + col * K2 + (unsigned long long)row + 1ULL);
  const unsigned long long z2 = mix64(z1 + K3);
  const double s = 2.3283064365386963e-10;  // 2^-32
  const double u0 = (double)(unsigned int)(z1 >> 32) * s;
  const double u1 = (double)(unsigned int)(z1 & 0xffffffffULL) * s;
  const double u2 = (double)(unsigned int)(z2 >> 32) * s;
  const double u3 = (double)(unsigned int)(z2 & 0xffffffffULL) * s;
  const double x = __dmul_rn(__dsub_rn(__dadd_rn(__dadd_rn(u0, u1), __dadd_rn(u2, u3)), 2.0), 1.7320508075688772);
  const unsigned long long z3 = mix64(z2 + K3);
  if ((int)(z3 & 0xffffULL) < nan_per_64k) return __longlong_as_double(0x7ff8000000000000LL);
  return x;
}

__global__ void gen_f64_kernel(double* __restrict__ out, long long n, unsigned long long seed,
                               unsigned long long col, long long row_offset, int nan_per_64k) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = synth_f64(seed, col, row_offset + i, nan_per_64k);
}

// STATS: the generator also leaves the column's key statistics {min, max, sampled, duplicated} in stats[4]
// (KeyStatsAcc, common.cuh) -- a groupby on a generated key column then needs no pre-pass over it.
template <bool STATS>
__global__ void gen_i64_kernel(long long* __restrict__ out, long long n, unsigned long long seed,
                               unsigned long long col, long long row_offset, unsigned long long modulus,
                               long long* stats) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  KeyStatsAcc st;
  unsigned int iter = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride, ++iter) {
    const unsigned long long z1 = mix64(seed * K1 + col * K2 + (unsigned long long)(row_offset + i) + 1ULL);
    const long long k = (long long)(((z1 >> 32) * modulus) >> 32);
    out[i] = k;
    if (STATS) {
      st.add(k);
      if ((iter & 15u) == 0 && __activemask() == 0xffffffffu) st.sample_warp(k);
    }
  }
  if (STATS) st.flush(stats);
}

// Skewed keys (the Zipf-like variant of SURVEY 8d): level t in [0, T] is drawn with weight w_t, w_0 = 2^20,
// w_{t+1} = w_t + (w_t >> 4) + (w_t >> 7) (x 1.0703 per level -- integer recurrence, so numpy reproduces it
// exactly), T = floor(log2(modulus)); the key is uniform in [0, modulus >> t).  Mass per octave grows towards
// small keys like k^-1.1: for modulus = 1e6 key 0 takes ~11 % of the rows, the top 256 keys ~60 %.
struct SkewLevels {
  unsigned long long cum[66];  // cum[t] = w_0 + ... + w_{t-1};  cum[T + 1] = total
  int nlevels;                 // T + 1
};

template <bool STATS>
__global__ void gen_i64_skew_kernel(long long* __restrict__ out, long long n, unsigned long long seed,
                                    unsigned long long col, long long row_offset, unsigned long long modulus,
                                    const __grid_constant__ SkewLevels lv, long long* stats) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const unsigned long long total = lv.cum[lv.nlevels];
  KeyStatsAcc st;
  unsigned int iter = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride, ++iter) {
    const unsigned long long z1 = mix64(seed * K1 + col * K2 + (unsigned long long)(row_offset + i) + 1ULL);
    const unsigned long long z2 = mix64(z1 + K3);
    // pick in [0, total): total < 2^31, so (32-bit piece * total) >> 32 is exact in 64 bits
    const unsigned long long pick = ((z2 >> 32) * total) >> 32;
    int t = 0;
    while (t + 1 < lv.nlevels && lv.cum[t + 1] <= pick) ++t;
    unsigned long long range = modulus >> t;
    if (range == 0) range = 1;
    const long long k = (long long)(((z1 >> 32) * range) >> 32);
    out[i] = k;
    if (STATS) {
      st.add(k);
      if ((iter & 15u) == 0 && __activemask() == 0xffffffffu) st.sample_warp(k);
    }
  }
  if (STATS) st.flush(stats);
}

// out[i] = start + i (row labels of a RangeIndex block as a device column) / out[i] = bits (constant column)
__global__ void iota_i64_kernel(long long* __restrict__ out, long long n, long long start) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = start + i;
}
__global__ void fill_u64_kernel(unsigned long long* __restrict__ out, long long n, unsigned long long bits) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = bits;
}

__global__ void flush_kernel(unsigned long long* __restrict__ buf, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) buf[i] = (unsigned long long)i;
}

}  // namespace mb200

using namespace mb200;

extern "C" int mb200_gen_f64(double* out, int64_t nrows, uint64_t seed, uint64_t col, int64_t row_offset,
                             int nan_per_64k, mb200_stream_t stream) {
  if (nrows < 0) return fail("mb200_gen_f64", "negative nrows");
  if (nrows == 0) return 0;
  if (!out) return fail("mb200_gen_f64", "null output");
  DevProps dp;
  if (int rc = dev_props(&dp)) return rc;
  long long grid = (nrows + 255) / 256;
  if (grid > (long long)dp.sm_count * 16) grid = (long long)dp.sm_count * 16;
  gen_f64_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(out, nrows, seed, col, row_offset, nan_per_64k);
  MB_LAUNCH_CHECK("gen_f64_kernel");
  return 0;
}

extern "C" int mb200_gen_i64(int64_t* out, int64_t nrows, uint64_t seed, uint64_t col, int64_t row_offset,
                             uint64_t modulus, int64_t* stats_dev, mb200_stream_t stream) {
  if (nrows < 0) return fail("mb200_gen_i64", "negative nrows");
  if (modulus == 0 || modulus > 0xffffffffULL) return fail("mb200_gen_i64", "modulus must be in [1, 2^32)");
  if (stats_dev)
    if (int rc = key_stats_init(reinterpret_cast<long long*>(stats_dev), (cudaStream_t)stream)) return rc;
  if (nrows == 0) return 0;
  if (!out) return fail("mb200_gen_i64", "null output");
  DevProps dp;
  if (int rc = dev_props(&dp)) return rc;
  long long grid = (nrows + 255) / 256;
  if (grid > (long long)dp.sm_count * 16) grid = (long long)dp.sm_count * 16;
  if (stats_dev)
    gen_i64_kernel<true><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<long long*>(out), nrows, seed, col, row_offset, modulus, reinterpret_cast<long long*>(stats_dev));
  else
    gen_i64_kernel<false><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<long long*>(out), nrows,
                                                                            seed, col, row_offset, modulus, nullptr);
  MB_LAUNCH_CHECK("gen_i64_kernel");
  return 0;
}

extern "C" int mb200_gen_i64_skew(int64_t* out, int64_t nrows, uint64_t seed, uint64_t col, int64_t row_offset,
                                  uint64_t modulus, int64_t* stats_dev, mb200_stream_t stream) {
  if (nrows < 0) return fail("mb200_gen_i64_skew", "negative nrows");
  if (modulus == 0 || modulus > 0xffffffffULL) return fail("mb200_gen_i64_skew", "modulus must be in [1, 2^32)");
  if (stats_dev)
    if (int rc = key_stats_init(reinterpret_cast<long long*>(stats_dev), (cudaStream_t)stream)) return rc;
  if (nrows == 0) return 0;
  if (!out) return fail("mb200_gen_i64_skew", "null output");
  DevProps dp;
  if (int rc = dev_props(&dp)) return rc;
  SkewLevels lv;
  memset(&lv, 0, sizeof(lv));
  int T = 0;
  while ((modulus >> (T + 1)) != 0) ++T;
  unsigned long long w = 1ULL << 20, acc = 0;
  for (int t = 0; t <= T; ++t) {
    lv.cum[t] = acc;
    acc += w;
    w = w + (w >> 4) + (w >> 7);
  }
  lv.cum[T + 1] = acc;
  lv.nlevels = T + 1;
  long long grid = (nrows + 255) / 256;
  if (grid > (long long)dp.sm_count * 16) grid = (long long)dp.sm_count * 16;
  if (stats_dev)
    gen_i64_skew_kernel<true><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<long long*>(out), nrows, seed, col, row_offset, modulus, lv,
        reinterpret_cast<long long*>(stats_dev));
  else
    gen_i64_skew_kernel<false><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<long long*>(out), nrows, seed, col, row_offset, modulus, lv, nullptr);
  MB_LAUNCH_CHECK("gen_i64_skew_kernel");
  return 0;
}

extern "C" int mb200_flush_l2(void* buf, size_t bytes, mb200_stream_t stream) {
  if (!buf || bytes < 8) return fail("mb200_flush_l2", "need a buffer");
  DevProps dp;
  if (int rc = dev_props(&dp)) return rc;
  flush_kernel<<<dp.sm_count * 8, 256, 0, (cudaStream_t)stream>>>(static_cast<unsigned long long*>(buf),
                                                                  (long long)(bytes / 8));
  MB_LAUNCH_CHECK("flush_kernel");
  return 0;
}

extern "C" int mb200_iota_i64(int64_t* out, int64_t nrows, int64_t start, mb200_stream_t stream) {
  if (nrows < 0) return fail("mb200_iota_i64", "negative nrows");
  if (nrows == 0) return 0;
  if (!out) return fail("mb200_iota_i64", "null output");
  DevProps dp;
  if (int rc = dev_props(&dp)) return rc;
  long long grid = (nrows + 255) / 256;
  if (grid > (long long)dp.sm_count * 16) grid = (long long)dp.sm_count * 16;
  iota_i64_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<long long*>(out), nrows, start);
  MB_LAUNCH_CHECK("iota_i64_kernel");
  return 0;
}

extern "C" int mb200_fill_u64(void* out, int64_t n, uint64_t bits, mb200_stream_t stream) {
  if (n < 0) return fail("mb200_fill_u64", "negative n");
  if (n == 0) return 0;
  if (!out) return fail("mb200_fill_u64", "null output");
  DevProps dp;
  if (int rc = dev_props(&dp)) return rc;
  long long grid = (n + 255) / 256;
  if (grid > (long long)dp.sm_count * 16) grid = (long long)dp.sm_count * 16;
  fill_u64_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(static_cast<unsigned long long*>(out), n, bits);
  MB_LAUNCH_CHECK("fill_u64_kernel");
  return 0;
}
