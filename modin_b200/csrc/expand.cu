// expand.cu — many-to-many broadcast merge: the row expansion that follows the probe when the broadcast (right)
// side holds DUPLICATE keys.
//
// Reference: MergeImpl.row_axis_merge hands every row partition of the left frame plus the whole right frame to
// `pandas.merge(left_block, right, how, on, sort=False)` (storage_formats/pandas/merge.py:139-168); with repeated
// right keys every left row yields one output row per matching right row, left order preserved and the matches in
// their order of appearance on the right.  On the device:
//   1. the right keys are stably sorted together with their row ids (sort.cu), `run_heads_kernel` marks where a new
//      key starts, the compaction of those marks (join.cu) gives `starts[U]`, the U distinct keys go into a join
//      table (dense or hashed, join.cu) -- so the probe stays many-to-one;
//   2. `expand_counts_kernel`: per left row, how many output rows it produces and where its run of right rows starts;
//   3. `scan_*`: exclusive prefix sum of the counts (int64, 2048 rows per block, one single-block pass over the
//      block sums) -> output position of every left row and the total, which sizes the result;
//   4. `expand_rows_kernel`: one thread per left row writes its (left row, right row) pairs; the result columns are
//      then two gathers (`mb200_take`).
// All passes are streaming except the final gathers; integer / index work only, bit-exact.
#include "common.cuh"

namespace mb200 {

constexpr int kScanBlock = 256;
constexpr int kScanItems = 2048;  // rows per block

__global__ void __launch_bounds__(256) run_heads_kernel(const long long* __restrict__ sorted_keys, long long n,
                                                        long long* __restrict__ idx_out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    idx_out[i] = (i == 0 || sorted_keys[i] != sorted_keys[i - 1]) ? i : -1;
}

__global__ void __launch_bounds__(256) expand_counts_kernel(const long long* __restrict__ u, long long n,
                                                            const long long* __restrict__ starts, long long nuniq,
                                                            long long nright, int keep_misses,
                                                            long long* __restrict__ cnt, long long* __restrict__ first) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long g = u[i];
    if (g >= 0 && g < nuniq) {
      const long long s = starts[g];
      const long long e = (g + 1 < nuniq) ? starts[g + 1] : nright;
      cnt[i] = e - s;
      first[i] = s;
    } else {
      cnt[i] = keep_misses ? 1 : 0;
      first[i] = -1;
    }
  }
}

__global__ void __launch_bounds__(kScanBlock) scan_block_sums_kernel(const long long* __restrict__ v, long long n,
                                                                     long long* __restrict__ block_sums) {
  __shared__ long long s[kScanBlock / 32];
  const long long base = (long long)blockIdx.x * kScanItems;
  long long c = 0;
  for (int j = threadIdx.x; j < kScanItems; j += kScanBlock) {
    const long long i = base + j;
    if (i < n) c += v[i];
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < kScanBlock / 32; ++w) t += s[w];
    block_sums[blockIdx.x] = t;
  }
}

// single block: exclusive scan of block_sums[nblocks] in place; total -> *total
__global__ void __launch_bounds__(1024) scan_sums_kernel(long long* __restrict__ block_sums, long long nblocks,
                                                         long long* total) {
  __shared__ long long part[1024];
  const int t = threadIdx.x;
  const long long per = (nblocks + 1023) / 1024;
  const long long lo = (long long)t * per;
  long long hi = lo + per;
  if (hi > nblocks) hi = nblocks;
  long long s = 0;
  for (long long i = lo; i < hi; ++i) s += block_sums[i];
  part[t] = s;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const long long v = t >= off ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  long long run = t ? part[t - 1] : 0;
  for (long long i = lo; i < hi; ++i) {
    const long long c = block_sums[i];
    block_sums[i] = run;
    run += c;
  }
  if (t == 1023 && total) *total = part[1023];
}

// block b: exclusive scan of its 2048 values, offset by block_sums[b] (each thread owns 8 consecutive values)
__global__ void __launch_bounds__(kScanBlock) scan_apply_kernel(const long long* __restrict__ v, long long n,
                                                                const long long* __restrict__ block_offsets,
                                                                long long* __restrict__ out) {
  __shared__ long long warp_tot[kScanBlock / 32];
  constexpr int kPer = kScanItems / kScanBlock;  // 8
  const long long base = (long long)blockIdx.x * kScanItems + (long long)threadIdx.x * kPer;
  long long x[kPer];
  long long sum = 0;
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    x[j] = (base + j < n) ? v[base + j] : 0;
    sum += x[j];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long incl = sum;
#pragma unroll
  for (int m = 1; m < 32; m <<= 1) {
    const long long y = __shfl_up_sync(0xffffffffu, incl, m);
    if (lane >= m) incl += y;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  long long run = block_offsets[blockIdx.x] + incl - sum;
  for (int w = 0; w < warp; ++w) run += warp_tot[w];
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    if (base + j < n) out[base + j] = run;
    run += x[j];
  }
}

__global__ void __launch_bounds__(256) expand_rows_kernel(const long long* __restrict__ offsets,
                                                          const long long* __restrict__ cnt,
                                                          const long long* __restrict__ first,
                                                          const long long* __restrict__ order, long long n,
                                                          long long* __restrict__ out_left,
                                                          long long* __restrict__ out_right) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long o = offsets[i], c = cnt[i], f = first[i];
    for (long long j = 0; j < c; ++j) {
      out_left[o + j] = i;
      out_right[o + j] = f >= 0 ? order[f + j] : -1;
    }
  }
}

// bins[i] = number of pivots <= values[i] (np.digitize(values, pivots, right=False) for ascending pivots): the split
// step of the range-partitioning shuffle (ShuffleSortFunctions.split_partitions, dfutils.py:355-475).  Up to 1023
// pivots live in shared memory; binary search per row.
__global__ void __launch_bounds__(256) digitize_kernel(const long long* __restrict__ values, long long n,
                                                       const long long* __restrict__ pivots, int npivots,
                                                       long long* __restrict__ bins) {
  __shared__ long long s_piv[1024];
  for (int i = threadIdx.x; i < npivots; i += blockDim.x) s_piv[i] = pivots[i];
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long v = values[i];
    int lo = 0, hi = npivots;  // first pivot > v
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_piv[mid] <= v) lo = mid + 1;
      else hi = mid;
    }
    bins[i] = lo;
  }
}

// Row-wise concatenation of column pieces (pandas.concat of the blocks of an axis partition, axpart.py:445-452; the
// gather before a full-axis function): up to 64 source buffers are copied into ONE destination by one launch.  Work
// is cut into 64 KiB chunks over all sources; a CTA finds the source of its chunk in the (<= 64-entry) prefix table in
// kernel parameters and copies 16 bytes per thread per step (bytes for ragged heads / tails).
constexpr int kConcatMax = 64;
constexpr long long kConcatChunk = 64 << 10;
struct ConcatParams {
  const char* src[kConcatMax];
  long long dst_off[kConcatMax];    // byte offset of the source in the destination
  long long chunk_lo[kConcatMax + 1];  // first chunk id of every source
  long long bytes[kConcatMax];
  int nsrc;
};

__global__ void __launch_bounds__(256) concat_kernel(const __grid_constant__ ConcatParams p, char* __restrict__ dst,
                                                     long long nchunks) {
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    int s = 0;
    while (s + 1 < p.nsrc && p.chunk_lo[s + 1] <= ch) ++s;
    const long long off = (ch - p.chunk_lo[s]) * kConcatChunk;
    long long n = p.bytes[s] - off;
    if (n > kConcatChunk) n = kConcatChunk;
    const char* in = p.src[s] + off;
    char* out = dst + p.dst_off[s] + off;
    if ((((uintptr_t)in | (uintptr_t)out) & 15u) == 0) {
      const long long nv = n >> 4;
      const int4* vi = reinterpret_cast<const int4*>(in);
      int4* vo = reinterpret_cast<int4*>(out);
      for (long long i = threadIdx.x; i < nv; i += blockDim.x) vo[i] = vi[i];
      for (long long i = (nv << 4) + threadIdx.x; i < n; i += blockDim.x) out[i] = in[i];
    } else if ((((uintptr_t)in | (uintptr_t)out) & 7u) == 0) {
      const long long nv = n >> 3;
      const long long* vi = reinterpret_cast<const long long*>(in);
      long long* vo = reinterpret_cast<long long*>(out);
      for (long long i = threadIdx.x; i < nv; i += blockDim.x) vo[i] = vi[i];
      for (long long i = (nv << 3) + threadIdx.x; i < n; i += blockDim.x) out[i] = in[i];
    } else {
      for (long long i = threadIdx.x; i < n; i += blockDim.x) out[i] = in[i];
    }
  }
}

static int grid_for(long long n, const DevProps& dp) {
  long long g = (n + 255) / 256;
  const long long cap = (long long)dp.sm_count * 16;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace mb200

using namespace mb200;

extern "C" int mb200_run_heads(const int64_t* sorted_keys, int64_t n, int64_t* idx_out, mb200_stream_t stream) {
  if (n < 0) return fail("mb200_run_heads", "negative n");
  if (n == 0) return 0;
  if (!sorted_keys || !idx_out) return fail("mb200_run_heads", "null argument");
  DevProps dp;
  if (int rc = dev_props(&dp)) return rc;
  run_heads_kernel<<<grid_for(n, dp), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(sorted_keys), n,
                                                                      reinterpret_cast<long long*>(idx_out));
  MB_LAUNCH_CHECK("run_heads_kernel");
  return 0;
}

extern "C" int mb200_expand_counts(const int64_t* u, int64_t n, const int64_t* starts, int64_t nuniq, int64_t nright,
                                   int keep_misses, int64_t* cnt, int64_t* first, mb200_stream_t stream) {
  if (n < 0 || nuniq < 0 || nright < 0) return fail("mb200_expand_counts", "negative size");
  if (n == 0) return 0;
  if (!u || !cnt || !first || (nuniq > 0 && !starts)) return fail("mb200_expand_counts", "null argument");
  DevProps dp;
  if (int rc = dev_props(&dp)) return rc;
  expand_counts_kernel<<<grid_for(n, dp), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const long long*>(u), n, reinterpret_cast<const long long*>(starts), nuniq, nright, keep_misses,
      reinterpret_cast<long long*>(cnt), reinterpret_cast<long long*>(first));
  MB_LAUNCH_CHECK("expand_counts_kernel");
  return 0;
}

extern "C" size_t mb200_scan_scratch_bytes(int64_t n) {
  if (n < 1) n = 1;
  return (size_t)((n + kScanItems - 1) / kScanItems) * 8 + 256;
}

extern "C" int mb200_scan_i64(const int64_t* values, int64_t n, int64_t* out_offsets, int64_t* out_total_dev,
                              void* scratch, size_t scratch_bytes, mb200_stream_t stream) {
  if (n < 0) return fail("mb200_scan_i64", "negative n");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    if (out_total_dev) MB_CUDA(cudaMemsetAsync(out_total_dev, 0, 8, st));
    return 0;
  }
  if (!values || !out_offsets || !scratch) return fail("mb200_scan_i64", "null argument");
  if (scratch_bytes < mb200_scan_scratch_bytes(n)) return fail("mb200_scan_i64", "scratch too small");
  const long long nblocks = (n + kScanItems - 1) / kScanItems;
  long long* sums = static_cast<long long*>(scratch);
  const long long* v = reinterpret_cast<const long long*>(values);
  scan_block_sums_kernel<<<(unsigned)nblocks, kScanBlock, 0, st>>>(v, n, sums);
  MB_LAUNCH_CHECK("scan_block_sums_kernel");
  scan_sums_kernel<<<1, 1024, 0, st>>>(sums, nblocks, reinterpret_cast<long long*>(out_total_dev));
  MB_LAUNCH_CHECK("scan_sums_kernel");
  scan_apply_kernel<<<(unsigned)nblocks, kScanBlock, 0, st>>>(v, n, sums, reinterpret_cast<long long*>(out_offsets));
  MB_LAUNCH_CHECK("scan_apply_kernel");
  return 0;
}

extern "C" int mb200_expand_rows(const int64_t* offsets, const int64_t* cnt, const int64_t* first, const int64_t* order,
                                 int64_t n, int64_t* out_left, int64_t* out_right, mb200_stream_t stream) {
  if (n < 0) return fail("mb200_expand_rows", "negative n");
  if (n == 0) return 0;
  if (!offsets || !cnt || !first || !out_left || !out_right) return fail("mb200_expand_rows", "null argument");
  DevProps dp;
  if (int rc = dev_props(&dp)) return rc;
  expand_rows_kernel<<<grid_for(n, dp), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const long long*>(offsets), reinterpret_cast<const long long*>(cnt),
      reinterpret_cast<const long long*>(first), reinterpret_cast<const long long*>(order), n,
      reinterpret_cast<long long*>(out_left), reinterpret_cast<long long*>(out_right));
  MB_LAUNCH_CHECK("expand_rows_kernel");
  return 0;
}

extern "C" int mb200_digitize_i64(const int64_t* values, int64_t n, const int64_t* pivots_dev, int npivots,
                                  int64_t* out_bins, mb200_stream_t stream) {
  if (n < 0 || npivots < 0 || npivots > 1023) return fail("mb200_digitize_i64", "bad sizes (0 <= npivots <= 1023)");
  if (n == 0) return 0;
  if (!values || !out_bins || (npivots > 0 && !pivots_dev)) return fail("mb200_digitize_i64", "null argument");
  DevProps dp;
  if (int rc = dev_props(&dp)) return rc;
  digitize_kernel<<<grid_for(n, dp), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const long long*>(values), n, reinterpret_cast<const long long*>(pivots_dev), npivots,
      reinterpret_cast<long long*>(out_bins));
  MB_LAUNCH_CHECK("digitize_kernel");
  return 0;
}

extern "C" int mb200_concat(int nsrc, const void* const* src, const int64_t* src_bytes, void* dst,
                            mb200_stream_t stream) {
  if (nsrc < 0) return fail("mb200_concat", "negative nsrc");
  if (nsrc == 0) return 0;
  if (!src || !src_bytes) return fail("mb200_concat", "null argument");
  {
    long long total = 0;
    for (int i = 0; i < nsrc; ++i) total += src_bytes[i] > 0 ? src_bytes[i] : 0;
    if (total == 0) return 0;  // nothing to copy: an empty destination has no buffer
  }
  if (!dst) return fail("mb200_concat", "null destination");
  DevProps dp;
  if (int rc = dev_props(&dp)) return rc;
  long long dst_off = 0;
  for (int base = 0; base < nsrc; base += kConcatMax) {
    ConcatParams p;
    memset(&p, 0, sizeof(p));
    const int m = (nsrc - base) < kConcatMax ? (nsrc - base) : kConcatMax;
    long long chunks = 0;
    int k = 0;
    for (int i = 0; i < m; ++i) {
      const long long b = src_bytes[base + i];
      if (b < 0) return fail("mb200_concat", "negative source size");
      if (b == 0) continue;
      if (!src[base + i]) return fail("mb200_concat", "null source");
      p.src[k] = static_cast<const char*>(src[base + i]);
      p.bytes[k] = b;
      p.dst_off[k] = dst_off;
      p.chunk_lo[k] = chunks;
      chunks += (b + kConcatChunk - 1) / kConcatChunk;
      dst_off += b;
      ++k;
    }
    if (k == 0) continue;
    p.chunk_lo[k] = chunks;
    p.nsrc = k;
    long long grid = chunks;
    const long long cap = (long long)dp.sm_count * 16;
    if (grid > cap) grid = cap;
    concat_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(p, static_cast<char*>(dst), chunks);
    MB_LAUNCH_CHECK("concat_kernel");
  }
  return 0;
}
