// sort.cu — stable LSD radix sort of (int64 key, int64 payload) pairs, 8 bits per pass.
//
// Used to put group keys in ascending order (pandas groupby(sort=True), alg/groupby.py:174-206
// forces as_index=True; the reduce phase output is key-sorted) and to order partial tables
// for the multi-GPU range exchange.  Sorting happens on group tables (<= a few 1e7 rows), never
// on the 1e9-row fact data, so a simple 3-kernel-per-pass scheme is enough; passes are
// limited to the bytes that differ between min and max key.
#include "common.cuh"

namespace mb200 {

constexpr int kSortThreads = 256;
constexpr int kSortSub = 256;      // items ranked per sub-tile (one per thread, in order)
constexpr int kSortMaxBlocks = 2048;

struct SortGeom {
  long long n;
  long long items_per_block;  // multiple of kSortSub
  int nblocks;
};

static SortGeom sort_geom(long long n) {
  SortGeom g;
  g.n = n;
  long long ipb = 2048;
  while ((n + ipb - 1) / ipb > kSortMaxBlocks) ipb *= 2;
  g.items_per_block = ipb;
  g.nblocks = (int)((n + ipb - 1) / ipb);
  if (g.nblocks < 1) g.nblocks = 1;
  return g;
}

__device__ __forceinline__ uint32_t digit_of(long long key, unsigned long long bias, int shift) {
  return (uint32_t)((((unsigned long long)key - bias) >> shift) & 0xffu);
}

__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const long long* __restrict__ keys, long long n,
                                                                  long long items_per_block,
                                                                  unsigned long long bias, int shift,
                                                                  unsigned int* __restrict__ counts) {
  __shared__ unsigned int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const long long start = (long long)blockIdx.x * items_per_block;
  long long end = start + items_per_block;
  if (end > n) end = n;
  for (long long i = start + threadIdx.x; i < end; i += kSortThreads) atomicAdd(&h[digit_of(keys[i], bias, shift)], 1u);
  __syncthreads();
  counts[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of counts[256 * nblocks] (bin-major) in place; single block.
__global__ void __launch_bounds__(1024) sort_scan_kernel(unsigned int* __restrict__ counts, int total) {
  __shared__ unsigned int part[1024];
  const int t = threadIdx.x;
  const int per = (total + 1023) / 1024;
  const int lo = t * per;
  int hi = lo + per;
  if (hi > total) hi = total;
  unsigned int s = 0;
  for (int i = lo; i < hi; ++i) s += counts[i];
  part[t] = s;
  __syncthreads();
  // Hillis-Steele inclusive scan over 1024 partials
  for (int off = 1; off < 1024; off <<= 1) {
    unsigned int v = t >= off ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  unsigned int run = t ? part[t - 1] : 0;
  for (int i = lo; i < hi; ++i) {
    const unsigned int c = counts[i];
    counts[i] = run;
    run += c;
  }
}

__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(
    const long long* __restrict__ keys_in, const long long* __restrict__ pay_in, long long* __restrict__ keys_out,
    long long* __restrict__ pay_out, long long n, long long items_per_block, unsigned long long bias, int shift,
    const unsigned int* __restrict__ offsets) {
  __shared__ unsigned int offs[256];
  __shared__ unsigned int warp_cnt[kSortThreads / 32][256];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  offs[t] = offsets[(size_t)t * gridDim.x + blockIdx.x];
  const long long start = (long long)blockIdx.x * items_per_block;
  long long end = start + items_per_block;
  if (end > n) end = n;
  for (long long sub = start; sub < end; sub += kSortSub) {
#pragma unroll
    for (int w = 0; w < kSortThreads / 32; ++w) warp_cnt[w][t] = 0;
    __syncthreads();
    const long long i = sub + t;
    const bool valid = i < end;
    long long k = 0, pl = 0;
    uint32_t d = 0x100u + (uint32_t)lane;  // invalid lanes never match a real digit or each other
    if (valid) {
      k = keys_in[i];
      pl = pay_in[i];
      d = digit_of(k, bias, shift);
    }
    const unsigned int peers = __match_any_sync(0xffffffffu, d);
    const unsigned int rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) warp_cnt[warp][d] = __popc(peers);
    __syncthreads();
    {  // thread t owns digit t: turn per-warp counts into per-warp bases, advance the running offset
      unsigned int run = offs[t];
#pragma unroll
      for (int w = 0; w < kSortThreads / 32; ++w) {
        const unsigned int c = warp_cnt[w][t];
        warp_cnt[w][t] = run;
        run += c;
      }
      offs[t] = run;
    }
    __syncthreads();
    if (valid) {
      const unsigned int pos = warp_cnt[warp][d] + rank;
      keys_out[pos] = k;
      pay_out[pos] = pl;
    }
    __syncthreads();
  }
}

// Sort n pairs by (key - bias) ascending over `nbits` significant bits.  Result ends in
// (keys, pay) (copied back if the pass count is odd).  tmp_* are n-element buffers,
// counts holds 256 * kSortMaxBlocks uint32.
int sort_pairs_device(long long* keys, long long* pay, long long* tmp_keys, long long* tmp_pay,
                      unsigned int* counts, long long n, unsigned long long bias, int nbits, cudaStream_t st) {
  if (n <= 1) return 0;
  if (n > 0x7fffffffLL) return fail("sort_pairs", "more than 2^31-1 pairs");
  const SortGeom g = sort_geom(n);
  const int passes = (nbits + 7) / 8;
  long long *ki = keys, *pi = pay, *ko = tmp_keys, *po = tmp_pay;
  for (int ps = 0; ps < passes; ++ps) {
    const int shift = ps * 8;
    sort_hist_kernel<<<g.nblocks, kSortThreads, 0, st>>>(ki, n, g.items_per_block, bias, shift, counts);
    MB_LAUNCH_CHECK("sort_hist_kernel");
    sort_scan_kernel<<<1, 1024, 0, st>>>(counts, 256 * g.nblocks);
    MB_LAUNCH_CHECK("sort_scan_kernel");
    sort_scatter_kernel<<<g.nblocks, kSortThreads, 0, st>>>(ki, pi, ko, po, n, g.items_per_block, bias, shift, counts);
    MB_LAUNCH_CHECK("sort_scatter_kernel");
    long long* tk = ki;
    ki = ko;
    ko = tk;
    long long* tp = pi;
    pi = po;
    po = tp;
  }
  if (ki != keys) {
    MB_CUDA(cudaMemcpyAsync(keys, ki, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    MB_CUDA(cudaMemcpyAsync(pay, pi, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

size_t sort_scratch_bytes(long long n) { return (size_t)n * 16 + (size_t)256 * kSortMaxBlocks * 4 + 256; }

}  // namespace mb200

using namespace mb200;

extern "C" size_t mb200_sort_scratch_bytes(int64_t n) { return sort_scratch_bytes(n < 1 ? 1 : n); }

extern "C" int mb200_sort_pairs_i64(int64_t* keys, int64_t* payload, int64_t n, void* scratch,
                                    size_t scratch_bytes, mb200_stream_t stream) {
  if (n < 0) return fail("mb200_sort_pairs_i64", "negative n");
  if (n <= 1) return 0;
  if (!keys || !payload || !scratch) return fail("mb200_sort_pairs_i64", "null argument");
  if (scratch_bytes < sort_scratch_bytes(n)) return fail("mb200_sort_pairs_i64", "scratch too small");
  char* s = static_cast<char*>(scratch);
  long long* tk = reinterpret_cast<long long*>(s);
  long long* tp = reinterpret_cast<long long*>(s + (size_t)n * 8);
  size_t off = ((size_t)n * 16 + 255) & ~(size_t)255;
  unsigned int* counts = reinterpret_cast<unsigned int*>(s + off);
  // signed order == unsigned order of (key - INT64_MIN)
  return sort_pairs_device(reinterpret_cast<long long*>(keys), reinterpret_cast<long long*>(payload), tk, tp, counts,
                           n, 0x8000000000000000ULL, 64, (cudaStream_t)stream);
}
