// groupby.cu — GroupByReduce: open-addressed hash aggregation of one block into an
// L2-resident table, plus the sorted emit.
//
// Reference path: GroupByReduce.map (alg/groupby.py:124-208) runs `df.groupby(by).sum()` per
// row block (pandas: factorize = hash of the key column, then _libs.groupby.group_sum);
// GroupByReduce.reduce (alg/groupby.py:211-300) concatenates the partial tables and regroups.
// Here one table per GPU absorbs every row block resident on that GPU (map + local reduce
// fused), and partial tables from other GPUs are merged with the same kernel.
//
// Data structure (all device resident):
//   slots[cap]   16-byte {int64 key, int32 gid, pad}; cap = pow2 >= 2 * group_capacity.
//                gid: -1 empty, -2 being inserted, [0, gcap) dense group id, gcap = overflowed.
//   acc[gcap][vstride]  float64 sums, row-major so one group's V sums share 64-byte segments.
//   cnt[gcap][vstride]  int64 non-NaN counts (optional), size[gcap] int64 rows (optional).
// For G = 1e6, V = 8 the table is 32 MiB of slots + 64 MiB of sums: it lives in the 126 MB L2
// while the 72 B/row input streams past it with L2 evict-first loads.
//
// Kernel (one warp = 32 consecutive rows per iteration):
//   1. coalesced loads of the key and the V values of the 32 rows (all issued up front);
//   2. warp-cooperative probe: __match_any_sync groups lanes holding the same key, the lowest
//      lane of each group probes / inserts (linear probing, 128-bit slot loads, CAS on gid as
//      the insertion lock), the gid is broadcast back with __shfl_sync;
//   3. accumulate with RED.ADD.F64:  variant 0 re-lays the 32x8 value tile through padded
//      shared memory so that 8 consecutive lanes update the 8 sums of ONE group (one 64-byte
//      segment per row); variant 1 keeps lane == row (32 different segments per instruction).
// Float atomics make the summation order run-dependent: results agree with pandas to the
// tolerance stated in tests (|err| <= 4 log2(n) eps sum|x|), not bit for bit; counts/sizes
// and keys are exact.
#include "common.cuh"

namespace mb200 {

int sort_pairs_device(long long* keys, long long* pay, long long* tmp_keys, long long* tmp_pay,
                      unsigned int* counts, long long n, unsigned long long bias, int nbits, cudaStream_t st);
size_t sort_scratch_bytes(long long n);

struct __align__(16) Slot {
  long long key;
  int gid;
  int pad;
};

struct GbMeta {  // device-side bookkeeping
  int ngroups;
  int overflow;
  long long kmin;
  long long kmax;
};

}  // namespace mb200

struct mb200_gb_table {
  mb200::Slot* slots;
  long long cap;  // power of two
  double* acc;
  long long* cnt;
  long long* size;
  long long gcap;
  int nvals;
  int vstride;
  int flags;
  mb200::GbMeta* meta;
};

namespace mb200 {

constexpr int kGbThreads = 256;
constexpr int kGbWarps = kGbThreads / 32;
constexpr int kColStride = 34;  // doubles; 34 = 2 (mod 16) -> conflict-free transposed reads

struct GbParams {
  Slot* slots;
  unsigned int mask;
  long long cap;
  double* acc;
  long long* cnt;
  long long* size;
  long long gcap;
  int nvals;
  int vstride;
  int flags;
  GbMeta* meta;
  const long long* keys;
  const void* vals[MB200_MAX_COLS];   // raw values, or partial sums when PARTIAL
  const void* pcnt[MB200_MAX_COLS];   // partial counts (PARTIAL only)
  const long long* psize;             // partial sizes (PARTIAL only)
  long long nrows;
};

__device__ __forceinline__ void red_add_f64(double* p, double v) {
  asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void red_add_u64(long long* p, long long v) {
  asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void ld_slot(const Slot* s, long long& key, int& gid) {
  unsigned long long a, b;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(s) : "memory");
  key = (long long)a;
  gid = (int)(unsigned int)(b & 0xffffffffULL);
}

// Find or insert `k`; returns the dense group id, or gcap if the table overflowed.
__device__ __forceinline__ int probe_insert(const GbParams& p, long long k) {
  unsigned int slot = hash_key(k) & p.mask;
  long long probes = 0;
  for (;;) {
    long long sk;
    int g;
    ld_slot(&p.slots[slot], sk, g);
    if (g >= 0) {
      if (sk == k) return g;
      slot = (slot + 1) & p.mask;
      if (++probes > p.cap) {  // table full of other keys
        p.meta->overflow = 1;
        return (int)p.gcap;
      }
      continue;
    }
    if (g == -1) {
      const int old = atomicCAS(&p.slots[slot].gid, -1, -2);
      if (old == -1) {  // we own the slot: write the key, take a dense id, publish
        *reinterpret_cast<volatile long long*>(&p.slots[slot].key) = k;
        int ng = atomicAdd(&p.meta->ngroups, 1);
        if (ng >= p.gcap) {
          p.meta->overflow = 1;
          ng = (int)p.gcap;
        } else {
          atomicMin(&p.meta->kmin, k);
          atomicMax(&p.meta->kmax, k);
        }
        __threadfence();
        *reinterpret_cast<volatile int*>(&p.slots[slot].gid) = ng;
        return ng;
      }
    }
    // g == -2 (or we lost the CAS): another thread is publishing this slot; look again
  }
}

// Read-only lookup: dense group id of `k`, or -1 if the probe chain ends at an empty or
// in-flight slot (the caller then takes the serialized insert path).
__device__ __forceinline__ int probe_find(const GbParams& p, long long k) {
  unsigned int slot = hash_key(k) & p.mask;
  for (long long probes = 0; probes <= p.cap; ++probes) {
    long long sk;
    int g;
    ld_slot(&p.slots[slot], sk, g);
    if (g < 0) return -1;
    if (sk == k) return g;
    slot = (slot + 1) & p.mask;
  }
  return -1;
}

template <int VARIANT, bool PARTIAL>
__global__ void __launch_bounds__(kGbThreads) gb_accumulate_kernel(const __grid_constant__ GbParams p) {
  // variant 0: [8 cols][34] column tiles; variant 2: [32 rows][10] row tiles (80-byte stride: 16-byte
  // aligned rows, conflict-free 128-bit stores)
  __shared__ __align__(16) double s_tile[kGbWarps][32 * 10];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long nchunks = (p.nrows + 31) >> 5;
  const long long wstride = (long long)gridDim.x * kGbWarps;
  const int gcap = (int)p.gcap;
  const uint64_t pol = l2_policy_evict_first();
  for (long long ch = (long long)blockIdx.x * kGbWarps + warp; ch < nchunks; ch += wstride) {
    const long long base = ch << 5;
    const long long row = base + lane;
    const bool valid = row < p.nrows;
    long long k = valid ? ldg_stream_i64(p.keys + row, pol) : 0;
    // value loads of the first 8 columns are issued before the (dependent, random) probe
    double x[8];
    {
      const int nc = p.nvals < 8 ? p.nvals : 8;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        x[c] = (c < nc && valid) ? ldg_stream_f64(static_cast<const double*>(p.vals[c]) + row, pol) : 0.0;
    }
    // ---- probe (one lane per distinct key in the warp)
    {
      // rows past the end borrow lane 0's key (lane 0 is always valid in a live chunk) so that they join
      // its peer group and never become leaders; the shuffle is executed by ALL lanes.
      const long long k0 = __shfl_sync(0xffffffffu, k, 0);
      k = valid ? k : k0;
    }
    const unsigned int peers = __match_any_sync(0xffffffffu, (unsigned long long)k);
    const int leader = __ffs(peers) - 1;
    const bool is_leader = (lane == leader);
    // fast path: read-only lookup by every leader in parallel (the steady state once the table is warm)
    int gid = -1;
    if (is_leader) gid = probe_find(p, k);
    // slow path: leaders that met an empty / in-flight slot insert ONE LANE AT A TIME, so a lane that
    // spins on a slot lock can only be waiting for another warp (which makes progress independently),
    // never for a diverged lane of its own warp.
    unsigned int need = __ballot_sync(0xffffffffu, is_leader && gid < 0);
    while (need) {
      const int l = __ffs(need) - 1;
      if (lane == l) gid = probe_insert(p, k);
      need &= need - 1;
      __syncwarp();
    }
    gid = __shfl_sync(0xffffffffu, gid, leader);
    const bool live = valid && gid < gcap;

    if ((p.flags & MB200_GB_SIZE) && live) {
      red_add_u64(p.size + gid, PARTIAL ? p.psize[row] : 1LL);
    }
    // ---- accumulate, 8 value columns at a time
    for (int c0 = 0; c0 < p.nvals; c0 += 8) {
      const int nc = (p.nvals - c0) < 8 ? (p.nvals - c0) : 8;
      if (c0 > 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
          x[c] = (c < nc && valid) ? ldg_stream_f64(static_cast<const double*>(p.vals[c0 + c]) + row, pol) : 0.0;
      }
      if (VARIANT == 2) {
        // One TMA bulk reduction per row: the row's (<= 8) values are laid out contiguously in shared
        // memory and added to the group's 64-byte accumulator row by the copy engine
        // (cp.reduce.async.bulk ... .add.f64 -> UBLKRED), i.e. ONE async op per row instead of 8 RED
        // lanes through the LSU.  NaNs are replaced by +0.0 (adding zero == skipping).
        double* myrow = s_tile[warp] + lane * 10;
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // my previous row has been read
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
          double2 d;
          d.x = (c < nc && x[c] == x[c]) ? x[c] : 0.0;
          d.y = (c + 1 < nc && x[c + 1] == x[c + 1]) ? x[c + 1] : 0.0;
          *reinterpret_cast<double2*>(myrow + c) = d;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (live) {
          const unsigned int bytes = (unsigned int)(((nc + 1) >> 1) * 16);
          double* dst = p.acc + (size_t)gid * p.vstride + c0;
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst),
                       "r"(smem_u32(myrow)), "r"(bytes)
                       : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      } else if (VARIANT == 1) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c < nc && live) {
            const size_t o = (size_t)gid * p.vstride + c0 + c;
            if (p.flags & MB200_GB_SUM) {
              if (x[c] == x[c]) red_add_f64(p.acc + o, x[c]);
            }
            if (p.flags & MB200_GB_COUNT) {
              if (PARTIAL) red_add_u64(p.cnt + o, static_cast<const long long*>(p.pcnt[c0 + c])[row]);
              else if (x[c] == x[c]) red_add_u64(p.cnt + o, 1LL);
            }
          }
        }
      } else {
        double* tile = s_tile[warp];
#pragma unroll
        for (int c = 0; c < 8; ++c) tile[c * kColStride + lane] = x[c];
        __syncwarp();
        const int c = lane & 7;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const int r = 4 * kk + (lane >> 3);
          const int g = __shfl_sync(0xffffffffu, gid, r);
          const double xv = tile[c * kColStride + r];
          const bool ok = (base + r < p.nrows) && (c < nc) && (g < gcap);
          if (ok) {
            const size_t o = (size_t)g * p.vstride + c0 + c;
            if (p.flags & MB200_GB_SUM) {
              if (xv == xv) red_add_f64(p.acc + o, xv);
            }
            if (p.flags & MB200_GB_COUNT) {
              if (PARTIAL) red_add_u64(p.cnt + o, static_cast<const long long*>(p.pcnt[c0 + c])[base + r]);
              else if (xv == xv) red_add_u64(p.cnt + o, 1LL);
            }
          }
        }
        __syncwarp();
      }
    }
  }
  if (VARIANT == 2) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void gb_init_kernel(Slot* slots, long long cap, GbMeta* meta) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) {
    Slot s;
    s.key = 0;
    s.gid = -1;
    s.pad = 0;
    slots[i] = s;
  }
  if (i == 0) {
    meta->ngroups = 0;
    meta->overflow = 0;
    meta->kmin = 0x7fffffffffffffffLL;
    meta->kmax = (long long)0x8000000000000000ULL;
  }
}

// keys_by_gid[gid] = key ; perm[gid] = gid
__global__ void gb_collect_kernel(const Slot* __restrict__ slots, long long cap, long long gcap,
                                  long long* __restrict__ keys_by_gid, long long* __restrict__ perm) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) {
    const Slot s = slots[i];
    if (s.gid >= 0 && s.gid < gcap) {
      keys_by_gid[s.gid] = s.key;
      perm[s.gid] = s.gid;
    }
  }
}

struct EmitParams {
  const long long* keys_sorted;
  const long long* perm;
  const double* acc;
  const long long* cnt;
  const long long* size;
  int nvals;
  int vstride;
  long long ngroups;
  long long* out_keys;
  void* out_sums[MB200_MAX_COLS];
  void* out_cnts[MB200_MAX_COLS];
  long long* out_sizes;
};

__global__ void gb_emit_kernel(const __grid_constant__ EmitParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.ngroups) return;
  const long long g = p.perm[i];
  if (p.out_keys) p.out_keys[i] = p.keys_sorted[i];
  if (p.out_sizes && p.size) p.out_sizes[i] = p.size[g];
  for (int v = 0; v < p.nvals; ++v) {
    if (p.out_sums[v] && p.acc) static_cast<double*>(p.out_sums[v])[i] = p.acc[(size_t)g * p.vstride + v];
    if (p.out_cnts[v] && p.cnt) static_cast<long long*>(p.out_cnts[v])[i] = p.cnt[(size_t)g * p.vstride + v];
  }
}

static long long next_pow2(long long v) {
  long long p = 1;
  while (p < v) p <<= 1;
  return p;
}

static int gb_variant_from_env() {
  const char* e = getenv("MB200_GB_VARIANT");  // read per call: lets one process compare variants
  if (e && e[0] == '1') return 1;
  if (e && e[0] == '2') return 2;
  return 0;
}

static int gb_launch(mb200_gb_table* t, const long long* keys, const void* const* vals, const void* const* pcnt,
                     const long long* psize, long long nrows, bool partial, cudaStream_t st) {
  if (nrows == 0) return 0;
  DevProps dp;
  if (int rc = dev_props(&dp)) return rc;
  GbParams p;
  memset(&p, 0, sizeof(p));
  p.slots = t->slots;
  p.mask = (unsigned int)(t->cap - 1);
  p.cap = t->cap;
  p.acc = t->acc;
  p.cnt = t->cnt;
  p.size = t->size;
  p.gcap = t->gcap;
  p.nvals = t->nvals;
  p.vstride = t->vstride;
  p.flags = t->flags;
  p.meta = t->meta;
  p.keys = keys;
  for (int c = 0; c < t->nvals; ++c) {
    p.vals[c] = vals ? vals[c] : nullptr;
    p.pcnt[c] = pcnt ? pcnt[c] : nullptr;
    if ((t->flags & (MB200_GB_SUM | MB200_GB_COUNT)) && !p.vals[c]) return fail("groupby", "null value column");
    if (partial && (t->flags & MB200_GB_COUNT) && !p.pcnt[c]) return fail("groupby", "null partial count column");
  }
  p.psize = psize;
  if (partial && (t->flags & MB200_GB_SIZE) && !psize) return fail("groupby", "null partial size column");
  p.nrows = nrows;
  int variant = gb_variant_from_env();
  // the bulk-reduce variant covers plain sums of raw rows; counts / partial merges use variant 0
  if (variant == 2 && (partial || (t->flags & MB200_GB_COUNT) || !(t->flags & MB200_GB_SUM))) variant = 0;
  int occ = 0;
  const long long nchunks = (nrows + 31) / 32;
#define MB_GB_LAUNCH(V, P)                                                                                  \
  do {                                                                                                      \
    MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gb_accumulate_kernel<V, P>, kGbThreads, 0)); \
    long long grid = (long long)dp.sm_count * (occ < 1 ? 1 : occ);                                          \
    const long long need = (nchunks + kGbWarps - 1) / kGbWarps;                                             \
    if (grid > need) grid = need;                                                                           \
    gb_accumulate_kernel<V, P><<<(unsigned)grid, kGbThreads, 0, st>>>(p);                                   \
  } while (0)
  if (variant == 2) {
    MB_GB_LAUNCH(2, false);
  } else if (variant == 0) {
    if (partial) MB_GB_LAUNCH(0, true);
    else MB_GB_LAUNCH(0, false);
  } else {
    if (partial) MB_GB_LAUNCH(1, true);
    else MB_GB_LAUNCH(1, false);
  }
#undef MB_GB_LAUNCH
  MB_LAUNCH_CHECK("gb_accumulate_kernel");
  return 0;
}

}  // namespace mb200

using namespace mb200;

extern "C" int mb200_gb_create(mb200_gb_table** table, int64_t group_capacity, int nvals, int flags,
                               mb200_stream_t stream) {
  if (!table) return fail("mb200_gb_create", "null out pointer");
  if (nvals < 0 || nvals > MB200_MAX_COLS) return fail("mb200_gb_create", "nvals out of range (0..32)");
  if (group_capacity < 1) group_capacity = 1;
  if (group_capacity > (1LL << 29)) return fail("mb200_gb_create", "group capacity above 2^29");
  DevProps dp;
  if (int rc = dev_props(&dp)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  mb200_gb_table* t = new mb200_gb_table();
  memset(t, 0, sizeof(*t));
  t->gcap = group_capacity;
  t->cap = next_pow2(2 * group_capacity < 1024 ? 1024 : 2 * group_capacity);
  t->nvals = nvals;
  t->vstride = (nvals + 3) & ~3;  // 32-byte sector aligned rows
  if (t->vstride == 0) t->vstride = 4;
  t->flags = flags;
  cudaError_t e;
  const size_t accb = (size_t)t->gcap * t->vstride * 8;
#define MB_TRY(call)                \
  do {                              \
    e = (call);                     \
    if (e != cudaSuccess) goto bad; \
  } while (0)
  MB_TRY(cudaMallocAsync((void**)&t->slots, (size_t)t->cap * sizeof(Slot), st));
  MB_TRY(cudaMallocAsync((void**)&t->meta, sizeof(GbMeta), st));
  if (flags & MB200_GB_SUM) {
    MB_TRY(cudaMallocAsync((void**)&t->acc, accb, st));
    MB_TRY(cudaMemsetAsync(t->acc, 0, accb, st));
  }
  if (flags & MB200_GB_COUNT) {
    MB_TRY(cudaMallocAsync((void**)&t->cnt, accb, st));
    MB_TRY(cudaMemsetAsync(t->cnt, 0, accb, st));
  }
  if (flags & MB200_GB_SIZE) {
    MB_TRY(cudaMallocAsync((void**)&t->size, (size_t)t->gcap * 8, st));
    MB_TRY(cudaMemsetAsync(t->size, 0, (size_t)t->gcap * 8, st));
  }
#undef MB_TRY
  gb_init_kernel<<<(unsigned)((t->cap + 255) / 256), 256, 0, st>>>(t->slots, t->cap, t->meta);
  e = cudaGetLastError();
  if (e != cudaSuccess) goto bad;
  g_launches.fetch_add(1);
  *table = t;
  return 0;
bad:
  mb200_gb_destroy(t, stream);
  return cuda_fail("mb200_gb_create", e);
}

extern "C" int mb200_gb_destroy(mb200_gb_table* t, mb200_stream_t stream) {
  if (!t) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (t->slots) cudaFreeAsync(t->slots, st);
  if (t->meta) cudaFreeAsync(t->meta, st);
  if (t->acc) cudaFreeAsync(t->acc, st);
  if (t->cnt) cudaFreeAsync(t->cnt, st);
  if (t->size) cudaFreeAsync(t->size, st);
  delete t;
  return 0;
}

extern "C" int mb200_gb_accumulate(mb200_gb_table* t, const int64_t* keys, const void* const* vals,
                                   int64_t nrows, mb200_stream_t stream) {
  if (!t) return fail("mb200_gb_accumulate", "null table");
  if (nrows < 0) return fail("mb200_gb_accumulate", "negative nrows");
  if (nrows > 0 && !keys) return fail("mb200_gb_accumulate", "null keys");
  return gb_launch(t, reinterpret_cast<const long long*>(keys), vals, nullptr, nullptr, nrows, false,
                   (cudaStream_t)stream);
}

extern "C" int mb200_gb_merge_partial(mb200_gb_table* t, const int64_t* keys, const void* const* sums,
                                      const void* const* cnts, const int64_t* sizes, int64_t npartial,
                                      mb200_stream_t stream) {
  if (!t) return fail("mb200_gb_merge_partial", "null table");
  if (npartial < 0) return fail("mb200_gb_merge_partial", "negative npartial");
  if (npartial > 0 && !keys) return fail("mb200_gb_merge_partial", "null keys");
  return gb_launch(t, reinterpret_cast<const long long*>(keys), sums, cnts, reinterpret_cast<const long long*>(sizes),
                   npartial, true, (cudaStream_t)stream);
}

extern "C" int mb200_gb_ngroups(mb200_gb_table* t, int64_t* ngroups, int* overflow, mb200_stream_t stream) {
  if (!t) return fail("mb200_gb_ngroups", "null table");
  GbMeta m;
  MB_CUDA(cudaMemcpyAsync(&m, t->meta, sizeof(m), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  MB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  long long n = m.ngroups;
  if (n > t->gcap) n = t->gcap;
  if (ngroups) *ngroups = n;
  if (overflow) *overflow = m.overflow;
  return 0;
}

extern "C" size_t mb200_gb_emit_scratch_bytes(int64_t ngroups) {
  if (ngroups < 1) ngroups = 1;
  // keys_by_gid + perm (2 * 8 * n) + sort scratch
  return (size_t)ngroups * 16 + 512 + sort_scratch_bytes(ngroups);
}

extern "C" int mb200_gb_emit(mb200_gb_table* t, int64_t ngroups, int sort, int64_t* out_keys,
                             void* const* out_sums, void* const* out_cnts, int64_t* out_sizes, void* scratch,
                             mb200_stream_t stream) {
  if (!t) return fail("mb200_gb_emit", "null table");
  if (ngroups < 0 || ngroups > t->gcap) return fail("mb200_gb_emit", "ngroups out of range");
  if (ngroups == 0) return 0;
  if (!scratch) return fail("mb200_gb_emit", "null scratch");
  cudaStream_t st = (cudaStream_t)stream;
  GbMeta m;
  MB_CUDA(cudaMemcpyAsync(&m, t->meta, sizeof(m), cudaMemcpyDeviceToHost, st));
  MB_CUDA(cudaStreamSynchronize(st));
  if (m.overflow) return fail("mb200_gb_emit", "table overflowed: recreate with a larger group capacity");

  char* s = static_cast<char*>(scratch);
  long long* keys_by_gid = reinterpret_cast<long long*>(s);
  long long* perm = reinterpret_cast<long long*>(s + (size_t)ngroups * 8);
  size_t off = ((size_t)ngroups * 16 + 255) & ~(size_t)255;
  char* sort_s = s + off;
  gb_collect_kernel<<<(unsigned)((t->cap + 255) / 256), 256, 0, st>>>(t->slots, t->cap, t->gcap, keys_by_gid, perm);
  MB_LAUNCH_CHECK("gb_collect_kernel");
  if (sort) {
    long long* tk = reinterpret_cast<long long*>(sort_s);
    long long* tp = reinterpret_cast<long long*>(sort_s + (size_t)ngroups * 8);
    size_t coff = ((size_t)ngroups * 16 + 255) & ~(size_t)255;
    unsigned int* counts = reinterpret_cast<unsigned int*>(sort_s + coff);
    const unsigned long long range = (unsigned long long)m.kmax - (unsigned long long)m.kmin;
    int nbits = 0;
    while (nbits < 64 && (range >> nbits) != 0) ++nbits;
    if (int rc = sort_pairs_device(keys_by_gid, perm, tk, tp, counts, ngroups, (unsigned long long)m.kmin, nbits, st))
      return rc;
  }
  EmitParams p;
  memset(&p, 0, sizeof(p));
  p.keys_sorted = keys_by_gid;
  p.perm = perm;
  p.acc = t->acc;
  p.cnt = t->cnt;
  p.size = t->size;
  p.nvals = t->nvals;
  p.vstride = t->vstride;
  p.ngroups = ngroups;
  p.out_keys = reinterpret_cast<long long*>(out_keys);
  for (int v = 0; v < t->nvals; ++v) {
    p.out_sums[v] = out_sums ? out_sums[v] : nullptr;
    p.out_cnts[v] = out_cnts ? out_cnts[v] : nullptr;
  }
  p.out_sizes = reinterpret_cast<long long*>(out_sizes);
  gb_emit_kernel<<<(unsigned)((ngroups + 255) / 256), 256, 0, st>>>(p);
  MB_LAUNCH_CHECK("gb_emit_kernel");
  return 0;
}
