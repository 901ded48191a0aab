// hostpipe.cu — mb200_map_host: the Map/Binary sweep for HOST-resident blocks.
//
// This is the end-to-end form of pm.map_partitions (pm.py:708-769) for a frame that still lives
// in host memory (pandas/Arrow buffers): row chunks are copied H2D, swept by the same kernel as
// mb200_map, and copied back D2H, with the three stages overlapped on three streams over a
// ring of device staging buffers.  PCIe-bound by construction (8 B in + 8 B out per element);
// pinned host buffers (mb200_alloc_host) make the copies truly asynchronous.
#include <memory>
#include <vector>

#include "common.cuh"

using namespace mb200;

namespace {
constexpr int kRing = 3;

struct Pipe {
  cudaStream_t s_in = nullptr, s_k = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[kRing] = {}, ev_k[kRing] = {}, ev_out[kRing] = {};
  std::vector<void*> dev;  // every staging allocation
  ~Pipe() {
    for (void* p : dev) cudaFree(p);
    for (int i = 0; i < kRing; ++i) {
      if (ev_in[i]) cudaEventDestroy(ev_in[i]);
      if (ev_k[i]) cudaEventDestroy(ev_k[i]);
      if (ev_out[i]) cudaEventDestroy(ev_out[i]);
    }
    if (s_in) cudaStreamDestroy(s_in);
    if (s_k) cudaStreamDestroy(s_k);
    if (s_out) cudaStreamDestroy(s_out);
  }
};
}  // namespace

extern "C" int mb200_map_host(int op, int dtype, int ncols, const void* const* in0_host, const void* const* in1_host,
                              const void* const* in2_host, void* const* out_host, int64_t nrows, const uint64_t* s0,
                              const uint64_t* s1, int64_t chunk_rows) {
  if (ncols < 0 || ncols > MB200_MAX_COLS) return fail("mb200_map_host", "ncols out of range (0..32)");
  if (nrows < 0) return fail("mb200_map_host", "negative nrows");
  if (ncols == 0 || nrows == 0) return 0;
  if (!in0_host || !out_host) return fail("mb200_map_host", "null column array");
  const int nin = op >= 64 ? 3 : (op >= 32 ? 2 : 1);
  if ((nin >= 2 && !in1_host) || (nin >= 3 && !in2_host)) return fail("mb200_map_host", "missing operand frame");
  if (chunk_rows <= 0) chunk_rows = 1 << 22;
  chunk_rows = (chunk_rows + 4095) / 4096 * 4096;
  if (chunk_rows > nrows) chunk_rows = (nrows + 4095) / 4096 * 4096;
  const bool pred = (op == MB200_OP_ISNA || op == MB200_OP_NOTNA || (op >= MB200_OP_EQ_S && op <= MB200_OP_GE_S) ||
                     (op >= MB200_OP_EQ && op <= MB200_OP_GE));
  const bool int_div = dtype == MB200_I64 && (op == MB200_OP_DIV || op == MB200_OP_DIV_S || op == MB200_OP_RDIV_S);
  (void)int_div;
  const size_t in_es = 8, out_es = pred ? 1 : 8;

  DevProps dp;
  if (int rc = dev_props(&dp)) return rc;
  // The pipe (3 streams, 9 events, ring x (nin + 1) x ncols staging buffers -- 1.5 GB at the default chunk size
  // and 8 columns) is kept between calls with the same shape: allocating and freeing it costs more than a chunk.
  struct Cached {
    std::unique_ptr<Pipe> pipe;
    int dev = -1, ncols = 0, nin = 0;
    long long chunk_rows = 0;
    size_t out_es = 0;
    void* d_in[kRing][3][MB200_MAX_COLS];
    void* d_out[kRing][MB200_MAX_COLS];
  };
  static thread_local Cached cache;
  int dev = 0;
  MB_CUDA(cudaGetDevice(&dev));
  if (!cache.pipe || cache.dev != dev || cache.ncols != ncols || cache.nin != nin || cache.chunk_rows != chunk_rows ||
      cache.out_es != out_es) {
    cache.pipe.reset();  // frees the previous staging buffers first
    std::unique_ptr<Pipe> np(new Pipe());
    MB_CUDA(cudaStreamCreateWithFlags(&np->s_in, cudaStreamNonBlocking));
    MB_CUDA(cudaStreamCreateWithFlags(&np->s_k, cudaStreamNonBlocking));
    MB_CUDA(cudaStreamCreateWithFlags(&np->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < kRing; ++i) {
      MB_CUDA(cudaEventCreateWithFlags(&np->ev_in[i], cudaEventDisableTiming));
      MB_CUDA(cudaEventCreateWithFlags(&np->ev_k[i], cudaEventDisableTiming));
      MB_CUDA(cudaEventCreateWithFlags(&np->ev_out[i], cudaEventDisableTiming));
    }
    for (int r = 0; r < kRing; ++r) {
      for (int c = 0; c < ncols; ++c) {
        for (int k = 0; k < nin; ++k) {
          void* p = nullptr;
          MB_CUDA(cudaMalloc(&p, (size_t)chunk_rows * in_es));
          np->dev.push_back(p);
          cache.d_in[r][k][c] = p;
        }
        void* p = nullptr;
        MB_CUDA(cudaMalloc(&p, (size_t)chunk_rows * out_es));
        np->dev.push_back(p);
        cache.d_out[r][c] = p;
      }
    }
    cache.pipe = std::move(np);
    cache.dev = dev;
    cache.ncols = ncols;
    cache.nin = nin;
    cache.chunk_rows = chunk_rows;
    cache.out_es = out_es;
  }
  Pipe& pp = *cache.pipe;
  auto& d_in = cache.d_in;
  auto& d_out = cache.d_out;
  const void* const* hin[3] = {in0_host, in1_host, in2_host};
  const long long nchunks = (nrows + chunk_rows - 1) / chunk_rows;
  for (long long j = 0; j < nchunks; ++j) {
    const int r = (int)(j % kRing);
    const long long row0 = j * chunk_rows;
    const long long rows = (row0 + chunk_rows <= nrows) ? chunk_rows : (nrows - row0);
    if (j >= kRing) {
      // the D2H of the chunk that used this ring slot must be done before we overwrite its inputs/outputs
      MB_CUDA(cudaStreamWaitEvent(pp.s_in, pp.ev_out[r], 0));
    }
    for (int k = 0; k < nin; ++k)
      for (int c = 0; c < ncols; ++c)
        MB_CUDA(cudaMemcpyAsync(d_in[r][k][c], static_cast<const char*>(hin[k][c]) + (size_t)row0 * in_es,
                                (size_t)rows * in_es, cudaMemcpyHostToDevice, pp.s_in));
    MB_CUDA(cudaEventRecord(pp.ev_in[r], pp.s_in));
    MB_CUDA(cudaStreamWaitEvent(pp.s_k, pp.ev_in[r], 0));
    if (int rc = mb200_map(op, dtype, ncols, d_in[r][0], nin >= 2 ? d_in[r][1] : nullptr,
                           nin >= 3 ? d_in[r][2] : nullptr, d_out[r], rows, s0, s1, pp.s_k))
      return rc;
    MB_CUDA(cudaEventRecord(pp.ev_k[r], pp.s_k));
    MB_CUDA(cudaStreamWaitEvent(pp.s_out, pp.ev_k[r], 0));
    for (int c = 0; c < ncols; ++c)
      MB_CUDA(cudaMemcpyAsync(static_cast<char*>(out_host[c]) + (size_t)row0 * out_es, d_out[r][c],
                              (size_t)rows * out_es, cudaMemcpyDeviceToHost, pp.s_out));
    MB_CUDA(cudaEventRecord(pp.ev_out[r], pp.s_out));
  }
  MB_CUDA(cudaStreamSynchronize(pp.s_in));
  MB_CUDA(cudaStreamSynchronize(pp.s_k));
  MB_CUDA(cudaStreamSynchronize(pp.s_out));
  return 0;
}
