// ktb_push.cu — the push/push form of scatter → exec → gather, synchronised INSIDE the kernels.
//
// ktb_scatter_map_gather (ktb_dispatch.cu) has every rank PULL its shard from the root (peer loads)
// and PUSH its result back (peer stores): one kernel per rank, but the rank's NVLink port then carries
// read requests + write data one way and read data + write acks the other, and measured only ~0.6 of
// the link per direction (profiles/r1f_sweep_peer_2gpu.jsonl).  Here both directions carry posted writes only:
//
//   root  : push_scatter_kernel  — ONE launch per call; CTAs run chunk-major: peer-STORE chunk c of every rank's
//           shard into that rank's staging buffer; the CTA that completes chunk c (per-chunk counter) publishes
//           ready[r][c] = seq (st.release.sys) in every rank's memory
//   rank r: push_consume_kernel  — ONE launch per call; the CTAs of chunk c spin (ld.acquire.sys, local memory)
//           until ready[c] >= seq, apply the op to the staged chunk (local HBM read) and peer-STORE the result into
//           the root's result arena; the last CTA publishes ack[r] = seq in the root's memory
//   root  : push_wait_kernel     — stream-ordered completion: spin until every ack[r] >= seq
//
// Why this shape (profiles/r2_summary.md, 2-GPU duplex matrix + NVLink counters): when ONE GPU drives both
// directions of its port (a kernel that loads from the peer and stores to it) the port delivers ~490 GB/s per
// direction; when each side only issues posted WRITES towards the other, both directions run at 689 GB/s at the same
// time (raw 818 GB/s per direction incl. 18.75 % write headers = 0.91 of the 900 GB/s links, acks negligible).
// Round 1 launched one kernel per chunk on both sides: the launch boundaries drained the link 2 x 8 times per call;
// here the chunk hand-off happens inside one resident grid per side.
//
// The flags replace host-side events, so the same code serves one controller process driving N GPUs
// and one process per GPU (CUDA IPC arenas): there is no host synchronisation on the data path at all.
// Staging is double-buffered by call parity; the root does not overwrite buffer (seq & 1) before the
// rank has acknowledged call seq-2.  Every spin carries a wall-clock timeout that raises a sticky
// status word instead of hanging the GPU.
//
// Replaces: kt/serving/spmd/spmd_supervisor.py:366-570 (fan-out, wait loop, concat) and
// kt/serving/remote_worker_pool.py:254-395 (per-pod POST + gather).
#include "ktb_common.cuh"

#include <algorithm>
#include <atomic>
#include <mutex>

namespace ktb {

constexpr int kPushMaxRanks = 16;
constexpr int kPushThreads = 256;
constexpr uint32_t kPushTile = 16384;          // bytes per CTA
struct PushScatterArgs {
  const uint8_t* src[kPushMaxRanks];             // root-local base of this rank's shard
  uint8_t* dst[kPushMaxRanks];                   // peer staging base (call-parity half already applied)
  unsigned long long shard_bytes[kPushMaxRanks];
  unsigned long long chunk_bytes[kPushMaxRanks]; // bytes per chunk of this rank's shard (same rule as the consumer)
  uint32_t tile_prefix[kPushMaxRanks + 1];       // tiles per chunk, prefix over ranks (identical for every chunk)
  unsigned long long* ready[kPushMaxRanks];      // peer: ready_r[0..n_chunks)
  const unsigned long long* ack[kPushMaxRanks];  // root-local: &ack[r]
  int n;
  int n_chunks;
  unsigned long long seq;
};

__global__ void __launch_bounds__(kPushThreads)
    push_scatter_kernel(const __grid_constant__ PushScatterArgs a, unsigned int* chunk_done, unsigned int* status,
                        uint32_t slice) {
  const uint32_t tiles_per_chunk = a.tile_prefix[a.n];
  const uint32_t total_tiles = tiles_per_chunk * (uint32_t)a.n_chunks;
  __shared__ bool last;
  // Every CTA streams a SLICE of `slice` consecutive 16 KiB tiles (chunk-major tile order) and reports them with ONE
  // system fence + one atomic.  slice = 1 measured best (see g_push_slice); grid = ceil(tiles / slice) by default, a
  // smaller grid walks further slices.
  uint32_t cur_chunk = 0xffffffffu, n_done = 0;
  auto flush = [&]() {   // all threads call it together
    __syncthreads();     // every thread's stores of the finished tiles are issued
    if (threadIdx.x == 0) {
      __threadfence_system();
      last = (atomicAdd(&chunk_done[cur_chunk], n_done) + n_done == tiles_per_chunk);
    }
    __syncthreads();
    if (last) {          // this CTA completed the chunk: publish "chunk landed" to every rank
      if (threadIdx.x < a.n) {
        __threadfence_system();
        st_release_sys(a.ready[threadIdx.x] + cur_chunk, a.seq);
      }
      if (threadIdx.x == 0) chunk_done[cur_chunk] = 0;
    }
    __syncthreads();     // `last` is reused
  };
  for (uint32_t base = blockIdx.x * slice; base < total_tiles; base += gridDim.x * slice)
  for (uint32_t tile = base; tile < min(total_tiles, base + slice); ++tile) {
    const uint32_t c = tile / tiles_per_chunk;                // chunk 0 of every rank goes out first
    const uint32_t rem = tile % tiles_per_chunk;
    if (c != cur_chunk) {
      if (n_done) flush();
      cur_chunk = c;
      n_done = 0;
    }
    int seg = 0;
    while (seg + 1 < a.n && a.tile_prefix[seg + 1] <= rem) ++seg;
    // do not overwrite staging buffer (seq & 1) before the rank consumed call seq-2
    if (a.seq > 2) {
      if (threadIdx.x == 0) (void)spin_until(a.ack[seg], a.seq - 2, status);   // a timeout is recorded in *status
      __syncthreads();
    }
    const size_t shard = (size_t)a.shard_bytes[seg];
    const size_t cb = min(shard, (size_t)c * (size_t)a.chunk_bytes[seg]);
    const size_t ce = min(shard, cb + (size_t)a.chunk_bytes[seg]);
    const size_t off = cb + (size_t)(rem - a.tile_prefix[seg]) * kPushTile;
    if (off < ce) {
      const size_t len = (ce - off) < (size_t)kPushTile ? (ce - off) : (size_t)kPushTile;
      const uint8_t* s = a.src[seg] + off;
      uint8_t* d = a.dst[seg] + off;
      if (((((uintptr_t)s) | ((uintptr_t)d)) & 31) == 0) {
        const size_t nv = len >> 5;   // a full tile = 512 packets = 2 per thread, both loads first
        uint32_t w0[8], w1[8];
        const size_t v0 = threadIdx.x, v1 = threadIdx.x + kPushThreads;
        if (v0 < nv) ldg256_stream(s + (v0 << 5), w0);
        if (v1 < nv) ldg256_stream(s + (v1 << 5), w1);
        if (v0 < nv) stg256(d + (v0 << 5), w0);
        if (v1 < nv) stg256(d + (v1 << 5), w1);
        for (size_t e = (nv << 5) + threadIdx.x; e < len; e += kPushThreads) d[e] = s[e];
      } else {
        for (size_t e = threadIdx.x; e < len; e += kPushThreads) d[e] = s[e];
      }
    }
    ++n_done;
  }
  if (n_done) flush();
}

// Publish-only launch for calls with nothing to move (keeps the flag protocol uniform).
__global__ void push_publish_kernel(const __grid_constant__ PushScatterArgs a) {
  if (threadIdx.x < a.n)
    for (int c = 0; c < a.n_chunks; ++c) st_release_sys(a.ready[threadIdx.x] + c, a.seq);
}

template <int DT, int OP>
__global__ void __launch_bounds__(kPushThreads)
    push_consume_kernel(const uint8_t* stage, uint8_t* dst, size_t shard_bytes, size_t chunk_bytes,
                        uint32_t tiles_per_chunk, uint32_t n_chunks, uint32_t slice, MapParams p,
                        const unsigned long long* ready, unsigned long long seq, unsigned long long* ack,
                        unsigned int* ticket, unsigned int* status) {
  constexpr size_t ES = (DT == KTB_U8) ? 1 : ((DT == KTB_BF16 || DT == KTB_F16) ? 2 : (DT == KTB_I64 ? 8 : 4));
  const uint32_t total_tiles = tiles_per_chunk * n_chunks;
  __shared__ bool flag;
  uint32_t have = 0;   // chunks [0, have) are known to have landed
  // a slice of `slice` consecutive tiles per CTA (1 by default), tiles in increasing order = chunk order; one system
  // fence per CTA before the ack ticket
  for (uint32_t base = blockIdx.x * slice; base < total_tiles; base += gridDim.x * slice)
  for (uint32_t tile = base; tile < min(total_tiles, base + slice); ++tile) {
    const uint32_t c = tile / tiles_per_chunk;
    const uint32_t t = tile % tiles_per_chunk;
    if (c >= have) {
      if (threadIdx.x == 0) flag = spin_until(ready + c, seq, status);
      __syncthreads();
      // the piece never arrived (the wait timed out and raised the sticky status word): do NOT map stale staging
      // data into the caller's result and do NOT acknowledge — the root's wait then times out too and the host maps
      // the status to PodTerminatedError; the session is torn down, so the unbalanced ticket does not matter
      if (!flag) return;
      have = c + 1;
      __syncthreads();   // `flag` is reused
    }
    const size_t cb = min(shard_bytes, (size_t)c * chunk_bytes);
    const size_t ce = min(shard_bytes, cb + chunk_bytes);
    const size_t off = cb + (size_t)t * kPushTile;
    if (off < ce) {
      const size_t len = (ce - off) < (size_t)kPushTile ? (ce - off) : (size_t)kPushTile;
      const uint8_t* s = stage + off;
      uint8_t* d = dst + off;
      if (((((uintptr_t)s) | ((uintptr_t)d)) & 31) == 0) {
        const size_t nv = len >> 5;
        uint32_t w0[8], w1[8];
        const size_t v0 = threadIdx.x, v1 = threadIdx.x + kPushThreads;
        if (v0 < nv) ldg256_stream(s + (v0 << 5), w0);
        if (v1 < nv) ldg256_stream(s + (v1 << 5), w1);
        if (v0 < nv) {
          apply_words<DT, OP, 8>(w0, p);
          stg256(d + (v0 << 5), w0);
        }
        if (v1 < nv) {
          apply_words<DT, OP, 8>(w1, p);
          stg256(d + (v1 << 5), w1);
        }
        const size_t tail = nv << 5;
        for (size_t e = threadIdx.x; e < (len - tail) / ES; e += kPushThreads)
          apply_elem<DT, OP>(s + tail + e * ES, d + tail + e * ES, p);
      } else {
        for (size_t e = threadIdx.x; e < len / ES; e += kPushThreads) apply_elem<DT, OP>(s + e * ES, d + e * ES, p);
      }
    }
  }
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence_system();
    st_release_sys(ack, seq);   // results of the whole shard are in the root's memory
    *ticket = 0;
  }
}

__global__ void push_wait_kernel(const unsigned long long* ack, int n_ranks, int root_rank, unsigned long long seq,
                                 unsigned int* status) {
  const int r = threadIdx.x;
  if (r < n_ranks && r != root_rank) spin_until(&ack[r], seq, status);
}

// Chunk c of a shard of `shard_elems` elements: element range [b, e), chunk size rounded to 64 elements.
static inline void chunk_bounds(size_t shard_elems, int n_chunks, int c, size_t* b, size_t* e) {
  size_t per = (shard_elems + (size_t)n_chunks - 1) / (size_t)n_chunks;
  per = (per + 63) / 64 * 64;
  size_t lo = std::min(shard_elems, per * (size_t)c);
  size_t hi = std::min(shard_elems, lo + per);
  *b = lo;
  *e = hi;
}

}  // namespace ktb

using namespace ktb;

extern "C" {

size_t ktb_push_control_bytes(void) { return 4096; }

// (control-block layout: KTB_CTRL_* in ktb_common.cuh)

// grid cap of the scatter in CTAs per SM (0 = no cap: ceil(tiles / slice) CTAs).  Caps were measured slower for the
// element-wise call (profiles/r2_summary.md §2); ktb_push_scatter_chunked takes its own cap for the MLP.
std::atomic<int> g_push_scatter_ctas_per_sm{0};   // ktb_set_tuning(21, n)
// tiles (16 KiB) per CTA per system fence, both sides.  Measured (profiles/r2f_push_sweep_2gpu.jsonl): 1 is best —
// 256 MiB at N=2: slice 1 0.317 ms, 2 0.377, 4 0.409, 8-16 0.413; many short-lived CTAs keep more stores in flight
// than fewer CTAs streaming longer slices, and the per-CTA fence is not what limits them
std::atomic<int> g_push_slice{1};                 // ktb_set_tuning(23, n)

// chunk_elems == 0: n_chunks pieces per shard (chunk_bounds); otherwise pieces of exactly chunk_elems elements
// (the consumer of ktb_mlp_bf16_pushed wants whole GEMM row chunks), n_chunks = ceil(largest shard / chunk_elems).
static int push_scatter_impl(const char* who, int root_dev, const void* src_root, size_t n_elems, size_t granule, int dtype,
                             int n_ranks, int root_rank, void* const* stage_peer, size_t stage_stride,
                             void* const* ctrl_peer, void* ctrl_root, int n_chunks, size_t chunk_elems,
                             unsigned long long seq, int ctas_per_sm, uintptr_t stream) {
  int rc = require_device(root_dev);
  if (rc) return rc;
  const size_t es = dtype_size(dtype);
  KTB_REQUIRE(es != 0, KTB_ERR_ARG, "%s: unknown dtype %d", who, dtype);
  KTB_REQUIRE(n_ranks > 0 && n_ranks <= kPushMaxRanks && root_rank >= 0 && root_rank < n_ranks, KTB_ERR_ARG,
              "%s: bad ranks %d/%d", who, root_rank, n_ranks);
  KTB_REQUIRE(src_root && stage_peer && ctrl_peer && ctrl_root && seq > 0, KTB_ERR_ARG, "%s: null argument", who);
  KTB_REQUIRE(granule > 0 && n_elems % granule == 0, KTB_ERR_ARG, "%s: n_elems not a multiple of granule", who);
  if (chunk_elems) {
    size_t b0 = 0, e0 = 0;
    ktb_shard_bounds(n_elems / granule, n_ranks, 0, &b0, &e0);   // rank 0 holds the largest shard
    n_chunks = (int)std::max<size_t>(1, ((e0 - b0) * granule + chunk_elems - 1) / chunk_elems);
  }
  KTB_REQUIRE(n_chunks > 0 && n_chunks <= KTB_PUSH_MAX_CHUNKS, KTB_ERR_ARG, "%s: n_chunks %d out of range", who, n_chunks);
  KTB_GUARD(root_dev);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* croot = static_cast<uint8_t*>(ctrl_root);
  // per call parity: the scatters of two consecutive calls may overlap when the caller alternates streams
  unsigned int* chunk_done =
      reinterpret_cast<unsigned int*>(croot + KTB_CTRL_CHUNK_DONE) + (size_t)(seq & 1) * KTB_PUSH_MAX_CHUNKS;
  unsigned int* status = reinterpret_cast<unsigned int*>(croot + KTB_CTRL_STATUS);
  const size_t buf_off = (size_t)(seq & 1) * stage_stride;
  PushScatterArgs a;
  a.n = 0;
  a.n_chunks = n_chunks;
  a.seq = seq;
  a.tile_prefix[0] = 0;
  for (int r = 0; r < n_ranks; ++r) {
    if (r == root_rank) continue;
    size_t sb = 0, se = 0, c0b = 0, c0e = 0, per = chunk_elems, dummy = 0;
    ktb_shard_bounds(n_elems / granule, n_ranks, r, &sb, &se);
    sb *= granule;
    se *= granule;
    if (!chunk_elems) {
      chunk_bounds(se - sb, n_chunks, 0, &c0b, &c0e);
      chunk_bounds(se - sb, n_chunks, 1, &per, &dummy);   // start of chunk 1 = elements per chunk
      if (per == 0) per = c0e - c0b;                      // single-chunk (or empty) shard
    }
    if (!stage_peer[r]) continue;   // not served by this call (a hybrid scatter splits the ranks between two engines)
    KTB_REQUIRE(ctrl_peer[r], KTB_ERR_ARG, "%s: rank %d has no control block", who, r);
    KTB_REQUIRE((se - sb) * es <= stage_stride, KTB_ERR_ARG, "%s: shard of rank %d exceeds stage_stride", who, r);
    const int i = a.n++;
    a.src[i] = static_cast<const uint8_t*>(src_root) + sb * es;
    a.dst[i] = static_cast<uint8_t*>(stage_peer[r]) + buf_off;
    a.shard_bytes[i] = (se - sb) * es;
    a.chunk_bytes[i] = per * es;
    a.tile_prefix[i + 1] = a.tile_prefix[i] + (uint32_t)((per * es + kPushTile - 1) / kPushTile);
    a.ready[i] = reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(ctrl_peer[r]) + KTB_CTRL_READY);
    a.ack[i] = reinterpret_cast<const unsigned long long*>(croot + KTB_CTRL_ACK) + r;
  }
  if (a.n == 0) return KTB_OK;
  const uint32_t tiles_per_chunk = a.tile_prefix[a.n];
  if (tiles_per_chunk == 0) {
    push_publish_kernel<<<1, 32, 0, st>>>(a);
  } else {
    const uint32_t slice = (uint32_t)std::max(1, std::min<int>(g_push_slice.load(), (int)tiles_per_chunk));
    const size_t total = ((size_t)tiles_per_chunk * (size_t)n_chunks + slice - 1) / slice;
    if (ctas_per_sm <= 0) ctas_per_sm = g_push_scatter_ctas_per_sm.load();
    const size_t cap = ctas_per_sm > 0 ? (size_t)device_info(root_dev)->sm_count * (size_t)ctas_per_sm : total;
    push_scatter_kernel<<<(unsigned)std::min(total, cap), kPushThreads, 0, st>>>(a, chunk_done, status, slice);
  }
  KTB_CK(cudaGetLastError());
  return KTB_OK;
}

int ktb_push_scatter(int root_dev, const void* src_root, size_t n_elems, size_t granule, int dtype, int n_ranks,
                     int root_rank, void* const* stage_peer, size_t stage_stride, void* const* ctrl_peer,
                     void* ctrl_root, int n_chunks, unsigned long long seq, uintptr_t stream) {
  return push_scatter_impl("ktb_push_scatter", root_dev, src_root, n_elems, granule, dtype, n_ranks, root_rank, stage_peer,
                           stage_stride, ctrl_peer, ctrl_root, n_chunks, 0, seq, 0, stream);
}

int ktb_push_scatter_chunked(int root_dev, const void* src_root, size_t n_elems, size_t granule, int dtype, int n_ranks,
                             int root_rank, void* const* stage_peer, size_t stage_stride, void* const* ctrl_peer,
                             void* ctrl_root, size_t chunk_elems, int ctas_per_sm, unsigned long long seq,
                             uintptr_t stream) {
  KTB_REQUIRE(chunk_elems > 0, KTB_ERR_ARG, "ktb_push_scatter_chunked: chunk_elems must be positive");
  return push_scatter_impl("ktb_push_scatter_chunked", root_dev, src_root, n_elems, granule, dtype, n_ranks, root_rank,
                           stage_peer, stage_stride, ctrl_peer, ctrl_root, 0, chunk_elems, seq, ctas_per_sm, stream);
}

// ---- copy-engine form of the root side: no SM of the root is used ------------------------------------------------
namespace ktb {
__global__ void push_publish_one_kernel(unsigned long long* ready, unsigned long long seq) {
  if (threadIdx.x == 0) {
    __threadfence_system();
    st_release_sys(ready, seq);
  }
}
__global__ void push_wait_ack_kernel(const unsigned long long* ack, unsigned long long want, unsigned int* status) {
  if (threadIdx.x == 0) (void)spin_until(ack, want, status);
}
}  // namespace ktb

int ktb_push_scatter_ce(int root_dev, const void* src_root, size_t n_elems, size_t granule, int dtype, int n_ranks,
                        int root_rank, const int* devs, void* const* stage_peer, size_t stage_stride,
                        void* const* ctrl_peer, void* ctrl_root, size_t chunk_elems, unsigned long long seq,
                        uintptr_t stream) {
  int rc = require_device(root_dev);
  if (rc) return rc;
  const size_t es = dtype_size(dtype);
  KTB_REQUIRE(es != 0, KTB_ERR_ARG, "ktb_push_scatter_ce: unknown dtype %d", dtype);
  KTB_REQUIRE(n_ranks > 0 && n_ranks <= kPushMaxRanks && root_rank >= 0 && root_rank < n_ranks && devs, KTB_ERR_ARG,
              "ktb_push_scatter_ce: bad ranks %d/%d", root_rank, n_ranks);
  KTB_REQUIRE(src_root && stage_peer && ctrl_peer && ctrl_root && seq > 0 && chunk_elems > 0, KTB_ERR_ARG,
              "ktb_push_scatter_ce: null/zero argument");
  KTB_REQUIRE(granule > 0 && n_elems % granule == 0, KTB_ERR_ARG, "ktb_push_scatter_ce: n_elems not a multiple of granule");
  size_t b0 = 0, e0 = 0;
  ktb_shard_bounds(n_elems / granule, n_ranks, 0, &b0, &e0);
  const size_t n_chunks = std::max<size_t>(1, ((e0 - b0) * granule + chunk_elems - 1) / chunk_elems);
  KTB_REQUIRE(n_chunks <= KTB_PUSH_MAX_CHUNKS, KTB_ERR_ARG, "ktb_push_scatter_ce: %zu chunks exceed %d", n_chunks,
              KTB_PUSH_MAX_CHUNKS);
  KTB_GUARD(root_dev);
  // one library stream (and fork/join event pair) per destination rank on the root device, created on first use:
  // every peer's copies ride their own copy-engine queue, chunk-major issue order
  static std::mutex mu;
  static cudaStream_t ce_stream[kMaxDevices][kPushMaxRanks] = {};
  static cudaEvent_t ce_done[kMaxDevices][kPushMaxRanks] = {};
  static cudaEvent_t ce_start[kMaxDevices] = {};
  std::lock_guard<std::mutex> lk(mu);   // also serialises concurrent CE scatters of one root (they share the streams)
  if (!ce_start[root_dev]) KTB_CK(cudaEventCreateWithFlags(&ce_start[root_dev], cudaEventDisableTiming));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* croot = static_cast<uint8_t*>(ctrl_root);
  unsigned int* status = reinterpret_cast<unsigned int*>(croot + KTB_CTRL_STATUS);
  const size_t buf_off = (size_t)(seq & 1) * stage_stride;
  KTB_CK(cudaEventRecord(ce_start[root_dev], st));   // the observations are ready once prior work on `st` is done
  size_t sb[kPushMaxRanks], sl[kPushMaxRanks];
  for (int r = 0; r < n_ranks; ++r) {
    size_t b = 0, e = 0;
    ktb_shard_bounds(n_elems / granule, n_ranks, r, &b, &e);
    sb[r] = b * granule;
    sl[r] = (e - b) * granule;
    if (r == root_rank || !stage_peer[r]) continue;   // null staging = rank served by another scatter call
    KTB_REQUIRE(ctrl_peer[r], KTB_ERR_ARG, "ktb_push_scatter_ce: rank %d has no control block", r);
    KTB_REQUIRE(sl[r] * es <= stage_stride, KTB_ERR_ARG, "ktb_push_scatter_ce: shard of rank %d exceeds stage_stride", r);
    if (!ce_stream[root_dev][r]) {
      KTB_CK(cudaStreamCreateWithFlags(&ce_stream[root_dev][r], cudaStreamNonBlocking));
      KTB_CK(cudaEventCreateWithFlags(&ce_done[root_dev][r], cudaEventDisableTiming));
    }
    KTB_CK(cudaStreamWaitEvent(ce_stream[root_dev][r], ce_start[root_dev], 0));
    if (seq > 2) {   // do not overwrite staging half (seq & 1) before the rank consumed call seq-2
      push_wait_ack_kernel<<<1, 32, 0, ce_stream[root_dev][r]>>>(
          reinterpret_cast<const unsigned long long*>(croot + KTB_CTRL_ACK) + r, seq - 2, status);
      KTB_CK(cudaGetLastError());
    }
  }
  for (size_t c = 0; c < n_chunks; ++c) {
    for (int r = 0; r < n_ranks; ++r) {
      if (r == root_rank || !stage_peer[r]) continue;
      cudaStream_t cs = ce_stream[root_dev][r];
      const size_t lo = std::min(sl[r], c * chunk_elems), hi = std::min(sl[r], lo + chunk_elems);
      if (hi > lo)
        KTB_CK(cudaMemcpyPeerAsync(static_cast<uint8_t*>(stage_peer[r]) + buf_off + lo * es, devs[r],
                                   static_cast<const uint8_t*>(src_root) + (sb[r] + lo) * es, root_dev, (hi - lo) * es, cs));
      unsigned long long* ready =
          reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(ctrl_peer[r]) + KTB_CTRL_READY) + c;
      push_publish_one_kernel<<<1, 32, 0, cs>>>(ready, seq);   // stream-ordered behind the copy
      KTB_CK(cudaGetLastError());
    }
  }
  for (int r = 0; r < n_ranks; ++r) {   // the caller's stream "contains" the scatter (src may be reused after it)
    if (r == root_rank || !stage_peer[r]) continue;
    KTB_CK(cudaEventRecord(ce_done[root_dev][r], ce_stream[root_dev][r]));
    KTB_CK(cudaStreamWaitEvent(st, ce_done[root_dev][r], 0));
  }
  return KTB_OK;
}

int ktb_push_consume(int dev, int op, int dtype, const void* stage_local, size_t stage_stride, void* dst_root_shard,
                     size_t shard_elems, double alpha, double beta, void* ctrl_local, void* ctrl_root_peer, int rank,
                     int n_chunks, unsigned long long seq, uintptr_t stream) {
  int rc = require_device(dev);
  if (rc) return rc;
  const size_t es = dtype_size(dtype);
  KTB_REQUIRE(es != 0, KTB_ERR_ARG, "ktb_push_consume: unknown dtype %d", dtype);
  KTB_REQUIRE(op >= KTB_OP_IDENTITY && op <= KTB_OP_AFFINE && !(dtype == KTB_U8 && op != KTB_OP_IDENTITY), KTB_ERR_ARG,
              "ktb_push_consume: bad op/dtype %d/%d", op, dtype);
  KTB_REQUIRE(n_chunks > 0 && n_chunks <= KTB_PUSH_MAX_CHUNKS && rank >= 0 && rank < kPushMaxRanks && seq > 0, KTB_ERR_ARG,
              "ktb_push_consume: bad chunk/rank arguments");
  KTB_REQUIRE(stage_local && ctrl_local && ctrl_root_peer && (dst_root_shard || shard_elems == 0), KTB_ERR_ARG,
              "ktb_push_consume: null argument");
  KTB_GUARD(dev);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const MapParams p = make_params(alpha, beta, dtype);
  uint8_t* cl = static_cast<uint8_t*>(ctrl_local);
  const unsigned long long* ready = reinterpret_cast<const unsigned long long*>(cl + KTB_CTRL_READY);
  unsigned int* ticket = reinterpret_cast<unsigned int*>(cl + KTB_CTRL_TICKET);
  unsigned int* status = reinterpret_cast<unsigned int*>(cl + KTB_CTRL_STATUS);
  unsigned long long* ack =
      reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(ctrl_root_peer) + KTB_CTRL_ACK) + rank;
  const uint8_t* stage = static_cast<const uint8_t*>(stage_local) + (size_t)(seq & 1) * stage_stride;
  size_t per = 0, dummy = 0, c0b = 0, c0e = 0;
  chunk_bounds(shard_elems, n_chunks, 0, &c0b, &c0e);
  chunk_bounds(shard_elems, n_chunks, 1, &per, &dummy);
  if (per == 0) per = c0e - c0b;
  const size_t chunk_bytes = per * es, shard_bytes = shard_elems * es;
  const uint32_t tpc = (uint32_t)std::max<size_t>(1, (chunk_bytes + kPushTile - 1) / kPushTile);
  const uint32_t slice = (uint32_t)std::max(1, std::min<int>(g_push_slice.load(), (int)tpc));
  const unsigned grid = (unsigned)(((size_t)tpc * (size_t)n_chunks + slice - 1) / slice);
  uint8_t* d = static_cast<uint8_t*>(dst_root_shard);
#define KTB_PC(DT, OPC)                                                                                       \
  push_consume_kernel<DT, OPC><<<grid, kPushThreads, 0, st>>>(stage, d, shard_bytes, chunk_bytes, tpc, (uint32_t)n_chunks, slice, p, \
                                                              ready, seq, ack, ticket, status)
  if (op == KTB_OP_IDENTITY) {
    KTB_PC(KTB_U8, KTB_OP_IDENTITY);
  } else {
    switch (dtype) {
      case KTB_F32: if (op == KTB_OP_SCALE) KTB_PC(KTB_F32, KTB_OP_SCALE); else KTB_PC(KTB_F32, KTB_OP_AFFINE); break;
      case KTB_BF16: if (op == KTB_OP_SCALE) KTB_PC(KTB_BF16, KTB_OP_SCALE); else KTB_PC(KTB_BF16, KTB_OP_AFFINE); break;
      case KTB_I32: if (op == KTB_OP_SCALE) KTB_PC(KTB_I32, KTB_OP_SCALE); else KTB_PC(KTB_I32, KTB_OP_AFFINE); break;
      case KTB_F16: if (op == KTB_OP_SCALE) KTB_PC(KTB_F16, KTB_OP_SCALE); else KTB_PC(KTB_F16, KTB_OP_AFFINE); break;
      default: if (op == KTB_OP_SCALE) KTB_PC(KTB_I64, KTB_OP_SCALE); else KTB_PC(KTB_I64, KTB_OP_AFFINE); break;
    }
  }
#undef KTB_PC
  KTB_CK(cudaGetLastError());
  return KTB_OK;
}

int ktb_push_wait(int root_dev, void* ctrl_root, int n_ranks, int root_rank, unsigned long long seq, uintptr_t stream) {
  int rc = require_device(root_dev);
  if (rc) return rc;
  KTB_REQUIRE(ctrl_root && n_ranks > 0 && n_ranks <= kPushMaxRanks, KTB_ERR_ARG, "ktb_push_wait: bad arguments");
  KTB_GUARD(root_dev);
  uint8_t* croot = static_cast<uint8_t*>(ctrl_root);
  push_wait_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const unsigned long long*>(croot + KTB_CTRL_ACK), n_ranks, root_rank, seq,
      reinterpret_cast<unsigned int*>(croot + KTB_CTRL_STATUS));
  KTB_CK(cudaGetLastError());
  return KTB_OK;
}

// Reads (synchronously) the sticky timeout status of a control block: 0 = healthy.
int ktb_push_status(int dev, const void* ctrl, unsigned int* out) {
  int rc = require_device(dev);
  if (rc) return rc;
  KTB_REQUIRE(ctrl && out, KTB_ERR_ARG, "ktb_push_status: null argument");
  KTB_GUARD(dev);
  KTB_CK(cudaMemcpy(out, static_cast<const uint8_t*>(ctrl) + KTB_CTRL_STATUS, sizeof(unsigned int),
                    cudaMemcpyDeviceToHost));
  return KTB_OK;
}

}  // extern "C"
