// ktb_dispatch.cu — multi-GPU data movement of the remote-map path over NVLink 5 / NVSwitch,
// and the host-resident (PCIe) form of the call.
//
// Replaces the reference's fan-out/fan-in:
//   broadcast of the same params to every rank   kt/serving/spmd/spmd_supervisor.py:341,439-455
//   per-rank mp.Queue put / HTTP POST            kt/serving/process_pool.py:125-212,
//                                                kt/serving/remote_worker_pool.py:254-316
//   concatenation of per-rank results            kt/serving/spmd/spmd_supervisor.py:547-570
// On this route the "wire" is peer-mapped HBM: a rank's kernel loads its shard directly from
// the root GPU's arg arena and stores its result directly into the root's result arena, so
// scatter, exec and gather are one kernel per rank and root HBM is read once / written once.
#include "ktb_common.cuh"

#include <algorithm>
#include <mutex>

namespace ktb {


// ---- broadcast: one read, n peer stores ---------------------------------------------------------
constexpr int kBcastMax = 15;
constexpr int kBcastThreads = 256;
constexpr int kBcastUnroll = 2;

struct BcastDsts {
  uint8_t* d[kBcastMax];
  int n;
};

__global__ void __launch_bounds__(kBcastThreads)
    bcast_kernel(const uint8_t* src, const __grid_constant__ BcastDsts dsts, size_t n_bytes) {
  constexpr size_t VB = 32;
  constexpr size_t ROW = (size_t)kBcastThreads * VB;
  constexpr size_t TILE = ROW * kBcastUnroll;
  const size_t n_full = n_bytes / TILE;
  for (size_t t = blockIdx.x; t < n_full; t += gridDim.x) {
    const size_t off = t * TILE + (size_t)threadIdx.x * VB;
    uint32_t w[kBcastUnroll][8];
#pragma unroll
    for (int j = 0; j < kBcastUnroll; ++j) ldg256(src + off + j * ROW, w[j]);
    for (int k = 0; k < dsts.n; ++k) {
#pragma unroll
      for (int j = 0; j < kBcastUnroll; ++j) stg256(dsts.d[k] + off + j * ROW, w[j]);
    }
  }
  if (blockIdx.x == (unsigned)(n_full % gridDim.x)) {
    const size_t base = n_full * TILE;
    const size_t n_vec = (n_bytes - base) / VB;
    for (size_t v = threadIdx.x; v < n_vec; v += kBcastThreads) {
      uint32_t w[8];
      ldg256(src + base + v * VB, w);
      for (int k = 0; k < dsts.n; ++k) stg256(dsts.d[k] + base + v * VB, w);
    }
    const size_t tail = base + n_vec * VB;
    for (size_t e = tail + threadIdx.x; e < n_bytes; e += kBcastThreads) {
      uint8_t b = src[e];
      for (int k = 0; k < dsts.n; ++k) dsts.d[k][e] = b;
    }
  }
}

// Unaligned fallback.
__global__ void __launch_bounds__(256)
    bcast_bytes_kernel(const uint8_t* src, const __grid_constant__ BcastDsts dsts, size_t n_bytes) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_bytes; e += stride) {
    uint8_t b = src[e];
    for (int k = 0; k < dsts.n; ++k) dsts.d[k][e] = b;
  }
}

}  // namespace ktb

using namespace ktb;

extern "C" {

int ktb_broadcast(int root, const void* src, void* const* dsts, int n_dst, size_t nbytes,
                  uintptr_t stream) {
  int rc = require_device(root);
  if (rc) return rc;
  KTB_REQUIRE(n_dst >= 0 && n_dst <= kBcastMax, KTB_ERR_ARG, "ktb_broadcast: n_dst %d out of range [0,%d]",
              n_dst, kBcastMax);
  if (nbytes == 0 || n_dst == 0) return KTB_OK;
  KTB_REQUIRE(src && dsts, KTB_ERR_ARG, "ktb_broadcast: null argument");
  BcastDsts b;
  b.n = 0;
  uintptr_t align = (uintptr_t)src;
  for (int k = 0; k < n_dst; ++k) {
    KTB_REQUIRE(dsts[k], KTB_ERR_ARG, "ktb_broadcast: dsts[%d] is null", k);
    if (dsts[k] == src) continue;  // the root's own copy
    b.d[b.n++] = static_cast<uint8_t*>(dsts[k]);
    align |= (uintptr_t)dsts[k];
  }
  if (b.n == 0) return KTB_OK;
  KTB_GUARD(root);
  const DeviceInfo* di = device_info(root);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((align & 31) == 0) {
    size_t tiles = nbytes / ((size_t)kBcastThreads * 32 * kBcastUnroll);
    int grid = (int)std::min<size_t>(std::max<size_t>(tiles, 1), (size_t)di->sm_count * 4);
    bcast_kernel<<<grid, kBcastThreads, 0, st>>>(static_cast<const uint8_t*>(src), b, nbytes);
  } else {
    size_t blocks = (nbytes + 255) / 256;
    int grid = (int)std::min<size_t>(std::max<size_t>(blocks, 1), (size_t)di->sm_count * 8);
    bcast_bytes_kernel<<<grid, 256, 0, st>>>(static_cast<const uint8_t*>(src), b, nbytes);
  }
  KTB_CK(cudaGetLastError());
  return KTB_OK;
}

// Per-call events (created on the device that records them, destroyed right after the waits are enqueued:
// CUDA releases an event's resources once pending work on it has completed). Shared per-device events would let two
// host threads using different streams wait on each other's records.
struct CallEvents {
  cudaEvent_t ev[kMaxDevices + 1];
  int dev[kMaxDevices + 1];
  int n = 0;
  cudaEvent_t make(int device) {
    DeviceGuard g(device);
    cudaEvent_t e = nullptr;
    if (n > kMaxDevices || !g.ok || cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    ev[n] = e;
    dev[n] = device;
    ++n;
    return e;
  }
  ~CallEvents() {
    for (int i = 0; i < n; ++i) {
      DeviceGuard g(dev[i]);
      cudaEventDestroy(ev[i]);
    }
  }
};

static int check_ranks(const char* who, int n_ranks, const int* devs, int root_rank) {
  KTB_REQUIRE(n_ranks > 0 && n_ranks <= kMaxDevices && devs, KTB_ERR_ARG, "%s: bad n_ranks %d", who, n_ranks);
  KTB_REQUIRE(root_rank >= 0 && root_rank < n_ranks, KTB_ERR_ARG, "%s: root_rank %d out of range", who, root_rank);
  for (int r = 0; r < n_ranks; ++r) {
    int rc = require_device(devs[r]);
    if (rc) return rc;
    if (devs[r] != devs[root_rank]) {
      KTB_REQUIRE(ktb_peer_enabled(devs[r], devs[root_rank]) == 1, KTB_ERR_UNSUPPORTED,
                  "%s: device %d has no peer access to root device %d", who, devs[r], devs[root_rank]);
    }
  }
  return KTB_OK;
}

int ktb_scatter_map_gather(int op, int dtype, const void* src_root, void* dst_root, size_t n_elems,
                           size_t granule, double alpha, double beta, int n_ranks, const int* devs,
                           int root_rank, int variant, const uintptr_t* streams) {
  int rc = check_ranks("ktb_scatter_map_gather", n_ranks, devs, root_rank);
  if (rc) return rc;
  const size_t es = dtype_size(dtype);
  KTB_REQUIRE(es != 0, KTB_ERR_ARG, "ktb_scatter_map_gather: unknown dtype %d", dtype);
  if (n_elems == 0) return KTB_OK;
  KTB_REQUIRE(src_root && dst_root, KTB_ERR_ARG, "ktb_scatter_map_gather: null src/dst");
  KTB_REQUIRE(granule > 0 && n_elems % granule == 0, KTB_ERR_ARG,
              "ktb_scatter_map_gather: n_elems %zu is not a multiple of granule %zu", n_elems, granule);
  const MapParams p = make_params(alpha, beta, dtype);
  const int root_dev = devs[root_rank];
  DeviceInfo* root = device_info(root_dev);
  // streams == NULL → library streams; otherwise streams[r] verbatim (0 is the legacy default stream)
  auto stream_of = [&](int r) {
    return streams ? reinterpret_cast<cudaStream_t>(streams[r]) : device_info(devs[r])->stream_rank;
  };
  cudaStream_t root_stream = stream_of(root_rank);
  CallEvents events;
  cudaEvent_t done[kMaxDevices] = {nullptr};
  cudaEvent_t args_ready = events.make(root_dev);
  KTB_REQUIRE(args_ready, KTB_ERR_CUDA, "ktb_scatter_map_gather: cudaEventCreate failed");
  (void)root;
  {
    KTB_GUARD(root_dev);
    KTB_CK(cudaEventRecord(args_ready, root_stream));  // args are ready on the root
  }
  for (int r = 0; r < n_ranks; ++r) {
    size_t b = 0, e = 0;
    ktb_shard_bounds(n_elems / granule, n_ranks, r, &b, &e);
    b *= granule;
    e *= granule;
    if (e == b) continue;
    const int dev = devs[r];
    KTB_GUARD(dev);
    cudaStream_t st = stream_of(r);
    // same ordering domain as the root only if it is the same stream ON the same device (handle 0 is
    // "the default stream of whichever device is current", so handles alone do not identify a stream)
    const bool is_root = (r == root_rank) || (dev == root_dev && st == root_stream);
    if (!is_root) KTB_CK(cudaStreamWaitEvent(st, args_ready, 0));
    rc = launch_map(dev, op, dtype, static_cast<const uint8_t*>(src_root) + b * es,
                    static_cast<uint8_t*>(dst_root) + b * es, e - b, p, variant, st);
    if (rc) return rc;
    if (!is_root) {
      done[r] = events.make(dev);
      KTB_REQUIRE(done[r], KTB_ERR_CUDA, "ktb_scatter_map_gather: cudaEventCreate failed");
      KTB_CK(cudaEventRecord(done[r], st));
    }
  }
  // Join on the root stream with the ROOT device current: stream handle 0 names the default stream of
  // whichever device is current, so the wait must be issued under the root's guard.
  {
    KTB_GUARD(root_dev);
    for (int r = 0; r < n_ranks; ++r)
      if (done[r]) KTB_CK(cudaStreamWaitEvent(root_stream, done[r], 0));
  }
  return KTB_OK;
}

int ktb_scatter_map_reduce(int op, int dtype, const void* src_root, size_t n_elems, size_t granule,
                           double alpha, double beta, int n_ranks, const int* devs, int root_rank,
                           void* partials_root, void* out_root, void* const* workspaces,
                           const uintptr_t* streams) {
  int rc = check_ranks("ktb_scatter_map_reduce", n_ranks, devs, root_rank);
  if (rc) return rc;
  const size_t es = dtype_size(dtype);
  KTB_REQUIRE(es != 0 && dtype != KTB_U8, KTB_ERR_ARG, "ktb_scatter_map_reduce: dtype %d not reducible", dtype);
  KTB_REQUIRE(partials_root && out_root && workspaces, KTB_ERR_ARG, "ktb_scatter_map_reduce: null argument");
  KTB_REQUIRE(src_root || n_elems == 0, KTB_ERR_ARG, "ktb_scatter_map_reduce: null src");
  KTB_REQUIRE(granule > 0 && n_elems % granule == 0, KTB_ERR_ARG,
              "ktb_scatter_map_reduce: n_elems %zu is not a multiple of granule %zu", n_elems, granule);
  const MapParams p = make_params(alpha, beta, dtype);
  const size_t acc_size = (dtype == KTB_F32 || dtype == KTB_BF16 || dtype == KTB_F16) ? 4 : 8;
  const int root_dev = devs[root_rank];
  DeviceInfo* root = device_info(root_dev);
  // streams == NULL → library streams; otherwise streams[r] verbatim (0 is the legacy default stream)
  auto stream_of = [&](int r) {
    return streams ? reinterpret_cast<cudaStream_t>(streams[r]) : device_info(devs[r])->stream_rank;
  };
  cudaStream_t root_stream = stream_of(root_rank);
  CallEvents events;
  cudaEvent_t done[kMaxDevices] = {nullptr};
  cudaEvent_t args_ready = events.make(root_dev);
  KTB_REQUIRE(args_ready, KTB_ERR_CUDA, "ktb_scatter_map_reduce: cudaEventCreate failed");
  (void)root;
  {
    KTB_GUARD(root_dev);
    KTB_CK(cudaEventRecord(args_ready, root_stream));
  }
  for (int r = 0; r < n_ranks; ++r) {
    size_t b = 0, e = 0;
    ktb_shard_bounds(n_elems / granule, n_ranks, r, &b, &e);
    b *= granule;
    e *= granule;
    const int dev = devs[r];
    KTB_GUARD(dev);
    cudaStream_t st = stream_of(r);
    // same ordering domain as the root only if it is the same stream ON the same device (handle 0 is
    // "the default stream of whichever device is current", so handles alone do not identify a stream)
    const bool is_root = (r == root_rank) || (dev == root_dev && st == root_stream);
    if (!is_root) KTB_CK(cudaStreamWaitEvent(st, args_ready, 0));
    KTB_REQUIRE(workspaces[r], KTB_ERR_ARG, "ktb_scatter_map_reduce: workspaces[%d] is null", r);
    // empty shards still write a zero partial (n_elems = 0 → kernel stores 0)
    rc = launch_map_reduce(dev, op, dtype, static_cast<const uint8_t*>(src_root) + b * es, e - b, p,
                           static_cast<uint8_t*>(partials_root) + (size_t)r * acc_size, workspaces[r], st);
    if (rc) return rc;
    if (!is_root) {
      done[r] = events.make(dev);
      KTB_REQUIRE(done[r], KTB_ERR_CUDA, "ktb_scatter_map_reduce: cudaEventCreate failed");
      KTB_CK(cudaEventRecord(done[r], st));
    }
  }
  KTB_GUARD(root_dev);
  for (int r = 0; r < n_ranks; ++r)
    if (done[r]) KTB_CK(cudaStreamWaitEvent(root_stream, done[r], 0));
  return launch_reduce_partials(root_dev, dtype, partials_root, n_ranks, out_root, root_stream);
}

int ktb_map_host(int dev, int op, int dtype, const void* src_host, void* dst_host, size_t n_elems,
                 double alpha, double beta, size_t chunk_bytes, void* stage_in, void* stage_out) {
  int rc = require_device(dev);
  if (rc) return rc;
  const size_t es = dtype_size(dtype);
  KTB_REQUIRE(es != 0, KTB_ERR_ARG, "ktb_map_host: unknown dtype %d", dtype);
  if (n_elems == 0) return KTB_OK;
  KTB_REQUIRE(src_host && dst_host && stage_in && stage_out, KTB_ERR_ARG, "ktb_map_host: null argument");
  KTB_REQUIRE(chunk_bytes >= 4096 && chunk_bytes % 256 == 0, KTB_ERR_ARG,
              "ktb_map_host: chunk_bytes must be a multiple of 256 and >= 4096 (got %zu)", chunk_bytes);
  KTB_GUARD(dev);
  DeviceInfo* di = device_info(dev);
  const MapParams p = make_params(alpha, beta, dtype);
  const size_t n_bytes = n_elems * es;
  const size_t n_chunks = (n_bytes + chunk_bytes - 1) / chunk_bytes;
  CallEvents events;   // per-call events, destroyed on every exit path
  cudaEvent_t h2d_done[2], exec_done[2], d2h_done[2];
  for (int i = 0; i < 2; ++i) {
    h2d_done[i] = events.make(dev);
    exec_done[i] = events.make(dev);
    d2h_done[i] = events.make(dev);
    KTB_REQUIRE(h2d_done[i] && exec_done[i] && d2h_done[i], KTB_ERR_CUDA, "ktb_map_host: cudaEventCreate failed");
  }
  int status = KTB_OK;
  for (size_t c = 0; c < n_chunks && status == KTB_OK; ++c) {
    const int b = (int)(c & 1);
    const size_t off = c * chunk_bytes;
    const size_t len = std::min(chunk_bytes, n_bytes - off);
    uint8_t* sin = static_cast<uint8_t*>(stage_in) + (size_t)b * chunk_bytes;
    uint8_t* sout = static_cast<uint8_t*>(stage_out) + (size_t)b * chunk_bytes;
    cudaError_t e = cudaSuccess;
    if (c >= 2) e = cudaStreamWaitEvent(di->stream_h2d, exec_done[b], 0);  // stage_in[b] consumed
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(sin, static_cast<const uint8_t*>(src_host) + off, len, cudaMemcpyHostToDevice,
                          di->stream_h2d);
    if (e == cudaSuccess) e = cudaEventRecord(h2d_done[b], di->stream_h2d);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(di->stream_exec, h2d_done[b], 0);
    if (e == cudaSuccess && c >= 2) e = cudaStreamWaitEvent(di->stream_exec, d2h_done[b], 0);  // stage_out[b] drained
    if (e == cudaSuccess) {
      status = launch_map(dev, op, dtype, sin, sout, len / es, p, KTB_VARIANT_AUTO, di->stream_exec);
      if (status != KTB_OK) break;
      e = cudaEventRecord(exec_done[b], di->stream_exec);
    }
    if (e == cudaSuccess) e = cudaStreamWaitEvent(di->stream_d2h, exec_done[b], 0);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(static_cast<uint8_t*>(dst_host) + off, sout, len, cudaMemcpyDeviceToHost,
                          di->stream_d2h);
    if (e == cudaSuccess) e = cudaEventRecord(d2h_done[b], di->stream_d2h);
    if (e != cudaSuccess) {
      set_error("ktb_map_host: chunk %zu failed: %s", c, cudaGetErrorString(e));
      status = KTB_ERR_CUDA;
    }
  }
  cudaError_t es1 = cudaStreamSynchronize(di->stream_d2h);
  cudaError_t es2 = cudaStreamSynchronize(di->stream_exec);
  cudaError_t es3 = cudaStreamSynchronize(di->stream_h2d);
  if (status == KTB_OK && (es1 != cudaSuccess || es2 != cudaSuccess || es3 != cudaSuccess)) {
    cudaError_t e = es1 != cudaSuccess ? es1 : (es2 != cudaSuccess ? es2 : es3);
    set_error("ktb_map_host: stream sync failed: %s", cudaGetErrorString(e));
    status = KTB_ERR_CUDA;
  }
  return status;
}

int ktb_map_host_multi(int op, int dtype, const void* src_host, void* dst_host, size_t n_elems, size_t granule,
                       double alpha, double beta, int n_ranks, const int* devs, size_t chunk_bytes,
                       void* const* stage_in, void* const* stage_out) {
  const size_t es = dtype_size(dtype);
  KTB_REQUIRE(es != 0, KTB_ERR_ARG, "ktb_map_host_multi: unknown dtype %d", dtype);
  KTB_REQUIRE(n_ranks > 0 && n_ranks <= kMaxDevices && devs && stage_in && stage_out, KTB_ERR_ARG,
              "ktb_map_host_multi: bad rank arguments");
  if (n_elems == 0) return KTB_OK;
  KTB_REQUIRE(src_host && dst_host, KTB_ERR_ARG, "ktb_map_host_multi: null host buffer");
  KTB_REQUIRE(granule > 0 && n_elems % granule == 0, KTB_ERR_ARG, "ktb_map_host_multi: n_elems not a multiple of granule");
  KTB_REQUIRE(chunk_bytes >= 4096 && chunk_bytes % 256 == 0, KTB_ERR_ARG,
              "ktb_map_host_multi: chunk_bytes must be a multiple of 256 and >= 4096 (got %zu)", chunk_bytes);
  for (int r = 0; r < n_ranks; ++r) {
    int rc = require_device(devs[r]);
    if (rc) return rc;
    for (int q = 0; q < r; ++q)
      KTB_REQUIRE(devs[q] != devs[r], KTB_ERR_ARG, "ktb_map_host_multi: devices must be distinct (use ktb_map_host per rank)");
    KTB_REQUIRE(stage_in[r] && stage_out[r], KTB_ERR_ARG, "ktb_map_host_multi: rank %d has no staging buffers", r);
  }
  static std::mutex host_multi_mu;   // the per-device copy/exec streams and events carry one call at a time
  std::lock_guard<std::mutex> lk(host_multi_mu);
  const MapParams p = make_params(alpha, beta, dtype);
  size_t sb[kMaxDevices], sbytes[kMaxDevices], max_chunks = 0;
  for (int r = 0; r < n_ranks; ++r) {
    size_t b = 0, e = 0;
    ktb_shard_bounds(n_elems / granule, n_ranks, r, &b, &e);
    sb[r] = b * granule * es;
    sbytes[r] = (e - b) * granule * es;
    max_chunks = std::max(max_chunks, (sbytes[r] + chunk_bytes - 1) / chunk_bytes);
  }
  int status = KTB_OK;
  // chunk-major issue order: every GPU's PCIe link starts moving data before any link gets its second chunk
  for (size_t c = 0; c < max_chunks && status == KTB_OK; ++c) {
    for (int r = 0; r < n_ranks && status == KTB_OK; ++r) {
      const size_t off = c * chunk_bytes;
      if (off >= sbytes[r]) continue;
      const int dev = devs[r];
      KTB_GUARD(dev);
      DeviceInfo* di = device_info(dev);
      const int b = (int)(c & 1);
      cudaEvent_t h2d_done = di->host_ev[b], exec_done = di->host_ev[2 + b], d2h_done = di->host_ev[4 + b];
      const size_t len = std::min(chunk_bytes, sbytes[r] - off);
      uint8_t* sin = static_cast<uint8_t*>(stage_in[r]) + (size_t)b * chunk_bytes;
      uint8_t* sout = static_cast<uint8_t*>(stage_out[r]) + (size_t)b * chunk_bytes;
      cudaError_t e = cudaSuccess;
      if (c >= 2) e = cudaStreamWaitEvent(di->stream_h2d, exec_done, 0);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(sin, static_cast<const uint8_t*>(src_host) + sb[r] + off, len, cudaMemcpyHostToDevice,
                            di->stream_h2d);
      if (e == cudaSuccess) e = cudaEventRecord(h2d_done, di->stream_h2d);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(di->stream_exec, h2d_done, 0);
      if (e == cudaSuccess && c >= 2) e = cudaStreamWaitEvent(di->stream_exec, d2h_done, 0);
      if (e == cudaSuccess) {
        status = launch_map(dev, op, dtype, sin, sout, len / es, p, KTB_VARIANT_AUTO, di->stream_exec);
        if (status != KTB_OK) break;
        e = cudaEventRecord(exec_done, di->stream_exec);
      }
      if (e == cudaSuccess) e = cudaStreamWaitEvent(di->stream_d2h, exec_done, 0);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(static_cast<uint8_t*>(dst_host) + sb[r] + off, sout, len, cudaMemcpyDeviceToHost,
                            di->stream_d2h);
      if (e == cudaSuccess) e = cudaEventRecord(d2h_done, di->stream_d2h);
      if (e != cudaSuccess) {
        set_error("ktb_map_host_multi: rank %d chunk %zu failed: %s", r, c, cudaGetErrorString(e));
        status = KTB_ERR_CUDA;
      }
    }
  }
  for (int r = 0; r < n_ranks; ++r) {
    DeviceGuard g(devs[r]);
    DeviceInfo* di = device_info(devs[r]);
    cudaError_t e1 = cudaStreamSynchronize(di->stream_d2h);
    cudaError_t e2 = cudaStreamSynchronize(di->stream_exec);
    cudaError_t e3 = cudaStreamSynchronize(di->stream_h2d);
    if (status == KTB_OK && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)) {
      cudaError_t e = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
      set_error("ktb_map_host_multi: stream sync on device %d failed: %s", devs[r], cudaGetErrorString(e));
      status = KTB_ERR_CUDA;
    }
  }
  return status;
}

}  // extern "C"
