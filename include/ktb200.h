/*
 * ktb200.h — C-ABI of libktb200.so, the B200 (sm_100a) dispatch backend for
 * kubetorch's data-parallel remote-call path.
 *
 * The reference (run-house/kubetorch @ 96fac95, python_client/kubetorch, "kt/"
 * below) has no FFI: its hot path is CPython pickle/base64/JSON + HTTP +
 * multiprocessing.Queue.  Every entry point here therefore cites the reference
 * *Python* function whose work it replaces on this route:
 *
 *   pack / unpack          kt/serving/utils.py:730-749 (_serialize_body: pickle+b64 of args)
 *                          kt/serving/http_server.py:1768-1822 (_parse_callable_params: b64decode+unpickle)
 *                          kt/serving/http_server.py:1825-1842 (_serialize_result)
 *                          kt/serving/utils.py:787-813 (_deserialize_response)
 *   broadcast / scatter    kt/serving/spmd/spmd_supervisor.py:341,439-455 (params_list=[params]*P → call_all)
 *                          kt/serving/process_pool.py:125-212 (mp.Queue.put per rank)
 *                          kt/serving/remote_worker_pool.py:254-316 (HTTP POST per pod)
 *   map (exec)             kt/serving/http_server.py:1845-1891 (execute_callable_async → user fn)
 *   gather / gather-reduce kt/serving/spmd/spmd_supervisor.py:547-570 (local_responses + worker_responses)
 *                          kt/serving/process_pool.py:214-234 (_response_router)
 *   shard bounds           user-side `x.chunk(WORLD_SIZE)[RANK]` driven by the env contract of
 *                          kt/serving/process_worker.py:75-102
 *
 * Conventions
 *   - All functions return 0 on success, a negative ktb_status on failure; the
 *     message is available from ktb_last_error() (thread-local).
 *   - Buffers are caller-owned (PyTorch tensors: tensor.data_ptr()); the library
 *     borrows them for the duration of the enqueued work and never frees them.
 *     Memory from ktb_arena_alloc / ktb_host_alloc is library-owned until the
 *     matching free or ktb_shutdown.
 *   - `stream` is a cudaStream_t passed as uintptr_t (0 = legacy default stream).
 *     Calls enqueue and return without a host sync unless stated otherwise.
 *   - Device pointers may be local, peer-mapped (NVLink P2P / CUDA IPC) or
 *     mapped pinned host memory; the kernels only require what each entry states.
 *   - Thread-safe. No torch types cross this boundary.
 */
#ifndef KTB200_H
#define KTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KTB_VERSION 100 /* 0.1.0 */

typedef enum {
  KTB_OK = 0,
  KTB_ERR_CUDA = -1,     /* a CUDA runtime call failed; see ktb_last_error() */
  KTB_ERR_ARG = -2,      /* invalid argument (null pointer, bad enum, misaligned, too large) */
  KTB_ERR_STATE = -3,    /* library not initialised / device not registered */
  KTB_ERR_UNSUPPORTED = -4
} ktb_status;

/* Mapped-callable vocabulary: the closed set of user callables that run as kernels. */
typedef enum {
  KTB_OP_IDENTITY = 0,   /* y = x                         (BASELINE config C5)       */
  KTB_OP_SCALE = 1,      /* y = x * alpha                 (BASELINE config C2: x→2x) */
  KTB_OP_AFFINE = 2      /* y = (x * alpha) + beta, each step rounded to dtype; for BF16/F16 beta is
                          * itself rounded to the dtype first (torch CPU eager scalar semantics)  */
} ktb_op;

typedef enum {
  KTB_U8 = 0,            /* raw bytes; identity only */
  KTB_F32 = 1,
  KTB_BF16 = 2,
  KTB_I32 = 3,           /* wrapping two's-complement arithmetic (torch semantics) */
  KTB_I64 = 4,
  KTB_F16 = 5            /* op-math in fp32, rounded to half (RNE) after each step, like ATen */
} ktb_dtype;

/* Kernel variant selector for the element-wise map (all bit-identical in output). */
typedef enum {
  KTB_VARIANT_AUTO = 0,  /* widest vector path the pointers allow */
  KTB_VARIANT_VEC = 1,   /* register path: 256-/128-bit LDG/STG, unrolled, persistent grid */
  KTB_VARIANT_TMA = 2,   /* cp.async.bulk global→shared ring, compute in shared, bulk store */
  KTB_VARIANT_SCALAR = 3 /* element-at-a-time (any alignment) */
} ktb_variant;

/* ---- runtime -------------------------------------------------------------------------- */

/* Register devices dev_ids[0..n_dev) with the library and enable peer access between every
 * ordered pair that supports it.  Replaces the rendezvous of kt/serving/distributed_supervisor.py:90-174
 * (pod_ips quorum) by a static local membership.  Idempotent; devices accumulate across calls. */
int ktb_init(int n_dev, const int* dev_ids);
int ktb_shutdown(void);
const char* ktb_last_error(void);
int ktb_version(void);
/* Number of SMs of a registered device (148 on B200), or a negative status. */
int ktb_sm_count(int dev);
/* 1 if `dev` can read/write `peer` memory directly (after ktb_init), else 0; negative on error. */
int ktb_peer_enabled(int dev, int peer);

/* ---- arenas: library-owned memory (cudaMalloc / cudaHostAlloc), IPC-exportable ----------- */

int ktb_arena_alloc(int dev, size_t nbytes, void** out);
int ktb_arena_free(int dev, void* ptr);
/* Pinned, device-mapped host memory (portable across registered devices). */
int ktb_host_alloc(size_t nbytes, void** out);
int ktb_host_free(void* ptr);
/* NUMA node of a registered device (sysfs, via its PCI bus id), or -1 if unknown. */
int ktb_device_numa_node(int dev);
/* Pinned host block whose byte range (part_end[i-1], part_end[i]] is FIRST-TOUCHED on the NUMA node of
 * device part_dev[i] (by that device's issue thread) before the block is page-locked (cudaHostRegister,
 * portable + mapped), so shard i's DMA never crosses the socket interconnect.  n_parts == 0 → plain block.
 * Replaces the client-side argument/result buffers of kt/serving/http_client.py:1041-1111. */
int ktb_host_alloc_sharded(size_t nbytes, int n_parts, const size_t* part_end, const int* part_dev, void** out);
int ktb_host_free_sharded(void* ptr);
/* CUDA IPC for process-per-rank workers (the reference's ProcessWorker model,
 * kt/serving/process_worker.py:15-60).  `ptr` must be the base of a ktb_arena_alloc block. */
#define KTB_IPC_HANDLE_BYTES 64
int ktb_ipc_export(int dev, void* ptr, unsigned char handle[KTB_IPC_HANDLE_BYTES]);
int ktb_ipc_open(int dev, const unsigned char handle[KTB_IPC_HANDLE_BYTES], void** out);
int ktb_ipc_close(int dev, void* ptr);

/* ---- shard partition ------------------------------------------------------------------- */

/* Bounds [begin,end) in elements of rank `rank`'s shard of an n-element dim-0 split over
 * `world` ranks, following torch.chunk: chunk = ceil(n/world); ranks past the data get an
 * empty shard (begin == end == n).  This is what the reference's user functions compute
 * from RANK/WORLD_SIZE (`x.chunk(w)[r]`). */
int ktb_shard_bounds(size_t n, int world, int rank, size_t* begin, size_t* end);

/* ---- element-wise map: "execute the mapped callable" -------------------------------------- */

/* dst[i] = op(src[i]) for i in [0, n_elems).  src/dst may be local, peer or mapped-host
 * pointers valid on `dev`; they must be element-aligned and must not partially overlap
 * (src == dst is allowed).  alpha/beta are converted to the dtype's op-math type
 * (float for F32/BF16, int64 for I32/I64).  KTB_U8 supports KTB_OP_IDENTITY only.
 * KTB_VARIANT_TMA / _VEC need 16-byte aligned src and dst and silently use the next narrower
 * path otherwise (all variants are bit-identical). */
int ktb_map(int dev, int op, int dtype, const void* src, void* dst, size_t n_elems,
            double alpha, double beta, int variant, uintptr_t stream);

/* Named wrappers (SURVEY.md §8(b) B4 naming). */
int ktb_map_identity_u8(int dev, const void* src, void* dst, size_t nbytes, uintptr_t stream);
int ktb_map_scale_f32(int dev, const float* src, float* dst, size_t n, float alpha, uintptr_t stream);
int ktb_map_affine_f32(int dev, const float* src, float* dst, size_t n, float alpha, float beta, uintptr_t stream);
int ktb_map_scale_bf16(int dev, const void* src, void* dst, size_t n, float alpha, uintptr_t stream);
int ktb_map_affine_bf16(int dev, const void* src, void* dst, size_t n, float alpha, float beta, uintptr_t stream);

/* ---- gather-reduce variant ------------------------------------------------------------------ */

/* Bytes of zero-initialised device workspace ktb_map_reduce_sum needs (per concurrent call). */
size_t ktb_reduce_workspace_bytes(void);
/* out[0] = sum_i op(src[i]).  Accumulator/out type: float for F32 and BF16 (per-thread fp32,
 * warp-shuffle tree, fp64 across CTAs — deterministic for a given n and device), int64 for
 * I32/I64 (exact, wrapping).  `out` may be a peer pointer (rank r writes root_out[r]).
 * `workspace` must be zero before first use; the kernel restores it to zero. */
int ktb_map_reduce_sum(int dev, int op, int dtype, const void* src, size_t n_elems,
                       double alpha, double beta, void* out, void* workspace, uintptr_t stream);
/* out[0] = sum of n partials (float or int64, per dtype rule above); the root-side final step. */
int ktb_reduce_partials(int dev, int dtype, const void* partials, int n, void* out, uintptr_t stream);

/* ---- pack / unpack: many tensors <-> one arena ------------------------------------------------- */

#define KTB_PACK_ALIGN 256
/* Computes offsets[i] (KTB_PACK_ALIGN-aligned, in order) for n segments; returns total bytes
 * needed in *total.  Pure host arithmetic (the layout half of pack). */
int ktb_pack_layout(const size_t* nbytes, int n, size_t* offsets, size_t* total);
/* arena[offsets[i] .. +nbytes[i]) = srcs[i][0..nbytes[i]) for all i, as one or more segmented
 * copy launches.  If `offsets` was not pre-filled pass compute_layout=1 to fill it here. */
int ktb_pack(int dev, const void* const* srcs, const size_t* nbytes, int n, void* arena,
             size_t arena_bytes, size_t* offsets, int compute_layout, uintptr_t stream);
int ktb_unpack(int dev, const void* arena, const size_t* offsets, const size_t* nbytes, int n,
               void* const* dsts, uintptr_t stream);
/* Batched map: n independent calls dst_i = op(src_i) (n_elems[i] elements each) in as few
 * launches as possible — the coalesced form of many small remote calls. */
int ktb_map_batch(int dev, int op, int dtype, const void* const* srcs, void* const* dsts,
                  const size_t* n_elems, int n, double alpha, double beta, uintptr_t stream);

/* ---- multi-GPU data movement over NVLink / NVSwitch ---------------------------------------------- */

/* SPMD broadcast (reference semantics: every rank sees the full args): one kernel on `root`
 * reads src once and peer-stores it into every dsts[k] (k < n_dst; entries equal to src are
 * skipped).  All dsts must be mapped on `root`. */
int ktb_broadcast(int root, const void* src, void* const* dsts, int n_dst, size_t nbytes,
                  uintptr_t stream);

/* Fused scatter → map → gather for a registered op, single controller process:
 * rank r (device devs[r]) pulls its torch.chunk shard of src_root straight out of the root
 * GPU's memory, applies op, and pushes the result into dst_root at the same offset.  No
 * staging copies: root HBM is read once and written once.  streams[r] is the stream on
 * devs[r], used verbatim (0 = legacy default stream); streams == NULL → library-owned streams.  The call is ordered after prior work on
 * streams[root_rank] and that stream is ordered after all ranks' work on return.
 * `granule` = elements per indivisible unit (a dim-0 row): shards are ktb_shard_bounds over
 * n_elems/granule units, i.e. exactly `x.chunk(world)` along dim 0; n_elems % granule must be 0. */
int ktb_scatter_map_gather(int op, int dtype, const void* src_root, void* dst_root, size_t n_elems,
                           size_t granule, double alpha, double beta, int n_ranks, const int* devs,
                           int root_rank, int variant, const uintptr_t* streams);
/* Gather-reduce variant: rank r reduces op(shard r) and writes partials_root[r]; the root then
 * reduces the n_ranks partials into out_root[0].  workspaces[r] as in ktb_map_reduce_sum. */
int ktb_scatter_map_reduce(int op, int dtype, const void* src_root, size_t n_elems, size_t granule,
                           double alpha, double beta, int n_ranks, const int* devs, int root_rank,
                           void* partials_root, void* out_root, void* const* workspaces,
                           const uintptr_t* streams);

/* ---- push/push pipeline: in-kernel flag synchronisation, no host sync on the data path ----------- */

/* Bytes of a control block (zero-initialised device memory; one per rank, one for the root). */
size_t ktb_push_control_bytes(void);
/* ROOT side of call number `seq` (1,2,3,... — every participant counts calls identically): for each
 * of n_chunks pieces, peer-store the piece of every non-root rank's shard into
 * stage_peer[r] + (seq&1)*stage_stride and publish ready[chunk] = seq in ctrl_peer[r].  The kernel
 * itself waits for ack[r] >= seq-2 (in ctrl_root) before overwriting a staging half.
 * stage_peer[r] / ctrl_peer[r] are pointers valid on root_dev (peer access or CUDA IPC). */
int ktb_push_scatter(int root_dev, const void* src_root, size_t n_elems, size_t granule, int dtype,
                     int n_ranks, int root_rank, void* const* stage_peer, size_t stage_stride,
                     void* const* ctrl_peer, void* ctrl_root, int n_chunks, unsigned long long seq,
                     uintptr_t stream);
/* The same root side with pieces of EXACTLY chunk_elems elements per rank (n_chunks = ceil(largest shard /
 * chunk_elems) <= 64): for consumers that need whole work units per piece (ktb_mlp_bf16_pushed: GEMM row chunks).
 * ctas_per_sm caps the persistent grid (0 = library default) so the root's own compute keeps its share of every SM.
 * A rank whose stage_peer[r] is NULL is skipped (also in ktb_push_scatter / ktb_push_scatter_ce): two calls with
 * complementary NULL masks split the ranks between engines (hybrid copy-engine + SM scatter). */
int ktb_push_scatter_chunked(int root_dev, const void* src_root, size_t n_elems, size_t granule, int dtype,
                             int n_ranks, int root_rank, void* const* stage_peer, size_t stage_stride,
                             void* const* ctrl_peer, void* ctrl_root, size_t chunk_elems, int ctas_per_sm,
                             unsigned long long seq, uintptr_t stream);
/* Copy-engine form of the chunked root side: every piece is a cudaMemcpyPeerAsync on a per-destination library
 * stream of the root followed by a one-thread flag publish, issued chunk-major; NO SM of the root moves data, so its
 * own compute (rank 0's GEMMs in ktb_mlp_bf16_pushed deployments) keeps the whole GPU.  `stream` is ordered before
 * (sources ready) and after (sources reusable) the copies.  devs[r] = device of rank r. */
int ktb_push_scatter_ce(int root_dev, const void* src_root, size_t n_elems, size_t granule, int dtype, int n_ranks,
                        int root_rank, const int* devs, void* const* stage_peer, size_t stage_stride,
                        void* const* ctrl_peer, void* ctrl_root, size_t chunk_elems, unsigned long long seq,
                        uintptr_t stream);
/* RANK side: for each piece, spin in-kernel until ready[chunk] >= seq, then
 * dst_root_shard[piece] = op(stage_local[piece]) (peer stores into the root's result arena); after the
 * last piece publish ack[rank] = seq in the root's control block (ctrl_root_peer). */
int ktb_push_consume(int dev, int op, int dtype, const void* stage_local, size_t stage_stride,
                     void* dst_root_shard, size_t shard_elems, double alpha, double beta,
                     void* ctrl_local, void* ctrl_root_peer, int rank, int n_chunks,
                     unsigned long long seq, uintptr_t stream);
/* ROOT: stream-ordered completion of call `seq` (spins until every ack[r] >= seq). */
int ktb_push_wait(int root_dev, void* ctrl_root, int n_ranks, int root_rank, unsigned long long seq,
                  uintptr_t stream);
/* Synchronously read a control block's sticky status: 0 healthy, 1 = an in-kernel wait timed out. */
int ktb_push_status(int dev, const void* ctrl, unsigned int* out);

/* ---- host-resident args/results (the reference's client lives outside the GPU) -------------------- */

/* dst_host = op(src_host) with both buffers in pinned host memory, chunked through
 * device staging buffers (each >= 2*chunk_bytes) on three library streams so that H2D copy,
 * kernel and D2H copy of successive chunks overlap (PCIe is full duplex).  Synchronous:
 * returns when dst_host is complete. */
int ktb_map_host(int dev, int op, int dtype, const void* src_host, void* dst_host, size_t n_elems,
                 double alpha, double beta, size_t chunk_bytes, void* stage_in, void* stage_out);

/* The same for a sharded call on n_ranks DISTINCT devices: rank r's `x.chunk(n_ranks)[r]` (granule as in
 * ktb_scatter_map_gather) moves host → devs[r] → host over that GPU's own PCIe link, each pipeline enqueued by a
 * persistent library issue thread bound to the CPUs of that GPU's NUMA node (all links run concurrently).
 * stage_in[r] / stage_out[r] are device buffers of >= 2*chunk_bytes on devs[r].  Synchronous. */
int ktb_map_host_multi(int op, int dtype, const void* src_host, void* dst_host, size_t n_elems,
                       size_t granule, double alpha, double beta, int n_ranks, const int* devs,
                       size_t chunk_bytes, void* const* stage_in, void* const* stage_out);

/* ---- bf16 MLP policy (BASELINE config C4) ---------------------------------------------------------- */

/* logits[M,d_out] = W3·relu(W2·relu(W1·obs^T)) with bf16 storage, fp32 accumulation on tcgen05
 * tensor cores (TMEM accumulators, TMA-fed CTA pairs), activations rounded to bf16 between layers.
 * W_l is [d_l, d_{l-1}] row-major (nn.Linear layout).  This build: M % 128 == 0, d_in % 64 == 0,
 * d_hidden % 256 == 0, d_out == 64 (KTB_ERR_UNSUPPORTED otherwise).  Rows are processed in chunks;
 * for chunks with rows % 256 == 0 layer 2 and the head run as ONE kernel (the second hidden activation
 * never leaves the SM), other chunks use three GEMM launches — results agree within one bf16 ulp.
 * `scratch` holds 2*M_chunk*d_hidden bf16 (see ktb_mlp_scratch_bytes).  Replaces the user's
 * nn.Sequential policy inside execute_callable_async (kt/serving/http_server.py:1845-1891). */
size_t ktb_mlp_scratch_bytes(size_t M, int d_hidden);
int ktb_mlp_bf16(int dev, const void* obs, size_t M, int d_in, int d_hidden, int d_out,
                 const void* W1, const void* W2, const void* W3, void* logits, void* scratch,
                 uintptr_t stream);
/* Scatter-fused form for a rank whose observations live on the ROOT GPU: row chunks are pulled
 * peer → local into `stage` (ktb_mlp_stage_bytes, double-buffered) on a library side stream while
 * the previous chunk computes; `logits` may be a peer pointer (fused gather). */
size_t ktb_mlp_stage_bytes(size_t M, int d_in);
int ktb_mlp_bf16_staged(int dev, const void* obs_peer, size_t M, int d_in, int d_hidden, int d_out,
                        const void* W1, const void* W2, const void* W3, void* logits, void* scratch,
                        void* stage, uintptr_t stream);

/* Push-fed form for a rank whose observation rows are PUSHED by the root (ktb_push_scatter_chunked with
 * chunk_elems = chunk_rows * d_in into stage_local, double-buffered by call parity with stride stage_stride):
 * before each row chunk's GEMMs a one-warp kernel waits in-stream for ready[chunk] >= seq; the logits are stored
 * straight into `logits` (a peer pointer into the root's result: fused gather) and ack[rank] = seq is published in
 * the root's control block behind the last chunk.  Root NVLink egress carries posted writes only (689 GB/s measured
 * with the result stream flowing the other way, against 403 GB/s when seven ranks pull). */
int ktb_mlp_bf16_pushed(int dev, const void* stage_local, size_t stage_stride, size_t M, int d_in, int d_hidden,
                        int d_out, const void* W1, const void* W2, const void* W3, void* logits, void* scratch,
                        void* ctrl_local, void* ctrl_root_peer, int rank, size_t chunk_rows, unsigned long long seq,
                        uintptr_t stream);

#ifdef __cplusplus
}
#endif
#endif /* KTB200_H */
