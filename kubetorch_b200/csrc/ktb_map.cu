// ktb_map.cu — the element-wise mapped callable ("execute" step of the remote-map path).
//
// Replaces the per-rank user-function execution of kt/serving/http_server.py:1845-1891
// (execute_callable_async) for the closed set of registered ops (identity / scale / affine).
// The same kernels serve as the fused scatter→exec→gather when src/dst are peer-mapped
// pointers into the root GPU's arg/result arenas (ktb_dispatch.cu).
//
// HBM-bound: algorithmic bytes = n*(sizeof in + sizeof out); no data reuse, so the only
// levers are coalesced wide accesses, enough bytes in flight, and a grid that is a multiple
// of the SM count.  Three bit-identical variants:
//   VEC    — 256-bit (or 128-bit) LDG/STG, UNROLL loads in flight per thread, persistent grid
//   TMA    — cp.async.bulk global→shared ring (mbarrier complete_tx), compute in shared,
//            cp.async.bulk shared→global; one elected thread drives the TMA engine
//   SCALAR — any alignment
#include "ktb_common.cuh"

#include <algorithm>
#include <atomic>

namespace ktb {

extern int g_red_ctas_per_sm;  // ktb_reduce.cu
extern int g_red_loads;        // ktb_reduce.cu
extern int g_red_fold;         // ktb_reduce.cu
extern int g_mlp_cluster4;     // ktb_mlp.cu
extern int g_mlp_stages;       // ktb_mlp.cu
extern int g_mlp_fuse_head;    // ktb_mlp.cu
extern std::atomic<int> g_mlp_cluster4_max[kMaxDevices];
extern int g_seg_large;        // ktb_pack.cu
extern int g_mlp_persistent;   // ktb_mlp.cu
extern int g_mlp_chunk_rows;   // ktb_mlp.cu
extern int g_mlp_epi_groups;   // ktb_mlp.cu
extern int g_mlp_tma_store;    // ktb_mlp.cu
extern int g_mlp_2sm;          // ktb_mlp.cu
extern int g_mlp_l1_bres;      // ktb_mlp.cu
extern int g_mlp_arrive_mode;  // ktb_mlp.cu
}  // namespace ktb
extern int g_mlp_stage_ce;     // ktb_mlp.cu (file scope there)
namespace ktb {
extern std::atomic<int> g_host_zero_copy;   // ktb_host.cu
}  // namespace ktb
extern std::atomic<int> g_push_scatter_ctas_per_sm;   // ktb_push.cu (file scope there)
extern std::atomic<int> g_push_slice;
namespace ktb {

// ---- tunables (ktb_set_tuning) ---------------------------------------------------------------
static std::atomic<int> g_vec_ctas_per_sm{0};  // 0 = one tile per CTA (measured best, profiles/r1_sweep.md)
static std::atomic<int> g_tma_ctas_per_sm{1};
static std::atomic<int> g_tma_cfg{0};       // 0: 4 x 32 KiB stages, 1: 6 x 32 KiB, 2: 8 x 16 KiB (2 CTA/SM)
static std::atomic<int> g_auto_variant{KTB_VARIANT_VEC};
static std::atomic<int> g_vec_flavor{2};    // cache flavor (tunable for F32-scale and identity only)
static std::atomic<int> g_vec_unroll{2};    // 2 | 4 | 8 (tunable for the same kernels)

constexpr int kVecThreads = 256;
constexpr int kVecUnroll = 2;   // defaults measured on B200: 2 x 256-bit loads in flight per thread,
constexpr int kVecFlavor = 2;   // L2::evict_first streaming hints, one 16 KiB tile per CTA

// ---- VEC -----------------------------------------------------------------------------------------
// FL (cache flavor, 256-bit path only): 0 = L1::no_allocate, 1 = default caching,
// 2 = L1::no_allocate + L2::evict_first on loads and stores, 3 = evict_first loads only.
template <int VB, int FL>
__device__ __forceinline__ void ld_vec(const uint8_t* p, uint32_t (&w)[VB / 4]) {
  if constexpr (VB == 32) {
#define KTB_LD256(MOD)                                                                          \
  asm volatile("ld.global" MOD ".v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                       \
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]),        \
                 "=r"(w[6]), "=r"(w[7])                                                         \
               : "l"(p)                                                                         \
               : "memory")
    if constexpr (FL == 1) {
      KTB_LD256("");
    } else if constexpr (FL == 2 || FL == 3) {
      KTB_LD256(".L1::no_allocate.L2::evict_first");
    } else {
      KTB_LD256(".L1::no_allocate");
    }
#undef KTB_LD256
  } else {
    asm volatile("ld.global.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                 : "l"(p)
                 : "memory");
  }
}
template <int VB, int FL>
__device__ __forceinline__ void st_vec(uint8_t* p, const uint32_t (&w)[VB / 4]) {
  if constexpr (VB == 32) {
#define KTB_ST256(MOD)                                                                          \
  asm volatile("st.global" MOD ".v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]),  \
               "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])      \
               : "memory")
    if constexpr (FL == 1) {
      KTB_ST256("");
    } else if constexpr (FL == 2) {
      KTB_ST256(".L1::no_allocate.L2::evict_first");
    } else {
      KTB_ST256(".L1::no_allocate");
    }
#undef KTB_ST256
  } else {
    stg128(p, w);
  }
}

template <int DT, int OP, int VB, int FL = kVecFlavor, int UNROLL = kVecUnroll>
__global__ void __launch_bounds__(kVecThreads)
    map_vec_kernel(const uint8_t* src, uint8_t* dst, size_t n_bytes, MapParams p) {
  constexpr int NW = VB / 4;
  constexpr int kVecUnroll = UNROLL;
  constexpr size_t TILE = (size_t)kVecThreads * kVecUnroll * VB;  // bytes per CTA iteration
  constexpr size_t ROW = (size_t)kVecThreads * VB;                // bytes per unrolled step
  const size_t n_full = n_bytes / TILE;

  for (size_t t = blockIdx.x; t < n_full; t += gridDim.x) {
    const size_t off = t * TILE + (size_t)threadIdx.x * VB;
    uint32_t w[kVecUnroll][NW];
#pragma unroll
    for (int j = 0; j < kVecUnroll; ++j) ld_vec<VB, FL>(src + off + j * ROW, w[j]);
#pragma unroll
    for (int j = 0; j < kVecUnroll; ++j) {
      apply_words<DT, OP, NW>(w[j], p);
      st_vec<VB, FL>(dst + off + j * ROW, w[j]);
    }
  }

  // Remainder (< TILE bytes): one CTA, bounds-checked vectors then an element tail.
  if (blockIdx.x == (unsigned)(n_full % gridDim.x)) {
    const size_t base = n_full * TILE;
    const size_t n_vec = (n_bytes - base) / VB;
    for (size_t v = threadIdx.x; v < n_vec; v += kVecThreads) {
      uint32_t w[NW];
      ld_vec<VB, FL>(src + base + v * VB, w);
      apply_words<DT, OP, NW>(w, p);
      st_vec<VB, FL>(dst + base + v * VB, w);
    }
    constexpr size_t ES = (DT == KTB_U8) ? 1 : ((DT == KTB_BF16 || DT == KTB_F16) ? 2 : (DT == KTB_I64 ? 8 : 4));
    const size_t tail = base + n_vec * VB;
    const size_t n_tail = (n_bytes - tail) / ES;
    for (size_t e = threadIdx.x; e < n_tail; e += kVecThreads)
      apply_elem<DT, OP>(src + tail + e * ES, dst + tail + e * ES, p);
  }
}

// ---- SCALAR ------------------------------------------------------------------------------------
template <int DT, int OP>
__global__ void __launch_bounds__(256)
    map_scalar_kernel(const uint8_t* src, uint8_t* dst, size_t n_elems, MapParams p) {
  constexpr size_t ES = (DT == KTB_U8) ? 1 : ((DT == KTB_BF16 || DT == KTB_F16) ? 2 : (DT == KTB_I64 ? 8 : 4));
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride)
    apply_elem<DT, OP>(src + i * ES, dst + i * ES, p);
}

// ---- TMA -----------------------------------------------------------------------------------------
// One CTA owns a ring of STAGES shared-memory tiles.  Thread 0 is the TMA driver: it arms
// full[s] with the tile's byte count and issues the bulk load; all threads wait on full[s],
// transform the tile in place (16-byte shared accesses, conflict-free), fence the generic→async
// proxy, and thread 0 issues the bulk store.  A stage is re-armed once the bulk store that
// read it has finished reading (cp.async.bulk.wait_group.read), one iteration later, so
// STAGES-1 loads stay in flight per CTA.
template <int DT, int OP, int STAGES, int STAGE_BYTES, int NT>
__global__ void __launch_bounds__(NT)
    map_tma_kernel(const uint8_t* src, uint8_t* dst, size_t n_bytes, MapParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* buf = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
  constexpr bool kCompute = !(OP == KTB_OP_IDENTITY || DT == KTB_U8);

  const size_t n16 = n_bytes & ~(size_t)15;  // bulk copies move multiples of 16 bytes
  const size_t n_tiles = (n16 + STAGE_BYTES - 1) / STAGE_BYTES;
  const size_t my_n =
      (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  auto tile_off = [&](size_t k) { return ((size_t)blockIdx.x + k * gridDim.x) * STAGE_BYTES; };
  auto tile_len = [&](size_t k) {
    size_t r = n16 - tile_off(k);
    return (uint32_t)(r < (size_t)STAGE_BYTES ? r : (size_t)STAGE_BYTES);
  };

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  __syncthreads();

  if (threadIdx.x == 0) {
    const size_t pre = my_n < (size_t)STAGES ? my_n : (size_t)STAGES;
    for (size_t k = 0; k < pre; ++k) {
      const uint32_t len = tile_len(k);
      mbar_expect_tx(&full[k], len);
      bulk_g2s(buf + k * STAGE_BYTES, src + tile_off(k), len, &full[k]);
    }
  }

  if (kCompute || threadIdx.x == 0) {
    for (size_t k = 0; k < my_n; ++k) {
      const int s = (int)(k % STAGES);
      const uint32_t parity = (uint32_t)((k / STAGES) & 1);
      const uint32_t len = tile_len(k);
      mbar_wait(&full[s], parity);
      if constexpr (kCompute) {
        uint4* tile = reinterpret_cast<uint4*>(buf + (size_t)s * STAGE_BYTES);
        const uint32_t nv = len >> 4;
        for (uint32_t v = threadIdx.x; v < nv; v += NT) {
          uint4 q = tile[v];
          uint32_t w[4] = {q.x, q.y, q.z, q.w};
          apply_words<DT, OP, 4>(w, p);
          tile[v] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_proxy_async_smem();  // make generic-proxy writes visible to the bulk store
        __syncthreads();
      }
      if (threadIdx.x == 0) {
        bulk_s2g(dst + tile_off(k), buf + (size_t)s * STAGE_BYTES, len);
        bulk_commit();
        if (k >= 1 && (k - 1 + STAGES) < my_n) {
          bulk_wait_read<1>();  // store of tile k-1 has finished reading its stage
          const size_t kn = k - 1 + STAGES;
          const int sn = (int)((k - 1) % STAGES);
          const uint32_t ln = tile_len(kn);
          mbar_expect_tx(&full[sn], ln);
          bulk_g2s(buf + (size_t)sn * STAGE_BYTES, src + tile_off(kn), ln, &full[sn]);
        }
      }
    }
    if (threadIdx.x == 0) bulk_wait_all<0>();
  }

  // < 16 trailing bytes: element tail by CTA 0.
  if (blockIdx.x == 0) {
    constexpr size_t ES = (DT == KTB_U8) ? 1 : ((DT == KTB_BF16 || DT == KTB_F16) ? 2 : (DT == KTB_I64 ? 8 : 4));
    const size_t n_tail = (n_bytes - n16) / ES;
    for (size_t e = threadIdx.x; e < n_tail; e += NT)
      apply_elem<DT, OP>(src + n16 + e * ES, dst + n16 + e * ES, p);
  }
}

// ---- launch plumbing --------------------------------------------------------------------------
template <int DT, int OP>
static int launch_typed(int dev, const uint8_t* src, uint8_t* dst, size_t n_elems, size_t es,
                        const MapParams& p, int variant, cudaStream_t stream) {
  const DeviceInfo* di = device_info(dev);
  const size_t n_bytes = n_elems * es;
  const uintptr_t both = (uintptr_t)src | (uintptr_t)dst;

  if (variant == KTB_VARIANT_AUTO) variant = g_auto_variant.load();
  // bulk (TMA) and vector paths need 16-byte aligned pointers; ragged shard boundaries fall back
  if (variant == KTB_VARIANT_TMA && (both & 15)) variant = KTB_VARIANT_VEC;
  if (variant == KTB_VARIANT_VEC && (both & 15)) variant = KTB_VARIANT_SCALAR;

  if (variant == KTB_VARIANT_VEC) {
    // ctas_per_sm == 0 → one tile per CTA (the hardware scheduler balances the tail);
    // otherwise a persistent grid of sm_count * ctas_per_sm CTAs striding over tiles.
    const int per_sm = g_vec_ctas_per_sm.load();
    auto grid_for = [&](size_t tile_bytes) {
      size_t tiles = std::max<size_t>(n_bytes / tile_bytes, 1);
      if (per_sm <= 0) return (int)std::min<size_t>(tiles, 0x7fffffffULL);
      return (int)std::min<size_t>(tiles, (size_t)di->sm_count * per_sm);
    };
    if ((both & 31) == 0) {
      constexpr bool kTunable = (DT == KTB_U8) || (DT == KTB_F32 && OP == KTB_OP_SCALE);
      const int fl = kTunable ? g_vec_flavor.load() : kVecFlavor;
      const int un = kTunable ? g_vec_unroll.load() : kVecUnroll;
#define KTB_VEC32(FL, UN)                                                                          \
  map_vec_kernel<DT, OP, 32, FL, UN>                                                               \
      <<<grid_for((size_t)kVecThreads * (UN)*32), kVecThreads, 0, stream>>>(src, dst, n_bytes, p)
      if constexpr (kTunable) {
        if (un == 2) {
          if (fl == 1) KTB_VEC32(1, 2); else if (fl == 2) KTB_VEC32(2, 2); else if (fl == 3) KTB_VEC32(3, 2); else KTB_VEC32(0, 2);
        } else if (un == 8) {
          if (fl == 1) KTB_VEC32(1, 8); else if (fl == 2) KTB_VEC32(2, 8); else if (fl == 3) KTB_VEC32(3, 8); else KTB_VEC32(0, 8);
        } else {
          if (fl == 1) KTB_VEC32(1, 4); else if (fl == 2) KTB_VEC32(2, 4); else if (fl == 3) KTB_VEC32(3, 4); else KTB_VEC32(0, 4);
        }
      } else {
        (void)fl; (void)un;
        KTB_VEC32(kVecFlavor, kVecUnroll);
      }
#undef KTB_VEC32
    } else {
      map_vec_kernel<DT, OP, 16>
          <<<grid_for((size_t)kVecThreads * kVecUnroll * 16), kVecThreads, 0, stream>>>(src, dst, n_bytes, p);
    }
  } else if (variant == KTB_VARIANT_TMA) {
    const int cfg = g_tma_cfg.load();
    const int per_sm = std::max(1, g_tma_ctas_per_sm.load());
#define KTB_LAUNCH_TMA(ST, SB, NT)                                                              \
  do {                                                                                          \
    constexpr size_t smem_bytes = (size_t)(ST) * (SB) + (ST) * sizeof(uint64_t);                \
    auto kfn = map_tma_kernel<DT, OP, ST, SB, NT>;                                              \
    KTB_CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,               \
                                (int)smem_bytes));                                              \
    size_t tiles = ((n_bytes & ~(size_t)15) + (SB)-1) / (SB);                                   \
    int grid = (int)std::min<size_t>(std::max<size_t>(tiles, 1), (size_t)di->sm_count * per_sm); \
    kfn<<<grid, NT, smem_bytes, stream>>>(src, dst, n_bytes, p);                                \
  } while (0)
    if (cfg == 1) {
      KTB_LAUNCH_TMA(6, 32768, 256);
    } else if (cfg == 2) {
      KTB_LAUNCH_TMA(6, 16384, 256);
    } else if (cfg == 3) {
      KTB_LAUNCH_TMA(3, 65536, 512);
    } else {
      KTB_LAUNCH_TMA(4, 32768, 256);
    }
#undef KTB_LAUNCH_TMA
  } else if (variant == KTB_VARIANT_SCALAR) {
    size_t blocks = (n_elems + 255) / 256;
    int grid = (int)std::min<size_t>(std::max<size_t>(blocks, 1), (size_t)di->sm_count * 8);
    map_scalar_kernel<DT, OP><<<grid, 256, 0, stream>>>(src, dst, n_elems, p);
  } else {
    set_error("ktb_map: unknown variant %d", variant);
    return KTB_ERR_ARG;
  }
  KTB_CK(cudaGetLastError());
  return KTB_OK;
}

int launch_map(int dev, int op, int dtype, const void* src_, void* dst_, size_t n_elems,
               const MapParams& p, int variant, cudaStream_t stream) {
  const uint8_t* src = static_cast<const uint8_t*>(src_);
  uint8_t* dst = static_cast<uint8_t*>(dst_);
  const size_t es = dtype_size(dtype);
  KTB_REQUIRE(es != 0, KTB_ERR_ARG, "ktb_map: unknown dtype %d", dtype);
  KTB_REQUIRE(op >= KTB_OP_IDENTITY && op <= KTB_OP_AFFINE, KTB_ERR_ARG, "ktb_map: unknown op %d", op);
  KTB_REQUIRE(!(dtype == KTB_U8 && op != KTB_OP_IDENTITY), KTB_ERR_ARG,
              "ktb_map: KTB_U8 supports KTB_OP_IDENTITY only");
  if (n_elems == 0) return KTB_OK;
  KTB_REQUIRE(src && dst, KTB_ERR_ARG, "ktb_map: null src/dst with n_elems=%zu", n_elems);
  KTB_REQUIRE((((uintptr_t)src | (uintptr_t)dst) & (es - 1)) == 0, KTB_ERR_ARG,
              "ktb_map: src/dst not aligned to the element size %zu", es);
  KTB_REQUIRE(n_elems <= (SIZE_MAX / 8), KTB_ERR_ARG, "ktb_map: n_elems too large");
  if (src != dst) {
    const uint8_t* se = src + n_elems * es;
    const uint8_t* de = dst + n_elems * es;
    KTB_REQUIRE(se <= dst || de <= src, KTB_ERR_ARG, "ktb_map: src and dst partially overlap");
  }

  // identity is a byte copy whatever the dtype
  if (op == KTB_OP_IDENTITY)
    return launch_typed<KTB_U8, KTB_OP_IDENTITY>(dev, src, dst, n_elems * es, 1, p, variant, stream);

#define KTB_CASE(DT)                                                                              \
  case DT:                                                                                        \
    return (op == KTB_OP_SCALE)                                                                   \
               ? launch_typed<DT, KTB_OP_SCALE>(dev, src, dst, n_elems, es, p, variant, stream)   \
               : launch_typed<DT, KTB_OP_AFFINE>(dev, src, dst, n_elems, es, p, variant, stream);
  switch (dtype) {
    KTB_CASE(KTB_F32)
    KTB_CASE(KTB_BF16)
    KTB_CASE(KTB_I32)
    KTB_CASE(KTB_I64)
    KTB_CASE(KTB_F16)
  }
#undef KTB_CASE
  set_error("ktb_map: unsupported dtype/op %d/%d", dtype, op);
  return KTB_ERR_UNSUPPORTED;
}

}  // namespace ktb

using namespace ktb;

extern "C" {

// Experiment knobs (not part of the stable ABI; used by the bench sweep).
//   key 0: VEC CTAs per SM (0 = one tile per CTA)   key 1: TMA CTAs per SM   key 2: TMA stage config
//   key 3: AUTO variant   key 4: VEC cache flavor   key 5: VEC unroll
int ktb_set_tuning(int key, int value) {
  switch (key) {
    case 0: g_vec_ctas_per_sm = value; return KTB_OK;
    case 1: g_tma_ctas_per_sm = value; return KTB_OK;
    case 2: g_tma_cfg = value; return KTB_OK;
    case 4: g_vec_flavor = value; return KTB_OK;
    case 5: g_vec_unroll = value; return KTB_OK;
    case 6: g_red_ctas_per_sm = value > 0 ? value : 4; return KTB_OK;
    case 7: g_mlp_persistent = value ? 1 : 0; return KTB_OK;
    case 9: g_mlp_epi_groups = (value == 2) ? 2 : 1; return KTB_OK;
    case 10: g_mlp_tma_store = value ? 1 : 0; return KTB_OK;
    case 11: g_mlp_2sm = value ? 1 : 0; return KTB_OK;
    case 12: g_seg_large = value ? 1 : 0; return KTB_OK;
    case 13: g_red_loads = value == 4 ? 4 : 8; return KTB_OK;
    case 14: g_red_fold = value ? 1 : 0; return KTB_OK;
    case 18: g_mlp_fuse_head = value ? 1 : 0; return KTB_OK;
    case 20: g_host_zero_copy = value ? 1 : 0; return KTB_OK;
    case 25: g_mlp_arrive_mode = (value >= 0 && value <= 63 && (value & 3) != 3 && !(value & 8)) ? value : 1; return KTB_OK;   // bit 2: pipelined TMEM loads in the fused epilogue
    case 24: g_mlp_l1_bres = (value >= 0 && value <= 4) ? value : 2; return KTB_OK;   // 2 = per-warp stores, 3 = + 8 epilogue warps
    case 23: g_push_slice = value > 0 ? value : 1; return KTB_OK;
    case 22: g_mlp_stage_ce = value ? 1 : 0; return KTB_OK;
    case 21: g_push_scatter_ctas_per_sm = value > 0 ? value : 0; return KTB_OK;
    case 17: g_mlp_stages = value == 5 ? 5 : 4; return KTB_OK;
    case 15: g_mlp_cluster4 = value ? 1 : 0; return KTB_OK;
    case 16: return (value >= 0 && value < kMaxDevices) ? g_mlp_cluster4_max[value].load() : 0;   // query, after a cluster-4 launch
    case 8:
      KTB_REQUIRE(value > 0 && value % 128 == 0, KTB_ERR_ARG, "ktb_set_tuning: MLP chunk rows must be a multiple of 128");
      g_mlp_chunk_rows = value;
      return KTB_OK;
    case 3:
      KTB_REQUIRE(value >= KTB_VARIANT_VEC && value <= KTB_VARIANT_SCALAR, KTB_ERR_ARG,
                  "ktb_set_tuning: bad auto variant %d", value);
      g_auto_variant = value;
      return KTB_OK;
    default: set_error("ktb_set_tuning: unknown key %d", key); return KTB_ERR_ARG;
  }
}

int ktb_map(int dev, int op, int dtype, const void* src, void* dst, size_t n_elems, double alpha,
            double beta, int variant, uintptr_t stream) {
  int rc = require_device(dev);
  if (rc) return rc;
  KTB_GUARD(dev);
  return launch_map(dev, op, dtype, src, dst, n_elems, make_params(alpha, beta, dtype), variant,
                    reinterpret_cast<cudaStream_t>(stream));
}

int ktb_map_identity_u8(int dev, const void* src, void* dst, size_t nbytes, uintptr_t stream) {
  return ktb_map(dev, KTB_OP_IDENTITY, KTB_U8, src, dst, nbytes, 1.0, 0.0, KTB_VARIANT_AUTO, stream);
}
int ktb_map_scale_f32(int dev, const float* src, float* dst, size_t n, float alpha, uintptr_t stream) {
  return ktb_map(dev, KTB_OP_SCALE, KTB_F32, src, dst, n, alpha, 0.0, KTB_VARIANT_AUTO, stream);
}
int ktb_map_affine_f32(int dev, const float* src, float* dst, size_t n, float alpha, float beta,
                       uintptr_t stream) {
  return ktb_map(dev, KTB_OP_AFFINE, KTB_F32, src, dst, n, alpha, beta, KTB_VARIANT_AUTO, stream);
}
int ktb_map_scale_bf16(int dev, const void* src, void* dst, size_t n, float alpha, uintptr_t stream) {
  return ktb_map(dev, KTB_OP_SCALE, KTB_BF16, src, dst, n, alpha, 0.0, KTB_VARIANT_AUTO, stream);
}
int ktb_map_affine_bf16(int dev, const void* src, void* dst, size_t n, float alpha, float beta,
                        uintptr_t stream) {
  return ktb_map(dev, KTB_OP_AFFINE, KTB_BF16, src, dst, n, alpha, beta, KTB_VARIANT_AUTO, stream);
}

}  // extern "C"
