// ktb_mlp.cu — the bf16 MLP policy callable of BASELINE config C4 on 5th-gen tensor cores.
//
// The mapped callable (oracle/cases.py:mlp_policy; what the reference would run per rank inside
// kt/serving/http_server.py:1845-1891) is   logits = W3·relu(W2·relu(W1·obsᵀ))   in bf16.
// It is the one place on this path where the user function is itself a dense GEMM, so it is the
// one place tensor cores are used: each layer is   C[M,N] = act(A[M,K] · B[N,K]ᵀ)   with
//   * A and B tiles staged global→shared by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B, K-major),
//   * tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BLOCK_N, K=16) issued by ONE thread,
//     fp32 accumulator in TMEM (BLOCK_N columns),
//   * epilogue warps reading TMEM with tcgen05.ld (32 lanes x 32 columns), fused ReLU + bf16
//     rounding, 16-byte global stores — the last layer's C may be a peer pointer into the root
//     GPU's result arena, which fuses the gather into the epilogue.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2-5 = epilogue (warp w reads TMEM lanes 32*(w%4)..+31).
// Activations are rounded to bf16 between layers (like the eager torch module); rows are processed
// in chunks whose two hidden activations stay resident in the 126 MB L2.
//
// Roofline: tensor-bound on one GPU (5.77e12 flop per C4 call vs 1.34 GB of arg+result);
// transfer-bound through the root's NVLink port at 8 GPUs.
#include "ktb_common.cuh"

#include <cuda.h>
#include <algorithm>
#include <atomic>
#include <mutex>

namespace ktb {

constexpr int kMlpBlockM = 128;
constexpr int kMlpBlockK = 64;   // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int kMlpUmmaK = 16;
constexpr int kMlpThreads = 192;
int g_mlp_chunk_rows = 75776;  // ktb_set_tuning key 8: rows per chunk = 4 x 74 CTA pairs x 256 rows: the fused layer-2+head kernel
                               // schedules whole 256-row blocks per pair, so a chunk is a whole number of waves on 148 SMs
                               // (measured 1150 TFLOP/s; 65536 rows: 1059 fused, 940-964 unfused)
int g_mlp_epi_groups = 1;      // ktb_set_tuning key 9: epilogue warpgroups (1 or 2); 2 measured 3% slower
std::atomic<int> g_mlp_cluster4_max[kMaxDevices];   // co-resident clusters of 4 per device (occupancy query, cached)
int g_mlp_fuse_head = 1;       // ktb_set_tuning key 18: 1 = layer 2 and the 64-wide head in one kernel (h2 stays on chip; default)
int g_mlp_stages = 4;          // ktb_set_tuning key 17: TMA ring depth of the CTA-pair kernel (4 or 5)
int g_mlp_cluster4 = 0;        // ktb_set_tuning key 15: 1 = cluster-of-4 multicast form of the CTA-pair kernel (opt-in)
unsigned long long* g_mlp_trace = nullptr;   // ktb_debug_set_ptr(0, p): device buffer of 16 x 64 clock64() stamps (layer-1 kernel)
int g_mlp_debug_flags = 0;     // ktb_debug_set_ptr(1, flags): bit 3 = layer-1 kernel issues no TMA stores (timing diagnostic, wrong results)
int g_mlp_arrive_mode = 1;     // ktb_set_tuning key 25: semantics of the epilogue's remote mbarrier arrives (see mbar_arrive_remote):
                               // 0 = release.cluster everywhere, 1 = CTA-scope release for the TMEM hand-backs (default: +6 %,
                               // profiles/r2n_probe_arrive.log), 2 = also for c_ready; bit 2 = pipelined TMEM loads in the fused epilogue
int g_mlp_l1_bres = 2;         // ktb_set_tuning key 24: 1 = layer-1 form of the CTA-pair kernel for K = 256 (default): W1 slice
                               // resident, half-tile bulk groups + pipelined TMEM loads in the epilogue; bit-identical, +3 %
int g_mlp_2sm = 1;             // ktb_set_tuning key 11: 1 = CTA-pair (cta_group::2) kernel for the 256-wide layers (default)
int g_mlp_tma_store = 1;       // ktb_set_tuning key 10: 1 = TMA-store epilogue for the 256-wide layers (default)
int g_mlp_persistent = 1;      // ktb_set_tuning key 7: 1 = persistent double-buffered kernel, 0 = one tile per CTA

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// the same load WITHOUT the wait: the caller overlaps it with work on registers of an earlier load and issues
// tmem_wait_ld() before touching r
__device__ __forceinline__ void tmem_ld_32x32b_x32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major) | [32,46) SBO >> 4
//   [46,48) version = 1 (sm_100) | [61,64) layout type = 2 (SWIZZLE_128B)
// SBO = 1024 bytes: 8 rows x 128 bytes per swizzle atom, atoms stacked along M/N.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(const void* smem_tile) {
  const uint32_t addr = smem_u32(smem_tile);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, K-major both.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int BLOCK_N, int STAGES>
struct MlpSmem {
  static constexpr int kABytes = kMlpBlockM * kMlpBlockK * 2;   // 16 KiB
  static constexpr int kBBytes = BLOCK_N * kMlpBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarrierOff = STAGES * kStageBytes;
  static constexpr int kTotal = kBarrierOff + (2 * STAGES + 4) * 8 + 16 + 1024 /* alignment slack */;
};

// C[m0:m0+128, n0:n0+BLOCK_N] = act(A[m0:.., :K] · B[n0:.., :K]ᵀ), one output tile per CTA.
template <int BLOCK_N, int STAGES, bool RELU>
__global__ void __launch_bounds__(kMlpThreads)
    gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        __nv_bfloat16* __restrict__ C, int ldc, int K) {
  using S = MlpSmem<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles must be 1024-byte aligned
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBarrierOff);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kMlpBlockM;
  const int n0 = blockIdx.y * BLOCK_N;
  const int num_kb = K / kMlpBlockK;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
        uint8_t* a_dst = smem + (size_t)s * S::kStageBytes;
        uint8_t* b_dst = a_dst + S::kABytes;
        mbar_expect_tx(&full[s], S::kStageBytes);
        tma_load_2d(a_dst, &map_a, kb * kMlpBlockK, m0, &full[s]);
        tma_load_2d(b_dst, &map_b, kb * kMlpBlockK, n0, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kMlpBlockM, BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&full[s], (kb / STAGES) & 1);
        tc_fence_after();
        const uint8_t* a_src = smem + (size_t)s * S::kStageBytes;
        const uint64_t adesc = make_smem_desc_sw128(a_src);
        const uint64_t bdesc = make_smem_desc_sw128(a_src + S::kABytes);
#pragma unroll
        for (int k = 0; k < kMlpBlockK / kMlpUmmaK; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the >>4 address field
          umma_f16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
        }
        umma_commit(&empty[s]);  // frees the stage when these MMAs have read it
      }
      umma_commit(tmem_full);    // accumulator complete
    }
  } else {
    // ===== epilogue: TMEM → registers → (ReLU) → bf16 → global =====
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int quarter = warp & 3;                // TMEM lanes this warp may access
    const int row = m0 + quarter * 32 + lane;
    __nv_bfloat16* crow = C + (size_t)row * ldc + n0;
#pragma unroll 1
    for (int c = 0; c < BLOCK_N; c += 32) {
      uint32_t acc[32];
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c, acc);
      uint32_t packed[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float lo = __uint_as_float(acc[2 * j]);
        float hi = __uint_as_float(acc[2 * j + 1]);
        if (RELU) {
          lo = fmaxf(lo, 0.f);
          hi = fmaxf(hi, 0.f);
        }
        __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
        packed[j] = *reinterpret_cast<uint32_t*>(&v);
      }
      uint4* dst = reinterpret_cast<uint4*>(crow + c);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BLOCK_N);
  }
}

// Persistent form: one CTA per SM loops over output tiles (n fastest, so the CTAs running at the same
// time share A rows in L2); the fp32 accumulator is DOUBLE-BUFFERED in TMEM (2 x BLOCK_N columns) so the
// epilogue of tile t (tcgen05.ld → ReLU → bf16 → global) overlaps the TMA/MMA main loop of tile t+1.
//   tmem_full[a]  : MMA warp → epilogue   (tcgen05.commit, count 1)
//   tmem_empty[a] : epilogue → MMA warp   (one arrive per epilogue warp, count 4)
template <int BLOCK_N, int STAGES, bool RELU, int EPI_GROUPS>
__global__ void __launch_bounds__(64 + 128 * EPI_GROUPS)
    gemm_bf16_tn_persistent_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                   __nv_bfloat16* __restrict__ C, int ldc, int K, int tiles_m, int tiles_n) {
  using S = MlpSmem<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::kBarrierOff);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;        // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = K / kMlpBlockK;
  const int num_tiles = tiles_m * tiles_n;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], 4 * EPI_GROUPS);
    mbar_init(&tmem_empty[1], 4 * EPI_GROUPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, 2 * BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;  // running k-block counter across tiles → stage / parity
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * kMlpBlockM;
        const int n0 = (tile % tiles_n) * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          uint8_t* a_dst = smem + (size_t)s * S::kStageBytes;
          mbar_expect_tx(&full[s], S::kStageBytes);
          tma_load_2d(a_dst, &map_a, kb * kMlpBlockK, m0, &full[s]);
          tma_load_2d(a_dst + S::kABytes, &map_b, kb * kMlpBlockK, n0, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kMlpBlockM, BLOCK_N);
      int it = 0, t = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
        const int as = t & 1;
        mbar_wait(&tmem_empty[as], ((t >> 1) & 1) ^ 1);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint8_t* a_src = smem + (size_t)s * S::kStageBytes;
          const uint64_t adesc = make_smem_desc_sw128(a_src);
          const uint64_t bdesc = make_smem_desc_sw128(a_src + S::kABytes);
#pragma unroll
          for (int k = 0; k < kMlpBlockK / kMlpUmmaK; ++k)
            umma_f16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
          umma_commit(&empty[s]);
        }
        umma_commit(&tmem_full[as]);
      }
    }
  } else {
    // EPI_GROUPS warpgroups of 4 warps: warp w reads TMEM lanes 32*(w%4)..+31; group g takes the g-th slice
    // of the tile's columns, so two groups halve the epilogue time of a tile
    const int quarter = warp & 3;
    const int group = (warp - 2) >> 2;
    constexpr int kColsPerGroup = BLOCK_N / EPI_GROUPS;
    int t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int as = t & 1;
      const int m0 = (tile / tiles_n) * kMlpBlockM;
      const int n0 = (tile % tiles_n) * BLOCK_N;
      mbar_wait(&tmem_full[as], (t >> 1) & 1);
      tc_fence_after();
      __nv_bfloat16* crow = C + (size_t)(m0 + quarter * 32 + lane) * ldc + n0;
#pragma unroll 1
      for (int c = group * kColsPerGroup; c < (group + 1) * kColsPerGroup; c += 32) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BLOCK_N + c), acc);
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float lo = __uint_as_float(acc[2 * j]);
          float hi = __uint_as_float(acc[2 * j + 1]);
          if (RELU) {
            lo = fmaxf(lo, 0.f);
            hi = fmaxf(hi, 0.f);
          }
          __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
          packed[j] = *reinterpret_cast<uint32_t*>(&v);
        }
        uint4* dst = reinterpret_cast<uint4*>(crow + c);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);  // this warp's quarter of the accumulator is free
    }
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BLOCK_N);
  }
}

// Persistent kernel with a TMA-STORE epilogue (BLOCK_N = 256 layers, whose output is a plain local matrix):
// the epilogue warps convert the accumulator to bf16 into a 128x256 shared-memory tile laid out as four
// SWIZZLE_128B boxes of 64 columns, and one thread writes it out with cp.async.bulk.tensor (full 128-byte lines
// instead of 16-byte stores 2 KiB apart).  Three operand stages (144 KiB) + the 64 KiB C tile fit in shared memory.
constexpr int kEpiBarrier = 1;   // named barrier of the 128 epilogue threads

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d_s(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_src),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync %0, 128;" ::"n"(kEpiBarrier) : "memory"); }
template <int NTHREADS>
__device__ __forceinline__ void epi_barrier_n() {
  asm volatile("bar.sync %0, %1;" ::"n"(kEpiBarrier), "n"(NTHREADS) : "memory");
}

template <int STAGES, bool RELU>
__global__ void __launch_bounds__(kMlpThreads)
    gemm_bf16_tn_tmastore_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                 const __grid_constant__ CUtensorMap map_c, int K, int tiles_m, int tiles_n) {
  constexpr int BLOCK_N = 256;
  using S = MlpSmem<BLOCK_N, STAGES>;
  constexpr int kCBytes = kMlpBlockM * BLOCK_N * 2;        // 64 KiB
  constexpr int kBoxBytes = kMlpBlockM * 64 * 2;           // one 128 x 64 box
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ctile = smem + S::kBarrierOff;                  // stages end on a 1024-byte boundary
  uint64_t* full = reinterpret_cast<uint64_t*>(ctile + kCBytes);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = K / kMlpBlockK;
  const int num_tiles = tiles_m * tiles_n;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
    prefetch_tensormap(&map_c);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], 4);
    mbar_init(&tmem_empty[1], 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_holder, 2 * BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * kMlpBlockM;
        const int n0 = (tile % tiles_n) * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
          uint8_t* a_dst = smem + (size_t)s * S::kStageBytes;
          mbar_expect_tx(&full[s], S::kStageBytes);
          tma_load_2d(a_dst, &map_a, kb * kMlpBlockK, m0, &full[s]);
          tma_load_2d(a_dst + S::kABytes, &map_b, kb * kMlpBlockK, n0, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kMlpBlockM, BLOCK_N);
      int it = 0, t = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
        const int as = t & 1;
        mbar_wait(&tmem_empty[as], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint8_t* a_src = smem + (size_t)s * S::kStageBytes;
          const uint64_t adesc = make_smem_desc_sw128(a_src);
          const uint64_t bdesc = make_smem_desc_sw128(a_src + S::kABytes);
#pragma unroll
          for (int k = 0; k < kMlpBlockK / kMlpUmmaK; ++k)
            umma_f16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
          umma_commit(&empty[s]);
        }
        umma_commit(&tmem_full[as]);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;                    // row of the tile this thread owns
    const bool issuer = (warp == 2 && lane == 0);
    int t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int as = t & 1;
      const int m0 = (tile / tiles_n) * kMlpBlockM;
      const int n0 = (tile % tiles_n) * BLOCK_N;
      mbar_wait(&tmem_full[as], (t >> 1) & 1);
      tc_fence_after();
      if (issuer) bulk_wait_read<0>();                      // the previous tile's stores have read the C tile
      epi_barrier();
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BLOCK_N + c), acc);
        uint8_t* box = ctile + (c >> 6) * kBoxBytes + row * 128;   // this row inside the 64-column box
        const int chunk0 = (c & 63) >> 3;                          // first 16-byte chunk (0 or 4)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float lo = __uint_as_float(acc[8 * q + 2 * j]);
            float hi = __uint_as_float(acc[8 * q + 2 * j + 1]);
            if (RELU) {
              lo = fmaxf(lo, 0.f);
              hi = fmaxf(hi, 0.f);
            }
            __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
            pk[j] = *reinterpret_cast<uint32_t*>(&v);
          }
          const int phys = (chunk0 + q) ^ (row & 7);               // SWIZZLE_128B: chunk index XOR (row mod 8)
          *reinterpret_cast<uint4*>(box + phys * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);          // accumulator drained: the MMA warp may refill it
      fence_proxy_async_smem();                             // generic-proxy writes → visible to the TMA store
      epi_barrier();
      if (issuer) {
#pragma unroll
        for (int b = 0; b < BLOCK_N / 64; ++b) tma_store_2d(&map_c, ctile + b * kBoxBytes, n0 + 64 * b, m0);
        bulk_commit();
      }
    }
    if (issuer) bulk_wait_all<0>();
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BLOCK_N);
  }
}

// ---- 2-SM form: a CTA PAIR (cluster of 2) computes a 256 x 256 tile with tcgen05.mma.cta_group::2 -------------
// Why: with one CTA per tile the 128x256x16 UMMA reads 12 KiB of operands from shared memory per 128 cycles while TMA
// refills 48 KiB per k-block — 192 B/clk against a 128 B/clk shared-memory port (tensor pipe <= 67 %; measured 58 %).
// In a pair each CTA holds ITS 128 rows of A and HALF of B (128 of the 256 N rows): 64 + 64 B/clk, the port's limit.
//   * both CTAs run a TMA producer: own A rows + own half of B into own shared memory, all transaction bytes
//     reported to the LEADER's full barrier (cp.async.bulk.tensor...cta_group::2, barrier address with the peer
//     bit cleared);
//   * only the leader issues tcgen05.mma.cta_group::2 (M = 256); the hardware reads A/B from both CTAs at the same
//     shared-memory offsets and writes rows 0-127 to the leader's TMEM, rows 128-255 to the peer's;
//   * tcgen05.commit...multicast::cluster releases the stage in BOTH CTAs and publishes the accumulator to both
//     epilogues; each CTA's epilogue drains its own TMEM half through a swizzled shared C tile + TMA store and
//     arrives on the leader's tmem_empty barrier (the peer via mapa).
// Every wait is bounded (trap instead of hanging the GPU if the protocol were ever violated).
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  __syncwarp();   // .aligned: the whole warp must arrive together (role branches leave lanes diverged)
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Execution-only cluster barrier (no release/acquire: ptxas emits no MEMBAR.ALL.GPU): enough for "nobody exits while
// the peer may still touch my shared memory / barriers".
__device__ __forceinline__ void cluster_sync_relaxed() {
  __syncwarp();
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Remote arrive whose only job is to hand a TMEM buffer back (ordering comes from tcgen05.fence::before_thread_sync):
// `.release.cluster` makes ptxas emit MEMBAR.ALL.GPU + ERRBAR in front of the arrive — the issuing lane then waits for
// every global write it has in flight, TMA stores included — while the default semantics (release at CTA scope) cost a
// MEMBAR.ALL.CTA.  mode 0 keeps the cluster-scope release (ktb_set_tuning key 25).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr, int cta_sem) {
  if (cta_sem)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
  else
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// two fp32 -> packed bf16 (round to nearest even), then ReLU on the PACKED pair (one HMNMX2 for two values instead of
// two FMNMX: the epilogue warps are ALU-bound).  Same bits as rounding relu(x): rounding is monotonic and keeps the sign,
// max(-0, +0) = +0 and max(NaN, 0) = 0 in both forms.
template <bool RELU>
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  if (RELU) v = __hmax2(v, __floats2bfloat162_rn(0.f, 0.f));
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                                uint32_t leader_bar_addr) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(map), "r"(leader_bar_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of a CONVERGED warp (elect.sync).  The MMA issuer runs its loops with the whole warp so that the operands of
// tcgen05.mma stay warp-uniform for the compiler: issued from an `if (lane == 0)` region, every UTCHMMA / UTCBAR is
// wrapped in an ELECT + 5 x R2UR.BROADCAST + BRA.U.ANY waterfall loop and the single thread issues one MMA per ~177
// cycles against the 128 the tensor pipe needs for a 256x256x16 step (tools/probe_trace.py).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n.reg .b32 rx;\n.reg .pred px;\n"
      "elect.sync rx|px, 0xFFFFFFFF;\n"
      "selp.b32 %0, 1, 0, px;\n}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {   // arrives on `bar` in BOTH CTAs of the pair
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

template <int STAGES, bool RELU, int EPI_GROUPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 128 * EPI_GROUPS)
    gemm_bf16_tn_2sm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                            const __grid_constant__ CUtensorMap map_c, int K, int tiles_m, int tiles_n) {
  constexpr int BLOCK_N = 256;                              // per pair; each CTA stages 128 of these B rows
  constexpr int kABytes = kMlpBlockM * kMlpBlockK * 2;      // 16 KiB: this CTA's 128 rows of A
  constexpr int kBBytes = 128 * kMlpBlockK * 2;             // 16 KiB: this CTA's half of B
  constexpr int kStageBytes = kABytes + kBBytes;            // 32 KiB per CTA per stage
  constexpr int kCBytes = kMlpBlockM * BLOCK_N * 2;         // 64 KiB C tile of this CTA (128 rows x 256 cols)
  constexpr int kBoxBytes = kMlpBlockM * 64 * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ctile = smem + STAGES * kStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(ctile + kCBytes);   // used on the leader only
  uint64_t* empty = full + STAGES;                                  // per CTA
  uint64_t* tmem_full = empty + STAGES;                             // [2], per CTA
  uint64_t* tmem_empty = tmem_full + 2;                             // [2], used on the leader only
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform (see the layer-1 kernel)
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int num_kb = K / kMlpBlockK;
  const int num_tiles = tiles_m * tiles_n;                 // 256 x 256 tiles
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
    prefetch_tensormap(&map_c);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);          // the leader's producer arms it with the bytes of BOTH CTAs' loads
      mbar_init(&empty[s], 1);         // one multicast commit
    }
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], 8 * EPI_GROUPS);      // 4 epilogue warps per group x 2 CTAs
    mbar_init(&tmem_empty[1], 8 * EPI_GROUPS);
    fence_barrier_init();
  }
  cluster_sync_all();                  // barriers of both CTAs are initialised before anyone signals across
  if (warp == 1) tmem_alloc_2sm(tmem_holder, 2 * BLOCK_N);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    {                                 // whole warp walks the loop, one elected lane issues (see elect_one_sync)
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m0 = (tile / tiles_n) * 256 + (int)rank * 128;
        const int n0 = (tile % tiles_n) * BLOCK_N + (int)rank * 128;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait_bounded(&empty[s], ((it / STAGES) & 1) ^ 1);
          uint8_t* a_dst = smem + (size_t)s * kStageBytes;
          // The leader expects the bytes of all four loads of this k-block (its own two and the peer's two); the
          // peer's complete_tx may land first (the tx-count goes transiently negative, the phase cannot complete
          // before the leader's arrive).  No remote arrive sits on the peer's critical path.
          const uint32_t leader_full_tma = smem_u32(&full[s]) & 0xFEFFFFFFu;   // peer bit cleared → CTA 0's barrier
          if (elect_one_sync()) {
            if (leader) mbar_expect_tx(&full[s], 2 * kStageBytes);
            tma_load_2d_2sm(a_dst, &map_a, kb * kMlpBlockK, m0, leader_full_tma);
            tma_load_2d_2sm(a_dst + kABytes, &map_b, kb * kMlpBlockK, n0, leader_full_tma);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (leader) {                     // whole warp, one elected lane issues
      constexpr uint32_t idesc = make_idesc_bf16(256, BLOCK_N);
      int it = 0, t = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++t) {
        const int as = t & 1;
        mbar_wait_bounded(&tmem_empty[as], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait_bounded(&full[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint8_t* a_src = smem + (size_t)s * kStageBytes;
          const uint64_t adesc = make_smem_desc_sw128(a_src);
          const uint64_t bdesc = make_smem_desc_sw128(a_src + kABytes);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < kMlpBlockK / kMlpUmmaK; ++k)
              umma_f16_2sm(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
            umma_commit_2sm(&empty[s]);          // stage free in both CTAs
            if (kb == num_kb - 1) umma_commit_2sm(&tmem_full[as]);       // accumulator ready in both CTAs
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===== epilogue (both CTAs): own TMEM half → bf16 → swizzled shared C tile → TMA store =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int group = (warp - 2) >> 2;                     // each warpgroup converts its slice of the columns
    constexpr int kColsPerGroup = BLOCK_N / EPI_GROUPS;
    const bool issuer = (warp == 2 && lane == 0);
    int t = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++t) {
      const int as = t & 1;
      const int m0 = (tile / tiles_n) * 256 + (int)rank * 128;
      const int n0 = (tile % tiles_n) * BLOCK_N;
      mbar_wait_bounded(&tmem_full[as], (t >> 1) & 1);
      tc_fence_after();
      if (issuer) bulk_wait_read<0>();
      epi_barrier_n<128 * EPI_GROUPS>();
#pragma unroll 1
      for (int c = group * kColsPerGroup; c < (group + 1) * kColsPerGroup; c += 32) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BLOCK_N + c), acc);
        uint8_t* box = ctile + (c >> 6) * kBoxBytes + row * 128;
        const int chunk0 = (c & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float lo = __uint_as_float(acc[8 * q + 2 * j]);
            float hi = __uint_as_float(acc[8 * q + 2 * j + 1]);
            if (RELU) {
              lo = fmaxf(lo, 0.f);
              hi = fmaxf(hi, 0.f);
            }
            __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
            pk[j] = *reinterpret_cast<uint32_t*>(&v);
          }
          const int phys = (chunk0 + q) ^ (row & 7);
          *reinterpret_cast<uint4*>(box + phys * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(map_to_cta(smem_u32(&tmem_empty[as]), 0), 1);   // leader's barrier, CTA-scope release
      fence_proxy_async_smem();
      epi_barrier_n<128 * EPI_GROUPS>();
      if (issuer) {
#pragma unroll
        for (int b = 0; b < BLOCK_N / 64; ++b) tma_store_2d(&map_c, ctile + b * kBoxBytes, n0 + 64 * b, m0);
        bulk_commit();
      }
    }
    if (issuer) bulk_wait_all<0>();
  }

  tc_fence_before();
  cluster_sync_all();                  // both CTAs are done with TMEM and with each other's barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * BLOCK_N);
  }
}

// ---- layer-1 form of the CTA-pair kernel: K = 256 (4 k-blocks), B RESIDENT, A ring two tiles deep -------------------
// With K = 256 a tile is four k-blocks and a 4-stage A+B ring is exactly one tile deep: stage s can be refilled for
// tile t+1 only after MMA s of tile t has retired, so every tile exposes a TMA round trip (ncu r2e: 3.6 us per tile,
// of which the MMAs are 1.1 us; DRAM 28 %, L2 34 %).  Here the pair walks tiles n-SLOWEST over a contiguous range, so
// the 256-row slice of W1 it multiplies with stays in shared memory (64 KiB per CTA, reloaded only when the column
// block changes, at most twice per launch) and the freed space makes the A ring AST x 16 KiB = 1.5 tiles deep (AST = 6)
// beside the unchanged 64 KiB C staging tile.  MMA order inside a tile is the pair kernel's: results are bit-identical.
template <int AST, bool RELU, int STORE, int EPI_WARPS = 4>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EPI_WARPS, 1)
    gemm_bf16_tn_2sm_bres_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                 const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_c32,
                                 int tiles_m, int tiles_n, int arrive_mode, unsigned long long* trace) {
  // trace (developer tool, tools/probe_trace.py; nullptr in every product call): the leader CTA of pair 0 records
  // clock64() at the hand-offs of its first 64 tiles, trace[event * 64 + tile]
#define KTB_TR(ev, tile) \
  do { if (trace != nullptr && blockIdx.x == 0) trace[(ev) * 64 + ((tile) & 63)] = (unsigned long long)clock64(); } while (0)
#define KTB_TG(k) \
  do { if (trace != nullptr && threadIdx.x == 0) trace[1024 + blockIdx.x * 4 + (k)] = globaltimer_ns(); } while (0)
  KTB_TG(0);
  constexpr int BLOCK_N = 256;
  constexpr int KB = 4;                                     // K = 256
  constexpr int kABytes = kMlpBlockM * kMlpBlockK * 2;      // 16 KiB: this CTA's 128 rows of one A k-block
  constexpr int kBBytes = 128 * kMlpBlockK * 2;             // 16 KiB: this CTA's half of one B k-block
  // STORE: 0 = one thread stores 128 x 64 boxes, 1 = every epilogue warp stores its own 32 rows in two 128-column
  // halves (64 KiB staging), 2 = in four 64-column boxes rotating through 8 KiB per warp (32 KiB staging: the other
  // 32 KiB go to the A ring, which the measured 1.9 us TMA latency under this kernel's own write stream needs)
  constexpr bool WARP_STORE = STORE >= 1;
  constexpr int kCBytes = (STORE == 2) ? kMlpBlockM * BLOCK_N : kMlpBlockM * BLOCK_N * 2;   // 32 / 64 KiB C staging
  constexpr int kBoxBytes = kMlpBlockM * 64 * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_ring = smem;
  uint8_t* b_res = smem + AST * kABytes;
  uint8_t* ctile = b_res + KB * kBBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(ctile + kCBytes);   // [AST], used on the leader only
  uint64_t* empty = full + AST;                                     // [AST], per CTA
  uint64_t* tmem_full = empty + AST;                                // [2], per CTA
  uint64_t* tmem_empty = tmem_full + 2;                             // [2], used on the leader only
  uint64_t* b_full = tmem_empty + 2;                                // [1], used on the leader only
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(b_full + 1);

  // warp index through a shuffle: provably warp-uniform for the compiler (uniform branches, uniform-register operands)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  // contiguous range of the n-slowest tile enumeration t = n * tiles_m + m
  const long long total = (long long)tiles_m * tiles_n;
  const int t_begin = (int)(total * pair / num_pairs);
  const int t_end = (int)(total * (pair + 1) / num_pairs);

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
    prefetch_tensormap(&map_c);
    prefetch_tensormap(&map_c32);
#pragma unroll
    for (int s = 0; s < AST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], 2 * EPI_WARPS);     // one arrival per epilogue warp of both CTAs
    mbar_init(&tmem_empty[1], 2 * EPI_WARPS);
    mbar_init(b_full, 1);
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 1) tmem_alloc_2sm(tmem_holder, 2 * BLOCK_N);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  KTB_TG(1);

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    {                                 // whole warp walks the loop, one elected lane issues (see elect_one_sync)
      int it = 0, cur_n = -1;
      for (int t = t_begin; t < t_end; ++t) {
        const int n = t / tiles_m;
        const int m0 = (t % tiles_m) * 256 + (int)rank * 128;
        if (lane == 0) KTB_TR(8, t - t_begin);
        if (n != cur_n) {
          // every MMA that reads the resident B has retired once the commit of the previous tile's last k-block arrived
          if (it > 0) mbar_wait_bounded(&empty[(it - 1) % AST], ((it - 1) / AST) & 1);
          const uint32_t leader_b_full = smem_u32(b_full) & 0xFEFFFFFFu;
          const int n0 = n * BLOCK_N + (int)rank * 128;
          if (elect_one_sync()) {
            if (leader) mbar_expect_tx(b_full, 2 * KB * kBBytes);
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) tma_load_2d_2sm(b_res + kb * kBBytes, &map_b, kb * kMlpBlockK, n0, leader_b_full);
          }
          __syncwarp();
          cur_n = n;
        }
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % AST;
          mbar_wait_bounded(&empty[s], ((it / AST) & 1) ^ 1);
          const uint32_t leader_full_tma = smem_u32(&full[s]) & 0xFEFFFFFFu;
          if (elect_one_sync()) {
            if (leader) mbar_expect_tx(&full[s], 2 * kABytes);
            tma_load_2d_2sm(a_ring + (size_t)s * kABytes, &map_a, kb * kMlpBlockK, m0, leader_full_tma);
          }
          __syncwarp();
        }
        if (lane == 0) KTB_TR(9, t - t_begin);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (leader) {                     // the WHOLE warp walks the loops (uniform operands), one elected lane issues
      constexpr uint32_t idesc = make_idesc_bf16(256, BLOCK_N);
      int it = 0, tt = 0, cur_n = -1, b_loads = 0;
      for (int t = t_begin; t < t_end; ++t, ++tt) {
        const int n = t / tiles_m;
        const int as = tt & 1;
        mbar_wait_bounded(&tmem_empty[as], ((tt >> 1) & 1) ^ 1);
        tc_fence_after();
        if (lane == 0) KTB_TR(0, tt);
        if (n != cur_n) {
          mbar_wait_bounded(b_full, b_loads & 1);
          tc_fence_after();
          ++b_loads;
          cur_n = n;
        }
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int s = it % AST;
          mbar_wait_bounded(&full[s], (it / AST) & 1);
          tc_fence_after();
          if (trace != nullptr && lane == 0) {
            if (kb == 0) KTB_TR(1, tt);
            if (kb == 1) KTB_TR(12, tt);
            if (kb == 2) KTB_TR(13, tt);
            if (kb == KB - 1) KTB_TR(2, tt);
          }
          const uint64_t adesc = make_smem_desc_sw128(a_ring + (size_t)s * kABytes);
          const uint64_t bdesc = make_smem_desc_sw128(b_res + kb * kBBytes);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < kMlpBlockK / kMlpUmmaK; ++k)
              umma_f16_2sm(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
            umma_commit_2sm(&empty[s]);          // A stage free in both CTAs
            if (kb == KB - 1) umma_commit_2sm(&tmem_full[as]);       // accumulator ready in both CTAs
          }
          __syncwarp();
        }
        if (lane == 0) KTB_TR(3, tt);
      }
    }
  } else {
    // ===== epilogue (both CTAs): own TMEM half → bf16 → swizzled shared C tile → TMA store =====
    // Two changes against the pair kernel's epilogue (same bytes, same rounding): (1) the tile leaves in two 128-column
    // halves, each its own bulk group: the conversion of one half overlaps the TMA store of the other (before a half's
    // boxes are rewritten only the store that read THEM, two groups back, has to be done: wait_group.read 1);
    // (2) the TMEM loads are software-pipelined: the load of chunk j+1 is in flight while chunk j is converted.
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const bool issuer = (warp == 2 && lane == 0);
    const uint32_t ctile_s = smem_u32(ctile);
    int tt = 0;
    auto convert = [&](const uint32_t (&acc)[32], int c) {
      const uint32_t box = ctile_s + (uint32_t)((c >> 6) * kBoxBytes + row * 128);
      const int chunk0 = (c & 63) >> 3;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t pk[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          pk[jj] = pack_bf16x2_relu<RELU>(__uint_as_float(acc[8 * q + 2 * jj]), __uint_as_float(acc[8 * q + 2 * jj + 1]));
        }
        const int phys = (chunk0 + q) ^ (row & 7);
        st_shared_v4(box + (uint32_t)(phys * 16), pk[0], pk[1], pk[2], pk[3]);     // STS.128, not a generic store
      }
    };
    if constexpr (STORE == 2) {
      // four bulk groups per tile and warp: group g = columns [64g, 64g + 64) -> box g & 1 of this warp's 8 KiB
      // (32 rows x 128 B, SWIZZLE_128B) -> one TMA store; a box is rewritten once the group two back has been read
      const uint32_t cwarp = ctile_s + (uint32_t)(quarter * 8192 + lane * 128);
      auto convert_q = [&](const uint32_t (&acc)[32], uint32_t boxrow, int chunk0) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t pk[4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            pk[jj] = pack_bf16x2_relu<RELU>(__uint_as_float(acc[8 * q + 2 * jj]), __uint_as_float(acc[8 * q + 2 * jj + 1]));
          }
          const int phys = (chunk0 + q) ^ (lane & 7);
          st_shared_v4(boxrow + (uint32_t)(phys * 16), pk[0], pk[1], pk[2], pk[3]);
        }
      };
      for (int t = t_begin; t < t_end; ++t, ++tt) {
        const int as = tt & 1;
        const int m0 = (t % tiles_m) * 256 + (int)rank * 128 + quarter * 32;
        const int n0 = (t / tiles_m) * BLOCK_N;
        mbar_wait_bounded(&tmem_full[as], (tt >> 1) & 1);
        tc_fence_after();
        if (warp == 2 && lane == 0) KTB_TR(4, tt);
        const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BLOCK_N);
        uint32_t acc0[32], acc1[32];
        tmem_ld_32x32b_x32_nowait(tbase, acc0);
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          bulk_wait_read<1>();
          __syncwarp();
          const uint32_t boxrow = cwarp + (uint32_t)((g & 1) * 4096);
          tmem_wait_ld();
          tmem_ld_32x32b_x32_nowait(tbase + (uint32_t)(64 * g + 32), acc1);
          convert_q(acc0, boxrow, 0);
          tmem_wait_ld();
          if (g < 3) tmem_ld_32x32b_x32_nowait(tbase + (uint32_t)(64 * g + 64), acc0);
          convert_q(acc1, boxrow, 4);
          if (g == 3) {                  // every TMEM read of this warp from this accumulator buffer is done
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(map_to_cta(smem_u32(&tmem_empty[as]), 0), arrive_mode & 3);
            if (warp == 2 && lane == 0) KTB_TR(6, tt);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (!(arrive_mode & 8)) {
            const uint32_t src = ctile_s + (uint32_t)(quarter * 8192 + (g & 1) * 4096);
            if (elect_one_sync()) {
              tma_store_2d_s(&map_c32, src, n0 + 64 * g, m0);
              bulk_commit();
            }
            __syncwarp();
          }
          if (warp == 2 && lane == 0 && g == 3) KTB_TR(7, tt);
        }
      }
      if (arrive_mode & 16) bulk_wait_all<0>(); else bulk_wait_read<0>();
    } else {
    for (int t = t_begin; t < t_end; ++t, ++tt) {
      const int as = tt & 1;
      const int m0 = (t % tiles_m) * 256 + (int)rank * 128;
      const int n0 = (t / tiles_m) * BLOCK_N;
      mbar_wait_bounded(&tmem_full[as], (tt >> 1) & 1);
      tc_fence_after();
      if (warp == 2 && lane == 0) KTB_TR(4, tt);
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BLOCK_N);
      // EPI_WARPS == 8: warps 2-5 own the first 128 columns, warps 6-9 the second (a warp reaches the TMEM lanes of
      // quarter warp % 4 only, so warps w and w + 4 share a quarter); each warp does ONE half per tile
      const int h_begin = (EPI_WARPS == 8) ? ((warp - 2) >> 2) : 0;
      const int h_end = (EPI_WARPS == 8) ? h_begin + 1 : 2;
#pragma unroll 1
      for (int h = h_begin; h < h_end; ++h) {
        uint32_t acc0[32], acc1[32];
        tmem_ld_32x32b_x32_nowait(tbase + (uint32_t)(h * 128), acc0);      // in flight while the staging half drains
        if constexpr (WARP_STORE) {
          // every epilogue warp owns its 32 rows end to end (own bulk groups, own TMA stores of 64 x 32 boxes): no
          // CTA-wide barrier in the epilogue, a slow warp delays nobody
          // every lane executes the wait (only the elected issuer owns bulk groups; elect.sync picks the same lane for
          // the same member mask every time), then the warp re-converges
          if constexpr (EPI_WARPS == 8) bulk_wait_read<0>(); else bulk_wait_read<1>();
          __syncwarp();
          if (warp == 2 && lane == 0) KTB_TR(10 + h, tt);
        } else {
          if (issuer) bulk_wait_read<1>();
          epi_barrier_n<128>();
        }
        tmem_wait_ld();
        tmem_ld_32x32b_x32_nowait(tbase + (uint32_t)(h * 128 + 32), acc1);
        convert(acc0, h * 128);
        tmem_wait_ld();
        tmem_ld_32x32b_x32_nowait(tbase + (uint32_t)(h * 128 + 64), acc0);
        convert(acc1, h * 128 + 32);
        tmem_wait_ld();
        tmem_ld_32x32b_x32_nowait(tbase + (uint32_t)(h * 128 + 96), acc1);
        convert(acc0, h * 128 + 64);
        tmem_wait_ld();
        convert(acc1, h * 128 + 96);
        if (warp == 2 && lane == 0 && h == 0) KTB_TR(5, tt);
        if (h == h_end - 1) {            // every TMEM read of this warp from this accumulator buffer is done
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(map_to_cta(smem_u32(&tmem_empty[as]), 0), arrive_mode & 3);   // leader's barrier
          if (warp == 2 && lane == 0) KTB_TR(6, tt);
        }
        fence_proxy_async_smem();
        if constexpr (WARP_STORE) {
          __syncwarp();
          // operands computed by the whole warp (uniform registers), issued by lane 0, which owns the bulk groups
          const int r0 = quarter * 32;            // this warp's rows inside the CTA's 128
          const uint32_t src0 = ctile_s + (uint32_t)((2 * h) * kBoxBytes + r0 * 128);
          const int sc0 = n0 + 128 * h, sc1 = m0 + r0;
          if (!(arrive_mode & 8)) {                // bit 3: developer diagnostic (no stores), set through ktb_debug_set_ptr only
            if (elect_one_sync()) {                // one active lane: plain uniform-register operands, no waterfall loop
              tma_store_2d_s(&map_c32, src0, sc0, sc1);
              tma_store_2d_s(&map_c32, src0 + (uint32_t)kBoxBytes, sc0 + 64, sc1);
              bulk_commit();
            }
            __syncwarp();
          }
          if (warp == 2 && lane == 0 && h == 1) KTB_TR(7, tt);
        } else {
          epi_barrier_n<128>();
          if (issuer) {
            tma_store_2d(&map_c, ctile + (2 * h) * kBoxBytes, n0 + 128 * h, m0);
            tma_store_2d(&map_c, ctile + (2 * h + 1) * kBoxBytes, n0 + 128 * h + 64, m0);
            bulk_commit();
          }
        }
      }
    }
    // shared memory must outlive the TMA engine's READS only; the writes complete with the grid (bit 4 of the mode
    // restores the full completion wait: it held every CTA ~6 us at the end of the kernel, tools/probe_trace.py)
    if (WARP_STORE || issuer) { if (arrive_mode & 16) bulk_wait_all<0>(); else bulk_wait_read<0>(); }
    }
  }

  KTB_TG(2);
  if (trace != nullptr && lane == 0) trace[2048 + blockIdx.x * 8 + warp] = globaltimer_ns();   // per-warp arrival at the end
  tc_fence_before();
  if (arrive_mode & 32) cluster_sync_relaxed(); else cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * BLOCK_N);
  }
  KTB_TG(3);
}
#undef KTB_TR
#undef KTB_TG

__device__ __forceinline__ void tma_load_2d_2sm_mc(void* smem_dst, const CUtensorMap* map, int c0, int c1,
                                                   uint32_t leader_bar_addr, uint16_t cta_mask) {
  // same CTA-relative destination offset and barrier offset in every CTA of cta_mask; with cta_group::2 and the peer bit
  // of the barrier address cleared, each destination signals the barrier of ITS pair's leader
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(map), "r"(leader_bar_addr), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mask(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}

// Cluster-of-4 form of the CTA-pair kernel (opt-in, ktb_set_tuning key 15): two pairs share every B tile through TMA
// multicast.  Everything else (TMEM double buffering, TMA-store epilogue) is the pair kernel's.
template <int STAGES, bool RELU, int EPI_GROUPS>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(64 + 128 * EPI_GROUPS)
    gemm_bf16_tn_2sm_mc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                            const __grid_constant__ CUtensorMap map_c, int K, int tiles_m, int tiles_n) {
  constexpr int BLOCK_N = 256;                              // per pair; each CTA stages 128 of these B rows
  constexpr int kABytes = kMlpBlockM * kMlpBlockK * 2;      // 16 KiB: this CTA's 128 rows of A
  constexpr int kBBytes = 128 * kMlpBlockK * 2;             // 16 KiB: this CTA's half of B
  constexpr int kStageBytes = kABytes + kBBytes;            // 32 KiB per CTA per stage
  constexpr int kCBytes = kMlpBlockM * BLOCK_N * 2;         // 64 KiB C tile of this CTA (128 rows x 256 cols)
  constexpr int kBoxBytes = kMlpBlockM * 64 * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ctile = smem + STAGES * kStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(ctile + kCBytes);   // used on the leader only
  uint64_t* empty = full + STAGES;                                  // per CTA
  uint64_t* tmem_full = empty + STAGES;                             // [2], per CTA
  uint64_t* tmem_empty = tmem_full + 2;                             // [2], used on the leader only
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // cluster of two CTA pairs: pair `p` = rank >> 1 takes rows [p*256, p*256+256) of a 512 x 256 cluster tile, `r` =
  // rank & 1 is the CTA's place inside its pair.  Both pairs need the same 256 x 64 B tile per k-block: CTA (p, r)
  // fetches ONE QUARTER of it (64 rows) and multicasts it into the B half of CTAs (0, r) and (1, r), so each B byte
  // leaves the L2 once per cluster instead of once per pair (layer 2 is bound by the L2 -> SM fabric, not by the MMA).
  const uint32_t rank = cluster_ctarank();
  const uint32_t p = rank >> 1, r = rank & 1;
  const uint32_t leader_rank = rank & ~1u;
  const bool leader = (r == 0);
  const int num_kb = K / kMlpBlockK;
  const int num_tiles = tiles_m * tiles_n;                 // 512 x 256 cluster tiles
  const int pair = blockIdx.x >> 2;                        // (cluster index: both pairs walk the same tile sequence)
  const int num_pairs = gridDim.x >> 2;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
    prefetch_tensormap(&map_c);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);          // the leader's producer arms it with the bytes of BOTH CTAs' loads
      mbar_init(&empty[s], 2);         // one multicast commit from EACH pair: the stage is free cluster-wide
    }
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], 8 * EPI_GROUPS);      // 4 epilogue warps per group x 2 CTAs
    mbar_init(&tmem_empty[1], 8 * EPI_GROUPS);
    fence_barrier_init();
  }
  cluster_sync_all();                  // barriers of both CTAs are initialised before anyone signals across
  if (warp == 1) tmem_alloc_2sm(tmem_holder, 2 * BLOCK_N);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (lane == 0) {
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m0 = (tile / tiles_n) * 512 + (int)p * 256 + (int)r * 128;
        const int n0 = (tile % tiles_n) * BLOCK_N + (int)r * 128 + (int)p * 64;   // this CTA's quarter of B
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait_bounded(&empty[s], ((it / STAGES) & 1) ^ 1);
          uint8_t* a_dst = smem + (size_t)s * kStageBytes;
          // The pair leader expects every byte that lands in ITS pair for this k-block: two A tiles (one per CTA) and
          // four B quarters (two per CTA, one of them multicast from the other pair) = 2 * kStageBytes.
          if (leader) mbar_expect_tx(&full[s], 2 * kStageBytes);
          const uint32_t leader_full_tma = smem_u32(&full[s]) & 0xFEFFFFFFu;   // peer bit cleared → CTA 0's barrier
          tma_load_2d_2sm(a_dst, &map_a, kb * kMlpBlockK, m0, leader_full_tma);
          tma_load_2d_2sm_mc(a_dst + kABytes + p * (kBBytes / 2), &map_b, kb * kMlpBlockK, n0, leader_full_tma,
                             (uint16_t)((1u << r) | (1u << (2 + r))));
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(256, BLOCK_N);
      int it = 0, t = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++t) {
        const int as = t & 1;
        mbar_wait_bounded(&tmem_empty[as], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait_bounded(&full[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint8_t* a_src = smem + (size_t)s * kStageBytes;
          const uint64_t adesc = make_smem_desc_sw128(a_src);
          const uint64_t bdesc = make_smem_desc_sw128(a_src + kABytes);
#pragma unroll
          for (int k = 0; k < kMlpBlockK / kMlpUmmaK; ++k)
            umma_f16_2sm(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
          umma_commit_2sm_mask(&empty[s], (uint16_t)0xF);          // this pair is done with stage s, tell all 4 CTAs
        }
        umma_commit_2sm_mask(&tmem_full[as], (uint16_t)(3u << leader_rank));   // accumulator ready in both CTAs of the pair
      }
    }
  } else {
    // ===== epilogue (both CTAs): own TMEM half → bf16 → swizzled shared C tile → TMA store =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int group = (warp - 2) >> 2;                     // each warpgroup converts its slice of the columns
    constexpr int kColsPerGroup = BLOCK_N / EPI_GROUPS;
    const bool issuer = (warp == 2 && lane == 0);
    int t = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++t) {
      const int as = t & 1;
      const int m0 = (tile / tiles_n) * 512 + (int)p * 256 + (int)r * 128;
      const int n0 = (tile % tiles_n) * BLOCK_N;
      mbar_wait_bounded(&tmem_full[as], (t >> 1) & 1);
      tc_fence_after();
      if (issuer) bulk_wait_read<0>();
      epi_barrier_n<128 * EPI_GROUPS>();
#pragma unroll 1
      for (int c = group * kColsPerGroup; c < (group + 1) * kColsPerGroup; c += 32) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BLOCK_N + c), acc);
        uint8_t* box = ctile + (c >> 6) * kBoxBytes + row * 128;
        const int chunk0 = (c & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float lo = __uint_as_float(acc[8 * q + 2 * j]);
            float hi = __uint_as_float(acc[8 * q + 2 * j + 1]);
            if (RELU) {
              lo = fmaxf(lo, 0.f);
              hi = fmaxf(hi, 0.f);
            }
            __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
            pk[j] = *reinterpret_cast<uint32_t*>(&v);
          }
          const int phys = (chunk0 + q) ^ (row & 7);
          *reinterpret_cast<uint4*>(box + phys * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&tmem_empty[as]), leader_rank));   // pair leader's barrier
      fence_proxy_async_smem();
      epi_barrier_n<128 * EPI_GROUPS>();
      if (issuer) {
#pragma unroll
        for (int b = 0; b < BLOCK_N / 64; ++b) tma_store_2d(&map_c, ctile + b * kBoxBytes, n0 + 64 * b, m0);
        bulk_commit();
      }
    }
    if (issuer) bulk_wait_all<0>();
  }

  tc_fence_before();
  cluster_sync_all();                  // both CTAs are done with TMEM and with each other's barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * BLOCK_N);
  }
}

// ---- layer 2 + head in ONE kernel (opt-in, ktb_set_tuning key 18) -----------------------------------------------------
// h2 = relu(h1 @ W2^T) never goes to memory: the swizzled shared C tile the pair kernel stages for its TMA store is
// already a valid K-major SWIZZLE_128B A operand, so after the epilogue has written it the MMA thread multiplies it
// with the matching 256-column slice of W3 (logits[256 x 64] partial = C[256 x 256] @ W3[:, n0:n0+256]^T) into the
// first 64 TMEM columns of the accumulator buffer that was just drained.  The partial products of the four N tiles of
// a row block are summed in fp32 registers by the epilogue threads (one row each) and stored once as bf16.
//   schedule (per pair, tiles walk N fastest):  main(t) | epi-1(t): drain + C tile | head MMA(t) issued in the MIDDLE of
//   main(t+1) (the tensor pipe never waits for the epilogue) | epi-2(t): 64 columns -> registers, buffer released.
// Barriers beyond the pair kernel's: c_ready (leader, 8 epilogue warps), l3_full (both CTAs, multicast commit),
// w3_full (leader, TMA bytes of both CTAs' W3 rows).
__device__ __forceinline__ void mbar_wait_bounded_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}

template <int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 128)
    mlp_l2_head_fused_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                             const __grid_constant__ CUtensorMap map_w3, __nv_bfloat16* __restrict__ out, int ldo,
                             int K, int tiles_m, int tiles_n, int arrive_mode) {
  constexpr int BLOCK_N = 256;
  constexpr int kABytes = kMlpBlockM * kMlpBlockK * 2;      // 16 KiB: this CTA's 128 rows of A
  constexpr int kBBytes = 128 * kMlpBlockK * 2;             // 16 KiB: this CTA's half of the W2 tile
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kCBytes = kMlpBlockM * BLOCK_N * 2;         // 64 KiB: relu(h2) tile of this CTA, 4 boxes of 128 x 64
  constexpr int kBoxBytes = kMlpBlockM * 64 * 2;
  constexpr int kW3Box = 32 * 64 * 2;                       // 4 KiB: this CTA's 32 W3 rows x 64 K columns
  constexpr int kW3Bytes = 4 * kW3Box;                      // 16 KiB: 256 K columns
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ctile = smem + STAGES * kStageBytes;
  uint8_t* w3s = ctile + kCBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(w3s + kW3Bytes);    // leader only
  uint64_t* empty = full + STAGES;                                  // per CTA
  uint64_t* tmem_full = empty + STAGES;                             // [2], per CTA
  uint64_t* tmem_empty = tmem_full + 2;                             // [2], leader only
  uint64_t* c_ready = tmem_empty + 2;                               // leader only
  uint64_t* l3_full = c_ready + 1;                                  // per CTA
  uint64_t* w3_full = l3_full + 1;                                  // leader only
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(w3_full + 1);

  // warp index through a shuffle: provably warp-uniform for the compiler (uniform branches, uniform-register operands)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int num_kb = K / kMlpBlockK;
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int n_units = pair < tiles_m ? (tiles_m - pair + num_pairs - 1) / num_pairs : 0;   // 256-row blocks of this pair
  const int T = n_units * tiles_n;                                                          // tiles, N fastest

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
    prefetch_tensormap(&map_w3);
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tmem_full[0], 1);
    mbar_init(&tmem_full[1], 1);
    mbar_init(&tmem_empty[0], 8);      // 4 epilogue warps x 2 CTAs
    mbar_init(&tmem_empty[1], 8);
    mbar_init(c_ready, 8);
    mbar_init(l3_full, 1);
    mbar_init(w3_full, 1);
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 1) tmem_alloc_2sm(tmem_holder, 2 * BLOCK_N);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    {                                 // whole warp walks the loop, one elected lane issues (see elect_one_sync)
      const uint32_t leader_w3 = smem_u32(w3_full) & 0xFEFFFFFFu;
      auto load_w3 = [&](int t) {      // this CTA's 32 rows of W3[:, tn*256 : tn*256+256]
        const int tn = t % tiles_n;
        if (elect_one_sync()) {
          if (leader) mbar_expect_tx(w3_full, 2 * kW3Bytes);
#pragma unroll
          for (int b = 0; b < 4; ++b)
            tma_load_2d_2sm(w3s + b * kW3Box, &map_w3, tn * BLOCK_N + 64 * b, (int)rank * 32, leader_w3);
        }
        __syncwarp();
      };
      int it = 0;
      for (int t = 0; t < T; ++t) {
        const int unit = pair + (t / tiles_n) * num_pairs;
        const int m0 = unit * 256 + (int)rank * 128;
        const int n0 = (t % tiles_n) * BLOCK_N + (int)rank * 128;
        if (t == 0) load_w3(0);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait_bounded(&empty[s], ((it / STAGES) & 1) ^ 1);
          uint8_t* a_dst = smem + (size_t)s * kStageBytes;
          const uint32_t leader_full_tma = smem_u32(&full[s]) & 0xFEFFFFFFu;
          if (elect_one_sync()) {
            if (leader) mbar_expect_tx(&full[s], 2 * kStageBytes);
            tma_load_2d_2sm(a_dst, &map_a, kb * kMlpBlockK, m0, leader_full_tma);
            tma_load_2d_2sm(a_dst + kABytes, &map_b, kb * kMlpBlockK, n0, leader_full_tma);
          }
          __syncwarp();
        }
        if (t >= 1) {
          // the W3 buffer is free once the head MMA of tile t-1 has completed; that MMA is issued in the middle of
          // main(t), whose loads are all in flight by now, so this wait cannot starve the pipeline
          mbar_wait_bounded(l3_full, (uint32_t)((t - 1) & 1));
          load_w3(t);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (leader) {                     // whole warp, one elected lane issues (see elect_one_sync)
      constexpr uint32_t idesc = make_idesc_bf16(256, BLOCK_N);
      constexpr uint32_t idesc3 = make_idesc_bf16(256, 64);
      auto issue_head = [&](int j) {   // logits partial of tile j into the first 64 columns of its drained buffer
        mbar_wait_bounded_cluster(c_ready, (uint32_t)(j & 1));
        mbar_wait_bounded(w3_full, (uint32_t)(j & 1));
        tc_fence_after();
        const uint32_t d3 = tmem_base + (uint32_t)((j & 1) * BLOCK_N);
        if (elect_one_sync()) {
#pragma unroll
          for (int k = 0; k < BLOCK_N / kMlpUmmaK; ++k) {
            const uint64_t adesc = make_smem_desc_sw128(ctile + (k >> 2) * kBoxBytes) + (uint64_t)(2 * (k & 3));
            const uint64_t bdesc = make_smem_desc_sw128(w3s + (k >> 2) * kW3Box) + (uint64_t)(2 * (k & 3));
            umma_f16_2sm(d3, adesc, bdesc, idesc3, (uint32_t)(k != 0));
          }
          umma_commit_2sm(l3_full);
        }
        __syncwarp();
      };
      int it = 0;
      for (int t = 0; t < T; ++t) {
        const int as = t & 1;
        mbar_wait_bounded(&tmem_empty[as], ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          if (t >= 1 && kb == num_kb / 2) issue_head(t - 1);
          const int s = it % STAGES;
          mbar_wait_bounded(&full[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint8_t* a_src = smem + (size_t)s * kStageBytes;
          const uint64_t adesc = make_smem_desc_sw128(a_src);
          const uint64_t bdesc = make_smem_desc_sw128(a_src + kABytes);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < kMlpBlockK / kMlpUmmaK; ++k)
              umma_f16_2sm(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
            umma_commit_2sm(&empty[s]);
            if (kb == num_kb - 1) umma_commit_2sm(&tmem_full[as]);
          }
          __syncwarp();
        }
      }
      if (T >= 1) issue_head(T - 1);
    }
  } else {
    // ===== epilogue (both CTAs) =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    float acc3[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc3[i] = 0.f;
    const uint32_t leader_c_ready = map_to_cta(smem_u32(c_ready), 0);
    const uint32_t ctile_s = smem_u32(ctile);
    for (int t = 0; t < T; ++t) {
      const int as = t & 1;
      const int tn = t % tiles_n;
      // ---- part 1: relu(accumulator) -> bf16 -> swizzled C tile (the head's A operand) ----
      mbar_wait_bounded(&tmem_full[as], (t >> 1) & 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BLOCK_N);
      auto convert = [&](const uint32_t (&acc)[32], int c) {
        const uint32_t box = ctile_s + (uint32_t)((c >> 6) * kBoxBytes + row * 128);
        const int chunk0 = (c & 63) >> 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            pk[j] = pack_bf16x2_relu<true>(__uint_as_float(acc[8 * q + 2 * j]), __uint_as_float(acc[8 * q + 2 * j + 1]));
          }
          const int phys = (chunk0 + q) ^ (row & 7);
          st_shared_v4(box + (uint32_t)(phys * 16), pk[0], pk[1], pk[2], pk[3]);
        }
      };
      if (arrive_mode & 4) {
        // software-pipelined TMEM loads: chunk j+1 is in flight while chunk j is converted (same bytes, same rounding)
        uint32_t acc0[32], acc1[32];
        tmem_ld_32x32b_x32_nowait(tacc, acc0);
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 64) {
          tmem_wait_ld();
          tmem_ld_32x32b_x32_nowait(tacc + (uint32_t)(c + 32), acc1);
          convert(acc0, c);
          tmem_wait_ld();
          if (c + 64 < BLOCK_N) tmem_ld_32x32b_x32_nowait(tacc + (uint32_t)(c + 64), acc0);
          convert(acc1, c + 32);
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
          uint32_t acc[32];
          tmem_ld_32x32b_x32(tacc + (uint32_t)c, acc);
          convert(acc, c);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();        // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      // mode 2 also drops the cluster-scope release here: the C tile is read by THIS CTA's tensor core (async proxy,
      // ordered by the proxy fence above), the leader only needs to learn that it may issue
      if (lane == 0) mbar_arrive_remote(leader_c_ready, (arrive_mode & 3) >= 2);
      // ---- part 2: head partial (64 columns) -> fp32 registers; the accumulator buffer is free afterwards ----
      mbar_wait_bounded(l3_full, (uint32_t)(t & 1));
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t part[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BLOCK_N + 32 * h), part);
#pragma unroll
        for (int i = 0; i < 32; ++i) acc3[32 * h + i] += __uint_as_float(part[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(map_to_cta(smem_u32(&tmem_empty[as]), 0), (arrive_mode & 3) >= 1);
      if (tn == tiles_n - 1) {
        const int unit = pair + (t / tiles_n) * num_pairs;
        const size_t grow = (size_t)unit * 256 + (size_t)rank * 128 + (size_t)row;
        uint4* dst = reinterpret_cast<uint4*>(out + grow * (size_t)ldo);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 v = __floats2bfloat162_rn(acc3[8 * q + 2 * j], acc3[8 * q + 2 * j + 1]);
            pk[j] = *reinterpret_cast<uint32_t*>(&v);
          }
          dst[q] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
#pragma unroll
        for (int i = 0; i < 64; ++i) acc3[i] = 0.f;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * BLOCK_N);
  }
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*PFN_tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                             const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                             CUtensorMapFloatOOBfill);
static PFN_tensorMapEncodeTiled g_encode = nullptr;
static std::once_flag g_encode_once;

static int get_encoder() {
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<PFN_tensorMapEncodeTiled>(fn);
  });
  KTB_REQUIRE(g_encode != nullptr, KTB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  return KTB_OK;
}

// Row-major bf16 [rows, cols] matrix, box = [box_rows, 64 cols], 128-byte swizzle.
static int encode_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);

// The descriptor depends only on (base, rows, cols, box_rows): weights, scratch and staging repeat every chunk and
// every call, so each host thread keeps a small cache (the per-rank issue loop is host-bound at 8 GPUs).
static int make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  struct Key {
    const void* base;
    uint64_t rows, cols;
    uint32_t box;
    bool operator==(const Key& o) const { return base == o.base && rows == o.rows && cols == o.cols && box == o.box; }
  };
  struct Slot {
    Key key;
    CUtensorMap map;
    bool used = false;
  };
  constexpr int kSlots = 64;
  thread_local Slot cache[kSlots];
  const Key k{base, rows, cols, box_rows};
  const size_t h = ((uintptr_t)base >> 8) * 0x9E3779B97F4A7C15ull + rows * 31 + cols * 7 + box_rows;
  Slot& s = cache[h % kSlots];
  if (s.used && s.key == k) {
    *map = s.map;
    return KTB_OK;
  }
  int rc = encode_map(map, base, rows, cols, box_rows);
  if (rc) return rc;
  s.key = k;
  s.map = *map;
  s.used = true;
  return KTB_OK;
}

static int encode_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)kMlpBlockK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KTB_REQUIRE(r == CUDA_SUCCESS, KTB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return KTB_OK;
}

// cudaFuncSetAttribute once per (kernel, device) instead of on every launch.
template <typename KernelT>
static int ensure_smem_attr(KernelT kfn, int smem_bytes, std::atomic<unsigned>& done_mask, int dev) {
  if (dev >= 0 && dev < 32 && (done_mask.load(std::memory_order_acquire) & (1u << dev))) return KTB_OK;
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(max dynamic shared memory = %d) failed: %s", smem_bytes, cudaGetErrorString(e));
    return KTB_ERR_CUDA;
  }
  if (dev >= 0 && dev < 32) done_mask.fetch_or(1u << dev, std::memory_order_release);
  return KTB_OK;
}

template <int BLOCK_N, int STAGES, bool RELU>
static int launch_gemm(int dev, const void* A, const void* B, void* C, size_t M, int N, int K, int ldc,
                       cudaStream_t stream) {
  using S = MlpSmem<BLOCK_N, STAGES>;
  CUtensorMap ma, mb;
  int rc = make_map(&ma, A, M, (uint64_t)K, kMlpBlockM);
  if (rc) return rc;
  rc = make_map(&mb, B, (uint64_t)N, (uint64_t)K, BLOCK_N);
  if (rc) return rc;
  if (g_mlp_persistent && g_mlp_2sm && BLOCK_N == 256 && ldc == N && M % 256 == 0) {
    constexpr int ST = 4;
    constexpr int smem_bytes = ST * 32768 + kMlpBlockM * 256 * 2 + (2 * ST + 4) * 8 + 16 + 1024;
    CUtensorMap mb2, mc;
    rc = make_map(&mb2, B, (uint64_t)N, (uint64_t)K, 128);     // each CTA of the pair loads half of the B tile
    if (rc) return rc;
    rc = make_map(&mc, C, M, (uint64_t)N, kMlpBlockM);
    if (rc) return rc;
    const int sms = device_info(dev) ? device_info(dev)->sm_count : 148;
    if (g_mlp_cluster4 && M % 512 == 0) {
      CUtensorMap mb4;
      rc = make_map(&mb4, B, (uint64_t)N, (uint64_t)K, 64);    // each CTA loads a quarter of the B tile and multicasts it
      if (rc) return rc;
      auto kfn = gemm_bf16_tn_2sm_mc_kernel<ST, RELU, 1>;
      static std::atomic<unsigned> attr_done{0};
      rc = ensure_smem_attr(kfn, smem_bytes, attr_done, dev);
      if (rc) return rc;
      const int tiles_m4 = (int)(M / 512), tiles_n4 = N / 256;
      std::atomic<int>* max_clusters = g_mlp_cluster4_max;
      int mc4 = (dev >= 0 && dev < kMaxDevices) ? max_clusters[dev].load(std::memory_order_relaxed) : 0;
      if (mc4 <= 0) {
        // clusters of 4 cannot always cover all 148 SMs (GPC boundaries): ask the occupancy calculator once
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(sms & ~3));
        cfg.blockDim = dim3(64 + 128);
        cfg.dynamicSmemBytes = smem_bytes;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 4;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int n_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&n_clusters, kfn, &cfg) != cudaSuccess || n_clusters <= 0) {
          (void)cudaGetLastError();
          n_clusters = sms / 4;
        }
        mc4 = n_clusters;
        if (dev >= 0 && dev < kMaxDevices) max_clusters[dev].store(mc4, std::memory_order_relaxed);
      }
      const int grid4 = 4 * std::max(1, std::min(tiles_m4 * tiles_n4, mc4));
      kfn<<<grid4, 64 + 128, smem_bytes, stream>>>(ma, mb4, mc, K, tiles_m4, tiles_n4);
      KTB_CK(cudaGetLastError());
      return KTB_OK;
    }
    const int tiles_m = (int)(M / 256), tiles_n = N / 256;
    const int grid = std::max(2, std::min(2 * tiles_m * tiles_n, sms & ~1));
    if (K == 256 && g_mlp_l1_bres && g_mlp_epi_groups != 2) {
      // layer 1 (K = 256): W1 slice resident, A ring 1.5 tiles deep (see gemm_bf16_tn_2sm_bres_kernel)
      constexpr int AST = 6;
      constexpr int smem_l1 = AST * 16384 + 4 * 16384 + kMlpBlockM * 256 * 2 + (2 * AST + 5) * 8 + 16 + 1024;
      static_assert(smem_l1 <= 232448, "the B-resident layer-1 kernel must fit the opt-in shared memory limit");
      CUtensorMap mc32;
      rc = make_map(&mc32, C, M, (uint64_t)N, 32);      // per-warp stores: 64 columns x 32 rows
      if (rc) return rc;
      if (g_mlp_l1_bres == 4) {
        constexpr int AST8 = 8;          // 2 tiles deep; C staging 32 KiB
        constexpr int smem_l1q = AST8 * 16384 + 4 * 16384 + kMlpBlockM * 256 + (2 * AST8 + 5) * 8 + 16 + 1024;
        static_assert(smem_l1q <= 232448, "the 8-stage layer-1 kernel must fit the opt-in shared memory limit");
        auto kfn = gemm_bf16_tn_2sm_bres_kernel<AST8, RELU, 2>;
        static std::atomic<unsigned> attr_done{0};
        rc = ensure_smem_attr(kfn, smem_l1q, attr_done, dev);
        if (rc) return rc;
        kfn<<<grid, 64 + 128, smem_l1q, stream>>>(ma, mb2, mc, mc32, tiles_m, tiles_n, g_mlp_arrive_mode | g_mlp_debug_flags, g_mlp_trace);
      } else if (g_mlp_l1_bres == 3) {
        auto kfn = gemm_bf16_tn_2sm_bres_kernel<AST, RELU, 1, 8>;
        static std::atomic<unsigned> attr_done{0};
        rc = ensure_smem_attr(kfn, smem_l1, attr_done, dev);
        if (rc) return rc;
        kfn<<<grid, 64 + 256, smem_l1, stream>>>(ma, mb2, mc, mc32, tiles_m, tiles_n, g_mlp_arrive_mode | g_mlp_debug_flags, g_mlp_trace);
      } else if (g_mlp_l1_bres == 2) {
        auto kfn = gemm_bf16_tn_2sm_bres_kernel<AST, RELU, 1>;
        static std::atomic<unsigned> attr_done{0};
        rc = ensure_smem_attr(kfn, smem_l1, attr_done, dev);
        if (rc) return rc;
        kfn<<<grid, 64 + 128, smem_l1, stream>>>(ma, mb2, mc, mc32, tiles_m, tiles_n, g_mlp_arrive_mode | g_mlp_debug_flags, g_mlp_trace);
      } else {
        auto kfn = gemm_bf16_tn_2sm_bres_kernel<AST, RELU, 0>;
        static std::atomic<unsigned> attr_done{0};
        rc = ensure_smem_attr(kfn, smem_l1, attr_done, dev);
        if (rc) return rc;
        kfn<<<grid, 64 + 128, smem_l1, stream>>>(ma, mb2, mc, mc32, tiles_m, tiles_n, g_mlp_arrive_mode | g_mlp_debug_flags, g_mlp_trace);
      }
      KTB_CK(cudaGetLastError());
      return KTB_OK;
    }
    if (g_mlp_epi_groups == 2) {
      auto kfn = gemm_bf16_tn_2sm_kernel<ST, RELU, 2>;
      {
      static std::atomic<unsigned> attr_done{0};
      rc = ensure_smem_attr(kfn, smem_bytes, attr_done, dev);
      if (rc) return rc;
    }
      kfn<<<grid, 64 + 256, smem_bytes, stream>>>(ma, mb2, mc, K, tiles_m, tiles_n);
    } else {
      if (g_mlp_stages == 5) {
        // five 32 KiB stages + the 64 KiB C tile = 225 KiB of the 227 KiB a CTA may own
        constexpr int ST5 = 5;
        constexpr int smem5 = ST5 * 32768 + kMlpBlockM * 256 * 2 + (2 * ST5 + 4) * 8 + 16 + 1024;
        static_assert(smem5 <= 232448, "5-stage ring must fit the opt-in shared memory limit");
        auto kfn = gemm_bf16_tn_2sm_kernel<ST5, RELU, 1>;
        static std::atomic<unsigned> attr_done{0};
        rc = ensure_smem_attr(kfn, smem5, attr_done, dev);
        if (rc) return rc;
        kfn<<<grid, 64 + 128, smem5, stream>>>(ma, mb2, mc, K, tiles_m, tiles_n);
      } else {
        auto kfn = gemm_bf16_tn_2sm_kernel<ST, RELU, 1>;
        static std::atomic<unsigned> attr_done{0};
        rc = ensure_smem_attr(kfn, smem_bytes, attr_done, dev);
        if (rc) return rc;
        kfn<<<grid, 64 + 128, smem_bytes, stream>>>(ma, mb2, mc, K, tiles_m, tiles_n);
      }
    }
  } else if (g_mlp_persistent && g_mlp_tma_store && BLOCK_N == 256 && ldc == N) {
    constexpr int ST = 3;
    using S3 = MlpSmem<256, ST>;
    constexpr int smem_bytes = S3::kBarrierOff + kMlpBlockM * 256 * 2 + (2 * ST + 4) * 8 + 16 + 1024;
    CUtensorMap mc;
    rc = make_map(&mc, C, M, (uint64_t)N, kMlpBlockM);
    if (rc) return rc;
    auto kfn = gemm_bf16_tn_tmastore_kernel<ST, RELU>;
    {
      static std::atomic<unsigned> attr_done{0};
      rc = ensure_smem_attr(kfn, smem_bytes, attr_done, dev);
      if (rc) return rc;
    }
    const int tiles_m = (int)(M / kMlpBlockM), tiles_n = N / 256;
    const int sms = device_info(dev) ? device_info(dev)->sm_count : 148;
    const int grid = std::min(tiles_m * tiles_n, sms);
    kfn<<<grid, kMlpThreads, smem_bytes, stream>>>(ma, mb, mc, K, tiles_m, tiles_n);
  } else if (g_mlp_persistent) {
    const int tiles_m = (int)(M / kMlpBlockM), tiles_n = N / BLOCK_N;
    const int sms = device_info(dev) ? device_info(dev)->sm_count : 148;
    const int grid = std::min(tiles_m * tiles_n, sms);
    if (g_mlp_epi_groups == 2) {
      auto kfn = gemm_bf16_tn_persistent_kernel<BLOCK_N, STAGES, RELU, 2>;
      {
      static std::atomic<unsigned> attr_done{0};
      rc = ensure_smem_attr(kfn, S::kTotal, attr_done, dev);
      if (rc) return rc;
    }
      kfn<<<grid, 64 + 256, S::kTotal, stream>>>(ma, mb, static_cast<__nv_bfloat16*>(C), ldc, K, tiles_m, tiles_n);
    } else {
      auto kfn = gemm_bf16_tn_persistent_kernel<BLOCK_N, STAGES, RELU, 1>;
      {
      static std::atomic<unsigned> attr_done{0};
      rc = ensure_smem_attr(kfn, S::kTotal, attr_done, dev);
      if (rc) return rc;
    }
      kfn<<<grid, 64 + 128, S::kTotal, stream>>>(ma, mb, static_cast<__nv_bfloat16*>(C), ldc, K, tiles_m, tiles_n);
    }
  } else {
    auto kfn = gemm_bf16_tn_kernel<BLOCK_N, STAGES, RELU>;
    {
      static std::atomic<unsigned> attr_done{0};
      rc = ensure_smem_attr(kfn, S::kTotal, attr_done, dev);
      if (rc) return rc;
    }
    dim3 grid((unsigned)(M / kMlpBlockM), (unsigned)(N / BLOCK_N));
    kfn<<<grid, kMlpThreads, S::kTotal, stream>>>(ma, mb, static_cast<__nv_bfloat16*>(C), ldc, K);
  }
  KTB_CK(cudaGetLastError());
  return KTB_OK;
}


// Layer 2 + head fused (see mlp_l2_head_fused_kernel). h1 [M, d_hidden], W2 [d_hidden, d_hidden], W3 [64, d_hidden].
static int launch_l2_head_fused(int dev, const void* h1, const void* W2, const void* W3, void* logits, size_t M,
                                int d_hidden, int d_out, cudaStream_t stream) {
  constexpr int ST = 4;
  constexpr int smem_bytes = ST * 32768 + kMlpBlockM * 256 * 2 + 16384 + (2 * ST + 7) * 8 + 16 + 1024;
  static_assert(smem_bytes <= 232448, "fused kernel must fit the opt-in shared memory limit");
  CUtensorMap ma, mb, mw3;
  int rc = make_map(&ma, h1, M, (uint64_t)d_hidden, kMlpBlockM);
  if (rc) return rc;
  rc = make_map(&mb, W2, (uint64_t)d_hidden, (uint64_t)d_hidden, 128);
  if (rc) return rc;
  rc = make_map(&mw3, W3, (uint64_t)d_out, (uint64_t)d_hidden, 32);
  if (rc) return rc;
  auto kfn = mlp_l2_head_fused_kernel<ST>;
  static std::atomic<unsigned> attr_done{0};
  rc = ensure_smem_attr(kfn, smem_bytes, attr_done, dev);
  if (rc) return rc;
  const int tiles_m = (int)(M / 256), tiles_n = d_hidden / 256;
  const int sms = device_info(dev) ? device_info(dev)->sm_count : 148;
  const int grid = 2 * std::max(1, std::min(tiles_m, sms / 2));
  kfn<<<grid, 64 + 128, smem_bytes, stream>>>(ma, mb, mw3, static_cast<__nv_bfloat16*>(logits), d_out, d_hidden, tiles_m,
                                              tiles_n, g_mlp_arrive_mode);
  KTB_CK(cudaGetLastError());
  return KTB_OK;
}
}  // namespace ktb

using namespace ktb;

extern "C" {

// Developer hook (not part of include/ktb200.h): key 0 = trace buffer of the layer-1 kernel, NULL switches it off.
int ktb_debug_set_ptr(int key, void* p) {
  if (key == 0) { g_mlp_trace = static_cast<unsigned long long*>(p); return KTB_OK; }
  if (key == 1) { g_mlp_debug_flags = (int)(reinterpret_cast<uintptr_t>(p) & 8); return KTB_OK; }
  return KTB_ERR_ARG;
}

size_t ktb_mlp_scratch_bytes(size_t M, int d_hidden) {
  const size_t rows = std::min<size_t>(M, (size_t)g_mlp_chunk_rows);
  return 2 * rows * (size_t)d_hidden * 2;
}

size_t ktb_mlp_stage_bytes(size_t M, int d_in) {
  const size_t rows = std::min<size_t>(M, (size_t)g_mlp_chunk_rows);
  return 2 * rows * (size_t)d_in * 2;
}

// ktb_set_tuning(22, v): staged pulls by a pull kernel (0, default) or by copy engine (1).  Measured at 8 GPUs
// (profiles/r2j_c4_pull_ce_8gpu.log): copy-engine pulls 3.13 ms per call against 2.38 for the pull kernel
int g_mlp_stage_ce = 0;

static int mlp_run(int dev, const void* obs, size_t M, int d_in, int d_hidden, int d_out, const void* W1,
                   const void* W2, const void* W3, void* logits, void* scratch, void* stage, uintptr_t stream) {
  int rc = require_device(dev);
  if (rc) return rc;
  if (M == 0) return KTB_OK;
  KTB_REQUIRE(obs && W1 && W2 && W3 && logits && scratch, KTB_ERR_ARG, "ktb_mlp_bf16: null argument");
  KTB_REQUIRE(M % kMlpBlockM == 0, KTB_ERR_ARG, "ktb_mlp_bf16: M=%zu must be a multiple of %d", M, kMlpBlockM);
  KTB_REQUIRE(d_in > 0 && d_in % kMlpBlockK == 0, KTB_ERR_ARG, "ktb_mlp_bf16: d_in=%d must be a multiple of 64", d_in);
  KTB_REQUIRE(d_hidden > 0 && d_hidden % 256 == 0, KTB_ERR_ARG, "ktb_mlp_bf16: d_hidden=%d must be a multiple of 256",
              d_hidden);
  KTB_REQUIRE(d_out == 64, KTB_ERR_UNSUPPORTED, "ktb_mlp_bf16: d_out=%d (this build carries the 64-wide head)", d_out);
  KTB_REQUIRE((((uintptr_t)obs | (uintptr_t)W1 | (uintptr_t)W2 | (uintptr_t)W3 | (uintptr_t)logits |
                (uintptr_t)scratch | (uintptr_t)stage) & 15) == 0,
              KTB_ERR_ARG, "ktb_mlp_bf16: all pointers must be 16-byte aligned");
  KTB_REQUIRE(g_mlp_chunk_rows % kMlpBlockM == 0 && g_mlp_chunk_rows > 0, KTB_ERR_ARG, "ktb_mlp_bf16: bad chunk rows");
  rc = get_encoder();
  if (rc) return rc;
  KTB_GUARD(dev);
  DeviceInfo* di = device_info(dev);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaStream_t side = di->stream_exec;   // pulls the next chunk while this one computes
  // staged (NVLink pull) calls use smaller chunks: the first pull is exposed and more chunks overlap better
  const size_t chunk = std::min<size_t>(M, stage ? std::min<size_t>((size_t)g_mlp_chunk_rows, 37888) : (size_t)g_mlp_chunk_rows);
  __nv_bfloat16* h1 = static_cast<__nv_bfloat16*>(scratch);
  __nv_bfloat16* h2 = h1 + chunk * (size_t)d_hidden;
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(obs);
  __nv_bfloat16* y = static_cast<__nv_bfloat16*>(logits);
  __nv_bfloat16* stg = static_cast<__nv_bfloat16*>(stage);
  const MapParams ident = make_params(1, 0);
  // per-call events: [0] start, [1..2] staged chunk landed (by buffer), [3..4] staging buffer consumed
  cudaEvent_t evs[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  struct EvGuard {
    cudaEvent_t* e;
    ~EvGuard() {
      for (int i = 0; i < 5; ++i)
        if (e[i]) cudaEventDestroy(e[i]);
    }
  } ev_guard{evs};
  static std::mutex side_mu[kMaxDevices];   // the side stream of a device carries one staged call at a time
  std::unique_lock<std::mutex> side_lock;
  if (stg) {
    side_lock = std::unique_lock<std::mutex>(side_mu[dev]);
    for (int i = 0; i < 5; ++i) KTB_CK(cudaEventCreateWithFlags(&evs[i], cudaEventDisableTiming));
    KTB_CK(cudaEventRecord(evs[0], st));            // obs is ready once prior work on `st` is done
    KTB_CK(cudaStreamWaitEvent(side, evs[0], 0));
  }
  size_t c = 0;
  for (size_t r0 = 0; r0 < M; r0 += chunk, ++c) {
    const size_t rows = std::min(chunk, M - r0);
    const __nv_bfloat16* a1 = x + r0 * d_in;
    if (stg) {
      // staged pull: rows of this chunk travel peer → local ONCE (the layer-1 GEMM would otherwise fetch every
      // A tile d_hidden/256 times over NVLink, peer reads being uncached in the local L2)
      const int b = (int)(c & 1);
      __nv_bfloat16* dstb = stg + (size_t)b * chunk * d_in;
      if (c >= 2) KTB_CK(cudaStreamWaitEvent(side, evs[3 + b], 0));   // GEMM 1 of chunk c-2 consumed it
      if (g_mlp_stage_ce) {
        // copy engine pull: no SM of this rank moves observation rows (the persistent GEMM CTAs own the SMs, a pull
        // KERNEL only gets the gaps between them)
        KTB_CK(cudaMemcpyAsync(dstb, a1, rows * (size_t)d_in * 2, cudaMemcpyDefault, side));
      } else {
        rc = launch_map(dev, KTB_OP_IDENTITY, KTB_U8, a1, dstb, rows * (size_t)d_in * 2, ident, KTB_VARIANT_AUTO, side);
        if (rc) return rc;
      }
      KTB_CK(cudaEventRecord(evs[1 + b], side));
      KTB_CK(cudaStreamWaitEvent(st, evs[1 + b], 0));
      a1 = dstb;
    }
    rc = launch_gemm<256, 4, true>(dev, a1, W1, h1, rows, d_hidden, d_in, d_hidden, st);
    if (rc) return rc;
    if (stg) KTB_CK(cudaEventRecord(evs[3 + (int)(c & 1)], st));
    if (g_mlp_fuse_head && g_mlp_persistent && g_mlp_2sm && rows % 256 == 0 && d_out == 64) {
      rc = launch_l2_head_fused(dev, h1, W2, W3, y + r0 * d_out, rows, d_hidden, d_out, st);
      if (rc) return rc;
      continue;
    }
    rc = launch_gemm<256, 4, true>(dev, h1, W2, h2, rows, d_hidden, d_hidden, d_hidden, st);
    if (rc) return rc;
    rc = launch_gemm<64, 4, false>(dev, h2, W3, y + r0 * d_out, rows, d_out, d_hidden, d_out, st);
    if (rc) return rc;
  }
  return KTB_OK;
}

// ---- pushed form: the root PUSHES this rank's observation rows chunk by chunk (ktb_push_scatter_chunked) ------------
namespace ktb {
// Stream-ordered "chunk c has landed": one warp spins on the local ready flag (a timeout raises the sticky status).
__global__ void mlp_wait_ready_kernel(const unsigned long long* ready, unsigned long long seq, unsigned int* status) {
  if (threadIdx.x == 0) (void)spin_until(ready, seq, status);
}
// Stream-ordered completion of the rank's call: everything before it on the stream (incl. the peer-stored logits)
// is visible system-wide before the root sees ack[rank] = seq.
__global__ void mlp_ack_kernel(unsigned long long* ack, unsigned long long seq) {
  if (threadIdx.x == 0) {
    __threadfence_system();
    st_release_sys(ack, seq);
  }
}
}  // namespace ktb

int ktb_mlp_bf16_pushed(int dev, const void* stage_local, size_t stage_stride, size_t M, int d_in, int d_hidden, int d_out,
                        const void* W1, const void* W2, const void* W3, void* logits, void* scratch, void* ctrl_local,
                        void* ctrl_root_peer, int rank, size_t chunk_rows, unsigned long long seq, uintptr_t stream) {
  int rc = require_device(dev);
  if (rc) return rc;
  KTB_REQUIRE(stage_local && ctrl_local && ctrl_root_peer && W1 && W2 && W3 && scratch && (logits || M == 0), KTB_ERR_ARG,
              "ktb_mlp_bf16_pushed: null argument");
  KTB_REQUIRE(rank >= 0 && rank < 16 && seq > 0, KTB_ERR_ARG, "ktb_mlp_bf16_pushed: bad rank/seq");
  KTB_REQUIRE(M % kMlpBlockM == 0, KTB_ERR_ARG, "ktb_mlp_bf16_pushed: M=%zu must be a multiple of %d", M, kMlpBlockM);
  KTB_REQUIRE(chunk_rows > 0 && chunk_rows % kMlpBlockM == 0, KTB_ERR_ARG,
              "ktb_mlp_bf16_pushed: chunk_rows=%zu must be a positive multiple of %d", chunk_rows, kMlpBlockM);
  KTB_REQUIRE(d_in > 0 && d_in % kMlpBlockK == 0 && d_hidden > 0 && d_hidden % 256 == 0, KTB_ERR_ARG,
              "ktb_mlp_bf16_pushed: d_in %% 64 and d_hidden %% 256 must be 0");
  KTB_REQUIRE(d_out == 64, KTB_ERR_UNSUPPORTED, "ktb_mlp_bf16_pushed: d_out=%d (this build carries the 64-wide head)", d_out);
  const size_t n_chunks = (M + chunk_rows - 1) / chunk_rows;
  KTB_REQUIRE(n_chunks <= KTB_PUSH_MAX_CHUNKS, KTB_ERR_ARG, "ktb_mlp_bf16_pushed: %zu chunks exceed %d", n_chunks,
              KTB_PUSH_MAX_CHUNKS);
  KTB_REQUIRE(M * (size_t)d_in * 2 <= stage_stride, KTB_ERR_ARG, "ktb_mlp_bf16_pushed: shard exceeds stage_stride");
  rc = get_encoder();
  if (rc) return rc;
  KTB_GUARD(dev);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* cl = static_cast<uint8_t*>(ctrl_local);
  const unsigned long long* ready = reinterpret_cast<const unsigned long long*>(cl + KTB_CTRL_READY);
  unsigned int* status = reinterpret_cast<unsigned int*>(cl + KTB_CTRL_STATUS);
  unsigned long long* ack =
      reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(ctrl_root_peer) + KTB_CTRL_ACK) + rank;
  const __nv_bfloat16* x =
      reinterpret_cast<const __nv_bfloat16*>(static_cast<const uint8_t*>(stage_local) + (size_t)(seq & 1) * stage_stride);
  __nv_bfloat16* h1 = static_cast<__nv_bfloat16*>(scratch);
  __nv_bfloat16* h2 = h1 + std::min(chunk_rows, M) * (size_t)d_hidden;
  __nv_bfloat16* y = static_cast<__nv_bfloat16*>(logits);
  size_t c = 0;
  for (size_t r0 = 0; r0 < M; r0 += chunk_rows, ++c) {
    const size_t rows = std::min(chunk_rows, M - r0);
    mlp_wait_ready_kernel<<<1, 32, 0, st>>>(ready + c, seq, status);
    KTB_CK(cudaGetLastError());
    const __nv_bfloat16* a1 = x + r0 * d_in;
    rc = launch_gemm<256, 4, true>(dev, a1, W1, h1, rows, d_hidden, d_in, d_hidden, st);
    if (rc) return rc;
    if (g_mlp_fuse_head && g_mlp_persistent && g_mlp_2sm && rows % 256 == 0) {
      rc = launch_l2_head_fused(dev, h1, W2, W3, y + r0 * d_out, rows, d_hidden, d_out, st);
      if (rc) return rc;
      continue;
    }
    rc = launch_gemm<256, 4, true>(dev, h1, W2, h2, rows, d_hidden, d_hidden, d_hidden, st);
    if (rc) return rc;
    rc = launch_gemm<64, 4, false>(dev, h2, W3, y + r0 * d_out, rows, d_out, d_hidden, d_out, st);
    if (rc) return rc;
  }
  mlp_ack_kernel<<<1, 32, 0, st>>>(ack, seq);
  KTB_CK(cudaGetLastError());
  return KTB_OK;
}

int ktb_mlp_bf16(int dev, const void* obs, size_t M, int d_in, int d_hidden, int d_out, const void* W1,
                 const void* W2, const void* W3, void* logits, void* scratch, uintptr_t stream) {
  return mlp_run(dev, obs, M, d_in, d_hidden, d_out, W1, W2, W3, logits, scratch, nullptr, stream);
}

int ktb_mlp_bf16_staged(int dev, const void* obs_peer, size_t M, int d_in, int d_hidden, int d_out, const void* W1,
                        const void* W2, const void* W3, void* logits, void* scratch, void* stage, uintptr_t stream) {
  KTB_REQUIRE(stage, KTB_ERR_ARG, "ktb_mlp_bf16_staged: null stage buffer");
  return mlp_run(dev, obs_peer, M, d_in, d_hidden, d_out, W1, W2, W3, logits, scratch, stage, stream);
}

}  // extern "C"
