// ktb_reduce.cu — gather-reduce variant of the mapped call: out = sum_i op(x_i).
//
// In the reference each rank returns `op(shard).sum()` and the caller receives a list of
// per-rank scalars (kt/serving/spmd/spmd_supervisor.py:547-570) which user code then sums.
// Here rank r reduces its shard in one launch (per-thread accumulators → warp-shuffle tree →
// shared memory across warps → one partial per CTA → last-arriving CTA folds the partials in
// fixed order) and stores the scalar straight into the root's partials[r] (peer store).
//
// HBM-bound: algorithmic bytes = n * sizeof(elem) (read once; the output is 4–8 bytes).
#include "ktb_common.cuh"

#include <algorithm>

namespace ktb {

constexpr int kRedThreads = 256;
constexpr int kRedMaxGrid = 8192;
constexpr size_t kRedHeader = 64;  // counter lives in the first 64 bytes of the workspace
int g_red_ctas_per_sm = 8;         // ktb_set_tuning key 6

template <int DT>
struct Acc {
  using type = float;
};
template <>
struct Acc<KTB_I32> {
  using type = long long;
};
template <>
struct Acc<KTB_I64> {
  using type = long long;
};

template <int DT, int OP>
__device__ __forceinline__ typename Acc<DT>::type elem_value(const uint8_t* p, const MapParams& mp) {
  if constexpr (DT == KTB_F32) {
    return apply_f32<OP>(*reinterpret_cast<const float*>(p), mp);
  } else if constexpr (DT == KTB_BF16) {
    uint16_t h = *reinterpret_cast<const uint16_t*>(p);
    return apply_bf16_as_f32<OP>(__uint_as_float((uint32_t)h << 16), mp);
  } else if constexpr (DT == KTB_F16) {
    return apply_f16_as_f32<OP>(__half2float(*reinterpret_cast<const __half*>(p)), mp);
  } else if constexpr (DT == KTB_I32) {
    return (long long)(int)apply_i32<OP>(*reinterpret_cast<const uint32_t*>(p), mp);
  } else {
    return (long long)apply_i64<OP>(*reinterpret_cast<const unsigned long long*>(p), mp);
  }
}

// Sum of the op-mapped elements held in NW 32-bit words.
template <int DT, int OP, int NW>
__device__ __forceinline__ typename Acc<DT>::type words_value(const uint32_t (&w)[NW], const MapParams& mp) {
  if constexpr (DT == KTB_F32) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) s += apply_f32<OP>(__uint_as_float(w[i]), mp);
    return s;
  } else if constexpr (DT == KTB_BF16) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      s += apply_bf16_as_f32<OP>(__uint_as_float(w[i] << 16), mp);
      s += apply_bf16_as_f32<OP>(__uint_as_float(w[i] & 0xffff0000u), mp);
    }
    return s;
  } else if constexpr (DT == KTB_F16) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      s += apply_f16_as_f32<OP>(f.x, mp);
      s += apply_f16_as_f32<OP>(f.y, mp);
    }
    return s;
  } else if constexpr (DT == KTB_I32) {
    long long s = 0;
#pragma unroll
    for (int i = 0; i < NW; ++i) s += (long long)(int)apply_i32<OP>(w[i], mp);
    return s;
  } else {
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < NW; i += 2)
      s += apply_i64<OP>(((unsigned long long)w[i + 1] << 32) | w[i], mp);
    return (long long)s;
  }
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem /* >= 32 entries */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  T r = 0;
  if (warp == 0) {
    r = (lane < (int)(blockDim.x >> 5)) ? smem[lane] : (T)0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

template <int DT, int OP>
__global__ void __launch_bounds__(kRedThreads)
    map_reduce_kernel(const uint8_t* src, size_t n_elems, MapParams mp, void* out, uint8_t* ws) {
  using A = typename Acc<DT>::type;
  // cross-CTA partials are kept in fp64 for float sums, int64 for integer sums
  using P = typename std::conditional<std::is_same<A, float>::value, double, long long>::type;
  constexpr size_t ES = (DT == KTB_BF16 || DT == KTB_F16) ? 2 : (DT == KTB_I64 ? 8 : 4);
  __shared__ P red[32];
  __shared__ bool is_last;

  const size_t n_bytes = n_elems * ES;
  const bool vec_ok = (((uintptr_t)src) & 31) == 0;
  const size_t n_vec = vec_ok ? (n_bytes >> 5) : 0;  // 32-byte packets

  A acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  const size_t stride = (size_t)gridDim.x * kRedThreads;
  // CTA-contiguous tiles of 4 x 256 packets (32 KiB): 4 independent 256-bit loads in flight per thread, one tile per
  // CTA when the grid covers the input (the hardware scheduler balances the tail), grid-stride beyond kRedMaxGrid
  constexpr size_t kTilePackets = 4 * (size_t)kRedThreads;
  const size_t n_tiles = n_vec / kTilePackets;
  for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const size_t v = t * kTilePackets + threadIdx.x;
    uint32_t w0[8], w1[8], w2[8], w3[8];
    ldg256_stream(src + (v << 5), w0);
    ldg256_stream(src + ((v + kRedThreads) << 5), w1);
    ldg256_stream(src + ((v + 2 * kRedThreads) << 5), w2);
    ldg256_stream(src + ((v + 3 * kRedThreads) << 5), w3);
    acc0 += words_value<DT, OP, 8>(w0, mp);
    acc1 += words_value<DT, OP, 8>(w1, mp);
    acc2 += words_value<DT, OP, 8>(w2, mp);
    acc3 += words_value<DT, OP, 8>(w3, mp);
  }
  for (size_t v = n_tiles * kTilePackets + (size_t)blockIdx.x * kRedThreads + threadIdx.x; v < n_vec; v += stride) {
    uint32_t w0[8];
    ldg256_stream(src + (v << 5), w0);
    acc0 += words_value<DT, OP, 8>(w0, mp);
  }
  // element tail (everything when the pointer is not 32-byte aligned)
  const size_t tail0 = (n_vec << 5) / ES;
  for (size_t e = tail0 + (size_t)blockIdx.x * kRedThreads + threadIdx.x; e < n_elems; e += stride)
    acc1 += elem_value<DT, OP>(src + e * ES, mp);

  A acc = (acc0 + acc1) + (acc2 + acc3);
  P part = block_sum<P>((P)acc, red);

  unsigned int* counter = reinterpret_cast<unsigned int*>(ws);
  P* partials = reinterpret_cast<P*>(ws + kRedHeader);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = part;
    __threadfence();
    unsigned int ticket = atomicAdd(counter, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    P s = 0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += kRedThreads)
      s += *reinterpret_cast<volatile P*>(&partials[i]);
    s = block_sum<P>(s, red);
    if (threadIdx.x == 0) {
      if constexpr (std::is_same<A, float>::value)
        *reinterpret_cast<float*>(out) = (float)s;
      else
        *reinterpret_cast<long long*>(out) = (long long)s;
      *counter = 0;  // leave the workspace ready for the next call
    }
  }
}

template <typename T, typename P>
__global__ void __launch_bounds__(kRedThreads) reduce_partials_kernel(const T* partials, int n, T* out) {
  __shared__ P red[32];
  P s = 0;
  for (int i = threadIdx.x; i < n; i += kRedThreads) s += (P)partials[i];
  s = block_sum<P>(s, red);
  if (threadIdx.x == 0) out[0] = (T)s;
}

template <int DT, int OP>
static int launch_reduce_typed(int dev, const uint8_t* src, size_t n_elems, const MapParams& p,
                               void* out, void* ws, cudaStream_t stream) {
  const DeviceInfo* di = device_info(dev);
  (void)di;
  constexpr size_t ES = (DT == KTB_BF16 || DT == KTB_F16) ? 2 : (DT == KTB_I64 ? 8 : 4);
  const size_t tiles = (n_elems * ES + 32767) / 32768;   // one 32 KiB tile per CTA
  const size_t cap = g_red_ctas_per_sm > 0 ? std::min<size_t>((size_t)kRedMaxGrid, (size_t)g_red_ctas_per_sm * 1024) : kRedMaxGrid;
  int grid = (int)std::min<size_t>(std::max<size_t>(tiles, 1), cap);
  map_reduce_kernel<DT, OP><<<grid, kRedThreads, 0, stream>>>(src, n_elems, p, out, (uint8_t*)ws);
  KTB_CK(cudaGetLastError());
  return KTB_OK;
}

int launch_map_reduce(int dev, int op, int dtype, const void* src, size_t n_elems, const MapParams& p,
                      void* out, void* workspace, cudaStream_t stream) {
  KTB_REQUIRE(dtype == KTB_F32 || dtype == KTB_BF16 || dtype == KTB_F16 || dtype == KTB_I32 || dtype == KTB_I64, KTB_ERR_ARG,
              "ktb_map_reduce_sum: dtype %d not reducible", dtype);
  KTB_REQUIRE(op >= KTB_OP_IDENTITY && op <= KTB_OP_AFFINE, KTB_ERR_ARG, "ktb_map_reduce_sum: unknown op %d", op);
  KTB_REQUIRE(out && workspace, KTB_ERR_ARG, "ktb_map_reduce_sum: null out/workspace");
  KTB_REQUIRE(src || n_elems == 0, KTB_ERR_ARG, "ktb_map_reduce_sum: null src");
  const size_t es = dtype_size(dtype);
  KTB_REQUIRE((((uintptr_t)src) & (es - 1)) == 0, KTB_ERR_ARG, "ktb_map_reduce_sum: src misaligned");
  const uint8_t* s = static_cast<const uint8_t*>(src);
#define KTB_RCASE(DT)                                                                         \
  case DT:                                                                                    \
    switch (op) {                                                                             \
      case KTB_OP_IDENTITY:                                                                   \
        return launch_reduce_typed<DT, KTB_OP_IDENTITY>(dev, s, n_elems, p, out, workspace, stream); \
      case KTB_OP_SCALE:                                                                      \
        return launch_reduce_typed<DT, KTB_OP_SCALE>(dev, s, n_elems, p, out, workspace, stream);    \
      default:                                                                                \
        return launch_reduce_typed<DT, KTB_OP_AFFINE>(dev, s, n_elems, p, out, workspace, stream);   \
    }
  switch (dtype) {
    KTB_RCASE(KTB_F32)
    KTB_RCASE(KTB_BF16)
    KTB_RCASE(KTB_I32)
    KTB_RCASE(KTB_I64)
    KTB_RCASE(KTB_F16)
  }
#undef KTB_RCASE
  return KTB_ERR_UNSUPPORTED;
}

int launch_reduce_partials(int dev, int dtype, const void* partials, int n, void* out, cudaStream_t stream) {
  (void)dev;
  KTB_REQUIRE(partials && out && n > 0, KTB_ERR_ARG, "ktb_reduce_partials: bad arguments");
  if (dtype == KTB_F32 || dtype == KTB_BF16 || dtype == KTB_F16)
    reduce_partials_kernel<float, double><<<1, kRedThreads, 0, stream>>>((const float*)partials, n, (float*)out);
  else if (dtype == KTB_I32 || dtype == KTB_I64)
    reduce_partials_kernel<long long, long long>
        <<<1, kRedThreads, 0, stream>>>((const long long*)partials, n, (long long*)out);
  else {
    set_error("ktb_reduce_partials: dtype %d not reducible", dtype);
    return KTB_ERR_ARG;
  }
  KTB_CK(cudaGetLastError());
  return KTB_OK;
}

}  // namespace ktb

using namespace ktb;

extern "C" {

size_t ktb_reduce_workspace_bytes(void) { return kRedHeader + sizeof(double) * kRedMaxGrid; }

int ktb_map_reduce_sum(int dev, int op, int dtype, const void* src, size_t n_elems, double alpha,
                       double beta, void* out, void* workspace, uintptr_t stream) {
  int rc = require_device(dev);
  if (rc) return rc;
  KTB_GUARD(dev);
  return launch_map_reduce(dev, op, dtype, src, n_elems, make_params(alpha, beta, dtype), out, workspace,
                           reinterpret_cast<cudaStream_t>(stream));
}

int ktb_reduce_partials(int dev, int dtype, const void* partials, int n, void* out, uintptr_t stream) {
  int rc = require_device(dev);
  if (rc) return rc;
  KTB_GUARD(dev);
  return launch_reduce_partials(dev, dtype, partials, n, out, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
