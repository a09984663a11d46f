// common.cuh — shared helpers for the sm_100a kernels of librse.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rse {

constexpr int kVec0Block = 1024;          // sqlite-vec default chunk_size
constexpr uint32_t kInvalidOKey = 0xFFFFFFFFu;

// Monotone map float -> uint32 (ascending float == ascending uint).  The KNN
// orders by the f32 distance first (vec0 stores/compares f32 distances).
__host__ __device__ __forceinline__ uint32_t f32_orderable(uint32_t bits) {
  return (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
}
__host__ __device__ __forceinline__ uint32_t f32_from_orderable(uint32_t o) {
  return (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
}
// Monotone map double -> uint64.
__host__ __device__ __forceinline__ uint64_t f64_orderable(uint64_t bits) {
  return (bits & 0x8000000000000000ull) ? ~bits : (bits | 0x8000000000000000ull);
}

// vec0 emit order (distance asc, block asc, slot desc) as ONE ascending 64-bit
// key: high word = orderable distance, low word = global position with the slot
// bits inverted (pos ^ 1023)  — SURVEY App. A.2.
__host__ __device__ __forceinline__ uint64_t knn_key(uint32_t okey, uint64_t global_pos) {
  return (static_cast<uint64_t>(okey) << 32) | static_cast<uint32_t>(global_pos ^ (kVec0Block - 1));
}
__host__ __device__ __forceinline__ uint64_t knn_key_pos(uint64_t key) {
  return static_cast<uint32_t>(key) ^ static_cast<uint32_t>(kVec0Block - 1);
}

// In-shared-memory bitonic sort of (key, payload) pairs, ascending by key.
// n must be a power of two; all threads of the block participate.
template <typename K, typename V>
__device__ __forceinline__ void block_bitonic_sort(K* keys, V* vals, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        int lo = 2 * t - (t & (stride - 1));
        int hi = lo + stride;
        bool up = ((lo & size) == 0);
        K a = keys[lo], b = keys[hi];
        if ((a > b) == up) {
          keys[lo] = b; keys[hi] = a;
          V va = vals[lo]; vals[lo] = vals[hi]; vals[hi] = va;
        }
      }
    }
  }
  __syncthreads();
}

__host__ __device__ __forceinline__ int next_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

}  // namespace rse
