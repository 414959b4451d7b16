// gather_probe — how large can the gathered window of a row-streaming SpMV pass be before the random 8-byte
// gathers start to miss the L2 of a B200, with the matrix entries streaming through the same L2?
//
// Model of one window pass of the banded operator ("random sparse LP"): thread-per-row CSR with k
// entries per row (int32 index + fp64 value, streamed once), a carried fp64 partial sum per row (read + written),
// and k gathers into a window of `window_mb` of a large vector.  Sweeps window size, k, rows per thread and the
// L2 eviction hints.  Standalone: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_probe gather_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

__global__ void k_init_idx(int32_t *idx, int64_t count, int64_t window_elems, int64_t base) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x)
    idx[e] = (int32_t)(base + (int64_t)(mix((uint64_t)e) % (uint64_t)window_elems));
}
__global__ void k_init_f64(double *v, int64_t count, double x) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x) v[e] = x;
}

__device__ __forceinline__ double ld_gather(const double *p, int hint, uint64_t pol) {
  double v;
  if (hint == 0) return __ldg(p);
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
  return v;
}

// K entries per row, R rows per thread (rows of one thread are 32*R apart inside a warp tile so that loads stay coalesced)
template <int K, int R>
__global__ void __launch_bounds__(256) k_pass(const int32_t *__restrict__ idx, const double *__restrict__ val,
                                              const double *__restrict__ vec, double *__restrict__ carry, int64_t rows,
                                              int use_carry, int hint) {
  uint64_t pol = 0;
  if (hint) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t row0 = warp * (32 * R) + lane;
  int32_t j[R][K];
  double a[R][K], acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int64_t row = row0 + 32 * r;
    acc[r] = 0.0;
    if (row < rows) {
      if (use_carry) acc[r] = __ldcs(carry + row);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        j[r][k] = __ldcs(idx + row * K + k);
        a[r][k] = __ldcs(val + row * K + k);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int64_t row = row0 + 32 * r;
    if (row < rows) {
      double g[K];
#pragma unroll
      for (int k = 0; k < K; ++k) g[k] = ld_gather(vec + j[r][k], hint, pol);
#pragma unroll
      for (int k = 0; k < K; ++k) acc[r] = __dadd_rn(acc[r], __dmul_rn(a[r][k], g[k]));
      __stcs(carry + row, acc[r]);
    }
  }
}

template <int K, int R>
float run(const int32_t *idx, const double *val, const double *vec, double *carry, int64_t rows, int use_carry, int hint,
          int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int64_t warps = (rows + 32 * R - 1) / (32 * R);
  const int grid = (int)((warps * 32 + 255) / 256);
  k_pass<K, R><<<grid, 256>>>(idx, val, vec, carry, rows, use_carry, hint);
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) k_pass<K, R><<<grid, 256>>>(idx, val, vec, carry, rows, use_carry, hint);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGetLastError());
  return ms / reps;
}

int main(int argc, char **argv) {
  const int64_t entries = 160ll << 20;              // 160 Mi entries: 640 MiB of indices + 1.25 GiB of values
  const int64_t vec_elems = (512ll << 20) / 8;      // 512 MiB vector
  int32_t *idx;
  double *val, *vec, *carry;
  CK(cudaMalloc(&idx, entries * 4));
  CK(cudaMalloc(&val, entries * 8));
  CK(cudaMalloc(&vec, vec_elems * 8));
  CK(cudaMalloc(&carry, (entries / 2) * 8));
  k_init_f64<<<2048, 256>>>(val, entries, 1.0);
  k_init_f64<<<2048, 256>>>(vec, vec_elems, 0.5);
  k_init_f64<<<2048, 256>>>(carry, entries / 2, 0.0);
  CK(cudaDeviceSynchronize());
  int dev = 0, l2 = 0, persist = 0;
  CK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev));
  CK(cudaDeviceGetAttribute(&persist, cudaDevAttrMaxPersistingL2CacheSize, dev));
  printf("{\"l2_bytes\": %d, \"max_persisting_l2\": %d}\n", l2, persist);
  const int windows_mb[] = {4, 16, 24, 32, 40, 48, 64, 80, 96, 128, 192, 320, 512};
  for (int wm : windows_mb) {
    const int64_t welems = ((int64_t)wm << 20) / 8;
    k_init_idx<<<4096, 256>>>(idx, entries, welems, (vec_elems - welems) / 2);
    CK(cudaDeviceSynchronize());
    for (int hint = 0; hint < 2; ++hint) {
      struct { int K, R; float ms; int64_t rows; } res[6];
      int n = 0;
      res[n] = {2, 1, run<2, 1>(idx, val, vec, carry, entries / 2, 1, hint, 5), entries / 2}; ++n;
      res[n] = {2, 2, run<2, 2>(idx, val, vec, carry, entries / 2, 1, hint, 5), entries / 2}; ++n;
      res[n] = {2, 4, run<2, 4>(idx, val, vec, carry, entries / 2, 1, hint, 5), entries / 2}; ++n;
      res[n] = {4, 2, run<4, 2>(idx, val, vec, carry, entries / 4, 1, hint, 5), entries / 4}; ++n;
      res[n] = {8, 1, run<8, 1>(idx, val, vec, carry, entries / 8, 1, hint, 5), entries / 8}; ++n;
      res[n] = {8, 2, run<8, 2>(idx, val, vec, carry, entries / 8, 1, hint, 5), entries / 8}; ++n;
      for (int i = 0; i < n; ++i) {
        const double streamed = 12.0 * entries + 16.0 * res[i].rows;
        printf("{\"window_mb\": %d, \"evict_last_hint\": %d, \"k\": %d, \"rows_per_thread\": %d, \"ms\": %.4f, "
               "\"streamed_GBs\": %.1f, \"ns_per_1e3_gathers\": %.3f}\n",
               wm, hint, res[i].K, res[i].R, res[i].ms, streamed / res[i].ms / 1e6, res[i].ms * 1e6 / entries * 1e3 / 1e3);
      }
      fflush(stdout);
    }
  }
  return 0;
}
