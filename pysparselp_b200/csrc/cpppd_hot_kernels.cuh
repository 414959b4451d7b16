// The two hot kernels: k_primal (:198-228) and k_dual (:231-240, :333-341).
// Part of libcpppd (single translation unit, included by cpppd.cu).
#pragma once

namespace {

// ------------------------------------------------------------------------------------------
// the two hot kernels
// ------------------------------------------------------------------------------------------
// Primal half-iteration (:198-228).  Thread j owns column j of A (row j of A^T).
// Loads that do not depend on the matrix (c, T, x) are issued first so that they are in flight
// together with the slice entries; matrix entries are read once (ld.global.cs).
// kDict: entries are single 32-bit words [pad][eq][code][index]; values come from a <= 256 entry
// dictionary staged in shared memory.
template <bool kWriteD, bool kDict, int kChunk>
__global__ void __launch_bounds__(kBlock, kMinBlocks)
k_primal(SellView AT, const double *__restrict__ y, Vec c, Vec T, Vec lb, Vec ub, double *__restrict__ x,
         double *__restrict__ xbar, double *__restrict__ d_out, int64_t n, int has_eq, int has_ineq,
         double theta, double one_plus_theta) {
  __shared__ double sdict[kDict ? 256 : 1];
  if (kDict) {
    if ((int)threadIdx.x < AT.ndict) sdict[threadIdx.x] = AT.dict[threadIdx.x];
    __syncthreads();
  }
  const int64_t j = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const int64_t s = j >> 5;
  if (s >= AT.nslices) return;
  const int lane = threadIdx.x & 31;
  int64_t p0, p1;
  slice_range(AT, s, p0, p1);
  const bool live = j < n;
  double cj = 0.0, tj = 0.0, xo = 0.0;
  if (live) {
    cj = c.at(j);
    tj = T.at(j);
    xo = __ldcs(x + j);
  }
  double s_eq = 0.0, s_in = 0.0;
  {
    const int32_t *ip = AT.idx + p0 + lane;
    const double *vp = AT.val + p0 + lane;
    const int width = (int)((p1 - p0) >> 5);
    const int32_t mask = AT.idx_mask;
    // entries are taken kChunk at a time: all index (and value) loads of a chunk are issued first,
    // then all gathers, then the sequential accumulation — so a row of any width (also 2 or 3) keeps
    // kChunk independent gathers in flight instead of one load-use chain per entry
#pragma unroll 1
    for (int k0 = 0; k0 < width; k0 += kChunk) {
      int32_t r[kChunk];
      double a[kChunk], g[kChunk];
#pragma unroll
      for (int u = 0; u < kChunk; ++u) {
        const bool ok = k0 + u < width;
        r[u] = ok ? __ldcs(ip + (k0 + u) * kSlice) : kPad;
        a[u] = (!kDict && ok) ? __ldcs(vp + (k0 + u) * kSlice) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kChunk; ++u) g[u] = r[u] >= 0 ? __ldg(y + (r[u] & mask)) : 0.0;
#pragma unroll
      for (int u = 0; u < kChunk; ++u) {
        if (r[u] >= 0) {
          const double av = kDict ? sdict[(r[u] >> AT.idx_bits) & AT.code_mask] : a[u];
          const double t = __dmul_rn(av, g[u]);
          if (r[u] & kEqBit) s_eq = __dadd_rn(s_eq, t); else s_in = __dadd_rn(s_in, t);
        }
      }
    }
  }
  if (!live) return;
  double d = cj;
  if (has_eq) d = __dadd_rn(d, s_eq);
  if (has_ineq) d = __dadd_rn(d, s_in);
  const double l = lb.at(j), u = ub.at(j);
  double x2 = __dsub_rn(xo, __dmul_rn(tj, d));
  x2 = (l > x2) ? l : x2;  // np.maximum(x2, lb)  (NaN in x2 propagates)
  x2 = (u < x2) ? u : x2;  // np.minimum(x2, ub)
  xbar[j] = __dsub_rn(__dmul_rn(one_plus_theta, x2), __dmul_rn(theta, xo));
  x[j] = x2;
  if (kWriteD) d_out[j] = d;
}

// Dual half-iteration (:231-240, :333-341).  Thread i owns row i of A.
template <bool kDict, int kChunk>
__global__ void __launch_bounds__(kBlock, kMinBlocks)
k_dual(SellView A, const double *__restrict__ xbar, Vec b, Vec sigma, double *__restrict__ y, int64_t m,
       int64_t m_eq) {
  __shared__ double sdict[kDict ? 256 : 1];
  if (kDict) {
    if ((int)threadIdx.x < A.ndict) sdict[threadIdx.x] = A.dict[threadIdx.x];
    __syncthreads();
  }
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  const int64_t s = i >> 5;
  if (s >= A.nslices) return;
  const int lane = threadIdx.x & 31;
  int64_t p0, p1;
  slice_range(A, s, p0, p1);
  const bool live = i < m;
  double bi = 0.0, si = 0.0, yi = 0.0;
  if (live) {
    bi = b.at(i);
    si = sigma.at(i);
    yi = __ldcs(y + i);
  }
  double acc = 0.0;
  {
    const int32_t *ip = A.idx + p0 + lane;
    const double *vp = A.val + p0 + lane;
    const int width = (int)((p1 - p0) >> 5);
    const int32_t mask = A.idx_mask;
#pragma unroll 1
    for (int k0 = 0; k0 < width; k0 += kChunk) {  // see k_primal: loads of a chunk first, then gathers, then sums
      int32_t jc[kChunk];
      double a[kChunk], g[kChunk];
#pragma unroll
      for (int u = 0; u < kChunk; ++u) {
        const bool ok = k0 + u < width;
        jc[u] = ok ? __ldcs(ip + (k0 + u) * kSlice) : kPad;
        a[u] = (!kDict && ok) ? __ldcs(vp + (k0 + u) * kSlice) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < kChunk; ++u) g[u] = jc[u] >= 0 ? __ldg(xbar + (jc[u] & mask)) : 0.0;
#pragma unroll
      for (int u = 0; u < kChunk; ++u) {
        if (jc[u] >= 0) {
          const double av = kDict ? sdict[(jc[u] >> A.idx_bits) & A.code_mask] : a[u];
          acc = __dadd_rn(acc, __dmul_rn(av, g[u]));
        }
      }
    }
  }
  if (!live) return;
  const double r = __dsub_rn(acc, bi);
  double yn = __dadd_rn(yi, __dmul_rn(si, r));
  if (i >= m_eq) yn = (yn < 0.0) ? 0.0 : yn;  // np.maximum(y_ineq, 0): NaN stays NaN, -0.0 stays
  y[i] = yn;
}

}  // namespace
