// TEST INFRASTRUCTURE — CPU driver for the real k_primal / k_dual sources (see cuda_shim.h).
#include "cuda_shim.h"

#include "../../pysparselp_b200/csrc/cpppd_device_types.cuh"
#include "../../pysparselp_b200/csrc/cpppd_hot_kernels.cuh"

namespace {

template <typename F>
void run_grid(int64_t nslices, bool has_barrier, F &&thread_body) {
  const int64_t threads = nslices * 32;
  blockDim.x = kBlock;
  gridDim.x = (unsigned)((threads + kBlock - 1) / kBlock);
  for (unsigned b = 0; b < gridDim.x; ++b) {
    blockIdx.x = b;
    for (int pass = has_barrier ? 0 : 1; pass < 2; ++pass) {
      g_emul_stop_at_barrier = pass == 0;
      for (unsigned t = 0; t < (unsigned)kBlock; ++t) {
        threadIdx.x = t;
        try {
          thread_body();
        } catch (const EmulBarrier &) {
        }
      }
    }
  }
  g_emul_stop_at_barrier = false;
}

SellView make_view(const int64_t *slice_ptr, const int32_t *idx, const double *val, int64_t nrows, int64_t nslices,
                   int64_t uniform_width, const double *dict, int idx_bits, int ndict) {
  const int code_bits = dict ? 30 - idx_bits : 0;
  return SellView{slice_ptr, idx, val, nrows, nslices, uniform_width, dict,
                  dict ? (int32_t)((1u << idx_bits) - 1) : kIdxMask, idx_bits, (int32_t)((1u << code_bits) - 1), ndict};
}

}  // namespace

extern "C" {

struct EmulVec { const double *p; double c; };

void emul_primal(int variant, int write_d, const int64_t *slice_ptr, const int32_t *idx, const double *val, int64_t nrows,
                 int64_t nslices, int64_t uniform_width, const double *dict, int idx_bits, int ndict, const double *y,
                 EmulVec c, EmulVec T, EmulVec lb, EmulVec ub, double *x, double *xbar, double *d_out, int has_eq,
                 int has_ineq, double theta, double one_plus_theta) {
  SellView AT = make_view(slice_ptr, idx, val, nrows, nslices, uniform_width, dict, idx_bits, ndict);
  Vec vc{c.p, c.c}, vT{T.p, T.c}, vlb{lb.p, lb.c}, vub{ub.p, ub.c};
  PrimalFn fn = primal_kernel(write_d != 0, dict != nullptr, variant);
  run_grid(nslices, dict != nullptr,
           [&] { fn(AT, y, vc, vT, vlb, vub, x, xbar, d_out, nrows, has_eq, has_ineq, theta, one_plus_theta, nullptr); });
}

void emul_dual(int variant, const int64_t *slice_ptr, const int32_t *idx, const double *val, int64_t nrows, int64_t nslices,
               int64_t uniform_width, const double *dict, int idx_bits, int ndict, const double *xbar, EmulVec b,
               EmulVec sigma, double *y, int64_t m_eq) {
  SellView A = make_view(slice_ptr, idx, val, nrows, nslices, uniform_width, dict, idx_bits, ndict);
  Vec vb{b.p, b.c}, vs{sigma.p, sigma.c};
  DualFn fn = dual_kernel(dict != nullptr, variant);
  run_grid(nslices, dict != nullptr, [&] { fn(A, xbar, vb, vs, y, nrows, m_eq, nullptr); });
}

int emul_num_variants() { return kNumVariants; }

int emul_constants(int which) {
  switch (which) {
    case 0: return kSlice;
    case 1: return kBlock;
    case 2: return kEqBit;
    case 3: return kPad;
    default: return 0;
  }
}

}  // extern "C"
