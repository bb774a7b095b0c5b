// Pieces shared by the two projection kernels (tensor-core and CUDA-core).
#pragma once
#include "common.cuh"

namespace psa {

// Exactness bound of one accumulation pass.  The worst digit-pair class (i + j == 3) adds at most
// 2*(128*64) + 2*(128*128) = 49152 per atom to an int32 accumulator, so 32768 atoms can never
// overflow.  Longer contractions are split into passes; later passes add into P in float32.
constexpr int64_t kMaxAtomsPerPass = 32768;

// Sum over kept digit-pair classes: phase * value = T * 2^(e-36) with
// T = c0 + 256 c1 + 256^2 c2 + 256^3 c3 (c_n = int32 accumulator of class i+j = 3+n), exact in
// int64, rounded once to float32.
// A row whose exponent is kExpPoison held a NaN or an infinity (or overflowed the fixed-point range): its
// projections come out as NaN, like the reference's float arithmetic would propagate them.
__device__ __forceinline__ float combine_classes(int32_t c0, int32_t c1, int32_t c2, int32_t c3, int e) {
  long long t = (long long)c0 + ((long long)c1 << 8) + ((long long)c2 << 16) + ((long long)c3 << 24);
  const float scale = e == kExpPoison ? __int_as_float(0x7fc00000) : __int_as_float((e - 36 + 127) << 23);
  return __ll2float_rn(t) * scale;
}

}  // namespace psa
