// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11) and the Box-Muller pair the engine draws its Gaussian test matrix
// with: counter p yields the normals of flat elements 2p and 2p+1 of the row-major n x l matrix Omega.  Restated on the
// CPU in oracle/ref_rsvd.py:philox_normal.  Stands in for random_mat_normal (reference mat_utils.rs:161-175).
#pragma once
#include <cstdint>

namespace corrla {

__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// normals (z0, z1) of pair p under key `seed`
__device__ __forceinline__ void philox_normal_pair(int64_t p, uint64_t seed, double* z0, double* z1) {
  uint32_t c[4] = {(uint32_t)(p & 0xffffffffu), (uint32_t)((uint64_t)p >> 32), 0u, 0u};
  philox4x32_10(c, (uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32));
  const uint64_t x = ((uint64_t)c[0] | ((uint64_t)c[1] << 32)) >> 11;
  const uint64_t y = ((uint64_t)c[2] | ((uint64_t)c[3] << 32)) >> 11;
  const double u1 = ((double)x + 1.0) * 0x1.0p-53;   // (0, 1]
  const double u2 = (double)y * 0x1.0p-53;           // [0, 1)
  const double r = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  *z0 = r * cs;
  *z1 = r * sn;
}

}  // namespace corrla
