// Host build of bftkv_b200/csrc/p256.cuh (the same __host__ __device__ code the kernels run).
#include "../../bftkv_b200/csrc/p256.cuh"
using namespace bftq::p256;
extern "C" {
// out = k * P  (P: 65-byte uncompressed); returns 0 if P is not on the curve, 2 for infinity
int p256_mul_host(const uint8_t* point, const uint8_t* k_be, uint8_t* out65) {
  pt p; if (!pt_from_uncompressed(p, point)) return 0;
  uint32_t k[8]; be_to_limbs(k, k_be);
  pt r; pt_mul(r, p, k);
  out65[0] = 4;
  return pt_to_affine(out65 + 1, out65 + 33, r) ? 1 : 2;
}
int p256_add_host(const uint8_t* a, const uint8_t* b, uint8_t* out65) {
  pt p, q; if (!pt_from_uncompressed(p, a) || !pt_from_uncompressed(q, b)) return 0;
  pt r; pt_add(r, p, q); out65[0] = 4;
  return pt_to_affine(out65 + 1, out65 + 33, r) ? 1 : 2;
}
int p256_ecdsa_verify_host(const uint8_t* pub65, const uint8_t* r_be, const uint8_t* s_be, const uint8_t* digest, int dlen) {
  pt q; if (!pt_from_uncompressed(q, pub65)) return -1;
  uint32_t r[8], s[8]; be_to_limbs(r, r_be); be_to_limbs(s, s_be);
  return ecdsa_verify_core(q, r, s, digest, dlen) ? 1 : 0;
}
}
