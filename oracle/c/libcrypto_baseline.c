/* CPU baseline stand-in prescribed by SURVEY §8(d)(2) / BASELINE.md §2 for boxes without a Go toolchain: the path
 * crypto/pgp -> x/crypto openpgp -> rsa.VerifyPKCS1v15 (crypto_pgp.go:324) restated on OpenSSL's libcrypto —
 * EVP_PKEY_verify with RSA_PKCS1_PADDING over a precomputed SHA-256 digest, one pthread per host core.
 *
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY (oracle/__init__.py): used by bench.py's cpu_baseline legs and by
 * `bench.py --impl reference`, never by the product.  It doubles as an independent check of the plain-C oracle
 * (tests/test_oracle_golden.py compares the two on config 2).
 *
 * Status bytes as the oracle's: 0 valid, 1 invalid, 4 unknown signer (key index out of range). */
#include <openssl/bn.h>
#include <openssl/core_names.h>
#include <openssl/evp.h>
#include <openssl/param_build.h>
#include <openssl/rsa.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static EVP_PKEY* make_key(const uint8_t* n_be, uint32_t e) {
  BIGNUM* n = BN_bin2bn(n_be, 256, NULL);
  BIGNUM* ee = BN_new();
  BN_set_word(ee, e);
  OSSL_PARAM_BLD* bld = OSSL_PARAM_BLD_new();
  OSSL_PARAM_BLD_push_BN(bld, OSSL_PKEY_PARAM_RSA_N, n);
  OSSL_PARAM_BLD_push_BN(bld, OSSL_PKEY_PARAM_RSA_E, ee);
  OSSL_PARAM* params = OSSL_PARAM_BLD_to_param(bld);
  EVP_PKEY_CTX* ctx = EVP_PKEY_CTX_new_from_name(NULL, "RSA", NULL);
  EVP_PKEY* pkey = NULL;
  if (EVP_PKEY_fromdata_init(ctx) <= 0 || EVP_PKEY_fromdata(ctx, &pkey, EVP_PKEY_PUBLIC_KEY, params) <= 0) pkey = NULL;
  EVP_PKEY_CTX_free(ctx);
  OSSL_PARAM_free(params);
  OSSL_PARAM_BLD_free(bld);
  BN_free(n);
  BN_free(ee);
  return pkey;
}

typedef struct {
  EVP_PKEY** keys; uint32_t nkeys;
  const uint32_t* key_idx; const uint8_t* sig; const uint8_t* digest;
  uint64_t lo, hi; uint8_t* status;
} job_t;

static void* worker(void* p) {
  job_t* j = (job_t*)p;
  /* one verify context per key and thread, reused across items (what a server keeps per peer) */
  EVP_PKEY_CTX** ctx = (EVP_PKEY_CTX**)calloc(j->nkeys ? j->nkeys : 1, sizeof(*ctx));
  for (uint64_t i = j->lo; i < j->hi; i++) {
    const uint32_t k = j->key_idx[i];
    if (k >= j->nkeys) { j->status[i] = 4; continue; }
    if (!ctx[k]) {
      ctx[k] = EVP_PKEY_CTX_new(j->keys[k], NULL);
      EVP_PKEY_verify_init(ctx[k]);
      EVP_PKEY_CTX_set_rsa_padding(ctx[k], RSA_PKCS1_PADDING);
      EVP_PKEY_CTX_set_signature_md(ctx[k], EVP_sha256());
    }
    j->status[i] = EVP_PKEY_verify(ctx[k], j->sig + i * 256, 256, j->digest + i * 32, 32) == 1 ? 0 : 1;
  }
  for (uint32_t k = 0; k < j->nkeys; k++) if (ctx[k]) EVP_PKEY_CTX_free(ctx[k]);
  free(ctx);
  return 0;
}

/* RSA-2048 / SHA-256 batch verify on `threads` host threads.  Returns 0, or -1 when a key cannot be built. */
int lcb_rsa_verify_batch(const uint8_t* keys_n_be, const uint32_t* keys_e, uint32_t nkeys, const uint32_t* key_idx, const uint8_t* sig_be,
                         const uint8_t* digest, uint64_t n_items, int threads, uint8_t* status) {
  EVP_PKEY** keys = (EVP_PKEY**)calloc(nkeys ? nkeys : 1, sizeof(*keys));
  for (uint32_t k = 0; k < nkeys; k++) {
    keys[k] = make_key(keys_n_be + (size_t)k * 256, keys_e[k]);
    if (!keys[k]) { for (uint32_t q = 0; q < k; q++) EVP_PKEY_free(keys[q]); free(keys); return -1; }
  }
  if (threads < 1) threads = 1;
  if ((uint64_t)threads > n_items) threads = n_items ? (int)n_items : 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * threads);
  for (int t = 0; t < threads; t++) {
    jobs[t] = (job_t){keys, nkeys, key_idx, sig_be, digest, n_items * t / threads, n_items * (t + 1) / threads, status};
    pthread_create(&th[t], 0, worker, &jobs[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(th[t], 0);
  for (uint32_t k = 0; k < nkeys; k++) EVP_PKEY_free(keys[k]);
  free(th); free(jobs); free(keys);
  return 0;
}
