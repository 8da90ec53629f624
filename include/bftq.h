/*
 * bftq.h — C ABI of libbftq.so, the B200-native batched Byzantine-quorum verification engine.
 *
 * This is the drop-in boundary for ONE hot path of yahoo/bftkv: the per-response OpenPGP
 * signature check (crypto/pgp), the web-of-trust quorum tally (quorum/wotqs) and the threshold
 * share-combine (crypto/sss, crypto/threshold).  bftkv is pure Go and has no FFI of its own; the
 * entry points below are what a cgo shim behind bftkv's crypto.Signature /
 * crypto.CollectiveSignature / quorum.Quorum interfaces binds (INTEGRATION.md shows the shim).
 * Each entry point cites the reference interface it replaces (file:line relative to the
 * reference tree).
 *
 * Conventions
 *   - plain pointers and sizes only; caller owns every buffer; nothing is retained after return
 *     (cgo pointer rules), except data explicitly registered (keys, quorums).
 *   - every function returns 0 (BFTQ_OK) or a negative BFTQ_ERR_* code; per-item results go to a
 *     caller-provided status array.
 *   - big integers are big-endian byte strings, exactly as Go's big.Int.Bytes()/SetBytes() and
 *     OpenPGP MPIs carry them; fixed-width fields are left-padded with zeros.
 *   - all *_batch functions are thread-safe and re-entrant (bftkv calls Message.Decrypt from one
 *     goroutine per peer, transport/transport.go:110-127).
 *   - there is NO CPU fallback: if no CUDA device is usable, bftq_init fails with
 *     BFTQ_ERR_NO_DEVICE and nothing else can be called.
 */
#ifndef BFTQ_H
#define BFTQ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFTQ_VERSION 1

/* ---- error codes --------------------------------------------------------------------------- */
#define BFTQ_OK                     0
#define BFTQ_ERR_NO_DEVICE         -1   /* no CUDA device / driver: the engine cannot exist      */
#define BFTQ_ERR_CUDA              -2   /* a CUDA call failed; bftq_last_error() has the text    */
#define BFTQ_ERR_INVALID_ARG       -3
#define BFTQ_ERR_UNSUPPORTED_KEY   -4   /* modulus size / exponent outside what is built         */
#define BFTQ_ERR_NOMEM             -5
#define BFTQ_ERR_INVALID_SIGNATURE -6   /* crypto.ErrInvalidSignature (crypto/crypto.go)         */
#define BFTQ_ERR_INSUFFICIENT_SIGS -7   /* crypto.ErrInsufficientNumberOfSignatures              */
#define BFTQ_ERR_MALFORMED         -8

/* ---- per-item status bytes (SURVEY §8b "Errors") --------------------------------------------
 * The reference collapses every failure to ErrInvalidSignature (crypto_pgp.go:325-327); the shim
 * maps any non-zero status to that sentinel.  The richer codes exist for statistics only. */
#define BFTQ_ST_OK              0
#define BFTQ_ST_BAD_SIGNATURE   1   /* s^e mod n != EMSA-PKCS1-v1_5(digest)                      */
#define BFTQ_ST_HASH_TAG        2   /* OpenPGP 16-bit hash-tag pre-check failed                  */
#define BFTQ_ST_MALFORMED       3
#define BFTQ_ST_UNKNOWN_SIGNER  4   /* key index out of range / issuer not in keyring            */
#define BFTQ_ST_UNSUPPORTED     5   /* algorithm the reference's library cannot verify either    */
#define BFTQ_ST_MISSING         6   /* no response from this replica (tally input only)          */

/* ---- hash algorithm ids = OpenPGP ids (RFC 4880 §9.4), as sig.Hash in x/crypto ---------------*/
#define BFTQ_HASH_MD5        1
#define BFTQ_HASH_SHA1       2
#define BFTQ_HASH_RIPEMD160  3
#define BFTQ_HASH_SHA256     8
#define BFTQ_HASH_SHA384     9
#define BFTQ_HASH_SHA512    10
#define BFTQ_HASH_SHA224    11

/* ---- flags --------------------------------------------------------------------------------- */
#define BFTQ_F_STRICT_RANGE  0x1u  /* reject s >= n.  Default (0) matches Go 1.13's
                                      rsa.VerifyPKCS1v15, which computes s^e mod n for any
                                      k-byte s (no range check before Go 1.20).                  */

typedef struct bftq_engine bftq_engine;

/* ---- engine life cycle --------------------------------------------------------------------- */
/* One engine per GPU.  Wired where the reference calls pgp.New() (cmd/bftkv/main.go:66,
 * api/api.go:37).  device = CUDA ordinal. */
int  bftq_init(int device, bftq_engine** out);
void bftq_shutdown(bftq_engine* e);
const char* bftq_last_error(void);          /* thread-local text of the last failure */
int  bftq_version(void);
int  bftq_device_sm_count(bftq_engine* e);

/* ---- key table ------------------------------------------------------------------------------
 * Replaces the keyring lookup inside openpgp.CheckDetachedSignature (EntityList.KeysByIdUsage),
 * reached from crypto/pgp/crypto_pgp.go:324,338,490.  Registers `count` RSA public keys:
 * n_be = count x 256 bytes (big-endian modulus, left-padded), e = count public exponents.
 * Precomputes the per-key Montgomery constants once.  Keys are appended; *first_index receives
 * the index of the first new key (indices are what key_idx[] refers to).  Round 1 accepts
 * moduli of 2041..2048 bits (k = 256, what gpg --quick-gen-key rsa2048 produces). */
int bftq_register_rsa_keys(bftq_engine* e, const uint8_t* n_be, const uint32_t* exps,
                           uint32_t count, uint32_t* first_index);
int bftq_key_count(bftq_engine* e);

/* ---- K1: batched RSA PKCS#1 v1.5 verify ------------------------------------------------------
 * Replaces x/crypto packet.PublicKey.VerifySignature -> rsa.VerifyPKCS1v15 (Go 1.13) as reached
 * from crypto_pgp.go:324 (Signature.Verify), :338 (VerifyWithCertificate), :490
 * (CollectiveSignature.Verify) and :454 (Message.Decrypt's m.SignatureError).
 *   key_idx[i]      index into the key table (>= key count -> BFTQ_ST_UNKNOWN_SIGNER)
 *   sig_be          n_items x 256 bytes, the signature MPI left-padded to the key size
 *                   (what x/crypto's padToKeySize hands to rsa.VerifyPKCS1v15)
 *   digest          n_items x digest_len(hash_alg) bytes, the OpenPGP v4 signature digest
 *   out_status[i]   BFTQ_ST_*
 * Host buffers; the call stages them through pinned memory, runs the kernel and copies the
 * status bytes back before returning. */
int bftq_rsa_verify_batch(bftq_engine* e, const uint32_t* key_idx, const uint8_t* sig_be,
                          const uint8_t* digest, uint32_t hash_alg, uint64_t n_items,
                          uint32_t flags, uint8_t* out_status);

/* Same, but every pointer is a DEVICE pointer and the work is enqueued on `cuda_stream`
 * (a cudaStream_t, may be NULL for the default stream) without synchronising. */
int bftq_rsa_verify_batch_dev(bftq_engine* e, const uint32_t* d_key_idx, const uint8_t* d_sig_be,
                              const uint8_t* d_digest, uint32_t hash_alg, uint64_t n_items,
                              uint32_t flags, uint8_t* d_status, void* cuda_stream);

/* ---- statistics ------------------------------------------------------------------------------
 * Counters since bftq_init (SURVEY §5 "metrics"): items verified, kernel launches. */
typedef struct {
  uint64_t items;          /* tuples pushed through bftq_rsa_verify_batch*            */
  uint64_t launches;       /* CUDA kernel launches issued by this engine              */
  uint64_t h2d_bytes;
  uint64_t d2h_bytes;
} bftq_stats_t;
int bftq_stats(bftq_engine* e, bftq_stats_t* out);

/* ---- integer-pipe peak (roofline denominator) -----------------------------------------------
 * Runs an unrolled dependency-free mad.wide.u32 micro-benchmark on the engine's device and
 * returns the measured rate in 32x32->64 multiply-accumulates per second ("measure
 * an unrolled mad.lo.u32 microbenchmark on the box; do not hard-code a datasheet number"). */
int bftq_measure_int_peak(bftq_engine* e, double* macs_per_second);

#ifdef __cplusplus
}
#endif
#endif /* BFTQ_H */
