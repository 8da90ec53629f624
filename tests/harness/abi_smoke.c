/* C harness for include/bftq.h: drives the call sequence of the Go shim (integration/bftkv/crypto/gpu/gpu.go) from plain
 * C, so that the header is exercised by a C compiler and not only through ctypes:
 *   bftq_init -> bftq_host_alloc -> bftq_keyring_create / _add -> bftq_signature_verify_batch -> bftq_collective_verify_batch
 *   -> bftq_message_verify_batch -> bftq_read_responses_batch -> bftq_graph_* -> bftq_stats -> bftq_shutdown
 * Input: a fixture file written by tests/test_abi_harness.py (sections of length-prefixed blobs + expectations).
 * Exit code 0 = every call returned what the fixture expects.  Without a CUDA device bftq_init must fail with
 * BFTQ_ERR_NO_DEVICE and the harness exits 77 (the CPU suite checks exactly that). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bftq.h"

static uint8_t* g_buf; static size_t g_len, g_pos;
static uint64_t rd64(void) { uint64_t v; memcpy(&v, g_buf + g_pos, 8); g_pos += 8; return v; }
static const uint8_t* rdbytes(size_t n) { const uint8_t* p = g_buf + g_pos; g_pos += n; return p; }
#define CHECK(cond, msg) do { if (!(cond)) { fprintf(stderr, "abi_smoke: %s (%s) [%s]\n", msg, #cond, bftq_last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: abi_smoke fixture.bin\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  fseek(f, 0, SEEK_END); g_len = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
  g_buf = (uint8_t*)malloc(g_len);
  if (fread(g_buf, 1, g_len, f) != g_len) return 2;
  fclose(f);
  if (bftq_version() != BFTQ_VERSION) return 1;

  bftq_engine* e = NULL;
  int rc = bftq_init(0, &e);
  if (rc == BFTQ_ERR_NO_DEVICE) { fprintf(stderr, "abi_smoke: no CUDA device (%s)\n", bftq_last_error()); return 77; }
  CHECK(rc == BFTQ_OK && e != NULL, "bftq_init");
  CHECK(bftq_device_sm_count(e) > 0, "sm count");

  /* keyring */
  bftq_keyring* kr = NULL;
  CHECK(bftq_keyring_create(e, &kr) == BFTQ_OK, "keyring_create");
  uint64_t ring_len = rd64();
  const uint8_t* ring = rdbytes(ring_len);
  uint32_t n_ent = 0;
  CHECK(bftq_keyring_add(kr, ring, ring_len, 0, &n_ent) == BFTQ_OK && n_ent == rd64(), "keyring_add");

  /* Signature.Verify batch: n items, blobs in page-locked memory as the shim's aggregator keeps them */
  uint64_t n = rd64(), tbs_len = rd64(), sig_len = rd64();
  const uint8_t* tbs = rdbytes(tbs_len);
  const uint64_t* tbs_off = (const uint64_t*)rdbytes((n + 1) * 8);
  const uint8_t* sig = rdbytes(sig_len);
  const uint64_t* sig_off = (const uint64_t*)rdbytes((n + 1) * 8);
  const uint8_t* expect_ok = rdbytes(n);
  void *p_tbs = NULL, *p_sig = NULL;
  CHECK(bftq_host_alloc(e, tbs_len, &p_tbs) == BFTQ_OK && bftq_host_alloc(e, sig_len, &p_sig) == BFTQ_OK, "host_alloc");
  memcpy(p_tbs, tbs, tbs_len); memcpy(p_sig, sig, sig_len);
  int32_t* err = (int32_t*)malloc(n * sizeof(int32_t));
  CHECK(bftq_signature_verify_batch(kr, (const uint8_t*)p_tbs, tbs_off, (const uint8_t*)p_sig, sig_off, n, err) == BFTQ_OK, "signature_verify_batch");
  for (uint64_t i = 0; i < n; i++) CHECK((err[i] == 0) == (expect_ok[i] != 0) && (err[i] == 0 || err[i] == BFTQ_ERR_INVALID_SIGNATURE), "verify verdict");

  /* CollectiveSignature.Verify batch: one clique by node id */
  uint64_t nmem = rd64();
  const uint64_t* members = (const uint64_t*)rdbytes(nmem * 8);
  bftq_qc_ids_t qc;
  qc.f = (int32_t)rd64(); qc.min = (int32_t)rd64(); qc.threshold = (int32_t)rd64(); qc.suff = (int32_t)rd64();
  qc.member_off = 0; qc.member_cnt = (uint32_t)nmem;
  uint64_t nc = rd64(), ctbs_len = rd64(), css_len = rd64();
  const uint8_t* ctbs = rdbytes(ctbs_len);
  const uint64_t* ctbs_off = (const uint64_t*)rdbytes((nc + 1) * 8);
  const uint8_t* css = rdbytes(css_len);
  const uint64_t* css_off = (const uint64_t*)rdbytes((nc + 1) * 8);
  const uint8_t* cexpect = rdbytes(nc);
  int32_t* cerr = (int32_t*)malloc(nc * sizeof(int32_t));
  CHECK(bftq_collective_verify_batch(kr, &qc, 1, members, (uint32_t)nmem, ctbs, ctbs_off, css, css_off, nc, cerr) == BFTQ_OK, "collective_verify_batch");
  for (uint64_t i = 0; i < nc; i++) CHECK((cerr[i] == 0) == (cexpect[i] != 0) && (cerr[i] == 0 || cerr[i] == BFTQ_ERR_INSUFFICIENT_SIGS), "collective verdict");

  /* Message.Decrypt's signature half + the read path from raw answers: one operation */
  uint64_t nr = rd64(), msg_len = rd64(), nonce_len = rd64();
  const uint8_t* msg = rdbytes(msg_len);
  const uint64_t* msg_off = (const uint64_t*)rdbytes((nr + 1) * 8);
  const uint64_t* peers = (const uint64_t*)rdbytes(nr * 8);
  const uint8_t* nonces = rdbytes(nr * nonce_len);
  const uint8_t* mexpect = rdbytes(nr);                 /* 1 = Decrypt returns err == nil */
  uint64_t want_decision = rd64(), want_winner = rd64(), want_at = rd64();
  int32_t* merr = (int32_t*)malloc(nr * sizeof(int32_t));
  uint64_t* by = (uint64_t*)malloc(nr * 8);
  uint8_t* fl = (uint8_t*)malloc(nr);
  uint8_t* plain = (uint8_t*)malloc(msg_len + 1);
  uint32_t* plen = (uint32_t*)malloc(nr * 4);
  CHECK(bftq_message_verify_batch(kr, msg, msg_off, nr, merr, by, fl, plain, plen, NULL, NULL) == BFTQ_OK, "message_verify_batch");
  for (uint64_t i = 0; i < nr; i++) CHECK((merr[i] == 0) == (mexpect[i] != 0), "message verdict");
  uint32_t op_off[2] = {0, (uint32_t)nr};
  uint8_t* rst = (uint8_t*)malloc(nr);
  uint8_t dec = 9; uint32_t win = 0, at = 0;
  CHECK(bftq_read_responses_batch(kr, &qc, 1, members, (uint32_t)nmem, op_off, 1, peers, msg, msg_off, NULL, nonces, (uint32_t)nonce_len, rst, NULL, NULL, NULL,
                                  &dec, &win, &at) == BFTQ_OK, "read_responses_batch");
  CHECK(dec == want_decision && win == (uint32_t)want_winner && at == (uint32_t)want_at, "read decision");
  for (uint64_t i = 0; i < nr; i++) CHECK(((rst[i] == BFTQ_ST_OK || rst[i] == BFTQ_ST_UNVERIFIED_SIGNER) ? 1 : 0) == (mexpect[i] != 0), "answer status");

  /* quorum-descriptor builder + cache */
  bftq_graph* g = NULL;
  CHECK(bftq_graph_create(&g) == BFTQ_OK, "graph_create");
  for (uint64_t i = 0; i < nmem; i++) CHECK(bftq_graph_add_node(g, members[i], members, (uint32_t)nmem) == BFTQ_OK, "graph_add_node");
  CHECK(bftq_graph_set_self(g, members[0]) == BFTQ_OK, "graph_set_self");
  bftq_qc_ids_t out_qc[4]; uint64_t out_m[64]; uint32_t nq = 0, nm = 0;
  CHECK(bftq_graph_choose_quorum(g, BFTQ_RW_AUTH, out_qc, 4, &nq, out_m, 64, &nm) == BFTQ_OK && nq == 1 && nm == nmem, "graph_choose_quorum");
  CHECK(out_qc[0].f == qc.f && out_qc[0].min == qc.min, "descriptor");
  uint64_t ver = 0, hits = 0, builds = 0;
  CHECK(bftq_graph_choose_quorum(g, BFTQ_RW_AUTH, out_qc, 4, &nq, out_m, 64, &nm) == BFTQ_OK, "graph_choose_quorum (cached)");
  CHECK(bftq_graph_version(g, &ver, &hits, &builds) == BFTQ_OK && hits == 1 && builds == 1, "descriptor cache");
  bftq_graph_destroy(g);

  bftq_stats_t st;
  CHECK(bftq_stats(e, &st) == BFTQ_OK && st.launches > 0 && st.items > 0, "stats");
  CHECK(bftq_host_free(e, p_tbs) == BFTQ_OK && bftq_host_free(e, p_sig) == BFTQ_OK, "host_free");
  bftq_keyring_destroy(kr);
  bftq_shutdown(e);
  printf("abi_smoke ok: %llu verifies, %llu collective, %llu answers\n", (unsigned long long)n, (unsigned long long)nc, (unsigned long long)nr);
  return 0;
}
