"""CPU oracle for the bftkv signature-verify + quorum-tally hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import, call, link or execute it, and there only
as the checker (or the timed CPU baseline), never as the thing shipped.

What it restates (file:line relative to the yahoo/bftkv reference tree):

* ``packet_oracle``  — packet/packet.go:35-235  (Serialize/Parse/TBS/TBSS)
* ``wotqs_oracle``   — node/graph/graph.go:46-75,117-125,279-393,420-438 and
                       quorum/wotqs/wotqs.go (whole file)
* ``sss_oracle``     — crypto/sss/sss.go:23-107, crypto/threshold/dsa/dsa_core.go:375-403,
                       crypto/threshold/dsa/dsa.go:33-52, crypto/threshold/rsa/rsa.go:318-393
* ``pgp_oracle``     — crypto/pgp/crypto_pgp.go:319-344,373-405,485-515 plus the
                       third-party arithmetic those lines call
                       (golang.org/x/crypto/openpgp @ v0.0.0-20191227163750-53104e6ec876 and
                       Go 1.13 crypto/rsa.VerifyPKCS1v15 — NOT in the reference tree; restated
                       from RFC 4880 / RFC 8017 and the module's published behaviour)
* ``c/bftq_oracle.c`` — the same RSA / tally / Lagrange arithmetic in plain C (own
                       big-number code, pthreads) for full-size batches and the CPU baseline.

Pinning status (see DESIGN.md "Oracle"):
* RSA PKCS#1 v1.5 verify: pinned against GnuPG 2.4.4-made detached signatures
  (tests/golden/), OpenSSL (`cryptography`) and the reference-owned key
  crypto/threshold/rsa/test.pkcs8 + rsa_test.go:165-206 (TestCombine KAT).
* Lagrange / SSS: pinned against crypto/auth/auth_test.go:121-155 (S = 1234) and
  crypto/sss/sss_test.go:15-75.
* wotqs tally and the OpenPGP accept/reject rules: the reference ships NO test or golden
  vector for them (crypto/pgp and quorum/wotqs have no _test.go; every integration test is
  t.Skip) and the Go toolchain is absent here, so these parts are "parity unpinned" —
  anchored on the closed-form thresholds of wotqs.go:55-66 and on independent tools (gpg).
"""
