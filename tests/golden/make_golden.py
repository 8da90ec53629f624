#!/usr/bin/env python3
"""Generates tests/golden/golden.json — run ONCE in the authoring container (needs gpg 2.4,
`cryptography`, and /root/reference for the reference-owned RSA test key).  The output is
committed; nothing at test time reads /root/reference or runs gpg.

Sources of truth, all independent of this repo's oracle and kernels:
  * GnuPG 2.4.4 creates the RSA-2048 keys (explicit rsa2048), the trust
    certifications (scripts/clique.sh / trust.sh topology, 4-node clique + client u01) and the
    detached signatures; `gpg --verify` confirms each one.
  * crypto/threshold/rsa/test.pkcs8 (reference-owned key) + OpenSSL: the deterministic PKCS#1
    v1.5 / SHA-256 signature over "tbs" that rsa_test.go:165-206 (TestCombine) must reproduce.
  * crypto/auth/auth_test.go:121-155 and crypto/sss/sss_test.go:15-75 numbers (Python ints).
"""
import hashlib, json, os, subprocess, sys, tempfile
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
from oracle import packet_oracle as pk, pgp_oracle as pgp

OUT = os.path.join(os.path.dirname(__file__), "golden.json")
home = tempfile.mkdtemp(prefix="gnupg-")
os.chmod(home, 0o700)
env = dict(os.environ, GNUPGHOME=home)


def gpg(*args, inp=None):
    r = subprocess.run(["gpg", "--batch", "--yes", "--no-tty", "--pinentry-mode", "loopback", "--passphrase", ""] + list(args),
                       input=inp, env=env, capture_output=True)
    if r.returncode != 0:
        raise RuntimeError("gpg %s failed: %s" % (args, r.stderr.decode()))
    return r.stdout


names = ["a01", "a02", "a03", "a04", "u01", "x99"]          # x99: a key that is NOT in the test keyring
ports = {"a01": 5701, "a02": 5702, "a03": 5703, "a04": 5704}
fpr = {}
for nme in names:
    uid = "%s (http://localhost:%d)" % (nme, ports[nme]) if nme in ports else "%s <foo@example.com>" % nme
    gpg("--quick-gen-key", uid, "rsa2048", "sign,cert", "never")
    out = gpg("--with-colons", "--list-keys", nme).decode()
    fpr[nme] = [l.split(":")[9] for l in out.splitlines() if l.startswith("fpr")][0]
    gpg("--quick-add-key", fpr[nme], "rsa2048", "encr", "never")
clique = ["a01", "a02", "a03", "a04"]
for s in clique:                                   # clique.sh: everyone signs everyone
    for t in clique:
        if s != t:
            gpg("-u", fpr[s], "--quick-sign-key", fpr[t])
for t in clique:                                   # trust.sh: u01 -> a0x, and a0x certify u01
    gpg("-u", fpr["u01"], "--quick-sign-key", fpr[t])
    gpg("-u", fpr[t], "--quick-sign-key", fpr["u01"])

keys = {}
for nme in names:
    pub = gpg("--export", fpr[nme])
    ent = pgp.read_entities(pub)[0]
    assert "%016X" % ent.primary_key.key_id == fpr[nme][-16:], "key id restatement disagrees with gpg"
    keys[nme] = {"pub": pub.hex(), "key_id": "%016x" % ent.primary_key.key_id,
                 "n": "%x" % ent.primary_key.n, "e": ent.primary_key.e,
                 "certifiers": ["%016x" % c for c in ent.certifier_ids]}

cases = []
msgs = [pk.serialize(b"test", b"test", 1), pk.serialize(b"x" * 16, b"v" * 32, 2**40 + 7), pk.serialize(b"k", None, 0),
        b"", b"line1\nline2\r\nline3\n"]
for mi, m in enumerate(msgs):
    for si, signer in enumerate(clique + ["u01", "x99"]):
        for algo in (["SHA256"] if (mi + si) % 3 else ["SHA256", "SHA512", "SHA1"]):
            with tempfile.NamedTemporaryFile(dir=home, delete=False) as f:
                f.write(m)
            sigf = f.name + ".sig"
            extra = ["--textmode"] if (mi == 4 and si == 0) else []
            gpg("-u", fpr[signer], "--digest-algo", algo, *extra, "-o", sigf, "--detach-sign", f.name)
            v = subprocess.run(["gpg", "--batch", "--verify", sigf, f.name], env=env, capture_output=True)
            assert v.returncode == 0, v.stderr
            sig = open(sigf, "rb").read()
            cases.append({"tbs": m.hex(), "sig": sig.hex(), "signer": signer, "hash": algo, "gpg": "good"})

# reference-owned RSA key: crypto/threshold/rsa/test.pkcs8 (rsa_test.go:165-206 TestCombine)
ref = {}
p8 = "/root/reference/crypto/threshold/rsa/test.pkcs8"
if os.path.exists(p8):
    from cryptography.hazmat.primitives import serialization, hashes
    from cryptography.hazmat.primitives.asymmetric import padding
    raw = open(p8, "rb").read()
    try:
        key = serialization.load_der_private_key(raw, None)
    except ValueError:
        key = serialization.load_pem_private_key(raw, None)
    sig = key.sign(b"tbs", padding.PKCS1v15(), hashes.SHA256())
    assert hashlib.sha256(sig).hexdigest() == "1d2cef7b44c674e771fdac4fb0f278c75e7b68fe40a835cd39ba5c23cd127998"
    pn = key.public_key().public_numbers()
    ref = {"n": "%x" % pn.n, "e": pn.e, "tbs": b"tbs".hex(), "sig": sig.hex(),
           "digest": hashlib.sha256(b"tbs").hexdigest(), "sig_sha256": hashlib.sha256(sig).hexdigest()}

sss = {"auth_test": {"q": 1237, "poly": [1234, 166, 94, 666], "xs": [2, 4, 5, 6], "lambda": [5, 1222, 16, 1232], "S": 1234,
                      "shares": [[1, 923], [2, 1085], [3, 768], [4, 257], [5, 1074], [6, 1030]]},
       "sss_test": {"secret": b"secret".hex(), "n": 10, "k": 7}}

dsa = {}
dt = "/root/reference/crypto/threshold/dsa/dsa_test.go"
if os.path.exists(dt):                      # P, Q, G constants of dsa_test.go:26-28 (numbers only)
    import re
    src = open(dt).read()
    dsa = {k: re.search(r'%sstr = "([0-9A-F]+)"' % k, src).group(1) for k in "PQG"}
    dsa["source"] = "crypto/threshold/dsa/dsa_test.go:26-28 (parameters of dsa/test.pkcs8)"

json.dump({"generator": "tests/golden/make_golden.py", "gpg": "2.4.4", "keys": keys, "cases": cases,
           "ref_rsa_kat": ref, "sss": sss, "dsa_test": dsa}, open(OUT, "w"), indent=0)
print("wrote", OUT, len(cases), "cases")
