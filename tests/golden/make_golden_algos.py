#!/usr/bin/env python3
"""Generates tests/golden/golden_algos.json: GnuPG-made ECDSA (NIST P-256) and DSA (1024/2048/3072) keys
and detached signatures, pinning the oracle's ECDSA / DSA arms (x/crypto VerifySignature algorithms 19 and
17).  Run once in the authoring container (needs gpg); the output is committed."""
import json, os, subprocess, sys, tempfile
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
from oracle import packet_oracle as pk, pgp_oracle as pgp

home = tempfile.mkdtemp(prefix="gnupg-")
os.chmod(home, 0o700)
env = dict(os.environ, GNUPGHOME=home)


def gpg(*args):
    r = subprocess.run(["gpg", "--batch", "--yes", "--no-tty", "--pinentry-mode", "loopback", "--passphrase", ""] + list(args),
                       env=env, capture_output=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr.decode())
    return r.stdout


keys, cases = {}, []
for name, algo, digs in [("p256a", "nistp256", ("SHA256", "SHA512", "SHA1")), ("p256b", "nistp256", ("SHA384", "SHA224")),
                         ("dsa1024", "dsa1024", ("SHA1", "SHA256")), ("dsa2048", "dsa2048", ("SHA256", "SHA512")),
                         ("dsa3072", "dsa3072", ("SHA256", "SHA384"))]:
    gpg("--quick-gen-key", "%s <t@example.com>" % name, algo, "sign,cert", "never")
    fpr = [l.split(":")[9] for l in gpg("--with-colons", "--list-keys", name).decode().splitlines() if l.startswith("fpr")][0]
    pub = gpg("--export", fpr)
    ent = pgp.read_entities(pub)[0]
    keys[name] = {"pub": pub.hex(), "key_id": "%016x" % ent.primary_key.key_id, "algo": ent.primary_key.algo}
    for i, m in enumerate([pk.serialize(b"test", b"test", 1), b"another message", b"", bytes(range(256)) * 3]):
        for dig in digs:
            with tempfile.NamedTemporaryFile(dir=home, delete=False) as f:
                f.write(m)
            try:
                gpg("-u", fpr, "--digest-algo", dig, "--weak-digest", "MD5", "-o", f.name + ".sig", "--detach-sign", f.name)
            except RuntimeError as ex:
                print("skip", name, dig, str(ex).strip().splitlines()[-1])
                continue
            ok = subprocess.run(["gpg", "--batch", "--verify", f.name + ".sig", f.name], env=env, capture_output=True).returncode == 0
            cases.append({"tbs": m.hex(), "sig": open(f.name + ".sig", "rb").read().hex(), "signer": name, "hash": dig, "gpg_ok": ok})
out = os.path.join(os.path.dirname(__file__), "golden_algos.json")
json.dump({"generator": "tests/golden/make_golden_algos.py", "keys": keys, "cases": cases}, open(out, "w"))
print("wrote", out, len(cases), {k: v["algo"] for k, v in keys.items()})
