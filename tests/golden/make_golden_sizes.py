#!/usr/bin/env python3
"""Generates tests/golden/golden_sizes.json: GnuPG-made RSA-3072 / RSA-4096 keys and detached
signatures (SHA-256 and SHA-512), for the key-size classes beyond RSA-2048.  Run once in the
authoring container (needs gpg); the output is committed."""
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
for name, algo in [("k3072", "rsa3072"), ("k4096", "rsa4096"), ("k2048", "rsa2048")]:
    gpg("--quick-gen-key", "%s <t@example.com>" % name, algo, "sign,cert", "never")
    fpr = [l.split(":")[9] for l in gpg("--with-colons", "--list-keys", name).decode().splitlines() if l.startswith("fpr")][0]
    pub = gpg("--export", fpr)
    ent = pgp.read_entities(pub)[0]
    keys[name] = {"pub": pub.hex(), "key_id": "%016x" % ent.primary_key.key_id, "bits": ent.primary_key.n.bit_length()}
    for i, m in enumerate([pk.serialize(b"test", b"test", 1), b"another message", b""]):
        for dig in ("SHA256", "SHA512"):
            with tempfile.NamedTemporaryFile(dir=home, delete=False) as f:
                f.write(m)
            gpg("-u", fpr, "--digest-algo", dig, "-o", f.name + ".sig", "--detach-sign", f.name)
            assert subprocess.run(["gpg", "--batch", "--verify", f.name + ".sig", f.name], env=env, capture_output=True).returncode == 0
            cases.append({"tbs": m.hex(), "sig": open(f.name + ".sig", "rb").read().hex(), "signer": name, "hash": dig})
out = os.path.join(os.path.dirname(__file__), "golden_sizes.json")
json.dump({"generator": "tests/golden/make_golden_sizes.py", "keys": keys, "cases": cases}, open(out, "w"))
print("wrote", out, len(cases), [k["bits"] for k in keys.values()])
