#!/usr/bin/env python3
"""Makes tests/golden/golden_messages.json: signed OpenPGP messages written by GnuPG (one-pass signature, literal data,
signature — the packet stream a bftkv transport message decrypts to, crypto_pgp.go:453-471) with GnuPG's own verdict
on each, as the external pin for the oracle's message_verify (the reference ships no such vector).

  python tests/golden/make_golden_messages.py        # needs gpg (2.4 here); writes the JSON next to this file
"""
import base64
import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))


def gpg(home, *args, stdin=None):
    r = subprocess.run(["gpg", "--homedir", home, "--batch", "--yes", "--quiet", "--pinentry-mode", "loopback", "--passphrase", ""] + list(args),
                       input=stdin, capture_output=True)
    return r


def main():
    home = tempfile.mkdtemp(prefix="bftq-gpg-")
    os.chmod(home, 0o700)
    for name in ("m01", "m02"):
        r = gpg(home, "--quick-gen-key", "%s (http://localhost:57%s) <%s@bftq.test>" % (name, name[1:], name), "rsa2048", "sign,cert", "never")
        assert r.returncode == 0, r.stderr
    pub = {n: gpg(home, "--export", n).stdout for n in ("m01", "m02")}
    nonce = bytes(range(8))
    fname = base64.b64encode(nonce).decode()
    body_small = b"\x00\x00\x00\x00\x00\x00\x00\x04test" + b"\x00\x00\x00\x00\x00\x00\x00\x05value" + (7).to_bytes(8, "big")
    body_big = bytes((i * 7 + 3) & 0xFF for i in range(20000))
    body_text = b"line one\nline two\r\nline three\n"
    cases = []

    def add(name, args, body, via_stdin=False, signer="m01", mutate=None, set_name=fname):
        out = os.path.join(home, "out.gpg")
        src = os.path.join(home, "in.bin")
        open(src, "wb").write(body)
        a = ["--sign", "-u", signer, "-o", out, "--set-filename", set_name] + args
        r = gpg(home, *a, stdin=body) if via_stdin else gpg(home, *(a + [src]))
        assert r.returncode == 0, r.stderr
        msg = bytearray(open(out, "rb").read())
        if mutate is not None:
            mutate(msg, body)
        open(out, "wb").write(bytes(msg))
        v = gpg(home, "--verify", out)
        cases.append({"name": name, "msg": bytes(msg).hex(), "body": body.hex(), "signer": signer, "gpg_good": v.returncode == 0,
                      "file_name": set_name})

    nc = ["--compress-algo", "none"]
    add("file-binary-sha256", nc + ["--digest-algo", "SHA256"], body_small)
    add("stdin-partial-sha256", nc + ["--digest-algo", "SHA256"], body_big, via_stdin=True)
    add("file-big-sha512", nc + ["--digest-algo", "SHA512"], body_big)
    add("textmode-sha256", nc + ["--digest-algo", "SHA256", "--textmode"], body_text)
    add("file-sha1", nc + ["--digest-algo", "SHA1"], body_small)
    add("other-signer", nc + ["--digest-algo", "SHA256"], body_small, signer="m02")
    add("name-not-base64", nc + ["--digest-algo", "SHA256"], body_small, set_name="reply.bin")
    add("compressed-default", [], body_small)

    def flip_body(msg, body):
        i = bytes(msg).find(body[:16])
        assert i > 0
        msg[i + 5] ^= 0x20
    add("tampered-body", nc + ["--digest-algo", "SHA256"], body_small, mutate=flip_body)

    def flip_sig(msg, body):
        msg[-20] ^= 0x01
    add("tampered-signature", nc + ["--digest-algo", "SHA256"], body_small, mutate=flip_sig)

    def truncate(msg, body):
        del msg[-40:]
    add("truncated-signature", nc + ["--digest-algo", "SHA256"], body_small, mutate=truncate)
    ver = subprocess.run(["gpg", "--version"], capture_output=True, text=True).stdout.splitlines()[0]
    json.dump({"made_by": ver, "keyring": pub["m01"].hex(), "outsider": pub["m02"].hex(), "nonce": nonce.hex(), "cases": cases},
              open(os.path.join(HERE, "golden_messages.json"), "w"), indent=1)
    print("wrote", len(cases), "cases;", ver)


if __name__ == "__main__":
    main()
