// Package gpu is the drop-in for bftkv's signature-verification + quorum-tally hot path on libbftq.so (B200).
//
//	crypt := gpu.New(0)                       // instead of pgp.New()      (cmd/bftkv/main.go:66, api/api.go:37,
//	                                          //                            protocol/test_utils/test_utils.go:35,47)
//	qs := gpu.NewQuorumSystem(g, crypt)       // instead of wotqs.New(g)
//
// Everything that signs, encrypts or holds private keys stays in crypto/pgp; what moves is the verification half of
// crypto.Signature / crypto.CollectiveSignature, the signature check inside crypto.Message.Decrypt, and the quorum
// predicates' descriptors.  Not compiled in this repository (the image has no Go toolchain): it is the reviewable
// reference-side binding INTEGRATION.md describes, written against include/bftq.h; tests/harness/abi_smoke.c drives the
// same call sequence from C.
package gpu

/*
#cgo CFLAGS: -I${SRCDIR}/../../../bftq/include
#cgo LDFLAGS: -L${SRCDIR}/../../../bftq/bftkv_b200 -lbftq -Wl,-rpath,${SRCDIR}/../../../bftq/bftkv_b200
#include <stdlib.h>
#include <string.h>
#include "bftq.h"
*/
import "C"

import (
	"bytes"
	"errors"
	"io"
	"io/ioutil"
	"sync"
	"time"
	"unsafe"

	"golang.org/x/crypto/openpgp"
	"golang.org/x/crypto/openpgp/packet"

	"github.com/yahoo/bftkv/crypto"
	"github.com/yahoo/bftkv/crypto/pgp"
	"github.com/yahoo/bftkv/node"
	"github.com/yahoo/bftkv/node/graph"
	bpacket "github.com/yahoo/bftkv/packet"
	"github.com/yahoo/bftkv/quorum"
	"github.com/yahoo/bftkv/quorum/wotqs"
)

// ErrUnsupported is returned for the forms libbftq reports as BFTQ_ERR_UNSUPPORTED / BFTQ_ST_UNSUPPORTED (compressed
// transport messages, RSA > 4096 bit, DSA-3072, ECDSA P-384/P-521): the caller falls back to crypto/pgp for those.
var ErrUnsupported = errors.New("gpu: form not built into libbftq, fall back to crypto/pgp")

type engine struct {
	e   *C.bftq_engine
	kr  *C.bftq_keyring
	agg *aggregator // coalesces concurrent single verifies into batches
}

func lastError() string { return C.GoString(C.bftq_last_error()) }

func boolInt(b bool) C.int {
	if b {
		return 1
	}
	return 0
}

// New returns pgp.New() with the verification half of Signature / CollectiveSignature / Message on the GPU.
func New(device int) *crypto.Crypto {
	c := pgp.New()
	var e *C.bftq_engine
	if rc := C.bftq_init(C.int(device), &e); rc != 0 {
		panic("bftq: " + lastError()) // no CPU fallback by design: a box without a B200 must not start silently slower
	}
	g := &engine{e: e}
	if rc := C.bftq_keyring_create(e, &g.kr); rc != 0 {
		panic("bftq: " + lastError())
	}
	g.agg = newAggregator(g, 200*time.Microsecond, 16384)
	c.Keyring = &keyring{Keyring: c.Keyring, g: g} // mirrors Register / Remove into bftq_keyring
	sig := &signature{Signature: c.Signature, g: g}
	c.Signature = sig // Sign / Certs / Issuer stay in pgp
	c.CollectiveSignature = &collective{CollectiveSignature: c.CollectiveSignature, sig: sig, g: g}
	c.Message = &message{Message: c.Message, g: g, kr: c.Keyring}
	return c
}

// ---- page-locked blobs ---------------------------------------------------------------------------------------------
// Go memory can never be handed to the DMA engine (cgo pointer rules, moving GC), so a batch is assembled in C memory
// anyway; taking that memory from bftq_host_alloc makes the library DMA it in place instead of staging it again.
type hostBuf struct {
	e   *C.bftq_engine
	p   unsafe.Pointer
	n   int // bytes used
	cap int
}

func (b *hostBuf) reset() { b.n = 0 }
func (b *hostBuf) grow(need int) {
	if b.n+need <= b.cap {
		return
	}
	nc := 2*b.cap + need + (1 << 16)
	var p unsafe.Pointer
	if rc := C.bftq_host_alloc(b.e, C.uint64_t(nc), &p); rc != 0 {
		panic("bftq_host_alloc: " + lastError())
	}
	if b.n > 0 {
		C.memcpy(p, b.p, C.size_t(b.n))
	}
	if b.p != nil {
		C.bftq_host_free(b.e, b.p)
	}
	b.p, b.cap = p, nc
}
func (b *hostBuf) write(d []byte) {
	b.grow(len(d))
	if len(d) > 0 {
		C.memcpy(unsafe.Pointer(uintptr(b.p)+uintptr(b.n)), unsafe.Pointer(&d[0]), C.size_t(len(d)))
	}
	b.n += len(d)
}
func (b *hostBuf) u8() *C.uint8_t { b.grow(1); return (*C.uint8_t)(b.p) }

// blob = concatenated items + (n+1) offsets, both page-locked
type blob struct {
	data hostBuf
	off  []C.uint64_t
}

func newBlob(e *C.bftq_engine) *blob { return &blob{data: hostBuf{e: e}, off: []C.uint64_t{0}} }
func (b *blob) reset()             { b.data.reset(); b.off = b.off[:1] }
func (b *blob) add(d []byte)       { b.data.write(d); b.off = append(b.off, C.uint64_t(b.data.n)) }
func (b *blob) offs() *C.uint64_t  { return &b.off[0] } // a Go slice of plain integers may cross for the duration of the call

// ---- keyring mirror (crypto/crypto.go:35-41) -----------------------------------------------------------------------
type keyring struct {
	crypto.Keyring
	g *engine
}

func (k *keyring) Register(nodes []node.Node, priv bool, self bool) error {
	if err := k.Keyring.Register(nodes, priv, self); err != nil {
		return err
	}
	for _, n := range nodes {
		pkt, err := n.Serialize() // OpenPGP public key block, crypto_pgp.go:90-98
		if err != nil {
			return err
		}
		if len(pkt) == 0 {
			continue
		}
		p := C.CBytes(pkt)
		rc := C.bftq_keyring_add(k.g.kr, (*C.uint8_t)(p), C.uint64_t(len(pkt)), boolInt(priv), nil)
		C.free(p)
		if rc != 0 {
			return errors.New("bftq_keyring_add: " + lastError())
		}
	}
	return nil
}

func (k *keyring) Remove(nodes []node.Node) {
	k.Keyring.Remove(nodes)
	ids := make([]C.uint64_t, len(nodes))
	for i, n := range nodes {
		ids[i] = C.uint64_t(n.Id())
	}
	if len(ids) > 0 {
		C.bftq_keyring_remove(k.g.kr, &ids[0], C.uint32_t(len(ids)))
	}
}

// ---- Signature (crypto/crypto.go:50-58) ----------------------------------------------------------------------------
type signature struct {
	crypto.Signature
	g *engine
}

// Verify replaces PGPSignature.Verify (crypto_pgp.go:319-330).  The call blocks until the batch it was coalesced into
// has been verified; many goroutines (transport.Multicast's one-per-peer workers, transport/transport.go:110-127, and
// net/http's one-per-request handlers) share one batch.
func (s *signature) Verify(tbs []byte, sig *bpacket.SignaturePacket) error {
	if sig == nil {
		return crypto.ErrInvalidSignature
	}
	if err := s.g.agg.verify(tbs, sig.Data, nil, false); err != ErrUnsupported {
		return err
	}
	return s.Signature.Verify(tbs, sig) // foreign key kinds (P-384, DSA-3072, RSA > 4096): the reference's own path
}

func (s *signature) VerifyWithCertificate(tbs []byte, sig *bpacket.SignaturePacket, cert node.Node) error {
	if sig == nil || cert == nil {
		return crypto.ErrInvalidSignature
	}
	c, err := cert.Serialize()
	if err != nil {
		return crypto.ErrInvalidSignature
	}
	if err := s.g.agg.verify(tbs, sig.Data, c, true); err != ErrUnsupported { // an EMPTY certificate must fail, not fall back to the shared keyring
		return err
	}
	return s.Signature.VerifyWithCertificate(tbs, sig, cert)
}

// ---- CollectiveSignature (crypto/crypto.go:66-71) ------------------------------------------------------------------
type collective struct {
	crypto.CollectiveSignature
	sig *signature
	g   *engine
}

// describe turns a quorum.Quorum into the flat descriptor libbftq takes.  Only quorums made by NewQuorumSystem carry one;
// for any other implementation the caller keeps the CPU path.
func describe(q quorum.Quorum) (qcs []C.bftq_qc_ids_t, members []C.uint64_t, ok bool) {
	gq, ok := q.(*gpuQuorum)
	if !ok {
		return nil, nil, false
	}
	return gq.qcs, gq.members, true
}
func qcsPtr(q []C.bftq_qc_ids_t) *C.bftq_qc_ids_t {
	if len(q) == 0 {
		return nil
	}
	return &q[0]
}
func idsPtr(m []C.uint64_t) *C.uint64_t {
	if len(m) == 0 {
		return nil
	}
	return &m[0]
}

func (cs *collective) Verify(tbs []byte, ss *bpacket.SignaturePacket, q quorum.Quorum) error {
	qcs, members, ok := describe(q)
	if !ok || ss == nil {
		return cs.CollectiveSignature.Verify(tbs, ss, q)
	}
	tb, sb := newBlob(cs.g.e), newBlob(cs.g.e)
	tb.add(tbs)
	sb.add(ss.Data)
	var rc C.int32_t
	r := C.bftq_collective_verify_batch(cs.g.kr, qcsPtr(qcs), C.uint32_t(len(qcs)), idsPtr(members), C.uint32_t(len(members)),
		tb.data.u8(), tb.offs(), sb.data.u8(), sb.offs(), 1, &rc)
	tb.free()
	sb.free()
	if r == 0 && rc == C.BFTQ_ERR_UNSUPPORTED {
		return cs.CollectiveSignature.Verify(tbs, ss, q) // a packet libbftq could not judge might have tipped the balance
	}
	if r != 0 || rc != 0 {
		return crypto.ErrInsufficientNumberOfSignatures
	}
	ss.Completed = true // crypto_pgp.go:494
	return nil
}

func (cs *collective) Combine(ss *bpacket.SignaturePacket, s *bpacket.SignaturePacket, q quorum.Quorum) bool {
	qcs, members, ok := describe(q)
	if !ok {
		return cs.CollectiveSignature.Combine(ss, s, q)
	}
	if ss.Type == bpacket.SignatureTypeNil { // crypto_pgp.go:507-512
		ss.Type = s.Type
	} else if ss.Type != s.Type {
		return false
	}
	ss.Data = append(ss.Data, s.Data...)
	if len(ss.Data) == 0 {
		return false
	}
	var out C.int32_t
	sb := C.CBytes(ss.Data)
	defer C.free(sb)
	if rc := C.bftq_collective_combine_sufficient(cs.g.kr, qcsPtr(qcs), C.uint32_t(len(qcs)), idsPtr(members), C.uint32_t(len(members)),
		(*C.uint8_t)(sb), C.uint64_t(len(ss.Data)), &out); rc != 0 {
		return false
	}
	return out != 0
}

func (b *blob) free() {
	if b.data.p != nil {
		C.bftq_host_free(b.data.e, b.data.p)
		b.data.p, b.data.cap, b.data.n = nil, 0, 0
	}
}

// ---- aggregator: where "tens of thousands of tuples" come from -----------------------------------------------------
// (libbftq also ships this coalescer natively — C.bftq_aggregator_verify blocks the calling thread until its batch is
// verified; the Go version below avoids pinning an OS thread per waiting goroutine.)
type job struct {
	tbs, sig, cert []byte
	withCert       bool
	done           chan error
}
type aggregator struct {
	g     *engine
	mu    sync.Mutex
	queue []*job
	kick  chan struct{}
	wait  time.Duration
	max   int
	// page-locked blobs, reused across flushes (only the flusher goroutine touches them)
	tbs, sig, cert *blob
}

func newAggregator(g *engine, wait time.Duration, max int) *aggregator {
	a := &aggregator{g: g, kick: make(chan struct{}, 1), wait: wait, max: max, tbs: newBlob(g.e), sig: newBlob(g.e), cert: newBlob(g.e)}
	go a.loop()
	return a
}
func (a *aggregator) verify(tbs, sig, cert []byte, withCert bool) error {
	j := &job{tbs, sig, cert, withCert, make(chan error, 1)}
	a.mu.Lock()
	a.queue = append(a.queue, j)
	full := len(a.queue) >= a.max
	a.mu.Unlock()
	if full {
		select {
		case a.kick <- struct{}{}:
		default:
		}
	}
	return <-j.done
}
func (a *aggregator) loop() {
	C.bftq_bind_thread(a.g.e) // best effort: the goroutine may migrate; the library's own workers are bound
	t := time.NewTicker(a.wait) // deadline- or size-triggered flush
	for {
		select {
		case <-t.C:
		case <-a.kick:
		}
		a.mu.Lock()
		batch := a.queue
		a.queue = nil
		a.mu.Unlock()
		if len(batch) > 0 {
			a.flush(batch)
		}
	}
}

// flush verifies one batch: the jobs without a certificate through bftq_signature_verify_batch, the others through
// bftq_signature_verify_with_cert_batch, and wakes every waiting caller with nil / ErrInvalidSignature.
func (a *aggregator) flush(batch []*job) {
	for pass := 0; pass < 2; pass++ {
		withCert := pass == 1
		a.tbs.reset()
		a.sig.reset()
		a.cert.reset()
		var sel []*job
		for _, j := range batch {
			if j.withCert != withCert {
				continue
			}
			sel = append(sel, j)
			a.tbs.add(j.tbs)
			a.sig.add(j.sig)
			if withCert {
				a.cert.add(j.cert)
			}
		}
		if len(sel) == 0 {
			continue
		}
		errs := make([]C.int32_t, len(sel))
		var rc C.int
		if withCert {
			rc = C.bftq_signature_verify_with_cert_batch(a.g.kr, a.tbs.data.u8(), a.tbs.offs(), a.sig.data.u8(), a.sig.offs(),
				a.cert.data.u8(), a.cert.offs(), C.uint64_t(len(sel)), &errs[0])
		} else {
			rc = C.bftq_signature_verify_batch(a.g.kr, a.tbs.data.u8(), a.tbs.offs(), a.sig.data.u8(), a.sig.offs(),
				C.uint64_t(len(sel)), &errs[0])
		}
		for i, j := range sel {
			switch {
			case rc == 0 && errs[i] == 0:
				j.done <- nil
			case rc == 0 && errs[i] == C.BFTQ_ERR_UNSUPPORTED:
				j.done <- ErrUnsupported // an algorithm / key size libbftq lacks: the caller below re-runs it on crypto/pgp
			default:
				j.done <- crypto.ErrInvalidSignature // every failure kind collapses to this sentinel (crypto_pgp.go:325-327)
			}
		}
	}
}

// ---- Message.Decrypt (crypto/crypto.go:60-64, crypto_pgp.go:453-471) -----------------------------------------------
// The private-key operation and the AES-CFB / MDC layer stay here; the signature check of the decrypted content — the
// "R verifies per read operation" — goes to bftq_message_verify_batch.
type message struct {
	crypto.Message
	g  *engine
	kr crypto.Keyring
}

// inner decrypts the message's SymmetricallyEncrypted packet and returns the packet stream inside.
func (m *message) inner(body io.Reader, priv *openpgp.Entity) ([]byte, error) {
	packets := packet.NewReader(body)
	var key []byte
	var cf packet.CipherFunction
	for {
		p, err := packets.Next()
		if err != nil {
			return nil, err
		}
		switch p := p.(type) {
		case *packet.EncryptedKey:
			for _, k := range priv.Subkeys {
				if k.PrivateKey != nil && k.PublicKey.KeyId == p.KeyId && p.Decrypt(k.PrivateKey, nil) == nil {
					key, cf = p.Key, p.CipherFunc
				}
			}
			if key == nil && priv.PrivateKey != nil && priv.PrimaryKey.KeyId == p.KeyId && p.Decrypt(priv.PrivateKey, nil) == nil {
				key, cf = p.Key, p.CipherFunc
			}
		case *packet.SymmetricallyEncrypted:
			if key == nil {
				return nil, crypto.ErrDecryptionFailed
			}
			r, err := p.Decrypt(cf, key)
			if err != nil {
				return nil, err
			}
			in, err := ioutil.ReadAll(r)
			if err != nil {
				return nil, err
			}
			if err := r.Close(); err != nil { // MDC check
				return nil, err
			}
			return in, nil
		}
	}
}

func (m *message) Decrypt(body io.Reader) (plain []byte, nonce []byte, peer node.Node, err error) {
	priv := privateEntity(m.kr)
	if priv == nil {
		return nil, nil, nil, crypto.ErrDecryptionFailed
	}
	raw, err := ioutil.ReadAll(body) // kept: the forms libbftq does not take go to crypto/pgp unchanged
	if err != nil {
		return nil, nil, nil, crypto.ErrDecryptionFailed
	}
	in, err := m.inner(bytes.NewReader(raw), priv)
	if err != nil {
		return nil, nil, nil, crypto.ErrDecryptionFailed
	}
	b := newBlob(m.g.e)
	defer b.free()
	b.add(in)
	var rc C.int32_t
	var by C.uint64_t
	var fl C.uint8_t
	var plen, nlen C.uint32_t
	pb := make([]byte, len(in)+1)
	nb := make([]byte, len(in)+1)
	r := C.bftq_message_verify_batch(m.g.kr, b.data.u8(), b.offs(), 1, &rc, &by, &fl,
		(*C.uint8_t)(unsafe.Pointer(&pb[0])), &plen, (*C.uint8_t)(unsafe.Pointer(&nb[0])), &nlen)
	if r != 0 {
		return nil, nil, nil, errors.New("bftq_message_verify_batch: " + lastError())
	}
	switch rc {
	case C.BFTQ_ERR_MALFORMED:
		return nil, nil, nil, crypto.ErrDecryptionFailed
	case C.BFTQ_ERR_NOT_SIGNED:
		return nil, nil, nil, crypto.ErrInvalidTransportSecurityData
	case C.BFTQ_ERR_MESSAGE_BODY:
		return nil, nil, nil, io.ErrUnexpectedEOF
	case C.BFTQ_ERR_UNSUPPORTED:
		return m.Message.Decrypt(bytes.NewReader(raw)) // compressed content: crypto/pgp handles the original message
	}
	peer = m.kr.GetCertById(uint64(by)) // may be nil, as in the reference
	if rc == C.BFTQ_ERR_INVALID_SIGNATURE {
		err = crypto.ErrInvalidSignature
	}
	return pb[:plen], nb[:nlen], peer, err
}

// privateEntity: PGPKeyring.getPrivateKey (crypto_pgp.go:199-204) through the exported surface.
func privateEntity(kr crypto.Keyring) *openpgp.Entity {
	if k, ok := kr.(*keyring); ok {
		kr = k.Keyring
	}
	type privGetter interface{ GetPrivateEntity() *openpgp.Entity } // one exported accessor to add to crypto/pgp
	if g, ok := kr.(privGetter); ok {
		return g.GetPrivateEntity()
	}
	return nil
}

// ---- quorum.QuorumSystem decorator (quorum/quorum.go:18-29) --------------------------------------------------------
// ChooseQuorum returns the reference's own quorum object (so Nodes() and the predicates on single lists stay exactly the
// reference's) together with the flat descriptor libbftq's batch calls take.  The descriptor comes from bftq_graph_*,
// a mirror of graph.Graph that is rebuilt only when the graph changed (the reference recomputes its cliques on every
// call); bftq_graph_choose_quorum caches per (rw, version).
type QuorumSystem struct {
	inner quorum.QuorumSystem
	g     *graph.Graph
	mu    sync.Mutex
	cg    *C.bftq_graph
	fp    uint64
}

type gpuQuorum struct {
	quorum.Quorum
	qcs     []C.bftq_qc_ids_t
	members []C.uint64_t
}

func NewQuorumSystem(g *graph.Graph, _ *crypto.Crypto) quorum.QuorumSystem {
	return &QuorumSystem{inner: wotqs.New(g), g: g}
}

// fingerprint of the graph's shape: vertices, edges, self nodes and revocations (graph.go:19-24 exports all of them)
func (qs *QuorumSystem) fingerprint() uint64 {
	h := uint64(1469598103934665603)
	mix := func(v uint64) { h ^= v; h *= 1099511628211 }
	var sv, se uint64
	for id, v := range qs.g.Vertices {
		x := id * 0x9E3779B97F4A7C15
		if v.Instance != nil {
			x ^= 0x5555
		}
		sv += x
		for to := range v.Edges {
			se += (id ^ (to * 0xC2B2AE3D27D4EB4F)) * 0x165667B19E3779F9
		}
	}
	mix(sv)
	mix(se)
	mix(uint64(len(qs.g.Vertices)))
	for id := range qs.g.Revoked {
		mix(id * 31)
	}
	for _, s := range qs.g.Self {
		if s.Instance != nil {
			mix(s.Instance.Id())
		}
	}
	return h
}

// mirror rebuilds the C-side graph from graph.Graph: AddNodes per vertex with its in-edges as signers, SetSelfNodes,
// Revoke (graph.go:46-88,131-140).
func (qs *QuorumSystem) mirror() {
	if qs.cg != nil {
		C.bftq_graph_destroy(qs.cg)
	}
	C.bftq_graph_create(&qs.cg)
	for id := range qs.g.Revoked {
		C.bftq_graph_revoke(qs.cg, C.uint64_t(id))
	}
	signers := make(map[uint64][]C.uint64_t)
	for from, v := range qs.g.Vertices {
		for to := range v.Edges {
			signers[to] = append(signers[to], C.uint64_t(from))
		}
	}
	for id, v := range qs.g.Vertices {
		if v.Instance == nil {
			continue
		}
		s := signers[id]
		var p *C.uint64_t
		if len(s) > 0 {
			p = &s[0]
		}
		C.bftq_graph_add_node(qs.cg, C.uint64_t(id), p, C.uint32_t(len(s)))
	}
	for _, s := range qs.g.Self {
		if s.Instance != nil {
			C.bftq_graph_set_self(qs.cg, C.uint64_t(s.Instance.Id()))
		}
	}
}

func (qs *QuorumSystem) ChooseQuorum(rw int) quorum.Quorum {
	q := qs.inner.ChooseQuorum(rw)
	qs.mu.Lock()
	defer qs.mu.Unlock()
	if fp := qs.fingerprint(); qs.cg == nil || fp != qs.fp {
		qs.mirror()
		qs.fp = fp
	}
	var nq, nm C.uint32_t
	C.bftq_graph_choose_quorum(qs.cg, C.int(rw), nil, 0, &nq, nil, 0, &nm)
	gq := &gpuQuorum{Quorum: q, qcs: make([]C.bftq_qc_ids_t, nq), members: make([]C.uint64_t, nm)}
	if nq > 0 {
		C.bftq_graph_choose_quorum(qs.cg, C.int(rw), &gq.qcs[0], nq, &nq, idsPtr(gq.members), nm, &nm)
	}
	return gq
}
