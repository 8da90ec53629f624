// Package gpu swaps bftkv's signature verification for libbftq.so (B200).
package gpu

/*
#cgo CFLAGS: -I${SRCDIR}/../../../bftq/include
#cgo LDFLAGS: -L${SRCDIR}/../../../bftq/bftkv_b200 -lbftq -Wl,-rpath,${SRCDIR}/../../../bftq/bftkv_b200
#include <stdlib.h>
#include "bftq.h"
*/
import "C"

import (
	"sync"
	"time"
	"unsafe"

	"github.com/yahoo/bftkv/crypto"
	"github.com/yahoo/bftkv/crypto/pgp"
	"github.com/yahoo/bftkv/node"
	"github.com/yahoo/bftkv/packet"
	"github.com/yahoo/bftkv/quorum"
)

type engine struct {
	e  *C.bftq_engine
	kr *C.bftq_keyring
	agg *aggregator // coalesces concurrent single verifies into batches
}

// New returns pgp.New() with the verification half of Signature / CollectiveSignature on the GPU.
// Wiring point: cmd/bftkv/main.go:66, api/api.go:37, protocol/test_utils/test_utils.go:35,47.
func New(device int) *crypto.Crypto {
	c := pgp.New()
	var e *C.bftq_engine
	if rc := C.bftq_init(C.int(device), &e); rc != 0 {
		panic("bftq: " + C.GoString(C.bftq_last_error())) // no CPU fallback by design
	}
	g := &engine{e: e}
	C.bftq_keyring_create(e, &g.kr)
	g.agg = newAggregator(g, 200*time.Microsecond, 16384)
	c.Keyring = &keyring{Keyring: c.Keyring, g: g}                     // mirrors Register/Remove into bftq_keyring
	sig := &signature{Signature: c.Signature, g: g}
	c.Signature = sig                                                   // Sign/Certs/Issuer stay in pgp
	c.CollectiveSignature = &collective{CollectiveSignature: c.CollectiveSignature, sig: sig, g: g}
	return c
}

// ---- keyring mirror (crypto/crypto.go:35-41) -------------------------------------------------
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
		p := C.CBytes(pkt)
		C.bftq_keyring_add(k.g.kr, (*C.uint8_t)(p), C.uint64_t(len(pkt)), boolInt(priv), nil)
		C.free(p)
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

// ---- Signature (crypto/crypto.go:50-58) --------------------------------------------------------
type signature struct {
	crypto.Signature
	g *engine
}

// Verify replaces PGPSignature.Verify (crypto_pgp.go:319-330).  The call blocks until the batch it
// was coalesced into has been verified; many goroutines (transport.Multicast's one-per-peer workers,
// transport/transport.go:110-127, and net/http's one-per-request handlers) share one batch.
func (s *signature) Verify(tbs []byte, sig *packet.SignaturePacket) error {
	if sig == nil {
		return crypto.ErrInvalidSignature
	}
	return s.g.agg.verify(tbs, sig.Data, nil)
}

func (s *signature) VerifyWithCertificate(tbs []byte, sig *packet.SignaturePacket, cert node.Node) error {
	c, err := cert.Serialize()
	if err != nil {
		return crypto.ErrInvalidSignature
	}
	return s.g.agg.verify(tbs, sig.Data, c)
}

// ---- CollectiveSignature (crypto/crypto.go:66-71) ----------------------------------------------
type collective struct {
	crypto.CollectiveSignature
	sig *signature
	g   *engine
}

func (cs *collective) Verify(tbs []byte, ss *packet.SignaturePacket, q quorum.Quorum) error {
	qcs, members := describe(q) // see "Quorum descriptors" below
	var rc C.int32_t
	tb, to := blob(tbs)
	sb, so := blob(ss.Data)
	r := C.bftq_collective_verify_batch(cs.g.kr, qcsPtr(qcs), C.uint32_t(len(qcs)), idsPtr(members), C.uint32_t(len(members)),
		tb, to, sb, so, 1, &rc)
	if r != 0 || rc != 0 {
		return crypto.ErrInsufficientNumberOfSignatures
	}
	ss.Completed = true // crypto_pgp.go:494
	return nil
}

func (cs *collective) Combine(ss *packet.SignaturePacket, s *packet.SignaturePacket, q quorum.Quorum) bool {
	if ss.Type == packet.SignatureTypeNil { // crypto_pgp.go:507-512
		ss.Type = s.Type
	} else if ss.Type != s.Type {
		return false
	}
	ss.Data = append(ss.Data, s.Data...)
	qcs, members := describe(q)
	var out C.int32_t
	sb := C.CBytes(ss.Data)
	defer C.free(sb)
	C.bftq_collective_combine_sufficient(cs.g.kr, qcsPtr(qcs), C.uint32_t(len(qcs)), idsPtr(members), C.uint32_t(len(members)),
		(*C.uint8_t)(sb), C.uint64_t(len(ss.Data)), &out)
	return out != 0
}

// ---- aggregator: where "tens of thousands of tuples" come from ----------------------------------
// (libbftq also ships this coalescer natively: C.bftq_aggregator_verify(agg, tbs, sig, cert) blocks the calling
//  goroutine's thread until its batch is verified; the Go version below avoids pinning an OS thread per call.)
type job struct {
	tbs, sig, cert []byte
	done           chan error
}
type aggregator struct {
	g     *engine
	mu    sync.Mutex
	queue []*job
	kick  chan struct{}
	wait  time.Duration
	max   int
}

func newAggregator(g *engine, wait time.Duration, max int) *aggregator {
	a := &aggregator{g: g, kick: make(chan struct{}, 1), wait: wait, max: max}
	go a.loop()
	return a
}
func (a *aggregator) verify(tbs, sig, cert []byte) error {
	j := &job{tbs, sig, cert, make(chan error, 1)}
	a.mu.Lock()
	a.queue = append(a.queue, j)
	full := len(a.queue) >= a.max
	a.mu.Unlock()
	if full {
		select { case a.kick <- struct{}{}: default: }
	}
	return <-j.done
}
func (a *aggregator) loop() {
	t := time.NewTicker(a.wait) // deadline- or size-triggered flush
	for {
		select { case <-t.C: case <-a.kick: }
		a.mu.Lock()
		batch := a.queue
		a.queue = nil
		a.mu.Unlock()
		if len(batch) > 0 {
			a.flush(batch)
		}
	}
}
func (a *aggregator) flush(batch []*job) {
	// concatenate tbs / sig / cert into three blobs with (n+1) offsets (C memory: cgo pointer rules),
	// split into with-cert and without-cert halves, then
	//   C.bftq_signature_verify_batch(a.g.kr, tbsBlob, tbsOff, sigBlob, sigOff, n, &errs[0])
	//   C.bftq_signature_verify_with_cert_batch(..., certBlob, certOff, n, &errs[0])
	// and for each job:  j.done <- (errs[i] == 0 ? nil : crypto.ErrInvalidSignature)
}
