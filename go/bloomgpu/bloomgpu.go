// Package bloomgpu is the cgo binding of libbloomgpu.so (include/bloomgpu.h) that a
// bloomsearch maintainer would vendor next to the engine to route the bloom build /
// probe hot path to a B200.  It is SOURCE ONLY in this repository: the build image
// has no Go toolchain, so the same C ABI is exercised from tests/ through ctypes.
//
// Call sites it replaces in danthegoodman1/bloomsearch @ 10735cf9 (see INTEGRATION.md):
//
//	ingest.go:127-145   buildFilters / buildSizedBloomFilter  -> BuildFilters
//	query_exec.go:75-159 evaluateBloomFilters (file + block)  -> Corpus.Probe
//	file_format.go:392-448 parseFilterSection per block/query -> LoadSections (once)
//
// Go pointers are never retained by C after a call returns (cgo rule); every
// []byte / []uint64 passed below is pinned only for the duration of the call.
package bloomgpu

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../bloomsearch_b200/_build -lbloomgpu
#include <stdlib.h>
#include "bloomgpu.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"runtime"
	"unsafe"
)

// Kind selects which of a unit's three filters a key is tested against (query.go:478-484).
type Kind uint8

const (
	KindField      Kind = C.BSG_KIND_FIELD
	KindToken      Kind = C.BSG_KIND_TOKEN
	KindFieldToken Kind = C.BSG_KIND_FIELDTOKEN
)

// FilterDesc mirrors bsg_filter_desc: m bits, k hashes, offset of the words (uint64 units).
type FilterDesc struct {
	M, K, WordOff uint64
}

// Op mirrors bsg_expr_op (postfix BloomExpression).
type Op struct {
	Op, Arg uint32
}

const (
	OpLeaf  = C.BSG_OP_LEAF
	OpAnd   = C.BSG_OP_AND
	OpOr    = C.BSG_OP_OR
	OpTrue  = C.BSG_OP_TRUE
	OpFalse = C.BSG_OP_FALSE
)

// check turns a status into an error.  bsg_last_error() is a THREAD-LOCAL detail string, and a goroutine may
// be rescheduled onto another OS thread between two cgo calls, so every entry point below pins its goroutine
// with pin() for the duration of "call + check".
func check(rc C.int) error {
	if rc == C.BSG_OK {
		return nil
	}
	return fmt.Errorf("bloomgpu: %s: %s", C.GoString(C.bsg_strerror(rc)), C.GoString(C.bsg_last_error()))
}

func pin() func() {
	runtime.LockOSThread()
	return runtime.UnlockOSThread
}

// Context owns one GPU (bsg_ctx).  Safe for concurrent use by many goroutines.
type Context struct{ h *C.bsg_ctx }

func NewContext(device int) (*Context, error) {
	defer pin()()
	var h *C.bsg_ctx
	if err := check(C.bsg_create(C.int(device), &h)); err != nil {
		return nil, err
	}
	c := &Context{h}
	runtime.SetFinalizer(c, func(c *Context) { c.Close() })
	return c, nil
}

func (c *Context) Close() {
	if c.h != nil {
		C.bsg_destroy(c.h)
		c.h = nil
	}
}

// PackedKeys is the ABI's key layout: bytes back to back + n+1 offsets.
type PackedKeys struct {
	Bytes []byte
	Off   []uint64
}

func Pack(keys []string) PackedKeys {
	p := PackedKeys{Off: make([]uint64, 1, len(keys)+1)}
	for _, k := range keys {
		p.Bytes = append(p.Bytes, k...) // []byte(string): no normalisation (row_matcher.go:228-231)
		p.Off = append(p.Off, uint64(len(p.Bytes)))
	}
	if len(p.Bytes) == 0 {
		p.Bytes = []byte{0}
	}
	return p
}

// Build is bsg_build: groups of keys -> filters described by desc, all in one launch.
// groupFilter2 may be nil; NoFilter marks "no secondary (file-level) filter".
const NoFilter = uint32(C.BSG_NO_FILTER)

func (c *Context) Build(keys PackedKeys, groupBegin []uint64, groupFilter, groupFilter2 []uint32,
	desc []FilterDesc, nWords uint64) ([]uint64, error) {
	out := make([]uint64, nWords+1)
	var gf2 *C.uint32_t
	if groupFilter2 != nil {
		gf2 = (*C.uint32_t)(unsafe.Pointer(&groupFilter2[0]))
	}
	if len(groupFilter) == 0 || len(desc) == 0 {
		return out[:nWords], nil
	}
	rc := C.bsg_build(c.h, (*C.uint8_t)(unsafe.Pointer(&keys.Bytes[0])), (*C.uint64_t)(unsafe.Pointer(&keys.Off[0])),
		C.uint64_t(len(keys.Off)-1), (*C.uint64_t)(unsafe.Pointer(&groupBegin[0])), C.uint32_t(len(groupFilter)),
		(*C.uint32_t)(unsafe.Pointer(&groupFilter[0])), gf2,
		(*C.bsg_filter_desc)(unsafe.Pointer(&desc[0])), C.uint32_t(len(desc)),
		(*C.uint64_t)(unsafe.Pointer(&out[0])), C.uint64_t(nWords))
	runtime.KeepAlive(keys)
	return out[:nWords], check(rc)
}

// Corpus is a set of units (data blocks, or files) whose filters are resident in HBM.
type Corpus struct {
	h     *C.bsg_corpus
	Units uint64
}

// LoadSections uploads raw filter sections exactly as they sit in a file's block filter
// region (file_format.go:343-385); framing, CRC32C and big-endian decode run on the GPU.
// status[u] != 0 marks a section that failed to parse.  As in the reference (query_exec.go:580-590: the
// error is recorded, the loop continues, the block is NOT scanned) such a unit never survives a probe —
// its mask bit is always 0 — and the caller records status[u] as that block's error.
func (c *Context) LoadSections(sections []byte, secOff []uint64, verifyCRC bool) (*Corpus, []int32, error) {
	defer pin()()
	n := uint64(len(secOff) - 1)
	status := make([]int32, n+1)
	var h *C.bsg_corpus
	var bad C.uint64_t
	v := C.int(0)
	if verifyCRC {
		v = 1
	}
	var sp *C.uint8_t
	if len(sections) > 0 {
		sp = (*C.uint8_t)(unsafe.Pointer(&sections[0]))
	}
	rc := C.bsg_corpus_load_sections(c.h, sp, (*C.uint64_t)(unsafe.Pointer(&secOff[0])), C.uint64_t(n), v,
		(*C.int32_t)(unsafe.Pointer(&status[0])), &bad, &h)
	if err := check(rc); err != nil {
		return nil, nil, err
	}
	cp := &Corpus{h, n}
	runtime.SetFinalizer(cp, func(cp *Corpus) { cp.Close() })
	return cp, status[:n], nil
}

// Load uploads already-decoded filters (desc[3*u+kind], native-endian words).
func (c *Context) Load(desc []FilterDesc, words []uint64) (*Corpus, error) {
	defer pin()()
	if len(desc)%3 != 0 {
		return nil, errors.New("bloomgpu: desc must hold 3 slots per unit")
	}
	var h *C.bsg_corpus
	var wp *C.uint64_t
	if len(words) > 0 {
		wp = (*C.uint64_t)(unsafe.Pointer(&words[0]))
	}
	var dp *C.bsg_filter_desc
	if len(desc) > 0 {
		dp = (*C.bsg_filter_desc)(unsafe.Pointer(&desc[0]))
	}
	rc := C.bsg_corpus_load(c.h, dp, C.uint64_t(len(desc)/3), wp, C.uint64_t(len(words)), 0, &h)
	if err := check(rc); err != nil {
		return nil, err
	}
	cp := &Corpus{h, uint64(len(desc) / 3)}
	runtime.SetFinalizer(cp, func(cp *Corpus) { cp.Close() })
	return cp, nil
}

func (cp *Corpus) Close() {
	if cp.h != nil {
		C.bsg_corpus_free(cp.h)
		cp.h = nil
	}
}

// Probe is bsg_probe: every key against every unit; prog (nil = no expression: every unit
// survives, query_exec.go:81-83) folds the leaf bits into the candidate mask.
// mask bit u = unit u survives; matrix (optional) row u, bit q = TestString(key q) on unit u.
func (c *Context) Probe(cp *Corpus, keys PackedKeys, kinds []Kind, prog []Op, wantMatrix bool) (mask []uint64, matrix []uint64, err error) {
	defer pin()()
	n := uint32(len(keys.Off) - 1)
	mask = make([]uint64, (cp.Units+63)/64+1)
	var mp *C.uint64_t
	if wantMatrix {
		matrix = make([]uint64, cp.Units*uint64((n+63)/64)+1)
		mp = (*C.uint64_t)(unsafe.Pointer(&matrix[0]))
	}
	var pp *C.bsg_expr_op
	if len(prog) > 0 {
		pp = (*C.bsg_expr_op)(unsafe.Pointer(&prog[0]))
	}
	var kp *C.uint8_t
	if n > 0 {
		kp = (*C.uint8_t)(unsafe.Pointer(&kinds[0]))
	}
	rc := C.bsg_probe(c.h, cp.h, (*C.uint8_t)(unsafe.Pointer(&keys.Bytes[0])), (*C.uint64_t)(unsafe.Pointer(&keys.Off[0])),
		C.uint32_t(n), kp, pp, C.uint32_t(len(prog)), mp, (*C.uint64_t)(unsafe.Pointer(&mask[0])))
	runtime.KeepAlive(keys)
	return mask[:(cp.Units+63)/64], matrix, check(rc)
}

// QuerySpec is one member of a ProbeMulti call: its keys (already packed), their kinds and its postfix
// program, whose leaf arguments index ITS OWN keys.
type QuerySpec struct {
	Keys  []string
	Kinds []Kind
	Prog  []Op
}

// ProbeMulti is bsg_probe_multi: several queries in ONE pass over the corpus (the staged kernels stream every
// filter byte once per pass of up to 1 024 keys, whatever the number of keys), one candidate mask per query —
// identical to len(qs) Probe calls.  masks[j] bit u = unit u survives query j.
func (c *Context) ProbeMulti(cp *Corpus, qs []QuerySpec) (masks [][]uint64, err error) {
	defer pin()()
	if len(qs) == 0 {
		return nil, nil
	}
	var all []string
	var kinds []Kind
	var progs []Op
	qBegin := make([]uint32, 1, len(qs)+1)
	pBegin := make([]uint32, 1, len(qs)+1)
	for _, q := range qs {
		all = append(all, q.Keys...)
		kinds = append(kinds, q.Kinds...)
		progs = append(progs, q.Prog...)
		qBegin = append(qBegin, uint32(len(all)))
		pBegin = append(pBegin, uint32(len(progs)))
	}
	keys := Pack(all)
	words := (cp.Units + 63) / 64
	flat := make([]uint64, uint64(len(qs))*words+1)
	var pp *C.bsg_expr_op
	if len(progs) > 0 {
		pp = (*C.bsg_expr_op)(unsafe.Pointer(&progs[0]))
	}
	var kp *C.uint8_t
	if len(kinds) > 0 {
		kp = (*C.uint8_t)(unsafe.Pointer(&kinds[0]))
	}
	rc := C.bsg_probe_multi(c.h, cp.h, (*C.uint8_t)(unsafe.Pointer(&keys.Bytes[0])), (*C.uint64_t)(unsafe.Pointer(&keys.Off[0])),
		C.uint32_t(len(all)), kp, C.uint32_t(len(qs)), (*C.uint32_t)(unsafe.Pointer(&qBegin[0])), pp,
		(*C.uint32_t)(unsafe.Pointer(&pBegin[0])), (*C.uint64_t)(unsafe.Pointer(&flat[0])))
	runtime.KeepAlive(keys)
	if err = check(rc); err != nil {
		return nil, err
	}
	masks = make([][]uint64, len(qs))
	for j := range qs {
		masks[j] = flat[uint64(j)*words : uint64(j+1)*words]
	}
	return masks, nil
}

// Batcher is bsg_batcher: goroutines that query the same resident corpus at the same time (one goroutine per
// query in the reference, query_exec.go:201-433) are merged into bsg_probe_multi launches by a group commit —
// a caller that finds no launch in flight launches at once, callers that arrive meanwhile share the next one.
type Batcher struct {
	h  *C.bsg_batcher
	cp *Corpus
}

func (c *Context) NewBatcher(cp *Corpus, maxKeys, maxQueries, windowMicros uint32) (*Batcher, error) {
	defer pin()()
	b := &Batcher{cp: cp}
	if err := check(C.bsg_batcher_create(c.h, cp.h, C.uint32_t(maxKeys), C.uint32_t(maxQueries), C.uint32_t(windowMicros), &b.h)); err != nil {
		return nil, err
	}
	return b, nil
}

func (b *Batcher) Close() {
	if b.h != nil {
		C.bsg_batcher_destroy(b.h)
		b.h = nil
	}
}

// Probe blocks until this query's candidate mask is ready; it may have shared its launch with other goroutines.
func (b *Batcher) Probe(keys PackedKeys, kinds []Kind, prog []Op) (mask []uint64, err error) {
	defer pin()()
	n := uint32(len(keys.Off) - 1)
	mask = make([]uint64, (b.cp.Units+63)/64+1)
	var pp *C.bsg_expr_op
	if len(prog) > 0 {
		pp = (*C.bsg_expr_op)(unsafe.Pointer(&prog[0]))
	}
	var kp *C.uint8_t
	if n > 0 {
		kp = (*C.uint8_t)(unsafe.Pointer(&kinds[0]))
	}
	rc := C.bsg_batcher_probe(b.h, (*C.uint8_t)(unsafe.Pointer(&keys.Bytes[0])), (*C.uint64_t)(unsafe.Pointer(&keys.Off[0])),
		C.uint32_t(n), kp, pp, C.uint32_t(len(prog)), (*C.uint64_t)(unsafe.Pointer(&mask[0])))
	runtime.KeepAlive(keys)
	return mask[:(b.cp.Units+63)/64], check(rc)
}

// SetParents records, for every unit (block) of cp, the index of its file in the files corpus;
// ProbeHierarchical then runs both pruning stages of Query (query_exec.go:399-406 and :572-615)
// in one call, compacting the surviving blocks on the device between them.
func (c *Context) SetParents(cp *Corpus, parent []uint32, nParentUnits uint64) error {
	defer pin()()
	if len(parent) == 0 {
		return nil
	}
	return check(C.bsg_corpus_set_parents(c.h, cp.h, (*C.uint32_t)(unsafe.Pointer(&parent[0])), C.uint64_t(len(parent)),
		C.uint64_t(nParentUnits)))
}

func (c *Context) ProbeHierarchical(files, blocks *Corpus, keys PackedKeys, kinds []Kind, prog []Op) (fileMask, blockMask []uint64, err error) {
	defer pin()()
	n := uint32(len(keys.Off) - 1)
	fileMask = make([]uint64, (files.Units+63)/64+1)
	blockMask = make([]uint64, (blocks.Units+63)/64+1)
	var pp *C.bsg_expr_op
	if len(prog) > 0 {
		pp = (*C.bsg_expr_op)(unsafe.Pointer(&prog[0]))
	}
	var kp *C.uint8_t
	if n > 0 {
		kp = (*C.uint8_t)(unsafe.Pointer(&kinds[0]))
	}
	rc := C.bsg_probe_hierarchical(c.h, files.h, blocks.h, (*C.uint8_t)(unsafe.Pointer(&keys.Bytes[0])),
		(*C.uint64_t)(unsafe.Pointer(&keys.Off[0])), C.uint32_t(n), kp, pp, C.uint32_t(len(prog)),
		(*C.uint64_t)(unsafe.Pointer(&fileMask[0])), (*C.uint64_t)(unsafe.Pointer(&blockMask[0])))
	runtime.KeepAlive(keys)
	return fileMask[:(files.Units+63)/64], blockMask[:(blocks.Units+63)/64], check(rc)
}

// CountDistinct is bsg_count_distinct: exact distinct counts per group and (if groupParent != nil)
// per parent union of emissions that may repeat — the counts the maps of bloomEntrySets provide
// (ingest.go:24-45,105-123), i.e. the n that sizes each filter (ingest.go:139-140).
func (c *Context) CountDistinct(keys PackedKeys, groupBegin []uint64, groupParent []uint32, nParents int) (groups, parents []uint64, err error) {
	defer pin()()
	nGroups := len(groupBegin) - 1
	groups = make([]uint64, nGroups+1)
	if nGroups <= 0 {
		return groups[:0], nil, nil
	}
	var gp *C.uint32_t
	var pc *C.uint64_t
	if groupParent != nil {
		parents = make([]uint64, nParents+1)
		gp = (*C.uint32_t)(unsafe.Pointer(&groupParent[0]))
		pc = (*C.uint64_t)(unsafe.Pointer(&parents[0]))
	}
	var kb *C.uint8_t
	if len(keys.Bytes) > 0 {
		kb = (*C.uint8_t)(unsafe.Pointer(&keys.Bytes[0]))
	}
	rc := C.bsg_count_distinct(c.h, kb, (*C.uint64_t)(unsafe.Pointer(&keys.Off[0])), C.uint64_t(len(keys.Off)-1),
		(*C.uint64_t)(unsafe.Pointer(&groupBegin[0])), C.uint32_t(nGroups), gp, C.uint32_t(nParents),
		(*C.uint64_t)(unsafe.Pointer(&groups[0])), pc)
	runtime.KeepAlive(keys)
	if parents != nil {
		parents = parents[:nParents]
	}
	return groups[:nGroups], parents, check(rc)
}

// BuildFieldTokens is bsg_build_fieldtokens: field::token entries as (path, token) index pairs into
// one string table; the joined key is hashed on the GPU and never materialised (ingest.go:95-102).
func (c *Context) BuildFieldTokens(strings PackedKeys, pairPath, pairToken []uint32, groupBegin []uint64,
	groupFilter, groupFilter2 []uint32, desc []FilterDesc, nWords uint64) ([]uint64, error) {
	out := make([]uint64, nWords+1)
	if len(groupFilter) == 0 || len(desc) == 0 || len(pairPath) == 0 {
		return out[:nWords], nil
	}
	var gf2 *C.uint32_t
	if groupFilter2 != nil {
		gf2 = (*C.uint32_t)(unsafe.Pointer(&groupFilter2[0]))
	}
	rc := C.bsg_build_fieldtokens(c.h, (*C.uint8_t)(unsafe.Pointer(&strings.Bytes[0])), (*C.uint64_t)(unsafe.Pointer(&strings.Off[0])),
		C.uint64_t(len(strings.Off)-1), (*C.uint32_t)(unsafe.Pointer(&pairPath[0])), (*C.uint32_t)(unsafe.Pointer(&pairToken[0])),
		C.uint64_t(len(pairPath)), (*C.uint64_t)(unsafe.Pointer(&groupBegin[0])), C.uint32_t(len(groupFilter)),
		(*C.uint32_t)(unsafe.Pointer(&groupFilter[0])), gf2, (*C.bsg_filter_desc)(unsafe.Pointer(&desc[0])),
		C.uint32_t(len(desc)), (*C.uint64_t)(unsafe.Pointer(&out[0])), C.uint64_t(nWords))
	runtime.KeepAlive(strings)
	return out[:nWords], check(rc)
}

// Cache is the resident filter cache (bsg_cache): corpora keyed by file id under a byte budget with
// least-recently-used eviction, pinned while a query uses them, invalidated when a merge replaces the
// file (merge.go:529-536) or it is tombstoned.  The reference decodes filters per query and drops them
// (query_exec.go:399-412, :572-615); this is what lets the GPU path keep them in HBM instead.
type Cache struct {
	h *C.bsg_cache
	c *Context
}

func (c *Context) NewCache(budgetBytes uint64) (*Cache, error) {
	defer pin()()
	var h *C.bsg_cache
	if err := check(C.bsg_cache_create(c.h, C.uint64_t(budgetBytes), &h)); err != nil {
		return nil, err
	}
	return &Cache{h, c}, nil
}

func (k *Cache) Close() {
	if k.h != nil {
		C.bsg_cache_destroy(k.h)
		k.h = nil
	}
}

// Acquire returns the pinned corpus of fileID, or nil on a miss.  Release it when the query is done with it;
// never Close it.
func (k *Cache) Acquire(fileID uint64) (*Corpus, error) {
	defer pin()()
	var h *C.bsg_corpus
	if err := check(C.bsg_cache_acquire(k.h, C.uint64_t(fileID), &h)); err != nil || h == nil {
		return nil, err
	}
	return &Corpus{h, uint64(C.bsg_corpus_units(h))}, nil
}

// InsertSections loads the file's raw filter sections (as LoadSections) into the cache and returns the pinned corpus.
func (k *Cache) InsertSections(fileID uint64, sections []byte, secOff []uint64, verifyCRC bool) (*Corpus, []int32, error) {
	defer pin()()
	n := uint64(len(secOff) - 1)
	status := make([]int32, n+1)
	var h *C.bsg_corpus
	var bad C.uint64_t
	v := C.int(0)
	if verifyCRC {
		v = 1
	}
	var sp *C.uint8_t
	if len(sections) > 0 {
		sp = (*C.uint8_t)(unsafe.Pointer(&sections[0]))
	}
	rc := C.bsg_cache_insert_sections(k.h, C.uint64_t(fileID), sp, (*C.uint64_t)(unsafe.Pointer(&secOff[0])), C.uint64_t(n), v,
		(*C.int32_t)(unsafe.Pointer(&status[0])), &bad, &h)
	if err := check(rc); err != nil {
		return nil, nil, err
	}
	return &Corpus{h, n}, status[:n], nil
}

// Insert hands an already loaded corpus (e.g. the file-level filters of one MetaStore generation) to the cache.
func (k *Cache) Insert(id uint64, cp *Corpus) (*Corpus, error) {
	defer pin()()
	var h *C.bsg_corpus
	if err := check(C.bsg_cache_insert(k.h, C.uint64_t(id), cp.h, &h)); err != nil {
		return nil, err
	}
	runtime.SetFinalizer(cp, nil)
	cp.h = nil // owned by the cache now
	return &Corpus{h, uint64(C.bsg_corpus_units(h))}, nil
}

func (k *Cache) Release(cp *Corpus) {
	if cp != nil && cp.h != nil {
		C.bsg_cache_release(k.h, cp.h)
		cp.h = nil
	}
}

// Invalidate: the file was replaced by a merge or tombstoned; its filters are never served again (a query
// that still holds them finishes on the old copy, which is freed on its last Release).
func (k *Cache) Invalidate(fileID uint64) error {
	defer pin()()
	return check(C.bsg_cache_invalidate(k.h, C.uint64_t(fileID)))
}
