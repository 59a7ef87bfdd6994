// engine_shim.go — the glue a maintainer adds INSIDE package bloomsearch (shown here in
// package bloomgpu's directory only so it travels with the binding; it references the
// engine's unexported types and therefore compiles only in the engine's package).
// SOURCE ONLY: no Go toolchain in the build image.  Every function names the reference
// line it replaces.
//
//go:build ignore

package bloomsearch

import (
	"sync"
	"github.com/bits-and-blooms/bloom/v3"

	"github.com/danthegoodman1/bloomsearch/bloomgpu"
)

// gpu is set by NewBloomSearchEngine when BloomSearchEngineConfig.GPUDevice >= 0.
var gpu *bloomgpu.Context

// ---------------------------------------------------------------------------
// BUILD — replaces bloomEntrySets.buildFilters (ingest.go:127-133) for ALL partition
// buffers of one flush (flush.go:138-282) plus the file-level union filter
// (flush.go:221,253) in ONE bsg_build call.  (m,k) still come from the bloom library's
// float64 formula on the host, exactly as ingest.go:139-140 does.
// ---------------------------------------------------------------------------
func buildFlushFiltersGPU(blocks []*bloomEntrySets, file *bloomEntrySets, fpr float64) (blockFilters []BloomFilters, fileFilters BloomFilters, err error) {
	var keys []string
	groupBegin := []uint64{0}
	var groupFilter, groupFilter2 []uint32
	var desc []bloomgpu.FilterDesc
	wordOff := uint64(0)
	newFilter := func(n int) uint32 {
		m, k := bloom.EstimateParameters(uint(max(n, 1)), fpr) // ingest.go:140
		desc = append(desc, bloomgpu.FilterDesc{M: uint64(max(m, 1)), K: uint64(max(k, 1)), WordOff: wordOff})
		wordOff += (uint64(max(m, 1)) + 63) / 64
		return uint32(len(desc) - 1)
	}
	fileIDs := [3]uint32{newFilter(len(file.fields)), newFilter(len(file.tokens)), newFilter(len(file.fieldTokens))}
	blockIDs := make([][3]uint32, len(blocks))
	for b, es := range blocks {
		for kind, set := range []map[string]struct{}{es.fields, es.tokens, es.fieldTokens} {
			id := newFilter(len(set))
			blockIDs[b][kind] = id
			for entry := range set { // ingest.go:141-143: the AddString loop moves to the GPU
				keys = append(keys, entry)
			}
			groupBegin = append(groupBegin, uint64(len(keys)))
			groupFilter = append(groupFilter, id)
			groupFilter2 = append(groupFilter2, fileIDs[kind]) // unionInto(fileEntries), flush.go:221
		}
	}
	words, err := gpu.Build(bloomgpu.Pack(keys), groupBegin, groupFilter, groupFilter2, desc, wordOff)
	if err != nil {
		return nil, BloomFilters{}, err
	}
	mk := func(id uint32) *bloom.BloomFilter {
		d := desc[id]
		return bloom.FromWithM(words[d.WordOff:d.WordOff+(d.M+63)/64], uint(d.M), uint(d.K)) // same words, no copy
	}
	for b := range blocks {
		blockFilters = append(blockFilters, BloomFilters{mk(blockIDs[b][0]), mk(blockIDs[b][1]), mk(blockIDs[b][2])})
	}
	fileFilters = BloomFilters{mk(fileIDs[0]), mk(fileIDs[1]), mk(fileIDs[2])}
	return
}

// ---------------------------------------------------------------------------
// PROBE — replaces the per-block loop of evaluateBlockFilters (query_exec.go:572-615) and the file-level
// test of the file stage (query_exec.go:399-406).  The reference decodes every filter per query and drops
// it again; here both levels stay RESIDENT in HBM in two caches:
//
//	blockCache  file pointer id -> the file's block filters (loaded from the raw <= 4 MiB chunks the
//	            blockFilterCursor already reads, file_format.go:618-662; parse + CRC + decode on the GPU)
//	fileCache   MetaStore generation -> the file-level filters of every current file as one corpus
//
// A merge commit (merge.go:529-536: MetaStore.Update adds the merged file and tombstones its sources) and
// DataStore.TombstoneFile call InvalidateFile for every retired file id and bump the generation.
// ---------------------------------------------------------------------------
var (
	blockCache *bloomgpu.Cache // gpu.NewCache(budget) at engine start
	fileCache  *bloomgpu.Cache
	blockErrMu sync.Mutex
	blockErr   = map[uint64][]int32{} // per-block parse status of the cached files (query_exec.go:580-590)
)

// evaluateBlockFiltersGPU: blocks of one file.  readSections is only called on a cache miss.
// keep[u] = block u is a candidate; perBlockErr[u] != 0 = its filter section failed to parse: the caller records
// the error for that block and does NOT scan it (keep[u] is false for such a block, as in the reference).
func evaluateBlockFiltersGPU(fileID uint64, readSections func() ([]byte, []uint64, error), q *BloomQuery) (keep []bool, perBlockErr []int32, err error) {
	corpus, err := blockCache.Acquire(fileID)
	if err != nil {
		return nil, nil, err
	}
	if corpus == nil { // miss: one read of the filter region, one upload; later queries of this file skip both
		sections, secOff, rerr := readSections()
		if rerr != nil {
			return nil, nil, rerr
		}
		var status []int32
		if corpus, status, err = blockCache.InsertSections(fileID, sections, secOff, true); err != nil {
			return nil, nil, err
		}
		blockErrMu.Lock()
		blockErr[fileID] = status
		blockErrMu.Unlock()
	}
	defer blockCache.Release(corpus)
	keys, kinds, prog := compileBloomQuery(q) // postfix lowering, see bloomsearch_b200/query.py:compile_bloom_query
	mask, _, err := gpu.Probe(corpus, bloomgpu.Pack(keys), kinds, prog, false)
	if err != nil {
		return nil, nil, err
	}
	keep = make([]bool, corpus.Units)
	for u := range keep {
		keep[u] = mask[u/64]>>(uint(u)%64)&1 == 1 // BloomFilterSkipped = !keep[u] (query_exec.go:599-606)
	}
	blockErrMu.Lock()
	perBlockErr = blockErr[fileID]
	blockErrMu.Unlock()
	return keep, perBlockErr, nil
}

// evaluateFileFiltersGPU: the file stage's bloom test for every candidate file of one MetaStore generation in
// one probe.  loadFiles is only called when that generation's file-level corpus is not resident yet.
func evaluateFileFiltersGPU(generation uint64, loadFiles func() (desc []bloomgpu.FilterDesc, words []uint64), q *BloomQuery) ([]bool, error) {
	corpus, err := fileCache.Acquire(generation)
	if err != nil {
		return nil, err
	}
	if corpus == nil {
		desc, words := loadFiles()
		fresh, lerr := gpu.Load(desc, words)
		if lerr != nil {
			return nil, lerr
		}
		if corpus, err = fileCache.Insert(generation, fresh); err != nil {
			return nil, err
		}
	}
	defer fileCache.Release(corpus)
	keys, kinds, prog := compileBloomQuery(q)
	// every concurrent Query() of this generation probes the SAME resident corpus: the batcher merges the ones
	// that arrive together into one bsg_probe_multi launch (group commit; an idle engine adds no latency)
	mask, err := fileBatcher(generation, corpus).Probe(bloomgpu.Pack(keys), kinds, prog)
	if err != nil {
		return nil, err
	}
	keep := make([]bool, corpus.Units)
	for u := range keep {
		keep[u] = mask[u/64]>>(uint(u)%64)&1 == 1
	}
	return keep, nil
}

// One batcher per resident file-level corpus (= per MetaStore generation); it pins the corpus for its lifetime.
var (
	batcherMu   sync.Mutex
	batchers    = map[uint64]*bloomgpu.Batcher{}
	batcherPins = map[uint64]*bloomgpu.Corpus{}
)

func fileBatcher(generation uint64, corpus *bloomgpu.Corpus) *bloomgpu.Batcher {
	batcherMu.Lock()
	defer batcherMu.Unlock()
	if b, ok := batchers[generation]; ok {
		return b
	}
	pinned, _ := fileCache.Acquire(generation) // a second pin, held until the generation is retired
	b, err := gpu.NewBatcher(corpus, 0, 0, 0)
	if err != nil || pinned == nil {
		panic("bloomgpu: batcher for a resident corpus") // both are programming errors: the caller holds a pin
	}
	batchers[generation], batcherPins[generation] = b, pinned
	return b
}

// InvalidateFile is called from the merge commit and from TombstoneFile for every retired file, with the
// generation the file-level corpus was built for.
func InvalidateFile(fileID, oldGeneration uint64) {
	batcherMu.Lock()
	if b, ok := batchers[oldGeneration]; ok { // queries in flight finish first: Close waits for the running launch
		b.Close()
		fileCache.Release(batcherPins[oldGeneration])
		delete(batchers, oldGeneration)
		delete(batcherPins, oldGeneration)
	}
	batcherMu.Unlock()
	_ = blockCache.Invalidate(fileID)
	_ = fileCache.Invalidate(oldGeneration)
	blockErrMu.Lock()
	delete(blockErr, fileID)
	blockErrMu.Unlock()
}

// compileBloomQuery lowers a BloomExpression tree (query.go:505-509) to distinct leaf keys,
// their kinds and a postfix program with the exact semantics of query_exec.go:89-159:
// nil Condition -> TRUE, unknown types -> FALSE, OR [] -> false, AND [] -> true,
// FieldToken keys joined by makeFieldTokenKey (tokenizer.go:508-511).
func compileBloomQuery(q *BloomQuery) (keys []string, kinds []bloomgpu.Kind, prog []bloomgpu.Op) {
	if q == nil || q.Expression == nil {
		return nil, nil, nil
	}
	index := map[string]uint32{}
	leaf := func(kind bloomgpu.Kind, key string) uint32 {
		id := string(rune('0'+kind)) + key
		if i, ok := index[id]; ok {
			return i
		}
		index[id] = uint32(len(keys))
		keys = append(keys, key)
		kinds = append(kinds, kind)
		return uint32(len(keys) - 1)
	}
	var emit func(e *BloomExpression)
	emit = func(e *BloomExpression) {
		switch {
		case e == nil:
			prog = append(prog, bloomgpu.Op{Op: bloomgpu.OpTrue})
		case e.ExpressionType == BloomExpressionCondition:
			c := e.Condition
			switch {
			case c == nil:
				prog = append(prog, bloomgpu.Op{Op: bloomgpu.OpTrue})
			case c.Type == BloomField:
				prog = append(prog, bloomgpu.Op{Op: bloomgpu.OpLeaf, Arg: leaf(bloomgpu.KindField, c.Field)})
			case c.Type == BloomToken:
				prog = append(prog, bloomgpu.Op{Op: bloomgpu.OpLeaf, Arg: leaf(bloomgpu.KindToken, c.Token)})
			case c.Type == BloomFieldToken:
				prog = append(prog, bloomgpu.Op{Op: bloomgpu.OpLeaf, Arg: leaf(bloomgpu.KindFieldToken, makeFieldTokenKey(c.Field, c.Token))})
			default:
				prog = append(prog, bloomgpu.Op{Op: bloomgpu.OpFalse})
			}
		case e.ExpressionType == BloomExpressionAnd || e.ExpressionType == BloomExpressionOr:
			op := uint32(bloomgpu.OpAnd)
			if e.ExpressionType == BloomExpressionOr {
				op = bloomgpu.OpOr
			}
			// fold pairwise (child, child, op 2, child, op 2, ...): the stack grows by one per nesting level,
			// not per child, so BSG_MAX_STACK = 64 covers any width
			for i := range e.Children {
				emit(&e.Children[i])
				if i >= 1 {
					prog = append(prog, bloomgpu.Op{Op: op, Arg: 2})
				}
			}
			if len(e.Children) <= 1 {
				prog = append(prog, bloomgpu.Op{Op: op, Arg: uint32(len(e.Children))})
			}
		default:
			prog = append(prog, bloomgpu.Op{Op: bloomgpu.OpFalse})
		}
	}
	emit(q.Expression)
	return
}
