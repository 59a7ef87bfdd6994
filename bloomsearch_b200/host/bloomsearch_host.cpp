// bloomsearch_host.cpp — see bloomsearch_host.hpp.  Host logic only; every hash / bit operation
// is a call into libbloomgpu.so.
#include "bloomsearch_host.hpp"

#include <cstring>
#include <unordered_map>

namespace bloomsearch {

static void check(int rc, const char* what) {
    if (rc != BSG_OK) throw Error(rc, std::string(what) + ": " + bsg_strerror(rc) + ": " + bsg_last_error());
}

// ------------------------------------------------------------------------------ codec ---
static void put_be64(std::vector<uint8_t>& b, uint64_t v) {
    for (int i = 7; i >= 0; --i) b.push_back(static_cast<uint8_t>(v >> (8 * i)));
}
static uint64_t get_be64(const uint8_t* p) {
    uint64_t v = 0;
    for (int i = 0; i < 8; ++i) v = (v << 8) | p[i];
    return v;
}
static void put_le32(std::vector<uint8_t>& b, uint32_t v) {
    for (int i = 0; i < 4; ++i) b.push_back(static_cast<uint8_t>(v >> (8 * i)));
}
static uint32_t get_le32(const uint8_t* p) {
    return static_cast<uint32_t>(p[0]) | static_cast<uint32_t>(p[1]) << 8 | static_cast<uint32_t>(p[2]) << 16 |
           static_cast<uint32_t>(p[3]) << 24;
}

uint32_t crc32c(const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int j = 0; j < 8; ++j) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xff] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

std::vector<uint8_t> BloomFilter::WriteTo() const {
    std::vector<uint8_t> b;
    b.reserve(24 + 8 * words.size());
    put_be64(b, m);
    put_be64(b, k);
    put_be64(b, m);
    for (uint64_t w : words) put_be64(b, w);
    return b;
}

BloomFilter BloomFilter::ReadFrom(const uint8_t* p, size_t n, size_t* used) {
    if (n < 24) throw Error(BSG_ERR_FORMAT, "bloom filter header truncated");
    BloomFilter f;
    f.m = get_be64(p);
    f.k = get_be64(p + 8);
    const uint64_t bitlen = get_be64(p + 16);
    const uint64_t nw = (bitlen + 63) >> 6;
    if (n < 24 + 8 * nw) throw Error(BSG_ERR_FORMAT, "bloom filter words truncated");
    f.words.resize(nw);
    for (uint64_t i = 0; i < nw; ++i) f.words[i] = get_be64(p + 24 + 8 * i);
    if (used) *used = 24 + 8 * nw;
    return f;
}

std::vector<uint8_t> encodeFilterSection(const BloomFilters& f) {
    const std::optional<BloomFilter>* slots[3] = {&f.FieldBloomFilter, &f.TokenBloomFilter, &f.FieldTokenBloomFilter};
    uint8_t flags = 0;
    for (int i = 0; i < 3; ++i)
        if (slots[i]->has_value()) flags |= static_cast<uint8_t>(1u << i);
    std::vector<uint8_t> b{flags};
    for (int i = 0; i < 3; ++i) {
        if (!slots[i]->has_value()) continue;
        const std::vector<uint8_t> raw = (*slots[i])->WriteTo();
        put_le32(b, static_cast<uint32_t>(raw.size()));
        b.insert(b.end(), raw.begin(), raw.end());
    }
    put_le32(b, crc32c(b.data(), b.size()));
    return b;
}

BloomFilters parseFilterSection(const std::vector<uint8_t>& s) {
    if (s.size() < 5) throw Error(BSG_ERR_FORMAT, "bloom filter section too small");
    const size_t plen = s.size() - 4;
    if (crc32c(s.data(), plen) != get_le32(s.data() + plen)) throw Error(BSG_ERR_FORMAT, "invalid hash");
    const uint8_t flags = s[0];
    if (flags & ~7u) throw Error(BSG_ERR_FORMAT, "unrecognized bloom filter section");
    BloomFilters out;
    std::optional<BloomFilter>* slots[3] = {&out.FieldBloomFilter, &out.TokenBloomFilter, &out.FieldTokenBloomFilter};
    size_t pos = 1;
    for (int i = 0; i < 3; ++i) {
        if (!(flags & (1u << i))) continue;
        if (plen - pos < 4) throw Error(BSG_ERR_FORMAT, "truncated bloom filter length prefix");
        const uint32_t len = get_le32(s.data() + pos);
        pos += 4;
        if (len > plen - pos) throw Error(BSG_ERR_FORMAT, "bloom filter length exceeds section remainder");
        *slots[i] = BloomFilter::ReadFrom(s.data() + pos, len, nullptr);
        pos += len;
    }
    if (pos != plen) throw Error(BSG_ERR_FORMAT, "bloom filter section has trailing bytes");
    return out;
}

// -------------------------------------------------------------------------------- AST ---
BloomExpression Field(const std::string& field) {
    BloomExpression e;
    e.Condition = BloomCondition{BloomConditionType::Field, field, ""};
    return e;
}
BloomExpression Token(const std::string& token) {
    BloomExpression e;
    e.Condition = BloomCondition{BloomConditionType::Token, "", token};
    return e;
}
BloomExpression FieldToken(const std::string& field, const std::string& token) {
    BloomExpression e;
    e.Condition = BloomCondition{BloomConditionType::FieldToken, field, token};
    return e;
}
static BloomExpression nary(BloomExpressionType t, std::vector<BloomExpression> ex) {
    BloomExpression e;
    e.ExpressionType = t;
    for (auto& c : ex) {  // flattenExpressions, query.go:600-610
        if (c.ExpressionType == t && !c.Condition) {
            for (auto& g : c.Children) e.Children.push_back(std::move(g));
        } else {
            e.Children.push_back(std::move(c));
        }
    }
    return e;
}
BloomExpression And(std::vector<BloomExpression> ex) { return nary(BloomExpressionType::And, std::move(ex)); }
BloomExpression Or(std::vector<BloomExpression> ex) { return nary(BloomExpressionType::Or, std::move(ex)); }
std::string makeFieldTokenKey(const std::string& field, const std::string& token) { return field + "::" + token; }

static std::optional<BloomExpression> regexToField(const RegexExpression& e) {  // query.go:651-694
    switch (e.ExpressionType) {
    case BloomExpressionType::Condition:
        if (!e.Condition) return std::nullopt;
        return Field(e.Condition->Field);
    case BloomExpressionType::And:
    case BloomExpressionType::Or: {
        BloomExpression out;
        out.ExpressionType = e.ExpressionType;
        for (const auto& c : e.Children)
            if (auto f = regexToField(c)) out.Children.push_back(std::move(*f));
        return out;
    }
    default: return std::nullopt;
    }
}
std::optional<BloomQuery> RegexFieldGuardBloomQuery(const RegexQuery* q) {
    if (!q || !q->Expression) return std::nullopt;
    auto e = regexToField(*q->Expression);
    if (!e) return std::nullopt;
    return BloomQuery{std::move(*e)};
}
std::optional<BloomQuery> AndBloomQueries(const BloomQuery* left, const BloomQuery* right) {
    if (!left || !left->Expression) return right ? std::optional<BloomQuery>(*right) : std::nullopt;
    if (!right || !right->Expression) return *left;
    return BloomQuery{And({*left->Expression, *right->Expression})};
}

CompiledQuery compileBloomQuery(const BloomQuery* q) {
    CompiledQuery cq;
    if (!q || !q->Expression) return cq;  // no expression: every unit survives (query_exec.go:81-83)
    cq.has_program = true;
    std::unordered_map<std::string, uint32_t> index;
    auto leaf = [&](uint8_t kind, const std::string& key) -> uint32_t {
        const std::string id = std::string(1, static_cast<char>('0' + kind)) + key;
        auto it = index.find(id);
        if (it != index.end()) return it->second;
        const uint32_t ix = static_cast<uint32_t>(cq.kinds.size());
        index.emplace(id, ix);
        cq.key_bytes.insert(cq.key_bytes.end(), key.begin(), key.end());
        cq.key_off.push_back(cq.key_bytes.size());
        cq.kinds.push_back(kind);
        return ix;
    };
    struct Emit {
        CompiledQuery& cq;
        decltype(leaf)& lf;
        void operator()(const BloomExpression& e) {
            switch (e.ExpressionType) {
            case BloomExpressionType::Condition:
                if (!e.Condition) { cq.prog.push_back({BSG_OP_TRUE, 0}); return; }
                switch (e.Condition->Type) {
                case BloomConditionType::Field: cq.prog.push_back({BSG_OP_LEAF, lf(BSG_KIND_FIELD, e.Condition->Field)}); return;
                case BloomConditionType::Token: cq.prog.push_back({BSG_OP_LEAF, lf(BSG_KIND_TOKEN, e.Condition->Token)}); return;
                case BloomConditionType::FieldToken:
                    cq.prog.push_back({BSG_OP_LEAF, lf(BSG_KIND_FIELDTOKEN, makeFieldTokenKey(e.Condition->Field, e.Condition->Token))});
                    return;
                default: cq.prog.push_back({BSG_OP_FALSE, 0}); return;
                }
            case BloomExpressionType::And:
            case BloomExpressionType::Or: {
                const uint32_t op = e.ExpressionType == BloomExpressionType::And ? BSG_OP_AND : BSG_OP_OR;
                // fold pairwise (child, child, op 2, child, op 2, ...): the stack grows by one per nesting level
                uint32_t n = 0;
                for (const auto& c : e.Children) {
                    (*this)(c);
                    if (++n >= 2) cq.prog.push_back({op, 2});
                }
                if (n <= 1) cq.prog.push_back({op, n});
                return;
            }
            default: cq.prog.push_back({BSG_OP_FALSE, 0}); return;
            }
        }
    } emit{cq, leaf};
    emit(*q->Expression);
    if (cq.key_bytes.empty()) cq.key_bytes.push_back(0);
    return cq;
}

// ------------------------------------------------------------------------------ device ---
Context::Context(int device) { check(bsg_create(device, &h_), "bsg_create"); }
Context::~Context() { bsg_destroy(h_); }

void BloomEntrySets::unionInto(BloomEntrySets& dst) const {
    dst.fields.insert(fields.begin(), fields.end());
    dst.tokens.insert(tokens.begin(), tokens.end());
    dst.fieldTokens.insert(fieldTokens.begin(), fieldTokens.end());
}

std::vector<BloomFilters> buildFiltersMany(Context& ctx, const std::vector<const BloomEntrySets*>& blocks,
                                           const BloomEntrySets* file, double fpr, BloomFilters* fileFilters) {
    std::vector<uint8_t> bytes;
    std::vector<uint64_t> off{0}, group_begin{0};
    std::vector<uint32_t> gf, gf2;
    std::vector<bsg_filter_desc> desc;
    uint64_t word_off = 0;
    auto new_filter = [&](size_t n) -> uint32_t {
        uint64_t m, k;
        bsg_estimate(n > 1 ? n : 1, fpr, &m, &k);  // ingest.go:139-140: empty set sized for one entry
        desc.push_back({m, k, word_off});
        word_off += (m + 63) / 64;
        return static_cast<uint32_t>(desc.size() - 1);
    };
    uint32_t file_ids[3] = {BSG_NO_FILTER, BSG_NO_FILTER, BSG_NO_FILTER};
    if (file) {
        file_ids[0] = new_filter(file->fields.size());
        file_ids[1] = new_filter(file->tokens.size());
        file_ids[2] = new_filter(file->fieldTokens.size());
    }
    std::vector<std::array<uint32_t, 3>> block_ids(blocks.size());
    for (size_t b = 0; b < blocks.size(); ++b) {
        const std::unordered_set<std::string>* sets[3] = {&blocks[b]->fields, &blocks[b]->tokens, &blocks[b]->fieldTokens};
        for (int kind = 0; kind < 3; ++kind) {
            const uint32_t id = new_filter(sets[kind]->size());
            block_ids[b][kind] = id;
            for (const auto& e : *sets[kind]) {
                bytes.insert(bytes.end(), e.begin(), e.end());
                off.push_back(bytes.size());
            }
            group_begin.push_back(off.size() - 1);
            gf.push_back(id);
            gf2.push_back(file_ids[kind]);
        }
    }
    if (bytes.empty()) bytes.push_back(0);
    std::vector<uint64_t> words(word_off ? word_off : 1);
    check(bsg_build(ctx.handle(), bytes.data(), off.data(), off.size() - 1, group_begin.data(),
                    static_cast<uint32_t>(gf.size()), gf.data(), file ? gf2.data() : nullptr, desc.data(),
                    static_cast<uint32_t>(desc.size()), words.data(), word_off),
          "bsg_build");
    auto mk = [&](uint32_t id) {
        BloomFilter f;
        f.m = desc[id].m;
        f.k = desc[id].k;
        f.words.assign(words.begin() + desc[id].word_off, words.begin() + desc[id].word_off + (f.m + 63) / 64);
        return f;
    };
    std::vector<BloomFilters> out(blocks.size());
    for (size_t b = 0; b < blocks.size(); ++b) {
        out[b].FieldBloomFilter = mk(block_ids[b][0]);
        out[b].TokenBloomFilter = mk(block_ids[b][1]);
        out[b].FieldTokenBloomFilter = mk(block_ids[b][2]);
    }
    if (file && fileFilters) {
        fileFilters->FieldBloomFilter = mk(file_ids[0]);
        fileFilters->TokenBloomFilter = mk(file_ids[1]);
        fileFilters->FieldTokenBloomFilter = mk(file_ids[2]);
    }
    return out;
}

BloomFilter buildBloomFilter(Context& ctx, const std::vector<std::string>& entries, uint64_t m, uint64_t k) {
    std::vector<uint8_t> bytes;
    std::vector<uint64_t> off{0};
    for (const auto& e : entries) {
        bytes.insert(bytes.end(), e.begin(), e.end());
        off.push_back(bytes.size());
    }
    if (bytes.empty()) bytes.push_back(0);
    BloomFilter f;
    f.m = m < 1 ? 1 : m;  // bloom.New clamps to >= 1
    f.k = k < 1 ? 1 : k;
    f.words.assign((f.m + 63) / 64, 0);
    const uint64_t group_begin[2] = {0, entries.size()};
    const uint32_t gf[1] = {0};
    const bsg_filter_desc d{f.m, f.k, 0};
    check(bsg_build(ctx.handle(), bytes.data(), off.data(), entries.size(), group_begin, 1, gf, nullptr, &d, 1,
                    f.words.data(), f.words.size()),
          "bsg_build");
    return f;
}

BloomFilters BloomEntrySets::buildFilters(Context& ctx, double fpr) const {
    return buildFiltersMany(ctx, {this}, nullptr, fpr, nullptr)[0];
}

std::unique_ptr<Corpus> Corpus::fromFilters(Context& ctx, const std::vector<BloomFilters>& units) {
    std::vector<bsg_filter_desc> desc(units.size() * 3, bsg_filter_desc{0, 0, 0});
    std::vector<uint64_t> words;
    for (size_t u = 0; u < units.size(); ++u) {
        const std::optional<BloomFilter>* slots[3] = {&units[u].FieldBloomFilter, &units[u].TokenBloomFilter,
                                                      &units[u].FieldTokenBloomFilter};
        for (int kind = 0; kind < 3; ++kind) {
            if (!slots[kind]->has_value()) continue;
            const BloomFilter& f = **slots[kind];
            desc[u * 3 + kind] = {f.m, f.k, words.size()};
            words.insert(words.end(), f.words.begin(), f.words.end());
        }
    }
    bsg_corpus* h = nullptr;
    check(bsg_corpus_load(ctx.handle(), desc.data(), units.size(), words.empty() ? nullptr : words.data(), words.size(), 0, &h),
          "bsg_corpus_load");
    return std::unique_ptr<Corpus>(new Corpus(ctx, h, units.size()));
}

std::unique_ptr<Corpus> Corpus::fromSections(Context& ctx, const std::vector<uint8_t>& sections,
                                             const std::vector<uint64_t>& sec_off, bool verify_crc,
                                             std::vector<int32_t>* status) {
    const uint64_t n = sec_off.size() - 1;
    std::vector<int32_t> st(n ? n : 1);
    uint64_t bad = 0;
    bsg_corpus* h = nullptr;
    check(bsg_corpus_load_sections(ctx.handle(), sections.empty() ? nullptr : sections.data(), sec_off.data(), n,
                                   verify_crc ? 1 : 0, st.data(), &bad, &h),
          "bsg_corpus_load_sections");
    st.resize(n);
    if (status) *status = st;
    return std::unique_ptr<Corpus>(new Corpus(ctx, h, n));
}

Corpus::~Corpus() { bsg_corpus_free(h_); }

std::vector<bool> Corpus::evaluateBloomFilters(const BloomQuery* q) const {
    const CompiledQuery cq = compileBloomQuery(q);
    std::vector<uint64_t> mask((n_units_ + 63) / 64 + 1);
    const uint8_t dummy = 0;
    check(bsg_probe(ctx_.handle(), h_, cq.key_bytes.empty() ? &dummy : cq.key_bytes.data(), cq.key_off.data(),
                    static_cast<uint32_t>(cq.kinds.size()), cq.kinds.empty() ? &dummy : cq.kinds.data(),
                    cq.has_program ? cq.prog.data() : nullptr, static_cast<uint32_t>(cq.prog.size()), nullptr, mask.data()),
          "bsg_probe");
    std::vector<bool> out(n_units_);
    for (uint64_t u = 0; u < n_units_; ++u) out[u] = (mask[u >> 6] >> (u & 63)) & 1;
    return out;
}

}  // namespace bloomsearch
