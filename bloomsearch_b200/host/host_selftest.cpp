// host_selftest — the C++ host layer exercised the way the reference's own tests exercise the
// Go call sites.  `--host-only` needs no GPU (AST, lowering, section codec);
// without it the GPU cases run too.  Exit code 0 = all passed.  Run by tests/test_host_cpp.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>

#include "bloomsearch_host.hpp"

using namespace bloomsearch;

static int g_fail = 0;
#define EXPECT(cond)                                                          \
    do {                                                                      \
        if (!(cond)) { std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); ++g_fail; } \
    } while (0)

static std::vector<uint8_t> from_hex(const std::string& h) {
    std::vector<uint8_t> b;
    for (size_t i = 0; i + 1 < h.size(); i += 2) b.push_back(static_cast<uint8_t>(std::stoul(h.substr(i, 2), nullptr, 16)));
    return b;
}
static std::string to_hex(const std::vector<uint8_t>& b) {
    static const char* d = "0123456789abcdef";
    std::string s;
    for (uint8_t x : b) { s.push_back(d[x >> 4]); s.push_back(d[x & 15]); }
    return s;
}

static void test_ast_and_lowering() {
    // flattenExpressions, query.go:600-610
    BloomExpression e = And({Field("a"), And({Field("b"), Token("c")}), Or({Field("d"), Field("e")})});
    EXPECT(e.ExpressionType == BloomExpressionType::And && e.Children.size() == 4);
    EXPECT(e.Children[3].ExpressionType == BloomExpressionType::Or);
    EXPECT(makeFieldTokenKey("service", "auth") == "service::auth");  // tokenizer.go:508-511
    // AndBloomQueries / RegexFieldGuardBloomQuery, query.go:696-716
    BloomQuery left{Token("x")};
    EXPECT(AndBloomQueries(nullptr, &left)->Expression->Condition->Token == "x");
    BloomQuery empty;
    EXPECT(AndBloomQueries(&left, &empty)->Expression->Condition->Token == "x");
    RegexQuery rq;
    RegexExpression r_and;
    r_and.ExpressionType = BloomExpressionType::And;
    RegexExpression c1, c2;
    c1.Condition = RegexCondition{"msg", "a.*"};
    c2.Condition = RegexCondition{"lvl", "e"};
    r_and.Children = {c1, c2};
    rq.Expression = r_and;
    auto guard = RegexFieldGuardBloomQuery(&rq);
    EXPECT(guard && guard->Expression->Children.size() == 2 &&
           guard->Expression->Children[0].Condition->Type == BloomConditionType::Field);
    auto both = AndBloomQueries(&left, &*guard);
    EXPECT(both->Expression->ExpressionType == BloomExpressionType::And && both->Expression->Children.size() == 3);
    EXPECT(!RegexFieldGuardBloomQuery(nullptr));
    // lowering: leaf dedup, kinds, postfix
    BloomQuery q{And({Token("a"), Or({Token("a"), Field("a")}), FieldToken("f", "a")})};
    CompiledQuery cq = compileBloomQuery(&q);
    EXPECT(cq.kinds.size() == 3 && cq.kinds[0] == BSG_KIND_TOKEN && cq.kinds[1] == BSG_KIND_FIELD && cq.kinds[2] == BSG_KIND_FIELDTOKEN);
    EXPECT(std::string(cq.key_bytes.begin(), cq.key_bytes.end()) == "aaf::a");
    // pairwise fold: a, b, x OR 2, [AND 2], a, AND 2  (And of three children -> two AND-2 ops)
    EXPECT(cq.prog.size() == 7 && cq.prog.back().op == BSG_OP_AND && cq.prog.back().arg == 2);
    EXPECT(!compileBloomQuery(nullptr).has_program);
    BloomQuery nilcond;
    nilcond.Expression = BloomExpression{};  // Condition == nil -> TRUE (query_exec.go:97-100)
    EXPECT(compileBloomQuery(&nilcond).prog.size() == 1 && compileBloomQuery(&nilcond).prog[0].op == BSG_OP_TRUE);
    BloomQuery wide{Or(std::vector<BloomExpression>(200, Token("t")))};
    uint32_t sp = 0, mx = 0;
    for (const auto& op : compileBloomQuery(&wide).prog) {
        sp = (op.op == BSG_OP_AND || op.op == BSG_OP_OR) ? sp - op.arg + 1 : sp + 1;
        mx = sp > mx ? sp : mx;
    }
    EXPECT(sp == 1 && mx <= BSG_MAX_STACK);
}

static void test_codec(const std::string& golden_section_hex) {
    EXPECT(crc32c(reinterpret_cast<const uint8_t*>("123456789"), 9) == 0xE3069283u);
    if (golden_section_hex.empty()) return;
    const std::vector<uint8_t> sec = from_hex(golden_section_hex);
    BloomFilters f = parseFilterSection(sec);  // file_format.go:392-448
    EXPECT(f.FieldBloomFilter && f.TokenBloomFilter && !f.FieldTokenBloomFilter);
    EXPECT(f.FieldBloomFilter->m == 29 && f.FieldBloomFilter->k == 11);  // NewWithEstimates(2, 0.001)
    EXPECT(encodeFilterSection(f) == sec);                                // file_format_test.go:439-443 round trip
    std::vector<uint8_t> bad = sec;
    bad[10] ^= 0x40;
    bool threw = false;
    try { parseFilterSection(bad); } catch (const Error&) { threw = true; }
    EXPECT(threw);
    threw = false;
    try { parseFilterSection({0, 0}); } catch (const Error&) { threw = true; }
    EXPECT(threw);
}

static void test_gpu(const std::string& golden_section_hex) {
    Context ctx(0);
    // ---- TestEvaluateBloomFilters, bloom_tree_engine_test.go:357-442 ----
    uint64_t m, k;
    bsg_estimate(100, 0.01, &m, &k);
    EXPECT(m == 959 && k == 7);
    BloomFilters file;
    file.FieldBloomFilter = buildBloomFilter(ctx, {"user.name", "user.age"}, m, k);
    file.TokenBloomFilter = buildBloomFilter(ctx, {"alice", "30"}, m, k);
    file.FieldTokenBloomFilter = buildBloomFilter(ctx, {makeFieldTokenKey("user.name", "alice"), makeFieldTokenKey("user.age", "30")}, m, k);
    auto corpus = Corpus::fromFilters(ctx, {file});
    struct Case { const char* name; std::optional<BloomQuery> q; bool expected; };
    std::vector<Case> cases = {
        {"nil query should return true", std::nullopt, true},
        {"field exists should return true", BloomQuery{Field("user.name")}, true},
        {"field does not exist should return false", BloomQuery{Field("nonexistent.field")}, false},
        {"token exists should return true", BloomQuery{Token("alice")}, true},
        {"field-token exists should return true", BloomQuery{FieldToken("user.name", "alice")}, true},
        {"OR condition with one match should return true", BloomQuery{Or({Field("nonexistent.field"), Field("user.name")})}, true},
        {"AND condition with one mismatch should return false", BloomQuery{And({Field("nonexistent.field"), Field("user.name")})}, false},
        {"multiple groups with OR combinator should return true", BloomQuery{Or({Field("nonexistent.field"), FieldToken("user.name", "alice")})}, true},
        {"empty OR is false", BloomQuery{Or({})}, false},
        {"empty AND is true", BloomQuery{And({})}, true},
    };
    for (auto& c : cases) {
        const bool got = corpus->evaluateBloomFilters(c.q ? &*c.q : nullptr)[0];
        if (got != c.expected) { std::fprintf(stderr, "FAIL case '%s': got %d\n", c.name, got); ++g_fail; }
    }
    // nil filters cannot disqualify (query_exec.go:137-151)
    BloomFilters partial;
    partial.TokenBloomFilter = file.TokenBloomFilter;
    auto c2 = Corpus::fromFilters(ctx, {partial});
    BloomQuery qf{Field("nonexistent.field")};
    EXPECT(c2->evaluateBloomFilters(&qf)[0]);
    BloomQuery qt{Token("nonexistent")};
    EXPECT(!c2->evaluateBloomFilters(&qt)[0]);

    // ---- sizing from distinct counts, file_format_test.go:28-94: counts {2,101,101} ----
    BloomEntrySets es;
    es.addField("a"); es.addField("b");
    for (int i = 0; i < 101; ++i) { es.addToken("t" + std::to_string(i)); es.addFieldToken("a", "t" + std::to_string(i)); }
    EXPECT(es.counts().Fields == 2 && es.counts().Tokens == 101 && es.counts().FieldTokens == 101);
    BloomFilters bf = es.buildFilters(ctx, 0.001);
    EXPECT(bf.FieldBloomFilter->m == 29 && bf.FieldBloomFilter->k == 11);
    EXPECT(bf.TokenBloomFilter->m == 1453 && bf.TokenBloomFilter->k == 10 && bf.FieldTokenBloomFilter->m == 1453);
    // empty set sized for one entry and testing negative (ingest.go:135-138)
    BloomEntrySets none;
    BloomFilters ef = none.buildFilters(ctx, 0.001);
    EXPECT(ef.TokenBloomFilter->m == 15 && ef.TokenBloomFilter->k == 11);
    bool any = false;
    for (uint64_t w : ef.TokenBloomFilter->words) any |= w != 0;
    EXPECT(!any);

    // ---- example_test.go:18-86: 2 rows, 1 block; FieldToken("service","auth") keeps the block ----
    BloomEntrySets ex;
    const char* rows[2][3][2] = {{{"id", "1"}, {"service", "auth"}, {"message", "login timeout for user"}},
                                 {{"id", "2"}, {"service", "payment"}, {"message", "charge succeeded"}}};
    for (auto& row : rows)
        for (auto& kv : row) {
            ex.addField(kv[0]);
            std::string text = kv[1], word;
            for (size_t i = 0; i <= text.size(); ++i) {
                if (i == text.size() || text[i] == ' ') {
                    if (!word.empty()) { ex.addToken(word); ex.addFieldToken(kv[0], word); }
                    word.clear();
                } else word.push_back(text[i]);
            }
        }
    BloomEntrySets fileUnion;
    ex.unionInto(fileUnion);
    BloomFilters fileF;
    std::vector<BloomFilters> blocks = buildFiltersMany(ctx, {&ex}, &fileUnion, 0.001, &fileF);
    EXPECT(blocks[0].TokenBloomFilter->Equal(*fileF.TokenBloomFilter));  // one block => file filter == block filter
    // write the block's section, read it back through the device-side section loader
    std::vector<uint8_t> sec = encodeFilterSection(blocks[0]);
    std::vector<int32_t> status;
    auto c3 = Corpus::fromSections(ctx, sec, {0, sec.size()}, true, &status);
    EXPECT(status.size() == 1 && status[0] == 0);
    BloomQuery hit{FieldToken("service", "auth")}, miss{FieldToken("service", "billing")};
    EXPECT(c3->evaluateBloomFilters(&hit)[0]);
    EXPECT(!c3->evaluateBloomFilters(&miss)[0]);
    // corrupt: the block is an error, never a candidate (query_exec.go:580-590 records the error and continues
    // without scanning the block); status says why
    sec[sec.size() / 2] ^= 0x08;
    auto c4 = Corpus::fromSections(ctx, sec, {0, sec.size()}, true, &status);
    EXPECT(status[0] == -2 && !c4->evaluateBloomFilters(&miss)[0] && !c4->evaluateBloomFilters(&hit)[0]);

    // ---- cross-language golden: same entries as tests/golden (oracle-generated) section ----
    BloomEntrySets g1, g2;
    g1.addField("service"); g1.addField("id");
    g2.addToken("auth"); g2.addToken("1");
    BloomFilters gf;
    gf.FieldBloomFilter = g1.buildFilters(ctx, 0.001).FieldBloomFilter;
    gf.TokenBloomFilter = g2.buildFilters(ctx, 0.001).TokenBloomFilter;
    const std::string hex = to_hex(encodeFilterSection(gf));
    std::printf("section_hex %s\n", hex.c_str());
    if (!golden_section_hex.empty()) EXPECT(hex == golden_section_hex);
}

int main(int argc, char** argv) {
    bool host_only = false;
    std::string golden;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--host-only")) host_only = true;
        else if (!std::strcmp(argv[i], "--golden-section") && i + 1 < argc) golden = argv[++i];
    }
    try {
        test_ast_and_lowering();
        test_codec(golden);
        if (!host_only) test_gpu(golden);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "EXCEPTION: %s\n", e.what());
        return 2;
    }
    if (g_fail) { std::fprintf(stderr, "%d failure(s)\n", g_fail); return 1; }
    std::printf("host_selftest ok (%s)\n", host_only ? "host only" : "host + gpu");
    return 0;
}
