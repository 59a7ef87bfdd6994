// bloomsearch_host.hpp — C++ host layer above the C ABI (include/bloomgpu.h).
//
// The reference is Go; its toolchain is absent from the build image, so this is the compiled
// host-side mirror of the reference's interface for the hot path: same names, argument meaning
// and error behaviour as the Go call sites it stands in for, so the self-test
// (host_selftest.cpp) reads like the reference's own tests.
//
//   BloomEntrySets::buildFilters   ingest.go:24-145   (sets stay on the host, hashing on the GPU)
//   BloomFilter / BloomFilters      file_format.go:328-332, bloom WriteTo/ReadFrom framing
//   encodeFilterSection / parseFilterSection   file_format.go:343-448
//   Field / Token / FieldToken / And / Or / AndBloomQueries / RegexFieldGuardBloomQuery
//                                   query.go:478-718
//   Corpus::evaluateBloomFilters    query_exec.go:75-159 for every unit at once
//
// All bloom arithmetic (hashing, bit set / test) happens in libbloomgpu.so; nothing here hashes
// keys or tests bits on the CPU.
#pragma once

#include <array>
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/bloomgpu.h"

namespace bloomsearch {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

// ---- value types --------------------------------------------------------------------------
struct BloomFilter {  // *bloom.BloomFilter as a value: m bits, k hashes, native-endian words
    uint64_t m = 0, k = 0;
    std::vector<uint64_t> words;
    std::vector<uint8_t> WriteTo() const;                               // [u64 BE m][u64 BE k][u64 BE m][words BE]
    static BloomFilter ReadFrom(const uint8_t* p, size_t n, size_t* used);  // throws Error on truncation
    bool Equal(const BloomFilter& o) const { return m == o.m && k == o.k && words == o.words; }
};

struct BloomFilters {  // file_format.go:328-332; an absent member cannot disqualify anything
    std::optional<BloomFilter> FieldBloomFilter, TokenBloomFilter, FieldTokenBloomFilter;
};

struct BloomEntryCounts { size_t Fields, Tokens, FieldTokens; };

std::vector<uint8_t> encodeFilterSection(const BloomFilters& f);          // file_format.go:343-385
BloomFilters parseFilterSection(const std::vector<uint8_t>& section);     // file_format.go:392-448 (footer use)
uint32_t crc32c(const uint8_t* p, size_t n);                              // file_format.go:44

// ---- query AST (query.go:478-718) -----------------------------------------------------------
enum class BloomConditionType { Field, Token, FieldToken, Unknown };
enum class BloomExpressionType { Condition, And, Or, Unknown };

struct BloomCondition {
    BloomConditionType Type;
    std::string Field, Token;
};
struct BloomExpression {
    BloomExpressionType ExpressionType = BloomExpressionType::Condition;
    std::optional<BloomCondition> Condition;
    std::vector<BloomExpression> Children;
};
struct BloomQuery { std::optional<BloomExpression> Expression; };

BloomExpression Field(const std::string& field);
BloomExpression Token(const std::string& token);
BloomExpression FieldToken(const std::string& field, const std::string& token);
BloomExpression And(std::vector<BloomExpression> expressions);  // flattens same-type children
BloomExpression Or(std::vector<BloomExpression> expressions);
std::string makeFieldTokenKey(const std::string& field, const std::string& token);  // tokenizer.go:508-511

struct RegexCondition { std::string Field, Pattern; };
struct RegexExpression {
    BloomExpressionType ExpressionType = BloomExpressionType::Condition;
    std::optional<RegexCondition> Condition;
    std::vector<RegexExpression> Children;
};
struct RegexQuery { std::optional<RegexExpression> Expression; };
std::optional<BloomQuery> RegexFieldGuardBloomQuery(const RegexQuery* q);                    // query.go:696-705
std::optional<BloomQuery> AndBloomQueries(const BloomQuery* left, const BloomQuery* right);  // query.go:707-716

struct CompiledQuery {  // what the C ABI takes
    std::vector<uint8_t> key_bytes;
    std::vector<uint64_t> key_off{0};
    std::vector<uint8_t> kinds;
    std::vector<bsg_expr_op> prog;
    bool has_program = false;
};
CompiledQuery compileBloomQuery(const BloomQuery* q);

// ---- device objects ------------------------------------------------------------------------
class Context {
public:
    explicit Context(int device = 0);
    ~Context();
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    bsg_ctx* handle() const { return h_; }
private:
    bsg_ctx* h_ = nullptr;
};

class BloomEntrySets {  // ingest.go:24-45
public:
    std::unordered_set<std::string> fields, tokens, fieldTokens;
    void addField(const std::string& path) { fields.insert(path); }
    void addToken(const std::string& token) { tokens.insert(token); }
    void addFieldToken(const std::string& path, const std::string& token) { fieldTokens.insert(path + "::" + token); }
    void unionInto(BloomEntrySets& dst) const;        // ingest.go:105-115
    BloomEntryCounts counts() const { return {fields.size(), tokens.size(), fieldTokens.size()}; }
    BloomFilters buildFilters(Context& ctx, double falsePositiveRate) const;  // ingest.go:127-133
};

// One bsg_build for all partition buffers of a flush + the file-level union filter
// (flush.go:204,221,253).  file may be null.
std::vector<BloomFilters> buildFiltersMany(Context& ctx, const std::vector<const BloomEntrySets*>& blocks,
                                           const BloomEntrySets* file, double fpr, BloomFilters* fileFilters);

// bloom.New(m,k) followed by AddString(entry) for each entry — explicit sizing, for callers (and
// tests, bloom_tree_engine_test.go:364-378) that do not size from the entry count.
BloomFilter buildBloomFilter(Context& ctx, const std::vector<std::string>& entries, uint64_t m, uint64_t k);

class Corpus {
public:
    static std::unique_ptr<Corpus> fromFilters(Context& ctx, const std::vector<BloomFilters>& units);
    // raw filter sections as they sit in a block filter region; status[u] != 0 => unit fails open
    static std::unique_ptr<Corpus> fromSections(Context& ctx, const std::vector<uint8_t>& sections,
                                                const std::vector<uint64_t>& sec_off, bool verify_crc,
                                                std::vector<int32_t>* status);
    ~Corpus();
    uint64_t units() const { return n_units_; }
    // evaluateBloomFilters (query_exec.go:75-87) for every unit: true = cannot be disqualified
    std::vector<bool> evaluateBloomFilters(const BloomQuery* q) const;
private:
    Corpus(Context& c, bsg_corpus* h, uint64_t n) : ctx_(c), h_(h), n_units_(n) {}
    Context& ctx_;
    bsg_corpus* h_;
    uint64_t n_units_;
};

}  // namespace bloomsearch
