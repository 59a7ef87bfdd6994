// bsg_corpus_load_sections — probe straight from raw on-disk filter sections
// (file_format.go:343-385 framing).  Placeholder until the device-side CRC32C /
// parse kernels land; the symbol is exported so the ABI is complete.
#include "bsg_internal.h"

extern "C" int bsg_set_last_error_internal(int code, const char* msg);

extern "C" int bsg_corpus_load_sections(bsg_ctx*, const uint8_t*, const uint64_t*, uint64_t, int, int32_t*,
                                        uint64_t*, bsg_corpus** out) {
    if (out) *out = nullptr;
    return bsg_set_last_error_internal(BSG_ERR_UNSUPPORTED, "bsg_corpus_load_sections: not implemented yet");
}
