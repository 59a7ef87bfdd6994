/* THIRD-PARTY PIN, TEST INFRASTRUCTURE ONLY: CRC32C (Castagnoli, the polynomial of Go's
 * crc32.MakeTable(crc32.Castagnoli), /root/reference/file_format.go:44,379,399) computed by the CPU's own
 * SSE4.2 `crc32` instruction — an implementation nobody here wrote.  crc32.Checksum starts from ^0 and
 * returns ^state, which is what this does.  tests/test_oracle.py compares oracle/bloomref.c's table-driven
 * bref_crc32c_sw (and the Python twin) with it; the device CRC (csrc/kernels_sections.cu) is compared with the oracle by the GPU tests. */
#include <nmmintrin.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

uint32_t crc32c_hw(const uint8_t *p, size_t n) {
    uint64_t c = 0xffffffffu;
    while (n >= 8) {
        uint64_t v;
        memcpy(&v, p, 8);
        c = _mm_crc32_u64(c, v);
        p += 8;
        n -= 8;
    }
    uint32_t c32 = (uint32_t)c;
    while (n--) c32 = _mm_crc32_u8(c32, *p++);
    return ~c32;
}
