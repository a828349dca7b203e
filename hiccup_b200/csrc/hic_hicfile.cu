// hic_hicfile.cu -- host-side helpers for the `.hic` container (no device code): the rows of a Huffman table to and
// from the byte strings the reference pickles them as.
//
// reference hiccup/hicimage.py:57-60, 103-121: a table payload is pickle.dumps({"type": TupP, "data": [pickle.dumps((symbol,
// code)) for every row]}) -- one pickle PER ROW, thousands per image, which in Python costs more host time than the
// whole encode + decode of the image on the GPU.  A row is (int | numpy.int32, str of '0'/'1'); its protocol-4 pickle is
//
//     80 04 95 <u64 frame length> <symbol> 8c <len> <code chars> 94 86 94 2e
//
// with <symbol> = 4b <u8> | 4d <u16> | 4a <i32> for a Python int, or <np_pre> <i32> <np_mid> for a numpy.int32
// scalar, np_pre / np_mid being whatever THIS environment's numpy writes around the four value bytes
// (hiccup_b200/hicimage.py reads them off a sample pickle and passes them in; it also checks this code against
// pickle.dumps on a set of samples when it loads and falls back to pickle if anything differs).
// The parser accepts exactly these canonical forms and reports the first row that is anything else.
#include <stdint.h>
#include <string.h>
#include "hic_runtime.cuh"

namespace {

constexpr uint8_t HEAD[3] = {0x80, 0x04, 0x95};
constexpr uint8_t TAIL[4] = {0x94, 0x86, 0x94, 0x2e};

inline void put_u64(uint8_t* p, uint64_t v) {
    for (int i = 0; i < 8; ++i) p[i] = (uint8_t)(v >> (8 * i));
}
inline uint64_t get_u64(const uint8_t* p) {
    uint64_t v = 0;
    for (int i = 0; i < 8; ++i) v |= (uint64_t)p[i] << (8 * i);
    return v;
}

}  // namespace

extern "C" {

int hic_hicfile_pack_rows(const int32_t* symbols, const uint8_t* lens, const uint64_t* codes, uint64_t n, const uint8_t* numpy_scalar,
                          const uint8_t* np_pre, uint32_t np_pre_len, const uint8_t* np_mid, uint32_t np_mid_len,
                          uint8_t* out, uint64_t out_capacity, uint64_t* out_off) {
    HIC_REQUIRE(symbols && lens && codes && out && out_off, "NULL argument");
    HIC_REQUIRE(!numpy_scalar || (np_pre && np_mid), "numpy scalar rows need the environment's prefix and infix");
    uint64_t pos = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t len = lens[i];
        HIC_REQUIRE(len >= 1 && len <= 58, "code length %u out of range in row %llu", len, (unsigned long long)i);
        const int32_t v = symbols[i];
        const bool is_np = numpy_scalar && numpy_scalar[i];
        uint32_t sym_bytes;
        if (is_np) sym_bytes = np_pre_len + 4 + np_mid_len;
        else sym_bytes = (v >= 0 && v < 256) ? 2 : ((v >= 256 && v < 65536) ? 3 : 5);
        const uint64_t body = sym_bytes + 2 + len + 4;
        if (pos + 11 + body > out_capacity)
            return hic::fail(HIC_ERR_CAPACITY, "row buffer of %llu bytes is too small", (unsigned long long)out_capacity);
        out_off[i] = pos;
        uint8_t* p = out + pos;
        memcpy(p, HEAD, 3);
        put_u64(p + 3, body);
        p += 11;
        if (is_np) {
            memcpy(p, np_pre, np_pre_len);
            p += np_pre_len;
            const uint32_t u = (uint32_t)v;
            p[0] = (uint8_t)u; p[1] = (uint8_t)(u >> 8); p[2] = (uint8_t)(u >> 16); p[3] = (uint8_t)(u >> 24);
            p += 4;
            memcpy(p, np_mid, np_mid_len);
            p += np_mid_len;
        } else if (sym_bytes == 2) {
            p[0] = 0x4b; p[1] = (uint8_t)v;
            p += 2;
        } else if (sym_bytes == 3) {
            p[0] = 0x4d; p[1] = (uint8_t)v; p[2] = (uint8_t)(v >> 8);
            p += 3;
        } else {
            const uint32_t u = (uint32_t)v;
            p[0] = 0x4a; p[1] = (uint8_t)u; p[2] = (uint8_t)(u >> 8); p[3] = (uint8_t)(u >> 16); p[4] = (uint8_t)(u >> 24);
            p += 5;
        }
        p[0] = 0x8c;
        p[1] = (uint8_t)len;
        p += 2;
        const uint64_t code = codes[i];
        for (uint32_t k = 0; k < len; ++k) p[k] = (uint8_t)('0' + ((code >> (len - 1 - k)) & 1u));
        p += len;
        memcpy(p, TAIL, 4);
        pos += 11 + body;
    }
    out_off[n] = pos;
    return HIC_OK;
}

int hic_hicfile_parse_rows(const uint8_t* data, const uint64_t* off, uint64_t n, const uint8_t* np_pre, uint32_t np_pre_len,
                           const uint8_t* np_mid, uint32_t np_mid_len, int32_t* symbols, uint8_t* lens, uint64_t* codes,
                           uint8_t* numpy_scalar, int64_t* bad_row) {
    HIC_REQUIRE(data && off && symbols && lens && codes && numpy_scalar && bad_row, "NULL argument");
    *bad_row = -1;
    for (uint64_t i = 0; i < n; ++i) {
        const uint8_t* p = data + off[i];
        const uint64_t size = off[i + 1] - off[i];
        bool ok = size >= 20 && memcmp(p, HEAD, 3) == 0 && get_u64(p + 3) == size - 11 && memcmp(p + size - 4, TAIL, 4) == 0;
        uint64_t q = 11;
        int32_t v = 0;
        uint8_t is_np = 0;
        if (ok) {
            const uint8_t op = p[11];
            if (op == 0x4b) { v = p[12]; q = 13; }
            else if (op == 0x4d) { v = (int32_t)(p[12] | (p[13] << 8)); q = 14; }
            else if (op == 0x4a) { v = (int32_t)((uint32_t)p[12] | ((uint32_t)p[13] << 8) | ((uint32_t)p[14] << 16) | ((uint32_t)p[15] << 24)); q = 16; }
            else if (np_pre && np_mid && size >= 11 + (uint64_t)np_pre_len + 4 + np_mid_len + 7 && memcmp(p + 11, np_pre, np_pre_len) == 0 &&
                     memcmp(p + 11 + np_pre_len + 4, np_mid, np_mid_len) == 0) {
                const uint8_t* b = p + 11 + np_pre_len;
                v = (int32_t)((uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24));
                q = 11 + np_pre_len + 4 + np_mid_len;
                is_np = 1;
            } else ok = false;
        }
        // (canonical int forms only: pickle writes the shortest one)
        if (ok && !is_np && ((p[11] == 0x4d && v < 256) || (p[11] == 0x4a && v >= 0 && v < 65536))) ok = false;
        uint32_t len = 0;
        if (ok) {
            ok = q + 2 <= size - 4 && p[q] == 0x8c;
            if (ok) {
                len = p[q + 1];
                ok = len >= 1 && len <= 58 && q + 2 + len == size - 4;
            }
        }
        uint64_t code = 0;
        if (ok) {
            for (uint32_t k = 0; k < len; ++k) {
                const uint8_t c = p[q + 2 + k];
                if (c != '0' && c != '1') { ok = false; break; }
                code = (code << 1) | (uint64_t)(c - '0');
            }
        }
        if (!ok) {
            *bad_row = (int64_t)i;
            return HIC_OK;
        }
        symbols[i] = v;
        lens[i] = (uint8_t)len;
        codes[i] = code;
        numpy_scalar[i] = is_np;
    }
    return HIC_OK;
}

}  // extern "C"
