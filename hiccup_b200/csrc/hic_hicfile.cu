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
// hic_hicfile_pack_table writes the whole table payload -- the outer pickle with its frames, batches and memo opcodes
// around the rows -- so that a table costs one call and one bytes object instead of one per row.
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

// One row pickle at p (the caller has checked the room); returns its size.  row_size() is the same arithmetic without
// the writes.
struct RowFormat {
    const uint8_t* np_pre; uint32_t np_pre_len;
    const uint8_t* np_mid; uint32_t np_mid_len;
};

inline uint32_t symbol_bytes(int32_t v, bool is_np, const RowFormat& f) {
    if (is_np) return f.np_pre_len + 4 + f.np_mid_len;
    return (v >= 0 && v < 256) ? 2 : ((v >= 256 && v < 65536) ? 3 : 5);
}

inline uint64_t row_size(int32_t v, uint32_t len, bool is_np, const RowFormat& f) {
    return 11 + (uint64_t)symbol_bytes(v, is_np, f) + 2 + len + 4;
}

inline uint64_t write_row(uint8_t* p, int32_t v, uint32_t len, uint64_t code, bool is_np, const RowFormat& f) {
    const uint32_t sym_bytes = symbol_bytes(v, is_np, f);
    const uint64_t body = (uint64_t)sym_bytes + 2 + len + 4;
    memcpy(p, HEAD, 3);
    put_u64(p + 3, body);
    p += 11;
    const uint32_t u = (uint32_t)v;
    if (is_np) {
        memcpy(p, f.np_pre, f.np_pre_len);
        p += f.np_pre_len;
        p[0] = (uint8_t)u; p[1] = (uint8_t)(u >> 8); p[2] = (uint8_t)(u >> 16); p[3] = (uint8_t)(u >> 24);
        p += 4;
        memcpy(p, f.np_mid, f.np_mid_len);
        p += f.np_mid_len;
    } else if (sym_bytes == 2) {
        p[0] = 0x4b; p[1] = (uint8_t)v;
        p += 2;
    } else if (sym_bytes == 3) {
        p[0] = 0x4d; p[1] = (uint8_t)v; p[2] = (uint8_t)(v >> 8);
        p += 3;
    } else {
        p[0] = 0x4a; p[1] = (uint8_t)u; p[2] = (uint8_t)(u >> 8); p[3] = (uint8_t)(u >> 16); p[4] = (uint8_t)(u >> 24);
        p += 5;
    }
    p[0] = 0x8c;
    p[1] = (uint8_t)len;
    p += 2;
    for (uint32_t k = 0; k < len; ++k) p[k] = (uint8_t)('0' + ((code >> (len - 1 - k)) & 1u));
    p += len;
    memcpy(p, TAIL, 4);
    return 11 + body;
}

// The framing of a protocol-4 pickle as CPython's pickler does it (Modules/_pickle.c, the same rules as
// Lib/pickle.py's _Framer): after the PROTO opcode every write goes into a frame, opened on demand with a 9-byte
// header; at the start of every object's save() a frame that has reached 64 KiB is closed (its length written into
// the header, or the header dropped if the frame holds fewer than 4 bytes), and the last one is closed after STOP.
struct FramedWriter {
    static constexpr uint64_t FRAME_HEADER = 9, FRAME_MIN = 4, FRAME_TARGET = 64 * 1024;
    uint8_t* out;
    uint64_t cap, pos = 0;
    int64_t frame_start = -1;
    bool overflow = false;

    FramedWriter(uint8_t* o, uint64_t c) : out(o), cap(c) {}
    uint8_t* reserve(uint64_t n) {                        // room for n bytes of pickle data (opens a frame if none is open)
        const uint64_t need = n + (frame_start < 0 ? FRAME_HEADER : 0);
        if (pos + need > cap) { overflow = true; return nullptr; }
        if (frame_start < 0) {
            frame_start = (int64_t)pos;
            pos += FRAME_HEADER;
        }
        uint8_t* p = out + pos;
        pos += n;
        return p;
    }
    void byte(uint8_t b) { if (uint8_t* p = reserve(1)) *p = b; }
    void bytes(const uint8_t* s, uint64_t n) { if (uint8_t* p = reserve(n)) memcpy(p, s, n); }
    void commit() {
        if (frame_start < 0) return;
        uint8_t* q = out + frame_start;
        const uint64_t len = pos - (uint64_t)frame_start - FRAME_HEADER;
        if (len >= FRAME_MIN) {
            q[0] = 0x95;
            put_u64(q + 1, len);
        } else {
            memmove(q, q + FRAME_HEADER, len);
            pos -= FRAME_HEADER;
        }
        frame_start = -1;
    }
    void object_boundary() {                              // start of a save()
        if (frame_start >= 0 && pos - (uint64_t)frame_start - FRAME_HEADER >= FRAME_TARGET) commit();
    }
};

// The reverse of write_row: accepts exactly the canonical forms.
inline bool parse_row(const uint8_t* p, uint64_t size, const RowFormat& f, int32_t* sym, uint8_t* len_out, uint64_t* code_out, uint8_t* np_out) {
    if (!(size >= 20 && memcmp(p, HEAD, 3) == 0 && get_u64(p + 3) == size - 11 && memcmp(p + size - 4, TAIL, 4) == 0)) return false;
    uint64_t q;
    int32_t v;
    uint8_t is_np = 0;
    const uint8_t op = p[11];
    if (op == 0x4b) { v = p[12]; q = 13; }
    else if (op == 0x4d) { v = (int32_t)(p[12] | (p[13] << 8)); q = 14; }
    else if (op == 0x4a) { v = (int32_t)((uint32_t)p[12] | ((uint32_t)p[13] << 8) | ((uint32_t)p[14] << 16) | ((uint32_t)p[15] << 24)); q = 16; }
    else if (f.np_pre && f.np_mid && size >= 11 + (uint64_t)f.np_pre_len + 4 + f.np_mid_len + 7 && memcmp(p + 11, f.np_pre, f.np_pre_len) == 0 &&
             memcmp(p + 11 + f.np_pre_len + 4, f.np_mid, f.np_mid_len) == 0) {
        const uint8_t* b = p + 11 + f.np_pre_len;
        v = (int32_t)((uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24));
        q = 11 + f.np_pre_len + 4 + f.np_mid_len;
        is_np = 1;
    } else return false;
    // (canonical int forms only: pickle writes the shortest one)
    if (!is_np && ((op == 0x4d && v < 256) || (op == 0x4a && v >= 0 && v < 65536))) return false;
    if (!(q + 2 <= size - 4 && p[q] == 0x8c)) return false;
    const uint32_t len = p[q + 1];
    if (!(len >= 1 && len <= 58 && q + 2 + len == size - 4)) return false;
    uint64_t code = 0;
    for (uint32_t k = 0; k < len; ++k) {
        const uint8_t c = p[q + 2 + k];
        if (c != '0' && c != '1') return false;
        code = (code << 1) | (uint64_t)(c - '0');
    }
    *sym = v;
    *len_out = (uint8_t)len;
    *code_out = code;
    *np_out = is_np;
    return true;
}

// A cursor over a table payload.  FRAME opcodes (0x95 + u64) carry no data of their own: they are stepped over wherever
// an opcode is expected, after checking that the frame they announce fits what is left.
struct TableReader {
    const uint8_t* data;
    uint64_t size, pos;
    bool skip_frames() {
        while (pos < size && data[pos] == 0x95) {
            if (pos + 9 > size || get_u64(data + pos + 1) > size - pos - 9) return false;
            pos += 9;
        }
        return true;
    }
    bool opcode(uint8_t op) {
        if (!skip_frames() || pos >= size || data[pos] != op) return false;
        ++pos;
        return true;
    }
    bool expect(const uint8_t* s, uint64_t n) {
        for (uint64_t i = 0; i < n; ++i)
            if (!opcode(s[i])) return false;
        return true;
    }
    bool short_string(const uint8_t** text, uint32_t* n) {   // SHORT_BINUNICODE <u8> chars MEMOIZE
        if (!opcode(0x8c) || pos >= size) return false;
        *n = data[pos];
        if (pos + 1 + *n > size) return false;
        *text = data + pos + 1;
        pos += 1 + *n;
        return opcode(0x94);
    }
};

}  // namespace

extern "C" {

int hic_hicfile_pack_rows(const int32_t* symbols, const uint8_t* lens, const uint64_t* codes, uint64_t n, const uint8_t* numpy_scalar,
                          const uint8_t* np_pre, uint32_t np_pre_len, const uint8_t* np_mid, uint32_t np_mid_len,
                          uint8_t* out, uint64_t out_capacity, uint64_t* out_off) {
    HIC_REQUIRE(symbols && lens && codes && out && out_off, "NULL argument");
    HIC_REQUIRE(!numpy_scalar || (np_pre && np_mid), "numpy scalar rows need the environment's prefix and infix");
    const RowFormat fmt{np_pre, np_pre_len, np_mid, np_mid_len};
    uint64_t pos = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t len = lens[i];
        HIC_REQUIRE(len >= 1 && len <= 58, "code length %u out of range in row %llu", len, (unsigned long long)i);
        const bool is_np = numpy_scalar && numpy_scalar[i];
        if (pos + row_size(symbols[i], len, is_np, fmt) > out_capacity)
            return hic::fail(HIC_ERR_CAPACITY, "row buffer of %llu bytes is too small", (unsigned long long)out_capacity);
        out_off[i] = pos;
        pos += write_row(out + pos, symbols[i], len, codes[i], is_np, fmt);
    }
    out_off[n] = pos;
    return HIC_OK;
}

int hic_hicfile_pack_table(const int32_t* symbols, const uint8_t* lens, const uint64_t* codes, uint64_t n, const uint8_t* numpy_scalar,
                           const uint8_t* np_pre, uint32_t np_pre_len, const uint8_t* np_mid, uint32_t np_mid_len,
                           const uint8_t* head, uint32_t head_len, uint8_t* out, uint64_t out_capacity, uint64_t* out_len) {
    HIC_REQUIRE((n == 0 || (symbols && lens && codes)) && head && out && out_len, "NULL argument");
    HIC_REQUIRE(!numpy_scalar || (np_pre && np_mid), "numpy scalar rows need the environment's prefix and infix");
    HIC_REQUIRE(out_capacity >= 2, "output buffer too small");
    const RowFormat fmt{np_pre, np_pre_len, np_mid, np_mid_len};
    out[0] = 0x80;                                        // PROTO 4, outside the frames
    out[1] = 0x04;
    FramedWriter w(out + 2, out_capacity - 2);
    w.bytes(head, head_len);                              // {"type": <class>, "data": [  -- up to the list's own MEMOIZE
    const uint64_t BATCH = 1000;                          // pickle's batch_list: MARK, up to 1000 items, APPENDS
    for (uint64_t first = 0; first < n; first += BATCH) {
        const uint64_t last = first + BATCH < n ? first + BATCH : n;
        if (n > 1) w.byte(0x28);                          // MARK
        for (uint64_t i = first; i < last; ++i) {
            const uint32_t len = lens[i];
            HIC_REQUIRE(len >= 1 && len <= 58, "code length %u out of range in row %llu", len, (unsigned long long)i);
            const bool is_np = numpy_scalar && numpy_scalar[i];
            const uint64_t size = row_size(symbols[i], len, is_np, fmt);
            w.object_boundary();
            if (size < 256) {                             // SHORT_BINBYTES | BINBYTES, the row, MEMOIZE
                if (uint8_t* p = w.reserve(2)) { p[0] = 0x43; p[1] = (uint8_t)size; }
            } else if (uint8_t* p = w.reserve(5)) {
                p[0] = 0x42; p[1] = (uint8_t)size; p[2] = (uint8_t)(size >> 8); p[3] = (uint8_t)(size >> 16); p[4] = (uint8_t)(size >> 24);
            }
            if (uint8_t* p = w.reserve(size)) write_row(p, symbols[i], len, codes[i], is_np, fmt);
            w.byte(0x94);
            if (w.overflow) break;
        }
        w.byte(n > 1 ? 0x65 : 0x61);                      // APPENDS | APPEND (a one-row list)
        if (w.overflow) break;
    }
    w.byte(0x75);                                         // SETITEMS of the two-entry dict
    w.byte(0x2e);                                         // STOP
    if (w.overflow)
        return hic::fail(HIC_ERR_CAPACITY, "table buffer of %llu bytes is too small", (unsigned long long)out_capacity);
    w.commit();
    *out_len = 2 + w.pos;
    return HIC_OK;
}

int hic_hicfile_parse_rows(const uint8_t* data, const uint64_t* off, uint64_t n, const uint8_t* np_pre, uint32_t np_pre_len,
                           const uint8_t* np_mid, uint32_t np_mid_len, int32_t* symbols, uint8_t* lens, uint64_t* codes,
                           uint8_t* numpy_scalar, int64_t* bad_row) {
    HIC_REQUIRE(data && off && symbols && lens && codes && numpy_scalar && bad_row, "NULL argument");
    const RowFormat fmt{np_pre, np_pre_len, np_mid, np_mid_len};
    *bad_row = -1;
    for (uint64_t i = 0; i < n; ++i) {
        if (!parse_row(data + off[i], off[i + 1] - off[i], fmt, symbols + i, lens + i, codes + i, numpy_scalar + i)) {
            *bad_row = (int64_t)i;
            return HIC_OK;
        }
    }
    return HIC_OK;
}

int hic_hicfile_parse_table(const uint8_t* data, uint64_t size, const uint8_t* np_pre, uint32_t np_pre_len, const uint8_t* np_mid,
                            uint32_t np_mid_len, int32_t* symbols, uint8_t* lens, uint64_t* codes, uint8_t* numpy_scalar,
                            uint64_t row_capacity, uint64_t* n_rows, int32_t* canonical) {
    HIC_REQUIRE(data && symbols && lens && codes && numpy_scalar && n_rows && canonical, "NULL argument");
    const RowFormat fmt{np_pre, np_pre_len, np_mid, np_mid_len};
    *canonical = 0;
    *n_rows = 0;
    TableReader r{data, size, 0};
    // PROTO 4 | 5 (the opcodes used here are the same in both), then  } MEMOIZE (  "type" MEMOIZE  <module> <name> STACK_GLOBAL
    // MEMOIZE  "data" MEMOIZE  ] MEMOIZE
    if (size < 2 || data[0] != 0x80 || (data[1] != 0x04 && data[1] != 0x05)) return HIC_OK;
    r.pos = 2;
    static const uint8_t OPEN[3] = {0x7d, 0x94, 0x28};
    static const uint8_t LIST[2] = {0x5d, 0x94};
    const uint8_t* text;
    uint32_t n;
    if (!r.expect(OPEN, 3) || !r.short_string(&text, &n) || n != 4 || memcmp(text, "type", 4) != 0) return HIC_OK;
    if (!r.short_string(&text, &n) || n < 8 || memcmp(text + n - 8, "hicimage", 8) != 0) return HIC_OK;      // <any package>.hicimage
    if (n > 8 && text[n - 9] != '.') return HIC_OK;
    if (!r.short_string(&text, &n) || n != 4 || memcmp(text, "TupP", 4) != 0) return HIC_OK;
    if (!r.opcode(0x93) || !r.opcode(0x94)) return HIC_OK;
    if (!r.short_string(&text, &n) || n != 4 || memcmp(text, "data", 4) != 0) return HIC_OK;
    if (!r.expect(LIST, 2)) return HIC_OK;
    uint64_t rows = 0;
    bool in_batch = false, closed = false;
    while (!closed) {
        if (!r.skip_frames() || r.pos >= size) return HIC_OK;
        const uint8_t op = data[r.pos];
        if (op == 0x28 && !in_batch) {                    // MARK
            in_batch = true;
            ++r.pos;
        } else if (op == 0x65 && in_batch) {              // APPENDS
            in_batch = false;
            ++r.pos;
        } else if (op == 0x75 && !in_batch) {             // SETITEMS: the end of the dict
            closed = true;
            ++r.pos;
        } else if (op == 0x43 || op == 0x42) {            // a row: SHORT_BINBYTES | BINBYTES, MEMOIZE (, APPEND outside a batch)
            uint64_t len;
            if (op == 0x43) {
                if (r.pos + 2 > size) return HIC_OK;
                len = data[r.pos + 1];
                r.pos += 2;
            } else {
                if (r.pos + 5 > size) return HIC_OK;
                len = (uint64_t)data[r.pos + 1] | ((uint64_t)data[r.pos + 2] << 8) | ((uint64_t)data[r.pos + 3] << 16) | ((uint64_t)data[r.pos + 4] << 24);
                r.pos += 5;
            }
            if (len > size - r.pos || rows >= row_capacity) return HIC_OK;
            if (!parse_row(data + r.pos, len, fmt, symbols + rows, lens + rows, codes + rows, numpy_scalar + rows)) return HIC_OK;
            r.pos += len;
            if (!r.opcode(0x94)) return HIC_OK;
            ++rows;
            if (!in_batch && !r.opcode(0x61)) return HIC_OK;
        } else {
            return HIC_OK;
        }
    }
    if (!r.opcode(0x2e) || r.pos != size) return HIC_OK;  // STOP, nothing after it
    *n_rows = rows;
    *canonical = 1;
    return HIC_OK;
}

}  // extern "C"
