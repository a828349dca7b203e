// hic_hicfile.cu -- host-side helpers for the `.hic` container (no device code): Huffman table rows, whole table
// payloads and whole files to and from the byte strings the reference pickles them as.
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
// hic_hicfile_pack_files / scan_files / parse_files do the same for the whole file (hicimage.py:165-183:
// pickle.dump of the list [mode string, table payloads, framed bit strings, shapes]) and for every image of a batch,
// files shared out over host threads: from the encode result's packed tables and bit strings to file bytes and back,
// with no Python object per row, table or file.
#include <stdint.h>
#include <string.h>
#include <atomic>
#include <thread>
#include <vector>
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
    // A bytes object of 64 KiB or more goes out unframed: the open frame is closed first, the opcode, length and
    // data follow bare (_Pickler_write_bytes), and whatever comes next opens a new frame.
    uint8_t* unframed(uint64_t n) {
        commit();
        if (pos + n > cap) { overflow = true; return nullptr; }
        uint8_t* p = out + pos;
        pos += n;
        return p;
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

// The rows of one table, either as separate arrays (hic_hicfile_pack_table) or in the packed layout of
// hic_entropy_tables_packed (length << 58 | code).  flag_mode: 0 Python ints, 1 numpy.int32 scalars, 2 numpy.int32 unless the
// symbol is 0 (the wavelet value tables), 3 per row from `flags`.
struct TableRows {
    const int32_t* symbols;
    const uint8_t* lens;
    const uint64_t* codes;
    const uint64_t* packed;
    const uint8_t* flags;
    int flag_mode;
    uint64_t n;
    uint32_t len(uint64_t i) const { return packed ? (uint32_t)(packed[i] >> 58) : lens[i]; }
    uint64_t code(uint64_t i) const { return packed ? (packed[i] & ((1ull << 58) - 1)) : codes[i]; }
    bool is_np(uint64_t i) const {
        return flag_mode == 1 || (flag_mode == 2 && symbols[i] != 0) || (flag_mode == 3 && flags[i]);
    }
};

// pickle.dumps({"type": TupP, "data": [row pickles]}) at out; returns its size, 0 if it does not fit.
inline uint64_t write_table(const TableRows& t, const RowFormat& fmt, const uint8_t* head, uint32_t head_len, uint8_t* out, uint64_t cap) {
    if (cap < 2) return 0;
    out[0] = 0x80;                                        // PROTO 4, outside the frames
    out[1] = 0x04;
    FramedWriter w(out + 2, cap - 2);
    w.bytes(head, head_len);                              // {"type": <class>, "data": [  -- up to the list's own MEMOIZE
    const uint64_t BATCH = 1000;                          // pickle's batch_list: MARK, up to 1000 items, APPENDS
    for (uint64_t first = 0; first < t.n && !w.overflow; first += BATCH) {
        const uint64_t last = first + BATCH < t.n ? first + BATCH : t.n;
        if (t.n > 1) w.byte(0x28);                        // MARK
        for (uint64_t i = first; i < last && !w.overflow; ++i) {
            const uint32_t len = t.len(i);
            const bool is_np = t.is_np(i);
            const uint64_t size = row_size(t.symbols[i], len, is_np, fmt);
            w.object_boundary();
            if (size < 256) {                             // SHORT_BINBYTES | BINBYTES, the row, MEMOIZE
                if (uint8_t* p = w.reserve(2)) { p[0] = 0x43; p[1] = (uint8_t)size; }
            } else if (uint8_t* p = w.reserve(5)) {
                p[0] = 0x42; p[1] = (uint8_t)size; p[2] = (uint8_t)(size >> 8); p[3] = (uint8_t)(size >> 16); p[4] = (uint8_t)(size >> 24);
            }
            if (uint8_t* p = w.reserve(size)) write_row(p, t.symbols[i], len, t.code(i), is_np, fmt);
            w.byte(0x94);
        }
        w.byte(t.n > 1 ? 0x65 : 0x61);                    // APPENDS | APPEND (a one-row list)
    }
    w.byte(0x75);                                         // SETITEMS of the two-entry dict
    w.byte(0x2e);                                         // STOP
    if (w.overflow) return 0;
    w.commit();
    return 2 + w.pos;
}

// Room that always suffices for `payload` bytes of opcodes and data cut into frames, with `items` possible extra cuts.
inline uint64_t framed_bound(uint64_t payload, uint64_t items) {
    return 2 + payload + FramedWriter::FRAME_HEADER * (payload / FramedWriter::FRAME_TARGET + 2 * items + 2);
}

inline uint64_t table_bound(const TableRows& t, const RowFormat& fmt, uint32_t head_len) {
    uint64_t payload = head_len + 4 + 2 * (t.n / 1000 + 1);
    for (uint64_t i = 0; i < t.n; ++i) payload += row_size(t.symbols[i], t.len(i), t.is_np(i), fmt) + 6;
    return framed_bound(payload, 0);
}

inline TableRows table_rows(const hic_hicfile_batch* b, uint32_t s, uint32_t k) {
    const uint64_t first = b->index[2 * (uint64_t)s], n = b->index[2 * (uint64_t)s + 1];
    return TableRows{b->symbols + first, nullptr, nullptr, b->packed + first, nullptr, b->flag_mode[k], n};
}

inline bool check_batch(const hic_hicfile_env* env, const hic_hicfile_batch* b) {
    return env->np_pre && env->np_mid && env->head && (b->n_files == 0 || (b->stream_of && b->flag_mode && b->index && b->symbols &&
           b->packed && b->data && b->byte_off && b->byte_len)) && (b->lead_len == 0 || b->lead) && (b->n_trail == 0 || (b->trail && b->trail_len));
}

// One entry of the file's list: a bytes object, memoised.  Objects of 64 KiB or more bypass the frames.
inline void write_item(FramedWriter& w, const uint8_t* data, uint64_t n) {
    w.object_boundary();
    uint8_t hdr[5];
    uint32_t hn;
    if (n < 256) { hdr[0] = 0x43; hdr[1] = (uint8_t)n; hn = 2; }
    else { hdr[0] = 0x42; hdr[1] = (uint8_t)n; hdr[2] = (uint8_t)(n >> 8); hdr[3] = (uint8_t)(n >> 16); hdr[4] = (uint8_t)(n >> 24); hn = 5; }
    if (n >= FramedWriter::FRAME_TARGET) {
        if (uint8_t* p = w.unframed(hn + n)) { memcpy(p, hdr, hn); memcpy(p + hn, data, n); }
    } else if (uint8_t* p = w.reserve(hn + n)) {
        memcpy(p, hdr, hn);
        memcpy(p + hn, data, n);
    }
    w.byte(0x94);
}

// pickle.dumps([lead, table payloads..., framed bit strings..., trail entries...]) of file i at out.  Returns its size,
// 0 if the file is left to the caller (an entry of fewer than two bytes: CPython shares those objects, so pickle may
// write a memo reference instead of the bytes), -1 if it does not fit.
inline int64_t write_file(const hic_hicfile_env* env, const hic_hicfile_batch* b, const RowFormat& fmt, uint64_t i, uint8_t* out,
                          uint64_t cap, std::vector<uint8_t>& scratch) {
    const uint32_t T = b->tables_per_file;
    const uint64_t items = 1 + 2ull * T + b->n_trail;
    if (b->lead_len < 2 || b->lead_len >= (1ull << 32)) return 0;
    for (uint32_t k = 0; k < T; ++k) {
        const uint64_t n = b->byte_len[b->stream_of[i * T + k]];
        if (n < 2 || n >= (1ull << 32)) return 0;
    }
    for (uint32_t t = 0; t < b->n_trail; ++t)
        if (b->trail_len[t] < 2 || b->trail_len[t] >= (1ull << 32)) return 0;
    if (cap < 2) return -1;
    out[0] = 0x80;
    out[1] = 0x04;
    FramedWriter w(out + 2, cap - 2);
    w.byte(0x5d);                                         // EMPTY_LIST, MEMOIZE
    w.byte(0x94);
    for (uint64_t first = 0; first < items; first += 1000) {      // (items is 15 or 21 plus extensions: one batch in practice)
        const uint64_t last = first + 1000 < items ? first + 1000 : items;
        if (items > 1) w.byte(0x28);
        for (uint64_t e = first; e < last && !w.overflow; ++e) {
            if (e == 0) {
                write_item(w, b->lead, b->lead_len);
            } else if (e <= T) {
                const uint32_t k = (uint32_t)(e - 1);
                const TableRows rows = table_rows(b, b->stream_of[i * T + k], k);
                const uint64_t room = table_bound(rows, fmt, env->head_len);
                if (scratch.size() < room) scratch.resize(room + room / 2);
                const uint64_t size = write_table(rows, fmt, env->head, env->head_len, scratch.data(), scratch.size());
                if (size == 0 || size >= (1ull << 32)) return size == 0 ? -1 : 0;
                write_item(w, scratch.data(), size);
            } else if (e <= 2ull * T) {
                const uint32_t s = b->stream_of[i * T + (e - 1 - T)];
                write_item(w, b->data + b->byte_off[s], b->byte_len[s]);
            } else {
                uint64_t off = 0;
                for (uint64_t t = 0; t < e - 1 - 2ull * T; ++t) off += b->trail_len[t];
                write_item(w, b->trail + off, b->trail_len[e - 1 - 2ull * T]);
            }
        }
        w.byte(items > 1 ? 0x65 : 0x61);
    }
    w.byte(0x2e);
    if (w.overflow) return -1;
    w.commit();
    return (int64_t)(2 + w.pos);
}

// Walks a table payload (see hic_hicfile_parse_table) and hands every row's bytes to visit(row, size, index); returns the
// row count, or -1 if the payload is not of the canonical form or a visit declined.
template <typename Visit>
inline int64_t walk_table(const uint8_t* data, uint64_t size, Visit&& visit) {
    TableReader r{data, size, 0};
    if (size < 2 || data[0] != 0x80 || (data[1] != 0x04 && data[1] != 0x05)) return -1;
    r.pos = 2;
    static const uint8_t OPEN[3] = {0x7d, 0x94, 0x28};
    static const uint8_t LIST[2] = {0x5d, 0x94};
    const uint8_t* text;
    uint32_t n;
    if (!r.expect(OPEN, 3) || !r.short_string(&text, &n) || n != 4 || memcmp(text, "type", 4) != 0) return -1;
    if (!r.short_string(&text, &n) || n < 8 || memcmp(text + n - 8, "hicimage", 8) != 0) return -1;       // <any package>.hicimage
    if (n > 8 && text[n - 9] != '.') return -1;
    if (!r.short_string(&text, &n) || n != 4 || memcmp(text, "TupP", 4) != 0) return -1;
    if (!r.opcode(0x93) || !r.opcode(0x94)) return -1;
    if (!r.short_string(&text, &n) || n != 4 || memcmp(text, "data", 4) != 0) return -1;
    if (!r.expect(LIST, 2)) return -1;
    int64_t rows = 0;
    bool in_batch = false, closed = false;
    while (!closed) {
        if (!r.skip_frames() || r.pos >= size) return -1;
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
                if (r.pos + 2 > size) return -1;
                len = data[r.pos + 1];
                r.pos += 2;
            } else {
                if (r.pos + 5 > size) return -1;
                len = (uint64_t)data[r.pos + 1] | ((uint64_t)data[r.pos + 2] << 8) | ((uint64_t)data[r.pos + 3] << 16) | ((uint64_t)data[r.pos + 4] << 24);
                r.pos += 5;
            }
            if (len > size - r.pos) return -1;
            if (!visit(data + r.pos, len, (uint64_t)rows)) return -1;
            r.pos += len;
            if (!r.opcode(0x94)) return -1;
            ++rows;
            if (!in_batch && !r.opcode(0x61)) return -1;
        } else {
            return -1;
        }
    }
    if (!r.opcode(0x2e) || r.pos != size) return -1;      // STOP, nothing after it
    return rows;
}

// The entries of a `.hic` file, i.e. of pickle.dumps([bytes, bytes, ...]) under protocol 4 or 5: positions and sizes of
// the first `want` entries.  Returns the entry count, -1 if the file is anything else (memo references, other types).
inline int64_t walk_file(const uint8_t* data, uint64_t size, uint64_t want, uint64_t* item_off, uint64_t* item_len) {
    TableReader r{data, size, 0};
    if (size < 2 || data[0] != 0x80 || (data[1] != 0x04 && data[1] != 0x05)) return -1;
    r.pos = 2;
    static const uint8_t LIST[2] = {0x5d, 0x94};
    if (!r.expect(LIST, 2)) return -1;
    int64_t items = 0;
    bool in_batch = false;
    for (;;) {
        if (!r.skip_frames() || r.pos >= size) return -1;
        const uint8_t op = data[r.pos];
        if (op == 0x28 && !in_batch) { in_batch = true; ++r.pos; }
        else if (op == 0x65 && in_batch) { in_batch = false; ++r.pos; }
        else if (op == 0x2e && !in_batch) { ++r.pos; break; }
        else if (op == 0x43 || op == 0x42 || op == 0x8e) {
            const uint32_t hn = op == 0x43 ? 2 : (op == 0x42 ? 5 : 9);
            if (r.pos + hn > size) return -1;
            uint64_t len = 0;
            for (uint32_t k = 1; k < hn; ++k) len |= (uint64_t)data[r.pos + k] << (8 * (k - 1));
            r.pos += hn;
            if (len > size - r.pos) return -1;
            if ((uint64_t)items < want) { item_off[items] = r.pos; item_len[items] = len; }
            r.pos += len;
            if (!r.opcode(0x94)) return -1;
            ++items;
            if (!in_batch && !r.opcode(0x61)) return -1;
        } else return -1;
    }
    return r.pos == size ? items : -1;
}

// work(i) for i in [0, n) on up to `threads` host threads (the caller's included), items handed out one at a time.
// Nothing may escape a C entry point: an exception in a worker (std::bad_alloc from a scratch buffer) stops the hand-out
// and makes the call return false; a thread that cannot be started only means fewer workers.
template <typename Work>
inline bool run_threads(uint64_t n, uint32_t threads, Work&& work) {
    std::atomic<uint64_t> next{0};
    std::atomic<bool> threw{false};
    auto loop = [&]() {
        try {
            for (;;) {
                const uint64_t i = next.fetch_add(1);
                if (i >= n || threw.load()) break;
                work(i);
            }
        } catch (...) {
            threw.store(true);
        }
    };
    const uint64_t workers = threads < 1 ? 1 : (threads > n ? (n ? n : 1) : threads);
    std::vector<std::thread> pool;
    try {
        for (uint64_t t = 0; t + 1 < workers; ++t) pool.emplace_back(loop);
    } catch (...) {
    }
    loop();
    for (auto& t : pool) t.join();
    return !threw.load();
}

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
    for (uint64_t i = 0; i < n; ++i)
        HIC_REQUIRE(lens[i] >= 1 && lens[i] <= 58, "code length %u out of range in row %llu", (unsigned)lens[i], (unsigned long long)i);
    const RowFormat fmt{np_pre, np_pre_len, np_mid, np_mid_len};
    const TableRows rows{symbols, lens, codes, nullptr, numpy_scalar, numpy_scalar ? 3 : 0, n};
    const uint64_t size = write_table(rows, fmt, head, head_len, out, out_capacity);
    if (size == 0)
        return hic::fail(HIC_ERR_CAPACITY, "table buffer of %llu bytes is too small", (unsigned long long)out_capacity);
    *out_len = size;
    return HIC_OK;
}

int hic_hicfile_files_bound(const hic_hicfile_env* env, const hic_hicfile_batch* b, uint64_t* bound, uint32_t threads) {
    HIC_REQUIRE(env && b && bound, "NULL argument");
    HIC_REQUIRE(check_batch(env, b), "incomplete batch description");
    const RowFormat fmt{env->np_pre, env->np_pre_len, env->np_mid, env->np_mid_len};
    uint64_t trail = 0;
    for (uint32_t t = 0; t < b->n_trail; ++t) trail += b->trail_len[t];
    const bool done = run_threads(b->n_files, threads, [&](uint64_t i) {
        uint64_t payload = b->lead_len + trail;
        const uint64_t items = 1 + 2ull * b->tables_per_file + b->n_trail;
        for (uint32_t k = 0; k < b->tables_per_file; ++k) {
            const uint32_t s = b->stream_of[i * b->tables_per_file + k];
            payload += table_bound(table_rows(b, s, k), fmt, env->head_len) + b->byte_len[s];
        }
        bound[i] = framed_bound(payload + 6 * items + 16, items);
    });
    HIC_REQUIRE(done, "a worker thread failed");
    return HIC_OK;
}

int hic_hicfile_pack_files(const hic_hicfile_env* env, const hic_hicfile_batch* b, uint8_t* out, const uint64_t* out_off,
                           uint64_t* out_len, uint32_t threads) {
    HIC_REQUIRE(env && b && out && out_off && out_len, "NULL argument");
    HIC_REQUIRE(check_batch(env, b), "incomplete batch description");
    const RowFormat fmt{env->np_pre, env->np_pre_len, env->np_mid, env->np_mid_len};
    std::atomic<int> failed{0};
    const bool done = run_threads(b->n_files, threads, [&](uint64_t i) {
        thread_local std::vector<uint8_t> scratch;
        out_len[i] = 0;
        for (uint32_t k = 0; k < b->tables_per_file; ++k) {            // a code length outside 1..58 is the caller's error
            const TableRows r = table_rows(b, b->stream_of[i * b->tables_per_file + k], k);
            for (uint64_t j = 0; j < r.n; ++j) {
                const uint32_t len = (uint32_t)(r.packed[j] >> 58);
                if (len < 1 || len > 58) { failed.store(2); return; }
            }
        }
        const int64_t n = write_file(env, b, fmt, i, out + out_off[i], out_off[i + 1] - out_off[i], scratch);
        if (n < 0) failed.store(1);
        else out_len[i] = (uint64_t)n;
    });
    if (!done) return hic::fail(HIC_ERR_CAPACITY, "out of host memory while writing the batch's files");
    if (failed.load() == 2) return hic::fail(HIC_ERR_INVALID, "a code length outside 1..58 in the batch's tables");
    if (failed.load()) return hic::fail(HIC_ERR_CAPACITY, "a file did not fit the room hic_hicfile_files_bound gives it");
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
    const int64_t rows = walk_table(data, size, [&](const uint8_t* row, uint64_t len, uint64_t i) {
        return i < row_capacity && parse_row(row, len, fmt, symbols + i, lens + i, codes + i, numpy_scalar + i);
    });
    if (rows < 0) return HIC_OK;
    *n_rows = (uint64_t)rows;
    *canonical = 1;
    return HIC_OK;
}

int hic_hicfile_scan_files(const uint8_t* const* files, const uint64_t* file_len, uint64_t n_files, uint32_t tables_per_file, uint32_t n_items,
                           uint64_t* item_off, uint64_t* item_len, uint32_t* rows, uint8_t* canonical, uint32_t threads) {
    HIC_REQUIRE(files && file_len && item_off && item_len && rows && canonical, "NULL argument");
    HIC_REQUIRE(n_items >= 1 + 2 * tables_per_file, "a file holds the mode entry, the tables and as many bit strings");
    const bool done = run_threads(n_files, threads, [&](uint64_t i) {
        uint64_t* off = item_off + i * n_items;
        uint64_t* len = item_len + i * n_items;
        canonical[i] = 0;
        if (!files[i]) return;
        const int64_t items = walk_file(files[i], file_len[i], n_items, off, len);
        if (items < (int64_t)n_items) return;
        for (uint32_t k = 0; k < tables_per_file; ++k) {
            const int64_t r = walk_table(files[i] + off[1 + k], len[1 + k], [](const uint8_t*, uint64_t, uint64_t) { return true; });
            if (r < 0 || r > 0x7fffffff) return;
            rows[i * tables_per_file + k] = (uint32_t)r;
        }
        canonical[i] = 1;
    });
    HIC_REQUIRE(done, "a worker thread failed");
    return HIC_OK;
}

int hic_hicfile_parse_files(const hic_hicfile_env* env, const uint8_t* const* files, uint64_t n_files, uint32_t tables_per_file,
                            uint32_t n_items, const uint64_t* item_off, const uint64_t* item_len, const uint32_t* stream_of,
                            const uint32_t* index, int32_t* symbols, uint64_t* packed, uint8_t* data, const uint64_t* byte_off,
                            uint64_t* nbits, uint8_t* ok, uint32_t threads) {
    HIC_REQUIRE(env && files && item_off && item_len && stream_of && index && symbols && packed && data && byte_off && nbits && ok,
                "NULL argument");
    const RowFormat fmt{env->np_pre, env->np_pre_len, env->np_mid, env->np_mid_len};
    const bool done = run_threads(n_files, threads, [&](uint64_t i) {
        const uint64_t* off = item_off + i * n_items;
        const uint64_t* len = item_len + i * n_items;
        ok[i] = 0;
        for (uint32_t k = 0; k < tables_per_file; ++k) {
            const uint32_t s = stream_of[i * tables_per_file + k];
            const uint64_t first = index[2 * (uint64_t)s], n = index[2 * (uint64_t)s + 1];
            const int64_t r = walk_table(files[i] + off[1 + k], len[1 + k], [&](const uint8_t* row, uint64_t size, uint64_t j) {
                uint8_t bits, is_np;
                uint64_t code;
                if (j >= n || !parse_row(row, size, fmt, symbols + first + j, &bits, &code, &is_np)) return false;
                packed[first + j] = ((uint64_t)bits << 58) | code;
                return true;
            });
            if (r != (int64_t)n) return;
            // the framed bit string: a pad byte (read as int8 by the reference, iohelper.py) and the bytes
            const uint8_t* framed = files[i] + off[1 + tables_per_file + k];
            const uint64_t fl = len[1 + tables_per_file + k];
            memcpy(data + byte_off[s], framed, fl);
            int64_t bits = 0;
            if (fl) {
                const int64_t pad = framed[0] >= 128 ? (int64_t)framed[0] - 256 : (int64_t)framed[0];
                bits = 8 * ((int64_t)fl - 1) - pad;
            }
            nbits[s] = bits > 0 ? (uint64_t)bits : 0;
        }
        ok[i] = 1;
    });
    HIC_REQUIRE(done, "a worker thread failed");
    return HIC_OK;
}

}  // extern "C"
