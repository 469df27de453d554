// INGEST: interaction.csv text -> device-resident COO, the wire format of the drop-in.
//
// Replaces tf.data make_csv_dataset(select_columns=[row, col, weight, target]) [ref src/models/data_utils.py:4-26] and the
// per-step StaticHashTable string -> id lookup with default 0 [ref src/models/model_utils.py:121-127,
// src/models/estimator.py:26-28]: the file bytes are parsed ONCE on the GPU, token strings are resolved against
// vocab.txt once, and the result is the (row, col, colA, colB) triple buffer that glove_prepare_batches reads.
//
//   tile_stats_kernel   one pass over the text: per 16 KB tile, the quote-parity flip and the number of record
//                       terminators under both hypotheses (tile starts outside / inside a quoted field)
//   cub InclusiveScan   composes the tile summaries (associative, non-commutative) -> parity and record index at every
//                       tile start
//   row_ends_kernel     second pass: position of the terminating '\n' of every non-empty record
//   parse_rows_kernel   one thread per record: split the fields (RFC-4180 quoting, "" escapes, \r\n), resolve the two
//                       token columns through the device hash table (or parse *_id columns), convert the two value
//                       columns with the correctly rounded decimal -> float32 routine of glove_strtof.cuh
//   vocab_build_kernel  open-addressing table of vocab line numbers keyed by the token bytes (first line wins)
//
// HBM-bound byte work: no tensor cores.  Errors never fall back to the host: the lowest offending record is reported.
#include <cub/cub.cuh>

#include "glove_common.cuh"
#include "glove_strtof.cuh"

namespace glove {

constexpr int kTileThreads = 256, kThreadBytes = 64, kTileBytes = kTileThreads * kThreadBytes;

// summary of a byte range: does it flip the quote parity; records terminated inside it if it starts outside / inside quotes
struct QState { uint32_t flip, n_out, n_in; };
struct QCombine {
    __host__ __device__ __forceinline__ QState operator()(const QState &a, const QState &b) const {
        QState r;
        r.flip = a.flip ^ b.flip;
        r.n_out = a.n_out + (a.flip ? b.n_in : b.n_out);
        r.n_in = a.n_in + (a.flip ? b.n_out : b.n_in);
        return r;
    }
};

enum { CSV_OK = 0, CSV_EFIELDS = 1, CSV_EQUOTE = 2, CSV_EFLOAT = 3, CSV_EDIGITS = 4, CSV_EINT = 5, CSV_ERANGE = 6 };
struct CsvStatus {
    unsigned long long first_error;   // (record << 8) | code, minimum over all records; ~0 = none
    long long n_terms;                // non-empty terminated records of the chunk
    long long tail_valid;             // final chunk: 1 when an unterminated last record exists
    long long last_end;               // position of the last terminator (-1: none)
    long long open_quote;             // parity at the end of the chunk
};

struct CsvWs {
    QState *tile;       // [n_tiles] inclusive scan of the tile summaries
    QState *tile_raw;   // [n_tiles]
    CsvStatus *status;
    void *cub_temp;
    size_t cub_bytes;
    size_t bytes;
};
static int64_t n_tiles_of(int64_t nbytes) { return (nbytes + kTileBytes - 1) / kTileBytes; }
static CsvWs csv_ws_view(void *base, int64_t nbytes) {
    CsvWs w;
    char *p = (char *)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char *r = p ? p + off : nullptr; off += align_up(bytes); return r; };
    const int64_t nt = n_tiles_of(nbytes) + 1;
    w.status = (CsvStatus *)take(sizeof(CsvStatus));
    w.tile = (QState *)take(sizeof(QState) * nt);
    w.tile_raw = (QState *)take(sizeof(QState) * nt);
    w.cub_bytes = 0;
    cub::DeviceScan::InclusiveScan(nullptr, w.cub_bytes, (QState *)nullptr, (QState *)nullptr, QCombine(), (int)nt);
    w.cub_temp = take(w.cub_bytes);
    w.bytes = off;
    return w;
}

// a '\n' outside quotes at `pos` terminates an EMPTY record when nothing but an optional '\r' lies between it and the
// previous terminator (or the start of the chunk); empty records are skipped like tf.data's CsvDataset does
__device__ __forceinline__ bool empty_record(const uint8_t *text, int64_t pos) {
    if (pos == 0) return true;
    uint8_t p1 = text[pos - 1];
    if (p1 == '\r') {
        if (pos == 1) return true;
        p1 = text[pos - 2];
    }
    return p1 == '\n';
}

// the thread's 64 bytes as 4 x 16 (zero beyond the end: NUL is neither a quote nor a newline)
__device__ __forceinline__ void load_span(const uint8_t *text, int64_t nbytes, int64_t base, uint4 (&v)[4]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int64_t o = base + 16 * u;
        if (o + 16 <= nbytes) v[u] = __ldg((const uint4 *)(text + o));
        else {
            uint8_t b[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) b[k] = o + k < nbytes ? text[o + k] : 0;
            v[u] = *(const uint4 *)b;
        }
    }
}
__device__ __forceinline__ uint8_t span_byte(const uint4 (&v)[4], int i) {
    const uint32_t w = ((const uint32_t *)v)[i >> 2];
    return (uint8_t)(w >> (8 * (i & 3)));
}

__global__ void __launch_bounds__(kTileThreads) tile_stats_kernel(const uint8_t *__restrict__ text, int64_t nbytes,
                                                                  QState *tile_raw) {
    const int64_t base = (int64_t)blockIdx.x * kTileBytes + (int64_t)threadIdx.x * kThreadBytes;
    QState s = {0u, 0u, 0u};
    if (base < nbytes) {
        uint4 v[4];
        load_span(text, nbytes, base, v);
        uint32_t par = 0, c[2] = {0u, 0u};
#pragma unroll
        for (int i = 0; i < kThreadBytes; ++i) {
            const uint8_t b = span_byte(v, i);
            if (b == '"') par ^= 1u;
            else if (b == '\n' && !empty_record(text, base + i)) ++c[par];
        }
        s.flip = par; s.n_out = c[0]; s.n_in = c[1];
    }
    typedef cub::BlockScan<QState, kTileThreads> Scan;
    __shared__ typename Scan::TempStorage tmp;
    QState incl;
    Scan(tmp).InclusiveScan(s, incl, QCombine());
    if (threadIdx.x == kTileThreads - 1) tile_raw[blockIdx.x] = incl;
}

__global__ void __launch_bounds__(kTileThreads) row_ends_kernel(const uint8_t *__restrict__ text, int64_t nbytes,
                                                                const QState *__restrict__ tile, int64_t n_tiles,
                                                                int64_t *row_ends, int64_t capacity, CsvStatus *status) {
    const int64_t base = (int64_t)blockIdx.x * kTileBytes + (int64_t)threadIdx.x * kThreadBytes;
    QState s = {0u, 0u, 0u};
    uint4 v[4];
    if (base < nbytes) {
        load_span(text, nbytes, base, v);
        uint32_t par = 0, c[2] = {0u, 0u};
#pragma unroll
        for (int i = 0; i < kThreadBytes; ++i) {
            const uint8_t b = span_byte(v, i);
            if (b == '"') par ^= 1u;
            else if (b == '\n' && !empty_record(text, base + i)) ++c[par];
        }
        s.flip = par; s.n_out = c[0]; s.n_in = c[1];
    }
    typedef cub::BlockScan<QState, kTileThreads> Scan;
    __shared__ typename Scan::TempStorage tmp;
    QState excl;
    const QState ident = {0u, 0u, 0u};
    Scan(tmp).ExclusiveScan(s, excl, ident, QCombine());
    const QState before = blockIdx.x ? QCombine()(tile[blockIdx.x - 1], excl) : excl;   // the chunk starts outside quotes
    if (base >= nbytes) return;
    uint32_t par = before.flip;
    int64_t r = before.n_out;
#pragma unroll
    for (int i = 0; i < kThreadBytes; ++i) {
        const uint8_t b = span_byte(v, i);
        if (b == '"') par ^= 1u;
        else if (b == '\n' && !par && !empty_record(text, base + i)) {
            if (r < capacity) row_ends[r] = base + i;
            ++r;
        }
    }
    if (base + kThreadBytes >= nbytes) status->open_quote = par;
}

__global__ void init_status_kernel(CsvStatus *st, const QState *tile, int64_t n_tiles) {
    st->first_error = ~0ull;
    st->n_terms = n_tiles ? tile[n_tiles - 1].n_out : 0;
    st->tail_valid = 0;
    st->last_end = -1;
    st->open_quote = 0;
}

// ---- vocab hash table ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t fnv1a(uint64_t h, uint8_t b) { return (h ^ b) * 0x100000001b3ull; }
constexpr uint64_t kFnvSeed = 0xcbf29ce484222325ull;
__device__ __forceinline__ uint32_t slot_of(uint64_t h, int64_t slots) {
    h ^= h >> 29;
    h *= 0xbf58476d1ce4e5b9ull;
    h ^= h >> 32;
    return (uint32_t)(h & (uint64_t)(slots - 1));
}

// vocab line v = vocab_bytes[vocab_off[v] .. vocab_off[v+1] - 1)   (every line is followed by one '\n')
__global__ void __launch_bounds__(256) vocab_build_kernel(int32_t *table, int64_t slots, const uint8_t *__restrict__ vb,
                                                          const int64_t *__restrict__ voff, int64_t n_vocab) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_vocab) return;
    const int64_t b0 = voff[v], len = voff[v + 1] - 1 - b0;
    uint64_t h = kFnvSeed;
    for (int64_t i = 0; i < len; ++i) h = fnv1a(h, vb[b0 + i]);
    uint32_t s = slot_of(h, slots);
    for (;;) {
        const int32_t old = atomicCAS(&table[s], -1, (int32_t)v);
        if (old == -1) return;
        // occupied: the same token (duplicate vocab line) keeps the LOWEST line number; another token -> next slot
        const int64_t o0 = voff[old], olen = voff[old + 1] - 1 - o0;
        bool same = olen == len;
        for (int64_t i = 0; same && i < len; ++i) same = vb[o0 + i] == vb[b0 + i];
        if (same) { atomicMin(&table[s], (int32_t)v); return; }
        s = (s + 1) & (uint32_t)(slots - 1);
    }
}

struct Schema {
    int32_t n_cols;
    int32_t sel[4];    // column index of row, col, colA, colB
    int32_t kind[4];   // GLOVE_CSV_TOKEN / _INT / _FLOAT
};
struct Field { int32_t b, e, esc; };   // content bytes [b, e); esc: contains "" escapes

// the bytes of one record, addressed relative to its first byte (32-bit offsets keep the address arithmetic short)
struct Bytes {
    const uint8_t *p;
    __device__ __forceinline__ uint8_t operator[](int i) const { return __ldg(p + i); }
};

// logical bytes of a field (collapses "" to ")
struct FieldBytes {
    const Bytes &text;
    int i, e;
    bool esc;
    __device__ __forceinline__ bool next(uint8_t &c) {
        if (i >= e) return false;
        c = text[i++];
        if (esc && c == '"') ++i;
        return true;
    }
};

__device__ int32_t lookup_token(const Bytes &text, const Field &f, const int32_t *__restrict__ table,
                                int64_t slots, const uint8_t *__restrict__ vb, const int64_t *__restrict__ voff) {
    uint64_t h = kFnvSeed;
    int len = 0;
    {
        FieldBytes it = {text, f.b, f.e, f.esc != 0};
        uint8_t c;
        while (it.next(c)) { h = fnv1a(h, c); ++len; }
    }
    uint32_t s = slot_of(h, slots);
    for (;;) {
        const int32_t v = table[s];
        if (v < 0) return 0;                       // StaticHashTable default_value = 0
        const int64_t o0 = voff[v];
        if (voff[v + 1] - 1 - o0 == len) {
            FieldBytes it = {text, f.b, f.e, f.esc != 0};
            uint8_t c;
            int k = 0;
            bool same = true;
            while (same && it.next(c)) same = vb[o0 + k++] == c;
            if (same) return v;
        }
        s = (s + 1) & (uint32_t)(slots - 1);
    }
}

__device__ __forceinline__ void report(CsvStatus *st, int64_t record, int code) {
    atomicMin(&st->first_error, ((unsigned long long)record << 8) | (unsigned)code);
}

__global__ void __launch_bounds__(128) parse_rows_kernel(const uint8_t *__restrict__ text_bytes, int64_t nbytes, int32_t final_chunk,
                                                         const int64_t *__restrict__ row_ends, CsvStatus *st, Schema sc,
                                                         const int32_t *__restrict__ table, int64_t slots,
                                                         const uint8_t *__restrict__ vb, const int64_t *__restrict__ voff,
                                                         int64_t n_vocab, int32_t *out_row, int32_t *out_col, float *out_a,
                                                         float *out_b, int64_t capacity, int64_t record0) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_terms = st->n_terms;
    if (r > n_terms || r >= capacity || (r == n_terms && !final_chunk)) return;
    int64_t pos0 = r ? row_ends[r - 1] + 1 : 0;
    int64_t end0 = r < n_terms ? row_ends[r] : nbytes;
    if (r == n_terms - 1) st->last_end = end0;
    // skipped empty records in front of this one
    while (pos0 < end0 && (text_bytes[pos0] == '\n' || (text_bytes[pos0] == '\r' && pos0 + 1 < end0 && text_bytes[pos0 + 1] == '\n')))
        ++pos0;
    if (end0 > pos0 && text_bytes[end0 - 1] == '\r') --end0;
    if (r == n_terms) {              // unterminated last record of the file
        if (pos0 >= end0) return;
        if (st->open_quote) { report(st, record0 + r, CSV_EQUOTE); return; }
        st->tail_valid = 1;
    }
    if (end0 - pos0 > 0x3fffffff) { report(st, record0 + r, CSV_EFIELDS); return; }
    const Bytes text = {text_bytes + pos0};
    int pos = 0;
    const int end = (int)(end0 - pos0);
    Field sel[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) sel[c] = Field{0, 0, 0};
    int32_t field = 0;
    bool bad_quote = false;
    for (;;) {
        Field f;
        f.esc = 0;
        if (pos < end && text[pos] == '"') {
            f.b = ++pos;
            for (;;) {
                if (pos >= end) { bad_quote = true; break; }
                if (text[pos] == '"') {
                    if (pos + 1 < end && text[pos + 1] == '"') { pos += 2; f.esc = 1; continue; }
                    break;
                }
                ++pos;
            }
            f.e = pos;
            if (!bad_quote) {
                ++pos;
                if (pos < end && text[pos] != ',') bad_quote = true;
            }
        } else {
            f.b = pos;
            while (pos < end && text[pos] != ',') ++pos;
            f.e = pos;
        }
        if (bad_quote) break;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (sc.sel[c] == field) sel[c] = f;
        ++field;
        if (pos >= end) break;
        ++pos;   // the comma
    }
    if (bad_quote) { report(st, record0 + r, CSV_EQUOTE); return; }
    if (field != sc.n_cols) { report(st, record0 + r, CSV_EFIELDS); return; }
    int32_t ids[2];
    float vals[2];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const Field f = sel[c];
        if (sc.kind[c] == GLOVE_CSV_TOKEN) {
            ids[c & 1] = lookup_token(text, f, table, slots, vb, voff);
        } else if (sc.kind[c] == GLOVE_CSV_INT) {
            int i = f.b;
            int64_t v = 0;
            bool neg = false, ok = f.e > f.b && !f.esc;
            if (ok && (text[i] == '-' || text[i] == '+')) { neg = text[i] == '-'; ++i; ok = i < f.e; }
            for (; ok && i < f.e; ++i) {
                const unsigned dgt = (unsigned)(text[i] - '0');
                ok = dgt <= 9 && v < (1ll << 40);
                v = v * 10 + dgt;
            }
            if (!ok) { report(st, record0 + r, CSV_EINT); return; }
            if (neg) v = -v;
            if (v < 0 || v >= n_vocab) { report(st, record0 + r, CSV_ERANGE); return; }
            ids[c & 1] = (int32_t)v;
        } else {
            uint32_t bits;
            const Bytes fb = {text.p + f.b};
            const int rc = f.esc ? (int)STRTOF_BAD : parse_f32([fb](int i) { return (int)fb[i]; }, f.e - f.b, bits);
            if (rc != STRTOF_OK) { report(st, record0 + r, rc == STRTOF_TOO_LONG ? CSV_EDIGITS : CSV_EFLOAT); return; }
            vals[c & 1] = __uint_as_float(bits);
        }
    }
    out_row[r] = ids[0];
    out_col[r] = ids[1];
    out_a[r] = vals[0];
    out_b[r] = vals[1];
}

}  // namespace glove

using namespace glove;

extern "C" {

int64_t glove_vocab_slots(int64_t n_vocab) {
    int64_t s = 1024;
    while (s < 2 * n_vocab) s <<= 1;
    return s;
}

int glove_vocab_build(int32_t *table, int64_t slots, const uint8_t *vocab_bytes, const int64_t *vocab_off,
                      int64_t n_vocab, void *stream) {
    GLOVE_REQUIRE(table && vocab_bytes && vocab_off && n_vocab > 0, "glove_vocab_build: bad arguments");
    GLOVE_REQUIRE(slots >= 2 * n_vocab && (slots & (slots - 1)) == 0 && n_vocab < (1ll << 31),
                  "glove_vocab_build: slots must be a power of two >= 2 * n_vocab");
    GLOVE_CHECK_CUDA(cudaMemsetAsync(table, 0xff, sizeof(int32_t) * slots, (cudaStream_t)stream));
    vocab_build_kernel<<<(unsigned)((n_vocab + 255) / 256), 256, 0, (cudaStream_t)stream>>>(table, slots, vocab_bytes,
                                                                                            vocab_off, n_vocab);
    GLOVE_CHECK_LAUNCH();
    return GLOVE_OK;
}

size_t glove_csv_workspace_bytes(int64_t nbytes) { return csv_ws_view(nullptr, nbytes < 1 ? 1 : nbytes).bytes; }

int glove_csv_index(const uint8_t *text, int64_t nbytes, void *workspace, size_t workspace_bytes, int64_t *n_records_host,
                    void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GLOVE_REQUIRE(text && workspace && n_records_host && nbytes > 0, "glove_csv_index: bad arguments");
    GLOVE_REQUIRE(((uintptr_t)text & 15) == 0, "glove_csv_index: text must be 16-byte aligned");
    GLOVE_REQUIRE(nbytes < (1ll << 40), "glove_csv_index: chunk too large");
    const CsvWs w = csv_ws_view(workspace, nbytes);
    if (workspace_bytes < w.bytes)
        return set_error(GLOVE_EWORKSPACE, "glove_csv_index: workspace %zu < required %zu", workspace_bytes, w.bytes);
    const int64_t nt = n_tiles_of(nbytes);
    tile_stats_kernel<<<(unsigned)nt, kTileThreads, 0, stream>>>(text, nbytes, w.tile_raw);
    GLOVE_CHECK_LAUNCH();
    size_t tb = w.cub_bytes;
    GLOVE_CHECK_CUDA(cub::DeviceScan::InclusiveScan(w.cub_temp, tb, w.tile_raw, w.tile, QCombine(), (int)nt, stream));
    init_status_kernel<<<1, 1, 0, stream>>>(w.status, w.tile, nt);
    GLOVE_CHECK_LAUNCH();
    CsvStatus st;
    GLOVE_CHECK_CUDA(cudaMemcpyAsync(&st, w.status, sizeof(st), cudaMemcpyDeviceToHost, stream));
    GLOVE_CHECK_CUDA(cudaStreamSynchronize(stream));
    *n_records_host = st.n_terms;
    return GLOVE_OK;
}

int glove_csv_parse(const uint8_t *text, int64_t nbytes, int32_t final_chunk, void *workspace, size_t workspace_bytes,
                    const glove_csv_schema *schema, const int32_t *vocab_table, int64_t vocab_slots,
                    const uint8_t *vocab_bytes, const int64_t *vocab_off, int64_t n_vocab, int64_t *row_ends,
                    int32_t *row, int32_t *col, float *colA, float *colB, int64_t capacity, int64_t first_record,
                    int64_t *n_rows_host, int64_t *consumed_host, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GLOVE_REQUIRE(text && workspace && schema && row_ends && row && col && colA && colB && n_rows_host && consumed_host,
                  "glove_csv_parse: null pointer");
    GLOVE_REQUIRE(nbytes > 0 && capacity > 0 && n_vocab > 0, "glove_csv_parse: bad sizes");
    Schema sc;
    sc.n_cols = schema->n_cols;
    bool need_vocab = false;
    for (int c = 0; c < 4; ++c) {
        sc.sel[c] = schema->column[c];
        sc.kind[c] = schema->kind[c];
        GLOVE_REQUIRE(sc.sel[c] >= 0 && sc.sel[c] < sc.n_cols, "glove_csv_parse: selected column %d out of range", sc.sel[c]);
        if (c < 2) GLOVE_REQUIRE(sc.kind[c] == GLOVE_CSV_TOKEN || sc.kind[c] == GLOVE_CSV_INT, "glove_csv_parse: id column kind");
        else GLOVE_REQUIRE(sc.kind[c] == GLOVE_CSV_FLOAT, "glove_csv_parse: value column kind");
        need_vocab |= sc.kind[c] == GLOVE_CSV_TOKEN;
    }
    if (need_vocab) GLOVE_REQUIRE(vocab_table && vocab_bytes && vocab_off && vocab_slots > 0, "glove_csv_parse: vocab table missing");
    const CsvWs w = csv_ws_view(workspace, nbytes);
    if (workspace_bytes < w.bytes)
        return set_error(GLOVE_EWORKSPACE, "glove_csv_parse: workspace %zu < required %zu", workspace_bytes, w.bytes);
    const int64_t nt = n_tiles_of(nbytes);
    row_ends_kernel<<<(unsigned)nt, kTileThreads, 0, stream>>>(text, nbytes, w.tile, nt, row_ends, capacity, w.status);
    GLOVE_CHECK_LAUNCH();
    parse_rows_kernel<<<(unsigned)((capacity + 127) / 128), 128, 0, stream>>>(
        text, nbytes, final_chunk, row_ends, w.status, sc, vocab_table, vocab_slots, vocab_bytes, vocab_off, n_vocab, row,
        col, colA, colB, capacity, first_record);
    GLOVE_CHECK_LAUNCH();
    CsvStatus st;
    GLOVE_CHECK_CUDA(cudaMemcpyAsync(&st, w.status, sizeof(st), cudaMemcpyDeviceToHost, stream));
    GLOVE_CHECK_CUDA(cudaStreamSynchronize(stream));
    if (st.n_terms + (final_chunk ? 1 : 0) > capacity)
        return set_error(GLOVE_EINVAL, "glove_csv_parse: capacity %lld < %lld records (+1 for a final chunk)",
                         (long long)capacity, (long long)st.n_terms);
    if (st.first_error != ~0ull) {
        static const char *what[] = {"", "wrong number of fields", "malformed quoting", "not a number",
                                     "more than 19 significant digits and the tail decides the rounding", "not an integer id",
                                     "id outside [0, vocab size)"};
        const int code = (int)(st.first_error & 0xff);
        return set_error(GLOVE_EINVAL, "interaction csv: record %lld: %s", (long long)(st.first_error >> 8),
                         code >= 1 && code <= 6 ? what[code] : "parse error");
    }
    if (final_chunk && st.open_quote)
        return set_error(GLOVE_EINVAL, "interaction csv: unterminated quoted field at end of file");
    *n_rows_host = st.n_terms + st.tail_valid;
    *consumed_host = final_chunk ? nbytes : st.last_end + 1;
    return GLOVE_OK;
}

// host-side entry to the SAME decimal -> float32 routine the kernels use (CPU tests check it against exact rounding)
int glove_parse_float32(const char *text, int32_t n, float *out) {
    GLOVE_REQUIRE(text && out && n >= 0, "glove_parse_float32: bad arguments");
    uint32_t bits;
    const int rc = parse_f32([text](int i) { return (int)(uint8_t)text[i]; }, n, bits);
    if (rc != STRTOF_OK) return set_error(GLOVE_EINVAL, "glove_parse_float32: %s", rc == STRTOF_TOO_LONG ? "too many digits" : "not a number");
    memcpy(out, &bits, 4);
    return GLOVE_OK;
}

}  // extern "C"
