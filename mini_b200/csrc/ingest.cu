// ingest.cu -- host-side graph ingest behind the C ABI: MatrixMarket files with the reference's
// semantics, and a binary CSR cache.  No device code; compiled with the rest of the
// library so that non-C++ hosts get the loader the header API has in include/gunrock/graph.hxx.
//
// What load_graph (gunrock/src/graph.hxx:96-223) does with a file, restated:
//   * lines starting with '%' are skipped; the next line is "rows cols entries" and num_nodes = rows (:104-112);
//   * an entry line "i j [w]" (1-based) becomes the arc (j-1) -> (i-1): CSR rows are built over the SECOND column
//     (:158-172), the stored column index is the first; w defaults to 1.0 (:120-127, _random_edge_value = false);
//   * undirected: every entry is followed (after all originals) by its reverse with the same weight (:130-137);
//   * arcs are ordered by (row, column); duplicates and self loops are kept; rows past the last arc get
//     offset = num_edges (:160 initialises offsets with num_edges).
// Differences on purpose: the reference's comparator returns true for equal keys (not a strict weak order: UB
// with duplicate entries, :139-157) -- here a stable sort keeps equal keys in file order; lines may be longer than
// its 100-byte buffer; nothing calls exit() -- errors are status codes.
#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include "b200_frontier.h"

namespace {

struct Arc {
    int32_t row, col;
    float w;
};

bool read_line(FILE *f, std::string &line) {
    line.clear();
    int c;
    while ((c = std::fgetc(f)) != EOF) {
        if (c == '\n') return true;
        if (c != '\r') line.push_back((char)c);
    }
    return !line.empty();
}

int alloc_csr(int64_t n, int64_t m, bool weights, b200_host_csr *out) {
    out->n = n;
    out->m = m;
    out->row_offsets = static_cast<uint32_t *>(std::malloc(sizeof(uint32_t) * (size_t)(n + 1)));
    out->col_indices = static_cast<int32_t *>(std::malloc(sizeof(int32_t) * (size_t)(m ? m : 1)));
    out->col_values = weights ? static_cast<float *>(std::malloc(sizeof(float) * (size_t)(m ? m : 1))) : nullptr;
    if (!out->row_offsets || !out->col_indices || (weights && !out->col_values)) {
        b200_host_csr_free(out);
        return B200_ERR_NOMEM;
    }
    return B200_OK;
}

constexpr char CACHE_MAGIC[8] = {'B', '2', '0', '0', 'C', 'S', 'R', '1'};
struct CacheHeader {
    char magic[8];
    int64_t n, m;
    uint32_t flags;      // bit 0: weights present
    uint32_t reserved;
};

uint64_t fnv1a(uint64_t h, const void *p, size_t bytes) {
    const unsigned char *b = static_cast<const unsigned char *>(p);
    for (size_t i = 0; i < bytes; ++i) {
        h ^= b[i];
        h *= 1099511628211ull;
    }
    return h;
}

}  // namespace

extern "C" {

int b200_host_csr_free(b200_host_csr *csr) {
    if (!csr) return B200_OK;
    std::free(csr->row_offsets);
    std::free(csr->col_indices);
    std::free(csr->col_values);
    csr->row_offsets = nullptr;
    csr->col_indices = nullptr;
    csr->col_values = nullptr;
    csr->n = csr->m = 0;
    return B200_OK;
}

static int mtx_load_impl(const char *path, int undirected, b200_host_csr *out);

int b200_mtx_load(const char *path, int undirected, b200_host_csr *out) {
    if (!path || !out) return B200_ERR_INVALID;
    std::memset(out, 0, sizeof *out);
    try {   // nothing may unwind through the extern "C" boundary
        return mtx_load_impl(path, undirected, out);
    } catch (const std::bad_alloc &) {
        b200_host_csr_free(out);
        return B200_ERR_NOMEM;
    } catch (...) {
        b200_host_csr_free(out);
        return B200_ERR_FORMAT;
    }
}

static int mtx_load_impl(const char *path, int undirected, b200_host_csr *out) {
    struct File {   // closed on every exit path, exceptions included
        FILE *f;
        ~File() { close(); }
        void close() { if (f) std::fclose(f); f = nullptr; }
    } fc{std::fopen(path, "r")};
    FILE *f = fc.f;
    if (!f) return B200_ERR_IO;
    std::string line;
    bool have = false;
    while ((have = read_line(f, line)) && !line.empty() && line[0] == '%') {}
    long long rows = 0, cols = 0, entries = 0;
    if (!have || std::sscanf(line.c_str(), "%lld %lld %lld", &rows, &cols, &entries) != 3 || rows < 1 || entries < 0 ||
        rows > (1ll << 31) || entries >= (1ll << 32) / (undirected ? 2 : 1)) {   // (compared before any multiply)
        fc.close();
        return B200_ERR_FORMAT;
    }
    std::vector<Arc> arcs;
    try {
        arcs.resize((size_t)entries * (undirected ? 2 : 1));
    } catch (const std::bad_alloc &) {
        fc.close();
        return B200_ERR_NOMEM;
    }
    for (long long e = 0; e < entries; ++e) {
        long long i = 0, j = 0;
        float w = 1.0f;
        int got = 0;
        do {   // (blank lines between entries are tolerated)
            if (!read_line(f, line)) {
                fc.close();
                return B200_ERR_FORMAT;
            }
        } while (line.empty());
        got = std::sscanf(line.c_str(), "%lld %lld %f", &i, &j, &w);
        if (got < 2 || i < 1 || j < 1 || i > rows || j > rows) {
            fc.close();
            return B200_ERR_FORMAT;
        }
        if (got == 2) w = 1.0f;
        arcs[(size_t)e] = Arc{(int32_t)(j - 1), (int32_t)(i - 1), w};
        if (undirected) arcs[(size_t)(e + entries)] = Arc{(int32_t)(i - 1), (int32_t)(j - 1), w};
    }
    fc.close();
    std::stable_sort(arcs.begin(), arcs.end(),
                     [](const Arc &a, const Arc &b) { return a.row != b.row ? a.row < b.row : a.col < b.col; });
    const int64_t n = rows, m = (int64_t)arcs.size();
    const int st = alloc_csr(n, m, true, out);
    if (st != B200_OK) return st;
    std::vector<uint32_t> count((size_t)n + 1, 0u);
    for (const Arc &a : arcs) count[(size_t)a.row + 1]++;
    uint32_t run = 0;
    for (int64_t v = 0; v <= n; ++v) {
        run += count[(size_t)v];
        out->row_offsets[v] = run;
    }
    for (int64_t e = 0; e < m; ++e) {
        out->col_indices[e] = arcs[(size_t)e].col;
        out->col_values[e] = arcs[(size_t)e].w;
    }
    return B200_OK;
}

int b200_csr_cache_write(const char *path, const b200_host_csr *csr) {
    if (!path || !csr || csr->n < 0 || csr->m < 0 || !csr->row_offsets || (csr->m && !csr->col_indices)) return B200_ERR_INVALID;
    FILE *f = std::fopen(path, "wb");
    if (!f) return B200_ERR_IO;
    CacheHeader h;
    std::memcpy(h.magic, CACHE_MAGIC, 8);
    h.n = csr->n;
    h.m = csr->m;
    h.flags = csr->col_values ? 1u : 0u;
    h.reserved = 0u;
    uint64_t sum = 14695981039346656037ull;
    bool ok = std::fwrite(&h, sizeof h, 1, f) == 1;
    auto put = [&](const void *p, size_t bytes) {
        if (bytes == 0) return;
        sum = fnv1a(sum, p, bytes);
        ok = ok && std::fwrite(p, 1, bytes, f) == bytes;
    };
    put(csr->row_offsets, sizeof(uint32_t) * (size_t)(csr->n + 1));
    put(csr->col_indices, sizeof(int32_t) * (size_t)csr->m);
    if (csr->col_values) put(csr->col_values, sizeof(float) * (size_t)csr->m);
    ok = ok && std::fwrite(&sum, sizeof sum, 1, f) == 1;
    ok = (std::fclose(f) == 0) && ok;
    return ok ? B200_OK : B200_ERR_IO;
}

int b200_csr_cache_read(const char *path, b200_host_csr *out) {
    if (!path || !out) return B200_ERR_INVALID;
    std::memset(out, 0, sizeof *out);
    FILE *f = std::fopen(path, "rb");
    if (!f) return B200_ERR_IO;
    CacheHeader h;
    if (std::fread(&h, sizeof h, 1, f) != 1 || std::memcmp(h.magic, CACHE_MAGIC, 8) != 0 || h.n < 0 || h.m < 0 ||
        h.n > (1ll << 31) || h.m >= (1ll << 32) || (h.flags & ~1u)) {
        std::fclose(f);
        return B200_ERR_FORMAT;
    }
    int st = alloc_csr(h.n, h.m, (h.flags & 1u) != 0, out);
    if (st != B200_OK) {
        std::fclose(f);
        return st;
    }
    uint64_t sum = 14695981039346656037ull, stored = 0;
    bool ok = true;
    auto get = [&](void *p, size_t bytes) {
        if (bytes == 0) return;
        ok = ok && std::fread(p, 1, bytes, f) == bytes;
        if (ok) sum = fnv1a(sum, p, bytes);
    };
    get(out->row_offsets, sizeof(uint32_t) * (size_t)(h.n + 1));
    get(out->col_indices, sizeof(int32_t) * (size_t)h.m);
    if (out->col_values) get(out->col_values, sizeof(float) * (size_t)h.m);
    ok = ok && std::fread(&stored, sizeof stored, 1, f) == 1 && stored == sum;
    std::fclose(f);
    // structural checks: a cache that passes them cannot send a kernel out of bounds
    if (ok) ok = out->row_offsets[0] == 0u && out->row_offsets[h.n] == (uint32_t)h.m;
    for (int64_t v = 0; ok && v < h.n; ++v) ok = out->row_offsets[v] <= out->row_offsets[v + 1];
    for (int64_t e = 0; ok && e < h.m; ++e) ok = out->col_indices[e] >= 0 && out->col_indices[e] < h.n;
    if (!ok) {
        b200_host_csr_free(out);
        return B200_ERR_FORMAT;
    }
    return B200_OK;
}

}  // extern "C"
