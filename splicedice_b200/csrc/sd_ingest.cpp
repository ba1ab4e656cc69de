// sd_ingest.cpp -- host-side sample-file ingest for `quant` (no CUDA).
//
// Replaces the two per-line python passes of the reference over every sample file:
//   SPLICEDICE.getAllJunctions   (SPLICEDICE.py:147-228)  -> sd_ingest_collect
//   SPLICEDICE.getJunctionCounts (SPLICEDICE.py:257-295)  -> sd_ingest_counts
// with the same admission rules per file type (STAR SJ.out.tab: strict length bounds, strand
// 1/2 only, motif set, unique(+multi) >= minUnique; bam_to_junc_bed BED: filters only when the
// annotation tag is "?"; plain BED / leafcutter: score and inclusive length bounds; strand must
// be + or -) and the same count semantics (assignment, last duplicate line wins; a low cell is
// any BED-family line scored below minUnique).  Files are memory-mapped and parsed on several
// threads; the junction union and the junction -> row index are open-addressing hash tables.
// Malformed lines are reported with file and line number (the reference raises
// ValueError / IndexError / KeyError at the same places).
#include <errno.h>
#include <fcntl.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "sd_common.cuh"

namespace {

struct Key {
    uint64_t a;   // chrom id << 8 | strand char
    uint64_t b;   // left << 32 | right
    bool operator==(const Key &o) const { return a == o.a && b == o.b; }
};
inline uint64_t mix(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
inline uint64_t hash_key(const Key &k) { return mix(k.a * 0x9E3779B97F4A7C15ull ^ mix(k.b)); }

// open addressing, linear probing; value = insertion order (set) or row (index)
struct Table {
    std::vector<Key> keys;
    std::vector<int64_t> vals;      // -1 = empty
    size_t used = 0, mask = 0;
    void reserve(size_t n)
    {
        size_t cap = 1024;
        while (cap < n * 2) cap <<= 1;
        if (cap <= keys.size()) return;
        std::vector<Key> ok;
        std::vector<int64_t> ov;
        ok.swap(keys); ov.swap(vals);
        keys.assign(cap, Key{0, 0});
        vals.assign(cap, -1);
        mask = cap - 1;
        used = 0;
        for (size_t i = 0; i < ok.size(); ++i)
            if (ov[i] >= 0) put(ok[i], ov[i]);
    }
    int64_t *find(const Key &k)
    {
        if (keys.empty()) return nullptr;
        for (size_t i = hash_key(k) & mask;; i = (i + 1) & mask) {
            if (vals[i] < 0) return nullptr;
            if (keys[i] == k) return &vals[i];
        }
    }
    // returns true if inserted
    bool put(const Key &k, int64_t v)
    {
        if ((used + 1) * 2 > keys.size()) reserve(std::max<size_t>(used * 2, 512));
        for (size_t i = hash_key(k) & mask;; i = (i + 1) & mask) {
            if (vals[i] < 0) { keys[i] = k; vals[i] = v; ++used; return true; }
            if (keys[i] == k) return false;
        }
    }
};

struct Filter {
    int32_t max_length, min_length, min_overhang, min_unique;
    int32_t no_multimap, low_coverage_nan;
    uint32_t motif_mask;          // bit m set: STAR motif code m is admitted
    double min_entropy;
};

struct Ingest {
    std::mutex mu;
    std::unordered_map<std::string, int32_t> chrom_id;
    std::vector<std::string> chrom_names;
    Table set;                            // union of admitted junctions, value = insertion index
    std::vector<Key> order;               // junctions in insertion order
    Table index;                          // junction -> output row
    int32_t intern(const std::string &name)
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = chrom_id.find(name);
        if (it != chrom_id.end()) return it->second;
        int32_t id = (int32_t)chrom_names.size();
        chrom_names.push_back(name);
        chrom_id.emplace(name, id);
        return id;
    }
};

struct Mapped {
    const char *data = nullptr;
    size_t size = 0;
    int fd = -1;
    bool open(const char *path, std::string *err)
    {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) { *err = std::string("cannot open ") + path + ": " + strerror(errno); return false; }
        struct stat st;
        if (fstat(fd, &st) != 0) { *err = std::string("cannot stat ") + path; return false; }
        size = (size_t)st.st_size;
        if (size == 0) return true;
        void *p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (p == MAP_FAILED) { *err = std::string("cannot map ") + path + ": " + strerror(errno); return false; }
        madvise(p, size, MADV_SEQUENTIAL);
        data = static_cast<const char *>(p);
        return true;
    }
    ~Mapped()
    {
        if (data) munmap(const_cast<char *>(data), size);
        if (fd >= 0) close(fd);
    }
};

struct Field {
    const char *p;
    size_t n;
};

// python: line.rstrip().split("\t")
inline int split_tabs(const char *p, const char *end, Field *out, int max_fields)
{
    while (end > p && (end[-1] == ' ' || end[-1] == '\t' || end[-1] == '\r' || end[-1] == '\n' || end[-1] == '\f' ||
                       end[-1] == '\v'))
        --end;
    int n = 0;
    const char *s = p;
    for (const char *q = p;; ++q) {
        if (q == end || *q == '\t') {
            if (n < max_fields) out[n] = Field{s, (size_t)(q - s)};
            ++n;
            if (q == end) break;
            s = q + 1;
        }
    }
    return n;
}

inline bool to_int(const Field &f, int64_t *out)
{
    const char *p = f.p, *e = f.p + f.n;
    while (p < e && (*p == ' ')) ++p;
    while (e > p && e[-1] == ' ') --e;
    if (p == e) return false;
    bool neg = false;
    if (*p == '-' || *p == '+') { neg = *p == '-'; ++p; }
    if (p == e) return false;
    int64_t v = 0;
    for (; p < e; ++p) {
        if (*p < '0' || *p > '9') return false;
        v = v * 10 + (*p - '0');
        if (v > (int64_t(1) << 40)) return false;
    }
    *out = neg ? -v : v;
    return true;
}

inline bool to_double(const char *p, size_t n, double *out)
{
    char buf[64];
    if (n == 0 || n >= sizeof buf) return false;
    memcpy(buf, p, n);
    buf[n] = 0;
    char *end = nullptr;
    *out = strtod(buf, &end);
    return end == buf + n;
}

// thread-local chromosome cache in front of the shared interner
struct ChromCache {
    Ingest *owner;
    std::unordered_map<std::string, int32_t> local;
    std::string scratch;
    int32_t id(const Field &f)
    {
        scratch.assign(f.p, f.n);
        auto it = local.find(scratch);
        if (it != local.end()) return it->second;
        int32_t v = owner->intern(scratch);
        local.emplace(scratch, v);
        return v;
    }
};

enum { kTypeSJ = 0, kTypeTaggedBed = 1, kTypeBed = 2 };

struct Parsed {
    Key key;
    int64_t score;
    bool admitted;     // passes this file type's filter (pass 1)
    bool bed_family;
};

// one line -> junction key, score, admission.  Returns false with `err` set on a malformed line.
inline bool parse_line(int type, const Filter &flt, ChromCache &chroms, const char *p, const char *end, Parsed *out,
                       std::string *err)
{
    Field f[16];
    const int n = split_tabs(p, end, f, 16);
    int64_t left, right, score;
    char strand;
    out->bed_family = type != kTypeSJ;
    if (type == kTypeSJ) {
        int64_t start, motif, uniq, multi = 0;
        if (n < 8) { *err = "expected at least 8 tab-separated fields"; return false; }
        if (!to_int(f[1], &start) || !to_int(f[2], &right) || !to_int(f[4], &motif) || !to_int(f[6], &uniq) ||
            !to_int(f[7], &multi)) { *err = "non-integer field"; return false; }
        left = start - 1;
        if (f[3].n != 1 || (f[3].p[0] != '0' && f[3].p[0] != '1' && f[3].p[0] != '2' && f[3].p[0] != '+' && f[3].p[0] != '-')) {
            *err = "unknown strand code";
            return false;
        }
        const char c = f[3].p[0];
        strand = c == '1' ? '+' : c == '2' ? '-' : c;      // '0' stays '0'
        score = flt.no_multimap ? uniq : uniq + multi;
        const int64_t span = right - left;
        out->admitted = span < flt.max_length && span > flt.min_length && strand != '0' && score >= flt.min_unique &&
                        motif >= 0 && motif < 32 && ((flt.motif_mask >> motif) & 1u);
    } else {
        if (n < 6) { *err = "expected at least 6 tab-separated fields"; return false; }
        if (!to_int(f[1], &left) || !to_int(f[2], &right) || !to_int(f[4], &score)) { *err = "non-integer field"; return false; }
        strand = f[5].n == 1 ? f[5].p[0] : '?';
        const int64_t span = right - left;
        bool ok = true;
        if (type == kTypeTaggedBed) {
            // e:<Hl>:<Hr>;o:<overhang>;m:<motif>;a:<gene or ?>
            Field tag[4];
            int nt = 0;
            const char *s = f[3].p, *e = f[3].p + f[3].n;
            for (const char *q = s;; ++q) {
                if (q == e || *q == ';') {
                    if (nt < 4) tag[nt] = Field{s, (size_t)(q - s)};
                    ++nt;
                    if (q == e) break;
                    s = q + 1;
                }
            }
            if (nt < 4) { *err = "name field without e:/o:/m:/a: tags"; return false; }
            const char *colon = (const char *)memchr(tag[3].p, ':', tag[3].n);
            if (!colon) { *err = "annotation tag without ':'"; return false; }
            const char *vbeg = colon + 1, *vend = tag[3].p + tag[3].n;
            const char *c2 = (const char *)memchr(vbeg, ':', (size_t)(vend - vbeg));
            if (c2) vend = c2;
            if (vend - vbeg == 1 && *vbeg == '?') {
                // unannotated: every filter applies
                const char *c1 = (const char *)memchr(tag[1].p, ':', tag[1].n);
                int64_t overhang;
                if (!c1) { *err = "overhang tag without ':'"; return false; }
                const char *oe = tag[1].p + tag[1].n;
                const char *o2 = (const char *)memchr(c1 + 1, ':', (size_t)(oe - c1 - 1));
                if (!to_int(Field{c1 + 1, (size_t)((o2 ? o2 : oe) - c1 - 1)}, &overhang)) { *err = "bad overhang"; return false; }
                const char *e1 = (const char *)memchr(tag[0].p, ':', tag[0].n);
                const char *te = tag[0].p + tag[0].n;
                const char *e2 = e1 ? (const char *)memchr(e1 + 1, ':', (size_t)(te - e1 - 1)) : nullptr;
                if (!e1 || !e2) { *err = "entropy tag needs two values"; return false; }
                const char *e3 = (const char *)memchr(e2 + 1, ':', (size_t)(te - e2 - 1));
                double hl, hr;
                if (!to_double(e1 + 1, (size_t)(e2 - e1 - 1), &hl) || !to_double(e2 + 1, (size_t)((e3 ? e3 : te) - e2 - 1), &hr)) {
                    *err = "bad entropy value";
                    return false;
                }
                ok = !(score < flt.min_unique || span > flt.max_length || span < flt.min_length ||
                       overhang < flt.min_overhang || hl < flt.min_entropy || hr < flt.min_entropy);
            }
        } else {
            ok = !(score < flt.min_unique || span > flt.max_length || span < flt.min_length);
        }
        out->admitted = ok && (strand == '+' || strand == '-');
    }
    if (left < 0 || right < 0 || left >= (int64_t(1) << 31) || right >= (int64_t(1) << 31)) {
        *err = "coordinate outside [0, 2^31)";
        return false;
    }
    if (score < 0 || score >= (int64_t(1) << 31)) { *err = "score outside [0, 2^31)"; return false; }
    // junction identity as the reference's tuple: SJ strand '0' and BED strands other than +/- can
    // never be in the index, the key only has to be distinct from every admitted key
    const uint64_t sbyte = (f[type == kTypeSJ ? 3 : 5].n == 1 || type == kTypeSJ) ? (uint64_t)(unsigned char)strand : 0xFFu;
    out->key = Key{((uint64_t)(uint32_t)chroms.id(f[0]) << 8) | sbyte, ((uint64_t)left << 32) | (uint64_t)right};
    out->score = score;
    return true;
}

template <class Fn>
bool for_each_line(const char *path, Fn fn, std::string *err)
{
    Mapped m;
    if (!m.open(path, err)) return false;
    const char *p = m.data, *end = m.data + m.size;
    int64_t line_no = 0;
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = nl ? nl : end;
        ++line_no;
        std::string why;
        if (!fn(p, le, &why)) {
            *err = std::string(path) + ":" + std::to_string(line_no) + ": " + why;
            return false;
        }
        p = nl ? nl + 1 : end;
    }
    return true;
}

int run_threads(int n_items, int n_threads, const std::function<bool(int, std::string *)> &job, std::string *first_err)
{
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    n_threads = std::max(1, std::min(n_threads, n_items));
    std::atomic<int> next{0};
    std::atomic<bool> failed{false};
    std::mutex emu;
    auto worker = [&]() {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= n_items || failed.load()) return;
            std::string err;
            if (!job(i, &err)) {
                std::lock_guard<std::mutex> lock(emu);
                if (!failed.exchange(true)) *first_err = err;
                return;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto &th : pool) th.join();
    return failed.load() ? SD_ERR_INVALID : SD_OK;
}

}  // namespace

extern "C" {


void *sd_ingest_create(void) { return new Ingest(); }
void sd_ingest_destroy(void *h) { delete static_cast<Ingest *>(h); }

static Filter to_filter(const sd_quant_filter *f)
{
    return Filter{f->max_length, f->min_length, f->min_overhang, f->min_unique, f->no_multimap,
                  f->low_coverage_nan, f->motif_mask, f->min_entropy};
}

// Pass 1 over n_files sample files (types[i]: 0 SJ.out.tab, 1 bam_to_junc_bed BED, 2 plain BED /
// leafcutter; any other value: the file is opened and ignored, as the reference does for .bam /
// unknown suffixes).  Adds the admitted junctions to the handle's union.
int sd_ingest_collect(void *handle, int32_t n_files, const char *const *paths, const int32_t *types,
                      const sd_quant_filter *filter, int32_t n_threads)
{
    SD_REQUIRE(handle && filter && n_files >= 0 && (n_files == 0 || (paths && types)), "sd_ingest_collect: bad arguments");
    Ingest *ing = static_cast<Ingest *>(handle);
    const Filter flt = to_filter(filter);
    std::vector<std::vector<Key>> found((size_t)n_files);
    std::string err;
    int rc = run_threads(n_files, n_threads, [&](int i, std::string *e) {
        if (types[i] < 0 || types[i] > 2) {
            Mapped m;
            return m.open(paths[i], e);
        }
        ChromCache chroms{ing, {}, {}};
        Table seen;
        std::vector<Key> &mine = found[(size_t)i];
        return for_each_line(paths[i], [&](const char *p, const char *le, std::string *why) {
            Parsed pr;
            if (!parse_line(types[i], flt, chroms, p, le, &pr, why)) return false;
            if (pr.admitted && seen.put(pr.key, 0)) mine.push_back(pr.key);
            return true;
        }, e);
    }, &err);
    if (rc != SD_OK) return sd::fail(SD_ERR_INVALID, "%s", err.c_str());
    for (auto &v : found)
        for (const Key &k : v)
            if (ing->set.put(k, (int64_t)ing->order.size())) ing->order.push_back(k);
    return SD_OK;
}

int64_t sd_ingest_junction_count(void *handle) { return handle ? (int64_t)static_cast<Ingest *>(handle)->order.size() : -1; }
int32_t sd_ingest_chrom_count(void *handle) { return handle ? (int32_t)static_cast<Ingest *>(handle)->chrom_names.size() : -1; }

// name of chromosome id `id` (NUL-terminated, owned by the handle)
const char *sd_ingest_chrom_name(void *handle, int32_t id)
{
    Ingest *ing = static_cast<Ingest *>(handle);
    if (!ing || id < 0 || id >= (int32_t)ing->chrom_names.size()) return nullptr;
    return ing->chrom_names[(size_t)id].c_str();
}

// the union as arrays, in insertion order: chromosome id, left, right, strand character
int sd_ingest_export(void *handle, int32_t *chrom_id, int32_t *left, int32_t *right, int8_t *strand)
{
    SD_REQUIRE(handle && chrom_id && left && right && strand, "sd_ingest_export: null pointer");
    Ingest *ing = static_cast<Ingest *>(handle);
    for (size_t i = 0; i < ing->order.size(); ++i) {
        const Key &k = ing->order[i];
        chrom_id[i] = (int32_t)(k.a >> 8);
        strand[i] = (int8_t)(k.a & 0xFF);
        left[i] = (int32_t)(k.b >> 32);
        right[i] = (int32_t)(k.b & 0xFFFFFFFFu);
    }
    return SD_OK;
}

// junction -> output row, for the n junctions given as (chromosome NAME index into `names`, ...)
int sd_ingest_index(void *handle, int64_t n, const char *const *chrom_names, const int32_t *chrom_of, const int32_t *left,
                    const int32_t *right, const int8_t *strand, const int32_t *row)
{
    SD_REQUIRE(handle && n >= 0 && (n == 0 || (chrom_names && chrom_of && left && right && strand && row)),
               "sd_ingest_index: bad arguments");
    Ingest *ing = static_cast<Ingest *>(handle);
    ing->index = Table();
    ing->index.reserve((size_t)n);
    std::unordered_map<int32_t, int32_t> ids;
    for (int64_t i = 0; i < n; ++i) {
        auto it = ids.find(chrom_of[i]);
        int32_t cid;
        if (it == ids.end()) {
            cid = ing->intern(chrom_names[chrom_of[i]]);
            ids.emplace(chrom_of[i], cid);
        } else
            cid = it->second;
        const Key k{((uint64_t)(uint32_t)cid << 8) | (uint64_t)(uint8_t)strand[i],
                    ((uint64_t)(uint32_t)left[i] << 32) | (uint64_t)(uint32_t)right[i]};
        ing->index.put(k, row[i]);
    }
    return SD_OK;
}

// Pass 2: counts[row, samples[i]] = score for every line of file i whose junction is indexed
// (last duplicate line wins); low_mask[row, sample] = 1 for BED-family lines scored below
// minUnique when filter->low_coverage_nan is set (low_mask may be NULL otherwise).
int sd_ingest_counts(void *handle, int32_t n_files, const char *const *paths, const int32_t *types, const int32_t *samples,
                     const sd_quant_filter *filter, int32_t *counts, int64_t ld_counts, uint8_t *low_mask, int64_t ld_mask,
                     int32_t n_threads)
{
    SD_REQUIRE(handle && filter && n_files >= 0 && counts && (n_files == 0 || (paths && types && samples)),
               "sd_ingest_counts: bad arguments");
    Ingest *ing = static_cast<Ingest *>(handle);
    const Filter flt = to_filter(filter);
    SD_REQUIRE(!flt.low_coverage_nan || low_mask, "sd_ingest_counts: lowCoverageNan needs a mask");
    std::string err;
    int rc = run_threads(n_files, n_threads, [&](int i, std::string *e) {
        if (types[i] < 0 || types[i] > 2) {
            Mapped m;
            return m.open(paths[i], e);
        }
        ChromCache chroms{ing, {}, {}};
        const int32_t s = samples[i];
        return for_each_line(paths[i], [&](const char *p, const char *le, std::string *why) {
            Parsed pr;
            if (!parse_line(types[i], flt, chroms, p, le, &pr, why)) return false;
            int64_t *row = ing->index.find(pr.key);
            if (!row) return true;
            counts[*row * ld_counts + s] = (int32_t)pr.score;
            if (flt.low_coverage_nan && pr.bed_family && pr.score < flt.min_unique) low_mask[*row * ld_mask + s] = 1;
            return true;
        }, e);
    }, &err);
    if (rc != SD_OK) return sd::fail(SD_ERR_INVALID, "%s", err.c_str());
    return SD_OK;
}

}  // extern "C"
