// Stage-boundary text matrices: the reference checkpoints every stage
// with np.savetxt ('%.18e' / '%i', space separated) and can resume from np.loadtxt files
// (wisp_w_mdanalysis.py:108-133).  At N = 2,000 the 3N x 3N covariance alone is 36 M numbers
// (~0.9 GB of text), which np.savetxt writes at ~1 M numbers/s.  These host routines produce
// byte-identical files (C's "%.18e" and Python's '%.18e' are both correctly rounded) by
// formatting row blocks on all host threads, and parse them back the same way.
#include "common.cuh"
#include <algorithm>
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <thread>

namespace {

int n_threads_for(int64_t work) {
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 4;
    int64_t t = std::min<int64_t>(hw, std::max<int64_t>(1, work / 65536));
    return (int)std::min<int64_t>(t, 64);
}

template <typename T, typename F>
int save_matrix(const char* path, const T* a, int64_t rows, int64_t cols, const char* header, F fmt_one,
                size_t bytes_per_value) {
    FILE* f = fopen(path, "wb");
    WISP_CHECK(f, "savetxt: cannot open %s: %s", path, strerror(errno));
    if (header && header[0]) fprintf(f, "# %s\n", header);
    const int nt = n_threads_for(rows * cols);
    // bounded memory: format blocks of rows, nt blocks at a time
    const int64_t rows_per_block = std::max<int64_t>(1, (int64_t)((size_t)8 << 20) / std::max<size_t>(1, cols * bytes_per_value));
    std::vector<std::string> bufs(nt);
    int rc = 0;
    for (int64_t r0 = 0; r0 < rows && rc == 0; r0 += rows_per_block * nt) {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) {
            const int64_t b0 = r0 + t * rows_per_block;
            const int64_t b1 = std::min(rows, b0 + rows_per_block);
            bufs[t].clear();
            if (b0 >= b1) continue;
            th.emplace_back([&, t, b0, b1]() {
                std::string& s = bufs[t];
                s.reserve((size_t)(b1 - b0) * cols * bytes_per_value);
                char tmp[64];
                for (int64_t r = b0; r < b1; ++r) {
                    for (int64_t c = 0; c < cols; ++c) {
                        const int n = fmt_one(tmp, a[r * cols + c]);
                        s.append(tmp, n);
                        s.push_back(c + 1 == cols ? '\n' : ' ');
                    }
                }
            });
        }
        for (auto& x : th) x.join();
        for (int t = 0; t < nt; ++t)
            if (!bufs[t].empty() && fwrite(bufs[t].data(), 1, bufs[t].size(), f) != bufs[t].size()) rc = 1;
    }
    if (fclose(f) != 0) rc = 1;
    WISP_CHECK(rc == 0, "savetxt: write to %s failed: %s", path, strerror(errno));
    return 0;
}

}  // namespace

extern "C" int wisp_savetxt_f64(const char* path, const double* a, int64_t rows, int64_t cols,
                                const char* header) {
    WISP_CHECK(path && a && rows >= 0 && cols >= 1, "savetxt_f64: bad argument");
    return save_matrix(path, a, rows, cols, header,
                       [](char* buf, double v) {
                           // std::to_chars(scientific, 18) is specified to print what printf("%.18e") prints,
                           // correctly rounded (Ryu-printf in libstdc++), at ~10x the speed of glibc's printf
                           if (std::isfinite(v)) {
                               auto r = std::to_chars(buf, buf + 64, v, std::chars_format::scientific, 18);
                               if (r.ec == std::errc()) return (int)(r.ptr - buf);
                           }
                           return snprintf(buf, 64, "%.18e", v);
                       }, 26);
}

extern "C" int wisp_savetxt_i64(const char* path, const int64_t* a, int64_t rows, int64_t cols,
                                const char* header) {
    WISP_CHECK(path && a && rows >= 0 && cols >= 1, "savetxt_i64: bad argument");
    return save_matrix(path, a, rows, cols, header,
                       [](char* buf, int64_t v) { return snprintf(buf, 64, "%lld", (long long)v); }, 3);
}

// Parse a whitespace-separated numeric text file ('#' comment lines skipped).  Call once with
// out == NULL to get the shape, then with a buffer of rows*cols doubles.
extern "C" int wisp_loadtxt_f64(const char* path, double* out, int64_t capacity, int64_t* rows_out,
                                int64_t* cols_out) {
    WISP_CHECK(path && rows_out && cols_out, "loadtxt_f64: bad argument");
    FILE* f = fopen(path, "rb");
    WISP_CHECK(f, "loadtxt: cannot open %s: %s", path, strerror(errno));
    fseek(f, 0, SEEK_END);
    const long long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::string text((size_t)size, '\0');
    const size_t got = fread(&text[0], 1, (size_t)size, f);
    fclose(f);
    WISP_CHECK((long long)got == size, "loadtxt: short read on %s", path);
    // data lines as [start, end): every row is parsed strictly inside its own line, so a short row can
    // never borrow numbers from the next one (strtod alone skips '\n' as white space)
    std::vector<std::pair<size_t, size_t>> lines;
    size_t pos = 0;
    while (pos < text.size()) {
        size_t end = text.find('\n', pos);
        if (end == std::string::npos) end = text.size();
        size_t p = pos;
        while (p < end && (text[p] == ' ' || text[p] == '\t' || text[p] == '\r')) ++p;
        if (p < end && text[p] != '#') lines.push_back({p, end});
        pos = end + 1;
    }
    // one value: locale-independent (std::from_chars), accepts what np.savetxt writes (incl. inf / nan,
    // a leading '+'); returns the position after the value or nullptr
    auto parse = [](const char* p, const char* e, double* v) -> const char* {
        while (p < e && (*p == ' ' || *p == '\t' || *p == ',' || *p == '\r')) ++p;
        if (p >= e) return nullptr;
        if (*p == '+') ++p;
        const auto r = std::from_chars(p, e, *v);
        if (r.ec != std::errc() && r.ec != std::errc::result_out_of_range) return nullptr;
        return r.ptr;
    };
    auto count_cols = [&](size_t a0, size_t a1) -> int64_t {
        const char *p = text.c_str() + a0, *e = text.c_str() + a1;
        int64_t n = 0;
        double v;
        for (;;) {
            const char* q = parse(p, e, &v);
            if (!q) break;
            ++n;
            p = q;
        }
        while (p < e && (*p == ' ' || *p == '\t' || *p == ',' || *p == '\r')) ++p;
        return p == e ? n : -1;          // trailing junk: not a numeric row
    };
    const int64_t rows = (int64_t)lines.size();
    int64_t cols = 0;
    if (rows) {
        cols = count_cols(lines[0].first, lines[0].second);
        WISP_CHECK(cols > 0, "loadtxt: non-numeric first row in %s", path);
    }
    *rows_out = rows;
    *cols_out = cols;
    if (!out) return 0;
    WISP_CHECK(capacity >= rows * cols, "loadtxt: buffer too small (%lld < %lld)", (long long)capacity, (long long)(rows * cols));
    const int nt = n_threads_for(rows * cols);
    std::vector<int64_t> bad(nt, -1);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) {
        const int64_t r0 = rows * t / nt, r1 = rows * (t + 1) / nt;
        th.emplace_back([&, t, r0, r1]() {
            for (int64_t r = r0; r < r1; ++r) {
                const char *p = text.c_str() + lines[r].first, *e = text.c_str() + lines[r].second;
                int64_t c = 0;
                double v;
                for (;;) {
                    const char* q = parse(p, e, &v);
                    if (!q) break;
                    if (c < cols) out[r * cols + c] = v;
                    ++c;
                    p = q;
                }
                while (p < e && (*p == ' ' || *p == '\t' || *p == ',' || *p == '\r')) ++p;
                if (c != cols || p != e) { bad[t] = r; return; }      // too few, too many, or junk
            }
        });
    }
    for (auto& x : th) x.join();
    for (int t = 0; t < nt; ++t)
        WISP_CHECK(bad[t] < 0, "loadtxt: data row %lld of %s does not hold exactly %lld numbers (ragged or non-numeric row)",
                   (long long)bad[t] + 1, path, (long long)cols);
    return 0;
}
