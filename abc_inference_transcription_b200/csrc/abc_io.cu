// abc_io.cu -- the reference's on-disk layouts written by the library, host code only.
//
// The reference formats every number with Julia's writedlm: compute_errors.jl:66-68 appends one 1 x G text row per
// particle (~65 KB of text each, 65 GB per 10^6 particles), process_error_files.jl:3-7 re-parses the whole text matrix
// into a JDF column store, abc_simulation.jl:47-61, 89-95 opens and appends eight files per trial.  Here the host hands
// the arrays the compute entry points returned straight to the library:
//   abc_format_float64        print(io, ::Float64): shortest round-trip digits (Ryu via std::to_chars), Julia's layout
//   abc_writedlm              writedlm(io, A) of a Float64 matrix, formatted on all host threads
//   abc_write_simulation      the seven files of abc_simulation.jl:47-61, 89-95 for a batch of trials
//   abc_write_accepted        data/posteriors/particles_<model>.txt (accepted_particles.jl:19-30, "0" sentinel)
//   abc_write_error_columns   one raw little-endian Float64 file per gene column x<g> (the layout idea of the .jdf the
//                             reference loads column by column, accepted_particles.jl:14-18; JDF.jl's own bytes are
//                             third-party and unpinned) + abc_read_error_column
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include <sys/stat.h>

#include "abc_internal.h"

// Julia's print(io, x::Float64) (base/ryu: shortest digits; fixed notation for 1e-5 <= |x| < 1e6, else d.ddde[-]x;
// always a decimal point; NaN, Inf, -Inf).  buf must hold 32 bytes.  Returns the length (no terminator counted).
static int jl_format(double x, char* out) {
    if (x != x) { memcpy(out, "NaN", 3); return 3; }
    if (std::isinf(x)) { if (x > 0) { memcpy(out, "Inf", 3); return 3; } memcpy(out, "-Inf", 4); return 4; }
    char* p = out;
    if (std::signbit(x)) { *p++ = '-'; x = -x; }
    if (x == 0.0) { memcpy(p, "0.0", 3); return (int)(p - out) + 3; }
    char sci[40];
    auto r = std::to_chars(sci, sci + sizeof(sci), x, std::chars_format::scientific);      // d[.ddd]e[+-]xx, shortest
    char* e = sci;
    while (e < r.ptr && *e != 'e') ++e;
    char digits[24];
    int nd = 0;
    for (char* q = sci; q < e; ++q) if (*q != '.') digits[nd++] = *q;
    while (nd > 1 && digits[nd - 1] == '0') --nd;
    int e10 = 0;
    {
        const char* q = e + 1;
        bool neg = false;
        if (*q == '-') { neg = true; ++q; } else if (*q == '+') ++q;
        for (; q < r.ptr; ++q) e10 = e10 * 10 + (*q - '0');
        if (neg) e10 = -e10;
    }
    if (e10 > -5 && e10 < 6) {
        if (e10 >= 0) {
            if (nd <= e10 + 1) {
                memcpy(p, digits, (size_t)nd); p += nd;
                for (int k = 0; k < e10 + 1 - nd; ++k) *p++ = '0';
                *p++ = '.'; *p++ = '0';
            } else {
                memcpy(p, digits, (size_t)(e10 + 1)); p += e10 + 1;
                *p++ = '.';
                memcpy(p, digits + e10 + 1, (size_t)(nd - e10 - 1)); p += nd - e10 - 1;
            }
        } else {
            *p++ = '0'; *p++ = '.';
            for (int k = 0; k < -e10 - 1; ++k) *p++ = '0';
            memcpy(p, digits, (size_t)nd); p += nd;
        }
    } else {
        *p++ = digits[0]; *p++ = '.';
        if (nd > 1) { memcpy(p, digits + 1, (size_t)(nd - 1)); p += nd - 1; } else *p++ = '0';
        *p++ = 'e';
        auto r2 = std::to_chars(p, p + 8, e10);
        p = r2.ptr;
    }
    return (int)(p - out);
}

extern "C" int abc_format_float64(double x, char* buf, size_t cap) {
    if (!buf || cap < 32) { abc_set_error("abc_format_float64: buffer of >= 32 bytes needed"); return ABC_ERR_ARG; }
    const int n = jl_format(x, buf);
    buf[n] = '\0';
    return n;
}

static void format_rows(const double* a, int64_t r0, int64_t r1, int64_t cols, int64_t pitch, std::string& out) {
    out.clear();
    out.reserve((size_t)((r1 - r0) * cols * 12));
    char tmp[32];
    for (int64_t r = r0; r < r1; ++r) {
        const double* row = a + r * pitch;
        for (int64_t c = 0; c < cols; ++c) {
            const int n = jl_format(row[c], tmp);
            out.append(tmp, (size_t)n);
            out.push_back(c + 1 < cols ? '\t' : '\n');
        }
    }
}

// rows x cols, row r at a + r * pitch; formatted in row blocks on all host threads, written in order
static int write_matrix(FILE* f, const double* a, int64_t rows, int64_t cols, int64_t pitch) {
    if (rows <= 0 || cols <= 0) return ABC_OK;
    const int64_t cells = rows * cols;
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)(hw ? hw : 4u);
    if (cells < (1 << 16)) nt = 1;
    if (nt > 64) nt = 64;
    const int64_t block = std::max<int64_t>(1, std::min<int64_t>((rows + nt - 1) / nt, std::max<int64_t>(1, (1 << 22) / cols)));
    std::vector<std::string> bufs((size_t)nt);
    for (int64_t r0 = 0; r0 < rows; r0 += block * nt) {
        std::vector<std::thread> th;
        int used = 0;
        for (int t = 0; t < nt; ++t) {
            const int64_t b0 = r0 + (int64_t)t * block, b1 = std::min(rows, b0 + block);
            if (b0 >= rows) break;
            ++used;
            if (nt == 1) format_rows(a, b0, b1, cols, pitch, bufs[0]);
            else th.emplace_back(format_rows, a, b0, b1, cols, pitch, std::ref(bufs[(size_t)t]));
        }
        for (auto& x : th) x.join();
        for (int t = 0; t < used; ++t)
            if (fwrite(bufs[(size_t)t].data(), 1, bufs[(size_t)t].size(), f) != bufs[(size_t)t].size()) {
                abc_set_error("write failed");
                return ABC_ERR_STATE;
            }
    }
    return ABC_OK;
}

extern "C" int abc_writedlm(const char* path, const double* a, int64_t rows, int64_t cols, int append) {
    if (!path || (!a && rows * cols > 0) || rows < 0 || cols < 0) { abc_set_error("abc_writedlm: bad arguments"); return ABC_ERR_ARG; }
    FILE* f = fopen(path, append ? "ab" : "wb");
    if (!f) { abc_set_error("cannot open %s", path); return ABC_ERR_STATE; }
    const int rc = write_matrix(f, a, rows, cols, cols);
    if (fclose(f) != 0 && rc == ABC_OK) { abc_set_error("close failed: %s", path); return ABC_ERR_STATE; }
    return rc;
}

static const char* model_name_of(int m) {
    static const char* names[5] = {"const", "const_const", "kon", "alpha", "gamma"};
    return (m >= 1 && m <= 5) ? names[m - 1] : nullptr;
}

static int mkdir_p(const std::string& dir) {
    std::string cur;
    for (size_t i = 0; i <= dir.size(); ++i) {
        if (i == dir.size() || dir[i] == '/') {
            if (!cur.empty() && mkdir(cur.c_str(), 0777) != 0) {
                struct stat st;
                if (stat(cur.c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) { abc_set_error("cannot create directory %s", cur.c_str()); return ABC_ERR_STATE; }
            }
        }
        if (i < dir.size()) cur.push_back(dir[i]);
    }
    return ABC_OK;
}

// abc_simulation.jl:47-61, 89-95 for n trials at once, appended to <dir>/<model>/...: progress_ (trial numbers
// first_trial .. first_trial+n-1), sets_ (1 x P), s_pulse_ / s_chase_ (2 rows x 5 per trial: means, Fano factors),
// s_ratios_, s_mean_corr_, s_corr_mean_ (1 x 11).  theta: n x P row-major, stats: n x 53 row-major.
extern "C" int abc_write_simulation(const char* dir, int m, int32_t submit, const double* theta, const double* stats, int64_t n,
                                    int64_t first_trial) {
    const char* name = model_name_of(m);
    if (!dir || !name || !theta || !stats || n < 0) { abc_set_error("abc_write_simulation: bad arguments"); return ABC_ERR_ARG; }
    const int P = (m <= 2) ? 5 : 9;
    const std::string base = std::string(dir) + "/" + name;
    int rc = mkdir_p(base);
    if (rc != ABC_OK) return rc;
    auto path = [&](const char* stem) { return base + "/" + stem + "_" + name + "_" + std::to_string(submit) + ".txt"; };
    {
        FILE* f = fopen(path("progress").c_str(), "ab");
        if (!f) { abc_set_error("cannot open %s", path("progress").c_str()); return ABC_ERR_STATE; }
        for (int64_t i = 0; i < n; ++i) fprintf(f, "%lld\n", (long long)(first_trial + i));
        fclose(f);
    }
    if ((rc = abc_writedlm(path("sets").c_str(), theta, n, P, 1)) != ABC_OK) return rc;
    // s_pulse / s_chase: per trial the row of means then the row of Fano factors = the 10 consecutive statistics as 2 x 5
    struct Part { const char* stem; int off, cols, rows_per_trial; };
    const Part parts[5] = {{"s_pulse", 0, 5, 2}, {"s_chase", 10, 5, 2}, {"s_ratios", 20, 11, 1}, {"s_mean_corr", 31, 11, 1}, {"s_corr_mean", 42, 11, 1}};
    for (const Part& pt : parts) {
        FILE* f = fopen(path(pt.stem).c_str(), "ab");
        if (!f) { abc_set_error("cannot open %s", path(pt.stem).c_str()); return ABC_ERR_STATE; }
        std::string buf;
        char tmp[32];
        for (int64_t i = 0; i < n; ++i) {
            const double* s = stats + i * ABC_NSTATS + pt.off;
            for (int r = 0; r < pt.rows_per_trial; ++r)
                for (int c = 0; c < pt.cols; ++c) {
                    const int k = jl_format(s[r * pt.cols + c], tmp);
                    buf.append(tmp, (size_t)k);
                    buf.push_back(c + 1 < pt.cols ? '\t' : '\n');
                }
            if (buf.size() > (1u << 20)) { fwrite(buf.data(), 1, buf.size(), f); buf.clear(); }
        }
        fwrite(buf.data(), 1, buf.size(), f);
        fclose(f);
    }
    return ABC_OK;
}

// accepted_particles.jl:19-30: one tab-separated line of 1-based particle indices per gene, "0" when none
extern "C" int abc_write_accepted(const char* path, const int64_t* offsets, const int64_t* idx, int32_t n_genes, int append) {
    if (!path || !offsets || n_genes < 0 || (!idx && offsets[n_genes] > 0)) { abc_set_error("abc_write_accepted: bad arguments"); return ABC_ERR_ARG; }
    FILE* f = fopen(path, append ? "ab" : "wb");
    if (!f) { abc_set_error("cannot open %s", path); return ABC_ERR_STATE; }
    std::string buf;
    char tmp[24];
    for (int g = 0; g < n_genes; ++g) {
        const int64_t b = offsets[g], e = offsets[g + 1];
        if (e <= b) buf += "0\n";
        for (int64_t k = b; k < e; ++k) {
            auto r = std::to_chars(tmp, tmp + sizeof(tmp), (long long)idx[k]);
            buf.append(tmp, (size_t)(r.ptr - tmp));
            buf.push_back(k + 1 < e ? '\t' : '\n');
        }
        if (buf.size() > (1u << 22)) { fwrite(buf.data(), 1, buf.size(), f); buf.clear(); }
    }
    fwrite(buf.data(), 1, buf.size(), f);
    if (fclose(f) != 0) { abc_set_error("close failed: %s", path); return ABC_ERR_STATE; }
    return ABC_OK;
}

// Column store of the error matrix: <dir>/x<g>.f64 = the raw little-endian Float64 column of gene g (1-based names like the
// DataFrame(:auto) columns of process_error_files.jl:5), <dir>/meta.txt = "n_genes n_rows".  err: gene-major, column g at
// err + g * pitch, n values.  append != 0 extends existing columns (batches of particles arrive one after the other).
extern "C" int abc_write_error_columns(const char* dir, const double* err, int64_t n, int64_t pitch, int32_t n_genes, int append) {
    if (!dir || (!err && n > 0) || n < 0 || pitch < n || n_genes <= 0) { abc_set_error("abc_write_error_columns: bad arguments"); return ABC_ERR_ARG; }
    int rc = mkdir_p(dir);
    if (rc != ABC_OK) return rc;
    long long have_g = 0, have_n = 0;
    const std::string meta = std::string(dir) + "/meta.txt";
    if (append) {
        FILE* f = fopen(meta.c_str(), "rb");
        if (f) {
            if (fscanf(f, "%lld %lld", &have_g, &have_n) != 2) { have_g = 0; have_n = 0; }
            fclose(f);
            if (have_g != 0 && have_g != n_genes) { abc_set_error("%s holds %lld genes, not %d", dir, have_g, n_genes); return ABC_ERR_ARG; }
        }
    }
    for (int g = 0; g < n_genes; ++g) {
        const std::string p = std::string(dir) + "/x" + std::to_string(g + 1) + ".f64";
        FILE* f = fopen(p.c_str(), append ? "ab" : "wb");
        if (!f) { abc_set_error("cannot open %s", p.c_str()); return ABC_ERR_STATE; }
        const size_t w = fwrite(err + (int64_t)g * pitch, sizeof(double), (size_t)n, f);
        if (fclose(f) != 0 || w != (size_t)n) { abc_set_error("write failed: %s", p.c_str()); return ABC_ERR_STATE; }
    }
    FILE* f = fopen(meta.c_str(), "wb");
    if (!f) { abc_set_error("cannot open %s", meta.c_str()); return ABC_ERR_STATE; }
    fprintf(f, "%d %lld\n", n_genes, (long long)(have_n + n));
    fclose(f);
    return ABC_OK;
}

// f["x<g>"] of the reference's JDFFile (accepted_particles.jl:14-18): g is 1-based; *n = rows stored
extern "C" int abc_read_error_column(const char* dir, int32_t g, double* out, int64_t cap, int64_t* n) {
    if (!dir || g < 1 || !n) { abc_set_error("abc_read_error_column: bad arguments"); return ABC_ERR_ARG; }
    const std::string p = std::string(dir) + "/x" + std::to_string(g) + ".f64";
    FILE* f = fopen(p.c_str(), "rb");
    if (!f) { abc_set_error("cannot open %s", p.c_str()); return ABC_ERR_STATE; }
    fseek(f, 0, SEEK_END);
    const long long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    *n = bytes / (long long)sizeof(double);
    int rc = ABC_OK;
    if (out) {
        if (cap < *n) { abc_set_error("abc_read_error_column: %lld rows, room for %lld", (long long)*n, (long long)cap); rc = ABC_ERR_ARG; }
        else if (fread(out, sizeof(double), (size_t)*n, f) != (size_t)*n) { abc_set_error("read failed: %s", p.c_str()); rc = ABC_ERR_STATE; }
    }
    fclose(f);
    return rc;
}
