// kernels_text.cu -- text block -> typed columns on the device: the per-block work of text.Reader.Block
// (go/text/text.go:181-200) = split at '\n' (go/text/parse.go:16-36), uncomment (:40-62), trim (:65-79), fields (:175-211)
// and the int64 / float32 column parsers (:81-172), which feed minh.Writer.Block in scripts/text_to_minh.go:166-214.
//
//   k_text_count / k_text_starts   newline positions -> the start of every line (warp ballots + a scan of tile counts)
//   k_text_lines                   per line: cut at the comment character, flag lines that hold a field at all
//   k_text_parse                   per kept line: walk its fields, parse the requested columns
//
// Integers: strconv.Atoi.  Floats: strconv.ParseFloat(s, 64) narrowed to float32 -- correctly rounded decimal -> float64 by
// Clinger's exact fast path (<= 15 digits, |exp10| <= 22: the common case by far) and otherwise the Eisel-Lemire
// algorithm with Go's own 128-bit power-of-ten table (strconv/eisel_lemire.go; tools/gen_pow10_table.py regenerates it and
// checks the transliteration against a correctly rounded conversion).  The few inputs Eisel-Lemire cannot decide
// (exact half-way cases, subnormal results, > 19 significant digits that straddle a rounding boundary, hexadecimal
// floats) are LISTED for the caller, who converts them with the host language's own parser: no approximate result is
// ever written.
#include "engine.cuh"
#include "launch.cuh"

namespace mnw {

namespace {

__device__ const unsigned long long MNW_POW10_128[696][2] = {
#include "pow10_table.inc"
};
__device__ const double MNW_POW10_D[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                           1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// strconv/eisel_lemire.go eiselLemire64.  ok = false: undecided.
__device__ bool eisel_lemire64(unsigned long long man, int exp10, double &out) {
    if (man == 0) { out = 0.0; return true; }
    if (exp10 < -348 || exp10 > 347) return false;
    const int clz = __clzll((long long)man);
    man <<= clz;
    long long ret_exp2 = ((217706LL * exp10) >> 16) + 64 + 1023 - clz;
    const unsigned long long p_lo = MNW_POW10_128[exp10 + 348][0], p_hi = MNW_POW10_128[exp10 + 348][1];
    unsigned long long x_hi = __umul64hi(man, p_hi), x_lo = man * p_hi;
    if ((x_hi & 0x1FFULL) == 0x1FFULL && x_lo + man < man) {
        const unsigned long long y_hi = __umul64hi(man, p_lo), y_lo = man * p_lo;
        unsigned long long m_hi = x_hi, m_lo = x_lo + y_hi;
        if (m_lo < x_lo) m_hi++;
        if ((m_hi & 0x1FFULL) == 0x1FFULL && m_lo + 1ULL == 0ULL && y_lo + man < man) return false;
        x_hi = m_hi; x_lo = m_lo;
    }
    const unsigned long long msb = x_hi >> 63;
    unsigned long long ret_man = x_hi >> (msb + 9);
    ret_exp2 -= (long long)(1ULL ^ msb);
    if (x_lo == 0 && (x_hi & 0x1FFULL) == 0 && (ret_man & 3ULL) == 1ULL) return false;
    ret_man += ret_man & 1ULL;
    ret_man >>= 1;
    if (ret_man >> 53) { ret_man >>= 1; ret_exp2 += 1; }
    if (ret_exp2 <= 0 || ret_exp2 >= 0x7FF) return false;   // subnormal or overflow: the caller's exact path
    out = __longlong_as_double((long long)(((unsigned long long)ret_exp2 << 52) | (ret_man & 0x000FFFFFFFFFFFFFULL)));
    return true;
}

__device__ __forceinline__ bool is_digit(unsigned char c) { return c >= '0' && c <= '9'; }
__device__ __forceinline__ unsigned char lower(unsigned char c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }

// strconv.ParseFloat(s, 64) narrowed to float32.  status: 0 = converted, 1 = the caller converts this field (see the header),
// 2 = syntax error (the reference panics).
__device__ int parse_float32(const unsigned char *s, int n, float &out) {
    int i = 0;
    bool neg = false;
    if (n == 0) return 2;
    if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; i = 1; }
    if (i >= n) return 2;
    if (!is_digit(s[i]) && s[i] != '.') {   // inf, infinity, nan (any case)
        const int r = n - i;
        const unsigned char a = lower(s[i]);
        if (a == 'i' && (r == 3 || r == 8)) {
            const char *w = "infinity";
            for (int k = 0; k < r; k++) if (lower(s[i + k]) != (unsigned char)w[k]) return 2;
            out = neg ? -INFINITY : INFINITY;
            return 0;
        }
        if (a == 'n' && r == 3 && lower(s[i + 1]) == 'a' && lower(s[i + 2]) == 'n' && i == 0) { out = NAN; return 0; }
        return 2;
    }
    if (s[i] == '0' && i + 1 < n && lower(s[i + 1]) == 'x') return 1;   // hexadecimal float: the host's parser
    unsigned long long man = 0;
    int nd = 0, ndm = 0, dp = 0;      // digits seen, digits in man, decimal point position
    bool saw_dot = false, saw_digits = false, trunc = false;
    for (; i < n; i++) {
        const unsigned char c = s[i];
        if (c == '.') {
            if (saw_dot) return 2;
            saw_dot = true;
            dp = nd;
            continue;
        }
        if (is_digit(c)) {
            saw_digits = true;
            if (c == '0' && nd == 0) { dp--; continue; }   // leading zeros
            nd++;
            if (ndm < 19) { man = man * 10ULL + (unsigned long long)(c - '0'); ndm++; }
            else if (c != '0') trunc = true;
            continue;
        }
        if (c == '_') return 1;   // (only legal with a base prefix; let the host's parser decide)
        break;
    }
    if (!saw_digits) return 2;
    if (!saw_dot) dp = nd;
    if (i < n && lower(s[i]) == 'e') {
        i++;
        if (i >= n) return 2;
        int esign = 1;
        if (s[i] == '+') i++;
        else if (s[i] == '-') { i++; esign = -1; }
        if (i >= n || !is_digit(s[i])) return 2;
        int e = 0;
        for (; i < n && is_digit(s[i]); i++) if (e < 10000) e = e * 10 + (s[i] - '0');
        dp += e * esign;
    }
    if (i != n) return 2;
    const int exp10 = dp - ndm;   // value = man * 10^exp10 (plus the truncated tail)
    double d;
    if (man == 0) {
        d = 0.0;
    } else if (!trunc && man <= (1ULL << 53) && exp10 >= -22 && exp10 <= 22) {   // Clinger: one exact operation
        d = __ull2double_rn(man);
        d = exp10 < 0 ? __ddiv_rn(d, MNW_POW10_D[-exp10]) : __dmul_rn(d, MNW_POW10_D[exp10]);
    } else if (!trunc && man <= (1ULL << 53) && exp10 > 22 && exp10 <= 22 + 15 &&
               __dmul_rn(__ull2double_rn(man), MNW_POW10_D[exp10 - 22]) <= 1e15) {
        d = __dmul_rn(__dmul_rn(__ull2double_rn(man), MNW_POW10_D[exp10 - 22]), 1e22);
    } else {
        if (!eisel_lemire64(man, exp10, d)) return 1;
        if (trunc) {   // > 19 digits: both neighbours of the truncated mantissa must agree (strconv/atof.go atof64)
            double d2;
            if (!eisel_lemire64(man + 1, exp10, d2) || d2 != d) return 1;
        }
    }
    out = __double2float_rn(neg ? -d : d);   // float32(x), go/text/parse.go:165
    return 0;
}

// strconv.Atoi.  false: syntax error or out of range (the reference panics).
__device__ bool parse_int64(const unsigned char *s, int n, long long &out) {
    int i = 0;
    bool neg = false;
    if (n == 0) return false;
    if (s[0] == '+' || s[0] == '-') { neg = s[0] == '-'; i = 1; }
    if (i >= n) return false;
    unsigned long long v = 0;
    for (; i < n; i++) {
        if (!is_digit(s[i])) return false;
        if (v > 922337203685477580ULL) return false;
        v = v * 10ULL + (unsigned long long)(s[i] - '0');
        if (v > 9223372036854775808ULL) return false;
    }
    if (!neg && v > 9223372036854775807ULL) return false;
    out = neg ? (long long)(0ULL - v) : (long long)v;
    return true;
}

}  // namespace

constexpr int TXT_TILE = 2048;   // bytes per warp

__global__ void __launch_bounds__(256) k_text_count(const unsigned char *buf, int64_t len, int64_t ntiles, int64_t *tile_nl) {
    const int lane = threadIdx.x & 31;
    const int64_t tile = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (tile >= ntiles) return;
    const int64_t b0 = tile * TXT_TILE;
    int cnt = 0;
    for (int64_t i = b0 + lane; i < b0 + TXT_TILE && i < len; i += 32) cnt += buf[i] == '\n';
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) tile_nl[tile] = cnt;
}

// line_start[0] = 0; the line after the r-th newline starts one byte behind it
__global__ void __launch_bounds__(256) k_text_starts(const unsigned char *buf, int64_t len, int64_t ntiles, const int64_t *tile_off,
                                                     int64_t *line_start) {
    const int lane = threadIdx.x & 31;
    const int64_t tile = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (tile >= ntiles) return;
    if (tile == 0 && lane == 0) line_start[0] = 0;
    const int64_t b0 = tile * TXT_TILE;
    int64_t r = tile_off[tile];
    for (int64_t ib = b0; ib < b0 + TXT_TILE && ib < len; ib += 32) {
        const int64_t i = ib + lane;
        const bool nl = i < len && buf[i] == '\n';
        const unsigned m = __ballot_sync(0xffffffffu, nl);
        if (nl) line_start[r + __popc(m & ((1u << lane) - 1u)) + 1] = i + 1;
        r += __popc(m);
    }
}

// uncomment + trim (go/text/parse.go:40-79): line_end = the comment character or the end of the line; keep = some byte is not
// the separator
__global__ void __launch_bounds__(256) k_text_lines(const unsigned char *buf, int64_t len, int64_t nlines, const int64_t *line_start,
                                                    unsigned char sep, unsigned char comm, int64_t *line_end, int64_t *keep) {
    const int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (l >= nlines) return;
    const int64_t s = line_start[l], e0 = l + 1 < nlines ? line_start[l + 1] - 1 : len;
    int64_t e = e0;
    bool any = false;
    for (int64_t i = s; i < e0; i++) {
        const unsigned char c = buf[i];
        if (c == comm) { e = i; break; }
        any = any || c != sep;
    }
    line_end[l] = e;
    keep[l] = any ? 1 : 0;
}

// len(bytes.Fields(lines[0])) of the first kept line (go/text/parse.go:92): fields separated by ASCII white space
__global__ void k_text_ncols(const unsigned char *buf, int64_t nlines, const int64_t *line_start, const int64_t *line_end,
                             const int64_t *keep, int *ncols) {
    for (int64_t l = 0; l < nlines; l++) {
        if (!keep[l]) continue;
        int n = 0;
        bool in = false;
        for (int64_t i = line_start[l]; i < line_end[l]; i++) {
            const unsigned char c = buf[i];
            const bool ws = c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r';
            if (!ws && !in) n++;
            in = !ws;
        }
        *ncols = n;
        return;
    }
    *ncols = 0;
}

struct TextCols {
    int n_i, n_f;
    int icol[64], fcol[64];   // column numbers, ascending
};

__global__ void __launch_bounds__(128) k_text_parse(const unsigned char *buf, int64_t nlines, const int64_t *line_start,
                                                    const int64_t *line_end, const int64_t *keep, const int64_t *row_of,
                                                    unsigned char sep, TextCols tc, const int *ncols, int64_t nrows, int64_t *iout,
                                                    float *fout, int64_t *fb_list, int fb_cap, int *fb_count, int *err,
                                                    long long *err_line) {
    const int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (l >= nlines || !keep[l]) return;
    const int64_t row = row_of[l];
    const unsigned char *p = buf + line_start[l];
    const int n = (int)(line_end[l] - line_start[l]);
    int field = 0, ii = 0, fi = 0, i = 0;
    const int maxcol = max(tc.n_i ? tc.icol[tc.n_i - 1] : -1, tc.n_f ? tc.fcol[tc.n_f - 1] : -1);
    if (maxcol >= *ncols) {   // "Data has %d columns, but column %d was requested."
        atomicExch(err, 5);
        return;
    }
    while (i < n) {
        while (i < n && p[i] == sep) i++;
        if (i >= n) break;
        const int f0 = i;
        while (i < n && p[i] != sep) i++;
        const int flen = i - f0;
        if (ii < tc.n_i && tc.icol[ii] == field) {
            long long v = 0;
            if (!parse_int64(p + f0, flen, v)) { if (atomicCAS(err, 0, 6) == 0) *err_line = l; }
            iout[(int64_t)ii * nrows + row] = v;
            ii++;
        }
        if (fi < tc.n_f && tc.fcol[fi] == field) {
            float v = 0.0f;
            const int st = parse_float32(p + f0, flen, v);
            if (st == 2) { if (atomicCAS(err, 0, 6) == 0) *err_line = l; }
            if (st == 1) {
                const int k = atomicAdd(fb_count, 1);
                if (k < fb_cap) { fb_list[3 * k] = row; fb_list[3 * k + 1] = fi; fb_list[3 * k + 2] = (line_start[l] + f0) | ((int64_t)flen << 40); }
            }
            fout[(int64_t)fi * nrows + row] = v;
            fi++;
        }
        field++;
    }
    if (field != *ncols) { if (atomicCAS(err, 0, 7) == 0) *err_line = l; }   // "Data on line %d has %d columns, not %d."
}

// ---- launchers ----
size_t text_tiles(int64_t len) { return (size_t)((len + TXT_TILE - 1) / TXT_TILE); }

void launch_text_count(Launcher &L, const unsigned char *buf, int64_t len, int64_t *tile_nl) {
    const int64_t ntiles = (int64_t)text_tiles(len);
    if (ntiles == 0) return;
    k_text_count<<<(unsigned)((ntiles * 32 + 255) / 256), 256, 0, L.stream>>>(buf, len, ntiles, tile_nl);
    L.count++;
}
void launch_text_starts(Launcher &L, const unsigned char *buf, int64_t len, const int64_t *tile_off, int64_t *line_start) {
    const int64_t ntiles = (int64_t)text_tiles(len);
    if (ntiles == 0) { cudaMemsetAsync(line_start, 0, 8, L.stream); return; }
    k_text_starts<<<(unsigned)((ntiles * 32 + 255) / 256), 256, 0, L.stream>>>(buf, len, ntiles, tile_off, line_start);
    L.count++;
}
void launch_text_lines(Launcher &L, const unsigned char *buf, int64_t len, int64_t nlines, const int64_t *line_start, unsigned char sep,
                       unsigned char comm, int64_t *line_end, int64_t *keep, int *ncols) {
    k_text_lines<<<(unsigned)((nlines + 255) / 256), 256, 0, L.stream>>>(buf, len, nlines, line_start, sep, comm, line_end, keep);
    k_text_ncols<<<1, 1, 0, L.stream>>>(buf, nlines, line_start, line_end, keep, ncols);
    L.count += 2;
}
void launch_text_parse(Launcher &L, const unsigned char *buf, int64_t nlines, const int64_t *line_start, const int64_t *line_end,
                       const int64_t *keep, const int64_t *row_of, unsigned char sep, int n_i, const int *icol, int n_f, const int *fcol,
                       const int *ncols, int64_t nrows, int64_t *iout, float *fout, int64_t *fb_list, int fb_cap, int *fb_count,
                       int *err, long long *err_line) {
    TextCols tc = {};
    tc.n_i = n_i; tc.n_f = n_f;
    for (int k = 0; k < n_i; k++) tc.icol[k] = icol[k];
    for (int k = 0; k < n_f; k++) tc.fcol[k] = fcol[k];
    L.begin("k_text_parse");
    k_text_parse<<<(unsigned)((nlines + 127) / 128), 128, 0, L.stream>>>(buf, nlines, line_start, line_end, keep, row_of, sep, tc, ncols, nrows,
                                                                        iout, fout, fb_list, fb_cap, fb_count, err, err_line);
    L.end();
    L.count++;
}

}  // namespace mnw
