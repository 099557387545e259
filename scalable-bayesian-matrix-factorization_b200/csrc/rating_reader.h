// rating_reader.h -- text rating files for the host program (main.cpp): `user SEP item SEP rating` triples with the line
// semantics of gibbs_sbpmf2.cpp ("[T]":35-73: sscanf("%u%c%u%c%lf"), a line counts iff all five conversions succeed) and libFM
// text lines `y user:v item:v` (sscanf("%lf %u:%lf %u:%lf")).
//
// [T] reads its input three times with getline + sscanf + std::map; at Netflix size (100M lines) that is minutes, and a plain
// one-pass getline + sscanf loop still takes about a minute -- far longer than the sampling job it feeds.  Here the file is
// read once, cut at line boundaries into one chunk per host thread, and every line first goes through a hand-written parser
// for the canonical spelling (decimal ids, std::from_chars for the numbers, which rounds exactly like strtod); any line that
// parser does not recognise is handed to the very sscanf call above, so the accepted language and the parsed values are
// those of sscanf by construction.  SBMF_SLOW_PARSER=1 sends every line through sscanf (the differential test compares both).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <charconv>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

namespace rating_reader {

struct Parsed {
    std::vector<uint32_t> user, item;
    std::vector<float> rating;
};

enum LineStatus { LINE_OK, LINE_BLANK, LINE_BAD };

inline bool is_blank(const char* b, const char* e)
{
    for (; b < e; ++b)
        if (*b != ' ' && *b != '\t' && *b != '\r' && *b != '\n') return false;
    return true;
}

// decimal id of at most 9 digits (anything longer, signed or otherwise unusual is left to sscanf)
inline bool fast_uint(const char*& p, const char* e, uint32_t& v)
{
    while (p < e && (*p == ' ' || *p == '\t')) ++p;
    const char* d = p;
    uint32_t x = 0;
    while (p < e && *p >= '0' && *p <= '9' && p - d < 10) x = x * 10 + (uint32_t)(*p++ - '0');
    if (p == d || p - d > 9) return false;
    v = x;
    return true;
}

// a number the way strtod reads it, restricted to spellings on which std::from_chars and strtod agree
inline bool fast_double(const char*& p, const char* e, double& v)
{
    while (p < e && (*p == ' ' || *p == '\t')) ++p;
    if (p == e || !((*p >= '0' && *p <= '9') || *p == '-' || *p == '.')) return false;   // '+', inf, nan, hex: sscanf
    const std::from_chars_result r = std::from_chars(p, e, v);
    if (r.ec != std::errc()) return false;
    if (r.ptr < e) {
        const char c = *r.ptr;   // "0x10", "1e5f", "3.5abc": strtod may read on (hex) or the line is odd anyway
        if ((c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '.' || c == '+' || c == '-') return false;
    }
    p = r.ptr;
    return true;
}

inline bool fast_triple(const char* p, const char* e, uint32_t& u, uint32_t& m, double& r)
{
    if (!fast_uint(p, e, u) || p == e) return false;
    ++p;                                          // %c: exactly one separator character, whatever it is
    if (!fast_uint(p, e, m) || p == e) return false;
    ++p;
    return fast_double(p, e, r);
}

inline bool fast_libfm(const char* p, const char* e, uint32_t& u, uint32_t& m, double& r)
{
    double v;
    if (!fast_double(p, e, r)) return false;
    if (!fast_uint(p, e, u) || p == e || *p != ':') return false;
    ++p;
    if (!fast_double(p, e, v)) return false;
    if (!fast_uint(p, e, m) || p == e || *p != ':') return false;
    ++p;
    return fast_double(p, e, v);
}

inline LineStatus slow_line(const char* b, const char* e, bool libfm, uint32_t& u, uint32_t& m, double& r)
{
    const std::string line(b, e);
    unsigned uu = 0, mm = 0;
    if (!libfm) {
        char w1, w2;
        if (sscanf(line.c_str(), "%u%c%u%c%lf", &uu, &w1, &mm, &w2, &r) >= 5) {
            u = uu;
            m = mm;
            return LINE_OK;
        }
    } else {
        double v1, v2;
        if (sscanf(line.c_str(), "%lf %u:%lf %u:%lf", &r, &uu, &v1, &mm, &v2) == 5) {
            u = uu;
            m = mm;
            return LINE_OK;
        }
    }
    return is_blank(b, e) ? LINE_BLANK : LINE_BAD;
}

inline LineStatus parse_line(const char* b, const char* e, bool libfm, bool slow_only, uint32_t& u, uint32_t& m, double& r)
{
    if (!slow_only && (libfm ? fast_libfm(b, e, u, m, r) : fast_triple(b, e, u, m, r))) return LINE_OK;
    return slow_line(b, e, libfm, u, m, r);
}

struct ChunkResult {
    Parsed p;
    uint64_t lines = 0;        // lines seen (all of the chunk, or up to and including the bad one)
    bool bad = false;
};

inline void parse_chunk(const char* b, const char* e, bool libfm, bool slow_only, ChunkResult& out)
{
    const size_t guess = (size_t)(e - b) / 12 + 16;
    out.p.user.reserve(guess);
    out.p.item.reserve(guess);
    out.p.rating.reserve(guess);
    while (b < e) {
        const char* nl = (const char*)memchr(b, '\n', (size_t)(e - b));
        const char* le = nl ? nl : e;
        ++out.lines;
        uint32_t u, m;
        double r;
        const LineStatus st = parse_line(b, le, libfm, slow_only, u, m, r);
        if (st == LINE_OK) {
            out.p.user.push_back(u);
            out.p.item.push_back(m);
            out.p.rating.push_back((float)r);
        } else if (st == LINE_BAD) {
            out.bad = true;
            return;
        }
        b = nl ? nl + 1 : e;
    }
}

// Reads the whole file.  libfm: set from the first non-blank line (a ':' means libFM text).  bad_line: 1-based number of the
// first line that is neither blank nor parseable (0 = none).  Returns false if the file cannot be opened.
inline bool read_text(const std::string& path, Parsed& out, bool& libfm, uint64_t& bad_line)
{
    bad_line = 0;
    libfm = false;
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f.is_open()) return false;
    std::string buf;
    f.seekg(0, std::ios::end);
    const std::streamoff sz = f.tellg();
    if (sz > 0) {
        buf.resize((size_t)sz);
        f.seekg(0, std::ios::beg);
        f.read(&buf[0], sz);
        buf.resize((size_t)f.gcount());
    } else {   // not seekable (a pipe): read to the end
        f.clear();
        f.seekg(0, std::ios::beg);
        f.clear();
        buf.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    }
    const char* b = buf.data();
    const char* e = b + buf.size();
    for (const char* p = b; p < e;) {   // format detection on the first non-blank line
        const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
        const char* le = nl ? nl : e;
        if (!is_blank(p, le)) {
            libfm = memchr(p, ':', (size_t)(le - p)) != nullptr;
            break;
        }
        p = nl ? nl + 1 : e;
    }
    const bool slow_only = getenv("SBMF_SLOW_PARSER") != nullptr;
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    nt = std::min(nt, 16u);
    if (buf.size() < (1u << 20)) nt = 1;
    std::vector<const char*> cut(nt + 1, e);
    cut[0] = b;
    for (unsigned t = 1; t < nt; ++t) {   // chunk boundaries just after a newline
        const char* p = b + buf.size() / nt * t;
        if (p < cut[t - 1]) p = cut[t - 1];
        const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
        cut[t] = nl ? nl + 1 : e;
    }
    std::vector<ChunkResult> res(nt);
    if (nt == 1) {
        parse_chunk(cut[0], cut[1], libfm, slow_only, res[0]);
    } else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(parse_chunk, cut[t], cut[t + 1], libfm, slow_only, std::ref(res[t]));
        for (auto& x : th) x.join();
    }
    uint64_t lines_before = 0;
    size_t total = 0;
    for (unsigned t = 0; t < nt; ++t) {
        if (res[t].bad) {
            bad_line = lines_before + res[t].lines;
            return true;
        }
        lines_before += res[t].lines;
        total += res[t].p.user.size();
    }
    out.user.resize(total);
    out.item.resize(total);
    out.rating.resize(total);
    size_t at = 0;
    for (unsigned t = 0; t < nt; ++t) {   // file order = chunk order
        const size_t n = res[t].p.user.size();
        if (n) {
            memcpy(&out.user[at], res[t].p.user.data(), n * 4);
            memcpy(&out.item[at], res[t].p.item.data(), n * 4);
            memcpy(&out.rating[at], res[t].p.rating.data(), n * 4);
        }
        at += n;
        ChunkResult().p.user.swap(res[t].p.user);   // release as we go
        ChunkResult().p.item.swap(res[t].p.item);
        ChunkResult().p.rating.swap(res[t].p.rating);
    }
    return true;
}

}  // namespace rating_reader
