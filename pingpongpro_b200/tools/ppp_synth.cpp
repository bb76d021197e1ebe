// ppp_synth.cpp -- writes the synthetic inputs of SURVEY.md App. E as SAM text or BGZF BAM
// (zlib level 1), plus the matching non-overlapping transposon annotation.
//
//   ppp_synth --genome dm6|mouse|human|tiny [--contigs n1:len1,n2:len2,...] --reads N --seed S
//             [--nh1] [--sorted] [--threads T] --out FILE.sam|FILE.bam
//             [--annot FILE.bed|.tsv|.csv|.gff|.gtf]
#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "synth_core.h"

using namespace ppp_synth;

static void put32(std::vector<uint8_t>& b, uint32_t v) { for (int i = 0; i < 4; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
static void put16(std::vector<uint8_t>& b, uint32_t v) { b.push_back((uint8_t)v); b.push_back((uint8_t)(v >> 8)); }

// One BGZF block (<= 0xff00 input bytes).
static void bgzfBlock(const uint8_t* in, size_t n, std::vector<uint8_t>& out, int level) {
    uint8_t cbuf[0x10000 + 1024];
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = (Bytef*)in;
    zs.avail_in = (uInt)n;
    zs.next_out = cbuf;
    zs.avail_out = sizeof cbuf;
    deflate(&zs, Z_FINISH);
    size_t clen = sizeof cbuf - zs.avail_out;
    deflateEnd(&zs);
    static const uint8_t hdr[12] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0};
    out.insert(out.end(), hdr, hdr + 12);
    out.push_back('B'); out.push_back('C'); put16(out, 2);
    put16(out, (uint32_t)(clen + 25));  // BSIZE = total block size - 1
    out.insert(out.end(), cbuf, cbuf + clen);
    put32(out, (uint32_t)crc32(crc32(0, nullptr, 0), in, (uInt)n));
    put32(out, (uint32_t)n);
}
static void bgzfAll(const std::vector<uint8_t>& raw, std::vector<uint8_t>& out, int level) {
    for (size_t o = 0; o < raw.size(); o += 0xff00) bgzfBlock(raw.data() + o, std::min<size_t>(0xff00, raw.size() - o), out, level);
}

static int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

static void emitSam(const Model& m, uint64_t i, std::string& out) {
    Read r;
    m.read(i, r);
    char seq[64];
    m.seq(i, r, seq);
    char tmp[256];
    int n;
    if (r.unmapped) n = snprintf(tmp, sizeof tmp, "r%llu\t4\t*\t0\t0\t*\t*\t0\t0\t", (unsigned long long)i);
    else {
        char cig[48];
        int k = 0;
        if (r.clipLead) k += snprintf(cig + k, sizeof cig - k, "%uS", r.clipLead);
        k += snprintf(cig + k, sizeof cig - k, "%uM", r.alnLen());
        if (r.clipTrail) k += snprintf(cig + k, sizeof cig - k, "%uS", r.clipTrail);
        n = snprintf(tmp, sizeof tmp, "r%llu\t%d\t%s\t%d\t255\t%s\t*\t0\t0\t", (unsigned long long)i, r.minus ? 16 : 0, m.names[r.contig].c_str(), r.beginPos + 1, cig);
    }
    out.append(tmp, n);
    out.append(seq, r.len);
    out.push_back('\t');
    out.append(r.len, 'I');
    n = snprintf(tmp, sizeof tmp, "\tNH:i:%u\n", r.nh);
    out.append(tmp, n);
}

static void emitBam(const Model& m, uint64_t i, std::vector<uint8_t>& out) {
    Read r;
    m.read(i, r);
    char seq[64];
    m.seq(i, r, seq);
    char name[32];
    int ln = snprintf(name, sizeof name, "r%llu", (unsigned long long)i) + 1;
    uint32_t cig[3];
    int nc = 0;
    if (!r.unmapped) {
        if (r.clipLead) cig[nc++] = (uint32_t)r.clipLead << 4 | 4;
        cig[nc++] = r.alnLen() << 4 | 0;
        if (r.clipTrail) cig[nc++] = (uint32_t)r.clipTrail << 4 | 4;
    }
    uint32_t blockSize = 32 + ln + 4 * nc + (r.len + 1) / 2 + r.len + 4;
    put32(out, blockSize);
    put32(out, (uint32_t)r.contig);
    put32(out, (uint32_t)r.beginPos);
    out.push_back((uint8_t)ln);
    out.push_back(r.unmapped ? 0 : 255);
    put16(out, r.unmapped ? 4680 : (uint32_t)reg2bin(r.beginPos, r.beginPos + (int64_t)r.alnLen()));
    put16(out, (uint32_t)nc);
    put16(out, r.unmapped ? 4 : (r.minus ? 16 : 0));
    put32(out, r.len);
    put32(out, 0xffffffffu);
    put32(out, 0xffffffffu);
    put32(out, 0);
    out.insert(out.end(), (uint8_t*)name, (uint8_t*)name + ln);
    for (int k = 0; k < nc; ++k) put32(out, cig[k]);
    for (uint32_t j = 0; j < r.len; j += 2) {
        auto code = [](char c) -> uint8_t { return c == 'A' ? 1 : c == 'C' ? 2 : c == 'G' ? 4 : c == 'T' ? 8 : 15; };
        uint8_t b = (uint8_t)(code(seq[j]) << 4);
        if (j + 1 < r.len) b |= code(seq[j + 1]);
        out.push_back(b);
    }
    out.insert(out.end(), r.len, (uint8_t)40);
    out.push_back('N'); out.push_back('H'); out.push_back('C'); out.push_back(r.nh);
}

int main(int argc, char** argv) {
    std::string genome = "tiny", outPath, annotPath, contigSpec;
    uint64_t nReads = 10000, seed = 1;
    bool nh1 = false, sorted = false;
    unsigned threads = std::max(1u, std::thread::hardware_concurrency());
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--genome") genome = val();
        else if (a == "--contigs") contigSpec = val();
        else if (a == "--reads") nReads = strtoull(val(), nullptr, 10);
        else if (a == "--seed") seed = strtoull(val(), nullptr, 10);
        else if (a == "--nh1") nh1 = true;
        else if (a == "--sorted") sorted = true;
        else if (a == "--threads") threads = (unsigned)atoi(val());
        else if (a == "--out") outPath = val();
        else if (a == "--annot") annotPath = val();
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    std::vector<std::string> names;
    std::vector<uint64_t> lens;
    if (!contigSpec.empty()) {
        size_t p = 0;
        while (p < contigSpec.size()) {
            size_t c = contigSpec.find(',', p);
            std::string item = contigSpec.substr(p, c == std::string::npos ? std::string::npos : c - p);
            size_t k = item.find(':');
            names.push_back(item.substr(0, k));
            lens.push_back(strtoull(item.c_str() + k + 1, nullptr, 10));
            if (c == std::string::npos) break;
            p = c + 1;
        }
    } else if (!preset(genome, names, lens)) { fprintf(stderr, "unknown genome preset %s\n", genome.c_str()); return 2; }
    Model m;
    m.init(names, lens, seed, nReads);
    m.nh1 = nh1;

    if (!annotPath.empty()) {
        FILE* f = fopen(annotPath.c_str(), "w");
        if (!f) { perror("annot"); return 1; }
        auto ends = [&](const char* e) { size_t n = strlen(e); return annotPath.size() >= n && annotPath.compare(annotPath.size() - n, n, e) == 0; };
        for (size_t k = 0; k < m.intervals.size(); ++k) {
            const Interval& iv = m.intervals[k];
            char strand = (rnd(seed, k, 103) & 1) ? '+' : '-';
            unsigned s = iv.start, e = iv.start + iv.len;  // 0-based, half-open
            const char* cn = names[iv.contig].c_str();
            if (ends(".bed")) fprintf(f, "%s\t%u\t%u\tTE%zu\t0\t%c\n", cn, s, e, k, strand);
            else if (ends(".csv")) fprintf(f, "TE%zu,%c,%s,%u,%u\n", k, strand, cn, s + 1, e);
            else if (ends(".gff") || ends(".gtf")) fprintf(f, "%s\t%s\ttransposon\t%u\t%u\t.\t%c\t.\tTE%zu\n", cn, cn, s + 1, e, strand, k);  // the reference reads the contig from column 2 (pingpongpro.cpp:1029)
            else fprintf(f, "TE%zu\t%c\t%s\t%u\t%u\n", k, strand, cn, s + 1, e);
        }
        fclose(f);
    }
    if (outPath.empty()) return 0;

    bool bam = outPath.size() >= 4 && outPath.compare(outPath.size() - 4, 4, ".bam") == 0;
    FILE* f = fopen(outPath.c_str(), "wb");
    if (!f) { perror("out"); return 1; }

    // header
    if (bam) {
        std::string text = "@HD\tVN:1.4\tSO:" + std::string(sorted ? "coordinate" : "unsorted") + "\n";
        for (size_t c = 0; c < names.size(); ++c) text += "@SQ\tSN:" + names[c] + "\tLN:" + std::to_string(lens[c]) + "\n";
        std::vector<uint8_t> raw, z;
        raw.insert(raw.end(), {'B', 'A', 'M', 1});
        put32(raw, (uint32_t)text.size());
        raw.insert(raw.end(), text.begin(), text.end());
        put32(raw, (uint32_t)names.size());
        for (size_t c = 0; c < names.size(); ++c) {
            put32(raw, (uint32_t)names[c].size() + 1);
            raw.insert(raw.end(), names[c].begin(), names[c].end());
            raw.push_back(0);
            put32(raw, (uint32_t)lens[c]);
        }
        bgzfAll(raw, z, 1);
        fwrite(z.data(), 1, z.size(), f);
    } else {
        fprintf(f, "@HD\tVN:1.4\tSO:%s\n", sorted ? "coordinate" : "unsorted");
        for (size_t c = 0; c < names.size(); ++c) fprintf(f, "@SQ\tSN:%s\tLN:%llu\n", names[c].c_str(), (unsigned long long)lens[c]);
    }

    // emission order
    std::vector<uint64_t> order;
    if (sorted) {
        struct K { uint64_t key, idx; };
        std::vector<K> ks(nReads);
        for (uint64_t i = 0; i < nReads; ++i) {
            Read r;
            m.read(i, r);
            ks[i].key = r.unmapped ? ~0ull : ((uint64_t)r.contig << 32 | (uint32_t)r.beginPos);
            ks[i].idx = i;
        }
        std::stable_sort(ks.begin(), ks.end(), [](const K& a, const K& b) { return a.key < b.key; });
        order.resize(nReads);
        for (uint64_t i = 0; i < nReads; ++i) order[i] = ks[i].idx;
    }

    const uint64_t chunk = 1u << 18;
    uint64_t nChunks = (nReads + chunk - 1) / chunk;
    for (uint64_t c0 = 0; c0 < nChunks; c0 += threads) {
        uint64_t nc = std::min<uint64_t>(threads, nChunks - c0);
        std::vector<std::vector<uint8_t> > outs(nc);
        std::vector<std::thread> pool;
        for (uint64_t t = 0; t < nc; ++t)
            pool.emplace_back([&, t]() {
                uint64_t lo = (c0 + t) * chunk, hi = std::min(nReads, lo + chunk);
                if (bam) {
                    std::vector<uint8_t> raw;
                    raw.reserve(0x10000);
                    for (uint64_t i = lo; i < hi; ++i) {
                        emitBam(m, sorted ? order[i] : i, raw);
                        if (raw.size() > 0xf000) { bgzfAll(raw, outs[t], 1); raw.clear(); }
                    }
                    if (!raw.empty()) bgzfAll(raw, outs[t], 1);
                } else {
                    std::string s;
                    for (uint64_t i = lo; i < hi; ++i) emitSam(m, sorted ? order[i] : i, s);
                    outs[t].assign(s.begin(), s.end());
                }
            });
        for (auto& th : pool) th.join();
        for (uint64_t t = 0; t < nc; ++t) fwrite(outs[t].data(), 1, outs[t].size(), f);
    }
    if (bam) {  // BGZF EOF marker
        static const uint8_t eof[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        fwrite(eof, 1, 28, f);
    }
    fclose(f);
    return 0;
}
