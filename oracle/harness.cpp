// harness.cpp -- include-harness around the UNMODIFIED reference (TEST INFRASTRUCTURE).
//
// It textually includes /root/reference/pingpongpro.cpp (path given by
// -DPPP_REFERENCE_SOURCE=...), renames its main(), and calls the reference's own
// functions in main()'s order (pingpongpro.cpp:1477-1606) so every intermediate of the
// hot path can be dumped (mode "dump") or timed stage by stage (mode "time").
// Nothing here restates the algorithm; it only drives and observes the reference.
//
//   ppp_harness dump OUT.bin [-m weighted|discard|unique] [-l MIN] [-L MAX] [-t ANNOT]... -i FILE...
//   ppp_harness time --genome G --reads N --seed S [--nh1] [--gpos LO:HI] [-m MODE]
//                    [--steps K] [--warmup W] [--from-packed]
//   ppp_harness digest --genome G --reads N --seed S [--nh1] [-m MODE] [-t ANNOT]... [-s S1,S2,...]
//                    [--pq OUT.f32]
//       full-size runs (BASELINE.json configs 2-5): the synthetic records are generated on the fly and fed to
//       the reference's own functions; SHA-256 digests of every intermediate go to stdout as one JSON object
//       (byte layouts in tests/fullsize_digest.py, which hashes the CUDA path's results the same way).
//
// Dump format (little endian): repeated sections  { char tag[4]; uint64 count; payload }.
#define main reference_main
#include PPP_REFERENCE_SOURCE
#undef main

#include <chrono>
#include <cstdio>
#include <set>
#include <thread>

#include "sha256.h"
#include "synth_core.h"

namespace {

struct Out {
    FILE* f;
    void sec(const char* tag, uint64_t n) { fwrite(tag, 1, 4, f); fwrite(&n, 8, 1, f); }
    void u32(uint32_t v) { fwrite(&v, 4, 1, f); }
    void f32(float v) { fwrite(&v, 4, 1, f); }
    void f64(double v) { fwrite(&v, 8, 1, f); }
};

TCountMultiHits parseMode(const std::string& m) {
    if (m == "unique") return multiHitsUnique;
    if (m == "discard") return multiHitsDiscard;
    return multiHitsWeighted;
}

const unsigned W = MAX_ARBITRARY_OVERLAP - MIN_ARBITRARY_OVERLAP + 1;

void dumpGrouped(Out& o, const char* tag, TGroupedStackCountsByOverlap& g) {
    uint64_t bins = g.begin()->size();
    o.sec(tag, bins);
    for (unsigned ov = 0; ov < g.size(); ++ov)
        for (unsigned b = 0; b < bins; ++b)
            for (unsigned i = 0; i < 2; ++i)
                for (unsigned j = 0; j < 2; ++j) o.f32(g[ov][b][i][j]);
}

void dumpTransposons(Out& o, const char* tag, TTransposonsPerGenome& t) {
    uint64_t n = 0;
    for (TTransposonsPerGenome::iterator c = t.begin(); c != t.end(); ++c) n += c->second.size();
    o.sec(tag, n);
    for (TTransposonsPerGenome::iterator c = t.begin(); c != t.end(); ++c)
        for (TTransposonsPerContig::iterator x = c->second.begin(); x != c->second.end(); ++x) {
            o.u32(c->first); o.u32(x->strand); o.u32(x->start); o.u32(x->end);
            o.f32(x->pValue); o.f32(x->qValue); o.f32(x->readsOnPlusStrand); o.f32(x->readsOnMinusStrand);
            for (unsigned k = 0; k < W; ++k) o.f32(k < x->histogram.size() ? x->histogram[k] : 0.0f);
        }
}

int runDump(int argc, char** argv) {
    if (argc < 3) return 2;
    std::string outPath = argv[2], mode = "weighted";
    unsigned minLen = 24, maxLen = 32, predictRange = 0;
    std::vector<std::string> inputs, annots;
    for (int i = 3; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-m") mode = argv[++i];
        else if (a == "-l") minLen = atoi(argv[++i]);
        else if (a == "-L") maxLen = atoi(argv[++i]);
        else if (a == "-t") annots.push_back(argv[++i]);
        else if (a == "-T") predictRange = atoi(argv[++i]);
        else if (a == "-i") inputs.push_back(argv[++i]);
        else inputs.push_back(a);
    }
    Out o{fopen(outPath.c_str(), "wb")};
    if (!o.f) { perror("out"); return 1; }
    o.sec("HDR ", 3);
    o.u32(MIN_ARBITRARY_OVERLAP); o.u32(MAX_ARBITRARY_OVERLAP); o.u32(PING_PONG_OVERLAP);

    TReadStacksPerGenome readStacks;
    TNameStore names;
    double total = 0;
    for (size_t k = 0; k < inputs.size(); ++k) {
        BamStream bam(inputs[k].c_str());
        if (!isGood(bam)) { fprintf(stderr, "Failed to open input file: %s\n", inputs[k].c_str()); return 1; }
        if (countReadsInBamFile(bam, readStacks, minLen, maxLen, parseMode(mode), total) != 0) return 1;
        if (length(names) == 0) names = nameStore(bam.bamIOContext);
        close(bam);
    }
    uint64_t nStacks = 0;
    for (unsigned s = 0; s < 2; ++s)
        for (TReadStacksPerStrand::iterator c = readStacks[s].begin(); c != readStacks[s].end(); ++c) nStacks += c->second.size();
    o.sec("STAK", nStacks);
    for (unsigned s = 0; s < 2; ++s)
        for (TReadStacksPerStrand::iterator c = readStacks[s].begin(); c != readStacks[s].end(); ++c)
            for (TReadStacksPerContig::iterator p = c->second.begin(); p != c->second.end(); ++p) {
                o.u32(s); o.u32(c->first); o.u32(p->first); o.f32(p->second.reads); o.u32(p->second.AAtPosition10 ? 1 : 0);
            }
    o.sec("TOTL", 1);
    o.f64(total);

    TTransposonsPerGenome transposons;
    for (size_t k = 0; k < annots.size(); ++k) {
        ifstream fs(annots[k].c_str());
        if (fs.fail()) { fprintf(stderr, "Failed to open transposon file \"%s\".\n", annots[k].c_str()); return 1; }
        TFileFormat ff = fileFormatTSV;
        if (_compareExtension(annots[k].c_str(), ".bed")) ff = fileFormatBED;
        else if (_compareExtension(annots[k].c_str(), ".csv")) ff = fileFormatCSV;
        else if (_compareExtension(annots[k].c_str(), ".gff")) ff = fileFormatGFF;
        else if (_compareExtension(annots[k].c_str(), ".gtf")) ff = fileFormatGTF;
        readTransposonsFromFile(fs, ff, transposons, names);
    }

    THeightScoreMap heightScoreMap;
    mapHeightsToScores(readStacks, heightScoreMap);
    o.sec("FREQ", heightScoreMap.size());
    for (THeightScoreMap::iterator h = heightScoreMap.begin(); h != heightScoreMap.end(); ++h) { o.u32(h->first); o.f32(h->second); }

    TGroupedStackCountsByOverlap grouped;
    TPingPongSignaturesByOverlap sigs;
    countStacksByGroup(readStacks, heightScoreMap, grouped, sigs);
    dumpGrouped(o, "GRP0", grouped);
    std::vector<unsigned> binPre;
    for (unsigned ov = 0; ov < sigs.size(); ++ov)
        for (TPingPongSignaturesPerGenome::iterator c = sigs[ov].begin(); c != sigs[ov].end(); ++c)
            for (TPingPongSignaturesPerContig::iterator s = c->second.begin(); s != c->second.end(); ++s) binPre.push_back(s->heightScoreBin);

    collapseBins(grouped, sigs);
    dumpGrouped(o, "GRP1", grouped);
    calculateFDRs(grouped, sigs);

    o.sec("SIGS", binPre.size());
    size_t q = 0;
    for (unsigned ov = 0; ov < sigs.size(); ++ov)
        for (TPingPongSignaturesPerGenome::iterator c = sigs[ov].begin(); c != sigs[ov].end(); ++c)
            for (TPingPongSignaturesPerContig::iterator s = c->second.begin(); s != c->second.end(); ++s) {
                o.u32(ov); o.u32(c->first); o.u32(s->position); o.u32(binPre[q++]); o.u32(s->heightScoreBin);
                o.u32(s->localHeightScoreBin); o.u32(s->baseBiasBin);
                o.f32(s->readsOnPlusStrand); o.f32(s->readsOnMinusStrand); o.f32(s->fdr);
            }

    if (!annots.empty()) {
        findSuppressedTransposons(sigs, transposons);
        dumpTransposons(o, "TRSP", transposons);
    }
    if (predictRange > 0) {
        TTransposonsPerGenome putative;
        predictSuppressedTransposons(sigs, putative, names, predictRange);
        dumpTransposons(o, "PRED", putative);
    }
    fclose(o.f);
    return 0;
}

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Times the reference stages on decoded records held in memory (BamStream "@mem").
int runTime(int argc, char** argv) {
    std::string genome = "dm6", mode = "weighted";
    uint64_t nReads = 1000000, seed = 2, gLo = 0, gHi = ~0ull;
    bool nh1 = false;
    int steps = 1, warmup = 0;
    unsigned minLen = 24, maxLen = 32;
    for (int i = 2; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--genome") genome = argv[++i];
        else if (a == "--reads") nReads = strtoull(argv[++i], 0, 10);
        else if (a == "--seed") seed = strtoull(argv[++i], 0, 10);
        else if (a == "--nh1") nh1 = true;
        else if (a == "--gpos") { std::string s = argv[++i]; size_t c = s.find(':'); gLo = strtoull(s.c_str(), 0, 10); gHi = strtoull(s.c_str() + c + 1, 0, 10); }
        else if (a == "-m") mode = argv[++i];
        else if (a == "--steps") steps = atoi(argv[++i]);
        else if (a == "--warmup") warmup = atoi(argv[++i]);
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    std::vector<std::string> names;
    std::vector<uint64_t> lens;
    if (!ppp_synth::preset(genome, names, lens)) return 2;
    ppp_synth::Model m;
    m.init(names, lens, seed, nReads);
    m.nh1 = nh1;
    seqan::shim::memNames() = names;
    std::vector<seqan::shim::MemRecord>& recs = seqan::shim::memRecords();
    // generate the sample in parallel, keep file (index) order
    unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    std::vector<std::vector<seqan::shim::MemRecord> > parts(nt);
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t)
        pool.emplace_back([&, t]() {
            uint64_t lo = nReads * t / nt, hi = nReads * (t + 1) / nt;
            char seq[64];
            for (uint64_t i = lo; i < hi; ++i) {
                ppp_synth::Read r;
                m.read(i, r);
                if (r.unmapped) continue;
                uint64_t g = m.cum[r.contig] + (uint64_t)r.beginPos;
                if (g < gLo || g >= gHi) continue;
                m.seq(i, r, seq);
                seqan::shim::MemRecord mr;
                mr.rID = r.contig; mr.beginPos = r.beginPos; mr.flag = r.minus ? 16 : 0;
                if (r.clipLead) mr.cigar.push_back(seqan::CigarElement{'S', r.clipLead});
                mr.cigar.push_back(seqan::CigarElement{'M', r.alnLen()});
                if (r.clipTrail) mr.cigar.push_back(seqan::CigarElement{'S', r.clipTrail});
                mr.seq.assign(seq, r.len);
                mr.hasNH = true; mr.nh = r.nh;
                parts[t].push_back(mr);
            }
        });
    for (auto& th : pool) th.join();
    size_t total = 0;
    for (auto& p : parts) total += p.size();
    recs.reserve(total);
    for (auto& p : parts) { recs.insert(recs.end(), p.begin(), p.end()); std::vector<seqan::shim::MemRecord>().swap(p); }

    uint64_t positions = std::min<uint64_t>(gHi, m.cum.back()) - gLo;
    for (int it = 0; it < warmup + steps; ++it) {
        TReadStacksPerGenome readStacks;
        double totalReads = 0;
        BamStream bam("@mem");
        double t0 = now();
        countReadsInBamFile(bam, readStacks, minLen, maxLen, parseMode(mode), totalReads);
        double t1 = now();
        THeightScoreMap hs;
        mapHeightsToScores(readStacks, hs);
        double t2 = now();
        TGroupedStackCountsByOverlap grouped;
        TPingPongSignaturesByOverlap sigs;
        countStacksByGroup(readStacks, hs, grouped, sigs);
        double t3 = now();
        collapseBins(grouped, sigs);
        double t4 = now();
        calculateFDRs(grouped, sigs);
        double t5 = now();
        size_t loci = 0, pairs = 0;
        for (unsigned ov = 0; ov < sigs.size(); ++ov)
            for (TPingPongSignaturesPerGenome::iterator c = sigs[ov].begin(); c != sigs[ov].end(); ++c) {
                pairs += c->second.size();
                if (ov == (unsigned)(PING_PONG_OVERLAP - MIN_ARBITRARY_OVERLAP)) loci += c->second.size();
            }
        printf("{\"step\": %d, \"timed\": %s, \"records\": %zu, \"positions\": %llu, \"t_count\": %.6f, \"t_heights\": %.6f, \"t_group\": %.6f, "
               "\"t_collapse\": %.6f, \"t_fdr\": %.6f, \"pairs\": %zu, \"loci\": %zu, \"total_weight\": %.3f}\n",
               it, it >= warmup ? "true" : "false", recs.size(), (unsigned long long)positions, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, pairs, loci, totalReads);
        fflush(stdout);
    }
    return 0;
}

// ---- digest mode ------------------------------------------------------------------------------------
// Records of the synthetic model, generated in parallel one batch at a time and served in index (= file) order.
struct SynthFeed {
    const ppp_synth::Model* m;
    uint64_t next = 0, end = 0;
    struct Lite {
        int32_t rID, beginPos;
        uint8_t minus, clipLead, clipTrail, len, nh;
        char seq[35];
    };
    std::vector<Lite> batch;
    size_t pos = 0;
    void refill() {
        const uint64_t kBatch = 1u << 22;
        uint64_t n = std::min<uint64_t>(kBatch, end - next);
        batch.resize(n);
        pos = 0;
        unsigned nt = std::max(1u, std::thread::hardware_concurrency());
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nt; ++t)
            pool.emplace_back([&, t]() {
                for (uint64_t k = n * t / nt; k < n * (t + 1) / nt; ++k) {
                    ppp_synth::Read r;
                    m->read(next + k, r);
                    Lite& l = batch[k];
                    l.rID = r.contig; l.beginPos = r.beginPos; l.minus = r.minus; l.clipLead = r.clipLead; l.clipTrail = r.clipTrail; l.len = r.len; l.nh = r.nh;
                    m->seq(next + k, r, l.seq);
                }
            });
        for (auto& th : pool) th.join();
        next += n;
    }
    static bool pull(seqan::shim::MemRecord& mr, void* user) {
        SynthFeed* f = static_cast<SynthFeed*>(user);
        if (f->pos >= f->batch.size()) {
            if (f->next >= f->end) return false;
            f->refill();
        }
        const Lite& l = f->batch[f->pos++];
        mr.rID = l.rID;
        mr.beginPos = l.beginPos;            // -1 when unmapped (flag 4): skipped at pingpongpro.cpp:415
        mr.flag = l.beginPos < 0 ? 4u : (l.minus ? 16u : 0u);
        mr.cigar.clear();
        if (l.clipLead) mr.cigar.push_back(seqan::CigarElement{'S', l.clipLead});
        mr.cigar.push_back(seqan::CigarElement{'M', (unsigned)l.len - l.clipLead - l.clipTrail});
        if (l.clipTrail) mr.cigar.push_back(seqan::CigarElement{'S', l.clipTrail});
        mr.seq.assign(l.seq, l.len);
        mr.hasNH = true;
        mr.nh = l.nh;
        return true;
    }
};

int runDigest(int argc, char** argv) {
    std::string genome = "dm6", mode = "weighted", pqPath;
    uint64_t nReads = 1000000, seed = 2;
    bool nh1 = false;
    unsigned minLen = 24, maxLen = 32;
    std::vector<std::string> annots;
    std::vector<unsigned> minHeights(1, 0u);
    for (int i = 2; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--genome") genome = argv[++i];
        else if (a == "--reads") nReads = strtoull(argv[++i], 0, 10);
        else if (a == "--seed") seed = strtoull(argv[++i], 0, 10);
        else if (a == "--nh1") nh1 = true;
        else if (a == "-m") mode = argv[++i];
        else if (a == "-t") annots.push_back(argv[++i]);
        else if (a == "--pq") pqPath = argv[++i];
        else if (a == "-s") {
            minHeights.clear();
            std::string v = argv[++i];
            for (size_t p = 0; p < v.size();) { size_t c = v.find(',', p); if (c == std::string::npos) c = v.size(); minHeights.push_back(atoi(v.substr(p, c - p).c_str())); p = c + 1; }
        } else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    std::vector<std::string> names;
    std::vector<uint64_t> lens;
    if (!ppp_synth::preset(genome, names, lens)) return 2;
    ppp_synth::Model m;
    m.init(names, lens, seed, nReads);
    m.nh1 = nh1;
    seqan::shim::memNames() = names;
    SynthFeed feed;
    feed.m = &m;
    feed.end = nReads;
    seqan::shim::genSource().fn = &SynthFeed::pull;
    seqan::shim::genSource().user = &feed;

    double t0 = now();
    TReadStacksPerGenome readStacks;
    TNameStore nameStoreCopy;
    double total = 0;
    {
        BamStream bam("@gen");
        if (countReadsInBamFile(bam, readStacks, minLen, maxLen, parseMode(mode), total) != 0) return 1;
        nameStoreCopy = nameStore(bam.bamIOContext);
    }
    double t1 = now();
    fprintf(stderr, "[digest] countReadsInBamFile %.1f s\n", t1 - t0);

    printf("{\"genome\": \"%s\", \"reads\": %llu, \"seed\": %llu, \"mode\": \"%s\", \"nh1\": %s, \"min_overlap\": %d, \"max_overlap\": %d",
           genome.c_str(), (unsigned long long)nReads, (unsigned long long)seed, mode.c_str(), nh1 ? "true" : "false", MIN_ARBITRARY_OVERLAP, MAX_ARBITRARY_OVERLAP);
    {   // stacks: {u32 strand, u32 contig, u32 key, f32 reads, u32 a10} in (strand, contig, key) order
        ppp_sha::Sha256 h;
        uint64_t n = 0;
        for (unsigned s = 0; s < 2; ++s)
            for (TReadStacksPerStrand::iterator c = readStacks[s].begin(); c != readStacks[s].end(); ++c)
                for (TReadStacksPerContig::iterator p = c->second.begin(); p != c->second.end(); ++p, ++n) {
                    h.put((uint32_t)s); h.put((uint32_t)c->first); h.put((uint32_t)p->first); h.put((float)p->second.reads); h.put((uint32_t)(p->second.AAtPosition10 ? 1 : 0));
                }
        printf(", \"n_stacks\": %llu, \"stacks\": \"%s\", \"total_weight\": %.17g", (unsigned long long)n, h.hex().c_str(), total);
    }

    TTransposonsPerGenome transposons;
    for (size_t k = 0; k < annots.size(); ++k) {
        ifstream fs(annots[k].c_str());
        if (fs.fail()) { fprintf(stderr, "Failed to open transposon file \"%s\".\n", annots[k].c_str()); return 1; }
        TFileFormat ff = fileFormatTSV;
        if (_compareExtension(annots[k].c_str(), ".bed")) ff = fileFormatBED;
        else if (_compareExtension(annots[k].c_str(), ".csv")) ff = fileFormatCSV;
        else if (_compareExtension(annots[k].c_str(), ".gff")) ff = fileFormatGFF;
        else if (_compareExtension(annots[k].c_str(), ".gtf")) ff = fileFormatGTF;
        readTransposonsFromFile(fs, ff, transposons, nameStoreCopy);
    }

    THeightScoreMap heightScoreMap;
    mapHeightsToScores(readStacks, heightScoreMap);
    double t2 = now();
    fprintf(stderr, "[digest] mapHeightsToScores %.1f s\n", t2 - t1);
    {   // {u32 height, f32 frequency} ascending; how many counters reached the 2^24 ceiling of a float counter (:508)
        ppp_sha::Sha256 h;
        uint64_t saturated = 0;
        float maxFreq = 0;
        for (THeightScoreMap::iterator x = heightScoreMap.begin(); x != heightScoreMap.end(); ++x) {
            h.put((uint32_t)x->first); h.put((float)x->second);
            if (x->second >= 16777216.0f) ++saturated;
            if (x->second > maxFreq) maxFreq = x->second;
        }
        printf(", \"n_heights\": %zu, \"freq\": \"%s\", \"freq_saturated\": %llu, \"max_freq\": %.9g, \"max_height\": %u", heightScoreMap.size(), h.hex().c_str(),
               (unsigned long long)saturated, maxFreq, heightScoreMap.empty() ? 0u : (unsigned)heightScoreMap.rbegin()->first);
    }

    TGroupedStackCountsByOverlap grouped;
    TPingPongSignaturesByOverlap sigs;
    countStacksByGroup(readStacks, heightScoreMap, grouped, sigs);
    double t3 = now();
    fprintf(stderr, "[digest] countStacksByGroup %.1f s\n", t3 - t2);
    for (unsigned s = 0; s < 2; ++s) TReadStacksPerStrand().swap(readStacks[s]);  // the stacks are not read again (memory for the signature lists)
    auto digestGrouped = [&](const char* key) {
        ppp_sha::Sha256 h;
        uint64_t bins = grouped.begin()->size(), saturated = 0;
        for (unsigned ov = 0; ov < grouped.size(); ++ov)
            for (unsigned b = 0; b < bins; ++b)
                for (unsigned i = 0; i < 2; ++i)
                    for (unsigned j = 0; j < 2; ++j) { h.put((float)grouped[ov][b][i][j]); if (grouped[ov][b][i][j] >= 16777216.0f) ++saturated; }
        printf(", \"%s_bins\": %llu, \"%s\": \"%s\", \"%s_cells_at_or_above_2p24\": %llu", key, (unsigned long long)bins, key, h.hex().c_str(), key, (unsigned long long)saturated);
    };
    digestGrouped("grouped_pre");
    // height-score bins before the collapse, per (overlap, contig) in list order
    std::vector<std::map<unsigned, std::vector<uint16_t> > > binPre(sigs.size());
    for (unsigned ov = 0; ov < sigs.size(); ++ov)
        for (TPingPongSignaturesPerGenome::iterator c = sigs[ov].begin(); c != sigs[ov].end(); ++c) {
            std::vector<uint16_t>& v = binPre[ov][c->first];
            v.reserve(c->second.size());
            for (TPingPongSignaturesPerContig::iterator s = c->second.begin(); s != c->second.end(); ++s) v.push_back((uint16_t)s->heightScoreBin);
        }
    collapseBins(grouped, sigs);
    digestGrouped("grouped_post");
    calculateFDRs(grouped, sigs);
    double t4 = now();
    fprintf(stderr, "[digest] collapseBins + calculateFDRs %.1f s\n", t4 - t3);
    {   // signatures in (contig, position, overlap) order, the layout of ppp_pair (include/ppp_b200.h):
        // {u32 contig, position, overlap, bin before collapse, collapsed bin, local bin, bias bin; f32 r+, r-, fdr}
        ppp_sha::Sha256 h;
        uint64_t n = 0;
        std::set<unsigned> contigs;
        for (unsigned ov = 0; ov < sigs.size(); ++ov)
            for (TPingPongSignaturesPerGenome::iterator c = sigs[ov].begin(); c != sigs[ov].end(); ++c) contigs.insert(c->first);
        for (std::set<unsigned>::iterator ci = contigs.begin(); ci != contigs.end(); ++ci) {
            const unsigned nOv = sigs.size();
            std::vector<TPingPongSignaturesPerContig::iterator> it(nOv), en(nOv);
            std::vector<const std::vector<uint16_t>*> pre(nOv, nullptr);
            std::vector<size_t> idx(nOv, 0);
            TPingPongSignaturesPerContig empty;
            for (unsigned ov = 0; ov < nOv; ++ov) {
                TPingPongSignaturesPerGenome::iterator f = sigs[ov].find(*ci);
                if (f == sigs[ov].end()) { it[ov] = en[ov] = empty.end(); continue; }
                it[ov] = f->second.begin();
                en[ov] = f->second.end();
                pre[ov] = &binPre[ov][*ci];
            }
            for (;;) {
                unsigned best = 0xffffffffu;
                for (unsigned ov = 0; ov < nOv; ++ov)
                    if (it[ov] != en[ov] && it[ov]->position < best) best = it[ov]->position;
                if (best == 0xffffffffu) {
                    bool any = false;
                    for (unsigned ov = 0; ov < nOv; ++ov) any = any || it[ov] != en[ov];
                    if (!any) break;
                }
                for (unsigned ov = 0; ov < nOv; ++ov)
                    while (it[ov] != en[ov] && it[ov]->position == best) {
                        h.put((uint32_t)*ci); h.put((uint32_t)it[ov]->position); h.put((uint32_t)(ov + MIN_ARBITRARY_OVERLAP)); h.put((uint32_t)(*pre[ov])[idx[ov]++]);
                        h.put((uint32_t)it[ov]->heightScoreBin); h.put((uint32_t)it[ov]->localHeightScoreBin); h.put((uint32_t)it[ov]->baseBiasBin);
                        h.put((float)it[ov]->readsOnPlusStrand); h.put((float)it[ov]->readsOnMinusStrand); h.put((float)it[ov]->fdr);
                        ++it[ov];
                        ++n;
                    }
            }
        }
        printf(", \"n_signatures\": %llu, \"signatures\": \"%s\"", (unsigned long long)n, h.hex().c_str());
    }
    {   // reported loci per -s (the filter of pingpongpro.cpp:903), the layout of ppp_locus: {u32 contig, position; f32 fdr, r+, r-}
        TPingPongSignaturesPerGenome& pp = sigs[PING_PONG_OVERLAP - MIN_ARBITRARY_OVERLAP];
        printf(", \"loci\": {");
        for (size_t k = 0; k < minHeights.size(); ++k) {
            const unsigned int minStackHeight = minHeights[k];
            ppp_sha::Sha256 h;
            uint64_t n = 0;
            for (TPingPongSignaturesPerGenome::iterator contig = pp.begin(); contig != pp.end(); ++contig)
                for (TPingPongSignaturesPerContig::iterator s = contig->second.begin(); s != contig->second.end(); ++s)
                    if ((s->readsOnPlusStrand >= minStackHeight) && (s->readsOnMinusStrand >= minStackHeight)) {
                        h.put((uint32_t)contig->first); h.put((uint32_t)s->position); h.put((float)s->fdr); h.put((float)s->readsOnPlusStrand); h.put((float)s->readsOnMinusStrand);
                        ++n;
                    }
            printf("%s\"%u\": {\"n\": %llu, \"sha256\": \"%s\"}", k ? ", " : "", minStackHeight, (unsigned long long)n, h.hex().c_str());
        }
        printf("}");
    }
    if (!annots.empty()) {
        findSuppressedTransposons(sigs, transposons);
        // per transposon in the reference's order (contig map, list sorted by start / end):
        //   sums    {u32 contig, strand, start, end; f32 R+, R-, hist[W]}   -- bit-exact contract
        //   pq      {f32 p, f32 q}                                          -- fp64 integral narrowed to float
        ppp_sha::Sha256 hs, hp;
        uint64_t n = 0;
        FILE* pq = pqPath.empty() ? nullptr : fopen(pqPath.c_str(), "wb");
        for (TTransposonsPerGenome::iterator c = transposons.begin(); c != transposons.end(); ++c)
            for (TTransposonsPerContig::iterator x = c->second.begin(); x != c->second.end(); ++x, ++n) {
                hs.put((uint32_t)c->first); hs.put((uint32_t)x->strand); hs.put((uint32_t)x->start); hs.put((uint32_t)x->end);
                hs.put((float)x->readsOnPlusStrand); hs.put((float)x->readsOnMinusStrand);
                for (unsigned k = 0; k < W; ++k) hs.put((float)(k < x->histogram.size() ? x->histogram[k] : 0.0f));
                hp.put((float)x->pValue); hp.put((float)x->qValue);
                if (pq) { float v[2] = {x->pValue, x->qValue}; fwrite(v, 4, 2, pq); }
            }
        if (pq) fclose(pq);
        printf(", \"n_transposons\": %llu, \"transposon_sums\": \"%s\", \"transposon_pq\": \"%s\"", (unsigned long long)n, hs.hex().c_str(), hp.hex().c_str());
    }
    printf(", \"seconds\": {\"count\": %.1f, \"heights\": %.1f, \"group\": %.1f, \"collapse_fdr\": %.1f}}\n", t1 - t0, t2 - t1, t3 - t2, t4 - t3);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc >= 2 && std::string(argv[1]) == "digest") return runDigest(argc, argv);
    if (argc >= 2 && std::string(argv[1]) == "dump") return runDump(argc, argv);
    if (argc >= 2 && std::string(argv[1]) == "time") return runTime(argc, argv);
    fprintf(stderr, "usage: ppp_harness dump OUT.bin [...] -i FILE... | ppp_harness time --genome G --reads N ...\n");
    return 2;
}
