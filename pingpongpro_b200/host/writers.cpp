// writers.cpp -- see writers.h.  Floats are printed like the reference prints them: std::ostream with
// ios::scientific and the default precision 6 (:883-889, :1387-1389).
#include "writers.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <thread>

namespace ppp {

namespace {
void replaceAll(std::string& s, const std::string& what, const std::string& with) {
    size_t i = 0;
    while ((i = s.find(what, i)) != std::string::npos) {
        s.replace(i, what.size(), with);
        i += with.size();
    }
}
}  // namespace

bool writeSignatures(const ppp_locus* loci, size_t n, const std::vector<std::string>& names, bool browserTracks) {
    std::ofstream tsv("ping-pong_signatures.tsv", std::ios_base::out);
    std::ofstream plus, minus, scores;
    if (browserTracks) {
        plus.open("ping-pong_signatures_read_stacks_on_plus_strand.bedGraph", std::ios_base::out);
        minus.open("ping-pong_signatures_read_stacks_on_minus_strand.bedGraph", std::ios_base::out);
        scores.open("ping-pong_signatures_scores.bedGraph", std::ios_base::out);
        if (plus.fail() || minus.fail() || scores.fail()) {
            std::cerr << "Failed to create browser track files for ping-pong signatures" << std::endl;  // :877
            return false;
        }
    }
    tsv.setf(std::ios::scientific, std::ios::floatfield);
    tsv << "contig\tposition\tFDR\tstackHeightOnPlusStrand\tstackHeightOnMinusStrand\n";  // :892
    if (browserTracks) {
        plus.setf(std::ios::scientific, std::ios::floatfield);
        minus.setf(std::ios::scientific, std::ios::floatfield);
        scores.setf(std::ios::scientific, std::ios::floatfield);
        plus << "track type=bedGraph name=\"read stacks on + strand\" description=\"height of read stacks on the + strand\" visibility=full\n";
        minus << "track type=bedGraph name=\"read stacks on - strand\" description=\"height of read stacks on the - strand\" visibility=full\n";
        scores << "track type=bedGraph name=\"scores\" description=\"scores of ping-pong signatures (1 - FDR)\" visibility=full viewLimits=0.0:1.0 autoScale=off\n";
    }
    // Rows are formatted on several threads (each with its own stream object carrying the same flags, so the text is
    // what the single stream would have produced) and appended to the files in order, one wave at a time.
    const size_t kChunk = 32768;
    const unsigned nThreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    struct Text { std::string tsv, plus, minus, scores; };
    std::vector<Text> wave(nThreads);
    for (size_t base = 0; base < n; base += kChunk * nThreads) {
        auto format = [&](unsigned t) {
            const size_t a = std::min(n, base + t * kChunk), b = std::min(n, a + kChunk);
            std::ostringstream o, p, m, sc;
            o.setf(std::ios::scientific, std::ios::floatfield);
            p.setf(std::ios::scientific, std::ios::floatfield);
            m.setf(std::ios::scientific, std::ios::floatfield);
            sc.setf(std::ios::scientific, std::ios::floatfield);
            for (size_t i = a; i < b; ++i) {
                const ppp_locus& l = loci[i];
                const std::string& contig = names[l.contig];
                o << contig << '\t' << l.position << '\t' << l.fdr << '\t' << l.reads_plus << '\t' << l.reads_minus << '\n';  // :905-910
                if (browserTracks) {
                    p << contig << '\t' << l.position << '\t' << (l.position + 1) << '\t' << l.reads_plus << '\n';
                    m << contig << '\t' << l.position << '\t' << (l.position + 1) << '\t' << l.reads_minus << '\n';
                    sc << contig << '\t' << l.position << '\t' << (l.position + 1) << '\t' << (1 - l.fdr) << '\n';  // :927
                }
            }
            wave[t].tsv = o.str();
            if (browserTracks) { wave[t].plus = p.str(); wave[t].minus = m.str(); wave[t].scores = sc.str(); }
        };
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nThreads && base + t * kChunk < n; ++t) th.emplace_back(format, t);
        format(0);
        for (std::thread& x : th) x.join();
        for (unsigned t = 0; t < nThreads && base + t * kChunk < n; ++t) {
            tsv << wave[t].tsv;
            if (browserTracks) { plus << wave[t].plus; minus << wave[t].minus; scores << wave[t].scores; }
        }
    }
    return true;
}

bool writeTransposons(const std::vector<ScoredTransposon>& rows, const std::vector<std::string>& names, bool browserTracks, std::string fileName,
                      double totalReadCount, int pp) {
    std::ofstream tsv((fileName + ".tsv").c_str(), std::ios_base::out);
    if (tsv.fail()) {
        std::cerr << "Failed to create TSV file for transposons" << std::endl;  // :1372
        return false;
    }
    std::ofstream bed;
    if (browserTracks) {
        bed.open((fileName + ".bed").c_str(), std::ios_base::out);
        if (bed.fail()) {
            std::cerr << "Failed to create browser track file for transposons" << std::endl;  // :1381
            return false;
        }
    }
    tsv.setf(std::ios::scientific, std::ios::floatfield);
    tsv << "identifier\tstrand\tcontig\tstart\tend\tpValue\tqValue\tpingPongReads\tnormalizedPingPongReads\tdiscardedPingPongReads\tstrandRatio\n";  // :1392
    if (browserTracks) {
        bed.setf(std::ios::scientific, std::ios::floatfield);
        replaceAll(fileName, "_", " ");  // :1396
        bed << "track name=\"" << fileName << "\" description=\"" << fileName << " shaded by ping-pong activity (1000 * (1 - q-value))\" useScore=1 visibility=dense\n";
    }
    for (const ScoredTransposon& s : rows) {
        const float pingPong = s.r.histogram[pp];
        const float rp = s.r.reads_plus, rm = s.r.reads_minus;
        tsv << s.t.identifier << '\t' << (s.t.strand == 0 ? '+' : '-') << '\t' << names[s.contig] << '\t' << s.t.start << '\t' << s.t.end << '\t'
            << s.r.p_value << '\t' << s.r.q_value << '\t' << pingPong << '\t'
            << pingPong / ((static_cast<float>(s.t.end) - s.t.start) / 1000) / (totalReadCount / 1000000) << '\t'  // :1413
            << ((rp + rm) - pingPong) << '\t'                                                                        // :1414
            << ((rm > 0) ? rp / rm : 1) << '\n';                                                                     // :1415
        if (browserTracks)
            bed << names[s.contig] << '\t' << s.t.start << '\t' << s.t.end << '\t' << s.t.identifier << '\t'
                << static_cast<int>(roundf((1 - s.r.q_value) * 1000)) << '\t' << (s.t.strand == 0 ? '+' : '-') << '\n';  // :1422
    }
    return true;
}

void plotHistograms(const std::string& fileName, const std::vector<std::string>& titles, const std::vector<std::vector<float> >& histograms, int minOverlap,
                    int maxOverlap, int pingpongOverlap) {
    std::ofstream r((fileName + ".R").c_str(), std::ios_base::out);
    if (r.fail()) {
        std::cerr << "Failed to create R script file" << std::endl;  // :783
        return;
    }
    const char nl = '\n';
    r << "histograms = data.frame(" << nl << "plotTitle = c(";
    for (size_t i = 0; i < titles.size(); ++i) {
        std::string title = titles[i];
        replaceAll(title, "\\", "\\\\");
        replaceAll(title, "'", "\\'");
        r << "'" << title << "'";
        if (i + 1 < titles.size()) r << "," << nl;
    }
    r << ")," << nl;
    for (int overlap = minOverlap; overlap <= maxOverlap; ++overlap) {
        r << nl << "overlap_";
        if (overlap < 0) r << "minus_";
        r << abs(overlap) << "=c(";
        for (size_t i = 0; i < histograms.size(); ++i) {
            if (i % 100 == 0) r << nl;
            r << histograms[i][overlap - minOverlap];
            if (i + 1 < histograms.size()) r << ",";
        }
        r << ")" << nl;
        if (overlap < maxOverlap) r << ",";
    }
    r << ")" << nl;
    const double minSd = 1E-10;  // MIN_STANDARD_DEVIATION, :226
    r << "# convert absolute values to z-scores" << nl
      << "means <- apply(histograms[,!colnames(histograms) %in% c('plotTitle', 'overlap_" << pingpongOverlap << "')], 1, mean)" << nl
      << "sds <- apply(histograms[,!colnames(histograms) %in% c('plotTitle', 'overlap_" << pingpongOverlap << "')], 1, sd)" << nl
      << "sds <- ifelse(sds < " << minSd << "," << minSd << ", sds)" << nl
      << "for (column in colnames(histograms[,colnames(histograms) != 'plotTitle'])) {" << nl
      << "\thistograms[,column] = (histograms[,column] - means) / sds" << nl
      << "}" << nl
      << "# save plots to a single PDF" << nl
      << "pdf('" << fileName << ".pdf', onefile=TRUE)" << nl
      << "par(font.lab=2, mar=c(5.1, 5.1, 5.1, 2.1))" << nl
      << "# draw a red bar for ping-pong signatures and a grey bar for arbitrary overlaps" << nl
      << "barColors <- ifelse(colnames(histograms[,colnames(histograms) != 'plotTitle']) != 'overlap_" << pingpongOverlap << "', rgb(0.7,0.7,0.7), rgb(0.8,0.4,0.4))" << nl
      << "# every row of the data frame <histograms> is rendered as a barplot" << nl
      << "for (i in 1:nrow(histograms)) {" << nl
      << "\thistogram <- histograms[i,]" << nl
      << "\tplotTitle <- histogram$plotTitle" << nl
      << "\thistogram <- c(t(histogram[, colnames(histogram) != 'plotTitle']))" << nl
      << "\tplot(0, 0, xlim=c(0,length(histogram)), type='n', xlab='overlap', ylim=c(min(c(0, histogram)), max(histogram)), ylab='z-score', xaxt='n', main=plotTitle, cex.axis=0.75)" << nl
      << "\taxis(1, at=1:length(histogram)-0.25, labels=" << minOverlap << ":" << maxOverlap << ", cex.axis=0.75)" << nl
      << "\tbarplot(histogram, col=barColors, border=NA, axes=FALSE, add=TRUE, width=0.5, space=1)" << nl
      << "}" << nl
      << "garbage <- dev.off()" << nl;
    r.close();
    std::string cmd = "Rscript -e 'source(\"" + fileName + ".R\")'";  // :853-854
    if (system(cmd.c_str()) != 0) { /* Rscript is optional: the script itself is the deliverable */ }
}

}  // namespace ppp
