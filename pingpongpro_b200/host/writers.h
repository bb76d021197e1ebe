// writers.h -- output tables of the drop-in CLI, byte-identical to writePingPongSignaturesToFile
// (pingpongpro.cpp:863-939) and writeTransposonsToFile (:1366-1430), plus the R scripts of :777-855.
#pragma once
#include <string>
#include <vector>

#include "annotation.h"
#include "ppp_b200.h"

namespace ppp {

// ping-pong_signatures.tsv (+ three bedGraph tracks with -b) in the current directory.
// `loci` are the overlap-10 signatures that pass -s, in (contig, position) order.  Returns false after printing
// the reference's message when a browser-track file cannot be created (:877).
bool writeSignatures(const ppp_locus* loci, size_t n, const std::vector<std::string>& names, bool browserTracks);

struct ScoredTransposon {
    Transposon t;
    unsigned contig = 0;
    ppp_tp_result r;
};

// <fileName>.tsv (+ .bed with -b), rows in (contig, start, end) order.
bool writeTransposons(const std::vector<ScoredTransposon>& rows, const std::vector<std::string>& names, bool browserTracks, std::string fileName,
                      double totalReadCount, int pingpongIndex);

// R script <fileName>.R that draws one z-score bar plot per histogram (:777-846); runs Rscript when it is installed.
void plotHistograms(const std::string& fileName, const std::vector<std::string>& titles, const std::vector<std::vector<float> >& histograms, int minOverlap,
                    int maxOverlap, int pingpongOverlap);

}  // namespace ppp
