// cli_args.h -- command line of the drop-in `pingpongpro` binary: the same flags, defaults, ranges and exit
// behaviour as parseCommandLine (pingpongpro.cpp:239-356, SeqAn ArgumentParser), re-implemented without SeqAn.
#pragma once
#include <string>
#include <vector>

namespace ppp {

struct AppOptions {
    bool browserTracks = false;                 // -b (:251)
    unsigned minStackHeight = 0;                // -s (:253-255)
    std::vector<std::string> inputFiles;        // -i, list, required, .bam/.sam (:257-259)
    unsigned minAlignmentLength = 24;           // -l (:261-263)
    unsigned maxAlignmentLength = 32;           // -L (:264-266)
    int countMultiHits = 0;                     // -m weighted|discard|unique (:268-270) -> PPP_MULTIHITS_*
    std::string output;                         // -o, trailing '/' appended (:327-329)
    bool plot = false;                          // -p (:274)
    std::vector<std::string> transposonFiles;   // -t, list, .bed .csv .gff .gtf .tsv (:276-277)
    unsigned predictTransposonsRange = 0;       // -T, >= 30 (:279-282)
    unsigned verbosity = 0;                     // -v -> 3 (:346-353)
    // extensions of the B200 build (not in the reference; its values are compile-time constants :108-113)
    int minOverlap = 3, maxOverlap = 23;        // --min-overlap / --max-overlap
    int device = 0;                             // --device
    int gpus = 1;                               // --gpus: position-range shards, one CUDA context each
    int threads = 0;                            // --threads: BGZF inflate workers (0 = all hardware threads)
};

enum ParseResult { PARSE_OK = 0, PARSE_ERROR, PARSE_HELP, PARSE_VERSION };

// Any result other than PARSE_OK makes main() return 1, help and version included (pingpongpro.cpp:1465-1466).
ParseResult parseCommandLine(AppOptions& options, int argc, char const** argv);

}  // namespace ppp
