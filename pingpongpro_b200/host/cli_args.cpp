// cli_args.cpp -- see cli_args.h.
#include "cli_args.h"

#include <cctype>
#include <cstdlib>
#include <cstring>
#include <iostream>

#include "ppp_types.h"

namespace ppp {

namespace {

const char* kApp = "pingpongpro";

bool endsWithNoCase(const std::string& s, const char* ext) {
    size_t n = s.size(), e = strlen(ext);
    if (e > n) return false;
    for (size_t i = 0; i < e; ++i)
        if (tolower((unsigned char)s[n - e + i]) != tolower((unsigned char)ext[i])) return false;
    return true;
}

bool hasExtension(const std::string& s, std::initializer_list<const char*> exts) {
    for (const char* e : exts)
        if (endsWithNoCase(s, e)) return true;
    return false;
}

bool parseInt(const std::string& text, long minValue, unsigned& out, const char* optionName) {
    char* end = nullptr;
    long v = strtol(text.c_str(), &end, 10);
    if (text.empty() || *end != 0) {
        std::cerr << kApp << ": the given value '" << text << "' cannot be casted to integer (option --" << optionName << ")" << std::endl;
        return false;
    }
    if (v < minValue) {
        std::cerr << kApp << ": the given value '" << text << "' is not in the interval [" << minValue << ":+inf] (option --" << optionName << ")" << std::endl;
        return false;
    }
    out = (unsigned)v;
    return true;
}

void printHelp(std::ostream& os) {
    os << "pingpongpro - Find ping-pong signatures like a pro\n"
          "==================================================\n\n"
          "SYNOPSIS\n    pingpongpro [OPTIONS] -i SAM_INPUT_FILE [-o OUTPUT_DIRECTORY]\n\n"
          "DESCRIPTION\n"
          "    PingPongPro scans piRNA-Seq data for signs of ping-pong cycle activity. The ping-pong cycle produces piRNA\n"
          "    molecules with complementary 5'-ends. These molecules appear as stacks of aligned reads whose 5'-ends overlap\n"
          "    with the 5'-ends of reads on the opposite strand by exactly 10 bases.\n\n"
          "    -h, --help\n          Displays this help message.\n"
          "    --version\n          Display version information\n"
          "    -b, --browserTracks\n          Generate genome browser tracks for loci with ping-pong signature and (if -t or -T is specified) for\n"
          "          transposons with ping-pong activity. Default: off.\n"
          "    -s, --min-stack-height NUMBER_OF_READS\n          Omit stacks with fewer than the specified number of reads from the output. In range [0..inf]. Default: 0.\n"
          "    -i, --input PATH\n          Input file in SAM/BAM format. Valid filetypes are: bam and sam.\n"
          "    -l, --min-alignment-length LENGTH\n          Ignore alignments in the input file that are shorter than the specified length. In range [1..inf]. Default: 24.\n"
          "    -L, --max-alignment-length LENGTH\n          Ignore alignments in the input file that are longer than the specified length. In range [1..inf]. Default: 32.\n"
          "    -m, --multi-hits METHOD\n          How to count multi-mapping reads. One of weighted, discard, and unique. Default: weighted.\n"
          "    -o, --output PATH\n          Write output to specified directory. Default: current working directory.\n"
          "    -p, --plot\n          Generate R plots on how z-scores are calculated for ping-pong signatures and (if -t or -T is specified) for\n"
          "          transposons. Requires Rscript. Default: off.\n"
          "    -t, --transposons PATH\n          Check if the transposons given in the file PATH are suppressed through ping-pong activity. Valid filetypes\n"
          "          are: bed, csv, gff, gtf, and tsv.\n"
          "    -T, --predict-transposons RANGE\n          Predict the location of suppressed transposons based on regions with high ping-pong activity. Consider\n"
          "          adjacent ping-pong signatures within a range of RANGE to belong to the same transposon. Default: off. In\n"
          "          range [30..inf].\n"
          "    -v, --verbose\n          Print messages about the current progress to stderr. Default: off.\n\n"
          "B200 BUILD\n"
          "    --max-overlap NT, --min-overlap NT\n          Background overlap range (compile-time constants 3 and 23 in the reference).\n"
          "    --device ORDINAL\n          CUDA device to run on. Default: 0.\n"
          "    --gpus N\n          Cut the genome into N position ranges, one per GPU starting at --device. Default: 1.\n"
          "    --threads N\n          Worker threads that inflate BGZF blocks of BAM input. Default: all hardware threads.\n\n"
          "VERSION\n    pingpongpro version: 1.0 (B200)\n    Last update Apr 2014\n";
}

}  // namespace

ParseResult parseCommandLine(AppOptions& o, int argc, char const** argv) {
    struct Spec { const char* shortName; const char* longName; bool takesValue; };
    static const Spec specs[] = {
        {"b", "browserTracks", false}, {"s", "min-stack-height", true}, {"i", "input", true}, {"l", "min-alignment-length", true},
        {"L", "max-alignment-length", true}, {"m", "multi-hits", true}, {"o", "output", true}, {"p", "plot", false},
        {"t", "transposons", true}, {"T", "predict-transposons", true}, {"v", "verbose", false},
        {"", "max-overlap", true}, {"", "min-overlap", true}, {"", "device", true}, {"", "gpus", true}, {"", "threads", true},
    };
    bool predictSet = false;
    std::string multiHits = "weighted";
    // combined short flags (-bv, -bs 2): expanded into single ones; only the last of a group may take a value
    std::vector<std::string> args;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        bool expanded = false;
        if (a.size() > 2 && a[0] == '-' && a[1] != '-') {
            bool ok = true;
            for (size_t k = 1; k < a.size() && ok; ++k) {
                const Spec* sp = nullptr;
                for (const Spec& s : specs)
                    if (s.shortName[0] && s.shortName[0] == a[k] && !s.shortName[1]) sp = &s;
                ok = sp && (!sp->takesValue || k + 1 == a.size());
            }
            if (ok) {
                for (size_t k = 1; k < a.size(); ++k) args.push_back(std::string("-") + a[k]);
                expanded = true;
            }
        }
        if (!expanded) args.push_back(a);
    }
    const int nArgs = (int)args.size();
    for (int i = 0; i < nArgs; ++i) {
        std::string a = args[i];
        if (a == "-h" || a == "--help") { printHelp(std::cout); return PARSE_HELP; }
        if (a == "--version") { std::cout << kApp << " version: 1.0 (B200)\nLast update Apr 2014" << std::endl; return PARSE_VERSION; }
        std::string name, value;
        bool inlineValue = false;
        if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            size_t eq = a.find('=');
            name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            if (eq != std::string::npos) { inlineValue = true; value = a.substr(eq + 1); }
        } else if (a.size() == 2 && a[0] == '-') {
            name = a.substr(1);
        } else {
            std::cerr << kApp << ": Too many arguments!" << std::endl;
            return PARSE_ERROR;
        }
        const Spec* spec = nullptr;
        for (const Spec& s : specs)
            if ((name.size() == 1 && name == s.shortName) || (name.size() > 1 && name == s.longName)) spec = &s;
        if (!spec) { std::cerr << kApp << ": Unknown option: " << a << std::endl; return PARSE_ERROR; }
        std::string ln = spec->longName;
        if (spec->takesValue && !inlineValue) {
            if (i + 1 >= nArgs) { std::cerr << kApp << ": option requires an argument -- " << name << std::endl; return PARSE_ERROR; }
            value = args[++i];
        }
        unsigned u = 0;
        if (ln == "browserTracks") o.browserTracks = true;
        else if (ln == "plot") o.plot = true;
        else if (ln == "verbose") o.verbosity = 3;
        else if (ln == "min-stack-height") { if (!parseInt(value, 0, o.minStackHeight, "min-stack-height")) return PARSE_ERROR; }
        else if (ln == "min-alignment-length") { if (!parseInt(value, 1, o.minAlignmentLength, "min-alignment-length")) return PARSE_ERROR; }
        else if (ln == "max-alignment-length") { if (!parseInt(value, 1, o.maxAlignmentLength, "max-alignment-length")) return PARSE_ERROR; }
        else if (ln == "predict-transposons") { if (!parseInt(value, 30, o.predictTransposonsRange, "predict-transposons")) return PARSE_ERROR; predictSet = true; }
        else if (ln == "max-overlap") { if (!parseInt(value, 0, u, "max-overlap")) return PARSE_ERROR; o.maxOverlap = (int)u; }
        else if (ln == "min-overlap") { if (!parseInt(value, 0, u, "min-overlap")) return PARSE_ERROR; o.minOverlap = (int)u; }
        else if (ln == "device") { if (!parseInt(value, 0, u, "device")) return PARSE_ERROR; o.device = (int)u; }
        else if (ln == "gpus") { if (!parseInt(value, 1, u, "gpus")) return PARSE_ERROR; o.gpus = (int)u; }
        else if (ln == "threads") { if (!parseInt(value, 0, u, "threads")) return PARSE_ERROR; o.threads = (int)u; }
        else if (ln == "input") {
            if (!hasExtension(value, {".bam", ".sam"})) {
                std::cerr << kApp << ": the given path '" << value << "' does not have one of the valid file extensions [*.bam, *.sam] (option --input)" << std::endl;
                return PARSE_ERROR;
            }
            o.inputFiles.push_back(value);
        } else if (ln == "transposons") {
            if (!hasExtension(value, {".bed", ".csv", ".gff", ".gtf", ".tsv"})) {
                std::cerr << kApp << ": the given path '" << value << "' does not have one of the valid file extensions [*.bed, *.csv, *.gff, *.gtf, *.tsv] (option --transposons)" << std::endl;
                return PARSE_ERROR;
            }
            o.transposonFiles.push_back(value);
        } else if (ln == "multi-hits") {
            if (value != "weighted" && value != "discard" && value != "unique") {
                std::cerr << kApp << ": the given value '" << value << "' is not in the list of allowed values [weighted, discard, unique] (option --multi-hits)" << std::endl;
                return PARSE_ERROR;
            }
            multiHits = value;
        } else if (ln == "output") o.output = value;
    }
    if (o.inputFiles.empty()) { std::cerr << kApp << ": Missing value for option: -i" << std::endl; return PARSE_ERROR; }
    o.countMultiHits = multiHits == "unique" ? PPP_MULTIHITS_UNIQUE : multiHits == "discard" ? PPP_MULTIHITS_DISCARD : PPP_MULTIHITS_WEIGHTED;  // :303-314
    if (o.minAlignmentLength > o.maxAlignmentLength) {  // :321-325
        std::cerr << kApp << ": maximum alignment length (" << o.maxAlignmentLength << ") must not be lower than minimum alignment length ("
                  << o.minAlignmentLength << ")" << std::endl;
        return PARSE_ERROR;
    }
    if (!o.output.empty() && o.output[o.output.size() - 1] != '/') o.output += '/';  // :328-329
    if (!predictSet) o.predictTransposonsRange = 0;                                   // :337-344
    if (o.minOverlap > 10 || o.maxOverlap < 10 || o.maxOverlap > 32 || o.maxOverlap - o.minOverlap < 2) {
        std::cerr << kApp << ": overlaps must satisfy min <= 10 <= max <= 32 with at least 3 overlaps" << std::endl;
        return PARSE_ERROR;
    }
    return PARSE_OK;
}

}  // namespace ppp
