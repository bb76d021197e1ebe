// annotation.h -- transposon annotation input (BED / CSV / GFF / GTF / TSV) with the exact field rules of
// readTransposonsFromFile (pingpongpro.cpp:993-1175), and the clustering of predictSuppressedTransposons
// (:1324-1354).  Host side; feeds ppp_transposons().
#pragma once
#include <cstdint>
#include <istream>
#include <map>
#include <string>
#include <vector>

namespace ppp {

enum FileFormat { FORMAT_BED, FORMAT_CSV, FORMAT_GFF, FORMAT_GTF, FORMAT_TSV };

// File format from the extension, as main() decides it (:1543-1554); anything else is TSV.
FileFormat formatFromPath(const std::string& path);

struct Transposon {         // TTransposon, :173-213
    std::string identifier;
    unsigned strand = 0;    // STRAND_PLUS 0 / STRAND_MINUS 1
    unsigned start = 0;
    unsigned end = 0;       // treated as inclusive downstream (:1224)
};

// contig id -> transposons in file order (std::list push_back order of :1137); sorted later by sortTransposons().
typedef std::map<unsigned, std::vector<Transposon> > TransposonsPerGenome;

// Appends the transposons of one file.  `names` is the contig name store; names that it does not contain are
// appended (:1101-1106), exactly when the contig field is read, whether or not the line survives.
void readTransposons(std::istream& in, FileFormat format, TransposonsPerGenome& transposons, std::vector<std::string>& names);

// Stable sort of every contig's list by (start, end) (:206-212, :1174); called after each file like the reference.
void sortTransposons(TransposonsPerGenome& transposons);

// predictSuppressedTransposons, :1326-1354: clusters the overlap-10 signature positions of each contig (ascending
// contig id, ascending position).  `positionsPerContig` holds ONLY contigs that have an entry in the reference's
// signature map; a contig with an empty list ends the walk (:1327).  The last cluster of a contig is never
// emitted.
void predictTransposons(const std::map<unsigned, std::vector<unsigned> >& positionsPerContig, const std::vector<std::string>& names,
                        unsigned range, TransposonsPerGenome& out);

}  // namespace ppp
