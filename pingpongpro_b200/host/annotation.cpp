// annotation.cpp -- see annotation.h.
#include "annotation.h"

#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <sstream>

namespace ppp {

namespace {

bool endsWithNoCase(const std::string& s, const char* ext) {
    size_t n = s.size(), e = strlen(ext);
    if (e > n) return false;
    for (size_t i = 0; i < e; ++i)
        if (tolower((unsigned char)s[n - e + i]) != tolower((unsigned char)ext[i])) return false;
    return true;
}

// Column layout of one annotation format (:1003-1047).
struct Layout {
    unsigned identifier, strand, contig, start, end;  // 1-based column numbers
    int offset;                                        // 1 when the format counts positions from 1
    int closedEnd;                                     // 1 when the end coordinate is inclusive in the file
    char delimiter;
};

Layout layoutOf(FileFormat f) {
    switch (f) {
        case FORMAT_BED: return Layout{4, 6, 1, 2, 3, 0, 0, ' '};
        case FORMAT_CSV: return Layout{1, 2, 3, 4, 5, 1, 0, ','};
        case FORMAT_GFF:
        case FORMAT_GTF: return Layout{9, 7, 2, 4, 5, 1, 1, '\t'};
        default: return Layout{1, 2, 3, 4, 5, 1, 0, '\t'};
    }
}

// The record being assembled from the fields of the current line.
struct PendingLine {
    std::string identifier;
    int strand = -1, contig = -1, start = -1, end = -1;
    void clear() { identifier.clear(); strand = contig = start = end = -1; }
    bool complete() const { return strand >= 0 && contig >= 0 && start >= 0 && end >= 0 && !identifier.empty(); }  // :1136
};

// atoi with the reference's error convention: 0 from anything but "0" is a conversion error (:1110-1114).
int coordinate(const std::string& text, int shift) {
    int v = atoi(text.c_str());
    if (v == 0 && text != "0") return -1;
    return v + shift;
}

}  // namespace

FileFormat formatFromPath(const std::string& path) {
    if (endsWithNoCase(path, ".bed")) return FORMAT_BED;
    if (endsWithNoCase(path, ".csv")) return FORMAT_CSV;
    if (endsWithNoCase(path, ".gff")) return FORMAT_GFF;
    if (endsWithNoCase(path, ".gtf")) return FORMAT_GTF;
    return FORMAT_TSV;
}

void readTransposons(std::istream& in, FileFormat format, TransposonsPerGenome& transposons, std::vector<std::string>& names) {
    const Layout L = layoutOf(format);
    const bool csv = format == FORMAT_CSV;
    PendingLine line;
    std::string field;
    unsigned column = 1;
    bool lineEnds = false;   // a line feed was seen; the record is closed by the next field boundary
    bool quoted = false;     // inside CSV double quotes

    auto storeField = [&]() {
        if (column == L.identifier) line.identifier = field;
        else if (column == L.strand) {
            if (field == "+") line.strand = 0;
            else if (field == "-") line.strand = 1;
        } else if (column == L.contig) {
            for (size_t i = 0; line.contig == -1 && i < names.size(); ++i)
                if (names[i] == field) line.contig = (int)i;
            if (line.contig == -1) {  // unknown contig: the name store grows (:1101-1106)
                names.push_back(field);
                line.contig = (int)names.size() - 1;
            }
        } else if (column == L.start) line.start = coordinate(field, -L.offset);
        else if (column == L.end) line.end = coordinate(field, -L.offset + L.closedEnd);
    };

    while (in.good()) {
        char ch = (char)in.get();
        if (!in.good()) ch = '\n';  // the last line is processed even without a terminating line feed (:1067-1068)
        if (ch == '\n') { lineEnds = true; ch = L.delimiter; }
        else if (L.delimiter == ' ' && ch == '\t') ch = ' ';  // BED: any blank separates (:1075-1078)
        const bool boundary = ch == L.delimiter && !(csv && quoted);
        if (boundary) {
            if (field.empty() && !lineEnds) continue;  // empty fields are skipped, not counted (:1082)
            storeField();
            if (lineEnds) {
                if (line.complete()) {
                    Transposon t;
                    t.identifier = line.identifier;
                    t.strand = (unsigned)line.strand;
                    t.start = (unsigned)line.start;
                    t.end = (unsigned)line.end;
                    transposons[(unsigned)line.contig].push_back(t);
                }
                lineEnds = false;
                line.clear();
                column = 1;
            } else {
                ++column;
            }
            field.clear();
            continue;
        }
        if (ch == '\r' && in.peek() == '\n') continue;  // CR of a CR LF pair (:1155)
        if (csv && ch == '"') {                         // :1157-1163
            if (!field.empty() && !quoted) field += ch;  // a doubled quote inside a field collapses to one
            quoted = !quoted;
        } else {
            field += ch;
        }
    }
    sortTransposons(transposons);
}

void sortTransposons(TransposonsPerGenome& transposons) {
    for (auto& kv : transposons)
        std::stable_sort(kv.second.begin(), kv.second.end(), [](const Transposon& a, const Transposon& b) {
            if (a.start == b.start) return a.end < b.end;
            return a.start < b.start;
        });
}

void predictTransposons(const std::map<unsigned, std::vector<unsigned> >& positionsPerContig, const std::vector<std::string>& names, unsigned range,
                        TransposonsPerGenome& out) {
    for (const auto& kv : positionsPerContig) {
        const std::vector<unsigned>& pos = kv.second;
        if (pos.empty()) break;  // :1327
        unsigned start = pos[0], end = start + 1;
        for (unsigned p : pos) {
            if (end + range >= p) {  // :1333 (unsigned arithmetic)
                end = p + 1;
            } else {
                if (start + 30u < end) {  // PREDICT_TRANSPOSONS_MIN_LENGTH, :220, :1339
                    Transposon t;
                    t.identifier = names[kv.first] + ":" + std::to_string(start) + "-" + std::to_string(end);  // :1341-1343
                    t.strand = 0;
                    t.start = start;
                    t.end = end;
                    out[kv.first].push_back(t);
                }
                start = p;
                end = start + 1;
            }
        }
    }
}

}  // namespace ppp
