// shim_impl.h -- minimal stand-in for the SeqAn 1.4.x API surface that
// /root/reference/pingpongpro.cpp uses.  TEST INFRASTRUCTURE ONLY:
// it exists so the UNMODIFIED reference source compiles here as the parity oracle
// (SeqAn is neither vendored in the reference nor installed in this image).
//
// No arithmetic of the hot path lives here: SeqAn is used by the reference for
// argument parsing and record decoding only (pingpongpro.cpp:239-356, :402-415,
// :440-444, :461, :1482-1520).  Record decoding is delegated to the project's own
// reader (pingpongpro_b200/host/aln_reader.h) so that the oracle and the new front
// end see identical records.
//
// An extra, non-SeqAn entry exists for the timing harness: a BamStream opened on the
// pseudo path "@mem" serves records from seqan::shim::memRecords() instead of a file,
// which lets the reference's countReadsInBamFile be timed without text parsing; "@gen"
// pulls them from a callback instead (full-size digests: 500 M records never fit in memory).
#pragma once

#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "aln_reader.h"

namespace seqan {

// ---------------------------------------------------------------- strings ------
struct CharString : public std::string {
    CharString() {}
    CharString(const char* s) : std::string(s) {}
    CharString(const std::string& s) : std::string(s) {}
};

inline const char* toCString(const CharString& s) { return s.c_str(); }
inline const char* toCString(const std::string& s) { return s.c_str(); }

template <typename T>
struct StringSet {
    std::vector<T> v;
    T& operator[](size_t i) { return v[i]; }
    const T& operator[](size_t i) const { return v[i]; }
};

template <typename T>
struct Iterator;
template <typename T>
struct Iterator<StringSet<T> > {
    typedef typename std::vector<T>::iterator Type;
};

template <typename T>
inline typename std::vector<T>::iterator begin(StringSet<T>& s) { return s.v.begin(); }
template <typename T>
inline typename std::vector<T>::iterator end(StringSet<T>& s) { return s.v.end(); }
template <typename TIt>
inline typename std::iterator_traits<TIt>::reference value(TIt it) { return *it; }
template <typename T, typename S>
inline void appendValue(StringSet<T>& set, const S& s) { set.v.push_back(T(s)); }

// SeqAn-style String<T>: only what the reference touches (operator[], length()).
template <typename T>
struct String {
    std::vector<T> v;
    T& operator[](size_t i) { return v[i]; }
    const T& operator[](size_t i) const { return v[i]; }
};

inline size_t length(const std::string& s) { return s.size(); }
template <typename T>
inline size_t length(const StringSet<T>& s) { return s.v.size(); }
template <typename T>
inline size_t length(const String<T>& s) { return s.v.size(); }

// Case-insensitive "does the file name end with this extension".
inline bool _compareExtension(const char* name, const char* ext) {
    size_t n = strlen(name), e = strlen(ext);
    if (e > n) return false;
    for (size_t i = 0; i < e; ++i)
        if (tolower((unsigned char)name[n - e + i]) != tolower((unsigned char)ext[i])) return false;
    return true;
}

// ---------------------------------------------------------------- BAM I/O ------
struct CigarElement {
    char operation;
    unsigned count;
};

struct BamTags {  // the shim carries the only tag the reference reads
    bool hasNH = false;
    unsigned nh = 1;
};

struct BamAlignmentRecord {
    static const int INVALID_POS = 2147483647;
    int rID = -1;
    int beginPos = -1;
    unsigned flag = 0;
    String<CigarElement> cigar;
    std::string seq;
    BamTags tags;
};

namespace shim {
struct MemRecord {
    int rID, beginPos;
    unsigned flag;
    std::vector<CigarElement> cigar;
    std::string seq;
    bool hasNH;
    unsigned nh;
};
inline std::vector<MemRecord>& memRecords() { static std::vector<MemRecord> r; return r; }
// "@gen": fills `rec` with the next record and returns true, or returns false at the end of the stream
typedef bool (*GenFn)(MemRecord& rec, void* user);
struct GenSource {
    GenFn fn = nullptr;
    void* user = nullptr;
};
inline GenSource& genSource() { static GenSource g; return g; }
inline std::vector<std::string>& memNames() { static std::vector<std::string> n; return n; }
}  // namespace shim

struct BamIOContext {
    StringSet<CharString> names;
};

struct BamStream {
    BamIOContext bamIOContext;
    ppp::AlnReader reader;
    bool ok = false;
    bool mem = false, gen = false, genHave = false;
    size_t memPos = 0;
    shim::MemRecord genRec;
    explicit BamStream(const char* path) {
        if (strcmp(path, "@gen") == 0) {
            gen = ok = true;
            for (const std::string& n : shim::memNames()) bamIOContext.names.v.push_back(CharString(n));
            return;
        }
        if (strcmp(path, "@mem") == 0) {
            mem = ok = true;
            for (const std::string& n : shim::memNames()) bamIOContext.names.v.push_back(CharString(n));
            return;
        }
        ok = reader.open(path);
        if (ok)
            for (const std::string& n : reader.names()) bamIOContext.names.v.push_back(CharString(n));
    }
};

inline bool isGood(BamStream& s) { return s.ok; }
inline bool atEnd(BamStream& s) {
    if (s.gen) {
        if (!s.genHave) s.genHave = shim::genSource().fn(s.genRec, shim::genSource().user);
        return !s.genHave;
    }
    return s.mem ? s.memPos >= shim::memRecords().size() : s.reader.atEnd();
}
inline void close(BamStream& s) { if (!s.mem && !s.gen) s.reader.close(); }
inline StringSet<CharString>& nameStore(BamIOContext& c) {
    return c.names;
}

inline int readRecord(BamAlignmentRecord& rec, BamStream& s) {
    if (s.mem || s.gen) {
        if (s.gen && !s.genHave && atEnd(s)) return 1;
        const shim::MemRecord& m = s.gen ? s.genRec : shim::memRecords()[s.memPos++];
        s.genHave = false;
        rec.rID = m.rID;
        rec.beginPos = m.beginPos;
        rec.flag = m.flag;
        rec.cigar.v = m.cigar;
        rec.seq = m.seq;
        rec.tags.hasNH = m.hasNH;
        rec.tags.nh = m.nh;
        return 0;
    }
    ppp::AlnView v;
    if (s.reader.next(v) != 0) return 1;
    // names may have been extended by contigs that are missing from the header
    while (s.bamIOContext.names.v.size() < s.reader.names().size())
        s.bamIOContext.names.v.push_back(CharString(s.reader.names()[s.bamIOContext.names.v.size()]));
    rec.rID = v.rID;
    rec.beginPos = v.beginPos;
    rec.flag = v.flag;
    rec.cigar.v.resize(v.nCigar);
    for (unsigned i = 0; i < v.nCigar; ++i) {
        rec.cigar.v[i].operation = ppp::cigar_op_char(v.cigar[i]);
        rec.cigar.v[i].count = ppp::cigar_op_count(v.cigar[i]);
    }
    rec.seq.resize(v.lSeq);
    for (unsigned i = 0; i < v.lSeq; ++i) rec.seq[i] = v.base(i);
    rec.tags.hasNH = v.hasNH;
    rec.tags.nh = v.nh;
    return 0;
}

inline bool hasFlagRC(const BamAlignmentRecord& r) { return (r.flag & 0x10u) != 0; }

struct BamTagsDict {
    const BamTags& t;
    explicit BamTagsDict(const BamTags& tags) : t(tags) {}
};
inline bool findTagKey(unsigned& idx, const BamTagsDict& d, const char* key) {
    if (key[0] == 'N' && key[1] == 'H' && d.t.hasNH) { idx = 0; return true; }
    return false;
}
inline bool extractTagValue(unsigned& dest, const BamTagsDict& d, unsigned /*idx*/) {
    dest = d.t.nh;
    return true;
}

// ------------------------------------------------------------- arg parsing -----
struct ArgParseArgument {
    enum ArgumentType { STRING, INTEGER, DOUBLE, INPUTFILE, OUTPUTFILE };
};

struct ArgParseOption {
    std::string shortName, longName, help, label;
    bool isFlag = true, isList = false, required = false;
    ArgParseArgument::ArgumentType type = ArgParseArgument::STRING;
    std::vector<std::string> validValues;
    bool hasMin = false;
    long minValue = 0;
    bool hasDefault = false;
    std::string defaultValue;
    std::vector<std::string> values;
    bool set = false;
    ArgParseOption(const char* s, const char* l, const char* h) : shortName(s), longName(l), help(h) {}
    ArgParseOption(const char* s, const char* l, const char* h, ArgParseArgument::ArgumentType t, const char* lab = "", bool list = false)
        : shortName(s), longName(l), help(h), label(lab), isFlag(false), isList(list), type(t) {}
};

struct ArgumentParser {
    enum ParseResult { PARSE_OK, PARSE_ERROR, PARSE_HELP, PARSE_VERSION, PARSE_WRITE_CTD, PARSE_EXPORT_HELP };
    std::string appName, usage, shortDescription, version, date;
    std::vector<std::string> description;
    std::vector<ArgParseOption> options;
    explicit ArgumentParser(const char* name) : appName(name) {}
    ArgParseOption* find(const std::string& n) {
        for (ArgParseOption& o : options)
            if (o.longName == n || o.shortName == n) return &o;
        return nullptr;
    }
};

inline void addUsageLine(ArgumentParser& p, const char* s) { p.usage = s; }
inline void setShortDescription(ArgumentParser& p, const char* s) { p.shortDescription = s; }
inline void addDescription(ArgumentParser& p, const char* s) { p.description.push_back(s); }
inline void setVersion(ArgumentParser& p, const char* s) { p.version = s; }
inline void setDate(ArgumentParser& p, const char* s) { p.date = s; }
inline void addOption(ArgumentParser& p, const ArgParseOption& o) { p.options.push_back(o); }
template <typename T>
inline void setDefaultValue(ArgumentParser& p, const char* name, const T& v) {
    std::ostringstream ss;
    ss << v;
    ArgParseOption* o = p.find(name);
    o->hasDefault = true;
    o->defaultValue = ss.str();
}
inline void setMinValue(ArgumentParser& p, const char* name, const std::string& v) {
    ArgParseOption* o = p.find(name);
    o->hasMin = true;
    o->minValue = atol(v.c_str());
}
inline void setValidValues(ArgumentParser& p, const char* name, const char* values) {
    std::istringstream ss(values);
    std::string tok;
    ArgParseOption* o = p.find(name);
    while (ss >> tok) o->validValues.push_back(tok);
}
inline void setRequired(ArgumentParser& p, const char* name) { p.find(name)->required = true; }
inline const std::string& getAppName(const ArgumentParser& p) { return p.appName; }
inline bool isSet(ArgumentParser& p, const char* name) { return p.find(name)->set; }
inline size_t getOptionValueCount(ArgumentParser& p, const char* name) {
    ArgParseOption* o = p.find(name);
    if (!o->values.empty()) return o->values.size();
    return o->hasDefault ? 1 : 0;
}
inline bool _optionText(std::string& out, ArgumentParser& p, const char* name, size_t idx) {
    ArgParseOption* o = p.find(name);
    if (idx < o->values.size()) { out = o->values[idx]; return true; }
    if (o->values.empty() && o->hasDefault && idx == 0) { out = o->defaultValue; return true; }
    return false;
}
inline bool getOptionValue(unsigned& v, ArgumentParser& p, const char* name, size_t idx = 0) {
    std::string t;
    if (!_optionText(t, p, name, idx)) return false;
    v = (unsigned)strtol(t.c_str(), nullptr, 10);
    return true;
}
inline bool getOptionValue(std::string& v, ArgumentParser& p, const char* name, size_t idx = 0) {
    std::string t;
    if (!_optionText(t, p, name, idx)) return false;
    v = t;
    return true;
}

inline void _printHelp(const ArgumentParser& p, std::ostream& os) {
    os << p.appName << " - " << p.shortDescription << "\n\nSYNOPSIS\n    " << p.appName << " " << p.usage << "\n\nOPTIONS\n";
    for (const ArgParseOption& o : p.options) os << "    -" << o.shortName << ", --" << o.longName << (o.isFlag ? "" : " " + o.label) << "\n          " << o.help << "\n";
    os << "\nVERSION\n    " << p.appName << " version: " << p.version << "\n    Last update " << p.date << "\n";
}

inline ArgumentParser::ParseResult parse(ArgumentParser& p, int argc, char const** argv) {
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        std::string name, inlineValue;
        bool hasInline = false;
        if (a == "-h" || a == "--help") { _printHelp(p, std::cout); return ArgumentParser::PARSE_HELP; }
        if (a == "--version") { std::cout << p.appName << " version: " << p.version << "\nLast update " << p.date << std::endl; return ArgumentParser::PARSE_VERSION; }
        if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            size_t eq = a.find('=');
            name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            if (eq != std::string::npos) { hasInline = true; inlineValue = a.substr(eq + 1); }
        } else if (a.size() == 2 && a[0] == '-') {
            name = a.substr(1);
        } else {
            std::cerr << p.appName << ": Too many arguments!" << std::endl;
            return ArgumentParser::PARSE_ERROR;
        }
        ArgParseOption* o = p.find(name);
        if (!o) { std::cerr << p.appName << ": Unknown option: " << a << std::endl; return ArgumentParser::PARSE_ERROR; }
        o->set = true;
        if (o->isFlag) continue;
        std::string val;
        if (hasInline) val = inlineValue;
        else if (i + 1 < argc) val = argv[++i];
        else { std::cerr << p.appName << ": option requires an argument -- " << name << std::endl; return ArgumentParser::PARSE_ERROR; }
        if (o->type == ArgParseArgument::INTEGER) {
            char* endp = nullptr;
            long x = strtol(val.c_str(), &endp, 10);
            if (val.empty() || *endp != 0) { std::cerr << p.appName << ": the given value '" << val << "' cannot be casted to integer" << std::endl; return ArgumentParser::PARSE_ERROR; }
            if (o->hasMin && x < o->minValue) { std::cerr << p.appName << ": the given value '" << val << "' is not in the interval [" << o->minValue << ":+inf]" << std::endl; return ArgumentParser::PARSE_ERROR; }
        }
        if (!o->validValues.empty()) {
            bool okv = false;
            for (const std::string& vv : o->validValues) {
                if (o->type == ArgParseArgument::INPUTFILE || o->type == ArgParseArgument::OUTPUTFILE) okv = okv || _compareExtension(val.c_str(), vv.c_str());
                else okv = okv || (val == vv);
            }
            if (!okv) { std::cerr << p.appName << ": the given value '" << val << "' is not in the list of allowed values" << std::endl; return ArgumentParser::PARSE_ERROR; }
        }
        if (!o->isList) o->values.clear();
        o->values.push_back(val);
    }
    for (ArgParseOption& o : p.options)
        if (o.required && o.values.empty()) { std::cerr << p.appName << ": Missing value for option: -" << o.shortName << std::endl; return ArgumentParser::PARSE_ERROR; }
    return ArgumentParser::PARSE_OK;
}

}  // namespace seqan
