// kcf_host.cpp — see kcf_host.hpp.  Everything here is host-side control flow of `kcftools getVariations`; the
// per-k-mer work is one call into libkcfgpu.so (include/kcf_b200.h).
#include "kcf_host.hpp"
#include "kcf_tools.hpp"

#include <algorithm>
#include <charconv>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fcntl.h>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>

namespace kcfh {

// ================================================================================================ Logger
namespace Logger {
static void log(const char *level, std::string cls, const std::string &msg)
{
    if (cls.size() < 20) cls.append(20 - cls.size(), ' '); // Logger.java:16-18
    using namespace std::chrono;
    const auto now = system_clock::now();
    const std::time_t t = system_clock::to_time_t(now);
    const int ms = (int)(duration_cast<milliseconds>(now.time_since_epoch()).count() % 1000);
    std::tm tm{};
    localtime_r(&t, &tm);
    char buf[64];
    std::snprintf(buf, sizeof buf, "%04d-%02d-%02d %02d:%02d:%02d:%03d", tm.tm_year + 1900, tm.tm_mon + 1, tm.tm_mday, tm.tm_hour,
                  tm.tm_min, tm.tm_sec, ms);
    std::printf("%s - %s - %s - %s\n", buf, level, cls.c_str(), msg.c_str());
    std::fflush(stdout);
}
void info(const std::string &cls, const std::string &msg) { log("INFO    ", cls, msg); }
void warning(const std::string &cls, const std::string &msg) { log("WARNING ", cls, msg); }
void error(const std::string &cls, const std::string &msg)
{
    log("ERROR   ", cls, msg);
    throw FatalError(msg); // System.exit(1)
}
} // namespace Logger

// ================================================================================================ Java text
namespace {
// shortest decimal digits that read back as the same binary value: digits d1 d2 ... dn and exponent e with
// value = 0.d1d2...dn * 10^e
struct Digits {
    bool neg = false;
    std::string d;
    int e = 0;
};

template <typename T> Digits shortest_digits(T v, int max_prec)
{
    Digits r;
    r.neg = std::signbit(v);
    if (v == 0) {
        r.d = "0";
        r.e = 1;
        return r;
    }
    // std::to_chars without a precision = the shortest digits that read back as the same value, the closest such
    // decimal when there are several (what Double.toString / Float.toString start from); scientific form d.ddde+XX
    (void)max_prec;
    char buf[64];
    const auto res = std::to_chars(buf, buf + sizeof buf, (T)std::fabs(v), std::chars_format::scientific);
    std::string s(buf, res.ptr);
    const size_t epos = s.find('e');
    std::string mant = s.substr(0, epos);
    const int ex = std::atoi(s.c_str() + epos + 1);
    mant.erase(std::remove(mant.begin(), mant.end(), '.'), mant.end());
    while (mant.size() > 1 && mant.back() == '0') mant.pop_back();
    r.d = mant;
    r.e = ex + 1;
    return r;
}

std::string java_fp_to_string(const Digits &g, bool is_zero, bool is_inf, bool is_nan)
{
    if (is_nan) return "NaN";
    if (is_inf) return g.neg ? "-Infinity" : "Infinity";
    std::string out = g.neg ? "-" : "";
    if (is_zero) return out + "0.0";
    const int n = (int)g.d.size();
    if (g.e > -3 && g.e <= 7) { // 10^-3 <= |v| < 10^7: plain notation, at least one digit after the point
        if (g.e <= 0) {
            out += "0.";
            out.append((size_t)(-g.e), '0');
            out += g.d;
        } else if (g.e >= n) {
            out += g.d;
            out.append((size_t)(g.e - n), '0');
            out += ".0";
        } else {
            out += g.d.substr(0, (size_t)g.e) + "." + g.d.substr((size_t)g.e);
        }
    } else { // computerised scientific notation: d.dddE[-]n
        out += g.d.substr(0, 1) + "." + (n > 1 ? g.d.substr(1) : std::string("0")) + "E" + std::to_string(g.e - 1);
    }
    return out;
}
} // namespace

std::string java_double_to_string(double v)
{
    return java_fp_to_string(shortest_digits<double>(v, 17), v == 0, std::isinf(v), std::isnan(v));
}

std::string java_float_to_string(float v)
{
    return java_fp_to_string(shortest_digits<float>(v, 9), v == 0, std::isinf(v), std::isnan(v));
}

// String.format("%.2f", v): java.util.Formatter rounds the SHORTEST decimal representation half-up (so 0.125 ->
// "0.13" and 2.675 -> "2.68", unlike printf which rounds the exact binary value half-even).
std::string java_format_2f(double v)
{
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-Infinity" : "Infinity";
    const Digits g = shortest_digits<double>(v, 17);
    // digit string of the integer part and the fraction
    std::string ip, fp;
    if (g.e <= 0) {
        ip = "0";
        fp = std::string((size_t)(-g.e), '0') + g.d;
    } else if (g.e >= (int)g.d.size()) {
        ip = g.d + std::string((size_t)(g.e - (int)g.d.size()), '0');
    } else {
        ip = g.d.substr(0, (size_t)g.e);
        fp = g.d.substr((size_t)g.e);
    }
    bool up = fp.size() > 2 && fp[2] >= '5';
    fp.resize(2, '0');
    if (up) { // propagate the carry through "ip.fp"
        std::string all = ip + fp;
        int i = (int)all.size() - 1;
        while (i >= 0) {
            if (all[i] == '9') {
                all[i] = '0';
                --i;
            } else {
                all[i]++;
                break;
            }
        }
        if (i < 0) all.insert(all.begin(), '1');
        ip = all.substr(0, all.size() - 2);
        fp = all.substr(all.size() - 2);
    }
    const bool zero = ip.find_first_not_of('0') == std::string::npos && fp == "00";
    return std::string(g.neg && !zero ? "-" : (g.neg ? "-" : "")) + ip + "." + fp;
}

// String.hashCode(): over the UTF-16 code units of the text.  The files are UTF-8: decode code points and feed Java's units
// (a surrogate pair above U+FFFF); malformed bytes go in as they are, as ISO-8859-1 would read them.
int32_t java_string_hash(const std::string &s)
{
    uint32_t h = 0;
    const size_t n = s.size();
    for (size_t i = 0; i < n;) {
        const unsigned char c = (unsigned char)s[i];
        uint32_t cp = c;
        size_t len = 1;
        if (c >= 0xC2 && c <= 0xDF && i + 1 < n && ((unsigned char)s[i + 1] & 0xC0) == 0x80) {
            cp = ((uint32_t)(c & 0x1F) << 6) | ((unsigned char)s[i + 1] & 0x3F);
            len = 2;
        } else if (c >= 0xE0 && c <= 0xEF && i + 2 < n && ((unsigned char)s[i + 1] & 0xC0) == 0x80 && ((unsigned char)s[i + 2] & 0xC0) == 0x80) {
            cp = ((uint32_t)(c & 0x0F) << 12) | (((uint32_t)(unsigned char)s[i + 1] & 0x3F) << 6) | ((unsigned char)s[i + 2] & 0x3F);
            len = 3;
        } else if (c >= 0xF0 && c <= 0xF4 && i + 3 < n && ((unsigned char)s[i + 1] & 0xC0) == 0x80 && ((unsigned char)s[i + 2] & 0xC0) == 0x80 &&
                   ((unsigned char)s[i + 3] & 0xC0) == 0x80) {
            cp = ((uint32_t)(c & 0x07) << 18) | (((uint32_t)(unsigned char)s[i + 1] & 0x3F) << 12) | (((uint32_t)(unsigned char)s[i + 2] & 0x3F) << 6) |
                 ((unsigned char)s[i + 3] & 0x3F);
            len = 4;
        }
        if (cp >= 0x10000) {
            const uint32_t v = cp - 0x10000;
            h = 31u * h + (0xD800u + (v >> 10));
            h = 31u * h + (0xDC00u + (v & 0x3FFu));
        } else {
            h = 31u * h + cp;
        }
        i += len;
    }
    return (int32_t)h;
}

// ================================================================================================ small helpers
namespace {
// BufferedReader.readLine over a memory image: lines end at \n, \r or \r\n
struct LineReader {
    const char *p, *end;
    explicit LineReader(const std::string &s) : p(s.data()), end(s.data() + s.size()) {}
    LineReader(const char *b, const char *e) : p(b), end(e) {}
    bool next(std::string &line)
    {
        if (p >= end) return false;
        const char *q = p;
        while (q < end && *q != '\n' && *q != '\r') ++q;
        line.assign(p, q);
        if (q < end) {
            if (*q == '\r' && q + 1 < end && q[1] == '\n') ++q;
            ++q;
        }
        p = q;
        return true;
    }
};

} // namespace

bool read_file(const std::string &path, std::string &out)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}

// String.split(regex of one literal char): trailing empty strings are removed, leading / inner ones kept
std::vector<std::string> java_split(const std::string &s, char sep)
{
    std::vector<std::string> out;
    size_t a = 0;
    for (;;) {
        const size_t b = s.find(sep, a);
        if (b == std::string::npos) {
            out.push_back(s.substr(a));
            break;
        }
        out.push_back(s.substr(a, b - a));
        a = b + 1;
    }
    while (!out.empty() && out.back().empty()) out.pop_back();
    if (out.empty() && s.empty()) out.push_back(""); // "".split(x) is [""]
    return out;
}

std::string java_trim(const std::string &s)
{
    size_t a = 0, b = s.size();
    while (a < b && (unsigned char)s[a] <= ' ') ++a;
    while (b > a && (unsigned char)s[b - 1] <= ' ') --b;
    return s.substr(a, b - a);
}

int java_parse_int(const std::string &s, const std::string &what)
{
    // Integer.parseInt: optional sign, decimal digits only, int range
    bool ok = !s.empty();
    size_t i = (s[0] == '-' || s[0] == '+') ? 1 : 0;
    if (i >= s.size()) ok = false;
    long long v = 0;
    for (; ok && i < s.size(); ++i) {
        if (s[i] < '0' || s[i] > '9') ok = false;
        else {
            v = v * 10 + (s[i] - '0');
            if (v > (1LL << 32)) ok = false;
        }
    }
    if (ok && s[0] == '-') v = -v;
    if (!ok || v > 2147483647LL || v < -2147483648LL)
        throw FatalError("java.lang.NumberFormatException: For input string: \"" + s + "\" (" + what + ")");
    return (int)v;
}

std::vector<std::string> java_lines(const std::string &text)
{
    std::vector<std::string> out;
    LineReader rd(text);
    std::string line;
    while (rd.next(line)) out.push_back(line);
    return out;
}

// ================================================================================================ FastaIndex
static const char *const FAI_CLASS = "FastaIndex";

void FastaIndex::generateIndexFile(const std::string &fasta, const std::string &fai)
{
    // FastaIndex.java:239-299.  Offsets count line.length() + 1 per line like the reference on Linux.
    int fd = ::open(fasta.c_str(), O_RDONLY);
    if (fd < 0) throw FatalError("java.io.FileNotFoundException: " + fasta + " (No such file or directory)");
    struct stat st;
    fstat(fd, &st);
    const size_t n = (size_t)st.st_size;
    const char *map = n ? (const char *)mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0) : "";
    if (n && map == (const char *)MAP_FAILED) {
        ::close(fd);
        throw FatalError("cannot map " + fasta);
    }
    if (n >= 2 && (unsigned char)map[0] == 0x1f && (unsigned char)map[1] == 0x8b) { // HelperFunctions.isCompressed
        munmap((void *)map, n);
        ::close(fd);
        Logger::error(FAI_CLASS, "Fasta file is compressed. Please decompress before indexing: " + fasta);
    }
    static bool valid[256];
    static bool init = false;
    if (!init) {
        for (const char *c = "ACGTYRWSMKHBVDNacgtyrwsmkhbvdn"; *c; ++c) valid[(unsigned char)*c] = true;
        init = true;
    }
    std::ostringstream out;
    LineReader rd(map, map + n);
    std::string line, currentName;
    long long offset = 0, sequenceStartOffset = 0;
    int lineNumber = 0, lineBases = 0, lineWidth = 0;
    long long seqLength = 0;
    bool inSequence = false;
    std::unordered_map<std::string, char> seen;
    try {
        while (rd.next(line)) {
            ++lineNumber;
            if (offset == 0 && (line.empty() || line[0] != '>')) Logger::error(FAI_CLASS, "Invalid fasta file: " + fasta);
            if (!line.empty() && line[0] == '>') {
                if (inSequence)
                    out << currentName << "\t" << seqLength << "\t" << sequenceStartOffset << "\t" << lineBases << "\t" << lineWidth << "\n";
                currentName = java_split(line.substr(1), ' ')[0];
                offset += (long long)line.size() + 1;
                sequenceStartOffset = offset;
                inSequence = true;
                seqLength = 0;
                if (seen.count(currentName))
                    Logger::error(FAI_CLASS, "Duplicate sequence name in fasta file: " + currentName + " at line " + std::to_string(lineNumber));
                seen[currentName] = 1;
            } else {
                for (unsigned char c : line)
                    if (!valid[c])
                        Logger::error(FAI_CLASS, std::string("Invalid character '") + (char)c + "' in fasta file: " + fasta + " at line " +
                                                     std::to_string(lineNumber));
                if (seqLength == 0) {
                    lineBases = (int)line.size();
                    lineWidth = (int)line.size() + 1;
                }
                seqLength += (long long)line.size();
                offset += (long long)line.size() + 1;
            }
        }
    } catch (...) {
        if (n) munmap((void *)map, n);
        ::close(fd);
        throw;
    }
    if (inSequence) out << currentName << "\t" << seqLength << "\t" << sequenceStartOffset << "\t" << lineBases << "\t" << lineWidth << "\n";
    if (n) munmap((void *)map, n);
    ::close(fd);
    std::ofstream f(fai, std::ios::binary);
    if (!f) throw FatalError("java.io.FileNotFoundException: " + fai + " (Permission denied)");
    f << out.str();
}

FastaIndex::FastaIndex(const std::string &fastaPath)
{
    const std::string fai = fastaPath + ".faidx";
    struct stat sf, si;
    if (stat(fastaPath.c_str(), &sf) != 0) throw FatalError("java.io.FileNotFoundException: " + fastaPath + " (No such file or directory)");
    const bool have = stat(fai.c_str(), &si) == 0;
    auto ms = [](const struct stat &s) { return (long long)s.st_mtim.tv_sec * 1000 + s.st_mtim.tv_nsec / 1000000; };
    if (!have || ms(si) < ms(sf)) { // HelperFunctions.isOlder
        Logger::info(FAI_CLASS, "Generating/Updating index file: " + fai);
        generateIndexFile(fastaPath, fai);
    } else {
        Logger::info(FAI_CLASS, "Using existing index file: " + fai);
    }
    std::string text;
    if (!read_file(fai, text)) throw FatalError("java.io.FileNotFoundException: " + fai);
    LineReader rd(text);
    std::string line;
    int seqId = 0;
    while (rd.next(line)) { // FastaIndex.java:82-103
        const std::vector<std::string> f = java_split(line, '\t');
        if (f.size() < 5) throw FatalError("java.lang.ArrayIndexOutOfBoundsException: malformed .faidx line: " + line);
        FastaIndexEntry e;
        e.seqId = seqId++;
        e.name = f[0];
        e.length = java_parse_int(f[1], ".faidx length");
        e.offset = std::atoll(f[2].c_str());
        e.lineBases = java_parse_int(f[3], ".faidx lineBases");
        e.lineWidth = java_parse_int(f[4], ".faidx lineWidth");
        if (byName_.count(e.name)) Logger::error(FAI_CLASS, "Duplicate sequence name in index: " + e.name);
        byName_[e.name] = (int)entries_.size();
        entries_.push_back(e); // file order == seqId order == the reference's sortByValue order
    }
    fd_ = ::open(fastaPath.c_str(), O_RDONLY);
    if (fd_ < 0) Logger::error(FAI_CLASS, "Error memory-mapping fasta file: " + fastaPath);
    mapLen_ = (uint64_t)sf.st_size;
    if (mapLen_) {
        void *m = mmap(nullptr, mapLen_, PROT_READ, MAP_PRIVATE, fd_, 0);
        if (m == MAP_FAILED) Logger::error(FAI_CLASS, "Error memory-mapping fasta file: " + fastaPath);
        map_ = (const uint8_t *)m;
    }
}

FastaIndex::~FastaIndex()
{
    if (map_) munmap((void *)map_, mapLen_);
    if (fd_ >= 0) ::close(fd_);
}

const FastaIndexEntry *FastaIndex::getEntry(const std::string &name) const
{
    auto it = byName_.find(name);
    return it == byName_.end() ? nullptr : &entries_[it->second];
}

int FastaIndex::getSequenceLength(const std::string &name) const
{
    const FastaIndexEntry *e = getEntry(name);
    if (!e) Logger::error(FAI_CLASS, "Sequence not found in index: " + name);
    return e->length;
}

const uint8_t *FastaIndex::seqBytes(int seqId, uint64_t *n) const
{
    // FastaIndex.java:54-68: from this entry's offset to the next entry's offset (or the end of the file)
    const FastaIndexEntry &e = entries_[seqId];
    const int64_t endOff = (size_t)seqId + 1 < entries_.size() ? entries_[seqId + 1].offset : (int64_t)mapLen_;
    int64_t len = endOff - e.offset;
    if (len < 0 || e.offset < 0 || (uint64_t)e.offset > mapLen_) len = 0;
    if ((uint64_t)(e.offset + len) > mapLen_) len = (int64_t)mapLen_ - e.offset;
    *n = (uint64_t)len;
    return map_ + e.offset;
}

// ================================================================================================ GTF
static const char *const GTF_CLASS = "GTF";

int GTF::addVertex(const std::string &v)
{
    auto it = vertices_.find(v);
    if (it != vertices_.end()) return it->second;
    const int id = (int)children_.size();
    vertices_[v] = id;
    children_.emplace_back();
    childSet_.emplace_back();
    return id;
}

void GTF::addEdge(const std::string &from, const std::string &to)
{
    const int a = addVertex(from);
    addVertex(to);
    if (childSet_[a].count(to)) return; // DefaultDirectedGraph holds no parallel edges
    childSet_[a][to] = 1;
    children_[a].push_back(to);
}

std::vector<std::string> GTF::getChildren(const std::string &parent) const
{
    std::vector<std::string> out;
    auto it = vertices_.find(parent);
    if (it == vertices_.end()) return out;
    for (const std::string &c : children_[it->second])
        if (c != parent) out.push_back(c);
    return out;
}

GTF::GTF(const std::string &path)
{
    Logger::info(GTF_CLASS, "Parsing GTF file at: " + path);
    std::string text;
    if (!read_file(path, text)) Logger::error(GTF_CLASS, "Error parsing GTF file: " + path + " (No such file or directory)");
    LineReader rd(text);
    std::string line;
    std::unordered_map<std::string, int> exonCounts;
    auto attr_of = [](const std::string &field) { // GTF.java:156-163
        std::unordered_map<std::string, std::string> m;
        for (const std::string &attr : java_split(field, ';')) {
            std::string t = java_trim(attr);
            t.erase(std::remove(t.begin(), t.end(), '"'), t.end());
            const std::vector<std::string> pair = java_split(t, ' ');
            if (pair.size() == 2) m[pair[0]] = pair[1];
        }
        return m;
    };
    while (rd.next(line)) {
        if ((!line.empty() && line[0] == '#') || java_trim(line).empty()) continue;
        const std::vector<std::string> f = java_split(line, '\t');
        if (f.size() < 9) Logger::error(GTF_CLASS, "Malformed line: " + line);
        auto attributes = attr_of(f[8]);
        const std::string &type = f[2], &chromID = f[0];
        std::string featureID, parentID;
        bool hasFeature = false, hasParent = false;
        addVertex(chromID);
        auto get = [&](const char *k, std::string &dst) {
            auto it = attributes.find(k);
            if (it == attributes.end()) return false;
            dst = it->second;
            return true;
        };
        if (type == "gene" || type == "pseudogene") {
            hasFeature = get("gene_id", featureID);
            parentID = chromID;
            hasParent = true;
        } else if (type == "transcript" || type == "mRNA" || type == "RNA" || type == "lnc_RNA" || type == "rRNA" || type == "tRNA" ||
                   type == "snRNA" || type == "snoRNA") {
            hasFeature = get("transcript_id", featureID);
            hasParent = get("gene_id", parentID);
            if (!hasFeature) throw FatalError("java.lang.NullPointerException: transcript line without transcript_id: " + line);
            if (hasParent && featureID == parentID)
                Logger::error(GTF_CLASS, "Transcript ID is the same as Gene ID: " + featureID + ". Fix the GTF file using AGAT.");
            if (!hasParent) throw FatalError("java.lang.NullPointerException: transcript line without gene_id: " + line);
            const int s = java_parse_int(f[3], "GTF start"), e = java_parse_int(f[4], "GTF end");
            if (!containsVertex(parentID)) {
                addVertex(parentID);
                addEdge(chromID, parentID);
                featureMap_[parentID] = Feature{chromID, s, e, f[6].empty() ? '\0' : f[6][0], "gene", parentID};
            }
            auto g = featureMap_.find(parentID);
            if (g != featureMap_.end()) {
                if (s < g->second.start) g->second.start = s;
                if (e > g->second.end) g->second.end = e;
            }
        } else if (type == "exon") {
            hasParent = get("transcript_id", parentID);
            const std::string key = hasParent ? parentID : std::string("null");
            const int count = ++exonCounts[key];
            featureID = key + "-e-" + std::to_string(count);
            hasFeature = true;
        } else {
            continue;
        }
        if (f[6].empty()) throw FatalError("java.lang.StringIndexOutOfBoundsException: empty strand field: " + line);
        if (!hasFeature) throw FatalError("java.lang.NullPointerException: feature without an id: " + line);
        featureMap_[featureID] = Feature{f[0], java_parse_int(f[3], "GTF start"), java_parse_int(f[4], "GTF end"), f[6][0], type, featureID};
        addVertex(featureID);
        if (hasParent) addEdge(parentID, featureID);
    }
}

Loci GTF::getLoci(const std::string &featureID) const
{
    auto it = featureMap_.find(featureID);
    if (it == featureMap_.end()) Logger::error(GTF_CLASS, "Feature ID not found: " + featureID);
    const Feature &f = it->second;
    return Loci{f.chromosome, f.start, f.end, std::string(1, f.strand)};
}

static bool loci_equal(const Loci &a, const Loci &b)
{
    return a.chromosome == b.chromosome && a.start == b.start && a.end == b.end && a.strand == b.strand;
}

// Iteration order of a java.util.HashSet<Loci> (a HashMap underneath): buckets in index order, insertion order
// inside a bucket; the table doubles from 16 whenever the size exceeds 3/4 of it.  Loci.hashCode: GTF.java:428-435.
std::vector<Loci> GTF::javaHashSetOrder(const std::vector<Loci> &ins)
{
    std::vector<Loci> uniq;
    for (const Loci &l : ins) {
        bool dup = false;
        for (const Loci &u : uniq)
            if (loci_equal(u, l)) {
                dup = true;
                break;
            }
        if (!dup) uniq.push_back(l);
    }
    size_t cap = 16;
    while (uniq.size() > cap * 3 / 4) cap *= 2;
    std::vector<std::pair<uint32_t, size_t>> key(uniq.size());
    for (size_t i = 0; i < uniq.size(); ++i) {
        uint32_t h = (uint32_t)java_string_hash(uniq[i].chromosome);
        h = 31u * h + (uint32_t)uniq[i].start;
        h = 31u * h + (uint32_t)uniq[i].end;
        h = 31u * h + (uint32_t)java_string_hash(uniq[i].strand);
        h ^= h >> 16; // HashMap.hash
        key[i] = {h & (uint32_t)(cap - 1), i};
    }
    std::sort(key.begin(), key.end());
    std::vector<Loci> out;
    for (auto &k : key) out.push_back(uniq[k.second]);
    return out;
}

static bool loci_less(const Loci &a, const Loci &b) // Loci.compareTo, GTF.java:407-413
{
    if (a.chromosome == b.chromosome) return a.start < b.start;
    return a.chromosome < b.chromosome;
}

std::vector<Loci> GTF::mergeOverlappingLoci(std::vector<Loci> sorted)
{
    std::stable_sort(sorted.begin(), sorted.end(), loci_less);
    std::vector<Loci> merged;
    for (const Loci &cur : sorted) {
        if (merged.empty()) {
            merged.push_back(cur);
            continue;
        }
        Loci &last = merged.back();
        const bool overlaps = last.chromosome == cur.chromosome && last.strand == cur.strand && last.start <= cur.end && cur.start <= last.end;
        if (overlaps) {
            last.start = std::min(last.start, cur.start);
            last.end = std::max(last.end, cur.end);
        } else {
            merged.push_back(cur);
        }
    }
    return merged;
}

std::vector<Loci> GTF::mergedLoci(const std::string &featureID, bool isGene) const
{
    // GTF.java:223-239
    std::vector<Loci> ins;
    if (!containsVertex(featureID)) return ins;
    const std::vector<std::string> targets = isGene ? getTranscripts(featureID) : getExons(featureID);
    for (const std::string &t : targets) {
        const std::vector<std::string> exons = isGene ? getExons(t) : std::vector<std::string>{t};
        for (const std::string &ex : exons) {
            auto it = featureMap_.find(ex);
            if (it != featureMap_.end()) ins.push_back(Loci{it->second.chromosome, it->second.start, it->second.end, std::string(1, it->second.strand)});
        }
    }
    if (ins.empty()) return ins;
    std::vector<Loci> merged = mergeOverlappingLoci(javaHashSetOrder(ins));
    std::stable_sort(merged.begin(), merged.end(), loci_less);
    return merged;
}

// ================================================================================================ GetVariants
static const char *const GV_CLASS = "GetVariants";

void validateCMD(const GetVariantsOptions &o)
{
    if (o.featureType == "window") {
        if (o.windowSize <= 0) Logger::error(GV_CLASS, "Window size is required for window model");
        if (o.hasGtf && !o.gtfFile.empty()) Logger::error(GV_CLASS, "GTF file is not valid for window model");
    } else if (o.featureType == "gene" || o.featureType == "transcript") {
        if (!o.hasGtf || o.gtfFile.empty()) Logger::error(GV_CLASS, "GTF file is required for targeted model");
        if (o.windowSize > 0) Logger::error(GV_CLASS, "Window size is not valid for targeted model");
    } else {
        Logger::error(GV_CLASS, "Invalid model type: " + o.featureType + ". Supported models are 'window' or 'gene' or 'transcript'");
    }
    if (o.nThreads <= 0) Logger::error(GV_CLASS, "Number of threads should be greater than 0");
    if (o.minKmerCount < 1) Logger::error(GV_CLASS, "Minimum kmer count should be at least 1");
}

std::string cleanSampleName(const std::string &s)
{
    std::string out = s;
    for (char &c : out)
        if (std::strchr("\\/:*?\"<>|", c) && c != '\0') c = '_';
    if (out != s) Logger::warning(GV_CLASS, "Sample name contains invalid characters, changed to: " + out);
    return out;
}

static std::vector<kcf_segment_t> segments_of_loci(const std::vector<Loci> &merged, const FastaIndex &index)
{
    std::vector<kcf_segment_t> segs;
    for (const Loci &l : merged) { // fastaIndex.getSequence(chrom, start - 1, length), GTF.java:240-244
        const FastaIndexEntry *e = index.getEntry(l.chromosome);
        if (!e) Logger::error(FAI_CLASS, "Sequence not found in index: " + l.chromosome);
        const long long start = (long long)l.start - 1, end = start + l.getLength();
        if (start < 0 || end > e->length || start >= end)
            Logger::error(FAI_CLASS, "Invalid range: " + std::to_string(start) + "-" + std::to_string(end) + " for sequence: " + l.chromosome);
        segs.push_back(kcf_segment_t{e->seqId, (int32_t)start, (int32_t)(end - start)});
    }
    return segs;
}

std::vector<Window> getWindows(const GetVariantsOptions &o, const FastaIndex &index, const GTF *gtf, const std::string &sequenceName,
                               int kmerSize)
{
    std::vector<Window> windows;
    const int sequenceLength = index.getSequenceLength(sequenceName);
    const int seqId = index.getEntry(sequenceName)->seqId;
    auto fixed = [&](int start, int end) {
        Window w;
        w.windowId = sequenceName + "_" + std::to_string(start);
        w.sequenceName = sequenceName;
        w.start = start;
        w.end = end;
        w.segments.push_back(kcf_segment_t{seqId, start, end - start});
        windows.push_back(std::move(w));
    };
    if (o.featureType == "window") {
        if (o.stepSize > 0) { // sliding windows
            long long lastPos = 0;
            while (lastPos < sequenceLength) {
                const int start = (int)lastPos;
                const int end = (int)std::min<long long>((long long)start + o.windowSize, sequenceLength);
                if (end - start >= kmerSize) fixed(start, end);
                lastPos += o.stepSize;
            }
        } else { // tiling windows that share k-1 bases
            if (o.windowSize <= kmerSize - 1)
                Logger::error(GV_CLASS, "Window size " + std::to_string(o.windowSize) + " does not exceed k-1 = " + std::to_string(kmerSize - 1) +
                                            ": the reference's tiling loop (GetVariants.java:309-319) never terminates");
            int lastEnd = 0;
            while (lastEnd < sequenceLength) {
                const int start = std::max(0, lastEnd - kmerSize + 1);
                const int end = (int)std::min<long long>((long long)start + o.windowSize, sequenceLength);
                if (end - start >= kmerSize) fixed(start, end);
                lastEnd = end;
            }
        }
    } else if (o.featureType == "gene" || o.featureType == "transcript") {
        const bool isGene = o.featureType == "gene";
        const std::vector<std::string> genes = gtf->getGenes(sequenceName);
        auto feature_window = [&](const std::string &id) {
            const Loci l = gtf->getLoci(id);
            Window w;
            w.windowId = id;
            w.sequenceName = l.chromosome;
            w.start = l.start;
            w.end = l.end;
            const std::vector<Loci> merged = gtf->mergedLoci(id, isGene);
            if (merged.empty()) w.noFasta = true; // GTF.getFasta returns null
            else w.segments = segments_of_loci(merged, index);
            windows.push_back(std::move(w));
        };
        if (isGene) {
            for (const std::string &g : genes) feature_window(g);
        } else {
            if (genes.empty()) {
                Logger::warning(GV_CLASS, "No genes found in GTF file for sequence: " + sequenceName);
                return windows;
            }
            for (const std::string &g : genes) {
                const std::vector<std::string> txs = gtf->getTranscripts(g);
                if (txs.empty())
                    Logger::error(GV_CLASS, "No transcripts found for gene: " + g + " in GTF file for sequence: " + sequenceName);
                for (const std::string &t : txs) feature_window(t);
            }
        }
    } else {
        Logger::error(GV_CLASS, "Invalid model type: " + o.featureType + ". Supported models are 'window' or 'gene' or 'transcript'");
    }
    return windows;
}

double computeScore(const kcf_result_t &r, const double w[3])
{
    // Data.java:95-107; this translation unit is compiled with -ffp-contract=off (Java has no fused multiply-add here)
    if (r.obs == 0 || r.total_kmers == 0 || r.eff_len == 0) return 0;
    if (w[0] + w[1] + w[2] != 1.0) Logger::error("Data", "Weights should sum to 1.0");
    return ((w[2] * ((double)r.obs / r.total_kmers)) + (w[0] * (1.0 - ((double)r.inner / r.eff_len))) +
            (w[1] * (1.0 - ((double)(r.left + r.right) / r.eff_len)))) *
           100.0;
}

std::string today()
{
    std::time_t t = std::time(nullptr);
    std::tm tm{};
    localtime_r(&t, &tm);
    char buf[32];
    std::snprintf(buf, sizeof buf, "%04d-%02d-%02d", tm.tm_year + 1900, tm.tm_mon + 1, tm.tm_mday);
    return buf;
}

static const char *const KCF_INFO_LINES[] = {
    "<ID=EFFLEN,Type=Integer,Description=\"Effective length of the window\">",
    "<ID=IS,Type=Float,Description=\"Minimum score for the window\">",
    "<ID=XS,Type=Float,Description=\"Maximum score for the window\">",
    "<ID=MS,Type=Float,Description=\"Mean score for the window\">",
    "<ID=IO,Type=Integer,Description=\"Minimum observed kmers in the window\">",
    "<ID=XO,Type=Integer,Description=\"Maximum observed kmers in the window\">",
    "<ID=MO,Type=Integer,Description=\"Mean observed kmers in the window\">",
    "<ID=IV,Type=Integer,Description=\"Minimum variations in the window\">",
    "<ID=XV,Type=Integer,Description=\"Maximum variations in the window\">",
    "<ID=MV,Type=Integer,Description=\"Mean variations in the window\">"};
static const char *const KCF_FORMAT_LINES[] = {
    "<ID=IB,Type=Integer,Description=\"IBS number\">",
    "<ID=VA,Type=Integer,Description=\"Variations\">",
    "<ID=OB,Type=Integer,Description=\"Observed kmers\">",
    "<ID=ID,Type=Integer,Description=\"Inner Distance\">",
    "<ID=LD,Type=Integer,Description=\"Kmer Variation Distance at the leftTail\">",
    "<ID=RD,Type=Integer,Description=\"Kmer Variation Distance at the rightTail\">",
    "<ID=KD,Type=Float,Description=\"Mean Kmer Depth\">",
    "<ID=SC,Type=Float,Description=\"Score\">"};

#ifndef KCF_FORMAT_VERSION
#define KCF_FORMAT_VERSION "0.4.0" /* the reference's pom.xml version, filtered into version.properties at build time */
#endif

std::string kcfStaticHeaderLines() // ##INFO and ##FORMAT lines (Configs.java:14-37)
{
    std::string s;
    for (const char *l : KCF_INFO_LINES) s += std::string("##INFO=") + l + "\n";
    for (const char *l : KCF_FORMAT_LINES) s += std::string("##FORMAT=") + l + "\n";
    return s;
}
const char *kcfFormatVersion() { return KCF_FORMAT_VERSION; }

std::string kcfHeaderText(const GetVariantsOptions &o, const std::string &sample, const FastaIndex &index, int kmerSize,
                          int totalWindows, const std::string &date)
{
    std::ostringstream sb;
    sb << "##format=KCF" << KCF_FORMAT_VERSION << "\n";
    sb << "##date=" << date << "\n";
    sb << "##source=kcftools\n";
    sb << "##reference=" << o.refFasta << "\n";
    for (const FastaIndexEntry &e : index.entries()) sb << "##contig=<ID=" << e.name << ",length=" << e.length << ">\n";
    for (const char *l : KCF_INFO_LINES) sb << "##INFO=" << l << "\n";
    for (const char *l : KCF_FORMAT_LINES) sb << "##FORMAT=" << l << "\n";
    auto param = [&](const char *k, const std::string &v) { sb << "##PARAM=<ID=" << k << ",value=" << v << ">\n"; };
    param("window", std::to_string(o.windowSize)); // KCFHeader.java:205-215, 139-191: params[0..7] in this order
    param("step", std::to_string(o.stepSize));
    param("kmer", std::to_string(kmerSize));
    param("IBS", "false");
    param("nwindow", std::to_string(totalWindows));
    param("wti", java_double_to_string(o.innerDistanceWeight));
    param("wtt", java_double_to_string(o.tailDistanceWeight));
    param("wtk", java_double_to_string(o.kmerRatioWeight));
    sb << "##CMD=" << o.commandLine << "\n";
    sb << "#CHROM\tSTART\tEND\tID\tTOTAL_KMERS\tINFO\tFORMAT\t" << sample << "\n";
    return sb.str();
}

std::string kcfRowText(const Window &w, const kcf_result_t &r, const double weights[3])
{
    // one sample per window: min = max = mean (Window.java:177-214).  XS starts at Float.MIN_VALUE, IS at Float.MAX_VALUE.
    const double score = computeScore(r, weights);
    const double minScore = std::min((double)3.4028234663852886e38f, score);
    const double maxScore = std::max((double)1.401298464324817e-45f, score);
    const float meanObs = (float)r.obs, meanVar = (float)r.variations;
    const double kd = r.obs > 0 ? (double)r.kmer_count_sum / r.obs : 0.0; // Data.java:87
    std::ostringstream sb;
    sb << w.sequenceName << "\t" << w.start << "\t" << w.end << "\t" << w.windowId << "\t" << r.total_kmers << "\t";
    const std::string sc2 = java_format_2f(score); // min = max = mean = the score unless it is 0 (XS then starts from Float.MIN_VALUE)
    sb << "EFFLEN=" << r.eff_len << ";IS=" << (minScore == score ? sc2 : java_format_2f(minScore)) << ";XS=" << (maxScore == score ? sc2 : java_format_2f(maxScore))
       << ";MS=" << sc2
       << ";IO=" << r.obs << ";XO=" << r.obs << ";MO=" << java_format_2f((double)meanObs) << ";IV=" << r.variations << ";XV=" << r.variations
       << ";MV=" << java_float_to_string(meanVar);
    sb << "\tGT:VA:OB:ID:LD:RD:KD:SC\t";
    sb << "N:" << r.variations << ":" << r.obs << ":" << r.inner << ":" << r.left << ":" << r.right << ":" << java_format_2f(kd) << ":"
       << sc2;
    return sb.str();
}

namespace {
struct Device {
    kcf_ctx *ctx = nullptr;
    kcf_db *db = nullptr;
    std::vector<kcf_plan *> plans;
    ~Device()
    {
        for (kcf_plan *p : plans) kcf_plan_destroy(p);
        if (db) kcf_db_close(db);
        if (ctx) kcf_shutdown(ctx);
    }
    [[noreturn]] void fail(const char *cls) const { Logger::error(cls, kcf_last_error(ctx)); }
};
} // namespace

// One database against every sequence: queue a plan per sequence.  Rows stay on the device until fetched.
static void screenAllSequences(Device &dev, const std::vector<std::vector<Window>> &perSeq, int kmerSize, int minKmerCount, const double weights[3])
{
    for (kcf_plan *p : dev.plans) kcf_plan_destroy(p);
    dev.plans.clear();
    for (size_t s = 0; s < perSeq.size(); ++s) {
        std::vector<kcf_window_t> wins;
        std::vector<kcf_segment_t> segs;
        for (const Window &w : perSeq[s]) {
            wins.push_back(kcf_window_t{(uint32_t)segs.size(), (uint32_t)w.segments.size()});
            segs.insert(segs.end(), w.segments.begin(), w.segments.end());
        }
        kcf_plan *plan = nullptr;
        if (kcf_plan_create(dev.ctx, kmerSize, wins.data(), wins.size(), segs.data(), segs.size(), &plan) != KCF_OK) dev.fail(FAI_CLASS);
        dev.plans.push_back(plan);
        if (kcf_plan_run(dev.ctx, dev.db, plan, minKmerCount, weights) != KCF_OK) dev.fail(GV_CLASS);
    }
}

// A command uses ONE GPU.  CUDA's start-up cost grows with the number of visible devices (3.5 s on an 8-GPU B200 box
// against 0.4 s with one device visible), so before the first CUDA call the process narrows CUDA_VISIBLE_DEVICES to the
// device it was asked for and addresses it as ordinal 0.  Returns the ordinal to pass to kcf_init.
int restrictToDevice(int device)
{
    static bool done = false;
    if (done) return 0;
    const char *cur = std::getenv("CUDA_VISIBLE_DEVICES");
    std::string pick = std::to_string(device);
    if (cur && *cur) { // already a list: take its device-th entry
        const std::vector<std::string> ids = java_split(cur, ',');
        if (device < 0 || (size_t)device >= ids.size()) return device; // let kcf_init report the bad ordinal
        pick = ids[(size_t)device];
    } else if (device < 0) {
        return device;
    }
    setenv("CUDA_VISIBLE_DEVICES", pick.c_str(), 1);
    done = true;
    return 0;
}

int getVariations(GetVariantsOptions o)
{
    // Extension: `-k a,b,c -s x,y,z` screens several databases against the same reference and writes
    // the COHORT file directly — the text `kcftools cohort` would write from the per-sample files, without those files.
    const std::vector<std::string> prefixes = java_split(o.kmcDBprefix, ',');
    std::vector<std::string> sampleNames = java_split(o.sampleName, ',');
    const bool multi = prefixes.size() > 1;
    const bool multiDevice = multi && o.devices.size() > 1;
    const bool shardedJob = !multi && o.devices.size() > 1; // ONE database, windows cut over the GPUs (GetVariants.java:129-151)
    if (multi && sampleNames.size() != prefixes.size())
        Logger::error(GV_CLASS, "Number of sample names (" + std::to_string(sampleNames.size()) + ") differs from the number of KMC databases (" +
                                    std::to_string(prefixes.size()) + ")");
    if (!multi) sampleNames = {o.sampleName};
    for (std::string &s : sampleNames) s = cleanSampleName(s);
    Device dev;
    // sharded job: the other GPUs open the same database while this thread opens it on the first one
    std::vector<std::unique_ptr<Device>> peers;
    std::vector<std::thread> openers;
    std::vector<std::string> openFailure(o.devices.size());
    if (shardedJob)
        for (size_t g = 1; g < o.devices.size(); ++g) {
            peers.emplace_back(new Device());
            Device *pd = peers.back().get();
            openers.emplace_back([&, g, pd] {
                if (kcf_init(o.devices[g], &pd->ctx) != KCF_OK) openFailure[g] = kcf_last_error(nullptr);
                else if (kcf_db_open(pd->ctx, prefixes[0].c_str(), 0, &pd->db) != KCF_OK) openFailure[g] = kcf_last_error(pd->ctx);
            });
        }
    struct JoinGuard {
        std::vector<std::thread> &t;
        ~JoinGuard()
        {
            for (std::thread &x : t)
                if (x.joinable()) x.join();
        }
    } joinGuard{openers};
    if (kcf_init((multiDevice || shardedJob) ? o.device : restrictToDevice(o.device), &dev.ctx) != KCF_OK) Logger::error("KMC", std::string(kcf_last_error(nullptr)));
    if (kcf_db_open(dev.ctx, prefixes[0].c_str(), 0, &dev.db) != KCF_OK) dev.fail("KMC");
    kcf_db_info_t info;
    kcf_db_info(dev.db, &info);
    const int kmerSize = info.kmer_length;
    {
        char buf[256];
        std::snprintf(buf, sizeof buf, "KMC database resident on device %d: %lld kmers (k=%d) in a %.2f GB table, loaded in %.3f s", o.device,
                      (long long)info.resident_kmers, kmerSize, (double)info.table_bytes / 1e9, info.load_seconds);
        Logger::info("KMC", buf);
    }

    FastaIndex index(o.refFasta);
    std::unique_ptr<GTF> gtf;
    if (o.featureType == "gene" || o.featureType == "transcript") gtf.reset(new GTF(o.gtfFile));

    Logger::info(GV_CLASS, "Generating windows...");
    std::vector<std::vector<Window>> perSeq;
    size_t totalWindows = 0;
    for (const FastaIndexEntry &e : index.entries()) {
        perSeq.push_back(getWindows(o, index, gtf.get(), e.name, kmerSize));
        totalWindows += perSeq.back().size();
    }
    Logger::info(GV_CLASS, "Number of windows: " + std::to_string(totalWindows));
    for (const auto &ws : perSeq)
        for (const Window &w : ws)
            if (w.noFasta) Logger::error(GV_CLASS, "Fasta object is null for window: " + w.windowId); // GetVariants.java:213-216

    // upload the sequences (queued; the copies overlap the 2-bit packing), then one plan per sequence
    auto uploadReference = [&index](Device &dv) {
        for (const FastaIndexEntry &e : index.entries()) {
            uint64_t n = 0;
            const uint8_t *bytes = index.seqBytes(e.seqId, &n);
            int sid = -1;
            if (kcf_ref_add_async(dv.ctx, bytes, n, (uint32_t)e.lineBases, (uint32_t)e.lineWidth, (uint64_t)e.length, &sid) != KCF_OK) dv.fail(FAI_CLASS);
        }
    };
    if (!shardedJob) uploadReference(dev);
    const double weights[3] = {o.innerDistanceWeight, o.tailDistanceWeight, o.kmerRatioWeight}; // getWeights(), :388-390
    // per sequence: the order the reference writes (stable sort by start, GetVariants.java:169-171)
    std::vector<std::vector<size_t>> sorted(perSeq.size());
    for (size_t s = 0; s < perSeq.size(); ++s) {
        sorted[s].resize(perSeq[s].size());
        for (size_t i = 0; i < sorted[s].size(); ++i) sorted[s][i] = i;
        std::stable_sort(sorted[s].begin(), sorted[s].end(), [&](size_t a, size_t b) { return perSeq[s][a].start < perSeq[s][b].start; });
    }
    std::ofstream out(o.outFile, std::ios::binary);
    if (!out) throw FatalError("java.io.FileNotFoundException: " + o.outFile + " (No such file or directory)");

    if (shardedJob) {
        // one job, several GPUs: the library cuts the window list into one contiguous range per device (balanced on bases),
        // every device uploads only the stretches of the reference its range touches and the rows come back in window order
        for (std::thread &t : openers) t.join();
        for (size_t g = 1; g < o.devices.size(); ++g)
            if (!openFailure[g].empty()) Logger::error("KMC", openFailure[g]);
        std::vector<kcf_ctx *> ctxs{dev.ctx};
        std::vector<kcf_db *> dbs{dev.db};
        for (auto &pd : peers) {
            ctxs.push_back(pd->ctx);
            dbs.push_back(pd->db);
        }
        std::vector<kcf_host_seq_t> hseqs;
        for (const FastaIndexEntry &e : index.entries()) {
            uint64_t n = 0;
            const uint8_t *bytes = index.seqBytes(e.seqId, &n);
            hseqs.push_back(kcf_host_seq_t{bytes, n, (uint32_t)e.lineBases, (uint32_t)e.lineWidth, (uint64_t)e.length});
        }
        std::vector<kcf_window_t> wins;
        std::vector<kcf_segment_t> segs;
        for (const auto &ws : perSeq)
            for (const Window &w : ws) {
                wins.push_back(kcf_window_t{(uint32_t)segs.size(), (uint32_t)w.segments.size()});
                segs.insert(segs.end(), w.segments.begin(), w.segments.end());
            }
        std::vector<kcf_result_t> rows(std::max<size_t>(wins.size(), 1));
        const int rc = kcf_screen_sharded(ctxs.data(), dbs.data(), (int)ctxs.size(), hseqs.data(), (uint32_t)hseqs.size(), wins.data(), wins.size(),
                                          segs.data(), segs.size(), o.minKmerCount, weights, rows.data());
        if (rc == KCF_ERR_WEIGHTS) Logger::error("Data", "Weights should sum to 1.0");
        if (rc != KCF_OK) dev.fail(rc == KCF_ERR_RANGE || rc == KCF_ERR_FASTA ? FAI_CLASS : GV_CLASS);
        Logger::info(GV_CLASS, "Screened " + std::to_string(totalWindows) + " windows on " + std::to_string(ctxs.size()) + " devices");
        out << kcfHeaderText(o, sampleNames[0], index, kmerSize, (int)totalWindows, today());
        size_t base = 0;
        for (size_t s = 0; s < perSeq.size(); ++s) {
            for (size_t i : sorted[s]) out << kcfRowText(perSeq[s][i], rows[base + i], weights) << "\n";
            base += perSeq[s].size();
        }
    } else if (!multi) {
        screenAllSequences(dev, perSeq, kmerSize, o.minKmerCount, weights);
        std::vector<std::vector<kcf_result_t>> results(perSeq.size());
        for (size_t s = 0; s < perSeq.size(); ++s) {
            results[s].resize(perSeq[s].size());
            const int rc = kcf_plan_fetch(dev.ctx, dev.plans[s], results[s].data());
            if (rc == KCF_ERR_WEIGHTS) Logger::error("Data", "Weights should sum to 1.0");
            if (rc != KCF_OK) dev.fail(GV_CLASS);
        }
        Logger::info(GV_CLASS, "Screened " + std::to_string(totalWindows) + " windows");
        out << kcfHeaderText(o, sampleNames[0], index, kmerSize, (int)totalWindows, today());
        for (size_t s = 0; s < perSeq.size(); ++s)
            for (size_t i : sorted[s]) out << kcfRowText(perSeq[s][i], results[s][i], weights) << "\n";
    } else if (multiDevice) {
        // configs[4] of BASELINE.json on one box: the SAMPLES are shared out over the GPUs (database d on device d mod G);
        // every device holds the reference, screens its databases in turn, and the host gathers the rows.  The scores of a
        // plan are bit-identical to the ones `cohort` recomputes from the integers (same formula, same roundings).
        const size_t G = o.devices.size(), D = prefixes.size();
        std::vector<std::vector<std::vector<kcf_result_t>>> rows(D, std::vector<std::vector<kcf_result_t>>(perSeq.size()));
        std::vector<std::string> failure(G);
        auto work = [&](size_t g, Device *dv) {
            try {
                std::unique_ptr<Device> own;
                if (!dv) { // device 0 of the list is `dev`: context, first database and reference are already there
                    own.reset(new Device());
                    dv = own.get();
                    if (kcf_init(o.devices[g], &dv->ctx) != KCF_OK) Logger::error("KMC", std::string(kcf_last_error(nullptr)));
                    uploadReference(*dv);
                }
                for (size_t d = g; d < D; d += G) {
                    if (!(dv == &dev && d == 0)) {
                        for (kcf_plan *p : dv->plans) kcf_plan_destroy(p);
                        dv->plans.clear();
                        if (dv->db) kcf_db_close(dv->db);
                        dv->db = nullptr;
                        if (kcf_db_open(dv->ctx, prefixes[d].c_str(), 0, &dv->db) != KCF_OK) dv->fail("KMC");
                        kcf_db_info_t inf;
                        kcf_db_info(dv->db, &inf);
                        if (inf.kmer_length != kmerSize) Logger::error("KCFHeader", "Kmer size mismatch between the KCFs");
                    }
                    screenAllSequences(*dv, perSeq, kmerSize, o.minKmerCount, weights);
                    for (size_t s = 0; s < perSeq.size(); ++s) {
                        rows[d][s].resize(perSeq[s].size());
                        const int rc = kcf_plan_fetch(dv->ctx, dv->plans[s], rows[d][s].data());
                        if (rc == KCF_ERR_WEIGHTS) Logger::error("Data", "Weights should sum to 1.0");
                        if (rc != KCF_OK) dv->fail(GV_CLASS);
                    }
                    Logger::info(GV_CLASS, "Sample " + sampleNames[d] + " screened on device " + std::to_string(o.devices[g]));
                }
            } catch (const std::exception &e) {
                failure[g] = e.what()[0] ? e.what() : "failed";
            }
        };
        std::vector<std::thread> threads;
        for (size_t g = 1; g < G; ++g) threads.emplace_back(work, g, (Device *)nullptr);
        work(0, &dev);
        for (std::thread &t : threads) t.join();
        for (size_t g = 0; g < G; ++g)
            if (!failure[g].empty()) throw FatalError(failure[g]); // already logged by the worker
        std::string sampleCols = sampleNames[0];
        for (size_t d = 1; d < D; ++d) sampleCols += "\t" + sampleNames[d];
        out << kcfHeaderText(o, sampleCols, index, kmerSize, (int)totalWindows, today());
        for (size_t s = 0; s < perSeq.size(); ++s)
            for (size_t i : sorted[s]) {
                const Window &w = perSeq[s][i];
                KcfRow row;
                row.seq = w.sequenceName;
                row.wid = w.windowId;
                row.start = w.start;
                row.end = w.end;
                row.total = rows[0][s][i].total_kmers;
                row.eff = rows[0][s][i].eff_len;
                for (size_t d = 0; d < D; ++d) {
                    const kcf_result_t &r = rows[d][s][i];
                    if (r.total_kmers != row.total || r.eff_len != row.eff) Logger::error("Cohort", "Windows mismatch found in sample: " + sampleNames[d]);
                    kcf_cell_t c{};
                    c.obs = r.obs;
                    c.variations = r.variations;
                    c.inner = r.inner;
                    c.left = r.left;
                    c.right = r.right;
                    c.ibs = -1;
                    const double kd = r.obs > 0 ? (double)r.kmer_count_sum / r.obs : 0.0;                     // Data.java:87, as written by getVariations
                    c.kmer_count = java_round(std::strtod(java_format_2f(kd).c_str(), nullptr) * (double)r.obs); // as re-read by cohort
                    c.score = r.score;
                    row.cells.push_back(c);
                }
                out << kcfRowTextMulti(row) << "\n";
            }
    } else {
        // every database in turn against the resident reference; its rows go device-to-device into the cohort matrix
        std::vector<size_t> offset(perSeq.size() + 1, 0);
        for (size_t s = 0; s < perSeq.size(); ++s) offset[s + 1] = offset[s] + perSeq[s].size();
        kcf_cohort *co = nullptr;
        if (kcf_cohort_create(dev.ctx, totalWindows, (uint32_t)prefixes.size(), nullptr, nullptr, &co) != KCF_OK) dev.fail("Cohort");
        struct CoGuard {
            kcf_cohort *c;
            ~CoGuard() { kcf_cohort_destroy(c); }
        } guard{co};
        for (size_t d = 0; d < prefixes.size(); ++d) {
            if (d > 0) {
                for (kcf_plan *p : dev.plans) kcf_plan_destroy(p);
                dev.plans.clear();
                kcf_db_close(dev.db);
                dev.db = nullptr;
                if (kcf_db_open(dev.ctx, prefixes[d].c_str(), 0, &dev.db) != KCF_OK) dev.fail("KMC");
                kcf_db_info_t inf;
                kcf_db_info(dev.db, &inf);
                if (inf.kmer_length != kmerSize) Logger::error("KCFHeader", "Kmer size mismatch between the KCFs");
            }
            screenAllSequences(dev, perSeq, kmerSize, o.minKmerCount, weights);
            for (size_t s = 0; s < perSeq.size(); ++s)
                if (kcf_cohort_add_plan(dev.ctx, co, (uint32_t)d, offset[s], dev.plans[s]) != KCF_OK) dev.fail("Cohort");
        }
        // what `cohort` does when it re-reads the per-sample files: the weights come back from their Double.toString text
        // (the same doubles), every score is recomputed from the integers, and KD has been through "%.2f"
        // (Window.java:70: kmerCount = Math.round(KD * OB))
        const int rc = kcf_cohort_scores(dev.ctx, co, weights);
        if (rc == KCF_ERR_WEIGHTS) Logger::error("Data", "Weights should sum to 1.0");
        if (rc != KCF_OK) dev.fail("Cohort");
        std::vector<std::vector<kcf_cell_t>> cols(prefixes.size(), std::vector<kcf_cell_t>(std::max<size_t>(totalWindows, 1)));
        std::vector<int32_t> total(std::max<size_t>(totalWindows, 1)), eff(std::max<size_t>(totalWindows, 1));
        for (size_t d = 0; d < prefixes.size(); ++d)
            if (kcf_cohort_fetch(dev.ctx, co, (uint32_t)d, cols[d].data(), total.data(), eff.data()) != KCF_OK) dev.fail("Cohort");
        std::string sampleCols = sampleNames[0];
        for (size_t d = 1; d < sampleNames.size(); ++d) sampleCols += "\t" + sampleNames[d];
        out << kcfHeaderText(o, sampleCols, index, kmerSize, (int)totalWindows, today());
        for (size_t s = 0; s < perSeq.size(); ++s)
            for (size_t i : sorted[s]) {
                const Window &w = perSeq[s][i];
                KcfRow row;
                row.seq = w.sequenceName;
                row.wid = w.windowId;
                row.start = w.start;
                row.end = w.end;
                row.total = total[offset[s] + i];
                row.eff = eff[offset[s] + i];
                for (size_t d = 0; d < prefixes.size(); ++d) {
                    kcf_cell_t c = cols[d][offset[s] + i];
                    const double kd = c.obs > 0 ? (double)c.kmer_count / c.obs : 0.0;                       // Data.java:87, as written by getVariations
                    c.kmer_count = java_round(std::strtod(java_format_2f(kd).c_str(), nullptr) * (double)c.obs); // as re-read by cohort
                    row.cells.push_back(c);
                }
                out << kcfRowTextMulti(row) << "\n";
            }
    }
    out.flush();
    if (!out) throw FatalError("Error writing KCF file window");
    Logger::info(GV_CLASS, "KCF file written: " + o.outFile);
    return 0;
}

// ================================================================================================ command line
namespace {
const char *const USAGE =
    "Usage: kcftools getVariations [-m] [-c=<minKmerCount>] -f=<featureType> [-g=<gtfFile>] -k=<kmcDBprefix>\n"
    "                              -o=<outFile> [-p=<stepSize>] -r=<refFasta> -s=<sampleName> [-t=<nThreads>]\n"
    "                              [-w=<windowSize>] [--wi=<innerDistanceWeight>] [--wr=<kmerRatioWeight>]\n"
    "                              [--wt=<tailDistanceWeight>] [--device=<cudaOrdinal>]\n"
    " Screen for reference kmers that are not present in the KMC database, and detect variation\n"
    "  -r, --reference=<refFasta>   Reference file name\n"
    "  -k, --kmc=<kmcDBprefix>      KMC database prefix\n"
    "  -o, --output=<outFile>       Output file name\n"
    "  -s, --sample=<sampleName>    Sample name\n"
    "  -f, --feature=<featureType>  Feature type (\"window\" or \"gene\" or \"transcript\")\n"
    "  -t, --threads=<nThreads>     Number of threads [2]\n"
    "  -m, --memory                 Load KMC database into memory\n"
    "      --wi=<innerDistanceWeight> Inner kmer distance weight [0.3]\n"
    "      --wt=<tailDistanceWeight>  Tail kmer distance weight [0.3]\n"
    "      --wr=<kmerRatioWeight>     Kmer ratio weight [0.4]\n"
    "  -w, --window=<windowSize>    Window size\n"
    "  -g, --gtf=<gtfFile>          GTF file name\n"
    "  -c, --min-k-count=<minKmerCount> Minimum kmer count to consider [1]\n"
    "  -p, --step=<stepSize>        Step size for sliding window [window size]\n"
    "      --device=<cudaOrdinal>   CUDA device (this build; the database always lives in HBM, -m and -t are accepted)\n"
    "      --devices=<a,b,...>      several CUDA devices: one database -> its windows are cut over them (one job, rows as with one\n"
    "                               device); several databases (-k x,y,... -s p,q,...) -> the databases are shared out over them\n";

struct OptSpec {
    const char *shortName, *longName;
    int kind; // 0 string, 1 int, 2 double, 3 flag
    bool required;
};
const OptSpec SPECS[] = {{"-r", "--reference", 0, true}, {"-k", "--kmc", 0, true},     {"-o", "--output", 0, true}, {"-s", "--sample", 0, true},
                         {"-f", "--feature", 0, true},   {"-t", "--threads", 1, false}, {"-m", "--memory", 3, false}, {nullptr, "--wi", 2, false},
                         {nullptr, "--wt", 2, false},    {nullptr, "--wr", 2, false},   {"-w", "--window", 1, false}, {"-g", "--gtf", 0, false},
                         {"-c", "--min-k-count", 1, false}, {"-p", "--step", 1, false}, {nullptr, "--device", 1, false}, {nullptr, "--kmer-size", 1, false}, {nullptr, "--devices", 0, false}};

struct Parsed {
    GetVariantsOptions o;
    std::map<std::string, std::string> seen; // long name -> raw value
    int kmerSizeOverride = 0;
};

Parsed parse_options(int argc, const char *const *argv, int first, bool need_kmc)
{
    Parsed ps;
    for (int i = first; i < argc; ++i) {
        std::string a = argv[i], val;
        bool hasVal = false;
        const size_t eq = a.find('=');
        if (a.size() > 1 && a[0] == '-' && eq != std::string::npos) {
            val = a.substr(eq + 1);
            a = a.substr(0, eq);
            hasVal = true;
        }
        const OptSpec *sp = nullptr;
        for (const OptSpec &s : SPECS)
            if ((s.shortName && a == s.shortName) || a == s.longName) sp = &s;
        if (!sp) throw UsageError("Unknown option: '" + std::string(argv[i]) + "'");
        if (sp->kind == 3) {
            ps.seen[sp->longName] = "true";
            continue;
        }
        if (!hasVal) {
            if (i + 1 >= argc) throw UsageError("Missing required parameter for option '" + std::string(sp->longName) + "'");
            val = argv[++i];
        }
        if (sp->kind == 1) {
            char *end = nullptr;
            const long v = std::strtol(val.c_str(), &end, 10);
            if (val.empty() || *end || v > 2147483647L || v < -2147483648L)
                throw UsageError("Invalid value for option '" + std::string(sp->longName) + "': '" + val + "' is not an int");
        } else if (sp->kind == 2) {
            char *end = nullptr;
            std::strtod(val.c_str(), &end);
            if (val.empty() || *end) throw UsageError("Invalid value for option '" + std::string(sp->longName) + "': '" + val + "' is not a double");
        }
        ps.seen[sp->longName] = val;
    }
    std::string missing;
    for (const OptSpec &s : SPECS)
        if (s.required && !ps.seen.count(s.longName) && (need_kmc || (std::string(s.longName) != "--kmc" && std::string(s.longName) != "--output" && std::string(s.longName) != "--sample")))
            missing += std::string(missing.empty() ? "" : ", ") + "'" + s.longName + "'";
    if (!missing.empty()) throw UsageError("Missing required options: " + missing);
    auto S = [&](const char *k, std::string &dst) {
        if (ps.seen.count(k)) dst = ps.seen[k];
    };
    auto I = [&](const char *k, int &dst) {
        if (ps.seen.count(k)) dst = std::atoi(ps.seen[k].c_str());
    };
    auto D = [&](const char *k, double &dst) {
        if (ps.seen.count(k)) dst = std::strtod(ps.seen[k].c_str(), nullptr);
    };
    GetVariantsOptions &o = ps.o;
    S("--reference", o.refFasta);
    S("--kmc", o.kmcDBprefix);
    S("--output", o.outFile);
    S("--sample", o.sampleName);
    S("--feature", o.featureType);
    S("--gtf", o.gtfFile);
    o.hasGtf = ps.seen.count("--gtf") != 0;
    o.loadMemory = ps.seen.count("--memory") != 0;
    I("--threads", o.nThreads);
    I("--window", o.windowSize);
    I("--min-k-count", o.minKmerCount);
    I("--step", o.stepSize);
    I("--device", o.device);
    if (ps.seen.count("--devices")) {
        for (const std::string &d : java_split(ps.seen["--devices"], ',')) {
            char *end = nullptr;
            const long v = std::strtol(d.c_str(), &end, 10);
            if (d.empty() || *end || v < 0 || v > 1023) throw UsageError("Invalid value for option '--devices': '" + d + "' is not a CUDA ordinal");
            o.devices.push_back((int)v);
        }
        if (!o.devices.empty()) o.device = o.devices[0];
    }
    I("--kmer-size", ps.kmerSizeOverride);
    D("--wi", o.innerDistanceWeight);
    D("--wt", o.tailDistanceWeight);
    D("--wr", o.kmerRatioWeight);
    return ps;
}

void printCommandLine(const GetVariantsOptions &o) // HelperFunctions.java:269-291: long names, declaration order, nulls skipped
{
    Logger::info(GV_CLASS, "========== CMD options - GetVariants ==========");
    auto row = [](const char *name, const std::string &v) {
        char buf[512];
        std::snprintf(buf, sizeof buf, "%-15s: %s", name, v.c_str());
        Logger::info(GV_CLASS, buf);
    };
    row("--reference", o.refFasta);
    row("--kmc", o.kmcDBprefix);
    row("--output", o.outFile);
    row("--sample", o.sampleName);
    row("--feature", o.featureType);
    row("--threads", std::to_string(o.nThreads));
    row("--memory", o.loadMemory ? "true" : "false");
    row("--wi", java_double_to_string(o.innerDistanceWeight));
    row("--wt", java_double_to_string(o.tailDistanceWeight));
    row("--wr", java_double_to_string(o.kmerRatioWeight));
    row("--window", std::to_string(o.windowSize));
    if (o.hasGtf) row("--gtf", o.gtfFile);
    row("--min-k-count", std::to_string(o.minKmerCount));
    row("--step", std::to_string(o.stepSize));
    Logger::info(GV_CLASS, "==================================================");
}
} // namespace

int cliMain(int argc, const char *const *argv)
{
    if (argc < 2 || std::string(argv[1]) == "-h" || std::string(argv[1]) == "--help") {
        std::fputs("Usage: kcftools [-h] [COMMAND]\nCommands (this build): getVariations, cohort, findIBS, kcf2gt (each with --help)\n\n", argc < 2 ? stderr : stdout);
        std::fputs(USAGE, argc < 2 ? stderr : stdout);
        return argc < 2 ? 2 : 0;
    }
    const std::string cmd = argv[1];
    const auto t_start = std::chrono::steady_clock::now();
    auto finish = [&](int rc) { // KCFTOOLS.java:44-62
        const long long ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t_start).count();
        char buf[64];
        std::snprintf(buf, sizeof buf, "%02lld:%02lld:%02lld", ms / 3600000, (ms % 3600000) / 60000, (ms % 60000) / 1000);
        if (rc == 0) Logger::info("KCFTOOLS", std::string("Total execution time: ") + buf);
        return rc;
    };
    std::string cmdline;
    for (int i = 0; i < argc; ++i) cmdline += std::string(i ? " " : "") + argv[i];
    try {
        if (cmd == "getVariations") {
            for (int i = 2; i < argc; ++i)
                if (std::string(argv[i]) == "-h" || std::string(argv[i]) == "--help") {
                    std::fputs(USAGE, stdout);
                    return 0;
                }
            Parsed ps = parse_options(argc, argv, 2, true);
            ps.o.commandLine = cmdline;
            printCommandLine(ps.o);
            validateCMD(ps.o);
            return finish(getVariations(ps.o));
        }
        if (cmd == "cohort") return finish(cohortMain(argc, argv, cmdline));
        if (cmd == "findIBS") return finish(findIBSMain(argc, argv, cmdline));
        if (cmd == "kcf2gt") return finish(kcf2gtMain(argc, argv, cmdline));
        if (cmd == "_kcfheader") { // test hook: parse a KCF file's header (and rows) on the host, print the header back
            if (argc < 3) throw UsageError("_kcfheader <file.kcf>");
            const KcfFile f = readKcf(argv[2]);
            std::fputs(f.header.text(today()).c_str(), stdout);
            return 0;
        }
        if (cmd == "_faidx") { // test hook: build / load <fasta>.faidx and print it
            if (argc < 3) throw UsageError("_faidx <fasta>");
            FastaIndex idx(argv[2]);
            for (const FastaIndexEntry &e : idx.entries())
                std::printf("%s\t%d\t%lld\t%d\t%d\n", e.name.c_str(), e.length, (long long)e.offset, e.lineBases, e.lineWidth);
            return 0;
        }
        if (cmd == "_windows") { // test hook: the window / segment lists getVariations would screen (no GPU, no database)
            Parsed ps = parse_options(argc, argv, 2, false);
            if (ps.kmerSizeOverride <= 0) throw UsageError("_windows needs --kmer-size");
            validateCMD(ps.o);
            FastaIndex index(ps.o.refFasta);
            std::unique_ptr<GTF> gtf;
            if (ps.o.featureType != "window") gtf.reset(new GTF(ps.o.gtfFile));
            for (const FastaIndexEntry &e : index.entries())
                for (const Window &w : getWindows(ps.o, index, gtf.get(), e.name, ps.kmerSizeOverride)) {
                    std::printf("W\t%s\t%s\t%d\t%d\t%d", w.windowId.c_str(), w.sequenceName.c_str(), w.start, w.end, w.noFasta ? 1 : 0);
                    for (const kcf_segment_t &s : w.segments) std::printf("\t%d:%d:%d", s.seq_id, s.start0, s.len);
                    std::printf("\n");
                }
            return 0;
        }
        if (cmd == "_hash") { // test hook: String.hashCode() of the arguments (UTF-8 in, UTF-16 code units hashed)
            for (int i = 2; i < argc; ++i) std::printf("%d\n", java_string_hash(argv[i]));
            return 0;
        }
        if (cmd == "_format") { // test hook: Java number formatting of the doubles given as hex bit patterns
            for (int i = 2; i < argc; ++i) {
                const unsigned long long bits = std::strtoull(argv[i], nullptr, 16);
                double d;
                std::memcpy(&d, &bits, 8);
                std::printf("%s\t%s\t%s\n", java_format_2f(d).c_str(), java_double_to_string(d).c_str(), java_float_to_string((float)d).c_str());
            }
            return 0;
        }
        throw UsageError("Unmatched argument at index 0: '" + cmd + "'");
    } catch (const UsageError &e) {
        const bool own = cmd == "cohort" || cmd == "findIBS" || cmd == "kcf2gt"; // their message carries their own usage text
        std::fprintf(stderr, "%s\n%s", e.what(), own ? "" : USAGE);
        return 2;
    } catch (const FatalError &) {
        return 1;
    }
}

} // namespace kcfh
