// kcf_tools.cpp — host side of the consumers of getVariations output: the KCF reader and the
// `cohort`, `findIBS` and `kcf2gt` commands with the reference's option names, defaults, messages and output text.
// The numbers (scores, IBS block numbers, allele codes, window filter) are computed on the device through the
// kcf_cohort_* calls of include/kcf_b200.h; this file is text in, text out.
//
//   KcfHeader   Data/KCFHeader.java:44-96 (parse), :291-330 (toString), :333-370 (equals), :420-432 (mergeHeader)
//   readKcf     Data/KCFReader.java:31-105; Data/Window.java:42-83
//   rowText     Data/Window.java:125-152, 170-214; Data/Data.java:120-132
//   cohortMain  Plugins/Cohort.java          findIBSMain  Plugins/FindIBS.java          kcf2gtMain  Plugins/KCFToGenotypeTable.java
#include "kcf_tools.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

namespace kcfh {

static const char *const PARAM_KEYS[8] = {"window", "step", "kmer", "IBS", "nwindow", "wti", "wtt", "wtk"}; // KCFHeader.java:26, 63-90

// Math.round(double): the closest long, ties towards positive infinity (exact, not floor(x + 0.5) in floating point)
long long java_round(double x)
{
    const double f = std::floor(x);
    return (long long)f + ((x - f) >= 0.5 ? 1 : 0);
}

static double java_parse_double(const std::string &s, const std::string &what)
{
    const std::string t = java_trim(s);
    char *end = nullptr;
    const double v = std::strtod(t.c_str(), &end);
    if (t.empty() || *end) throw FatalError("java.lang.NumberFormatException: For input string: \"" + s + "\" (" + what + ")");
    return v;
}

// ================================================================================================ KcfHeader
KcfHeader KcfHeader::parse(const std::string &headerLines)
{
    KcfHeader h;
    for (const std::string &line : java_split(headerLines, '\n')) {
        if (line.rfind("##reference=", 0) == 0) {
            h.reference = line.substr(12);
        } else if (line.rfind("##contig=", 0) == 0) {
            if (line.size() < 11) throw FatalError("java.lang.StringIndexOutOfBoundsException (" + line + ")");
            const std::vector<std::string> f = java_split(line.substr(10, line.size() - 11), ',');
            if (f.size() < 2 || f[0].size() < 3 || f[1].size() < 7) throw FatalError("java.lang.StringIndexOutOfBoundsException (" + line + ")");
            const std::string name = f[0].substr(3);
            const int len = java_parse_int(f[1].substr(7), "contig length");
            bool found = false; // LinkedHashMap.put: a repeated name keeps its position
            for (auto &c : h.contigs)
                if (c.first == name) {
                    c.second = len;
                    found = true;
                }
            if (!found) h.contigs.push_back({name, len});
            h.hasContigs = true;
        } else if (line.rfind("##CMD=", 0) == 0) {
            h.cmds.push_back(line.substr(6));
        } else if (line.rfind("#CHROM", 0) == 0) {
            const std::vector<std::string> f = java_split(line, '\t');
            if (f.size() < 7) throw FatalError("java.lang.NegativeArraySizeException (" + line + ")");
            h.samples.assign(f.begin() + 7, f.end());
            h.hasSamples = true;
        } else if (line.rfind("##PARAM=", 0) == 0) {
            if (line.size() < 10) throw FatalError("java.lang.StringIndexOutOfBoundsException (" + line + ")");
            const std::vector<std::string> f = java_split(line.substr(9, line.size() - 10), ',');
            if (f.size() < 2 || f[0].size() < 3 || f[1].size() < 6) throw FatalError("java.lang.StringIndexOutOfBoundsException (" + line + ")");
            const std::string key = f[0].substr(3), value = f[1].substr(6);
            for (int i = 0; i < 8; ++i)
                if (key == PARAM_KEYS[i]) {
                    h.hasParam[i] = true;
                    h.param[i] = value;
                }
        }
    }
    return h;
}

int KcfHeader::intParam(int i) const { return hasParam[i] ? java_parse_int(param[i], PARAM_KEYS[i]) : 0; }
double KcfHeader::dblParam(int i) const { return hasParam[i] ? java_parse_double(param[i], PARAM_KEYS[i]) : 0.0; }
bool KcfHeader::isIBS() const
{
    if (!hasParam[3]) return false;
    std::string v = param[3]; // Boolean.parseBoolean: equalsIgnoreCase("true")
    for (char &c : v) c = (char)std::tolower((unsigned char)c);
    return v == "true";
}
void KcfHeader::setParam(int i, const std::string &v)
{
    hasParam[i] = true;
    param[i] = v;
}

std::string KcfHeader::mismatch(const KcfHeader &o) const
{
    // KCFHeader.equals, :333-370: the first failing test is logged as a (fatal) error
    if (windowSize() != o.windowSize()) return "Window size mismatch between the KCFs";
    if (kmerSize() != o.kmerSize()) return "Kmer size mismatch between the KCFs";
    if (isIBS() != o.isIBS()) return "IBS processing mismatch between the KCFs";
    if (windowCount() != o.windowCount()) return "Number of windows mismatch between the KCFs";
    if (dblParam(5) != o.dblParam(5)) return "Weight Inner Distance mismatch between the KCFs";
    if (dblParam(6) != o.dblParam(6)) return "Weight Tail Distance mismatch between the KCFs";
    if (dblParam(7) != o.dblParam(7)) return "Weight Kmer Ratio mismatch between the KCFs";
    if (stepSize() != o.stepSize()) return "Step size mismatch between the KCFs";
    return "";
}

std::string KcfHeader::text(const std::string &date) const
{
    std::ostringstream sb;
    sb << "##format=KCF" << kcfFormatVersion() << "\n##date=" << date << "\n##source=kcftools\n##reference=" << reference << "\n";
    if (hasContigs)
        for (const auto &c : contigs) sb << "##contig=<ID=" << c.first << ",length=" << c.second << ">\n";
    sb << kcfStaticHeaderLines();
    for (int i = 0; i < 8; ++i)
        if (hasParam[i]) sb << "##PARAM=<ID=" << PARAM_KEYS[i] << ",value=" << param[i] << ">\n";
    for (const std::string &c : cmds) sb << "##CMD=" << c << "\n";
    sb << "#CHROM\tSTART\tEND\tID\tTOTAL_KMERS\tINFO\tFORMAT";
    if (hasSamples)
        for (const std::string &s : samples) sb << "\t" << s;
    sb << "\n";
    return sb.str();
}

int KcfHeader::contigId(const std::string &name) const
{
    for (size_t i = 0; i < contigs.size(); ++i)
        if (contigs[i].first == name) return (int)i;
    Logger::error("KCFHeader", "Contig " + name + " not found in the KCF header");
}

// ================================================================================================ reader
static kcf_cell_t parse_cell(const std::string &field)
{
    // Window.parseSampleData, Window.java:58-83 (the score is filled in on the device)
    const std::vector<std::string> s = java_split(field, ':');
    if (s.size() < 7) throw FatalError("java.lang.ArrayIndexOutOfBoundsException (" + field + ")");
    kcf_cell_t c{};
    c.ibs = s[0] == "N" ? -1 : java_parse_int(s[0], "GT");
    c.variations = java_parse_int(s[1], "VA");
    c.obs = java_parse_int(s[2], "OB");
    c.inner = java_parse_int(s[3], "ID");
    c.left = java_parse_int(s[4], "LD");
    c.right = java_parse_int(s[5], "RD");
    c.kmer_count = java_round(java_parse_double(s[6], "KD") * (double)c.obs);
    c.score = 0.0;
    return c;
}

KcfFile readKcf(const std::string &path)
{
    Logger::info("KCFReader", "Reading KCF file:" + path);
    std::string text;
    if (!read_file(path, text)) throw FatalError("java.io.FileNotFoundException: " + path + " (No such file or directory)");
    const std::vector<std::string> lines = java_lines(text);
    size_t i = 0;
    while (i < lines.size() && lines[i].rfind("##", 0) == 0) ++i;
    if (i >= lines.size()) throw FatalError("java.lang.NullPointerException: KCF file without a #CHROM line: " + path);
    std::string hl;
    for (size_t j = 0; j <= i; ++j) hl += lines[j] + "\n";
    KcfFile f;
    f.header = KcfHeader::parse(hl);
    // WindowIterator skips up to the #CHROM line (KCFReader.java:66-73)
    size_t r0 = 0;
    while (r0 < lines.size() && lines[r0].rfind("#CHROM", 0) != 0) ++r0;
    const size_t ns = f.header.samples.size();
    for (size_t j = r0 + 1; j < lines.size(); ++j) {
        const std::vector<std::string> fl = java_split(lines[j], '\t');
        if (fl.size() < 6) throw FatalError("java.lang.ArrayIndexOutOfBoundsException (row " + std::to_string(j + 1) + " of " + path + ")");
        KcfRow row;
        row.seq = fl[0];
        row.start = java_parse_int(fl[1], "START");
        row.end = java_parse_int(fl[2], "END");
        row.wid = fl[3];
        row.total = java_parse_int(fl[4], "TOTAL_KMERS");
        bool haveEff = false;
        for (const std::string &kv : java_split(fl[5], ';')) { // Window.getInfoFieldMap
            const std::vector<std::string> p = java_split(kv, '=');
            if (p.size() < 2) throw FatalError("java.lang.ArrayIndexOutOfBoundsException (INFO of row " + std::to_string(j + 1) + ")");
            if (p[0] == "EFFLEN") {
                row.eff = java_parse_int(p[1], "EFFLEN");
                haveEff = true;
            }
        }
        if (!haveEff) throw FatalError("java.lang.NumberFormatException: null (EFFLEN missing in row " + std::to_string(j + 1) + ")");
        if (fl.size() > 7 + ns) throw FatalError("java.lang.ArrayIndexOutOfBoundsException (more sample columns than header samples in row " + std::to_string(j + 1) + ")");
        if (fl.size() < 7 + ns) throw FatalError("row " + std::to_string(j + 1) + " of " + path + " has fewer sample columns than the header");
        for (size_t k = 7; k < fl.size(); ++k) row.cells.push_back(parse_cell(fl[k]));
        f.rows.push_back(std::move(row));
    }
    return f;
}

// Window.toString with calculateStats (Window.java:125-214) for any number of samples
std::string kcfRowTextMulti(const KcfRow &r)
{
    int mnO = 2147483647, mxO = -2147483647 - 1, mnV = 2147483647, mxV = -2147483647 - 1;
    float meanO = 0.f, meanV = 0.f;
    double mnS = (double)3.4028234663852886e38f, mxS = (double)1.401298464324817e-45f, meanS = 0.0;
    for (const kcf_cell_t &d : r.cells) {
        if (d.obs < mnO) mnO = d.obs;
        if (d.obs > mxO) mxO = d.obs;
        meanO += (float)d.obs;
        if (d.variations < mnV) mnV = d.variations;
        if (d.variations > mxV) mxV = d.variations;
        meanV += (float)d.variations;
        if (d.score < mnS) mnS = d.score;
        if (d.score > mxS) mxS = d.score;
        meanS += d.score;
    }
    const size_t n = r.cells.size();
    meanO /= (float)n; // 0 / 0 = NaN for a row without samples, as in Java
    meanV /= (float)n;
    meanS /= (double)n;
    std::ostringstream sb;
    sb << r.seq << "\t" << r.start << "\t" << r.end << "\t" << r.wid << "\t" << r.total << "\t";
    sb << "EFFLEN=" << r.eff << ";IS=" << java_format_2f(mnS) << ";XS=" << java_format_2f(mxS) << ";MS=" << java_format_2f(meanS) << ";IO=" << mnO
       << ";XO=" << mxO << ";MO=" << java_format_2f((double)meanO) << ";IV=" << mnV << ";XV=" << mxV << ";MV=" << java_float_to_string(meanV);
    sb << "\tGT:VA:OB:ID:LD:RD:KD:SC";
    for (const kcf_cell_t &d : r.cells) {
        const double kd = d.kmer_count > 0 ? (double)d.kmer_count / d.obs : 0.0; // Data.java:58-60
        sb << "\t" << (d.ibs == -1 ? std::string("N") : std::to_string(d.ibs)) << ":" << d.variations << ":" << d.obs << ":" << d.inner << ":" << d.left << ":"
           << d.right << ":" << java_format_2f(kd) << ":" << java_format_2f(d.score);
    }
    return sb.str();
}

// iteration order of a java.util.HashMap<String, ?> filled with distinct keys in the given order
std::vector<std::string> javaHashMapOrder(const std::vector<std::string> &keys)
{
    size_t cap = 16;
    while (keys.size() > cap * 3 / 4) cap *= 2;
    std::vector<std::pair<uint32_t, size_t>> k(keys.size());
    for (size_t i = 0; i < keys.size(); ++i) {
        uint32_t h = (uint32_t)java_string_hash(keys[i]);
        h ^= h >> 16;
        k[i] = {h & (uint32_t)(cap - 1), i};
    }
    std::sort(k.begin(), k.end());
    std::vector<std::string> out;
    for (auto &e : k) out.push_back(keys[e.second]);
    return out;
}

// ================================================================================================ device matrix
namespace {
struct Dev {
    kcf_ctx *ctx = nullptr;
    kcf_cohort *co = nullptr;
    ~Dev()
    {
        if (co) kcf_cohort_destroy(co);
        if (ctx) kcf_shutdown(ctx);
    }
    void open(int device, const char *cls)
    {
        if (kcf_init(restrictToDevice(device), &ctx) != KCF_OK) Logger::error(cls, std::string(kcf_last_error(nullptr)));
    }
    [[noreturn]] void fail(const char *cls) const { Logger::error(cls, kcf_last_error(ctx)); }
};

// rows (one file, all its samples) -> device matrix; scores computed there with the header's weights
void upload(Dev &dev, const std::vector<KcfRow *> &rows, size_t nSamples, const double w[3], const char *cls)
{
    const size_t n = rows.size();
    std::vector<int32_t> total(n), eff(n);
    for (size_t i = 0; i < n; ++i) {
        total[i] = rows[i]->total;
        eff[i] = rows[i]->eff;
    }
    if (kcf_cohort_create(dev.ctx, n, (uint32_t)std::max<size_t>(nSamples, 1), total.data(), eff.data(), &dev.co) != KCF_OK) dev.fail(cls);
    std::vector<kcf_cell_t> col(std::max<size_t>(n, 1));
    for (size_t s = 0; s < nSamples; ++s) {
        for (size_t i = 0; i < n; ++i) col[i] = rows[i]->cells[s];
        if (kcf_cohort_set_sample(dev.ctx, dev.co, (uint32_t)s, col.data()) != KCF_OK) dev.fail(cls);
    }
    if (nSamples == 0) return;
    const int rc = kcf_cohort_scores(dev.ctx, dev.co, w);
    if (rc == KCF_ERR_WEIGHTS) Logger::error("Data", "Weights should sum to 1.0");
    if (rc != KCF_OK) dev.fail(cls);
}

void download(Dev &dev, const std::vector<KcfRow *> &rows, size_t nSamples, const char *cls)
{
    std::vector<kcf_cell_t> col(std::max<size_t>(rows.size(), 1));
    for (size_t s = 0; s < nSamples; ++s) {
        if (kcf_cohort_fetch(dev.ctx, dev.co, (uint32_t)s, col.data(), nullptr, nullptr) != KCF_OK) dev.fail(cls);
        for (size_t i = 0; i < rows.size(); ++i) rows[i]->cells[s] = col[i];
    }
}

// ---- a small picocli-like option parser -------------------------------------------------------------------------------
struct Opt {
    const char *shortName, *longName;
    int kind; // 0 string, 1 int, 2 double / float, 3 flag
    bool required;
};

struct HelpShown {};

std::map<std::string, std::string> parseOpts(int argc, const char *const *argv, const std::vector<Opt> &specs, const char *usage)
{
    std::map<std::string, std::string> seen;
    for (int i = 2; i < argc; ++i)
        if (std::string(argv[i]) == "-h" || std::string(argv[i]) == "--help") {
            std::fputs(usage, stdout);
            throw HelpShown();
        }
    for (int i = 2; i < argc; ++i) {
        std::string a = argv[i], val;
        bool hasVal = false;
        const size_t eq = a.find('=');
        if (a.size() > 1 && a[0] == '-' && eq != std::string::npos) {
            val = a.substr(eq + 1);
            a = a.substr(0, eq);
            hasVal = true;
        }
        const Opt *sp = nullptr;
        for (const Opt &s : specs)
            if ((s.shortName && a == s.shortName) || a == s.longName) sp = &s;
        if (!sp) throw UsageError("Unknown option: '" + std::string(argv[i]) + "'\n" + usage);
        if (sp->kind == 3) {
            seen[sp->longName] = "true";
            continue;
        }
        if (!hasVal) {
            if (i + 1 >= argc) throw UsageError("Missing required parameter for option '" + std::string(sp->longName) + "'");
            val = argv[++i];
        }
        char *end = nullptr;
        if (sp->kind == 1) {
            const long v = std::strtol(val.c_str(), &end, 10);
            if (val.empty() || *end || v > 2147483647L || v < -2147483648L)
                throw UsageError("Invalid value for option '" + std::string(sp->longName) + "': '" + val + "' is not an int");
        } else if (sp->kind == 2) {
            std::strtod(val.c_str(), &end);
            if (val.empty() || *end) throw UsageError("Invalid value for option '" + std::string(sp->longName) + "': '" + val + "' is not a double");
        }
        seen[sp->longName] = val;
    }
    std::string missing;
    for (const Opt &s : specs)
        if (s.required && !seen.count(s.longName)) missing += std::string(missing.empty() ? "" : ", ") + "'" + s.longName + "'";
    if (!missing.empty()) throw UsageError("Missing required options: " + missing + "\n" + usage);
    return seen;
}

const char *const COHORT_USAGE =
    "Usage: kcftools cohort [-i=<inFiles>[,<inFiles>...]]... [-l=<listFile>] -o=<outFile> [--device=<cudaOrdinal>]\n"
    "Create a cohort of samples kcf files\n"
    "  -i, --input=<inFiles>[,<inFiles>...]   List of samples kcf files\n"
    "  -l, --list=<listFile>     File containing list of samples kcf files\n"
    "  -o, --output=<outFile>    Output file name\n"
    "      --device=<cudaOrdinal> CUDA device (this build)\n";
const char *const FINDIBS_USAGE =
    "Usage: kcftools findIBS [--bed] [--summary] [--var] -i=<inFile> [--min=<minConsecutive>] -o=<outFile>\n"
    "                        [--score=<scoreCutOff>] [--device=<cudaOrdinal>]\n"
    "Find IBS windows in a KCF file\n"
    "      --bed                 Write bed file [default: false]\n"
    "  -i, --input=<inFile>      Input KCF file name\n"
    "      --min=<minConsecutive> Minimum number of consecutive windows [default: 4]\n"
    "  -o, --output=<outFile>    Output KCF file name\n"
    "      --score=<scoreCutOff> Score cut-off [default: 95.00]\n"
    "      --summary             Write summary tsv file [default: false]\n"
    "      --var                 Detect Variable Regions instead of IBS [default: false]\n"
    "      --device=<cudaOrdinal> CUDA device (this build)\n";
const char *const KCF2GT_USAGE =
    "Usage: kcftools kcf2gt [--chrs=<chrsFile>] -i=<inFile> [--maf=<minMAF>] [--max-missing=<maxMissing>] -o=<outFile>\n"
    "                       [--score_a=<scoreA>] [--score_b=<scoreB>] [--score_n=<scoreN>] [--device=<cudaOrdinal>]\n"
    "Convert KCF to Genotype Table\n"
    "      --chrs=<chrsFile>     List file with chromosomes to include\n"
    "  -i, --input=<inFile>      Input KCF file\n"
    "      --maf=<minMAF>        minimum allele frequency to consider a window valid\n"
    "      --max-missing=<maxMissing> maximum proportion of missing data to consider a window valid\n"
    "  -o, --output=<outFile>    Output file\n"
    "      --score_a=<scoreA>    Lower score cut-off for reference allele (default = 95.0)\n"
    "      --score_b=<scoreB>    Lower score cut-off for alternate allele (default = 60.0)\n"
    "      --score_n=<scoreN>    Score value for missing data (default = 30.0)\n"
    "      --device=<cudaOrdinal> CUDA device (this build)\n";

void writeText(const std::string &path, const std::string &text)
{
    std::ofstream out(path, std::ios::binary);
    if (!out) throw FatalError("java.io.FileNotFoundException: " + path + " (No such file or directory)");
    out << text;
    out.flush();
    if (!out) throw FatalError("Error writing " + path);
}
} // namespace

// ================================================================================================ cohort
static int cohortMainImpl(int argc, const char *const *argv, const std::string &cmdline);
int cohortMain(int argc, const char *const *argv, const std::string &cmdline)
{
    try {
        return cohortMainImpl(argc, argv, cmdline);
    } catch (const HelpShown &) {
        return 0;
    }
}

static int cohortMainImpl(int argc, const char *const *argv, const std::string &cmdline)
{
    static const char *const CLS = "Cohort";
    const std::vector<Opt> specs = {{"-o", "--output", 0, true}, {"-i", "--input", 0, false}, {"-l", "--list", 0, false}, {nullptr, "--device", 1, false}};
    auto o = parseOpts(argc, argv, specs, COHORT_USAGE);
    std::vector<std::string> inFiles;
    if (!o.count("--input") && !o.count("--list")) Logger::error(CLS, "No input files provided");
    if (o.count("--input")) inFiles = java_split(o["--input"], ','); // picocli split = ","
    if (o.count("--list")) {
        std::string t;
        if (!read_file(o["--list"], t)) throw FatalError("java.io.FileNotFoundException: " + o["--list"] + " (No such file or directory)");
        inFiles = java_lines(t);
    }
    // Cohort.java:71-101
    KcfHeader header;
    std::vector<KcfRow> windows;                 // first file's rows, in order (LinkedHashMap)
    std::map<std::string, size_t> byId;
    std::vector<std::vector<std::string>> rowSamples; // samples present in each window, in arrival order
    for (size_t i = 0; i < inFiles.size(); ++i) {
        KcfFile f;
        try {
            f = readKcf(inFiles[i]);
        } catch (const FatalError &) {
            Logger::error(CLS, "Error reading KCF file: " + inFiles[i]);
        }
        // the reference recomputes every score while reading; done here per file with that file's weights
        {
            Dev dev;
            dev.open(o.count("--device") ? std::atoi(o["--device"].c_str()) : 0, CLS);
            std::vector<KcfRow *> rp;
            for (KcfRow &r : f.rows) rp.push_back(&r);
            const double w[3] = {f.header.dblParam(5), f.header.dblParam(6), f.header.dblParam(7)};
            upload(dev, rp, f.header.samples.size(), w, CLS);
            download(dev, rp, f.header.samples.size(), CLS);
        }
        if (i == 0) {
            header = f.header;
            for (KcfRow &r : f.rows) {
                auto it = byId.find(r.wid);
                if (it == byId.end()) {
                    byId[r.wid] = windows.size();
                    windows.push_back(r);
                    rowSamples.push_back(f.header.samples);
                } else { // LinkedHashMap.put on an existing key: position kept, value replaced
                    windows[it->second] = r;
                    rowSamples[it->second] = f.header.samples;
                }
            }
        } else {
            const std::string mm = header.mismatch(f.header);
            if (!mm.empty()) Logger::error("KCFHeader", mm);
            if (f.header.hasSamples) { // mergeHeader, KCFHeader.java:420-432
                header.samples.insert(header.samples.end(), f.header.samples.begin(), f.header.samples.end());
                header.hasSamples = true;
            }
            for (const std::string &c : f.header.cmds) header.cmds.push_back(c);
            for (KcfRow &r : f.rows) {
                auto it = byId.find(r.wid);
                if (it == byId.end()) Logger::error(CLS, "Windows mismatch found in sample: " + inFiles[i] + " at window: " + kcfRowTextMulti(r));
                for (size_t s = 0; s < f.header.samples.size(); ++s) {
                    std::vector<std::string> &have = rowSamples[it->second];
                    if (std::find(have.begin(), have.end(), f.header.samples[s]) != have.end())
                        Logger::error("Window", "Sample " + f.header.samples[s] + " already exists in window " + r.wid);
                    have.push_back(f.header.samples[s]);
                    windows[it->second].cells.push_back(r.cells[s]);
                }
            }
        }
    }
    if (inFiles.empty()) throw FatalError("java.lang.AssertionError: no header"); // `assert header != null`
    header.cmds.push_back(cmdline);
    // alignSamplesWithHeader (Window.java:262-268): a sample the window lacks becomes a null Data -> NullPointerException on write
    std::ostringstream out;
    out << header.text(today());
    for (size_t i = 0; i < windows.size(); ++i) {
        KcfRow aligned = windows[i];
        aligned.cells.clear();
        std::map<std::string, size_t> pos; // LinkedHashMap keyed by sample: a repeated header name collapses to one column
        std::vector<std::string> order;
        for (const std::string &s : header.samples) {
            if (pos.count(s)) continue;
            const auto it = std::find(rowSamples[i].begin(), rowSamples[i].end(), s);
            if (it == rowSamples[i].end()) throw FatalError("java.lang.NullPointerException: window " + windows[i].wid + " has no data for sample " + s);
            pos[s] = (size_t)(it - rowSamples[i].begin());
            order.push_back(s);
        }
        for (const std::string &s : order) aligned.cells.push_back(windows[i].cells[pos[s]]);
        out << kcfRowTextMulti(aligned) << "\n";
    }
    writeText(o["--output"], out.str());
    return 0;
}

// ================================================================================================ findIBS
static int findIBSMainImpl(int argc, const char *const *argv, const std::string &cmdline);
int findIBSMain(int argc, const char *const *argv, const std::string &cmdline)
{
    try {
        return findIBSMainImpl(argc, argv, cmdline);
    } catch (const HelpShown &) {
        return 0;
    }
}

static int findIBSMainImpl(int argc, const char *const *argv, const std::string &cmdline)
{
    static const char *const CLS = "FindIBS";
    const std::vector<Opt> specs = {{"-i", "--input", 0, true}, {"-o", "--output", 0, true}, {nullptr, "--var", 3, false},   {nullptr, "--min", 1, false},
                                    {nullptr, "--score", 2, false}, {nullptr, "--summary", 3, false}, {nullptr, "--bed", 3, false}, {nullptr, "--device", 1, false}};
    auto o = parseOpts(argc, argv, specs, FINDIBS_USAGE);
    std::string outFile = o["--output"];
    const bool detectVar = o.count("--var") != 0, writeSummary = o.count("--summary") != 0, writeBed = o.count("--bed") != 0;
    int minConsecutive = o.count("--min") ? std::atoi(o["--min"].c_str()) : 4;
    const float scoreCutOff = o.count("--score") ? std::strtof(o["--score"].c_str(), nullptr) : 95.0f;
    if (outFile.size() < 4 || outFile.compare(outFile.size() - 4, 4, ".kcf") != 0) outFile += ".kcf";

    KcfFile f = readKcf(o["--input"]);
    KcfHeader &header = f.header;
    if (header.stepSize() > 0) {
        minConsecutive = header.windowSize() / header.stepSize();
        Logger::warning(CLS, "Input KCF file is created with step size. Hence we are using the --min = windowSize/stepSize [" + std::to_string(minConsecutive) + "]");
    }
    // windows per chromosome, chromosomes in the iteration order of the reference's HashMap (FindIBS.java:84-115)
    std::vector<std::string> names;
    std::map<std::string, std::vector<size_t>> byChrom;
    for (size_t i = 0; i < f.rows.size(); ++i) {
        if (!byChrom.count(f.rows[i].seq)) names.push_back(f.rows[i].seq);
        byChrom[f.rows[i].seq].push_back(i);
    }
    const std::vector<std::string> chromOrder = javaHashMapOrder(names);
    std::vector<uint32_t> order, chrom;
    for (size_t c = 0; c < chromOrder.size(); ++c)
        for (size_t i : byChrom[chromOrder[c]]) {
            order.push_back((uint32_t)i);
            chrom.push_back((uint32_t)c);
        }
    if (!header.hasSamples) throw FatalError("java.lang.NullPointerException: no samples in the KCF header");
    const size_t ns = header.samples.size();
    for (const std::string &s : header.samples) Logger::info(CLS, "Finding IBS for sample: " + s);
    {
        Dev dev;
        dev.open(o.count("--device") ? std::atoi(o["--device"].c_str()) : 0, CLS);
        std::vector<KcfRow *> rp;
        for (KcfRow &r : f.rows) rp.push_back(&r);
        const double w[3] = {header.dblParam(5), header.dblParam(6), header.dblParam(7)};
        upload(dev, rp, ns, w, CLS);
        if (ns && kcf_cohort_find_ibs(dev.ctx, dev.co, order.data(), chrom.data(), order.size(), detectVar ? 1 : 0, minConsecutive, scoreCutOff) != KCF_OK)
            dev.fail(CLS);
        download(dev, rp, ns, CLS);
    }
    header.setParam(3, "true");
    header.cmds.push_back(cmdline);
    std::ostringstream out;
    out << header.text(today());
    for (uint32_t i : order) out << kcfRowTextMulti(f.rows[i]) << "\n";
    writeText(outFile, out.str());

    if (writeSummary) { // FindIBS.java:172-272
        std::ostringstream sm;
        sm << "Block\tSample\tChromosome\tStart\tEnd\tLength\tTotalBlocks\tIBSBlocks\tIBSProportion\tMeanScore\n";
        const std::string stem = outFile; // String.replace(".kcf", x) replaces EVERY occurrence
        auto replaceAll = [](std::string s, const std::string &from, const std::string &to) {
            size_t p = 0;
            while ((p = s.find(from, p)) != std::string::npos) {
                s.replace(p, from.size(), to);
                p += to.size();
            }
            return s;
        };
        for (size_t s = 0; s < ns; ++s) {
            std::vector<int> blockIds;                      // LinkedHashMap<Integer, List<Window>>
            std::map<int, std::vector<size_t>> blocks;
            size_t c0 = 0;
            for (size_t c = 0; c < chromOrder.size(); ++c) {
                std::vector<size_t> na;
                const size_t nrow = byChrom[chromOrder[c]].size();
                for (size_t q = 0; q < nrow; ++q) {
                    const size_t i = order[c0 + q];
                    const int v = f.rows[i].cells[s].ibs;
                    if (v == -1) {
                        na.push_back(i);
                    } else if (blocks.count(v)) {
                        blocks[v].insert(blocks[v].end(), na.begin(), na.end());
                        blocks[v].push_back(i);
                        na.clear();
                    } else {
                        blockIds.push_back(v);
                        blocks[v] = {i};
                        na.clear();
                    }
                }
                c0 += nrow;
            }
            if (writeBed) {
                std::ostringstream bed;
                for (int b : blockIds) {
                    const std::vector<size_t> &bl = blocks[b];
                    if (!bl.empty()) bed << f.rows[bl.front()].seq << "\t" << f.rows[bl.front()].start << "\t" << f.rows[bl.back()].end << "\n";
                }
                writeText(replaceAll(stem, ".kcf", "." + header.samples[s] + ".bed"), bed.str());
            }
            for (int b : blockIds) {
                const std::vector<size_t> &bl = blocks[b];
                if (bl.empty()) continue;
                float meanScore = 0.f;
                int ibsBlocks = 0;
                for (size_t i : bl) {
                    meanScore = (float)((double)meanScore + f.rows[i].cells[s].score); // float += double
                    if (f.rows[i].cells[s].ibs != -1) ++ibsBlocks;
                }
                meanScore /= (float)bl.size();
                const float prop = (float)ibsBlocks / (float)bl.size();
                const int start = f.rows[bl.front()].start, end = f.rows[bl.back()].end;
                sm << b << "\t" << header.samples[s] << "\t" << f.rows[bl.front()].seq << "\t" << start << "\t" << end << "\t" << (end - start) << "\t" << bl.size()
                   << "\t" << ibsBlocks << "\t" << java_format_2f((double)prop) << "\t" << java_format_2f((double)meanScore) << "\n";
            }
        }
        writeText(replaceAll(stem, ".kcf", ".summary.tsv"), sm.str());
    }
    return 0;
}

// ================================================================================================ kcf2gt
static int kcf2gtMainImpl(int argc, const char *const *argv, const std::string &cmdline);
int kcf2gtMain(int argc, const char *const *argv, const std::string &cmdline)
{
    try {
        return kcf2gtMainImpl(argc, argv, cmdline);
    } catch (const HelpShown &) {
        return 0;
    }
}

static int kcf2gtMainImpl(int argc, const char *const *argv, const std::string &cmdline)
{
    (void)cmdline;
    static const char *const CLS = "KCFToGenotypeTable";
    const std::vector<Opt> specs = {{"-i", "--input", 0, true},      {"-o", "--output", 0, true},   {nullptr, "--score_a", 2, false},     {nullptr, "--score_b", 2, false},
                                    {nullptr, "--score_n", 2, false}, {nullptr, "--maf", 2, false}, {nullptr, "--max-missing", 2, false}, {nullptr, "--chrs", 0, false},
                                    {nullptr, "--device", 1, false}};
    auto o = parseOpts(argc, argv, specs, KCF2GT_USAGE);
    auto D = [&](const char *k, double dflt) { return o.count(k) ? std::strtod(o[k].c_str(), nullptr) : dflt; };
    double scoreA = D("--score_a", 95.0), scoreB = D("--score_b", 60.0), scoreN = D("--score_n", 30.0);
    const double minMAF = D("--maf", 0.0), maxMissing = D("--max-missing", 1.0);
    // validateScores, KCFToGenotypeTable.java:174-196
    if (scoreA < 0.0 || scoreA > 100.0) Logger::error(CLS, "Score A must be between 0.0 and 100.0");
    if (scoreB < 0.0 || scoreB > 100.0) Logger::error(CLS, "Score B must be between 0.0 and 100.0");
    if (scoreN < 0.0 || scoreN > 100.0) Logger::error(CLS, "Score N must be between 0.0 and 100.0");
    if (scoreA <= scoreB) Logger::error(CLS, "Score A must be greater than Score B");
    if (scoreB == scoreN) {
        Logger::warning(CLS, "Score B is equal to Score N. There would be no alleles scored as het (1).");
        scoreN = scoreB;
    }
    if (scoreB == 0.0 && scoreN != 0.0) {
        Logger::warning(CLS, "Score B is not greater than Score N. There would be no alleles scored as missing (-1) or het (1).");
        scoreN = 0.0;
    }
    KcfFile f = readKcf(o["--input"]);
    const KcfHeader &header = f.header;
    if (!header.hasSamples) throw FatalError("java.lang.NullPointerException: no samples in the KCF header");
    const size_t ns = header.samples.size();
    bool haveChrs = false;
    std::map<std::string, char> chrs;
    if (o.count("--chrs")) {
        std::string t;
        if (!read_file(o["--chrs"], t)) throw FatalError("java.io.FileNotFoundException: " + o["--chrs"] + " (No such file or directory)");
        haveChrs = true;
        for (const std::string &line : java_lines(t)) {
            if ((!line.empty() && line[0] == '#') || java_trim(line).empty()) continue;
            chrs[java_trim(line)] = 1;
        }
    }
    std::vector<int8_t> alleles(std::max<size_t>(f.rows.size() * ns, 1));
    std::vector<uint8_t> bad(std::max<size_t>(f.rows.size(), 1));
    if (ns) {
        Dev dev;
        dev.open(o.count("--device") ? std::atoi(o["--device"].c_str()) : 0, CLS);
        std::vector<KcfRow *> rp;
        for (KcfRow &r : f.rows) rp.push_back(&r);
        const double w[3] = {header.dblParam(5), header.dblParam(6), header.dblParam(7)};
        upload(dev, rp, ns, w, CLS);
        if (kcf_cohort_genotypes(dev.ctx, dev.co, scoreA, scoreB, scoreN, minMAF, maxMissing, alleles.data(), bad.data()) != KCF_OK) dev.fail(CLS);
    } else {
        std::fill(bad.begin(), bad.end(), 1); // all four "count == alleles.length" tests hold for an empty array
    }
    std::ostringstream out;
    out << "# Genotype Table 0:" << java_double_to_string(scoreA) << " - 100.00, 2:" << java_double_to_string(scoreB) << " - " << java_double_to_string(scoreA)
        << ", 1:" << java_double_to_string(scoreN) << " - " << java_double_to_string(scoreB) << ", -1: <=" << java_double_to_string(scoreN) << "\n";
    out << "ID\tCHR\tSTART\tEND";
    for (const std::string &s : header.samples) out << "\t" << s;
    out << "\n";
    std::vector<std::string> contigsMap;
    for (size_t i = 0; i < f.rows.size(); ++i) {
        const KcfRow &r = f.rows[i];
        const int contigID = header.contigId(r.seq) + 1;
        const std::string ent = r.seq + "\t" + std::to_string(contigID);
        if (std::find(contigsMap.begin(), contigsMap.end(), ent) == contigsMap.end()) contigsMap.push_back(ent);
        if (haveChrs && !chrs.count(r.seq)) continue;
        if (bad[i] && (minMAF > 0.0 || maxMissing < 1.0)) continue;
        out << r.wid << "\t" << contigID << "\t" << r.start << "\t" << r.end;
        for (size_t s = 0; s < ns; ++s) out << "\t" << (int)alleles[i * ns + s];
        out << "\n";
    }
    writeText(o["--output"], out.str());
    Logger::info(CLS, "Genotype table written to: " + o["--output"]);
    std::ostringstream cm;
    cm << "contigName\tcontigID\n";
    for (const std::string &e : contigsMap) cm << e << "\n";
    writeText(o["--output"] + ".contigsMap.tsv", cm.str());
    Logger::info(CLS, "Generated Contigs Map file: " + o["--output"] + ".contigsMap.tsv");
    return 0;
}

} // namespace kcfh
