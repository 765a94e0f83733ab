// kcf_host.hpp — C++ host side of the getVariations drop-in (the reference's host side is Java; no JDK exists in
// the build image, so the host above the C ABI is C++ and mirrors the reference's classes for this path).
//
//   Logger        Utils/Logger.java:14-32           (ERROR is fatal: the CLI exits with status 1)
//   FastaIndex    Data/FastaIndex.java:26-103, 239-299 (+ FastaIndexEntry)
//   GTF           Data/GTF.java:26-100, 156-163, 207-248, 278-306, 372-444
//   GetVariants   Plugins/GetVariants.java:18-90, 92-183, 278-401 (options, validation, windows, driver)
//   KCF text      Data/KCFHeader.java:291-330, Data/Window.java:125-214, Data/Data.java:120-132, Utils/Configs.java:14-37
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/kcf_b200.h"

namespace kcfh {

// Logger.error => System.exit(1) in the reference; here a FatalError unwinds to main() which returns 1.
struct FatalError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
// picocli usage errors (unknown / missing option): message + usage on stderr, exit status 2
struct UsageError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

namespace Logger {
void info(const std::string &cls, const std::string &msg);
void warning(const std::string &cls, const std::string &msg);
[[noreturn]] void error(const std::string &cls, const std::string &msg);
} // namespace Logger

// ---- Java text formatting ----------------------------------------------------------------------
std::string java_double_to_string(double v); // Double.toString
std::string java_float_to_string(float v);   // Float.toString
std::string java_format_2f(double v);        // String.format("%.2f", v): HALF_UP on the shortest decimal digits
int32_t java_string_hash(const std::string &s); // String.hashCode (ASCII / Latin-1 input)
bool read_file(const std::string &path, std::string &out);
std::vector<std::string> java_split(const std::string &s, char sep);  // String.split(one literal char)
std::vector<std::string> java_lines(const std::string &text);         // BufferedReader.readLine over a whole file
std::string java_trim(const std::string &s);
int java_parse_int(const std::string &s, const std::string &what);    // Integer.parseInt; FatalError when malformed
std::string today();                                                  // HelperFunctions.getTodayDate
std::string kcfStaticHeaderLines();                                   // the ##INFO / ##FORMAT block (Configs.java:14-37)
const char *kcfFormatVersion();

// ---- FastaIndex -------------------------------------------------------------------------------------
struct FastaIndexEntry {
    int seqId = 0;
    std::string name;
    int32_t length = 0;
    int64_t offset = 0;
    int32_t lineBases = 0, lineWidth = 0;
};

class FastaIndex {
  public:
    explicit FastaIndex(const std::string &fastaPath); // builds <fasta>.faidx when missing or older, then mmaps the file
    ~FastaIndex();
    FastaIndex(const FastaIndex &) = delete;
    FastaIndex &operator=(const FastaIndex &) = delete;
    const std::vector<FastaIndexEntry> &entries() const { return entries_; }
    const FastaIndexEntry *getEntry(const std::string &name) const;
    int getSequenceLength(const std::string &name) const; // Logger.error when absent
    // the slice the reference maps for sequence i (FastaIndex.java:54-68)
    const uint8_t *seqBytes(int seqId, uint64_t *n) const;
    static void generateIndexFile(const std::string &fasta, const std::string &fai);

  private:
    std::vector<FastaIndexEntry> entries_;
    std::unordered_map<std::string, int> byName_;
    const uint8_t *map_ = nullptr;
    uint64_t mapLen_ = 0;
    int fd_ = -1;
};

// ---- GTF ------------------------------------------------------------------------------------------
struct Loci {
    std::string chromosome;
    int start = 0, end = 0; // 1-based inclusive
    std::string strand;
    int getLength() const { return end - start + 1; }
};

class GTF {
  public:
    explicit GTF(const std::string &path);
    std::vector<std::string> getGenes(const std::string &chrom) const { return getChildren(chrom); }
    std::vector<std::string> getTranscripts(const std::string &gene) const { return getChildren(gene); }
    std::vector<std::string> getExons(const std::string &tx) const { return getChildren(tx); }
    bool containsVertex(const std::string &id) const { return vertices_.count(id) != 0; }
    Loci getLoci(const std::string &featureID) const; // Logger.error when unknown
    // the merged, sorted loci GTF.getFasta concatenates; empty => the reference returns null (fatal in processWindow)
    std::vector<Loci> mergedLoci(const std::string &featureID, bool isGene) const;
    static std::vector<Loci> mergeOverlappingLoci(std::vector<Loci> lociInHashSetOrder);
    // iteration order of a java.util.HashSet<Loci> filled in the given order (ties of the stable sorts depend on it)
    static std::vector<Loci> javaHashSetOrder(const std::vector<Loci> &insertionOrder);

  private:
    struct Feature {
        std::string chromosome;
        int start, end;
        char strand;
        std::string type, id;
    };
    std::vector<std::string> getChildren(const std::string &parent) const;
    int addVertex(const std::string &v);
    void addEdge(const std::string &from, const std::string &to);
    std::unordered_map<std::string, int> vertices_;                // id -> index
    std::vector<std::vector<std::string>> children_;               // outgoing edges in insertion order
    std::vector<std::unordered_map<std::string, char>> childSet_;  // duplicate-edge guard (simple graph)
    std::unordered_map<std::string, Feature> featureMap_;
};

// ---- windows ---------------------------------------------------------------------------------------
struct Window {
    std::string windowId, sequenceName;
    int start = 0, end = 0;
    std::vector<kcf_segment_t> segments; // what Window.getFasta / GTF.getFasta would concatenate
    bool noFasta = false;                 // GTF.getFasta returned null: fatal when the window is processed
};

// ---- the plugin -------------------------------------------------------------------------------------
struct GetVariantsOptions {             // GetVariants.java:21-61, same names and defaults
    std::string refFasta, kmcDBprefix, outFile, sampleName, featureType, gtfFile;
    bool hasGtf = false;
    int nThreads = 2;
    bool loadMemory = false;
    double innerDistanceWeight = 0.3, tailDistanceWeight = 0.3, kmerRatioWeight = 0.4;
    int windowSize = 0;
    int minKmerCount = 1;
    int stepSize = 0;
    int device = 0;                     // extension: CUDA ordinal (--device)
    std::vector<int> devices;           // extension: --devices a,b,...: several databases (-k x,y,...) are shared out over these GPUs
    std::string commandLine;            // for ##CMD
};

std::vector<Window> getWindows(const GetVariantsOptions &o, const FastaIndex &index, const GTF *gtf,
                               const std::string &sequenceName, int kmerSize); // GetVariants.java:278-352
void validateCMD(const GetVariantsOptions &o);                                 // GetVariants.java:357-386
std::string cleanSampleName(const std::string &s);                             // GetVariants.java:392-401
std::string kcfHeaderText(const GetVariantsOptions &o, const std::string &sample, const FastaIndex &index, int kmerSize,
                          int totalWindows, const std::string &date);          // KCFHeader.java:291-330
std::string kcfRowText(const Window &w, const kcf_result_t &r, const double weights[3]); // Window.java:125-138, Data.java:120-132
double computeScore(const kcf_result_t &r, const double weights[3]);          // Data.java:95-107
int cohortMain(int argc, const char *const *argv, const std::string &cmdline);   // Plugins/Cohort.java
int findIBSMain(int argc, const char *const *argv, const std::string &cmdline);  // Plugins/FindIBS.java
int kcf2gtMain(int argc, const char *const *argv, const std::string &cmdline);   // Plugins/KCFToGenotypeTable.java
int restrictToDevice(int device);                                              // narrow CUDA_VISIBLE_DEVICES before the first CUDA call
int getVariations(GetVariantsOptions o);                                       // GetVariants.java:92-183; 0 or throws
int cliMain(int argc, const char *const *argv);                                // KCFTOOLS.main + picocli parsing

} // namespace kcfh
