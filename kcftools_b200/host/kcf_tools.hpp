// kcf_tools.hpp — KCF reader / multi-sample rows and the cohort, findIBS, kcf2gt commands (see kcf_tools.cpp).
#pragma once
#include <string>
#include <utility>
#include <vector>
#include "kcf_host.hpp"

namespace kcfh {

long long java_round(double x); // Math.round(double)

struct KcfHeader { // Data/KCFHeader.java
    std::string reference;
    std::vector<std::pair<std::string, int>> contigs; // LinkedHashMap<String, Integer>
    bool hasContigs = false;
    std::vector<std::string> cmds;
    std::vector<std::string> samples;
    bool hasSamples = false;
    bool hasParam[8] = {false, false, false, false, false, false, false, false}; // window, step, kmer, IBS, nwindow, wti, wtt, wtk
    std::string param[8];

    static KcfHeader parse(const std::string &headerLines); // KCFHeader(String), :44-96
    std::string text(const std::string &date) const;       // toString, :291-330
    std::string mismatch(const KcfHeader &o) const;        // equals, :333-370: "" or the message of the first difference
    int intParam(int i) const;
    double dblParam(int i) const;
    void setParam(int i, const std::string &v);
    int windowSize() const { return intParam(0); }
    int stepSize() const { return intParam(1); }
    int kmerSize() const { return intParam(2); }
    bool isIBS() const;
    int windowCount() const { return intParam(4); }
    int contigId(const std::string &name) const; // getContigID, :103-109 (fatal when absent)
};

struct KcfRow { // Data/Window.java as read from a line; cells in the order of the header's samples
    std::string seq, wid;
    int start = 0, end = 0, total = 0, eff = 0;
    std::vector<kcf_cell_t> cells;
};

struct KcfFile {
    KcfHeader header;
    std::vector<KcfRow> rows;
};

KcfFile readKcf(const std::string &path);                                      // KCFReader + Window(String[], ...); scores left at 0
std::string kcfRowTextMulti(const KcfRow &r);                                  // Window.toString, any number of samples
std::vector<std::string> javaHashMapOrder(const std::vector<std::string> &keys); // HashMap<String, ?>.keySet() order

} // namespace kcfh
