// Oracle harness (test infrastructure, NOT product code).
//
// The reference's MOIP constructor WITH the patch of INTEGRATION.md applied (oracle/patched/make_patched.py): the
// same reference code before and after the insertion-site search — y / x variables, C_{x,i,p} variables,
// define_problem_constraints, objectives — but the search itself done by bsh::InsertionSiteSearch (CUDA, no CPU path).
//
//   biorseo_patched ctor  <source> <library> <seqfile> <maskfile>    dump insertion_sites_ (as biorseo_ref ctor)
//   biorseo_patched model <source> <library> <seqfile> <maskfile>    sites + constraints  (as biorseo_ref model)
//
// <source> is jsonfolder | rinfolder | descfolder | csvfile. Without a usable GPU the library branches exit with
// "ERR: ..." and EXIT_FAILURE: there is nothing to fall back to.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <mutex>
#include <map>
#include <filesystem>

#define private public
#include "MOIP.h"
#undef private

using std::string;
using std::vector;

extern const uint32_t* g_oracle_mask_words;
extern size_t          g_oracle_mask_n;

static string read_file(const string& p)
{
    std::ifstream f(p, std::ios::binary);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", p.c_str()); std::exit(2); }
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

static string site_line(const Motif& m)
{
    string out = m.get_identifier() + '\t';
    for (size_t j = 0; j < m.comp.size(); j++) out += (j ? " " : "") + std::to_string(m.comp[j].pos.first) + "-" + std::to_string(m.comp[j].pos.second);
    out += '\t';
    for (size_t j = 0; j < m.links_.size(); j++) out += (j ? " " : "") + std::to_string(m.links_[j].nts.first) + "-" + std::to_string(m.links_[j].nts.second);
    return out + '\n';
}

int main(int argc, char** argv)
{
    if (argc < 6) {
        std::fprintf(stderr, "usage: %s ctor|model <source> <library> <seqfile> <maskfile>\n", argv[0]);
        return 2;
    }
    const string cmd = argv[1], source = argv[2], lib = argv[3];
    string       seq = read_file(argv[4]);
    if (!seq.empty() && seq.back() == '\n') seq.pop_back();
    const string maskbytes = read_file(argv[5]);
    const size_t n = seq.size(), W = (n + 31) / 32;
    if (maskbytes.size() != n * W * 4) {
        std::fprintf(stderr, "mask file holds %zu bytes, expected %zu\n", maskbytes.size(), n * W * 4);
        return 2;
    }
    vector<uint32_t> mask(n * W);
    std::memcpy(mask.data(), maskbytes.data(), maskbytes.size());
    g_oracle_mask_words = mask.data();
    g_oracle_mask_n     = n;

    Motif::delay = (source == "descfolder") ? 5 : 1;   // cppsrc/biorseo.cpp:151,156 (InsertionSiteSearch::open sets it too)
    RNA rna("query", seq, false);
    if (cmd == "model") ilo_shim::recording() = true;
    MOIP   moip(rna, source, lib, 0.001f, false);
    string out;
    if (cmd == "ctor") {
        for (const Motif& m : moip.insertion_sites_) out += site_line(m);
    } else if (cmd == "model") {
        for (const Motif& m : moip.insertion_sites_) out += "S\t" + site_line(m);
        char buf[64];
        for (const ilo_shim::Constraint& k : moip.model_.added) {
            bool has_c = false;
            for (const auto& t : k.terms) has_c = has_c || ilo_shim::names()[t.second][0] == 'C';
            if (!has_c) continue;
            std::snprintf(buf, sizeof buf, "K %c %g", k.op, k.rhs);
            out += buf;
            for (const auto& t : k.terms) {
                std::snprintf(buf, sizeof buf, " %g*", t.first);
                out += buf;
                out += ilo_shim::names()[t.second];
            }
            out += '\n';
        }
    } else {
        std::fprintf(stderr, "unknown command %s\n", cmd.c_str());
        return 2;
    }
    std::fwrite(out.data(), 1, out.size(), stdout);
    return 0;
}
