// Oracle harness (test infrastructure, NOT product code).
//
// Drives the reference's own, unmodified insertion-site search
// (/root/reference/cppsrc/{MOIP,Motif,Pool,SecondaryStructure}.cpp, compiled where they lie)
// and prints what it found, so that the CUDA path and the CPU restatement in oracle/port can
// be checked against the real thing.
//
//   biorseo_ref ctor  <source> <library> <seqfile> <maskfile>
//       runs MOIP::MOIP (cppsrc/MOIP.cpp:58) end to end with the stubbed solver and dumps
//       insertion_sites_.
//   biorseo_ref model <source> <library> <seqfile> <maskfile>
//       as `ctor`, with the recording solver shim switched on (oracle/shims/ilconcert/ilomodel.h):
//       dumps insertion_sites_ in the order the constructor numbered them ("S" lines) and every
//       constraint of define_problem_constraints (cppsrc/MOIP.cpp:467-694) that involves an
//       insertion variable C_{x,i,p} ("K <op> <rhs> <coef>*<name> ..." lines, terms of lhs - rhs).
//   biorseo_ref jobs  <source> <library> <seqfile> <maskfile> [threads]
//       runs only the search: the per-module jobs MOIP::allowed_motifs_from_{json,rin,desc}
//       (cppsrc/MOIP.cpp:912-1213) on the reference's Pool, dispatched the way
//       cppsrc/MOIP.cpp:146-296 does, and dumps insertion_sites_.
//   biorseo_ref fno   <delay> <seq> <component>...
//       the placement enumerator find_next_ones_in (cppsrc/Motif.cpp:432-494) alone.
//   biorseo_ref placements jsonfolder <library> <seqfile> <maskfile>
//       the reference's real work term: the number of placements find_next_ones_in returns (before the pair
//       filters), summed over the valid modules of a JSON library — components split the way
//       MOIP::allowed_motifs_from_json does (cppsrc/MOIP.cpp:1158-1172). Prints one number.
//   biorseo_ref fasta <file>
//       the reference's FASTA loader Fasta::load (cppsrc/fa.cpp:22-60): one line per record, name \t seq \t str.
//   biorseo_ref bench <source> <library> <seqfile> <maskfile> <threads> <repeat> [max_modules]
//       as `jobs`, timed from the first Pool::push to the last join (validation, JSON parse
//       and directory walk excluded), prints one JSON line.
//
// <source> is jsonfolder | rinfolder | descfolder; Motif::delay follows cppsrc/biorseo.cpp:151,156.
// <seqfile> holds the raw query bytes (one trailing newline is dropped); <maskfile> holds the
// pair mask as n rows of ceil(n/32) little-endian u32 words, bit v of row u = "pij(u,v) > theta".
// Output lines: identifier \t first-second ... \t a-b ... (components, then links).
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <mutex>
#include <map>
#include <stack>
#include <regex>
#include <filesystem>

#define private public
#include "MOIP.h"
#undef private
#include "Pool.h"
#include "fa.h"

using std::string;
using std::vector;
using json = nlohmann::json;

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

static void dump_sites(const vector<Motif>& sites)
{
    string out;
    for (const Motif& m : sites) {
        out += m.get_identifier();
        out += '\t';
        for (size_t j = 0; j < m.comp.size(); j++) {
            if (j) out += ' ';
            out += std::to_string(m.comp[j].pos.first) + "-" + std::to_string(m.comp[j].pos.second);
        }
        out += '\t';
        for (size_t j = 0; j < m.links_.size(); j++) {
            if (j) out += ' ';
            out += std::to_string(m.links_[j].nts.first) + "-" + std::to_string(m.links_[j].nts.second);
        }
        out += '\n';
    }
    std::fwrite(out.data(), 1, out.size(), stdout);
}

// The basepair table MOIP::allowed_basepair reads, built as at cppsrc/MOIP.cpp:77-90 but
// without creating solver variables.
static void fill_yuv(MOIP& moip, float theta)
{
    const size_t n       = moip.rna_.get_RNA_length();
    const size_t missing = n * n + 1;
    size_t       next    = 0;
    moip.index_of_yuv_.assign(n - 6, vector<size_t>());
    for (size_t u = 0; u + 6 < n; u++)
        for (size_t v = u + 4; v < n; v++)
            moip.index_of_yuv_[u].push_back(moip.rna_.get_pij(u, v) > theta ? next++ : missing);
}

struct Dispatch {
    vector<std::function<void()>> jobs;
    json                          library;   // keeps JSON iterators alive
    std::mutex                    sites_mutex;
};

static void collect_jobs(MOIP& moip, const string& source, const string& lib, Dispatch& d, size_t max_modules)
{
    if (source == "jsonfolder") {
        std::ifstream f(lib);
        d.library = json::parse(f);
        for (auto it = d.library.begin(); it != d.library.end(); ++it) {
            if (Motif::is_valid_JSON(it)) continue;
            if (d.jobs.size() >= max_modules) break;
            motif_and_mutex args(it, d.sites_mutex);
            d.jobs.push_back(std::bind(&MOIP::allowed_motifs_from_json, &moip, args));
        }
    } else {
        vector<string> files;
        for (auto& e : std::filesystem::recursive_directory_iterator(lib))
            if (e.is_regular_file()) files.push_back(e.path().string());
        std::sort(files.begin(), files.end());
        for (const string& p : files) {
            if (source == "rinfolder") {
                if (Motif::is_valid_RIN(p)) continue;
                if (d.jobs.size() >= max_modules) break;
                file_and_mutex args(path(p), d.sites_mutex);
                d.jobs.push_back(std::bind(&MOIP::allowed_motifs_from_rin, &moip, args));
            } else {
                if (Motif::is_valid_DESC(p)) continue;
                if (!is_desc_insertible(p, moip.rna_.get_seq())) continue;
                if (d.jobs.size() >= max_modules) break;
                file_and_mutex args(path(p), d.sites_mutex);
                d.jobs.push_back(std::bind(&MOIP::allowed_motifs_from_desc, &moip, args));
            }
        }
    }
}

static double run_jobs(Dispatch& d, int threads)
{
    Pool                pool;
    vector<std::thread> workers;
    for (int i = 0; i < threads; i++) workers.emplace_back(&Pool::infinite_loop_func, &pool);
    auto t0 = std::chrono::steady_clock::now();
    for (auto& j : d.jobs) pool.push(j);
    pool.done();
    for (auto& w : workers) w.join();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

// biorseo_ref fno <delay> <seq> <component>...  : the reference's find_next_ones_in
// (cppsrc/Motif.cpp:432-494) on its own; one placement per line, "first-second ...".
static int run_fno(int argc, char** argv)
{
    Motif::delay = (uint)std::atoi(argv[2]);
    vector<string> comps;
    for (int i = 4; i < argc; i++) comps.push_back(argv[i]);
    for (const vector<Component>& v : find_next_ones_in(argv[3], 0, comps)) {
        string line;
        for (size_t j = 0; j < v.size(); j++) line += (j ? " " : "") + std::to_string(v[j].pos.first) + "-" + std::to_string(v[j].pos.second);
        std::puts(line.c_str());
    }
    return 0;
}

extern const float* g_oracle_pij;

// pairmask <seqfile> <pijfile: n x n float32, row-major> <theta> <empty json library>
// Runs the reference's own MOIP constructor (cppsrc/MOIP.cpp:74-90 builds index_of_yuv_ from
// pij > theta) and prints MOIP::allowed_basepair(u, v) (cppsrc/MOIP.cpp:896-910) for every
// u < v as n rows of ceil(n/32) hexadecimal words.
static int run_pairmask(int argc, char** argv)
{
    if (argc < 6) return 2;
    string seq = read_file(argv[2]);
    if (!seq.empty() && seq.back() == '\n') seq.pop_back();
    const string bytes = read_file(argv[3]);
    const size_t n = seq.size(), W = (n + 31) / 32;
    if (bytes.size() != n * n * 4) {
        std::fprintf(stderr, "pij file holds %zu bytes, expected %zu\n", bytes.size(), n * n * 4);
        return 2;
    }
    vector<float> pij(n * n);
    std::memcpy(pij.data(), bytes.data(), bytes.size());
    g_oracle_pij    = pij.data();
    g_oracle_mask_n = n;
    const float theta = (float)std::atof(argv[4]);
    Motif::delay      = 1;
    RNA  rna("query", seq, false);
    MOIP moip(rna, "jsonfolder", argv[5], theta, false);
    for (size_t u = 0; u < n; u++) {
        vector<uint32_t> row(W, 0u);
        for (size_t v = u + 1; v < n; v++)
            if (moip.allowed_basepair(u, v)) row[v >> 5] |= 1u << (v & 31);
        string line;
        char   buf[16];
        for (size_t w = 0; w < W; w++) {
            std::snprintf(buf, sizeof buf, "%08x", row[w]);
            line += buf;
        }
        std::puts(line.c_str());
    }
    return 0;
}

int main(int argc, char** argv)
{
    if (argc >= 3 && string(argv[1]) == "fasta") {
        std::list<Fasta> records;
        Fasta::load(records, argv[2]);
        for (const Fasta& r : records) std::printf("%s\t%s\t%s\n", r.name().c_str(), r.seq().c_str(), r.str().c_str());
        return 0;
    }
    if (argc >= 5 && string(argv[1]) == "fno") return run_fno(argc, argv);
    if (argc >= 6 && string(argv[1]) == "pairmask") return run_pairmask(argc, argv);
    if (argc < 6) {
        std::fprintf(stderr, "usage: %s ctor|model|jobs|bench <source> <library> <seqfile> <maskfile> [threads] [repeat] [max_modules]\n", argv[0]);
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

    Motif::delay = (source == "descfolder") ? 5 : 1;
    const float theta = 0.001f;   // default of --theta, cppsrc/biorseo.cpp:93
    RNA         rna("query", seq, false);

    if (cmd == "ctor") {
        MOIP moip(rna, source, lib, theta, false);
        dump_sites(moip.insertion_sites_);
        return 0;
    }
    if (cmd == "model") {
        ilo_shim::recording() = true;
        MOIP moip(rna, source, lib, theta, false);
        string out;
        // "S" lines: the sites in the order of insertion_sites_ (= the x of C_{x,i,p})
        for (const Motif& m : moip.insertion_sites_) {
            out += "S\t" + m.get_identifier() + '\t';
            for (size_t j = 0; j < m.comp.size(); j++)
                out += (j ? " " : "") + std::to_string(m.comp[j].pos.first) + "-" + std::to_string(m.comp[j].pos.second);
            out += '\t';
            for (size_t j = 0; j < m.links_.size(); j++)
                out += (j ? " " : "") + std::to_string(m.links_[j].nts.first) + "-" + std::to_string(m.links_[j].nts.second);
            out += '\n';
        }
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
        std::fwrite(out.data(), 1, out.size(), stdout);
        return 0;
    }

    int    threads     = argc > 6 ? std::atoi(argv[6]) : std::max(1, (int)std::thread::hardware_concurrency() - 1);
    int    repeat      = argc > 7 ? std::atoi(argv[7]) : 1;
    size_t max_modules = argc > 8 ? (size_t)std::atoll(argv[8]) : (size_t)-1;
    if (threads < 1) threads = 1;

    MOIP moip;
    moip.verbose_ = false;
    moip.rna_     = rna;
    fill_yuv(moip, theta);

    Dispatch d;
    collect_jobs(moip, source, lib, d, max_modules);

    if (cmd == "placements") {
        if (source != "jsonfolder") { std::fprintf(stderr, "placements: jsonfolder only\n"); return 2; }
        unsigned long long total = 0;
        for (auto it = d.library.begin(); it != d.library.end(); ++it) {
            if (Motif::is_valid_JSON(it)) continue;
            string s = it.value()["sequence"];
            std::transform(s.begin(), s.end(), s.begin(), ::toupper);
            vector<string> comps;
            while (s.find("&") != string::npos) {
                const size_t end = s.find("&");
                comps.push_back(s.substr(0, end));
                s = s.substr(end + 1);
            }
            if (!s.empty()) comps.push_back(s);
            total += find_next_ones_in(seq, 0, comps).size();
        }
        std::printf("%llu\n", total);
        return 0;
    }
    if (cmd == "jobs") {
        run_jobs(d, threads);
        dump_sites(moip.insertion_sites_);
        return 0;
    }
    if (cmd == "bench") {
        double best = 1e300, total = 0;
        size_t sites = 0;
        for (int r = 0; r < repeat; r++) {
            moip.insertion_sites_.clear();
            double t = run_jobs(d, threads);
            best  = std::min(best, t);
            total += t;
            sites = moip.insertion_sites_.size();
        }
        std::printf("{\"seconds_best\": %.9f, \"seconds_mean\": %.9f, \"repeat\": %d, \"threads\": %d, "
                    "\"modules\": %zu, \"n\": %zu, \"sites\": %zu, \"host_cores\": %u}\n",
                    best, total / repeat, repeat, threads, d.jobs.size(), n, sites,
                    std::thread::hardware_concurrency());
        return 0;
    }
    std::fprintf(stderr, "unknown command %s\n", cmd.c_str());
    return 2;
}
