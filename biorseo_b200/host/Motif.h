// Host-side data model of an insertion site, interface-compatible with the reference's
// Motif / Component / Link (cppsrc/Motif.h:22-78) for everything MOIP reads after the
// search (cppsrc/MOIP.cpp:302-328, 347-374, 507-645; cppsrc/SecondaryStructure.cpp:101,139):
// comp[j].pos / .k, links_[k].nts, score_, reversed_, get_identifier(), pos_string(),
// sec_struct(), Motif::delay.
//
// Unlike the reference, a Motif here is never built by re-parsing library files per
// placement: the search returns (module, component starts) records and
// ModuleLibrary::rebuild() turns each one into a Motif from tables computed once at load.
#ifndef BIORSEO_B200_MOTIF_H_
#define BIORSEO_B200_MOTIF_H_

#include <cstddef>
#include <string>
#include <utility>
#include <vector>

typedef unsigned int uint;

typedef enum { RNA3DMOTIF = 1, CSV = 2, CARNAVAL = 3, JSON = 4 } source_type;

typedef struct Comp_ {
    std::pair<uint, uint> pos;    // first and last nucleotide, 0-based inclusive
    size_t                k;      // length
    std::string           seq_;   // unused on this path, kept for layout compatibility
    Comp_(std::pair<int, int> p) : pos(p) { k = 1u + pos.second - pos.first; }
    Comp_(uint start, uint length) : k(length)
    {
        pos.first  = start;
        pos.second = start + length - 1;
    }
} Component;

typedef struct Link {
    std::pair<uint, uint> nts;
    bool                  long_range;
} Link;

class Motif
{
  public:
    Motif(void) : contact_(0), tx_occurrences_(0), score_(0), reversed_(false), source_(RNA3DMOTIF) {}
    Motif(const std::vector<Component>& v, std::string name, source_type source)
    : comp(v), contact_(0), tx_occurrences_(0), score_(0), reversed_(false), id_(name), source_(source) {}

    std::string pos_string(void) const;
    std::string sec_struct(void) const;
    std::string get_identifier(void) const;

    std::vector<Component> comp;
    std::vector<Link>      links_;
    std::vector<uint>      pos_contacts;

    size_t      contact_;
    double      tx_occurrences_;
    double      score_;
    bool        reversed_;
    static uint delay;   // minimal shift between the end of a component and the start of the next

  private:
    std::string id_;
    source_type source_;
};

bool operator==(const Component& c1, const Component& c2);
bool operator!=(const Component& c1, const Component& c2);
bool operator==(const Motif& m1, const Motif& m2);
bool operator!=(const Motif& m1, const Motif& m2);

#endif   // BIORSEO_B200_MOTIF_H_
