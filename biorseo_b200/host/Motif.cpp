#include "Motif.h"

uint Motif::delay = 1;

// "id ( first-second first-second )", the text MOIP prints and SecondaryStructure stores
// (reference behaviour: cppsrc/Motif.cpp:182-189).
std::string Motif::pos_string(void) const
{
    std::string s = id_ + " ( ";
    for (const Component& c : comp) s += std::to_string(c.pos.first) + "-" + std::to_string(c.pos.second) + " ";
    s += ")";
    return s;
}

// "<C> components: .... ... basepairs:\ta-b\tc-d" (reference behaviour: cppsrc/Motif.cpp:191-207).
std::string Motif::sec_struct(void) const
{
    std::string s = std::to_string(comp.size()) + " components: ";
    for (const Component& c : comp) {
        for (uint k = c.pos.first; k <= c.pos.second; k++) s += ".";
        s += " ";
    }
    s += "basepairs:";
    for (const Link& l : links_) s += "\t" + std::to_string(l.nts.first) + "-" + std::to_string(l.nts.second);
    return s;
}

// RIN and JSON modules are prefixed with their source (cppsrc/Motif.cpp:209-216).
std::string Motif::get_identifier(void) const
{
    if (source_ == CARNAVAL) return "RIN" + id_;
    if (source_ == JSON) return "JSON" + id_;
    return id_;
}

bool operator==(const Component& c1, const Component& c2) { return c1.pos == c2.pos; }
bool operator!=(const Component& c1, const Component& c2) { return !(c1 == c2); }

bool operator==(const Motif& m1, const Motif& m2)
{
    if (m1.get_identifier() != m2.get_identifier() || m1.score_ != m2.score_ || m1.reversed_ != m2.reversed_) return false;
    for (size_t i = 0; i < m1.comp.size(); i++)
        if (m1.comp[i] != m2.comp[i]) return false;
    return true;
}
bool operator!=(const Motif& m1, const Motif& m2) { return !(m1 == m2); }
