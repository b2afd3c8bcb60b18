// Oracle build shim (test infrastructure, NOT product code).
// boost::split with token_compress_off semantics: one token boundary per character for
// which the predicate holds; the predicate is evaluated exactly once per character, left
// to right (the reference passes a stateful lambda, cppsrc/Motif.cpp:231-235).
#pragma once
#include <string>
#include <vector>
namespace boost {
struct is_any_of_pred {
    std::string set;
    bool operator()(char c) const { return set.find(c) != std::string::npos; }
};
inline is_any_of_pred is_any_of(const char* s) { return is_any_of_pred{std::string(s)}; }
template <class Pred>
inline std::vector<std::string>& split(std::vector<std::string>& out, const std::string& in, Pred pred)
{
    out.clear();
    std::string cur;
    for (char c : in) {
        if (pred(c)) { out.push_back(cur); cur.clear(); }
        else cur.push_back(c);
    }
    out.push_back(cur);
    return out;
}
}  // namespace boost
