// Oracle build shim (test infrastructure, NOT product code).
// Only boost::str(boost::format("%.7f") % double) is used (cppsrc/SecondaryStructure.cpp:102).
#pragma once
#include <cstdio>
#include <string>
namespace boost {
class format {
  public:
    explicit format(const char* f) : fmt_(f) {}
    format& operator%(double v) { char b[128]; std::snprintf(b, sizeof b, fmt_.c_str(), v); out_ += b; return *this; }
    std::string str() const { return out_; }
  private:
    std::string fmt_, out_;
};
inline std::string str(const format& f) { return f.str(); }
}  // namespace boost
