// Oracle build shim (test infrastructure, NOT product code).
// IBM CPLEX Concert is absent. The insertion-site search (the path under test) never
// touches the solver; these types only let cppsrc/MOIP.cpp compile unmodified so that its
// constructor can run the real search and fill insertion_sites_.
//
// By default every operation is a no-op. When ilo_shim::recording() is switched on (the
// harness's `model` mode) the types keep what the reference's constraint builder
// (cppsrc/MOIP.cpp:302-328 variables, :467-694 constraints) hands them: variables keep their
// names, expressions their linear terms, IloModel::add the finished constraint in the normal
// form "terms (op) constant". Nothing is solved.
#pragma once
#include <cstddef>
#include <deque>
#include <iostream>
#include <string>
#include <utility>
#include <vector>
typedef double IloNum;
namespace ilo_shim {
inline bool& recording() { static bool on = false; return on; }
inline std::deque<std::string>& names() { static std::deque<std::string> n; return n; }
struct Constraint {
    std::vector<std::pair<double, int>> terms;   // coefficient, variable id (index into names())
    char   op = '<';                             // '<' : <=, '>' : >=, '=' : ==
    double rhs = 0.0;
};
}   // namespace ilo_shim
struct IloEnv {
    std::ostream& getNullStream() { static std::ostream null(nullptr); return null; }
    void end() {}
};
struct IloNumExprArg {
    std::vector<std::pair<double, int>> t;
    double c = 0.0;
    IloNumExprArg& add(const IloNumExprArg& o, double f)
    {
        if (ilo_shim::recording()) {
            for (const auto& x : o.t) t.push_back({f * x.first, x.second});
            c += f * o.c;
        }
        return *this;
    }
};
struct IloNumVar : IloNumExprArg {
    enum Type { Float, Int, Bool };
    IloNumVar() {}
    IloNumVar(IloEnv, IloNum, IloNum, Type, const char* name = nullptr)
    {
        if (ilo_shim::recording()) {
            ilo_shim::names().push_back(name ? name : "?");
            t.push_back({1.0, (int)ilo_shim::names().size() - 1});
        }
    }
};
struct IloNumVarArray {
    IloNumVarArray() : n_(0) {}
    explicit IloNumVarArray(IloEnv) : n_(0) {}
    void add(const IloNumVar& v)
    {
        ++n_;
        if (ilo_shim::recording()) vars_.push_back(v);
    }
    IloNumVar& operator[](std::size_t i)
    {
        if (!ilo_shim::recording()) return v_;
        if (i >= vars_.size()) {   // the reference's "no such variable" sentinel (n*n+1): keep it visible
            static IloNumVar bad;
            if (bad.t.empty()) {
                ilo_shim::names().push_back("INVALID");
                bad.t.push_back({1.0, (int)ilo_shim::names().size() - 1});
            }
            return bad;
        }
        return vars_[i];
    }
    std::size_t getSize() const { return n_; }
  private:
    std::size_t n_;
    IloNumVar v_;
    std::deque<IloNumVar> vars_;
};
struct IloExpr : IloNumExprArg {
    IloExpr() {}
    explicit IloExpr(IloEnv) {}
    IloExpr& operator+=(const IloNumExprArg& o) { add(o, 1.0); return *this; }
    IloExpr& operator+=(IloNum v) { c += v; return *this; }
};
struct IloConstraint {
    ilo_shim::Constraint rec;
    const void* getImpl() const { return this; }
    IloConstraint asConstraint() const { return *this; }
};
struct IloRange : IloConstraint {
    IloRange() {}
    IloRange(IloEnv, IloNum, const IloNumExprArg&, IloNum) {}
};
struct IloObjective : IloConstraint {};
inline IloObjective IloMaximize(IloEnv, const IloNumExprArg&) { return IloObjective(); }
struct IloModel {
    IloModel() : n_(0) {}
    explicit IloModel(IloEnv) : n_(0) {}
    void add(const IloConstraint& k)
    {
        ++n_;
        if (ilo_shim::recording()) added.push_back(k.rec);
    }
    void remove(const IloConstraint&) {}
    struct Iterator {
        explicit Iterator(const IloModel& m) : left_(m.n_) {}
        bool ok() const { return left_ > 0; }
        IloConstraint operator*() const { return IloConstraint(); }
        Iterator& operator++() { --left_; return *this; }
        std::size_t left_;
    };
    std::size_t n_;
    std::vector<ilo_shim::Constraint> added;
};
struct IloConstraintArray { explicit IloConstraintArray(IloEnv) {} };
namespace ilo_shim {
inline IloNumExprArg lin(const IloNumExprArg& a, double fa, const IloNumExprArg* b, double fb, double k)
{
    IloNumExprArg r;
    if (recording()) {
        r.add(a, fa);
        if (b) r.add(*b, fb);
        r.c += k;
    }
    return r;
}
// lhs (op) rhs  ->  terms of (lhs - rhs) (op) constant of (rhs - lhs)
inline IloRange rel(const IloNumExprArg& l, char op, const IloNumExprArg& r)
{
    IloRange out;
    if (recording()) {
        IloNumExprArg d = lin(l, 1.0, &r, -1.0, 0.0);
        out.rec.terms = d.t;
        out.rec.op    = op;
        out.rec.rhs   = -d.c;
    }
    return out;
}
inline IloNumExprArg num(IloNum v) { IloNumExprArg e; e.c = v; return e; }
}   // namespace ilo_shim
inline IloNumExprArg operator+(const IloNumExprArg& a, const IloNumExprArg& b) { return ilo_shim::lin(a, 1.0, &b, 1.0, 0.0); }
inline IloNumExprArg operator+(const IloNumExprArg& a, IloNum v) { return ilo_shim::lin(a, 1.0, nullptr, 0.0, v); }
inline IloNumExprArg operator+(IloNum v, const IloNumExprArg& a) { return ilo_shim::lin(a, 1.0, nullptr, 0.0, v); }
inline IloNumExprArg operator-(const IloNumExprArg& a, const IloNumExprArg& b) { return ilo_shim::lin(a, 1.0, &b, -1.0, 0.0); }
inline IloNumExprArg operator-(const IloNumExprArg& a, IloNum v) { return ilo_shim::lin(a, 1.0, nullptr, 0.0, -v); }
inline IloNumExprArg operator-(IloNum v, const IloNumExprArg& a) { return ilo_shim::lin(a, -1.0, nullptr, 0.0, v); }
inline IloNumExprArg operator-(const IloNumExprArg& a) { return ilo_shim::lin(a, -1.0, nullptr, 0.0, 0.0); }
inline IloNumExprArg operator*(IloNum v, const IloNumExprArg& a) { return ilo_shim::lin(a, v, nullptr, 0.0, 0.0); }
inline IloNumExprArg operator*(const IloNumExprArg& a, IloNum v) { return ilo_shim::lin(a, v, nullptr, 0.0, 0.0); }
inline IloRange operator<=(const IloNumExprArg& a, const IloNumExprArg& b) { return ilo_shim::rel(a, '<', b); }
inline IloRange operator<=(const IloNumExprArg& a, IloNum v) { return ilo_shim::rel(a, '<', ilo_shim::num(v)); }
inline IloRange operator<=(IloNum v, const IloNumExprArg& a) { return ilo_shim::rel(ilo_shim::num(v), '<', a); }
inline IloRange operator>=(const IloNumExprArg& a, const IloNumExprArg& b) { return ilo_shim::rel(a, '>', b); }
inline IloRange operator>=(const IloNumExprArg& a, IloNum v) { return ilo_shim::rel(a, '>', ilo_shim::num(v)); }
inline IloRange operator>=(IloNum v, const IloNumExprArg& a) { return ilo_shim::rel(ilo_shim::num(v), '>', a); }
inline IloRange operator==(const IloNumExprArg& a, const IloNumExprArg& b) { return ilo_shim::rel(a, '=', b); }
inline std::ostream& operator<<(std::ostream& o, const IloConstraint&) { return o << "<constraint>"; }
