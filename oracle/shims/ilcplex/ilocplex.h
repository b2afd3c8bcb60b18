// Oracle build shim (test infrastructure, NOT product code). See ilconcert/ilomodel.h.
// solve() reports "no solution": the ILP is outside the path under test.
#pragma once
#include <ilconcert/ilomodel.h>
#include <exception>
struct IloCplex {
    typedef std::exception Exception;
    IloCplex() {}
    explicit IloCplex(const IloModel&) {}
    void setOut(std::ostream&) {}
    bool solve() { return false; }
    IloNum getValue(const IloNumExprArg&) const { return 0.0; }
};
