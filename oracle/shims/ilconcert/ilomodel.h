// Oracle build shim (test infrastructure, NOT product code).
// IBM CPLEX Concert is absent. The insertion-site search (the path under test) never
// touches the solver; these no-op types only let cppsrc/MOIP.cpp compile unmodified so
// that its constructor can run the real search and fill insertion_sites_.
#pragma once
#include <cstddef>
#include <iostream>
typedef double IloNum;
struct IloEnv {
    std::ostream& getNullStream() { static std::ostream null(nullptr); return null; }
    void end() {}
};
struct IloNumExprArg {};
struct IloNumVar : IloNumExprArg {
    enum Type { Float, Int, Bool };
    IloNumVar() {}
    IloNumVar(IloEnv, IloNum, IloNum, Type, const char* = nullptr) {}
};
struct IloNumVarArray {
    IloNumVarArray() : n_(0) {}
    explicit IloNumVarArray(IloEnv) : n_(0) {}
    void add(const IloNumVar&) { ++n_; }
    IloNumVar& operator[](std::size_t) { return v_; }
    std::size_t getSize() const { return n_; }
  private:
    std::size_t n_;
    IloNumVar v_;
};
struct IloExpr : IloNumExprArg {
    IloExpr() {}
    explicit IloExpr(IloEnv) {}
    IloExpr& operator+=(const IloNumExprArg&) { return *this; }
    IloExpr& operator+=(IloNum) { return *this; }
};
struct IloConstraint {
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
    void add(const IloConstraint&) { ++n_; }
    void remove(const IloConstraint&) {}
    struct Iterator {
        explicit Iterator(const IloModel& m) : left_(m.n_) {}
        bool ok() const { return left_ > 0; }
        IloConstraint operator*() const { return IloConstraint(); }
        Iterator& operator++() { --left_; return *this; }
        std::size_t left_;
    };
    std::size_t n_;
};
struct IloConstraintArray { explicit IloConstraintArray(IloEnv) {} };
inline IloNumExprArg operator+(const IloNumExprArg&, const IloNumExprArg&) { return IloNumExprArg(); }
inline IloNumExprArg operator+(const IloNumExprArg&, IloNum) { return IloNumExprArg(); }
inline IloNumExprArg operator+(IloNum, const IloNumExprArg&) { return IloNumExprArg(); }
inline IloNumExprArg operator-(const IloNumExprArg&, const IloNumExprArg&) { return IloNumExprArg(); }
inline IloNumExprArg operator-(const IloNumExprArg&, IloNum) { return IloNumExprArg(); }
inline IloNumExprArg operator-(IloNum, const IloNumExprArg&) { return IloNumExprArg(); }
inline IloNumExprArg operator-(const IloNumExprArg&) { return IloNumExprArg(); }
inline IloNumExprArg operator*(IloNum, const IloNumExprArg&) { return IloNumExprArg(); }
inline IloNumExprArg operator*(const IloNumExprArg&, IloNum) { return IloNumExprArg(); }
inline IloRange operator<=(const IloNumExprArg&, const IloNumExprArg&) { return IloRange(); }
inline IloRange operator<=(const IloNumExprArg&, IloNum) { return IloRange(); }
inline IloRange operator<=(IloNum, const IloNumExprArg&) { return IloRange(); }
inline IloRange operator>=(const IloNumExprArg&, const IloNumExprArg&) { return IloRange(); }
inline IloRange operator>=(const IloNumExprArg&, IloNum) { return IloRange(); }
inline IloRange operator>=(IloNum, const IloNumExprArg&) { return IloRange(); }
inline IloRange operator==(const IloNumExprArg&, const IloNumExprArg&) { return IloRange(); }
inline std::ostream& operator<<(std::ostream& o, const IloConstraint&) { return o << "<constraint>"; }
