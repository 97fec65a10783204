// dd.h — double-double arithmetic (about 31 significant digits) for the HOST-side construction of
// the evaluation tables (pptable.h).  The table coefficients are formed in this precision and
// rounded once to the table's element type, so that the only rounding left in a table evaluation
// is that of the Horner recurrence itself.  Error-free transformations after Dekker / Knuth; the
// translation unit must be compiled without floating-point contraction (build.sh passes
// -ffp-contract=off to the host compiler), std::fma is used explicitly where a fused product is meant.
#pragma once
#include <cmath>

namespace bsk {

struct dd {
    double hi, lo;
    dd() : hi(0.0), lo(0.0) {}
    dd(double h) : hi(h), lo(0.0) {}
    dd(double h, double l) : hi(h), lo(l) {}
    double value() const { return hi + lo; }
};

inline dd two_sum(double a, double b) {
    const double s = a + b;
    const double bb = s - a;
    return dd(s, (a - (s - bb)) + (b - bb));
}
inline dd quick_two_sum(double a, double b) {   // |a| >= |b|
    const double s = a + b;
    return dd(s, b - (s - a));
}
inline dd two_prod(double a, double b) {
    const double p = a * b;
    return dd(p, std::fma(a, b, -p));
}
inline dd operator+(const dd &a, const dd &b) {
    dd s = two_sum(a.hi, b.hi);
    dd t = two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return quick_two_sum(s.hi, s.lo);
}
inline dd operator-(const dd &a) { return dd(-a.hi, -a.lo); }
inline dd operator-(const dd &a, const dd &b) { return a + (-b); }
inline dd operator*(const dd &a, const dd &b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    return quick_two_sum(p.hi, p.lo);
}
inline dd operator/(const dd &a, const dd &b) {
    const double q1 = a.hi / b.hi;
    dd r = a - b * dd(q1);
    const double q2 = r.hi / b.hi;
    r = r - b * dd(q2);
    const double q3 = r.hi / b.hi;
    dd q = quick_two_sum(q1, q2);
    return q + dd(q3);
}
inline dd dd_abs(const dd &a) { return a.hi < 0.0 ? -a : a; }

}  // namespace bsk
