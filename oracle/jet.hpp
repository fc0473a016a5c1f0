// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Truncated Taylor-series ("jet") arithmetic of order 3, used by the oracle to obtain
// time derivatives of observed quantities along the ODE flow.  The reference obtains the
// same derivatives symbolically: ModelingToolkit expands D(expr) by the chain rule and
// substitutes D(unknown) -> RHS (reference src/models/cosmologies.jl:99-104 builds ST with
// D(b.v*b.u) and D(D(b.v*γ.Π)) |> expand_derivatives; src/solve.jl:637-657 shows the same
// expansion for user queries).  Picard iteration on jets reproduces that expansion exactly.
#pragma once
#include <cmath>

template <int ORD>
struct JetT {
    static constexpr int N = ORD + 1;
    double c[N]; // Taylor coefficients: x(t0+s) = sum_k c[k] s^k
    JetT() { for (int i = 0; i < N; i++) c[i] = 0; }
    JetT(double v) { for (int i = 0; i < N; i++) c[i] = 0; c[0] = v; }
};
using Jet = JetT<3>;

template <int O> inline JetT<O> operator+(const JetT<O>& a, const JetT<O>& b) { JetT<O> r; for (int i = 0; i <= O; i++) r.c[i] = a.c[i] + b.c[i]; return r; }
template <int O> inline JetT<O> operator-(const JetT<O>& a, const JetT<O>& b) { JetT<O> r; for (int i = 0; i <= O; i++) r.c[i] = a.c[i] - b.c[i]; return r; }
template <int O> inline JetT<O> operator-(const JetT<O>& a) { JetT<O> r; for (int i = 0; i <= O; i++) r.c[i] = -a.c[i]; return r; }
template <int O> inline JetT<O> operator*(const JetT<O>& a, const JetT<O>& b) {
    JetT<O> r;
    for (int k = 0; k <= O; k++) { double s = 0; for (int j = 0; j <= k; j++) s += a.c[j] * b.c[k - j]; r.c[k] = s; }
    return r;
}
template <int O> inline JetT<O> operator/(const JetT<O>& a, const JetT<O>& b) {
    JetT<O> r;
    for (int k = 0; k <= O; k++) { double s = a.c[k]; for (int j = 1; j <= k; j++) s -= b.c[j] * r.c[k - j]; r.c[k] = s / b.c[0]; }
    return r;
}
template <int O> inline JetT<O> operator+(const JetT<O>& a, double b) { JetT<O> r = a; r.c[0] += b; return r; }
template <int O> inline JetT<O> operator+(double b, const JetT<O>& a) { return a + b; }
template <int O> inline JetT<O> operator-(const JetT<O>& a, double b) { JetT<O> r = a; r.c[0] -= b; return r; }
template <int O> inline JetT<O> operator-(double b, const JetT<O>& a) { JetT<O> r = -a; r.c[0] += b; return r; }
template <int O> inline JetT<O> operator*(const JetT<O>& a, double b) { JetT<O> r; for (int i = 0; i <= O; i++) r.c[i] = a.c[i] * b; return r; }
template <int O> inline JetT<O> operator*(double b, const JetT<O>& a) { return a * b; }
template <int O> inline JetT<O> operator/(const JetT<O>& a, double b) { JetT<O> r; for (int i = 0; i <= O; i++) r.c[i] = a.c[i] / b; return r; }
template <int O> inline JetT<O> operator/(double b, const JetT<O>& a) { return JetT<O>(b) / a; }
template <int O> inline JetT<O>& operator+=(JetT<O>& a, const JetT<O>& b) { a = a + b; return a; }
template <int O> inline JetT<O>& operator-=(JetT<O>& a, const JetT<O>& b) { a = a - b; return a; }

template <int O> inline JetT<O> exp(const JetT<O>& a) {
    JetT<O> r; r.c[0] = std::exp(a.c[0]);
    for (int k = 1; k <= O; k++) { double s = 0; for (int j = 1; j <= k; j++) s += j * a.c[j] * r.c[k - j]; r.c[k] = s / k; }
    return r;
}
template <int O> inline JetT<O> log(const JetT<O>& a) {
    JetT<O> r; r.c[0] = std::log(a.c[0]);
    for (int k = 1; k <= O; k++) { double s = 0; for (int j = 1; j < k; j++) s += j * r.c[j] * a.c[k - j]; r.c[k] = (a.c[k] - s / k) / a.c[0]; }
    return r;
}
template <int O> inline JetT<O> sqrt(const JetT<O>& a) {
    JetT<O> r; r.c[0] = std::sqrt(a.c[0]);
    for (int k = 1; k <= O; k++) { double s = a.c[k]; for (int j = 1; j < k; j++) s -= r.c[j] * r.c[k - j]; r.c[k] = s / (2 * r.c[0]); }
    return r;
}
template <int O> inline JetT<O> pow(const JetT<O>& a, double p) {
    JetT<O> r; r.c[0] = std::pow(a.c[0], p);
    for (int k = 1; k <= O; k++) { double s = 0; for (int j = 1; j <= k; j++) s += (p * j - (k - j)) * a.c[j] * r.c[k - j]; r.c[k] = s / (k * a.c[0]); }
    return r;
}
template <int O> inline JetT<O> tanh(const JetT<O>& a) {
    JetT<O> t, w; // w = 1 - t^2
    t.c[0] = std::tanh(a.c[0]); w.c[0] = 1 - t.c[0] * t.c[0];
    for (int k = 1; k <= O; k++) {
        double s = 0; for (int j = 1; j <= k; j++) s += j * a.c[j] * w.c[k - j];
        t.c[k] = s / k;
        double q = 0; for (int j = 0; j <= k; j++) q += t.c[j] * t.c[k - j];
        w.c[k] = -q;
    }
    return t;
}
// d/ds of a jet (drops one order; top coefficient becomes 0)
template <int O> inline JetT<O> deriv(const JetT<O>& a) { JetT<O> r; for (int k = 0; k < O; k++) r.c[k] = (k + 1) * a.c[k + 1]; r.c[O] = 0; return r; }
