// kepler_device.cuh -- Cartesian -> Kepler elements and their exact 6x6 Jacobian on the device.
//
// Replaces, batched, what Kessler does once per (CDM, object) on the CPU
// (/root/reference/kessler/util.py:187-302): `from_cartesian_to_keplerian` (torch) and
// `keplerian_cartesian_partials`, which obtains d(a,e,i,Omega,omega,M)/d(x,y,z,vx,vy,vz) with six
// autograd backward passes.  Here the same formulas are evaluated once on forward-mode dual numbers
// (value + 6 partials), which gives the same derivatives as reverse-mode autograd up to rounding --
// no finite differences.  SURVEY.md Sec. 8f row 1 ("next" scope).
//
// __host__ __device__ like the other device headers (tests/hostcheck).
#pragma once
#include <math.h>

#include "sgp4_device.cuh"

namespace ksl {

struct Dual6 {
    double v;
    double d[6];
};

KSL_HD Dual6 dconst(double c) {
    Dual6 r;
    r.v = c;
#pragma unroll
    for (int k = 0; k < 6; ++k) r.d[k] = 0.0;
    return r;
}
KSL_HD Dual6 dvar(double c, int idx) {
    Dual6 r = dconst(c);
    r.d[idx] = 1.0;
    return r;
}
KSL_HD Dual6 operator+(const Dual6 &a, const Dual6 &b) {
    Dual6 r;
    r.v = a.v + b.v;
#pragma unroll
    for (int k = 0; k < 6; ++k) r.d[k] = a.d[k] + b.d[k];
    return r;
}
KSL_HD Dual6 operator-(const Dual6 &a, const Dual6 &b) {
    Dual6 r;
    r.v = a.v - b.v;
#pragma unroll
    for (int k = 0; k < 6; ++k) r.d[k] = a.d[k] - b.d[k];
    return r;
}
KSL_HD Dual6 operator-(const Dual6 &a) {
    Dual6 r;
    r.v = -a.v;
#pragma unroll
    for (int k = 0; k < 6; ++k) r.d[k] = -a.d[k];
    return r;
}
KSL_HD Dual6 operator*(const Dual6 &a, const Dual6 &b) {
    Dual6 r;
    r.v = a.v * b.v;
#pragma unroll
    for (int k = 0; k < 6; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k];
    return r;
}
KSL_HD Dual6 operator/(const Dual6 &a, const Dual6 &b) {
    Dual6 r;
    const double inv = 1.0 / b.v;
    r.v = a.v * inv;
#pragma unroll
    for (int k = 0; k < 6; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) * inv;
    return r;
}
KSL_HD Dual6 operator*(double s, const Dual6 &a) {
    Dual6 r;
    r.v = s * a.v;
#pragma unroll
    for (int k = 0; k < 6; ++k) r.d[k] = s * a.d[k];
    return r;
}
KSL_HD Dual6 operator+(const Dual6 &a, double s) {
    Dual6 r = a;
    r.v += s;
    return r;
}
KSL_HD Dual6 operator-(double s, const Dual6 &a) { return dconst(s) - a; }
KSL_HD Dual6 chain(const Dual6 &a, double f, double fp) {   // f(a) with derivative fp
    Dual6 r;
    r.v = f;
#pragma unroll
    for (int k = 0; k < 6; ++k) r.d[k] = fp * a.d[k];
    return r;
}
KSL_HD Dual6 dsqrt(const Dual6 &a) {
    const double s = sqrt(a.v);
    return chain(a, s, 0.5 / s);
}
KSL_HD Dual6 dacos(const Dual6 &a) { return chain(a, acos(a.v), -1.0 / sqrt(1.0 - a.v * a.v)); }
KSL_HD Dual6 datan(const Dual6 &a) { return chain(a, atan(a.v), 1.0 / (1.0 + a.v * a.v)); }
KSL_HD Dual6 dtan(const Dual6 &a) {
    const double t = tan(a.v);
    return chain(a, t, 1.0 + t * t);
}
KSL_HD Dual6 dsin(const Dual6 &a) { return chain(a, sin(a.v), cos(a.v)); }
// torch.clamp(x, -1, 1): gradient 1 inside (bounds included), 0 outside
KSL_HD Dual6 dclamp1(const Dual6 &a) {
    if (a.v > 1.0) return dconst(1.0);
    if (a.v < -1.0) return dconst(-1.0);
    return a;
}

KSL_HD double dval(const Dual6 &a) { return a.v; }
KSL_HD double dval(double a) { return a; }
// scalar counterparts so that one template serves both
KSL_HD double dsqrt(double a) { return sqrt(a); }
KSL_HD double dacos(double a) { return acos(a); }
KSL_HD double datan(double a) { return atan(a); }
KSL_HD double dtan(double a) { return tan(a); }
KSL_HD double dsin(double a) { return sin(a); }
KSL_HD double dclamp1(double a) { return a > 1.0 ? 1.0 : (a < -1.0 ? -1.0 : a); }

template <typename T>
KSL_HD T make_const(double c);
template <>
KSL_HD double make_const<double>(double c) { return c; }
template <>
KSL_HD Dual6 make_const<Dual6>(double c) { return dconst(c); }

// util.py:187-264 (elliptic branch).  r_vec, v_vec: position / velocity; out = (a, e, i, Omega, omega, M).
// Returns false for e >= 1 (the reference leaves M undefined there).
template <typename T>
KSL_HD bool cartesian_to_keplerian(const T r_vec[3], const T v_vec[3], double mu, T out[6]) {
    const double kPi2 = 6.283185307179586;
    const T r = dsqrt(r_vec[0] * r_vec[0] + r_vec[1] * r_vec[1] + r_vec[2] * r_vec[2]);
    const T v2 = v_vec[0] * v_vec[0] + v_vec[1] * v_vec[1] + v_vec[2] * v_vec[2];
    const T hx = r_vec[1] * v_vec[2] - r_vec[2] * v_vec[1];
    const T hy = r_vec[2] * v_vec[0] - r_vec[0] * v_vec[2];
    const T hz = r_vec[0] * v_vec[1] - r_vec[1] * v_vec[0];
    const T h = dsqrt(hx * hx + hy * hy + hz * hz);
    const T inc = dacos(hz / h);
    const T nx = -hy, ny = hx;                       // n = K x h = (-hy, hx, 0)
    const T n = dsqrt(nx * nx + ny * ny);
    const T rv = r_vec[0] * v_vec[0] + r_vec[1] * v_vec[1] + r_vec[2] * v_vec[2];
    const T mu_over_r = make_const<T>(mu) / r;
    const T f = v2 - mu_over_r;
    T e_vec[3];
    for (int k = 0; k < 3; ++k) e_vec[k] = (1.0 / mu) * (f * r_vec[k] - rv * v_vec[k]);
    const T e = dsqrt(e_vec[0] * e_vec[0] + e_vec[1] * e_vec[1] + e_vec[2] * e_vec[2]);
    const T energy = 0.5 * v2 - mu_over_r;
    const T a = make_const<T>(-mu) / (2.0 * energy);
    T raw = dacos(dclamp1(nx / n));
    T Omega = (dval(ny) >= 0.0) ? raw : (kPi2 - raw);
    raw = dacos(dclamp1((nx * e_vec[0] + ny * e_vec[1]) / (n * e + 1e-12)));
    T omega = (dval(e_vec[2]) >= 0.0) ? raw : (kPi2 - raw);
    raw = dacos(dclamp1((e_vec[0] * r_vec[0] + e_vec[1] * r_vec[1] + e_vec[2] * r_vec[2]) / (e * r + 1e-12)));
    const T nu = (dval(rv) >= 0.0) ? raw : (kPi2 - raw);
    if (!(dval(e) < 1.0)) return false;
    const T E = 2.0 * datan(dtan(0.5 * nu) * dsqrt((1.0 - e) / (e + 1.0)));
    T M = E - e * dsin(E);
    // wrap to [0, 2pi): a constant shift, derivative unchanged
    Omega = Omega + (-kPi2 * floor(dval(Omega) / kPi2));
    omega = omega + (-kPi2 * floor(dval(omega) / kPi2));
    M = M + (-kPi2 * floor(dval(M) / kPi2));
    out[0] = a; out[1] = e; out[2] = inc; out[3] = Omega; out[4] = omega; out[5] = M;
    return true;
}

// state[6] (x,y,z,vx,vy,vz) -> elements[6] and jac[36] row-major: jac[6*i + j] = d element_i / d state_j
KSL_HD bool kepler_elements_and_partials(const double state[6], double mu, double elements[6], double jac[36]) {
    Dual6 r[3], v[3], out[6];
    for (int k = 0; k < 3; ++k) {
        r[k] = dvar(state[k], k);
        v[k] = dvar(state[3 + k], 3 + k);
    }
    if (!cartesian_to_keplerian<Dual6>(r, v, mu, out)) return false;
    for (int i = 0; i < 6; ++i) {
        elements[i] = out[i].v;
        for (int j = 0; j < 6; ++j) jac[6 * i + j] = out[i].d[j];
    }
    return true;
}

}  // namespace ksl
