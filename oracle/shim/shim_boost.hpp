// shim_boost.hpp — TEST INFRASTRUCTURE.  The handful of Boost facilities the reference's Instance / Robot /
// Obstacle / PCCBlackmoreSVC sources use, restated so those sources compile here without Boost:
//   boost::geometry   point_xy, polygon<Point, false, true> (counter-clockwise, closed), exterior_ring, get<>,
//                     append, assign, correct (close the ring, make it counter-clockwise), transform with a
//                     2-D matrix_transformer (Boost evaluates  c1*m00 + c2*m01 + m02)
//   boost::tokenizer / char_separator (split on a separator set, empty tokens dropped)
//   boost::program_options::variables_map (just enough for Instance's constructor)
//   boost::math::erf_inv — an INDEPENDENT implementation (bisection on erfl in long double), deliberately not
//                     the oracle's Halley iteration, so the constant is cross-checked rather than copied.
// Round 2 additions (goal region, chi-squared and CDF-grid checkers):
//   boost::geometry   buffer(point, multi_polygon, distance_symmetric, side_straight, join_round, end_flat,
//                     point_circle(n)): the n-gon Boost 1.74's point_circle strategy emits (angle 0, then a -= 2pi/n,
//                     i.e. clockwise, closed) reversed to the polygon's declared orientation; disjoint / within on
//                     polygons and points - written with segment-intersection and point-in-ring tests on long double
//                     orientation predicates, deliberately NOT the oracle's separating-axis formulation
//   boost::math       chi_squared + quantile (2 degrees of freedom: -2 log(1-p) in long double), erf
#pragma once
#include <algorithm>
#include <any>
#include <cassert>
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost {
namespace geometry {
namespace model {
namespace d2 {
template <class T>
struct point_xy {
    T v[2] = {0, 0};
    point_xy() = default;
    point_xy(T x, T y) { v[0] = x; v[1] = y; }
    T x() const { return v[0]; }
    T y() const { return v[1]; }
};
}  // namespace d2
template <class P, bool ClockWise = true, bool Closed = true>
struct polygon {
    static const bool clockwise = ClockWise;
    using ring_type = std::vector<P>;
    ring_type ring;
    ring_type& outer() { return ring; }
    const ring_type& outer() const { return ring; }
};
}  // namespace model

template <int D, class T> T get(const model::d2::point_xy<T>& p) { return p.v[D]; }
template <class P, bool CW, bool CL> const std::vector<P>& exterior_ring(const model::polygon<P, CW, CL>& p) { return p.ring; }
template <class P, bool CW, bool CL> std::vector<P>& exterior_ring(model::polygon<P, CW, CL>& p) { return p.ring; }
template <class P> void append(std::vector<P>& ring, const P& pt) { ring.push_back(pt); }
template <class G> void assign(G& dst, const G& src) { dst = src; }

// correct(): a closed polygon type gets its ring closed, and the ring orientation is made to match the
// declared one (counter-clockwise when ClockWise == false) by reversing it when the signed area disagrees.
template <class P, bool CW, bool CL>
void correct(model::polygon<P, CW, CL>& p) {
    auto& r = p.ring;
    if (r.size() < 3) return;
    if (CL && (r.front().v[0] != r.back().v[0] || r.front().v[1] != r.back().v[1])) r.push_back(r.front());
    double a2 = 0;   // twice the signed area, positive = counter-clockwise
    for (size_t i = 0; i + 1 < r.size(); ++i) a2 += r[i].v[0] * r[i + 1].v[1] - r[i + 1].v[0] * r[i].v[1];
    const bool ccw = a2 > 0;
    if (a2 != 0 && ccw == CW) std::reverse(r.begin(), r.end());
}

// ---- round 2: buffer / disjoint / within ----
template <class Poly> using multi_polygon_base = std::vector<Poly>;
namespace model { template <class Poly> struct multi_polygon : std::vector<Poly> {}; }
namespace strategy { namespace buffer {
template <class T> struct distance_symmetric { T d; explicit distance_symmetric(T v) : d(v) {} };
struct join_round { std::size_t n; explicit join_round(std::size_t k = 90) : n(k) {} };
struct end_flat {};
struct side_straight {};
struct point_circle { std::size_t n; explicit point_circle(std::size_t k = 90) : n(k) {} };
}}  // namespace strategy::buffer

template <class T, class Poly, class D, class S, class J, class E>
void buffer(const model::d2::point_xy<T>& pt, model::multi_polygon<Poly>& out, const D& dist, const S&, const J&, const E&,
            const strategy::buffer::point_circle& circle) {
    // strategy::buffer::point_circle::apply (Boost 1.74): two_pi / n steps, clockwise from angle 0, ring closed
    typedef model::d2::point_xy<T> P;
    const double two_pi = 6.283185307179586476925286766559;   // boost::math::constants::two_pi<double>()
    const double diff = two_pi / double(circle.n);
    double a = 0;
    std::vector<P> ring;
    for (std::size_t i = 0; i < circle.n; i++, a -= diff) ring.push_back(P(pt.v[0] + dist.d * std::cos(a), pt.v[1] + dist.d * std::sin(a)));
    ring.push_back(ring.front());
    Poly poly;
    poly.ring = ring;
    // buffered_piece_collection::reverse(): output geometries that are counter-clockwise get their rings reversed
    if (!Poly::clockwise) std::reverse(poly.ring.begin(), poly.ring.end());
    out.clear();
    out.push_back(poly);
}

namespace detail {
typedef long double LD;
template <class P> inline int orient(const P& a, const P& b, const P& c) {   // sign of (b-a) x (c-a)
    const LD v = ((LD)b.v[0] - (LD)a.v[0]) * ((LD)c.v[1] - (LD)a.v[1]) - ((LD)b.v[1] - (LD)a.v[1]) * ((LD)c.v[0] - (LD)a.v[0]);
    return v > 0 ? 1 : (v < 0 ? -1 : 0);
}
template <class P> inline bool on_segment(const P& a, const P& b, const P& p) {   // p collinear with ab assumed
    return std::min(a.v[0], b.v[0]) <= p.v[0] && p.v[0] <= std::max(a.v[0], b.v[0]) && std::min(a.v[1], b.v[1]) <= p.v[1] && p.v[1] <= std::max(a.v[1], b.v[1]);
}
template <class P> inline bool segments_touch(const P& a, const P& b, const P& c, const P& d) {   // share at least one point
    const int o1 = orient(a, b, c), o2 = orient(a, b, d), o3 = orient(c, d, a), o4 = orient(c, d, b);
    if (o1 != o2 && o3 != o4) return true;
    if (o1 == 0 && on_segment(a, b, c)) return true;
    if (o2 == 0 && on_segment(a, b, d)) return true;
    if (o3 == 0 && on_segment(c, d, a)) return true;
    if (o4 == 0 && on_segment(c, d, b)) return true;
    return false;
}
// 1 inside, 0 on the boundary, -1 outside (crossing number; ring may be open or closed, any orientation)
template <class P> inline int point_in_ring(const P& p, const std::vector<P>& r) {
    const std::size_t n = r.size();
    bool in = false;
    for (std::size_t i = 0; i < n; ++i) {
        const P& a = r[i]; const P& b = r[(i + 1) % n];
        if (a.v[0] == b.v[0] && a.v[1] == b.v[1]) continue;
        if (orient(a, b, p) == 0 && on_segment(a, b, p)) return 0;
        if ((a.v[1] > p.v[1]) != (b.v[1] > p.v[1])) {
            const int o = orient(a, b, p);
            if ((b.v[1] > a.v[1]) ? (o > 0) : (o < 0)) in = !in;
        }
    }
    return in ? 1 : -1;
}
}  // namespace detail

template <class P, bool CW, bool CL>
bool disjoint(const model::polygon<P, CW, CL>& A, const model::polygon<P, CW, CL>& B) {
    const auto& a = A.ring; const auto& b = B.ring;
    for (std::size_t i = 0; i < a.size(); ++i)
        for (std::size_t j = 0; j < b.size(); ++j)
            if (detail::segments_touch(a[i], a[(i + 1) % a.size()], b[j], b[(j + 1) % b.size()])) return false;
    if (!a.empty() && detail::point_in_ring(a[0], b) >= 0) return false;
    if (!b.empty() && detail::point_in_ring(b[0], a) >= 0) return false;
    return true;
}
template <class P, bool CW, bool CL>
bool disjoint(const P& p, const model::polygon<P, CW, CL>& B) { return detail::point_in_ring(p, B.ring) < 0; }
// polygon within polygon: no point of A outside B and the interiors meet.  With every vertex of A in-or-on B and no
// PROPER edge crossing, A lies in the closure of B (the rings here are simple).
template <class P, bool CW, bool CL>
bool within(const model::polygon<P, CW, CL>& A, const model::polygon<P, CW, CL>& B) {
    const auto& a = A.ring; const auto& b = B.ring;
    bool any_inside = false;
    for (const P& p : a) { const int s = detail::point_in_ring(p, b); if (s < 0) return false; if (s > 0) any_inside = true; }
    for (std::size_t i = 0; i < a.size(); ++i)
        for (std::size_t j = 0; j < b.size(); ++j) {
            const P& p1 = a[i]; const P& p2 = a[(i + 1) % a.size()]; const P& q1 = b[j]; const P& q2 = b[(j + 1) % b.size()];
            const int o1 = detail::orient(p1, p2, q1), o2 = detail::orient(p1, p2, q2), o3 = detail::orient(q1, q2, p1), o4 = detail::orient(q1, q2, p2);
            if (o1 * o2 < 0 && o3 * o4 < 0) return false;   // proper crossing: part of A leaves B
        }
    if (!any_inside) {   // every vertex on the boundary: interiors meet iff the centroid of A's vertices is strictly inside
        long double cx = 0, cy = 0; std::size_t n = 0;
        for (const P& p : a) { cx += p.v[0]; cy += p.v[1]; ++n; }
        if (n == 0) return false;
        return detail::point_in_ring(P((double)(cx / n), (double)(cy / n)), b) > 0;
    }
    return true;
}

namespace strategy { namespace transform {
template <class T, int D1, int D2>
struct matrix_transformer {
    T m[3][3];
    matrix_transformer(T m00, T m01, T m02, T m10, T m11, T m12, T m20, T m21, T m22)
        : m{{m00, m01, m02}, {m10, m11, m12}, {m20, m21, m22}} {}
    template <class P> P apply(const P& p) const {
        const T c1 = p.v[0], c2 = p.v[1];
        return P(c1 * m[0][0] + c2 * m[0][1] + m[0][2], c1 * m[1][0] + c2 * m[1][1] + m[1][2]);
    }
};
}}  // namespace strategy::transform

template <class G, class S>
bool transform(const G& in, G& out, const S& s) {
    out.ring.clear();
    for (const auto& p : in.ring) out.ring.push_back(s.apply(p));
    return true;
}
}  // namespace geometry

// ---- tokenizer ----
template <class Ch> struct char_separator {
    std::basic_string<Ch> seps;
    char_separator() = default;
    explicit char_separator(const Ch* s) : seps(s) {}
};
template <class Sep> class tokenizer {
public:
    using iterator = std::vector<std::string>::const_iterator;
    tokenizer(const std::string& s, const Sep& sep) {
        std::string cur;
        for (char ch : s) {
            if (sep.seps.find(ch) != std::string::npos) { if (!cur.empty()) toks_.push_back(cur); cur.clear(); }
            else cur.push_back(ch);
        }
        if (!cur.empty()) toks_.push_back(cur);
    }
    iterator begin() const { return toks_.begin(); }
    iterator end() const { return toks_.end(); }
private:
    std::vector<std::string> toks_;
};

// ---- program_options ----
namespace program_options {
struct variable_value {
    std::any v;
    template <class T> const T& as() const { return *std::any_cast<T>(&v); }
};
struct variables_map {
    std::map<std::string, variable_value> m;
    const variable_value& operator[](const std::string& k) const {
        auto it = m.find(k);
        if (it == m.end()) throw std::runtime_error("variables_map: missing " + k);
        return it->second;
    }
    template <class T> void set(const std::string& k, T val) { m[k] = variable_value{std::any(val)}; }
};
}  // namespace program_options

// ---- math ----
namespace math {
template <class T> T erf_inv(T z) {
    if (z <= -1 || z >= 1) return z <= -1 ? -INFINITY : INFINITY;
    long double lo = -7.0L, hi = 7.0L;   // erfl(7) == 1 in long double
    for (int it = 0; it < 90; ++it) {          // 14 / 2^90 is far below the long double ulp
        const long double mid = 0.5L * (lo + hi);
        if (erfl(mid) < (long double)z) lo = mid; else hi = mid;
    }
    return (T)(0.5L * (lo + hi));
}
template <class T> T erf(T z) { return (T)erfl((long double)z); }
// chi_squared distribution, quantile for 2 degrees of freedom only (all the reference ever constructs):
// CDF = 1 - exp(-x/2)  =>  x = -2 log(1 - p)
struct chi_squared { double v; explicit chi_squared(double dof) : v(dof) {} };
inline double quantile(const chi_squared& d, double p) {
    if (d.v != 2) throw std::runtime_error("shim chi_squared: only 2 degrees of freedom");
    return (double)(-2.0L * log1pl(-(long double)p));
}
}  // namespace math
}  // namespace boost
