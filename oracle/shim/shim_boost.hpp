// shim_boost.hpp — TEST INFRASTRUCTURE.  The handful of Boost facilities the reference's Instance / Robot /
// Obstacle / PCCBlackmoreSVC sources use, restated so those sources compile here without Boost:
//   boost::geometry   point_xy, polygon<Point, false, true> (counter-clockwise, closed), exterior_ring, get<>,
//                     append, assign, correct (close the ring, make it counter-clockwise), transform with a
//                     2-D matrix_transformer (Boost evaluates  c1*m00 + c2*m01 + m02)
//   boost::tokenizer / char_separator (split on a separator set, empty tokens dropped)
//   boost::program_options::variables_map (just enough for Instance's constructor)
//   boost::math::erf_inv — an INDEPENDENT implementation (bisection on erfl in long double), deliberately not
//                     the oracle's Halley iteration, so the constant is cross-checked rather than copied.
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
}  // namespace math
}  // namespace boost
