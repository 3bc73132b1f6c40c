// Stand-in for <boost/geometry.hpp>: TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// Surface used by /root/reference/data/pillars.cpp:14-16,149-165:
//   bg::model::d2::point_xy<double>, bg::model::polygon<Point, ClockWise, Closed>
//   (brace-initialised from one open ring of 4 points), bg::intersection(a, b, out_vector),
//   bg::area(polygon).
//
// Boost.Geometry's overlay is a general polygon clipper; its version is unpinned in the
// reference (install_mods.sh:5).  For two convex rings the intersection is the convex
// polygon a Sutherland-Hodgman clip produces, so this stand-in clips `a` against the
// half-planes of `b` in double precision.  Result areas agree with exact rational
// arithmetic to ~1e-15 (tests/test_oracle_iou.py); Boost's own last-bit behaviour stays
// UNPINNED.  Like Boost, an intersection of zero area yields an empty output vector.
#pragma once
#include <cstddef>
#include <initializer_list>
#include <vector>
namespace boost {
namespace geometry {
namespace model {
namespace d2 {
template <class T>
struct point_xy {
  T x_, y_;
  point_xy() : x_(0), y_(0) {}
  point_xy(T x, T y) : x_(x), y_(y) {}
  T x() const { return x_; }
  T y() const { return y_; }
};
}  // namespace d2
template <class P>
struct ring : public std::vector<P> {
  ring() {}
  ring(std::initializer_list<P> l) : std::vector<P>(l) {}
};
template <class P, bool ClockWise = true, bool Closed = true>
struct polygon {
  typedef P point_type;
  static const bool clockwise = ClockWise;
  ring<P> outer_;
  polygon() {}
  polygon(std::initializer_list<ring<P>> l) {
    if (l.size() > 0) outer_ = *l.begin();
  }
  ring<P>& outer() { return outer_; }
  const ring<P>& outer() const { return outer_; }
};
}  // namespace model

namespace shim_detail {
// signed area, counter-clockwise positive, ring given open (last != first)
template <class R>
double ccw_area(const R& r) {
  const std::size_t n = r.size();
  if (n < 3) return 0.0;
  double s = 0.0;
  for (std::size_t i = 0; i < n; ++i) {
    const std::size_t j = (i + 1 == n) ? 0 : i + 1;
    s += r[i].x() * r[j].y() - r[j].x() * r[i].y();
  }
  return 0.5 * s;
}
}  // namespace shim_detail

template <class P, bool CW, bool Closed>
double area(const model::polygon<P, CW, Closed>& p) {
  const double a = shim_detail::ccw_area(p.outer());
  return CW ? -a : a;  // positive when the ring winds the way its type says
}

template <class PA, bool CWA, bool CA, class PB, bool CWB, bool CB, class PO, bool CWO, bool CO>
bool intersection(const model::polygon<PA, CWA, CA>& a, const model::polygon<PB, CWB, CB>& b,
                  std::vector<model::polygon<PO, CWO, CO>>& out) {
  // subject = a as a CCW ring; clip = b as a CCW ring
  std::vector<PA> subj(a.outer().begin(), a.outer().end());
  if (CWA) subj.assign(a.outer().rbegin(), a.outer().rend());
  std::vector<PB> clip(b.outer().begin(), b.outer().end());
  if (CWB) clip.assign(b.outer().rbegin(), b.outer().rend());
  const std::size_t m = clip.size();
  for (std::size_t e = 0; e < m && !subj.empty(); ++e) {
    const double cx0 = clip[e].x(), cy0 = clip[e].y();
    const double cx1 = clip[(e + 1) % m].x(), cy1 = clip[(e + 1) % m].y();
    const double ex = cx1 - cx0, ey = cy1 - cy0;
    std::vector<PA> next;
    const std::size_t n = subj.size();
    for (std::size_t i = 0; i < n; ++i) {
      const PA& cur = subj[i];
      const PA& prv = subj[(i + n - 1) % n];
      const double dc = ex * (cur.y() - cy0) - ey * (cur.x() - cx0);  // >=0: inside (left)
      const double dp = ex * (prv.y() - cy0) - ey * (prv.x() - cx0);
      const bool cin = dc >= 0.0, pin = dp >= 0.0;
      if (cin != pin) {
        const double t = dp / (dp - dc);
        next.push_back(PA(prv.x() + t * (cur.x() - prv.x()), prv.y() + t * (cur.y() - prv.y())));
      }
      if (cin) next.push_back(cur);
    }
    subj.swap(next);
  }
  if (subj.size() < 3) return true;
  const double ar = shim_detail::ccw_area(subj);
  if (!(ar > 0.0)) return true;  // touching / degenerate: Boost reports no output polygon
  model::polygon<PO, CWO, CO> res;
  if (CWO) {
    for (std::size_t i = subj.size(); i-- > 0;) res.outer().push_back(PO(subj[i].x(), subj[i].y()));
  } else {
    for (std::size_t i = 0; i < subj.size(); ++i) res.outer().push_back(PO(subj[i].x(), subj[i].y()));
  }
  out.push_back(res);
  return true;
}
}  // namespace geometry
}  // namespace boost
