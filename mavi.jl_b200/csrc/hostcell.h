// hostcell.h — update_particle_chunk! (src/chunks.jl:120-147) on the HOST, with exactly the device's arithmetic
// (common.cuh: julia_div_pos / cell_of_point): Base.div(x::Float64, y::Float64) is trunc of the REAL quotient; fl(x/y)
// can be one off when x is a rounded multiple of y, one FMA gives the sign of the exact remainder and fixes it.
// Used where the host has to route particles to the device that owns their cell column (multi.cu) and exported as
// mavi_cells_of_points so that host-side callers partition with the device's rule, not with an approximation of it.
#pragma once
#include <cmath>
#include <cstdint>

namespace mavi_host {

inline double julia_div_pos(double x, double y) {
  const double ax = std::fabs(x);
  double q = std::trunc(ax / y);
  const double rem = std::fma(-q, y, ax);  // exact sign of ax - q*y
  if (rem < 0.0) q -= 1.0;
  else if (rem >= y) q += 1.0;
  return std::copysign(q, x);
}

struct Grid {
  double bl[2], h, cl, ch;
  int cols, rows;
};

// 0-based (col, row) of a point, false when out of grid (BoundsError in the reference)
inline bool cell_of_point(const Grid &g, double x, double y, int *col, int *row) {
  const double rowf = julia_div_pos(-y + g.bl[1] + g.h, g.ch);
  const double colf = julia_div_pos(x - g.bl[0], g.cl);
  if (!(std::fabs(rowf) < 2.0e9) || !(std::fabs(colf) < 2.0e9)) return false;
  int r = (int)rowf + 1, c = (int)colf + 1;
  r -= (r == g.rows + 1) ? 1 : 0;
  c -= (c == g.cols + 1) ? 1 : 0;
  if (r < 1 || r > g.rows || c < 1 || c > g.cols) return false;
  *col = c - 1;
  *row = r - 1;
  return true;
}

}  // namespace mavi_host
