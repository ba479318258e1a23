// walls.cuh — per-particle wall handling: calc_walls_forces! (src/integration.jl:228-266) and walls!
// (src/integration.jl:268-412) as device functions fused into the integrate kernels.
#pragma once
#include "common.cuh"

namespace MAVI_NS {

// potential_force(dr, dist, potential) with an explicit (possibly signed) dist, as PotentialWalls uses it
// (src/configs.jl:354-368, :389-397).  Per particle, not per pair: written like the reference.
__device__ __forceinline__ void wall_potential_force(const DevSpace &sp, const real *pot, real drx, real dry, real dist,
                                                     real &fx, real &fy) {
  real c;
  if (sp.pot_kind == MAVI_POT_HARMTRUNC) {
    if (dist > pot[3]) return;
    real k = (dist < pot[2]) ? pot[0] : pot[1];
    real fmod_ = -k * (dist / pot[2] - 1.0);
    c = fmod_ / dist;
  } else {
    real sigma = pot[0], eps = pot[1];
    real s6 = sigma * sigma * sigma * sigma * sigma * sigma;
    real d2 = dist * dist, d6 = d2 * d2 * d2, d7 = d6 * dist;
    real fmod_ = 4.0 * eps * (12.0 * s6 * s6 / (d6 * d7) - 6.0 * s6 / d7);
    c = fmod_ / dist;
  }
  fx += c * drx;
  fy += c * dry;
}

// process_dist, src/configs.jl:259-261
__device__ __forceinline__ real process_dist(int mode, real dist, real flag) {
  if (mode == MAVI_WALLMODE_OUTSIDE) return flag * dist;
  if (mode == MAVI_WALLMODE_INSIDE) return -flag * dist;
  return dist;
}

// calc_walls_forces! for one particle: every PotentialWalls sub-space acts (ManyWalls loop, :258-264).
// ptype: get_particle_type(state, i) - 1 (the ring type; 0 for states without types), selects the PotentialVector entry.
__device__ __forceinline__ void wall_forces(const DevParams &p, real x, real y, real &fx, real &fy, int ptype = 0) {
  for (int k = 0; k < p.n_spaces; k++) {
    const DevSpace &sp = p.spaces[k];
    if (sp.wall != MAVI_WALL_POTENTIAL) continue;
    const real *pot = sp.n_pot_types > 0 ? sp.pot_t[ptype] : sp.pot;
    if (sp.geom == MAVI_GEOM_CIRCLE) {
      // signed_pos(point, ::CircleCfg), src/configs.jl:154-163
      real d0 = x - sp.cc[0], d1 = y - sp.cc[1];
      real dd = sqrt(d0 * d0 + d1 * d1);
      real drx = d0 - (d0 / dd) * sp.cr, dry = d1 - (d1 / dd) * sp.cr;
      real sd = dd - sp.cr;
      real dist = process_dist(sp.pot_mode, fabs(sd), sign_d(sd));
      wall_potential_force(sp, pot, drx, dry, dist, fx, fy);
    } else if (sp.geom == MAVI_GEOM_LINES) {
      for (int l = 0; l < sp.n_lines; l++) {
        // signed_pos(point, ::Line2D), src/configs.jl:119-135
        const DevLine &ln = sp.lines[l];
        real drx = x - ln.p1[0], dry = y - ln.p1[1];
        real delta_t = drx * ln.tangent[0] + dry * ln.tangent[1];
        if (delta_t > 0.0) {
          if (delta_t < ln.length) {
            drx = x - (ln.p1[0] + ln.tangent[0] * delta_t);
            dry = y - (ln.p1[1] + ln.tangent[1] * delta_t);
          } else {
            drx = x - ln.p2[0];
            dry = y - ln.p2[1];
          }
        }
        real dist = process_dist(sp.pot_mode, sqrt(drx * drx + dry * dry), 1.0);
        wall_potential_force(sp, pot, drx, dry, dist, fx, fy);
      }
    }
  }
}

// walls!(system) for one particle: loops the (wall, geometry) pairs in order (src/integration.jl:404-408).
// vx/vy are only meaningful for SecondLawState (HAS_VEL); pr is the particle's radius.
template <bool HAS_VEL>
__device__ __forceinline__ void apply_walls(const DevParams &p, real &x, real &y, real &vx, real &vy, real pr) {
  if (p.wall_fast == 1) {
    // the common space: ONE periodic rectangle (:309-324); same arithmetic as the generic branch below with the
    // centre bl + size/2 and size/2 taken from the parameter block ((size/2)*2 == size exactly)
    const real dx = x - p.wall_ctr[0], dy = y - p.wall_ctr[1];
    if (fabs(dx) > p.half[0]) x = x - sign_d(dx) * p.size[0];
    if (fabs(dy) > p.half[1]) y = y - sign_d(dy) * p.size[1];
    return;
  }
  for (int k = 0; k < p.n_spaces; k++) {
    const DevSpace &sp = p.spaces[k];
    if (sp.wall == MAVI_WALL_RIGID && sp.geom == MAVI_GEOM_RECT) {
      // :271-285 — velocity flip only while overlapping the wall; radius = particle_radius(dynamic_cfg)
      if (HAS_VEL) {
        real r = p.particle_radius;
        real relx = x - sp.rect_bl[0], rely = y - sp.rect_bl[1];
        bool ox = ((relx + r) > sp.rect_sz[0]) || ((relx - r) < 0.0);
        bool oy = ((rely + r) > sp.rect_sz[1]) || ((rely - r) < 0.0);
        if (ox) vx = -vx;
        if (oy) vy = -vy;
      }
    } else if (sp.wall == MAVI_WALL_RIGID && sp.geom == MAVI_GEOM_CIRCLE) {
      // :287-306 — cross terms use the raw position (kept)
      if (HAS_VEL) {
        real mr = sp.cr - p.particle_radius;
        real ex = x - sp.cc[0], ey = y - sp.cc[1];
        real dr2x = ex * ex, dr2y = ey * ey;
        real r2 = dr2x + dr2y;
        if (r2 > mr * mr) {
          real nvx = (vx * (dr2y - dr2x) - 2.0 * vy * x * y) / r2;
          real nvy = (-vy * (dr2y - dr2x) - 2.0 * vx * x * y) / r2;
          vx = nvx;
          vy = nvy;
        }
      }
    } else if (sp.wall == MAVI_WALL_PERIODIC && sp.geom == MAVI_GEOM_RECT) {
      // :309-324 — strict '>', single image
      real hx = sp.rect_sz[0] / 2.0, hy = sp.rect_sz[1] / 2.0;
      real dx = x - (sp.rect_bl[0] + hx), dy = y - (sp.rect_bl[1] + hy);
      if (fabs(dx) > hx) x = x - sign_d(dx) * (hx * 2.0);
      if (fabs(dy) > hy) y = y - sign_d(dy) * (hy * 2.0);
    } else if (sp.wall == MAVI_WALL_SLIPPERY && sp.geom == MAVI_GEOM_LINES) {
      // :327-378 — pos_i is read once; corrections accumulate
      const real pi0 = x, pi1 = y;
      for (int l = 0; l < sp.n_lines; l++) {
        const DevLine &ln = sp.lines[l];
        real dr0 = pi0 - ln.p1[0], dr1 = pi1 - ln.p1[1];
        real delta_s = dr0 * ln.normal[0] + dr1 * ln.normal[1];
        if (fabs(delta_s) > pr) continue;
        real delta_t = dr0 * ln.tangent[0] + dr1 * ln.tangent[1];
        bool is_corner = false;
        real cx = 0.0, cy = 0.0;
        if (delta_t > 0.0) {
          if (delta_t > ln.length) {
            if (delta_t > (ln.length + pr)) continue;
            is_corner = true;
            cx = ln.p2[0]; cy = ln.p2[1];
          }
        } else if (delta_t > -pr) {
          is_corner = true;
          cx = ln.p1[0]; cy = ln.p1[1];
        } else {
          continue;
        }
        if (is_corner) {
          dr0 = pi0 - cx; dr1 = pi1 - cy;
          real norm = sqrt(dr0 * dr0 + dr1 * dr1);
          if (norm > pr) continue;
          real alpha = pr / norm - 1.0;
          x += alpha * dr0;
          y += alpha * dr1;
        } else {
          real sgn = sign_d(delta_s);
          real alpha = sgn * (pr - sgn * delta_s);
          x += alpha * ln.normal[0];
          y += alpha * ln.normal[1];
        }
      }
    } else if (sp.wall == MAVI_WALL_SLIPPERY && sp.geom == MAVI_GEOM_CIRCLE) {
      // :380-401 — calc_diff with the SYSTEM space (minimum image when the main space is periodic)
      real mr = sp.cr + pr;
      real dx = x - sp.cc[0], dy = y - sp.cc[1];
      if (p.periodic) {
        dx = min_image<true>(dx, p.half[0], p.size[0]);
        dy = min_image<true>(dy, p.half[1], p.size[1]);
      }
      real dr_2 = dx * dx + dy * dy;
      if (dr_2 <= mr * mr) {
        real dr_norm = sqrt(dr_2);
        real kk = sp.cr + pr - dr_norm;
        x = x + kk * dx / dr_norm;
        y = y + kk * dy / dr_norm;
      }
    }
  }
}

}  // namespace MAVI_NS
