// rings.cu — Mavi.Rings on the device (reference: src/rings/integration.jl, src/rings/rings.jl, src/rings/states.jl).
//
// The ring state stays RING-ORDERED (scalar idx = ring*n_max + p, src/rings/states.jl:137-139) so that one warp owns one
// ring; only particle indices are binned (index tiles) for the pair pass.  Per step (src/rings/integration.jl:522-543):
//   k_rings_pair : calc_forces! with the Rings pair law (:32-77) + calc_walls_forces!          (thread per particle)
//   k_rings_ring : update_cms! -> update_continuos_pos! -> calc_area -> springs -> area_forces! -> update! -> walls!
//                                                                                              (warp per ring)
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "handle.cuh"
#include "walls.cuh"

namespace MAVI_NS {

#define RINGS_TRY(h, expr)                                                                              \
  do {                                                                                                  \
    cudaError_t e_ = (expr);                                                                            \
    if (e_ != cudaSuccess) {                                                                            \
      (h)->set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #expr); \
      return MAVI_ERR_CUDA;                                                                             \
    }                                                                                                   \
  } while (0)

#define RINGS_LAUNCH(h, kernel, grid, block, ...)                 \
  do {                                                            \
    kernel<<<(grid), (block), 0, (h)->stream>>>(__VA_ARGS__);     \
    (h)->launches++;                                              \
  } while (0)

constexpr int RING_NMAX = 128;  // particles per ring the warp kernel stages in shared memory
constexpr int RING_WARPS = 4;

__device__ __forceinline__ int ring_type(const DevRings &r, int ring) { return r.types ? r.types[ring] : 0; }

// cross_prod(p1, p2) of calc_area (src/rings/integration.jl:103-116) with the reference's roundings: fl(fl(x1 y2) - fl(y1 x2)),
// no FMA contraction.  The shoelace sum of a ring far from the origin cancels catastrophically (terms ~ |r|^2, area ~ 1):
// a fused multiply-add here moves the area by ~|r|^2 ulp and the area force with it (4e-10 relative in a 1500-wide box).
__device__ __forceinline__ real cross_exact(real2 a, real2 b) { return add_rn(mul_rn(a.x, b.y), -mul_rn(a.y, b.x)); }

// idflag for ring-ordered slots: padding slots of shorter ring types are inactive (FixRingsIds, src/rings/states.jl:45-61)
// ... and, with VarRingsIds, every slot of an inactive ring (calc_active_ids!, src/rings/states.jl:200-223)
__global__ void k_rings_ids(const __grid_constant__ DevParams p, const unsigned char *__restrict__ ring_mask,
                            unsigned int *__restrict__ idflag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const int ring = i / p.rings.n_max, q = i - ring * p.rings.n_max;
  const int np = p.rings.num_particles[ring_type(p.rings, ring)];
  const bool on = q < np && (!ring_mask || ring_mask[ring]);
  idflag[i] = (unsigned int)i | (on ? 0u : MAVI_INACTIVE_BIT);
}

// update_cms! (src/rings/integration.jl:366-372) on its own: with sources / sinks it must run BEFORE they are processed
// (sinks test info.cms, spawned rings get their cms primed by process_source!), so the ring kernel skips it (prime_cms = -1)
template <bool PER>
__global__ void k_rings_cms(const __grid_constant__ DevParams p, const unsigned char *__restrict__ ring_mask,
                            const real2 *__restrict__ pos, const real2 *__restrict__ cont_pos, real2 *__restrict__ cms) {
  const DevRings &R = p.rings;
  const int ring = blockIdx.x * blockDim.x + threadIdx.x;
  if (ring >= R.num_rings || (ring_mask && !ring_mask[ring])) return;
  const int np = R.num_particles[ring_type(R, ring)];
  const real2 *src = (PER ? cont_pos : pos) + (size_t)ring * R.n_max;
  real sx = src[0].x, sy = src[0].y;
  for (int i = 1; i < np; i++) { sx += src[i].x; sy += src[i].y; }
  cms[ring] = make_real2(sx / np, sy / np);
}

// update_area_empty! (src/rings/sources.jl:191-225): an area is empty when no ACTIVE particle (ids as of the last
// update_ids!: idflag has not been refreshed yet) lies inside its bounding box grown by pad (is_inside, src/configs.jl:89-93).
// The reference's ChunksChecker walks the (one step old) chunk lists of the cells that intersect the padded box, which
// finds the same particles unless one moved farther than pad + a cell in a single step.
__global__ void k_rings_area_empty(const __grid_constant__ DevParams p, const unsigned int *__restrict__ idflag,
                                   const real2 *__restrict__ pos, const double *__restrict__ areas, int n_areas,
                                   int *__restrict__ empty) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n || (idflag[i] & MAVI_INACTIVE_BIT)) return;
  const double x = pos[i].x, y = pos[i].y;
  for (int a = 0; a < n_areas; a++) {
    const double *b = areas + 5 * a;
    const double pad = b[4];
    if (b[0] - pad <= x && x <= b[0] + b[2] + pad && b[1] - pad <= y && y <= b[1] + b[3] + pad) empty[a] = 0;
  }
}

// calc_interaction + calc_interaction_force (src/rings/integration.jl:32-77) summed over the stencil, + wall forces.
// Thread per particle in RING order (neighbouring threads are neighbours in space: their windows overlap in L1).  The
// candidates come from the index tiles: spos[s] / perm[s] = position / particle index of slot s, so a candidate costs two
// independent loads instead of the chain perm -> pos.  A particle whose cell rows r-1 .. r+1 lie inside one tile (30 of 32
// rows) walks its three column runs in ONE flat loop (tstart[r-1] .. tstart[r+2] of each column, same visiting order as the
// generic 9-cell walker, so the sums are bit-identical to it); the others take the generic walker.  In round 1 the nested
// per-cell loops kept 12 of 32 lanes busy (profiles/r01_ncu_szabo_rings.md).
template <bool PER>
__global__ void __launch_bounds__(TPB) k_rings_pair(const __grid_constant__ DevParams p, const int *__restrict__ tstart,
                                                    const int *__restrict__ perm, const real2 *__restrict__ spos,
                                                    const int *__restrict__ cell,
                                                    const unsigned int *__restrict__ idflag,
                                                    const real2 *__restrict__ pos, real2 *__restrict__ fpair,
                                                    int with_walls, const int *__restrict__ flags) {
  extern __shared__ real s_inter[];  // [num_types^2][7]
  __shared__ int s_rb[9][TPB], s_re[9][TPB];  // candidate runs of each thread
  const DevRings &R = p.rings;
  if (flags[FLAG_OVERFLOW]) return;  // the index tiles of this step overflowed: the step does not run (rings_run_steps re-runs it)
  for (int t = threadIdx.x; t < R.num_types * R.num_types * 7; t += blockDim.x) s_inter[t] = R.interaction[t];
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  real fx = 0.0, fy = 0.0;
  if (!(idflag[i] & MAVI_INACTIVE_BIT)) {
    const int ring = i / R.n_max;
    const int ring_lo = ring * R.n_max, ring_hi = ring_lo + R.n_max;
    const int ti = ring_type(R, ring);
    const int np = R.num_particles[ti];
    const bool typed = R.num_types > 1;
    const real *ic_row = s_inter + 7 * (ti * R.num_types);
    const real2 ri = pos[i];
    // branch-free: a candidate that does not interact (itself, out of range, bonded neighbour) contributes c = 0.  With
    // early returns the warp ran the force branch for ~every candidate anyway and kept 13 of 32 lanes busy.
    auto visit = [&](int j, real2 rj) {
      const real dx = min_image<PER>(ri.x - rj.x, p.half[0], p.size[0]);
      const real dy = min_image<PER>(ri.y - rj.y, p.half[1], p.size[1]);
      const real r2 = dist2_exact(dx, dy);
      const bool same = j >= ring_lo && j < ring_hi;
      const real *ic = typed ? ic_row + 7 * ring_type(R, j / R.n_max) : ic_row;
      const int diff = i > j ? i - j : j - i;
      // dist > dist_max | the particle itself | bonded neighbours inside the ring
      const bool skip = (r2 > ic[4]) | (diff == 0) | (same & ((diff == 1) | (diff == np - 1)));
      const real k = (r2 < ic[5]) ? ic[0] : (same ? 0.0 : ic[1]);  // no intra-ring attraction
      const real c = skip ? real(0.0) : k * (rsqrt(r2) - ic[6]);
      fx = fma(c, dx, fx);
      fy = fma(c, dy, fy);
    };
    if (p.num_cells == 0) {
      // chunks === nothing: all pairs over the active ids (src/integration.jl:197-224)
      for (int j = 0; j < p.n; j++)
        if (j != i && !(idflag[j] & MAVI_INACTIVE_BIT)) visit(j, __ldg(pos + j));
    } else {
      // Candidate RUNS of slots (start, end) in the visiting order of the generic 9-cell walker (columns col-1 .. col+1,
      // rows row-1 .. row+1 inside each; cell rows that are adjacent in a tile merge into one run), kept in shared memory
      // [run][thread], then ONE flat loop over all of them: every lane of the warp stays in the same loop whatever its cell
      // (a separate generic path for rows at tile edges ran with 2-3 active lanes and dominated the kernel).
      const int c = cell[i];
      const int col = div_rows(p, c), row = c - col * p.num_rows;
      const int tr = row / MAVI_TR, lr = row - tr * MAVI_TR;
      const int tid = threadIdx.x;
      int nruns = 0;
      if (lr >= 1 && lr <= MAVI_TR - 2 && row + 1 < p.num_rows) {  // rows row-1 .. row+1 inside one tile: one run per column
#pragma unroll
        for (int d = 0; d < 3; d++) {
          int c2 = col + d - 1;
          bool ok = true;
          if (c2 < 0) { ok = p.wrap_cols; c2 = p.num_cols - 1; }
          else if (c2 >= p.num_cols) { ok = p.wrap_cols; c2 = 0; }
          if (ok) {
            const int *tt = tstart + (size_t)(c2 * p.tpc + tr) * (MAVI_TR + 1) + lr - 1;
            s_rb[nruns][tid] = __ldg(tt);
            s_re[nruns][tid] = __ldg(tt + 3);
            nruns++;
          }
        }
      } else {
        for (int d = 0; d < 3; d++) {
          int c2 = col + d - 1;
          if (c2 < 0) { if (!p.wrap_cols) continue; c2 = p.num_cols - 1; }
          else if (c2 >= p.num_cols) { if (!p.wrap_cols) continue; c2 = 0; }
          int prev_end = -1;
          for (int dr = -1; dr <= 1; dr++) {
            int r2 = row + dr;
            if (r2 < 0) { if (!p.wrap_rows) continue; r2 = p.num_rows - 1; }
            else if (r2 >= p.num_rows) { if (!p.wrap_rows) continue; r2 = 0; }
            const int q = tq_of(p, c2, r2);
            const int jb = __ldg(tstart + q), je = __ldg(tstart + q + 1);
            if (jb == prev_end) s_re[nruns - 1][tid] = je;
            else { s_rb[nruns][tid] = jb; s_re[nruns][tid] = je; nruns++; }
            prev_end = je;
          }
        }
      }
      int k = 0, cur = 0, end = 0;
      if (nruns > 0) { cur = s_rb[0][tid]; end = s_re[0][tid]; }
      for (;;) {
        while (cur >= end && ++k < nruns) { cur = s_rb[k][tid]; end = s_re[k][tid]; }
        if (cur >= end) break;
        visit(__ldg(perm + cur), __ldg(spos + cur));
        cur++;
      }
    }
    if (with_walls && p.has_force_walls) wall_forces(p, ri.x, ri.y, fx, fy, ti);
  }
  fpair[i] = make_real2(fx, fy);
}

// neigh_update!(::ParticleNeighbors, ...) (src/rings/neighbors.jl:125-134) over the pair set of calc_forces!, as a gather:
// particle i counts (and lists) every j of its stencil with  dist < max_dist * tol  and (type == :all or other ring),
// max_dist = 2 particle_radius(interaction_cfg) = dist_eq (src/rings/integration.jl:40-41).  The reference appends in
// pair-enumeration order and its own test compares sorted lists (test/tests_rings/tests_general.jl:84-95): the device
// list is ascending.  Separate from k_rings_pair so that runs without neighbour tracking are untouched.
template <bool PER>
__global__ void __launch_bounds__(TPB) k_rings_neighbors(const __grid_constant__ DevParams p, const int *__restrict__ tstart,
                                                         const int *__restrict__ perm, const int *__restrict__ cell,
                                                         const unsigned int *__restrict__ idflag,
                                                         const real2 *__restrict__ pos, int type_all, real tol,
                                                         int *__restrict__ count, int *__restrict__ list,
                                                         const int *__restrict__ flags) {
  const DevRings &R = p.rings;
  if (flags[FLAG_OVERFLOW]) return;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  int cnt = 0;
  int mine[MAVI_NEIGH_MAX];
  if (!(idflag[i] & MAVI_INACTIVE_BIT)) {
    const int ring = i / R.n_max;
    const int ti = ring_type(R, ring);
    const real2 ri = pos[i];
    auto visit = [&](int j) {
      const int rj_ring = j / R.n_max;
      if (!type_all && rj_ring == ring) return;
      const real2 rj = __ldg(pos + j);
      const real dx = min_image<PER>(ri.x - rj.x, p.half[0], p.size[0]);
      const real dy = min_image<PER>(ri.y - rj.y, p.half[1], p.size[1]);
      const real dist = sqrt(dist2_exact(dx, dy));
      const real max_dist = R.interaction[7 * (ti * R.num_types + ring_type(R, rj_ring)) + 2];  // 2 * (dist_eq / 2)
      if (dist < max_dist * tol) {
        if (list && cnt < MAVI_NEIGH_MAX) {  // insertion keeps the list ascending
          int q = cnt;
          while (q > 0 && mine[q - 1] > j) { mine[q] = mine[q - 1]; --q; }
          mine[q] = j;
        }
        ++cnt;
      }
    };
    if (p.num_cells == 0) {
      for (int j = 0; j < p.n; j++)
        if (j != i && !(idflag[j] & MAVI_INACTIVE_BIT)) visit(j);
    } else {
      for_each_neighbor(p, tstart, cell[i], -1, [&](int s) {
        const int j = __ldg(perm + s);
        if (j != i) visit(j);
      });
    }
  }
  count[i] = cnt;
  if (list)
    for (int q = 0; q < MAVI_NEIGH_MAX; q++) list[(size_t)i * MAVI_NEIGH_MAX + q] = (q < cnt) ? mine[q] : -1;
}

// One warp per ring.  MODE 0: forces! only (constructor priming / mavi_calc_forces); MODE 1: full step.
template <bool PER, int MODE>
__global__ void __launch_bounds__(RING_WARPS * 32) k_rings_ring(
    const __grid_constant__ DevParams p, real2 *__restrict__ pos, real *__restrict__ pol,
    const real2 *__restrict__ fpair, real2 *__restrict__ force, real2 *__restrict__ cont_pos,
    real *__restrict__ areas, real2 *__restrict__ cms, const real *__restrict__ noise, unsigned long long step,
    int prime_cms, int *__restrict__ flags, const unsigned char *__restrict__ ring_mask) {
  __shared__ real2 s_pos[RING_WARPS][RING_NMAX];
  __shared__ real2 s_cont[RING_WARPS][RING_NMAX];
  __shared__ real2 s_vel[RING_WARPS][RING_NMAX];
  const DevRings &R = p.rings;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int ring = blockIdx.x * RING_WARPS + w;
  if (ring >= R.num_rings || flags[FLAG_OVERFLOW]) return;
  if (MODE == 1 && ring == 0 && lane == 0) flags[FLAG_STEPS] += 1;  // steps that really ran (device-side step counter)
  if (ring_mask && !ring_mask[ring]) return;  // get_rings_ids(state): active rings only
  const int t = ring_type(R, ring);
  const int np = R.num_particles[t];
  const int base = ring * R.n_max;
  real2 *sp = s_pos[w], *sc = s_cont[w], *sv = s_vel[w];
  for (int i = lane; i < R.n_max; i += 32) sp[i] = pos[base + i];
  __syncwarp();
  // ---- update_cms! (src/rings/integration.jl:366-372) runs FIRST in step!: it still sees last step's continuos_pos
  if (MODE == 1 && prime_cms >= 0 && lane == 0) {
    real sx, sy;
    if (PER) { sx = cont_pos[base].x; sy = cont_pos[base].y; }
    else { sx = sp[0].x; sy = sp[0].y; }
    for (int i = 1; i < np; i++) {
      const real2 c = PER ? cont_pos[base + i] : sp[i];
      sx += c.x;
      sy += c.y;
    }
    cms[ring] = make_real2(sx / np, sy / np);
  }
  __syncwarp();
  // ---- update_continuos_pos! (:118-138): sequential unwrap, same accumulation order as the reference
  if (PER) {
    if (lane == 0) {
      sc[0] = sp[0];
      for (int i = 1; i < np; i++) {
        const real dx = min_image<true>(sp[i].x - sp[i - 1].x, p.half[0], p.size[0]);
        const real dy = min_image<true>(sp[i].y - sp[i - 1].y, p.half[1], p.size[1]);
        sc[i] = make_real2(sc[i - 1].x + dx, sc[i - 1].y + dy);
      }
      for (int i = np; i < R.n_max; i++) sc[i] = sp[i];  // continuos_pos[:, ring] .= rings_pos[:, ring] first
    }
    __syncwarp();
    for (int i = lane; i < R.n_max; i += 32) cont_pos[base + i] = sc[i];
  } else {
    for (int i = lane; i < R.n_max; i += 32) sc[i] = sp[i];
    __syncwarp();
  }
  // ---- calc_area (:103-116), shoelace, sequential
  real area = 0.0;
  if (lane == 0) {
    for (int i = 0; i < np - 1; i++) area = add_rn(area, cross_exact(sc[i], sc[i + 1]));
    area = add_rn(area, cross_exact(sc[np - 1], sc[0]));
    area = area / 2.0;
    areas[ring] = area;
    if (MODE == 0 && prime_cms > 0) {  // constructor: update_cms! right after the first unwrap (src/rings/rings.jl:280-283)
      real sx = sc[0].x, sy = sc[0].y;
      for (int i = 1; i < np; i++) { sx += sc[i].x; sy += sc[i].y; }
      cms[ring] = make_real2(sx / np, sy / np);
    }
  }
  area = __shfl_sync(0xffffffffu, area, 0);
  // ---- forces! (:197-226): pair forces + springs (:79-97) + area_forces! (:140-195)
  const real k_spring = R.k_spring[t], l_spring = R.l_spring[t];
  const real k_area = R.k_area[t], p0 = R.p0[t];
  const real a0s = np * l_spring / p0;
  const real fmod_area = k_area * (area - a0s * a0s);
  const real vo = R.vo[t], mu = R.mobility[t];
  const real theta = pol[ring];
  real sn, cs;
  sincos(theta, &sn, &cs);
  auto spring = [&](int a, int b, real &ox, real &oy) {  // springs_force(p1 = a, p2 = b)
    const real dx = min_image<PER>(sp[a].x - sp[b].x, p.half[0], p.size[0]);
    const real dy = min_image<PER>(sp[a].y - sp[b].y, p.half[1], p.size[1]);
    const real dist = sqrt(dx * dx + dy * dy);
    const real c = (-k_spring * (dist - l_spring)) / dist;
    ox = c * dx;
    oy = c * dy;
  };
  for (int i = lane; i < np; i += 32) {
    const int nxt = (i == np - 1) ? 0 : i + 1, prv = (i == 0) ? np - 1 : i - 1;
    real2 F = fpair[base + i];
    real ax, ay, bx, by;
    spring(i, nxt, ax, ay);  // spring i: +f on its first particle
    spring(prv, i, bx, by);  // spring i-1: -f on its second particle
    F.x += ax; F.y += ay;
    F.x -= bx; F.y -= by;
    const real dx = min_image<PER>(sp[nxt].x - sp[prv].x, p.half[0], p.size[0]);
    const real dy = min_image<PER>(sp[nxt].y - sp[prv].y, p.half[1], p.size[1]);
    F.x -= fmod_area * (dy / 2);
    F.y -= fmod_area * (-dx / 2);
    force[base + i] = F;
    if (MODE == 1) {
      // update! (:300-351): overdamped active motion
      const real vx = vo * cs + mu * F.x, vy = vo * sn + mu * F.y;
      sv[i] = make_real2(vx, vy);
      real x = sp[i].x + vx * p.dt, y = sp[i].y + vy * p.dt;
      real dummy_vx = 0.0, dummy_vy = 0.0;
      const real pr = R.interaction[7 * (t * R.num_types + t) + 2] / 2.0;  // get_particle_radius of the ring type
      apply_walls<false>(p, x, y, dummy_vx, dummy_vy, pr);  // walls!(system), generic walls over the active ids
      pos[base + i] = make_real2(x, y);
    }
  }
  for (int i = np + lane; i < R.n_max; i += 32) force[base + i] = make_real2(0.0, 0.0);
  if (MODE == 1) {
    __syncwarp();
    if (lane == 0) {
      real vcx = 0.0, vcy = 0.0;
      for (int i = 0; i < np; i++) { vcx += sv[i].x; vcy += sv[i].y; }
      vcx /= np;
      vcy /= np;
      const real speed = sqrt(vcx * vcx + vcy * vcy);
      real cross_prod = 0.0;
      if (speed != 0.0) {
        cross_prod = (cs * vcy - sn * vcx) / speed;
        if (fabs(cross_prod) > 1.0) cross_prod = sign_d(cross_prod);
      }
      const real drot = R.rot_diff[t];
      real nz = 0.0;
      if (drot != 0.0)
        nz = (p.rng_mode == MAVI_RNG_HOST_NOISE) ? (noise ? noise[ring] : 0.0) : philox_normal(p.seed, (unsigned int)ring, step);
      pol[ring] = theta + (1.0 / R.relax_time[t] * asin(cross_prod) * p.dt + sqrt(2.0 * drot * p.dt) * nz);
    }
  }
}


// THREAD per ring (rings of up to RING_T_NMAX particles; the default for the small rings of the reference's fixtures and
// BASELINE config C4).  The warp-per-ring kernel above keeps 10 of 32 lanes busy on 10-particle rings and runs its four
// order-sensitive sums (cms, unwrap, shoelace area, mean velocity) on ONE lane while 31 wait: 200 us for 100 k rings
// (profiles/r01_ncu_szabo_rings.md).  Here a lane owns a ring and walks its particles in the reference's order, so all
// sums keep the reference's accumulation order and all 32 lanes work; consecutive lanes read consecutive rings (stride
// n_max positions: every 32-byte sector is used by two consecutive iterations out of L1).
constexpr int RING_T_NMAX = 32;
constexpr int RING_T_TPB = 128;

template <bool PER, int MODE>
__global__ void __launch_bounds__(RING_T_TPB) k_rings_ring_t(
    const __grid_constant__ DevParams p, real2 *__restrict__ pos, real *__restrict__ pol,
    const real2 *__restrict__ fpair, real2 *__restrict__ force, real2 *__restrict__ cont_pos,
    real *__restrict__ areas, real2 *__restrict__ cms, const real *__restrict__ noise, unsigned long long step,
    int prime_cms, int *__restrict__ flags, const unsigned char *__restrict__ ring_mask) {
  const DevRings &R = p.rings;
  const int ring = blockIdx.x * blockDim.x + threadIdx.x;
  if (ring >= R.num_rings || flags[FLAG_OVERFLOW]) return;
  if (MODE == 1 && ring == 0) flags[FLAG_STEPS] += 1;  // steps that really ran (device-side step counter)
  if (ring_mask && !ring_mask[ring]) return;  // get_rings_ids(state): active rings only
  const int t = ring_type(R, ring);
  const int np = R.num_particles[t];
  const int base = ring * R.n_max;
  // ---- update_cms! (src/rings/integration.jl:366-372) runs FIRST in step!: it still sees last step's continuos_pos
  if (MODE == 1 && prime_cms >= 0) {
    const real2 *src = PER ? cont_pos : pos;
    real sx = src[base].x, sy = src[base].y;
    for (int i = 1; i < np; i++) {
      const real2 c = src[base + i];
      sx += c.x;
      sy += c.y;
    }
    cms[ring] = make_real2(sx / np, sy / np);
  }
  // ---- update_continuos_pos! (:118-138) + calc_area (:103-116) + (constructor) update_cms!: one walk, sequential sums
  const real2 p0v = pos[base];
  real2 c_prev = p0v, r_prev = p0v;  // continuos_pos / rings_pos of particle i-1
  const real2 c0 = p0v;
  real area = 0.0, sx = p0v.x, sy = p0v.y;
  if (PER) cont_pos[base] = c0;
  for (int i = 1; i < np; i++) {
    const real2 ri = pos[base + i];
    real2 ci = ri;
    if (PER) {
      const real dx = min_image<true>(ri.x - r_prev.x, p.half[0], p.size[0]);
      const real dy = min_image<true>(ri.y - r_prev.y, p.half[1], p.size[1]);
      ci = make_real2(c_prev.x + dx, c_prev.y + dy);
      cont_pos[base + i] = ci;
    }
    area = add_rn(area, cross_exact(c_prev, ci));
    sx += ci.x;
    sy += ci.y;
    c_prev = ci;
    r_prev = ri;
  }
  if (PER)
    for (int i = np; i < R.n_max; i++) cont_pos[base + i] = pos[base + i];  // continuos_pos[:, ring] .= rings_pos[:, ring] first
  area = add_rn(area, cross_exact(c_prev, c0));
  area = area / 2.0;
  areas[ring] = area;
  if (MODE == 0 && prime_cms > 0) cms[ring] = make_real2(sx / np, sy / np);  // src/rings/rings.jl:280-283
  // ---- forces! (:197-226): pair forces + springs (:79-97) + area_forces! (:140-195); update! (:300-351); walls!
  const real k_spring = R.k_spring[t], l_spring = R.l_spring[t];
  const real k_area = R.k_area[t], p0 = R.p0[t];
  const real a0s = np * l_spring / p0;
  const real fmod_area = k_area * (area - a0s * a0s);
  const real vo = R.vo[t], mu = R.mobility[t];
  const real theta = pol[ring];
  real sn, cs;
  sincos(theta, &sn, &cs);
  const real pr = R.interaction[7 * (t * R.num_types + t) + 2] / 2.0;  // get_particle_radius of the ring type
  auto spring = [&](real2 a, real2 b, real &ox, real &oy) {  // springs_force(p1 = a, p2 = b)
    const real dx = min_image<PER>(a.x - b.x, p.half[0], p.size[0]);
    const real dy = min_image<PER>(a.y - b.y, p.half[1], p.size[1]);
    const real dist = sqrt(dx * dx + dy * dy);
    const real c = (-k_spring * (dist - l_spring)) / dist;
    ox = c * dx;
    oy = c * dy;
  };
  const real2 r_last = pos[base + np - 1];
  real2 rp = r_last, rc = p0v;  // previous / current particle (old positions: every new position is written after its last use)
  real bx, by;                  // spring (i-1, i): computed once, used as "next" of i-1 and "previous" of i
  spring(rp, rc, bx, by);
  const real lbx = bx, lby = by;  // spring (np-1, 0)
  real vcx = 0.0, vcy = 0.0;
  real2 new_first = p0v;
  for (int i = 0; i < np; i++) {
    const real2 rn = (i == np - 1) ? p0v : pos[base + i + 1];
    real2 F = fpair[base + i];
    real ax, ay;
    if (i == np - 1) { ax = lbx; ay = lby; }
    else spring(rc, rn, ax, ay);  // spring i: +f on its first particle
    F.x += ax; F.y += ay;
    F.x -= bx; F.y -= by;         // spring i-1: -f on its second particle
    const real dx = min_image<PER>(rn.x - rp.x, p.half[0], p.size[0]);
    const real dy = min_image<PER>(rn.y - rp.y, p.half[1], p.size[1]);
    F.x -= fmod_area * (dy / 2);
    F.y -= fmod_area * (-dx / 2);
    force[base + i] = F;
    if (MODE == 1) {
      const real vx = vo * cs + mu * F.x, vy = vo * sn + mu * F.y;
      vcx += vx;
      vcy += vy;
      real x = rc.x + vx * p.dt, y = rc.y + vy * p.dt;
      real dummy_vx = 0.0, dummy_vy = 0.0;
      apply_walls<false>(p, x, y, dummy_vx, dummy_vy, pr);  // walls!(system), generic walls over the active ids
      // pos[base+i-1 .. ] are no longer read: particle i-1 was "previous" for i only, held in rp; particle 0 is held in p0v
      if (i == 0) new_first = make_real2(x, y);
      else pos[base + i] = make_real2(x, y);
    }
    rp = rc;
    rc = rn;
    bx = ax;
    by = ay;
  }
  if (MODE == 1) pos[base] = new_first;
  for (int i = np; i < R.n_max; i++) force[base + i] = make_real2(0.0, 0.0);
  if (MODE == 1) {
    vcx /= np;
    vcy /= np;
    const real speed = sqrt(vcx * vcx + vcy * vcy);
    real cross_prod = 0.0;
    if (speed != 0.0) {
      cross_prod = (cs * vcy - sn * vcx) / speed;
      if (fabs(cross_prod) > 1.0) cross_prod = sign_d(cross_prod);
    }
    const real drot = R.rot_diff[t];
    real nz = 0.0;
    if (drot != 0.0)
      nz = (p.rng_mode == MAVI_RNG_HOST_NOISE) ? (noise ? noise[ring] : 0.0) : philox_normal(p.seed, (unsigned int)ring, step);
    pol[ring] = theta + (1.0 / R.relax_time[t] * asin(cross_prod) * p.dt + sqrt(2.0 * drot * p.dt) * nz);
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
template <typename T>
static int up(Handle *h, const T **dst, const T *src, size_t n) {
  T *d = nullptr;
  if (cudaMalloc((void **)&d, (n ? n : 1) * sizeof(T)) != cudaSuccess) {
    h->set_error("cudaMalloc failed (rings parameters)");
    return MAVI_ERR_CUDA;
  }
  h->allocs.push_back((void *)d);
  if (n && cudaMemcpy(d, src, n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {
    h->set_error("cudaMemcpy failed (rings parameters)");
    return MAVI_ERR_CUDA;
  }
  *dst = d;
  return MAVI_OK;
}

// per-type parameter arrays arrive as Float64 (MaviRingsParams) and are stored in the arithmetic type of this build
static int upr(Handle *h, const real **dst, const double *src, size_t n) {
  std::vector<real> tmp(src, src + n);
  return up(h, dst, tmp.data(), n);
}

static real sqrt_le_thr(real d) {
  if (!(d > 0)) return d < 0 ? real(-1) : real(0);
  if (std::isinf(d)) return d;  // "no cutoff"
  real x = d * d;
  while (std::sqrt(x) <= d) x = std::nextafter(x, INFINITY);
  while (std::sqrt(x) > d) x = std::nextafter(x, -INFINITY);
  return x;
}
static real sqrt_ge_thr(real d) {
  if (!(d > 0)) return real(0);
  if (std::isinf(d)) return d;
  real x = d * d;
  while (std::sqrt(x) >= d && x > 0) x = std::nextafter(x, -INFINITY);
  while (std::sqrt(x) < d) x = std::nextafter(x, INFINITY);
  return x;
}

int rings_lower(Handle *h, const MaviParams *mp) {
  const MaviRingsParams *r = mp->rings;
  if (!r || r->num_types < 1 || r->n_max < 1 || r->num_rings < 0 || r->num_rings * r->n_max != mp->n) {
    h->set_error("bad MaviRingsParams (n must equal num_rings * n_max)");
    return MAVI_ERR_BAD_PARAMS;
  }
  if (r->n_max > RING_NMAX) {
    h->set_error("rings with more than %d particles are not supported by the warp kernel", RING_NMAX);
    return MAVI_ERR_UNSUPPORTED;
  }
  const int nt = r->num_types;
  for (int t = 0; t < nt; t++)
    if (r->num_particles[t] < 3 || r->num_particles[t] > r->n_max) {
      h->set_error("ring type %d: num_particles must be in [3, n_max]", t + 1);
      return MAVI_ERR_BAD_PARAMS;
    }
  std::vector<real> inter((size_t)nt * nt * 7);
  for (int a = 0; a < nt; a++)
    for (int b = 0; b < nt; b++) {
      const double *s = r->interaction + 4 * (a * nt + b), *st = r->interaction + 4 * (b * nt + a);
      for (int q = 0; q < 4; q++)
        if (s[q] != st[q]) {
          h->set_error("InteractionMatrix must be symmetric: the reference applies +f/-f of ONE evaluation to both particles");
          return MAVI_ERR_UNSUPPORTED;
        }
      real *d = inter.data() + 7 * (a * nt + b);
      d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
      d[4] = sqrt_le_thr(s[3]);  // dist > dist_max  <=>  r2 > d[4]
      d[5] = sqrt_ge_thr(s[2]);  // dist < dist_eq   <=>  r2 < d[5]
      d[6] = 1.0 / s[2];
    }
  DevRings &R = h->p.rings;
  R.num_types = nt;
  R.n_max = r->n_max;
  R.num_rings = (int)r->num_rings;
  int st;
  if ((st = upr(h, &R.p0, r->p0, nt)) || (st = upr(h, &R.relax_time, r->relax_time, nt)) || (st = upr(h, &R.vo, r->vo, nt)) ||
      (st = upr(h, &R.mobility, r->mobility, nt)) || (st = upr(h, &R.rot_diff, r->rot_diff, nt)) ||
      (st = upr(h, &R.k_area, r->k_area, nt)) || (st = upr(h, &R.k_spring, r->k_spring, nt)) ||
      (st = upr(h, &R.l_spring, r->l_spring, nt)) || (st = up(h, &R.num_particles, r->num_particles, nt)) ||
      (st = up(h, &R.interaction, (const real *)inter.data(), inter.size())))
    return st;
  R.types = nullptr;
  h->r.np_h.assign(r->num_particles, r->num_particles + nt);
  h->r.types_h.clear();
  long long n_active = 0;
  if (r->types) {
    std::vector<int> t0((size_t)r->num_rings);
    for (long long i = 0; i < r->num_rings; i++) {
      if (r->types[i] < 1 || r->types[i] > nt) {
        h->set_error("ring %lld has type %d outside 1..%d", i + 1, r->types[i], nt);
        return MAVI_ERR_BAD_PARAMS;
      }
      t0[i] = r->types[i] - 1;
      n_active += r->num_particles[t0[i]];
    }
    h->r.types_h = t0;
    if ((st = up(h, &R.types, t0.data(), t0.size()))) return st;
  } else {
    n_active = r->num_rings * (long long)r->num_particles[0];
  }
  h->rings_n_active = (int)n_active;
  return MAVI_OK;
}

static int rings_alloc_tiles(Handle *h, int cap);

int rings_allocate(Handle *h) {
  DevArrays &a = h->a;
  RingsArrays &r = h->r;
  const size_t n = (size_t)h->p.n, nr = (size_t)h->p.rings.num_rings;
  auto al = [&](void **ptr, size_t bytes) -> int {
    if (cudaMalloc(ptr, bytes ? bytes : 16) != cudaSuccess) {
      h->set_error("cudaMalloc failed (rings state)");
      return MAVI_ERR_CUDA;
    }
    h->allocs.push_back(*ptr);
    return MAVI_OK;
  };
  int st;
  if ((st = al((void **)&a.pos[0], n * sizeof(real2))) || (st = al((void **)&a.force, n * sizeof(real2))) ||
      (st = al((void **)&a.force_old, n * sizeof(real2))) || (st = al((void **)&a.idflag, n * sizeof(unsigned int))) ||
      (st = al((void **)&a.cell, n * sizeof(int))) || (st = al((void **)&r.cont_pos, n * sizeof(real2))) ||
      (st = al((void **)&r.areas, nr * sizeof(real))) || (st = al((void **)&r.cms, nr * sizeof(real2))) ||
      (st = al((void **)&r.pol, nr * sizeof(real))))
    return st;
  h->p.n_count = h->rings_n_active;
  h->ns = n;
  RINGS_TRY(h, cudaMemsetAsync(a.force, 0, n * sizeof(real2), h->stream));
  RINGS_TRY(h, cudaMemsetAsync(r.cont_pos, 0, n * sizeof(real2), h->stream));
  if (h->p.num_cells > 0) {
    const long long ntiles = (long long)h->p.num_cols * ((h->p.num_rows + MAVI_TR - 1) / MAVI_TR);
    int cap = ((int)std::ceil(2.0 * h->rings_n_active / (real)ntiles + 16.0) + 15) / 16 * 16;
    return rings_alloc_tiles(h, cap);
  }
  return MAVI_OK;
}

static int rings_alloc_tiles(Handle *h, int cap) {
  DevParams &p = h->p;
  DevArrays &a = h->a;
  void *old[] = {a.tstart, a.perm, h->r.spos};
  for (void *q : old)
    if (q) {
      for (auto &x : h->allocs)
        if (x == q) x = nullptr;
      cudaFree(q);
    }
  p.tpc = (p.num_rows + MAVI_TR - 1) / MAVI_TR;
  p.nt = p.num_cols * p.tpc;
  p.cap = cap;
  p.n_active = 0;
  p.tail_base = 0;
  p.gcols = p.num_cols;
  p.ord_cols = p.num_cols;
  p.ord_col0 = 0;
  p.nt_ord = p.nt;
  mavi_magic_div((unsigned int)p.num_rows, &p.rows_mul, &p.rows_shr);
  mavi_magic_div((unsigned int)p.tpc, &p.tpc_mul, &p.tpc_shr);
  mavi_magic_div((unsigned int)p.num_cols, &p.cols_mul, &p.cols_shr);
  const size_t slots = (size_t)p.nt * cap;
  if (slots > 0x7ffffff0ull) {
    h->set_error("index tiles need too many slots");
    return MAVI_ERR_BAD_PARAMS;
  }
  a.tstart = nullptr;
  a.perm = nullptr;
  h->r.spos = nullptr;
  if (cudaMalloc((void **)&a.tstart, ((size_t)p.nt * (MAVI_TR + 1) + 1) * sizeof(int)) != cudaSuccess ||
      cudaMalloc((void **)&a.perm, (slots + 2) * sizeof(int)) != cudaSuccess ||
      cudaMalloc((void **)&h->r.spos, (slots + 2) * sizeof(real2)) != cudaSuccess) {
    h->set_error("cudaMalloc failed (index tiles)");
    return MAVI_ERR_CUDA;
  }
  h->allocs.push_back(a.tstart);
  h->allocs.push_back(a.perm);
  h->allocs.push_back(h->r.spos);
  return MAVI_OK;
}

// update_chunks! for Rings (src/rings/integration.jl:18-23): bin the active particle indices
int rings_bin(Handle *h) {
  DevParams &p = h->p;
  DevArrays &a = h->a;
  if (p.num_cells == 0) return MAVI_OK;
  if (int st0 = h->pending_out_of_grid()) return st0;
  for (int attempt = 0; attempt < 8; attempt++) {
    RINGS_TRY(h, cudaMemsetAsync(a.count, 0, ((size_t)p.num_cells + 2) * sizeof(int), h->stream));
    RINGS_TRY(h, cudaMemsetAsync(a.flags + 1, 0, (FLAG_STEPS - 1) * sizeof(int), h->stream));
    launch_build_index_tiles(h->ctx(), p, a.pos[0], a.idflag, a.cell, a.count, a.tstart, a.perm, a.flags, h->r.spos);
    int st = h->check_device_flags();
    if (st) return st;
    if (!h->flags_host[FLAG_OVERFLOW]) return MAVI_OK;
    int cap = ((int)std::ceil(h->flags_host[FLAG_MAXCOUNT] * 1.25 + 8.0) + 15) / 16 * 16;
    if ((st = rings_alloc_tiles(h, cap))) return st;
  }
  h->set_error("index tile capacity did not converge");
  return MAVI_ERR_CAPACITY;
}

static void launch_pair(Handle *h, bool with_walls) {
  const DevParams &p = h->p;
  DevArrays &a = h->a;
  const size_t smem = (size_t)p.rings.num_types * p.rings.num_types * 7 * sizeof(real);
  const int grid = (p.n + TPB - 1) / TPB;
  if (p.periodic) {
    k_rings_pair<true><<<grid, TPB, smem, h->stream>>>(p, a.tstart, a.perm, h->r.spos, a.cell, a.idflag, a.pos[0], a.force_old, (int)with_walls, a.flags);
  } else {
    k_rings_pair<false><<<grid, TPB, smem, h->stream>>>(p, a.tstart, a.perm, h->r.spos, a.cell, a.idflag, a.pos[0], a.force_old, (int)with_walls, a.flags);
  }
  h->launches++;
  RingsArrays &r = h->r;
  if (r.neigh_mode != MAVI_NEIGH_OFF && r.neigh_count) {  // neigh_clean! + neigh_update! of this forces! call
    int *list = r.neigh_mode == MAVI_NEIGH_LIST ? r.neigh_list : nullptr;
    if (p.periodic) {
      RINGS_LAUNCH(h, (k_rings_neighbors<true>), grid, TPB, p, a.tstart, a.perm, a.cell, a.idflag, a.pos[0], r.neigh_all, (real)r.neigh_tol, r.neigh_count, list, a.flags);
    } else {
      RINGS_LAUNCH(h, (k_rings_neighbors<false>), grid, TPB, p, a.tstart, a.perm, a.cell, a.idflag, a.pos[0], r.neigh_all, (real)r.neigh_tol, r.neigh_count, list, a.flags);
    }
  }
}

static void launch_ring(Handle *h, int mode, const real *noise, int prime_cms) {
  const DevParams &p = h->p;
  DevArrays &a = h->a;
  RingsArrays &r = h->r;
  const int grid = (p.rings.num_rings + RING_WARPS - 1) / RING_WARPS;
  if (grid == 0) return;
  const unsigned long long step = (unsigned long long)h->num_steps;
  static const bool warp_kernel = getenv("MAVI_RINGS_WARP_KERNEL") != nullptr;  // A/B switch (tests): the warp-per-ring kernel
  if (p.rings.n_max <= RING_T_NMAX && !warp_kernel) {
    const int gt = (p.rings.num_rings + RING_T_TPB - 1) / RING_T_TPB;
#define RING_CALL_T(PER, MODE) \
  RINGS_LAUNCH(h, (k_rings_ring_t<PER, MODE>), gt, RING_T_TPB, p, a.pos[0], r.pol, a.force_old, a.force, r.cont_pos, r.areas, r.cms, noise, step, prime_cms, a.flags, r.mask_dev)
    if (p.periodic) {
      if (mode) RING_CALL_T(true, 1); else RING_CALL_T(true, 0);
    } else {
      if (mode) RING_CALL_T(false, 1); else RING_CALL_T(false, 0);
    }
#undef RING_CALL_T
    return;
  }
#define RING_CALL(PER, MODE) \
  RINGS_LAUNCH(h, (k_rings_ring<PER, MODE>), grid, RING_WARPS * 32, p, a.pos[0], r.pol, a.force_old, a.force, r.cont_pos, r.areas, r.cms, noise, step, prime_cms, a.flags, r.mask_dev)
  if (p.periodic) {
    if (mode) RING_CALL(true, 1); else RING_CALL(true, 0);
  } else {
    if (mode) RING_CALL(false, 1); else RING_CALL(false, 0);
  }
#undef RING_CALL
}



// ---------------------------------------------------------------------------------------------------------
// invasions (src/rings/integration.jl:379-520): every InvasionsCfg.steps_to_update steps, the particles of a ring that lie
// inside the polygon of another ring whose centre of mass sits in the same or an adjacent RING chunk
// ---------------------------------------------------------------------------------------------------------
// update_continuos_pos! on its own (the ring kernel redoes it): ring_points of THIS step for the polygon tests
template <bool PER>
__global__ void k_rings_unwrap(const __grid_constant__ DevParams p, const unsigned char *__restrict__ ring_mask,
                               const real2 *__restrict__ pos, real2 *__restrict__ cont_pos) {
  const DevRings &R = p.rings;
  const int ring = blockIdx.x * blockDim.x + threadIdx.x;
  if (!PER || ring >= R.num_rings || (ring_mask && !ring_mask[ring])) return;
  const int np = R.num_particles[ring_type(R, ring)];
  const size_t base = (size_t)ring * R.n_max;
  real2 c_prev = pos[base], r_prev = c_prev;
  cont_pos[base] = c_prev;
  for (int i = 1; i < np; i++) {
    const real2 ri = pos[base + i];
    const real dx = min_image<true>(ri.x - r_prev.x, p.half[0], p.size[0]);
    const real dy = min_image<true>(ri.y - r_prev.y, p.half[1], p.size[1]);
    c_prev = make_real2(c_prev.x + dx, c_prev.y + dy);
    cont_pos[base + i] = c_prev;
    r_prev = ri;
  }
  for (int i = np; i < R.n_max; i++) cont_pos[base + i] = pos[base + i];
}

// update_chunks!(r_chunks) (src/rings/integration.jl:18-23): ring chunk of every active ring's centre of mass
__global__ void k_ring_cells(const __grid_constant__ DevParams p, const unsigned char *__restrict__ ring_mask,
                             const real2 *__restrict__ cms, int r_cols, int r_rows, double r_cl, double r_ch,
                             int *__restrict__ rcell, int *__restrict__ rcount, int *__restrict__ flags) {
  const int ring = blockIdx.x * blockDim.x + threadIdx.x;
  if (ring >= p.rings.num_rings) return;
  int c = -1;
  if (!ring_mask || ring_mask[ring]) {
    const double rowf = julia_div_pos(-(double)cms[ring].y + p.grid_bl[1] + p.grid_h, r_ch);
    const double colf = julia_div_pos((double)cms[ring].x - p.grid_bl[0], r_cl);
    if (fabs(rowf) < 2.0e9 && fabs(colf) < 2.0e9) {
      int row = (int)rowf + 1, col = (int)colf + 1;
      row -= (row == r_rows + 1) ? 1 : 0;
      col -= (col == r_cols + 1) ? 1 : 0;
      if (row >= 1 && row <= r_rows && col >= 1 && col <= r_cols) c = (row - 1) + r_rows * (col - 1);
    }
    if (c < 0) atomicOr(&flags[FLAG_ERR], ERRBIT_OUT_OF_GRID);  // BoundsError in the reference
    else atomicAdd(&rcount[c], 1);
  }
  rcell[ring] = c;
}
__global__ void k_ring_scatter(int num_rings, const int *__restrict__ rcell, const int *__restrict__ rstart,
                               int *__restrict__ rcount, int *__restrict__ rperm) {
  const int ring = blockIdx.x * blockDim.x + threadIdx.x;
  if (ring >= num_rings) return;
  const int c = rcell[ring];
  if (c >= 0) rperm[rstart[c] + atomicSub(&rcount[c], 1) - 1] = ring;
}

// point_line_intersect, src/rings/integration.jl:379-420 (ray from p towards +x against the segment l1-l2)
__device__ __forceinline__ bool point_line_intersect(real2 p, real2 l1, real2 l2) {
  const real dx = l2.x - l1.x, dy = l2.y - l1.y;
  const real ylo = dy < 0 ? l2.y : l1.y, yhi = dy < 0 ? l1.y : l2.y;
  if (dx == 0) return (l1.x > p.x) & (ylo < p.y && p.y < yhi);
  if (dy == 0) return false;
  const real c = dy * l1.x - dx * l1.y;
  const real x_inter = (c + dx * p.y) / dy;
  const real xlo = dx < 0 ? l2.x : l1.x, xhi = dx < 0 ? l1.x : l2.x;
  return (x_inter > p.x) & ((xlo < x_inter && x_inter < xhi) & (ylo < p.y && p.y < yhi));
}

// find_invasions! for every (particle of ring r1, ring r2 of the same / an adjacent ring chunk): thread per particle
__global__ void k_invasions(const __grid_constant__ DevParams p, const unsigned char *__restrict__ ring_mask,
                            const unsigned int *__restrict__ idflag, const real2 *__restrict__ pts,
                            const int *__restrict__ rcell, const int *__restrict__ rstart, const int *__restrict__ rperm,
                            int r_cols, int r_rows, int wrap, int *__restrict__ inv_n, int *__restrict__ inv_list, int inv_cap) {
  const DevRings &R = p.rings;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n || (idflag[i] & MAVI_INACTIVE_BIT)) return;
  const int r1 = i / R.n_max;
  const real2 pt = pts[i];
  auto test_ring = [&](int r2) {
    if (r2 == r1) return;
    const int np2 = R.num_particles[ring_type(R, r2)];
    const real2 *poly = pts + (size_t)r2 * R.n_max;
    int count = 0;
    real2 a = poly[0];
    for (int e = 0; e < np2; e++) {
      const real2 b = poly[e == np2 - 1 ? 0 : e + 1];
      count += point_line_intersect(pt, a, b) ? 1 : 0;
      a = b;
    }
    if (count & 1) {
      const int k = atomicAdd(inv_n, 1);
      if (k < inv_cap) {
        inv_list[3 * k] = r1;
        inv_list[3 * k + 1] = r2;
        inv_list[3 * k + 2] = i;
      }
    }
  };
  if (r_cols <= 0) {  // check_invasions!(system, ::Nothing): every pair of rings
    for (int r2 = 0; r2 < R.num_rings; r2++)
      if (!ring_mask || ring_mask[r2]) test_ring(r2);
    return;
  }
  const int c = rcell[r1];
  if (c < 0) return;
  const int col = c / r_rows, row = c - col * r_rows;
  for (int dc = -1; dc <= 1; dc++) {
    int c2 = col + dc;
    if (c2 < 0) { if (!wrap) continue; c2 = r_cols - 1; }
    else if (c2 >= r_cols) { if (!wrap) continue; c2 = 0; }
    for (int dr = -1; dr <= 1; dr++) {
      int r2 = row + dr;
      if (r2 < 0) { if (!wrap) continue; r2 = r_rows - 1; }
      else if (r2 >= r_rows) { if (!wrap) continue; r2 = 0; }
      const int cell = r2 + r_rows * c2;
      for (int q = rstart[cell]; q < rstart[cell + 1]; q++) test_ring(rperm[q]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// sources / sinks / variable ring count (src/rings/sources.jl, src/rings/states.jl:173-227)
// ---------------------------------------------------------------------------------------------------------
int rings_set_sources(Handle *h, const MaviSourceSink *list, int n, const unsigned char *ring_active, const double *draws,
                      long long n_draws) {
  if (h->p.dynamics != MAVI_DYN_RINGS || n < 0 || (n > 0 && !list)) {
    h->set_error("sources / sinks are a Mavi.Rings feature (RingsSystem source_cfg)");
    return MAVI_ERR_BAD_PARAMS;
  }
  RingsArrays &r = h->r;
  const DevRings &R = h->p.rings;
  const size_t nr = (size_t)R.num_rings;
  if (!ring_active && n > 0) {
    h->set_error("sources / sinks need a RingsState with active_state (VarRingsIds)");
    return MAVI_ERR_BAD_PARAMS;
  }
  r.sources.clear();
  r.var_rings = ring_active != nullptr;
  r.has_sinks = false;
  std::vector<double> areas;
  for (int k = 0; k < n; k++) {
    const MaviSourceSink &c = list[k];
    RingsArrays::Source o;
    o.kind = c.kind;
    if (c.kind == MAVI_SRC_SINK) {
      o.sink_geom = c.sink_geom;
      if (c.sink_geom == MAVI_GEOM_RECT) {
        o.sink[0] = c.sink_rect_bl[0]; o.sink[1] = c.sink_rect_bl[1]; o.sink[2] = c.sink_rect_len; o.sink[3] = c.sink_rect_h;
      } else if (c.sink_geom == MAVI_GEOM_CIRCLE) {
        o.sink[0] = c.sink_circ_center[0]; o.sink[1] = c.sink_circ_center[1]; o.sink[2] = c.sink_circ_radius;
      } else {
        h->set_error("SinkCfg geometry must be a rectangle or a circle (is_inside, src/configs.jl:89-93,165-168)");
        return MAVI_ERR_UNSUPPORTED;
      }
      r.has_sinks = true;
      r.sources.push_back(o);
      continue;
    }
    if (c.kind != MAVI_SRC_SOURCE || !c.spawn_pos || c.num_spawn_pos != R.n_max || c.size[0] < 1 || c.size[1] < 1) {
      h->set_error("SourceCfg: spawn_pos must hold n_max = %d points (add_ring! writes rings_pos[:, ring] .= pos) and size >= (1, 1)", R.n_max);
      return MAVI_ERR_BAD_PARAMS;
    }
    // Source ctor, src/rings/sources.jl:134-190
    const int nsp = c.num_spawn_pos;
    const double pad = c.pad;
    o.nsp = nsp; o.pad = pad; o.spawn_pol = c.spawn_pol;
    double min_x = c.spawn_pos[0], max_x = min_x, min_y = c.spawn_pos[1], max_y = min_y;
    for (int i = 1; i < nsp; i++) {
      const double x = c.spawn_pos[2 * i], y = c.spawn_pos[2 * i + 1];
      min_x = x < min_x ? x : min_x; max_x = x > max_x ? x : max_x;
      min_y = y < min_y ? y : min_y; max_y = y > max_y ? y : max_y;
    }
    const double bl_len = max_x - min_x + 2 * pad, bl_h = max_y - min_y + 2 * pad;
    o.nspawn = c.size[0] * c.size[1];
    o.first_area = (int)(areas.size() / 5);
    for (int i = 1; i <= c.size[0]; i++)
      for (int j = 1; j <= c.size[1]; j++) {
        const double bx = c.bottom_left[0] + ((i - 1) * bl_len + i * c.offset[0]);
        const double by = c.bottom_left[1] + ((j - 1) * bl_h + j * c.offset[1]);
        o.bbox.insert(o.bbox.end(), {bx, by, bl_len, bl_h});
        areas.insert(areas.end(), {bx, by, bl_len, bl_h, pad});
        const double dx = bx - min_x + pad, dy = by - min_y + pad;
        for (int q = 0; q < nsp; q++) {
          o.spawn.push_back(c.spawn_pos[2 * q] + dx);
          o.spawn.push_back(c.spawn_pos[2 * q + 1] + dy);
        }
      }
    r.sources.push_back(o);
  }
  r.mask_h.assign(nr, 1);
  r.uids_h.resize(nr);
  r.ids_h.resize(nr);
  for (size_t i = 0; i < nr; i++) {
    if (ring_active) r.mask_h[i] = ring_active[i] != 0;
    r.uids_h[i] = (long long)i + 1;
    r.ids_h[i] = (long long)i;
  }
  r.num_active = 0;
  r.draws.assign(draws ? draws : nullptr, draws ? draws + (n_draws > 0 ? n_draws : 0) : nullptr);
  r.draw_pos = 0;
  r.spawn_count = 0;
  r.n_areas = (int)(areas.size() / 5);
  auto al = [&](void **ptr, size_t bytes) -> int {
    if (*ptr) return MAVI_OK;
    if (cudaMalloc(ptr, bytes ? bytes : 16) != cudaSuccess) {
      h->set_error("cudaMalloc failed (sources)");
      return MAVI_ERR_CUDA;
    }
    h->allocs.push_back(*ptr);
    return MAVI_OK;
  };
  int st;
  if (r.var_rings) {
    if ((st = al((void **)&r.mask_dev, nr))) return st;
    RINGS_TRY(h, cudaMemcpyAsync(r.mask_dev, r.mask_h.data(), nr, cudaMemcpyHostToDevice, h->stream));
  }
  if (r.n_areas > 0) {
    r.areas_dev = nullptr;
    r.empty_dev = nullptr;
    if ((st = al((void **)&r.areas_dev, areas.size() * sizeof(double))) || (st = al((void **)&r.empty_dev, (size_t)r.n_areas * sizeof(int))))
      return st;
    RINGS_TRY(h, cudaMemcpyAsync(r.areas_dev, areas.data(), areas.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  }
  RINGS_TRY(h, cudaStreamSynchronize(h->stream));
  return MAVI_OK;
}

// calc_active_ids!, src/rings/states.jl:200-223: ids[1:num_active] from the mask (the particle ids follow on the device)
static void rings_calc_active_ids(RingsArrays &r) {
  long long q = 0;
  for (size_t i = 0; i < r.mask_h.size(); i++)
    if (r.mask_h[i]) r.ids_h[(size_t)q++] = (long long)i;
  r.num_active = q;
}

int rings_download_active(Handle *h, unsigned char *mask, long long *uids, long long *num_active) {
  if (h->p.dynamics != MAVI_DYN_RINGS) return MAVI_ERR_BAD_PARAMS;
  const RingsArrays &r = h->r;
  const size_t nr = (size_t)h->p.rings.num_rings;
  for (size_t i = 0; i < nr; i++) {
    if (mask) mask[i] = r.var_rings ? r.mask_h[i] : 1;
    if (uids) uids[i] = r.var_rings ? r.uids_h[i] : (long long)i + 1;
  }
  if (num_active) *num_active = r.var_rings ? r.num_active : (long long)nr;
  return MAVI_OK;
}

// update_sources! + update_ids! of one step (src/rings/integration.jl:353-358,525-527).  info.cms is current (k_rings_cms ran).
static int rings_process_sources(Handle *h) {
  DevParams &p = h->p;
  DevArrays &a = h->a;
  RingsArrays &r = h->r;
  const DevRings &R = p.rings;
  const size_t nr = (size_t)R.num_rings;
  std::vector<real2> cms;
  std::vector<int> empty((size_t)r.n_areas, 1);
  if (r.has_sinks) {
    cms.resize(nr);
    RINGS_TRY(h, cudaMemcpyAsync(cms.data(), r.cms, nr * sizeof(real2), cudaMemcpyDeviceToHost, h->stream));
  }
  if (r.n_areas > 0) {
    RINGS_TRY(h, cudaMemcpyAsync(r.empty_dev, empty.data(), empty.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    RINGS_LAUNCH(h, k_rings_area_empty, (p.n + TPB - 1) / TPB, TPB, p, a.idflag, a.pos[0], r.areas_dev, r.n_areas, r.empty_dev);
    RINGS_TRY(h, cudaMemcpyAsync(empty.data(), r.empty_dev, empty.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  }
  RINGS_TRY(h, cudaStreamSynchronize(h->stream));
  bool changed = false;
  for (auto &o : r.sources) {
    if (o.kind == MAVI_SRC_SINK) {
      // process_sink! (src/rings/sources.jl:245-251): `for ring_id in get_rings_ids(state)` is the view ids[1:num_active]
      // taken once; remove_ring! only clears the mask and lowers num_active (src/rings/states.jl:189-193)
      const long long nview = r.num_active;
      for (long long q = 0; q < nview; q++) {
        const size_t ring = (size_t)r.ids_h[(size_t)q];
        const double cx = cms[ring].x, cy = cms[ring].y;
        bool in;
        if (o.sink_geom == MAVI_GEOM_RECT)
          in = o.sink[0] <= cx && cx <= o.sink[0] + o.sink[2] && o.sink[1] <= cy && cy <= o.sink[1] + o.sink[3];
        else {
          const double ex = cx - o.sink[0], ey = cy - o.sink[1];
          in = ex * ex + ey * ey <= o.sink[2] * o.sink[2];
        }
        if (in) {
          r.mask_h[ring] = 0;
          r.num_active -= 1;
          changed = true;
        }
      }
      continue;
    }
    for (int k = 0; k < o.nspawn; k++) {
      if (!empty[(size_t)(o.first_area + k)]) continue;
      double pol = o.spawn_pol;
      if (std::isnan(pol)) {  // get_spawn_pol(::RandomPol) = rand(rng) * 2 pi, drawn before add_ring! looks for a slot
        double u;
        if (!r.draws.empty()) {
          if (r.draw_pos >= r.draws.size()) {
            h->set_error("spawn_draws exhausted after %zu draws", r.draws.size());
            return MAVI_ERR_BAD_PARAMS;
          }
          u = r.draws[r.draw_pos++];
        } else {  // production mode: splitmix64 keyed (seed, spawn count)
          unsigned long long z = p.seed + 0x9E3779B97F4A7C15ull * (++r.spawn_count);
          z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
          z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
          z ^= z >> 31;
          u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
        }
        pol = u * 2 * 3.141592653589793;
      }
      // add_ring! (src/rings/states.jl:173-187): first free slot
      size_t slot = nr;
      for (size_t i = 0; i < nr; i++)
        if (!r.mask_h[i]) { slot = i; break; }
      if (slot == nr) continue;  // "Space not found to add a new Ring!" — silently skipped by the reference
      const int nm = R.n_max;
      std::vector<real2> pts((size_t)nm);
      const double *sp = o.spawn.data() + 2 * (size_t)k * o.nsp;
      for (int q = 0; q < nm; q++) pts[(size_t)q] = make_real2((real)sp[2 * q], (real)sp[2 * q + 1]);
      const real polr = (real)pol;
      long long mx = r.uids_h[0];
      for (size_t i = 1; i < nr; i++) mx = r.uids_h[i] > mx ? r.uids_h[i] : mx;
      r.mask_h[slot] = 1;
      r.uids_h[slot] = mx + 1;
      r.num_active += 1;
      changed = true;
      // system.info.cms[ring_id] = sum(state.rings_pos[1:num_p, ring_id]) / num_p  (src/rings/sources.jl:240-243)
      const int np = r.np_h[r.types_h.empty() ? 0 : (size_t)r.types_h[slot]];
      real sx = pts[0].x, sy = pts[0].y;
      for (int q = 1; q < np; q++) { sx += pts[(size_t)q].x; sy += pts[(size_t)q].y; }
      const real2 cm = make_real2(sx / np, sy / np);
      RINGS_TRY(h, cudaMemcpy(a.pos[0] + slot * (size_t)nm, pts.data(), (size_t)nm * sizeof(real2), cudaMemcpyHostToDevice));
      RINGS_TRY(h, cudaMemcpy(r.pol + slot, &polr, sizeof(real), cudaMemcpyHostToDevice));
      RINGS_TRY(h, cudaMemcpy(r.cms + slot, &cm, sizeof(real2), cudaMemcpyHostToDevice));
    }
  }
  rings_calc_active_ids(r);  // update_ids!
  if (changed) {
    RINGS_TRY(h, cudaMemcpyAsync(r.mask_dev, r.mask_h.data(), nr, cudaMemcpyHostToDevice, h->stream));
    RINGS_LAUNCH(h, k_rings_ids, (p.n + TPB - 1) / TPB, TPB, p, r.mask_dev, a.idflag);
  }
  return MAVI_OK;
}

// InvasionsCfg(steps_to_update) + RingsIntCfg(r_chunks_cfg), src/rings/configs.jl:334-352
int rings_set_invasions(Handle *h, int steps_to_update, int r_cols, int r_rows) {
  if (h->p.dynamics != MAVI_DYN_RINGS || steps_to_update < 0 || r_cols < 0 || r_rows < 0 || (r_cols > 0) != (r_rows > 0)) {
    h->set_error("bad InvasionsCfg / r_chunks_cfg");
    return MAVI_ERR_BAD_PARAMS;
  }
  RingsArrays &r = h->r;
  const size_t nr = (size_t)h->p.rings.num_rings, ncell = (size_t)r_cols * r_rows;
  r.inv_steps = steps_to_update;
  r.inv_last_check = 0;
  r.r_cols = r_cols;
  r.r_rows = r_rows;
  auto al = [&](int **ptr, size_t count) -> int {
    *ptr = nullptr;
    if (cudaMalloc((void **)ptr, (count ? count : 1) * sizeof(int)) != cudaSuccess) {
      h->set_error("cudaMalloc failed (invasions)");
      return MAVI_ERR_CUDA;
    }
    h->allocs.push_back((void *)*ptr);
    return MAVI_OK;
  };
  int st;
  r.inv_cap = (int)(2 * (size_t)h->p.n + 64);
  if ((st = al(&r.rcell, nr)) || (st = al(&r.rcount, ncell + 2)) || (st = al(&r.rstart, ncell + 2)) || (st = al(&r.rperm, nr + 1)) ||
      (st = al(&r.rpart, ncell / 4096 + 4)) || (st = al(&r.inv_list, 3 * (size_t)r.inv_cap)) || (st = al(&r.inv_n, 1)))
    return st;
  RINGS_TRY(h, cudaMemsetAsync(r.inv_n, 0, sizeof(int), h->stream));
  return MAVI_OK;
}

// update_invasions! (src/rings/integration.jl:509-520) of the step about to run; info.cms must be current
static int rings_check_invasions(Handle *h) {
  DevParams &p = h->p;
  DevArrays &a = h->a;
  RingsArrays &r = h->r;
  const int nr = p.rings.num_rings;
  const int gr = (nr + TPB - 1) / TPB;
  r.inv_last_check = h->num_steps;
  if (p.periodic) RINGS_LAUNCH(h, (k_rings_unwrap<true>), gr, TPB, p, r.mask_dev, a.pos[0], r.cont_pos);
  RINGS_TRY(h, cudaMemsetAsync(r.inv_n, 0, sizeof(int), h->stream));
  if (r.r_cols > 0) {
    const int ncell = r.r_cols * r.r_rows;
    RINGS_TRY(h, cudaMemsetAsync(r.rcount, 0, ((size_t)ncell + 2) * sizeof(int), h->stream));
    RINGS_LAUNCH(h, k_ring_cells, gr, TPB, p, r.mask_dev, r.cms, r.r_cols, r.r_rows, h->grid_len_host / (double)r.r_cols,
                 (double)p.grid_h / r.r_rows, r.rcell, r.rcount, a.flags);
    launch_exclusive_scan(h->ctx(), r.rcount, r.rstart, r.rpart, ncell + 1);
    RINGS_LAUNCH(h, k_ring_scatter, gr, TPB, nr, r.rcell, r.rstart, r.rcount, r.rperm);
  }
  RINGS_LAUNCH(h, k_invasions, (p.n + TPB - 1) / TPB, TPB, p, r.mask_dev, a.idflag, p.periodic ? r.cont_pos : a.pos[0], r.rcell, r.rstart,
               r.rperm, r.r_cols, r.r_rows, p.spaces[0].wall == MAVI_WALL_PERIODIC ? 1 : 0, r.inv_n, r.inv_list, r.inv_cap);
  return MAVI_OK;
}

// info.invasions.list of the last check as (invasor ring, invaded ring, scalar particle id) triples, 0-based, sorted
int rings_download_invasions(Handle *h, long long *n, int *triples, long long cap) {
  RingsArrays &r = h->r;
  if (h->p.dynamics != MAVI_DYN_RINGS || !r.inv_n) {
    h->set_error("invasions are off (mavi_rings_set_invasions)");
    return MAVI_ERR_BAD_PARAMS;
  }
  int cnt = 0;
  RINGS_TRY(h, cudaMemcpyAsync(&cnt, r.inv_n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  int st = h->check_device_flags();
  if (st) return st;
  if (cnt > r.inv_cap) {
    h->set_error("%d invasions exceed the list capacity %d", cnt, r.inv_cap);
    return MAVI_ERR_CAPACITY;
  }
  if (n) *n = cnt;
  if (triples && cnt > 0) {
    std::vector<int> tmp(3 * (size_t)cnt);
    RINGS_TRY(h, cudaMemcpy(tmp.data(), r.inv_list, tmp.size() * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<size_t> order((size_t)cnt);
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t x, size_t y) {
      for (int q = 0; q < 3; q++)
        if (tmp[3 * x + q] != tmp[3 * y + q]) return tmp[3 * x + q] < tmp[3 * y + q];
      return false;
    });
    const long long m = cnt < cap ? cnt : cap;
    for (long long i = 0; i < m; i++)
      for (int q = 0; q < 3; q++) triples[3 * i + q] = tmp[3 * order[(size_t)i] + q];
  }
  return MAVI_OK;
}

// RingsSystem ctor tail (src/rings/rings.jl:280-288): ids, continuos_pos, cms, chunks, forces!
int rings_upload_finish(Handle *h) {
  DevParams &p = h->p;
  DevArrays &a = h->a;
  RingsArrays &r = h->r;
  const size_t n = (size_t)p.n;
  RINGS_TRY(h, cudaMemcpyAsync(a.pos[0], a.st_pos, n * sizeof(real2), cudaMemcpyDeviceToDevice, h->stream));
  RINGS_TRY(h, cudaMemcpyAsync(r.pol, a.st_ang, (size_t)p.rings.num_rings * sizeof(real), cudaMemcpyDeviceToDevice, h->stream));
  if (r.var_rings) rings_calc_active_ids(r);  // RingsState ctor: update_ids!(state), src/rings/states.jl:121
  RINGS_LAUNCH(h, k_rings_ids, (p.n + TPB - 1) / TPB, TPB, p, r.mask_dev, a.idflag);
  RINGS_TRY(h, cudaMemcpyAsync(a.st_id, a.idflag, n * sizeof(unsigned int), cudaMemcpyDeviceToDevice, h->stream));
  if (p.n_spaces == 1) launch_check_inside(h->ctx(), p, a);
  int st = h->check_device_flags();
  if (st) return st;
  if ((st = rings_bin(h))) return st;
  launch_pair(h, false);
  launch_ring(h, 0, nullptr, 1);
  return h->check_device_flags();
}

int rings_calc_forces(Handle *h) {
  int st = rings_bin(h);
  if (st) return st;
  launch_pair(h, true);
  launch_ring(h, 0, nullptr, 0);
  return h->check_device_flags();
}

// One Rings step!, enqueued without host synchronisation.  If the index tiles overflow, the step latches FLAG_OVERFLOW and
// it and every later enqueued step turn into no-ops (the ring-ordered state is untouched); FLAG_STEPS counts the steps
// that really ran and Handle::run_steps grows the tiles and re-runs the rest.
int rings_step(Handle *h, const real *noise_dev) {
  DevParams &p = h->p;
  DevArrays &a = h->a;
  RingsArrays &r = h->r;
  if (r.var_rings) {
    // step! with a variable ring set (src/rings/integration.jl:522-527): update_cms!; update_sources!; update_ids! — one
    // small host round trip (add_ring! / remove_ring! are sequential by definition) — then the usual step, binning
    // synchronously (an index-tile overflow must not replay the sources)
    const int gr = (p.rings.num_rings + TPB - 1) / TPB;
    if (p.periodic) RINGS_LAUNCH(h, (k_rings_cms<true>), gr, TPB, p, r.mask_dev, a.pos[0], r.cont_pos, r.cms);
    else RINGS_LAUNCH(h, (k_rings_cms<false>), gr, TPB, p, r.mask_dev, a.pos[0], r.cont_pos, r.cms);
    int st = rings_process_sources(h);
    if (st) return st;
    if ((st = rings_bin(h))) return st;
    if (r.inv_steps > 0 && h->num_steps - r.inv_last_check >= r.inv_steps && (st = rings_check_invasions(h))) return st;
    launch_pair(h, true);
    launch_ring(h, 1, noise_dev, -1);  // -1: update_cms! already done
    h->num_steps += 1;
    h->time += h->dt_host;
    return MAVI_OK;
  }
  if (p.num_cells > 0) {  // update_chunks_all! (after update_cms!, which only reads last step's continuos_pos)
    RINGS_TRY(h, cudaMemsetAsync(a.count, 0, ((size_t)p.num_cells + 2) * sizeof(int), h->stream));
    launch_build_index_tiles(h->ctx(), p, a.pos[0], a.idflag, a.cell, a.count, a.tstart, a.perm, a.flags, h->r.spos);
  }
  int prime = 0;
  if (r.inv_steps > 0 && h->num_steps - r.inv_last_check >= r.inv_steps) {
    // a check step: update_cms! on its own first (the ring chunks bin THIS step's cms), then the polygon tests
    const int gr = (p.rings.num_rings + TPB - 1) / TPB;
    if (p.periodic) RINGS_LAUNCH(h, (k_rings_cms<true>), gr, TPB, p, r.mask_dev, a.pos[0], r.cont_pos, r.cms);
    else RINGS_LAUNCH(h, (k_rings_cms<false>), gr, TPB, p, r.mask_dev, a.pos[0], r.cont_pos, r.cms);
    int st = rings_check_invasions(h);
    if (st) return st;
    prime = -1;
  }
  launch_pair(h, true);
  launch_ring(h, 1, noise_dev, prime);
  h->num_steps += 1;  // src/rings/integration.jl:541-542
  h->time += h->dt_host;
  return MAVI_OK;
}

// after a latched overflow: larger index tiles, latch cleared
int rings_grow_tiles(Handle *h) {
  const int cap = ((int)std::ceil(h->flags_host[FLAG_MAXCOUNT] * 1.25 + 8.0) + 15) / 16 * 16;
  int st = rings_alloc_tiles(h, cap > h->p.cap ? cap : h->p.cap + 16);
  if (st) return st;
  RINGS_TRY(h, cudaMemsetAsync(h->a.flags + FLAG_OVERFLOW, 0, sizeof(int), h->stream));
  h->n_rebuilds++;
  return MAVI_OK;
}

int rings_download_state(Handle *h, void *pos, void *second) {
  const DevParams &p = h->p;
  if (pos) RINGS_TRY(h, cudaMemcpyAsync(pos, h->a.pos[0], (size_t)p.n * sizeof(real2), cudaMemcpyDeviceToHost, h->stream));
  if (second) RINGS_TRY(h, cudaMemcpyAsync(second, h->r.pol, (size_t)p.rings.num_rings * sizeof(real), cudaMemcpyDeviceToHost, h->stream));
  return h->check_device_flags();
}

int rings_download_forces(Handle *h, void *forces) {
  RINGS_TRY(h, cudaMemcpyAsync(forces, h->a.force, (size_t)h->p.n * sizeof(real2), cudaMemcpyDeviceToHost, h->stream));
  return h->check_device_flags();
}

int rings_download_info(Handle *h, void *areas, void *cms, void *cont_pos) {
  const DevParams &p = h->p;
  const size_t nr = (size_t)p.rings.num_rings;
  if (areas) RINGS_TRY(h, cudaMemcpyAsync(areas, h->r.areas, nr * sizeof(real), cudaMemcpyDeviceToHost, h->stream));
  if (cms) RINGS_TRY(h, cudaMemcpyAsync(cms, h->r.cms, nr * sizeof(real2), cudaMemcpyDeviceToHost, h->stream));
  if (cont_pos)
    RINGS_TRY(h, cudaMemcpyAsync(cont_pos, p.periodic ? h->r.cont_pos : h->a.pos[0], (size_t)p.n * sizeof(real2), cudaMemcpyDeviceToHost, h->stream));
  return h->check_device_flags();
}

// cell_of_particle / counts / CSR lists of the index tiles
__global__ void k_rings_cells_out(const __grid_constant__ DevParams p, const unsigned int *__restrict__ idflag,
                                  const int *__restrict__ cell, int *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p.n) out[i] = (idflag[i] & MAVI_INACTIVE_BIT) ? -1 : cell[i];
}

int rings_download_cells(Handle *h, int *cell_of_particle, int *counts, int *start, int *ids) {
  const DevParams &p = h->p;
  DevArrays &a = h->a;
  if (cell_of_particle) {
    RINGS_LAUNCH(h, k_rings_cells_out, (p.n + TPB - 1) / TPB, TPB, p, a.idflag, a.cell, a.st_cell);
    RINGS_TRY(h, cudaMemcpyAsync(cell_of_particle, a.st_cell, (size_t)p.n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  }
  if (counts || start || ids) {
    std::vector<int> ts((size_t)p.nt * (MAVI_TR + 1) + 1), perm;
    RINGS_TRY(h, cudaMemcpyAsync(ts.data(), a.tstart, ts.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (ids) {
      perm.resize((size_t)p.nt * p.cap);
      RINGS_TRY(h, cudaMemcpyAsync(perm.data(), a.perm, perm.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    }
    RINGS_TRY(h, cudaStreamSynchronize(h->stream));
    int acc = 0;
    for (int c = 0; c < p.num_cells; c++) {
      const int col = c / p.num_rows, row = c - col * p.num_rows;
      const int tr = row / MAVI_TR;
      const size_t q = (size_t)(col * p.tpc + tr) * (MAVI_TR + 1) + (row - tr * MAVI_TR);
      const int b = ts[q], e = ts[q + 1];
      if (counts) counts[c] = e - b;
      if (start) start[c] = acc;
      if (ids)
        for (int j = b; j < e; j++) ids[acc + (j - b)] = perm[j];
      acc += e - b;
    }
    if (start) start[p.num_cells] = acc;
  }
  return h->check_device_flags();
}

// NeighborsCfg (src/rings/neighbors.jl:11-15) for the particle contact lists; mode MAVI_NEIGH_OFF frees nothing and
// simply stops the tracking.  Takes effect at the next forces! (upload, mavi_calc_forces, mavi_step).
int rings_set_neighbors(Handle *h, int mode, int type_all, double tol) {
  if (h->p.dynamics != MAVI_DYN_RINGS) {
    h->set_error("contact lists are a Mavi.Rings feature (RingsSystem p_neighbors_cfg)");
    return MAVI_ERR_BAD_PARAMS;
  }
  if (mode < MAVI_NEIGH_OFF || mode > MAVI_NEIGH_LIST || !(tol > 0.0)) {
    h->set_error("bad neighbour mode / tol");
    return MAVI_ERR_BAD_PARAMS;
  }
  RingsArrays &r = h->r;
  const size_t n = (size_t)h->p.n;
  auto al = [&](int **ptr, size_t count) -> int {
    if (*ptr) return MAVI_OK;
    if (cudaMalloc((void **)ptr, (count ? count : 1) * sizeof(int)) != cudaSuccess) {
      h->set_error("cudaMalloc failed (neighbour lists)");
      return MAVI_ERR_CUDA;
    }
    h->allocs.push_back((void *)*ptr);
    return MAVI_OK;
  };
  int st;
  if (mode != MAVI_NEIGH_OFF && (st = al(&r.neigh_count, n))) return st;
  if (mode == MAVI_NEIGH_LIST && (st = al(&r.neigh_list, n * MAVI_NEIGH_MAX))) return st;
  if (mode != MAVI_NEIGH_OFF) RINGS_TRY(h, cudaMemsetAsync(r.neigh_count, 0, (n ? n : 1) * sizeof(int), h->stream));
  if (mode == MAVI_NEIGH_LIST) RINGS_TRY(h, cudaMemsetAsync(r.neigh_list, 0xff, (n ? n : 1) * MAVI_NEIGH_MAX * sizeof(int), h->stream));
  r.neigh_mode = mode;
  r.neigh_all = type_all ? 1 : 0;
  r.neigh_tol = tol;
  return MAVI_OK;
}

// get_neigh_count / get_neigh_list (src/rings/neighbors.jl:58-62) for every particle slot
int rings_download_neighbors(Handle *h, int *count, int *list) {
  RingsArrays &r = h->r;
  if (r.neigh_mode == MAVI_NEIGH_OFF || !r.neigh_count) {
    h->set_error("neighbour tracking is off (mavi_rings_set_neighbors)");
    return MAVI_ERR_BAD_PARAMS;
  }
  if (list && r.neigh_mode != MAVI_NEIGH_LIST) {
    h->set_error("only_count mode keeps no lists");
    return MAVI_ERR_BAD_PARAMS;
  }
  const size_t n = (size_t)h->p.n;
  std::vector<int> tmp;
  int *cdst = count;
  if (!cdst && list) { tmp.resize(n); cdst = tmp.data(); }
  if (cdst) RINGS_TRY(h, cudaMemcpyAsync(cdst, r.neigh_count, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (list) RINGS_TRY(h, cudaMemcpyAsync(list, r.neigh_list, n * MAVI_NEIGH_MAX * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  int st = h->check_device_flags();
  if (st) return st;
  if (list)
    for (size_t i = 0; i < n; i++)
      if (cdst[i] > MAVI_NEIGH_MAX) {  // the reference writes past its 15-entry table here (BoundsError)
        h->set_error("particle %zu has %d contacts, more than the %d a list holds (BoundsError in the reference)", i, cdst[i], MAVI_NEIGH_MAX);
        return MAVI_ERR_CAPACITY;
      }
  return MAVI_OK;
}

}  // namespace MAVI_NS
