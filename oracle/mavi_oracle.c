/* mavi_oracle.c — CPU oracle (plain C) for the Mavi.jl per-step hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * PARITY UNPINNED (see mavi_oracle.h): Julia cannot run in this image and the reference's tests hold
 * no portable golden vectors for this path; every function below cites the reference lines it restates.
 * All `file:line` citations are relative to /root/reference/.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared   (no FMA contraction: Julia does not contract)
 */
#include "mavi_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

struct OrSystem {
  MaviParams p;
  MaviLine *lines[MAVI_MAX_SPACES];
  MaviRingsParams rp; /* deep copy */
  int64_t n;
  double *pos;    /* [2n] */
  double *second; /* vel [2n] | pol_angle [n] | ring pol [num_rings] */
  uint8_t *mask;
  int64_t *ids;   /* active particle ids, 0-based, ascending (get_particles_ids) */
  int64_t n_ids;
  int32_t nthreads;
  double **forces; /* forces[t][2n]  (src/systems.jl:89-97) */
  /* Chunks (src/chunks.jl:10-25) */
  int has_chunks;
  int64_t num_cols, num_rows, nc;
  double cl, ch;
  int64_t *chunk_particles; /* [nc][rows][cols], first index fastest */
  int64_t *num_in_chunk;    /* [rows][cols], row fastest */
  int32_t *neigh;           /* [cells][4] linear cell ids (row + rows*col) */
  int8_t *neigh_n;
  /* RingsInfo (src/rings/rings.jl:118-128) */
  double *cont_pos, *areas, *cms;
  /* ParticleNeighbors (src/rings/neighbors.jl:52-56, src/rings/rings.jl:143-158): one slice (the thread slices of the
   * Threaded device are summed by neigh_sum_buffers, :76-100; here updates go to the main slice under a lock) */
  int32_t pn_mode, pn_all;
  double pn_tol;
  int32_t *pn_count; /* [n] */
  int32_t *pn_list;  /* [n][15], pair-enumeration order like the reference */
  int32_t pn_overflow;
  /* VarRingsIds (src/rings/states.jl:24-43) + sources / sinks (src/rings/sources.jl) */
  int var_rings;           /* RingsState built with active_state */
  uint8_t *ring_mask;      /* rings_ids.mask */
  int64_t *ring_ids;       /* rings_ids.ids[1:num_active] as of the last calc_active_ids! (0-based) */
  int64_t *ring_uids;
  int64_t num_active;      /* rings_ids.num_active (add_ring! / remove_ring! change it at once) */
  struct OrSource *src;
  int32_t nsrc;
  double *draws;
  int64_t n_draws, draw_pos;
  /* invasions (src/rings/integration.jl:379-520): InvasionsCfg.steps_to_update, InvasionsInfo, ring-level Chunks on the cms */
  int32_t inv_steps;         /* 0 = off */
  int64_t inv_last_check;
  int32_t *inv_list;         /* [n][3] = invasor ring, invaded ring, scalar particle id (0-based) */
  int64_t inv_n, inv_cap;
  int64_t r_cols, r_rows, r_nc; /* r_chunks (0 columns: check_invasions!(system, ::Nothing)) */
  double r_cl, r_ch;
  int64_t *r_particles, *r_num;
  int32_t *r_neigh;
  int8_t *r_neigh_n;
  int64_t num_steps;
  double time;
  char err[256];
};

/* one Source / Sink (src/rings/sources.jl:127-190, :232-251) */
typedef struct OrSource {
  int kind, nspawn, nsp;
  double pad, spawn_pol;
  double *bbox;       /* [nspawn][4]: bottom_left.x, .y, length, height */
  double *spawn;      /* bbox_spawn_pos: [nspawn][nsp][2] */
  uint8_t *is_empty;
  int32_t **cells;    /* ChunksChecker.ids: per spawn area, linear cell ids (row-1) + rows*(col-1) in col-major scan order */
  int32_t *ncells;
  int sink_geom;
  double sink[5];     /* rect: bl.x, bl.y, length, height | circle: c.x, c.y, radius */
} OrSource;

/* ------------------------------------------------------------------ helpers */

static double *dup_d(const double *src, int64_t n) {
  double *d = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  if (src) memcpy(d, src, sizeof(double) * (size_t)n);
  return d;
}

static int is_periodic(const OrSystem *s) {
  /* calc_diff dispatch: SpaceCfg{PeriodicWalls, RectangleCfg}; ManyWalls -> first pair (src/integration.jl:43-52) */
  return s->p.spaces[0].wall == MAVI_WALL_PERIODIC && s->p.spaces[0].geom == MAVI_GEOM_RECT;
}

/* Base.div(x::Float64, y::Float64) = round((x - rem(x, y)) / y)  (Julia Base, not under /root/reference;
 * call sites src/chunks.jl:129-130).  rem == C fmod, round == rint (ties to even). */
double mor_julia_div(double x, double y) { return rint((x - fmod(x, y)) / y); }

/* calc_diff, src/integration.jl:38-52 */
void mor_calc_diff(const OrSystem *s, const double *r1, const double *r2, double *dr) {
  dr[0] = r1[0] - r2[0];
  dr[1] = r1[1] - r2[1];
  if (is_periodic(s)) {
    const double sz[2] = {s->p.spaces[0].rect_len, s->p.spaces[0].rect_h};
    for (int d = 0; d < 2; d++) {
      /* dr - (abs(dr) > size/2) * copysign(size, dr) */
      double flag = fabs(dr[d]) > (sz[d] / 2) ? 1.0 : 0.0;
      dr[d] = dr[d] - flag * copysign(sz[d], dr[d]);
    }
  }
}

/* potential_force(dr, dist, potential): HarmTrunc src/configs.jl:354-368, LenJones :389-397 */
void mor_potential_force(int32_t kind, const double *par, const double *dr, double dist, double *f) {
  if (kind == MAVI_POT_HARMTRUNC) {
    const double k_rep = par[0], k_atr = par[1], dist_eq = par[2], dist_max = par[3];
    if (dist > dist_max) {
      f[0] = 0.0; f[1] = 0.0;
      return;
    }
    double fmod_;
    if (dist < dist_eq) fmod_ = -k_rep * (dist / dist_eq - 1);
    else fmod_ = -k_atr * (dist / dist_eq - 1);
    double c = fmod_ / dist;
    f[0] = c * dr[0]; f[1] = c * dr[1];
  } else {
    const double sigma = par[0], epsilon = par[1];
    double fmod_ = 4 * epsilon * (12 * pow(sigma, 12) / pow(dist, 13) - 6 * pow(sigma, 6) / pow(dist, 7));
    double c = fmod_ / dist;
    f[0] = c * dr[0]; f[1] = c * dr[1];
  }
}

/* calc_interaction(::SzaboCfg), src/integration.jl:68-87.  dr is NOT normalised (sic). */
void mor_szabo_interaction(const double *par, const double *dr, double *f) {
  const double k_rep = par[3], k_adh = par[4], r_eq = par[5], r_max = par[6];
  double dist = sqrt(dr[0] * dr[0] + dr[1] * dr[1]);
  if (dist > r_max) {
    f[0] = 0.0; f[1] = 0.0;
    return;
  }
  double f_mod;
  if (dist > r_eq) f_mod = k_adh / r_eq;
  else f_mod = k_rep / (r_max - r_eq);
  double r = dist - r_eq;
  double c = -f_mod * r;
  f[0] = c * dr[0]; f[1] = c * dr[1];
}

/* calc_interaction(::RunTumbleCfg), src/integration.jl:89-109 (WCA) */
void mor_rtp_interaction(const double *par, const double *dr, double *f) {
  const double sigma = par[1], epsilon = par[2];
  double dist = sqrt(dr[0] * dr[0] + dr[1] * dr[1]);
  double cutoff = pow(2.0, 1.0 / 6.0) * sigma;
  if (dist > cutoff) {
    f[0] = 0.0; f[1] = 0.0;
    return;
  }
  double fmod_ = -4 * epsilon * (-12 * pow(sigma, 12) / pow(dist, 13) + 6 * pow(sigma, 6) / pow(dist, 7));
  double c = fmod_ / dist;
  f[0] = c * dr[0]; f[1] = c * dr[1];
}

/* ring helpers, src/rings/states.jl:137-146, :195-198 */
static inline int64_t ring_of(const OrSystem *s, int64_t idx0) { return idx0 / s->rp.n_max; }
static inline int32_t ring_type(const OrSystem *s, int64_t ring0) { return s->rp.types ? s->rp.types[ring0] - 1 : 0; }
static inline int32_t ring_np(const OrSystem *s, int64_t ring0) { return s->rp.num_particles[ring_type(s, ring0)]; }

/* get_rings_ids(state): FixRingsIds -> every ring; VarRingsIds -> ids[1:num_active] (src/rings/states.jl:36-40) */
static inline int64_t n_ring_ids(const OrSystem *s) { return s->var_rings ? s->num_active : s->rp.num_rings; }
static inline int64_t ring_id_at(const OrSystem *s, int64_t q) { return s->var_rings ? s->ring_ids[q] : q; }

/* Rings calc_interaction + calc_interaction_force, src/rings/integration.jl:32-77 */
static void rings_interaction(const OrSystem *s, int64_t i, int64_t j, double *f) {
  int64_t ri = ring_of(s, i), rj = ring_of(s, j);
  const double *ic = s->rp.interaction + 4 * ((int64_t)ring_type(s, ri) * s->rp.num_types + ring_type(s, rj));
  double dr[2];
  mor_calc_diff(s, s->pos + 2 * i, s->pos + 2 * j, dr);
  double dist = sqrt(dr[0] * dr[0] + dr[1] * dr[1]);
  const double k_rep = ic[0], k_atr = ic[1], dist_eq = ic[2], dist_max = ic[3];
  if (s->pn_mode) {
    /* max_dist = 2 * particle_radius(interaction_cfg) (src/rings/integration.jl:40-41; HarmTruncCfg radius = dist_eq/2,
     * src/configs.jl:418-421); neigh_update!(::ParticleNeighbors, ...) src/rings/neighbors.jl:125-134 */
    double max_dist = 2 * (dist_eq / 2);
    if (dist < max_dist * s->pn_tol && (s->pn_all || ri != rj)) {
      OrSystem *m = (OrSystem *)s;
#pragma omp critical(mor_neigh)
      {
        /* neigh_update_data!, src/rings/neighbors.jl:102-116 */
        int32_t ci = ++m->pn_count[i], cj = ++m->pn_count[j];
        if (m->pn_mode == MAVI_NEIGH_LIST) {
          if (ci <= MAVI_NEIGH_MAX) m->pn_list[i * MAVI_NEIGH_MAX + ci - 1] = (int32_t)j; else m->pn_overflow = 1;
          if (cj <= MAVI_NEIGH_MAX) m->pn_list[j * MAVI_NEIGH_MAX + cj - 1] = (int32_t)i; else m->pn_overflow = 1;
        }
      }
    }
  }
  f[0] = 0.0; f[1] = 0.0;
  if (dist > dist_max) return;
  if (ri == rj) {
    int64_t diff = i > j ? i - j : j - i;
    int64_t num_p = ring_np(s, ri);
    if (diff == 1 || diff == num_p - 1) return;
  }
  double fmod_;
  if (dist < dist_eq) fmod_ = -k_rep * (dist / dist_eq - 1);
  else {
    if (ri == rj) fmod_ = 0.0;
    else fmod_ = -k_atr * (dist / dist_eq - 1);
  }
  double c = fmod_ / dist;
  f[0] = c * dr[0]; f[1] = c * dr[1];
}

/* calc_interaction dispatch, src/integration.jl:62-109, src/rings/integration.jl:32-45 */
static inline void interaction(const OrSystem *s, int64_t i, int64_t j, double *f) {
  double dr[2];
  switch (s->p.dynamics) {
    case MAVI_DYN_LJ:
      mor_calc_diff(s, s->pos + 2 * i, s->pos + 2 * j, dr);
      mor_potential_force(MAVI_POT_LJ, s->p.dyn, dr, sqrt(dr[0] * dr[0] + dr[1] * dr[1]), f);
      break;
    case MAVI_DYN_HARMTRUNC:
      mor_calc_diff(s, s->pos + 2 * i, s->pos + 2 * j, dr);
      mor_potential_force(MAVI_POT_HARMTRUNC, s->p.dyn, dr, sqrt(dr[0] * dr[0] + dr[1] * dr[1]), f);
      break;
    case MAVI_DYN_SZABO:
      mor_calc_diff(s, s->pos + 2 * i, s->pos + 2 * j, dr);
      mor_szabo_interaction(s->p.dyn, dr, f);
      break;
    case MAVI_DYN_RTP:
      mor_calc_diff(s, s->pos + 2 * i, s->pos + 2 * j, dr);
      mor_rtp_interaction(s->p.dyn, dr, f);
      break;
    default:
      rings_interaction(s, i, j, f);
  }
}

/* ------------------------------------------------------------------ chunks */

/* get_neighbors periodic, src/chunks.jl:61-87 */
static int64_t wrap_id(int64_t x, int64_t num_t) {
  if (x == 0) return num_t;
  if (x % (num_t + 1) == 0) return 1;
  return x;
}

typedef struct NeighTab { int32_t *neigh; int8_t *neigh_n; } NeighTab;
static void build_neighbors_tab(NeighTab *s, int64_t nr, int64_t ncl, int periodic);
static void build_neighbors(OrSystem *sys) {
  NeighTab t;
  build_neighbors_tab(&t, sys->num_rows, sys->num_cols, sys->p.spaces[0].wall == MAVI_WALL_PERIODIC);
  sys->neigh = t.neigh;
  sys->neigh_n = t.neigh_n;
}
static void build_neighbors_tab(NeighTab *s, int64_t nr, int64_t ncl, int periodic) {
  s->neigh = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)(nr * ncl));
  s->neigh_n = (int8_t *)calloc((size_t)(nr * ncl), 1);
#define CELL(i, j) ((int32_t)(((i)-1) + nr * ((j)-1)))
#define PUSH(i, j, ii, jj)                                          \
  do {                                                              \
    int32_t c_ = CELL(i, j);                                        \
    s->neigh[4 * c_ + s->neigh_n[c_]] = CELL(ii, jj);               \
    s->neigh_n[c_]++;                                               \
  } while (0)
  if (periodic) {
    for (int64_t i = 1; i <= nr; i++)
      for (int64_t j = 1; j <= ncl; j++) {
        PUSH(i, j, wrap_id(i + 1, nr), wrap_id(j, ncl));
        PUSH(i, j, wrap_id(i + 1, nr), wrap_id(j + 1, ncl));
        PUSH(i, j, wrap_id(i, nr), wrap_id(j + 1, ncl));
        PUSH(i, j, wrap_id(i - 1, nr), wrap_id(j + 1, ncl));
      }
  } else {
    /* get_neighbors walled, src/chunks.jl:89-118.  Later assignments overwrite earlier ones. */
    for (int64_t i = 1; i <= nr - 1; i++) {
      s->neigh_n[CELL(i, 1)] = 0;
      PUSH(i, 1, i + 1, 1);
      PUSH(i, 1, i + 1, 2);
      PUSH(i, 1, i, 2);
      for (int64_t j = 2; j <= ncl - 1; j++) {
        s->neigh_n[CELL(i, j)] = 0;
        PUSH(i, j, i + 1, j - 1);
        PUSH(i, j, i + 1, j);
        PUSH(i, j, i + 1, j + 1);
        PUSH(i, j, i, j + 1);
      }
      s->neigh_n[CELL(i, ncl)] = 0;
      PUSH(i, ncl, i + 1, ncl - 1);
      PUSH(i, ncl, i + 1, ncl);
    }
    for (int64_t j = 1; j <= ncl - 1; j++) {
      s->neigh_n[CELL(nr, j)] = 0;
      PUSH(nr, j, nr, j + 1);
    }
    s->neigh_n[CELL(nr, ncl)] = 0;
  }
#undef PUSH
#undef CELL
}

/* Chunks ctor, src/chunks.jl:26-40 (+ get_chunks, src/systems.jl:14-28) */
static int32_t chunks_init(OrSystem *s) {
  s->has_chunks = s->p.num_cols > 0;
  if (!s->has_chunks) return MAVI_OK;
  s->num_cols = s->p.num_cols;
  s->num_rows = s->p.num_rows;
  if (s->num_rows < 1) return MAVI_ERR_BAD_PARAMS;
  if (s->p.spaces[0].wall != MAVI_WALL_PERIODIC && s->num_cols < 2) {
    snprintf(s->err, sizeof s->err, "walled neighbour table needs num_cols >= 2 (reference indexes column 0)");
    return MAVI_ERR_BAD_PARAMS;
  }
  s->cl = s->p.grid_len / (double)s->num_cols;
  s->ch = s->p.grid_h / (double)s->num_rows;
  double r = s->p.particle_radius;
  double ncf = (ceil(0.5 * s->cl / r) + 1) * (ceil(0.5 * s->ch / r) + 1);
  s->nc = (int64_t)ceil(ncf * 2);
  size_t cells = (size_t)(s->num_rows * s->num_cols);
  s->chunk_particles = (int64_t *)malloc(sizeof(int64_t) * cells * (size_t)s->nc);
  s->num_in_chunk = (int64_t *)calloc(cells, sizeof(int64_t));
  build_neighbors(s);
  return MAVI_OK;
}

/* update_particle_chunk!, src/chunks.jl:120-147.  i is 0-based here; the table stores 0-based ids. */
static int32_t update_particle_chunk(OrSystem *s, int64_t i) {
  const double space_h = s->p.grid_h;
  const double *bl = s->p.grid_bl;
  const double x = s->pos[2 * i], y = s->pos[2 * i + 1];
  double rowf = mor_julia_div(-y + bl[1] + space_h, s->ch);
  double colf = mor_julia_div(x - bl[0], s->cl);
  if (!(fabs(rowf) < 9.0e15) || !(fabs(colf) < 9.0e15)) { /* trunc(Int, NaN/Inf) -> InexactError */
    snprintf(s->err, sizeof s->err, "particle %lld: non-finite cell index", (long long)i);
    return MAVI_ERR_OUT_OF_GRID;
  }
  int64_t row_id = (int64_t)rowf + 1;
  int64_t col_id = (int64_t)colf + 1;
  row_id -= row_id == (s->num_rows + 1) ? 1 : 0;
  col_id -= col_id == (s->num_cols + 1) ? 1 : 0;
  if (row_id < 1 || row_id > s->num_rows || col_id < 1 || col_id > s->num_cols) {
    snprintf(s->err, sizeof s->err, "particle %lld out of grid (row %lld col %lld): BoundsError in the reference",
             (long long)i, (long long)row_id, (long long)col_id);
    return MAVI_ERR_OUT_OF_GRID;
  }
  int64_t cell = (row_id - 1) + s->num_rows * (col_id - 1);
  int64_t p_i = s->num_in_chunk[cell];
  if (p_i >= s->nc) {
    snprintf(s->err, sizeof s->err, "cell %lld over capacity nc=%lld: BoundsError in the reference", (long long)cell,
             (long long)s->nc);
    return MAVI_ERR_CAPACITY;
  }
  s->chunk_particles[cell * s->nc + p_i] = i;
  s->num_in_chunk[cell] += 1;
  return MAVI_OK;
}

/* update_chunks!, State-aware: active ids only.  src/integration.jl:54-59, src/rings/integration.jl:18-23 */
int32_t mor_update_chunks(OrSystem *s) {
  if (!s->has_chunks) return MAVI_OK; /* update_chunks!(::Nothing), src/chunks.jl:165 */
  memset(s->num_in_chunk, 0, sizeof(int64_t) * (size_t)(s->num_rows * s->num_cols));
  for (int64_t k = 0; k < s->n_ids; k++) {
    int32_t st = update_particle_chunk(s, s->ids[k]);
    if (st) return st;
  }
  return MAVI_OK;
}

/* ------------------------------------------------------------------ forces */

/* clean_forces!, src/systems.jl:119-123 */
void mor_clean_forces(OrSystem *s) {
  for (int t = 0; t < s->nthreads; t++) memset(s->forces[t], 0, sizeof(double) * 2 * (size_t)s->n);
}

static inline void pair_scatter(const OrSystem *s, double *forces, int64_t p1, int64_t p2) {
  double f[2];
  interaction(s, p1, p2, f);
  forces[2 * p1] += f[0];
  forces[2 * p1 + 1] += f[1];
  forces[2 * p2] -= f[0];
  forces[2 * p2 + 1] -= f[1];
}

/* one cell column of calc_forces!(system, chunks, device), src/integration.jl:116-156 / :160-190 */
static void forces_column(const OrSystem *s, double *forces, int64_t col /*0-based*/) {
  for (int64_t row = 0; row < s->num_rows; row++) {
    int64_t cell = row + s->num_rows * col;
    int64_t np = s->num_in_chunk[cell];
    const int64_t *chunk = s->chunk_particles + cell * s->nc;
    for (int64_t i = 0; i < np; i++) {
      int64_t p1 = chunk[i];
      for (int64_t j = i + 1; j < np; j++) pair_scatter(s, forces, p1, chunk[j]);
      for (int k = 0; k < s->neigh_n[cell]; k++) {
        int64_t ncell = s->neigh[4 * cell + k];
        int64_t nei_np = s->num_in_chunk[ncell];
        const int64_t *nei_chunk = s->chunk_particles + ncell * s->nc;
        for (int64_t j = 0; j < nei_np; j++) pair_scatter(s, forces, p1, nei_chunk[j]);
      }
    }
  }
}

/* calc_forces!(system): dispatch of src/integration.jl:226 */
void mor_pair_forces(OrSystem *s) {
  if (!s->has_chunks) {
    /* all pairs over active ids, src/integration.jl:197-224 */
    double *forces = s->forces[0];
    for (int64_t i = 0; i < s->n_ids; i++)
      for (int64_t j = i + 1; j < s->n_ids; j++) pair_scatter(s, forces, s->ids[i], s->ids[j]);
    return;
  }
  if (s->nthreads <= 1) {
    /* Sequencial, src/integration.jl:112-157: for col, for row */
    for (int64_t col = 0; col < s->num_cols; col++) forces_column(s, s->forces[0], col);
    return;
  }
  /* Threaded, src/integration.jl:159-194: @threads :static over columns (contiguous blocks, the first
   * `rem` threads get one extra), private slices, then get_forces .= sum(system.forces). */
  const int T = s->nthreads;
  const int64_t len = s->num_cols / T, rem = s->num_cols % T;
#pragma omp parallel for schedule(static, 1) num_threads(T)
  for (int t = 0; t < T; t++) {
    int64_t lo = t * len + (t < rem ? t : rem);
    int64_t hi = lo + len + (t < rem ? 1 : 0);
    for (int64_t col = lo; col < hi; col++) forces_column(s, s->forces[t], col);
  }
  double *f0 = s->forces[0];
#pragma omp parallel for schedule(static) num_threads(T)
  for (int64_t k = 0; k < 2 * s->n; k++) {
    double acc = f0[k];
    for (int t = 1; t < T; t++) acc += s->forces[t][k];
    f0[k] = acc;
  }
}

/* signed_pos(point, ::CircleCfg), src/configs.jl:154-163 */
static void signed_pos_circle(const double *pt, const MaviSpace *sp, double *dr, double *dist, double *flag) {
  double d0 = pt[0] - sp->circ_center[0], d1 = pt[1] - sp->circ_center[1];
  double dd = sqrt(d0 * d0 + d1 * d1); /* sum(dr.^2)^.5 */
  double h0 = d0 / dd, h1 = d1 / dd;
  dr[0] = d0 - h0 * sp->circ_radius;
  dr[1] = d1 - h1 * sp->circ_radius;
  double sd = dd - sp->circ_radius;
  *dist = fabs(sd);
  *flag = sd > 0 ? 1.0 : (sd < 0 ? -1.0 : sd); /* sign() */
}

/* Line2D ctor, src/configs.jl:102-117 */
static void line_frame(const MaviLine *l, double *normal, double *tangent, double *length) {
  double d0 = l->p2[0] - l->p1[0], d1 = l->p2[1] - l->p1[1];
  double norm = sqrt(d0 * d0 + d1 * d1);
  normal[0] = -d1 / norm; normal[1] = d0 / norm;
  tangent[0] = d0 / norm; tangent[1] = d1 / norm;
  *length = norm;
}

/* signed_pos(point, ::Line2D), src/configs.jl:119-135 */
static void signed_pos_line(const double *pt, const MaviLine *l, double *dr, double *dist) {
  double nrm[2], tan[2], len;
  line_frame(l, nrm, tan, &len);
  dr[0] = pt[0] - l->p1[0];
  dr[1] = pt[1] - l->p1[1];
  double delta_t = dr[0] * tan[0] + dr[1] * tan[1];
  if (delta_t > 0) {
    if (delta_t < len) {
      double b0 = l->p1[0] + tan[0] * delta_t, b1 = l->p1[1] + tan[1] * delta_t;
      dr[0] = pt[0] - b0; dr[1] = pt[1] - b1;
    } else {
      dr[0] = pt[0] - l->p2[0]; dr[1] = pt[1] - l->p2[1];
    }
  }
  *dist = sqrt(dr[0] * dr[0] + dr[1] * dr[1]);
}

/* process_dist, src/configs.jl:259-261 */
static double process_dist(int32_t mode, double dist, double flag) {
  if (mode == MAVI_WALLMODE_OUTSIDE) return flag * dist;
  if (mode == MAVI_WALLMODE_INSIDE) return -flag * dist;
  return dist;
}

/* calc_walls_forces!, src/integration.jl:228-266 (ManyWalls: only ForceWalls sub-spaces act) */
void mor_walls_forces(OrSystem *s) {
  double *forces = s->forces[0];
  for (int k = 0; k < s->p.n_spaces; k++) {
    const MaviSpace *sp = &s->p.spaces[k];
    if (sp->wall != MAVI_WALL_POTENTIAL) continue;
    for (int64_t q = 0; q < s->n_ids; q++) {
      int64_t i = s->ids[q];
      const double *pt = s->pos + 2 * i;
      double dr[2], dist, flag, f[2];
      /* get_potential_cfg(wall_pot.potential, state, i), src/integration.jl:237,248: PotentialVector picks the entry of the
       * particle's type (src/configs.jl:454-463) = the type of its ring (src/rings/states.jl:148) */
      const double *pot = sp->pot;
      if (sp->n_pot_types > 0) pot = sp->pot_types[s->p.dynamics == MAVI_DYN_RINGS ? ring_type(s, i / s->rp.n_max) : 0];
      if (sp->geom == MAVI_GEOM_CIRCLE) {
        signed_pos_circle(pt, sp, dr, &dist, &flag);
        dist = process_dist(sp->pot_mode, dist, flag);
        mor_potential_force(sp->pot_kind, pot, dr, dist, f);
        forces[2 * i] += f[0]; forces[2 * i + 1] += f[1];
      } else if (sp->geom == MAVI_GEOM_LINES) {
        for (int l = 0; l < sp->n_lines; l++) {
          signed_pos_line(pt, &s->lines[k][l], dr, &dist);
          dist = process_dist(sp->pot_mode, dist, 1.0);
          mor_potential_force(sp->pot_kind, pot, dr, dist, f);
          forces[2 * i] += f[0]; forces[2 * i + 1] += f[1];
        }
      }
      /* PotentialWalls on a RectangleCfg: no signed_pos method exists in the reference -> MethodError;
       * rejected at create. */
    }
  }
}

/* per-particle radius: get_particle_radius, src/systems.jl:132-133, src/rings/rings.jl:47-56 */
static double particle_radius_of(const OrSystem *s, int64_t idx) {
  if (s->p.dynamics != MAVI_DYN_RINGS) return s->p.particle_radius;
  int32_t t = ring_type(s, ring_of(s, idx));
  return s->rp.interaction[4 * ((int64_t)t * s->rp.num_types + t) + 2] / 2.0;
}

/* walls!(system, SpaceCfg{W,G}) for one (wall, geometry) pair, src/integration.jl:268-401 */
static void walls_one(OrSystem *s, const MaviSpace *sp, const MaviLine *lines) {
  double *pos = s->pos;
  if (sp->wall == MAVI_WALL_RIGID && sp->geom == MAVI_GEOM_RECT) {
    /* :271-285 — velocity flip only, uses particle_radius(dynamic_cfg) */
    double *vel = s->second;
    const double r = s->p.particle_radius;
    const double size[2] = {sp->rect_len, sp->rect_h};
    for (int64_t q = 0; q < s->n_ids; q++) {
      int64_t i = s->ids[q];
      int out[2];
      for (int d = 0; d < 2; d++) {
        double rel = pos[2 * i + d] - sp->rect_bl[d];
        out[d] = ((rel + r) > size[d]) || ((rel - r) < 0);
      }
      if (out[0] || out[1])
        for (int d = 0; d < 2; d++) vel[2 * i + d] = vel[2 * i + d] * (double)(-2 * out[d] + 1);
    }
  } else if (sp->wall == MAVI_WALL_RIGID && sp->geom == MAVI_GEOM_CIRCLE) {
    /* :287-306 — cross terms use raw pos (exact only for centre 0; kept) */
    double *vel = s->second;
    double mr = sp->circ_radius - s->p.particle_radius;
    double max_r2 = mr * mr;
    for (int64_t q = 0; q < s->n_ids; q++) {
      int64_t i = s->ids[q];
      double px = pos[2 * i], py = pos[2 * i + 1];
      double ex = px - sp->circ_center[0], ey = py - sp->circ_center[1];
      double dr2x = ex * ex, dr2y = ey * ey;
      double r2 = dr2x + dr2y;
      if (r2 <= max_r2) continue;
      double vx = vel[2 * i], vy = vel[2 * i + 1];
      double nvx = (vx * (dr2y - dr2x) - 2 * vy * px * py) / r2;
      double nvy = (-vy * (dr2y - dr2x) - 2 * vx * px * py) / r2;
      vel[2 * i] = nvx; vel[2 * i + 1] = nvy;
    }
  } else if (sp->wall == MAVI_WALL_PERIODIC && sp->geom == MAVI_GEOM_RECT) {
    /* :309-324 */
    const double size[2] = {sp->rect_len, sp->rect_h};
    for (int64_t q = 0; q < s->n_ids; q++) {
      int64_t i = s->ids[q];
      int out[2], any = 0;
      double diff[2], half[2];
      for (int d = 0; d < 2; d++) {
        half[d] = size[d] / 2;
        double center = sp->rect_bl[d] + size[d] / 2;
        diff[d] = pos[2 * i + d] - center;
        out[d] = fabs(diff[d]) > half[d];
        any |= out[d];
      }
      if (any)
        for (int d = 0; d < 2; d++) {
          double sg = diff[d] > 0 ? 1.0 : (diff[d] < 0 ? -1.0 : diff[d]);
          pos[2 * i + d] = pos[2 * i + d] - sg * (half[d] * 2) * (double)out[d];
        }
    }
  } else if (sp->wall == MAVI_WALL_SLIPPERY && sp->geom == MAVI_GEOM_LINES) {
    /* :327-378 — pos_i is read once per particle, corrections accumulate into state.pos */
    for (int64_t q = 0; q < s->n_ids; q++) {
      int64_t pid = s->ids[q];
      double pr = particle_radius_of(s, pid);
      const double pi0 = pos[2 * pid], pi1 = pos[2 * pid + 1];
      for (int l = 0; l < sp->n_lines; l++) {
        double nrm[2], tan[2], len;
        line_frame(&lines[l], nrm, tan, &len);
        double dr0 = pi0 - lines[l].p1[0], dr1 = pi1 - lines[l].p1[1];
        double delta_s = dr0 * nrm[0] + dr1 * nrm[1];
        if (fabs(delta_s) > pr) continue;
        double delta_t = dr0 * tan[0] + dr1 * tan[1];
        int is_corner = 0;
        const double *corner = NULL;
        if (delta_t > 0) {
          if (delta_t > len) {
            if (delta_t > (len + pr)) continue;
            is_corner = 1;
            corner = lines[l].p2;
          }
        } else if (delta_t > -pr) {
          is_corner = 1;
          corner = lines[l].p1;
        } else {
          continue;
        }
        if (is_corner) {
          dr0 = pi0 - corner[0]; dr1 = pi1 - corner[1];
          double norm = sqrt(dr0 * dr0 + dr1 * dr1);
          if (norm > pr) continue;
          double alpha = pr / norm - 1;
          pos[2 * pid] += alpha * dr0;
          pos[2 * pid + 1] += alpha * dr1;
        } else {
          double sgn = delta_s > 0 ? 1.0 : (delta_s < 0 ? -1.0 : delta_s);
          double alpha = sgn * (pr - sgn * delta_s);
          pos[2 * pid] += alpha * nrm[0];
          pos[2 * pid + 1] += alpha * nrm[1];
        }
      }
    }
  } else if (sp->wall == MAVI_WALL_SLIPPERY && sp->geom == MAVI_GEOM_CIRCLE) {
    /* :380-401 — calc_diff uses the SYSTEM space_cfg (min image if the main space is periodic) */
    for (int64_t q = 0; q < s->n_ids; q++) {
      int64_t pid = s->ids[q];
      double pr = particle_radius_of(s, pid);
      double mr = sp->circ_radius + pr;
      double max_r_2 = mr * mr;
      double dr[2];
      mor_calc_diff(s, pos + 2 * pid, sp->circ_center, dr);
      double dr_2 = dr[0] * dr[0] + dr[1] * dr[1];
      if (dr_2 > max_r_2) continue;
      double dr_norm = sqrt(dr_2);
      double k = sp->circ_radius + pr - dr_norm;
      pos[2 * pid] = pos[2 * pid] + k * dr[0] / dr_norm;
      pos[2 * pid + 1] = pos[2 * pid + 1] + k * dr[1] / dr_norm;
    }
  }
  /* every other (wall, geometry) pair: generic no-op method walls!(system, ::SpaceCfg), :268 */
}

/* walls!(system): single space or ManyWalls loop, src/integration.jl:404-412 */
void mor_walls(OrSystem *s) {
  for (int k = 0; k < s->p.n_spaces; k++) walls_one(s, &s->p.spaces[k], s->lines[k]);
}

/* ------------------------------------------------------------------ integrators */

/* update_verlet!, src/integration.jl:415-431.  Pass 2 runs on the stale chunks and without wall forces;
 * the broadcasts cover every slot of pos/vel, active or not. */
void mor_update_verlet(OrSystem *s) {
  const int64_t m = 2 * s->n;
  double *forces = s->forces[0];
  double *old = dup_d(forces, m);
  const double dt = s->p.dt;
  const double term = dt * dt / 2; /* dt^2/2 */
  double *pos = s->pos, *vel = s->second;
  for (int64_t k = 0; k < m; k++) pos[k] = pos[k] + (vel[k] * dt + forces[k] * term);
  mor_clean_forces(s);
  mor_pair_forces(s);
  const double hdt = dt / 2;
  for (int64_t k = 0; k < m; k++) vel[k] = vel[k] + hdt * (forces[k] + old[k]);
  free(old);
}

static inline double sign_d(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : x); }

/* update_szabo!, src/integration.jl:433-465.  Loops 1:get_num_total_particles (slots, not ids);
 * "speed" is sqrt(|vx|+|vy|) (sic).  noise[i] stands for the reference's global randn(). */
static void update_szabo(OrSystem *s, const double *noise) {
  const double *par = s->p.dyn;
  const double vo = par[0], mu = par[1], relax_time = par[2], drot = par[7];
  const double dt = s->p.dt;
  double *forces = s->forces[0], *pos = s->pos, *ang = s->second;
  for (int64_t i = 0; i < s->n_ids; i++) {
    double theta = ang[i];
    double polx = cos(theta), poly = sin(theta);
    double velx = vo * polx + mu * forces[2 * i], vely = vo * poly + mu * forces[2 * i + 1];
    double speed = sqrt(fabs(velx) + fabs(vely));
    double cross_prod;
    if (speed > 0) cross_prod = (polx * vely - poly * velx) / speed;
    else cross_prod = 0;
    if (fabs(cross_prod) > 1) cross_prod = sign_d(cross_prod);
    double nz = noise ? noise[i] : 0.0;
    double d_theta = 1 / relax_time * asin(cross_prod) * dt + sqrt(2 * drot * dt) * nz;
    pos[2 * i] += velx * dt;
    pos[2 * i + 1] += vely * dt;
    ang[i] += d_theta;
  }
}

/* update_rtp!, src/integration.jl:467-498.  noise[2i] = u, noise[2i+1] = the second rand() (used on tumble). */
static void update_rtp(OrSystem *s, const double *noise) {
  const double *par = s->p.dyn;
  const double vo = par[0], tumble_rate = par[3];
  const double dt = s->p.dt;
  const double two_pi = 2 * M_PI;
  double *forces = s->forces[0], *pos = s->pos, *ang = s->second;
  for (int64_t i = 0; i < s->n_ids; i++) {
    double theta = ang[i];
    double polx = cos(theta), poly = sin(theta);
    double velx = vo * polx + forces[2 * i], vely = vo * poly + forces[2 * i + 1];
    pos[2 * i] += velx * dt;
    pos[2 * i + 1] += vely * dt;
    double u = noise ? noise[2 * i] : 1.0;
    if (u < tumble_rate * dt) ang[i] = two_pi * (noise ? noise[2 * i + 1] : 0.0);
  }
}

/* update_time!, src/integration.jl:500-503 */
static void update_time(OrSystem *s) {
  s->time += s->p.dt;
  s->num_steps += 1;
}

/* ------------------------------------------------------------------ rings */

/* update_continuos_pos!, src/rings/integration.jl:118-138 (periodic main wall only; otherwise the
 * "continuous" positions are rings_pos itself, src/rings/rings.jl:31-43) */
static void update_continuos_pos(OrSystem *s) {
  if (s->p.spaces[0].wall != MAVI_WALL_PERIODIC) return;
  const int64_t nm = s->rp.n_max;
  for (int64_t rq = 0; rq < n_ring_ids(s); rq++) {  /* get_rings_ids(state) */
    const int64_t ring = ring_id_at(s, rq);
    memcpy(s->cont_pos + 2 * ring * nm, s->pos + 2 * ring * nm, sizeof(double) * 2 * (size_t)nm);
    int32_t np = ring_np(s, ring);
    for (int32_t i = 1; i < np; i++) {
      int64_t a = ring * nm + i, b = a - 1;
      double dr[2];
      mor_calc_diff(s, s->pos + 2 * a, s->pos + 2 * b, dr);
      s->cont_pos[2 * a] = s->cont_pos[2 * b] + dr[0];
      s->cont_pos[2 * a + 1] = s->cont_pos[2 * b + 1] + dr[1];
    }
  }
}

static const double *ring_points(const OrSystem *s, int64_t ring) {
  const double *base = s->p.spaces[0].wall == MAVI_WALL_PERIODIC ? s->cont_pos : s->pos;
  return base + 2 * ring * s->rp.n_max;
}

/* update_cms!, src/rings/integration.jl:366-372 */
static void update_cms(OrSystem *s) {
  for (int64_t rq = 0; rq < n_ring_ids(s); rq++) {  /* get_rings_ids(state) */
    const int64_t ring = ring_id_at(s, rq);
    const double *pts = ring_points(s, ring);
    int32_t np = ring_np(s, ring);
    double sx = pts[0], sy = pts[1];
    for (int32_t i = 1; i < np; i++) { sx += pts[2 * i]; sy += pts[2 * i + 1]; }
    s->cms[2 * ring] = sx / np;
    s->cms[2 * ring + 1] = sy / np;
  }
}

/* calc_area, src/rings/integration.jl:103-116 (shoelace) */
static double calc_area(const double *pts, int32_t np) {
  double area = 0.0;
  for (int32_t i = 0; i < np - 1; i++) area += pts[2 * i] * pts[2 * (i + 1) + 1] - pts[2 * i + 1] * pts[2 * (i + 1)];
  area += pts[2 * (np - 1)] * pts[1] - pts[2 * (np - 1) + 1] * pts[0];
  return area / 2.0;
}

/* forces!, src/rings/integration.jl:197-226 = calc_forces! + springs (:79-97) + area_forces! (:140-195) */
static void rings_forces(OrSystem *s) {
  if (s->pn_mode) { /* neigh_clean!, src/rings/integration.jl:362 */
    memset(s->pn_count, 0, sizeof(int32_t) * (size_t)s->n);
    if (s->pn_list) memset(s->pn_list, 0xff, sizeof(int32_t) * (size_t)s->n * MAVI_NEIGH_MAX);
  }
  mor_pair_forces(s);
  double *forces = s->forces[0];
  const int64_t nm = s->rp.n_max;
  for (int64_t rq = 0; rq < n_ring_ids(s); rq++) {  /* get_rings_ids(state) */
    const int64_t ring = ring_id_at(s, rq);
    int32_t np = ring_np(s, ring);
    int32_t t = ring_type(s, ring);
    double k = s->rp.k_spring[t], l = s->rp.l_spring[t];
    for (int32_t sp = 0; sp < np; sp++) {
      int64_t p1 = ring * nm + sp, p2 = ring * nm + (sp == np - 1 ? 0 : sp + 1);
      double dr[2];
      mor_calc_diff(s, s->pos + 2 * p1, s->pos + 2 * p2, dr);
      double dist = sqrt(dr[0] * dr[0] + dr[1] * dr[1]);
      double fmod_ = -k * (dist - l);
      double c = fmod_ / dist;
      double f0 = c * dr[0], f1 = c * dr[1];
      forces[2 * p1] += f0; forces[2 * p1 + 1] += f1;
      forces[2 * p2] -= f0; forces[2 * p2 + 1] -= f1;
    }
  }
  for (int64_t rq = 0; rq < n_ring_ids(s); rq++) {  /* get_rings_ids(state) */
    const int64_t ring = ring_id_at(s, rq);
    int32_t np = ring_np(s, ring);
    int32_t t = ring_type(s, ring);
    double area = calc_area(ring_points(s, ring), np);
    s->areas[ring] = area;
    double k_area = s->rp.k_area[t], p0 = s->rp.p0[t], l0 = s->rp.l_spring[t];
    double a0s = np * l0 / p0;
    double area0 = a0s * a0s;
    for (int32_t i = 0; i < np; i++) {
      double fmod_ = k_area * (area - area0);
      int32_t id1 = i == 0 ? np - 1 : i - 1;
      int32_t id2 = i == np - 1 ? 0 : i + 1;
      double dr[2];
      mor_calc_diff(s, s->pos + 2 * (ring * nm + id2), s->pos + 2 * (ring * nm + id1), dr);
      double ax = dr[1] / 2, ay = -dr[0] / 2;
      forces[2 * (ring * nm + i)] -= fmod_ * ax;
      forces[2 * (ring * nm + i) + 1] -= fmod_ * ay;
    }
  }
}

/* update!, src/rings/integration.jl:300-351.  noise[ring] stands for randn(system.rng). */
static void rings_update(OrSystem *s, const double *noise) {
  const double dt = s->p.dt;
  double *forces = s->forces[0], *pos = s->pos, *pol_a = s->second;
  const int64_t nm = s->rp.n_max;
  for (int64_t rq = 0; rq < n_ring_ids(s); rq++) {  /* get_rings_ids(state) */
    const int64_t ring = ring_id_at(s, rq);
    int32_t np = ring_np(s, ring);
    int32_t t = ring_type(s, ring);
    double vo = s->rp.vo[t], relax_time = s->rp.relax_time[t], mu = s->rp.mobility[t], drot = s->rp.rot_diff[t];
    double theta = pol_a[ring];
    double polx = cos(theta), poly = sin(theta);
    double vcx = 0.0, vcy = 0.0;
    for (int32_t i = 0; i < np; i++) {
      int64_t pid = ring * nm + i;
      double velx = vo * polx + mu * forces[2 * pid], vely = vo * poly + mu * forces[2 * pid + 1];
      vcx += velx; vcy += vely;
      pos[2 * pid] += velx * dt;
      pos[2 * pid + 1] += vely * dt;
    }
    vcx /= np; vcy /= np;
    double speed = sqrt(vcx * vcx + vcy * vcy);
    double cross_prod;
    if (speed == 0) cross_prod = 0;
    else {
      cross_prod = (polx * vcy - poly * vcx) / speed;
      if (fabs(cross_prod) > 1) cross_prod = sign_d(cross_prod);
    }
    double nz = noise ? noise[ring] : 0.0;
    double d_theta = 1 / relax_time * asin(cross_prod) * dt + sqrt(2 * drot * dt) * nz;
    pol_a[ring] += d_theta;
  }
}


/* ------------------------------------------------------------------ sources / sinks / variable ring count */

/* calc_active_ids!, src/rings/states.jl:200-223 (update_ids!) */
static void calc_active_ids(OrSystem *s) {
  if (!s->var_rings) return;
  int64_t pointer = 0, p_pointer = 0;
  const int64_t nm = s->rp.n_max;
  memset(s->mask, 0, (size_t)s->n);
  for (int64_t ring = 0; ring < s->rp.num_rings; ring++) {
    if (!s->ring_mask[ring]) continue;
    s->ring_ids[pointer++] = ring;
    int32_t np = ring_np(s, ring);
    for (int32_t q = 0; q < np; q++) {
      s->ids[p_pointer++] = ring * nm + q;
      s->mask[ring * nm + q] = 1;
    }
  }
  s->n_ids = p_pointer;
  s->num_active = pointer;
}

/* is_inside, src/configs.jl:89-93 (RectangleCfg) */
static int inside_rect(const double *pt, const double *r /* bl.x bl.y len h */, double pad) {
  int is_x = r[0] - pad <= pt[0] && pt[0] <= r[0] + r[2] + pad;
  int is_y = r[1] - pad <= pt[1] && pt[1] <= r[1] + r[3] + pad;
  return is_x && is_y;
}

/* check_intersection(r1::RectangleCfg, r2::RectangleCfg), src/configs.jl:79-87 (only r1's edges are tested against r2) */
static int rect_intersect(const double *r1, const double *r2) {
  int x1 = r2[0] <= r1[0] && r1[0] <= r2[0] + r2[2];
  int x2 = r2[0] <= r1[0] + r1[2] && r1[0] + r1[2] <= r2[0] + r2[2];
  int y1 = r2[1] <= r1[1] && r1[1] <= r2[1] + r2[3];
  int y2 = r2[1] <= r1[1] + r1[3] && r1[1] + r1[3] <= r2[1] + r2[3];
  return (x1 || x2) && (y1 || y2);
}

/* Source ctor (src/rings/sources.jl:134-190) and, with chunks, ChunksChecker ctor (:52-119) */
static int32_t build_source(OrSystem *s, OrSource *o, const MaviSourceSink *c) {
  memset(o, 0, sizeof *o);
  o->kind = c->kind;
  if (c->kind == MAVI_SRC_SINK) {
    o->sink_geom = c->sink_geom;
    if (c->sink_geom == MAVI_GEOM_RECT) {
      o->sink[0] = c->sink_rect_bl[0]; o->sink[1] = c->sink_rect_bl[1]; o->sink[2] = c->sink_rect_len; o->sink[3] = c->sink_rect_h;
    } else {
      o->sink[0] = c->sink_circ_center[0]; o->sink[1] = c->sink_circ_center[1]; o->sink[2] = c->sink_circ_radius;
    }
    return MAVI_OK;
  }
  if (c->num_spawn_pos != s->rp.n_max) {
    snprintf(s->err, sizeof s->err, "SourceCfg.spawn_pos has %d points, rings_pos[:, ring] has %d", c->num_spawn_pos, s->rp.n_max);
    return MAVI_ERR_BAD_PARAMS;
  }
  const int nsp = c->num_spawn_pos;
  const double pad = c->pad;
  o->nsp = nsp; o->pad = pad; o->spawn_pol = c->spawn_pol;
  double min_x = c->spawn_pos[0], max_x = min_x, min_y = c->spawn_pos[1], max_y = min_y;
  for (int i = 1; i < nsp; i++) {
    double x = c->spawn_pos[2 * i], y = c->spawn_pos[2 * i + 1];
    if (x < min_x) min_x = x;
    if (x > max_x) max_x = x;
    if (y < min_y) min_y = y;
    if (y > max_y) max_y = y;
  }
  const double bl_len = max_x - min_x + 2 * pad, bl_h = max_y - min_y + 2 * pad;
  const int ns = c->size[0] * c->size[1];
  o->nspawn = ns;
  o->bbox = (double *)calloc((size_t)ns * 4, sizeof(double));
  double *bbox_pad = (double *)calloc((size_t)ns * 4, sizeof(double));
  o->spawn = (double *)calloc((size_t)ns * nsp * 2, sizeof(double));
  o->is_empty = (uint8_t *)calloc((size_t)ns, 1);
  int idx = 0;
  for (int i = 1; i <= c->size[0]; i++)
    for (int j = 1; j <= c->size[1]; j++, idx++) {
      double bx = c->bottom_left[0] + ((i - 1) * bl_len + i * c->offset[0]);
      double by = c->bottom_left[1] + ((j - 1) * bl_h + j * c->offset[1]);
      double *b = o->bbox + 4 * idx, *bp = bbox_pad + 4 * idx;
      b[0] = bx; b[1] = by; b[2] = bl_len; b[3] = bl_h;
      bp[0] = bx - pad; bp[1] = by - pad; bp[2] = bl_len + 2 * pad; bp[3] = bl_h + 2 * pad;
      /* desloc = bbox.bottom_left - spawn_bbox.bottom_left + (pad, pad); spawn_pos .+ desloc */
      double dx = bx - min_x + pad, dy = by - min_y + pad;
      for (int q = 0; q < nsp; q++) {
        o->spawn[2 * ((size_t)idx * nsp + q)] = c->spawn_pos[2 * q] + dx;
        o->spawn[2 * ((size_t)idx * nsp + q) + 1] = c->spawn_pos[2 * q + 1] + dy;
      }
    }
  if (s->has_chunks) { /* ChunksChecker(pos, chunks, bbox_pad_vec) */
    double g[4] = {bbox_pad[0], bbox_pad[1], bbox_pad[2], bbox_pad[3]}; /* global_bbox = sum(bbox_vec), src/configs.jl:68-77 */
    for (int k = 1; k < ns; k++) {
      const double *b = bbox_pad + 4 * k;
      double mx = fmax(g[0] + g[2], b[0] + b[2]), my = fmax(g[1] + g[3], b[1] + b[3]);
      double mnx = fmin(g[0], b[0]), mny = fmin(g[1], b[1]);
      g[0] = mnx; g[1] = mny; g[2] = mx - mnx; g[3] = my - mny;
    }
    const double tlx = s->p.grid_bl[0], tly = s->p.grid_bl[1] + s->p.grid_h; /* chunk_tl */
    int found = 0;
    int64_t sx = 0, ex = 0, sy = 0, ey = 0;
    for (int64_t i = 1; i <= s->num_cols; i++) {
      double x1 = tlx + (i - 1) * s->cl, x2 = x1 + s->cl;
      if (!found && x1 <= g[0] && g[0] <= x2) { found = 1; sx = i; }
      if (found && x1 >= g[0] + g[2]) { ex = i - 1; break; }
    }
    found = 0;
    for (int64_t i = 1; i <= s->num_rows; i++) {
      double y2 = tly - (i - 1) * s->ch, y1 = y2 - s->ch;
      if (!found && y1 <= g[1] + g[3] && g[1] + g[3] <= y2) { found = 1; sy = i; }
      if (found && y2 <= g[1]) { ey = i - 1; break; }
    }
    o->cells = (int32_t **)calloc((size_t)ns, sizeof(int32_t *));
    o->ncells = (int32_t *)calloc((size_t)ns, sizeof(int32_t));
    int64_t cap = (ex >= sx ? ex - sx + 1 : 0) * (ey >= sy ? ey - sy + 1 : 0);
    for (int k = 0; k < ns; k++) o->cells[k] = (int32_t *)calloc((size_t)(cap > 0 ? cap : 1), sizeof(int32_t));
    for (int64_t col = sx; col <= ex; col++)
      for (int64_t row = sy; row <= ey; row++) {
        /* get_chunk_rect, src/chunks.jl:50-59 */
        double rect[4] = {tlx + (col - 1) * s->cl, tly - row * s->ch, s->cl, s->ch};
        for (int k = 0; k < ns; k++)
          if (rect_intersect(rect, bbox_pad + 4 * k)) o->cells[k][o->ncells[k]++] = (int32_t)((row - 1) + s->num_rows * (col - 1));
      }
  }
  free(bbox_pad);
  return MAVI_OK;
}

int32_t mor_rings_set_sources(OrSystem *s, const MaviSourceSink *list, int32_t n, const uint8_t *ring_active,
                              const double *spawn_draws, int64_t n_draws) {
  if (s->p.dynamics != MAVI_DYN_RINGS || n < 0 || (n > 0 && !list)) return MAVI_ERR_BAD_PARAMS;
  const int64_t nr = s->rp.num_rings;
  if (ring_active) {
    s->var_rings = 1;
    s->ring_mask = (uint8_t *)calloc((size_t)nr + 1, 1);
    s->ring_ids = (int64_t *)calloc((size_t)nr + 1, sizeof(int64_t));
    s->ring_uids = (int64_t *)calloc((size_t)nr + 1, sizeof(int64_t));
    for (int64_t r = 0; r < nr; r++) {
      s->ring_mask[r] = ring_active[r] != 0;
      s->ring_ids[r] = r;
      s->ring_uids[r] = r + 1; /* uids = Vector(1:num_rings) */
    }
  } else if (n > 0) {
    snprintf(s->err, sizeof s->err, "sources / sinks need a RingsState with active_state (VarRingsIds)");
    return MAVI_ERR_BAD_PARAMS;
  }
  s->nsrc = n;
  s->src = (OrSource *)calloc((size_t)(n > 0 ? n : 1), sizeof(OrSource));
  for (int k = 0; k < n; k++) {
    int32_t st = build_source(s, &s->src[k], &list[k]);
    if (st) return st;
  }
  if (spawn_draws && n_draws > 0) {
    s->draws = dup_d(spawn_draws, n_draws);
    s->n_draws = n_draws;
  }
  s->draw_pos = 0;
  return MAVI_OK;
}

int32_t mor_rings_download_active(OrSystem *s, uint8_t *ring_active, int64_t *uids, int64_t *num_active) {
  if (s->p.dynamics != MAVI_DYN_RINGS) return MAVI_ERR_BAD_PARAMS;
  const int64_t nr = s->rp.num_rings;
  for (int64_t r = 0; r < nr; r++) {
    if (ring_active) ring_active[r] = s->var_rings ? s->ring_mask[r] : 1;
    if (uids) uids[r] = s->var_rings ? s->ring_uids[r] : r + 1;
  }
  if (num_active) *num_active = s->var_rings ? s->num_active : nr;
  return MAVI_OK;
}

/* update_area_empty!, src/rings/sources.jl:191-225 */
static void update_area_empty(OrSystem *s, OrSource *o) {
  for (int k = 0; k < o->nspawn; k++) {
    o->is_empty[k] = 1;
    const double *bbox = o->bbox + 4 * k;
    if (s->has_chunks) { /* ChunksChecker: the chunk lists of the LAST update_chunks! (stale by one step), current positions */
      for (int c = 0; c < o->ncells[k] && o->is_empty[k]; c++) {
        int64_t cell = o->cells[k][c];
        const int64_t *chunk = s->chunk_particles + cell * s->nc;
        for (int64_t q = 0; q < s->num_in_chunk[cell]; q++)
          if (inside_rect(s->pos + 2 * chunk[q], bbox, o->pad)) { o->is_empty[k] = 0; break; }
      }
    } else { /* PosChecker: get_ids(part_ids) as of the last update_ids! */
      for (int64_t q = 0; q < s->n_ids; q++)
        if (inside_rect(s->pos + 2 * s->ids[q], bbox, o->pad)) { o->is_empty[k] = 0; break; }
    }
  }
}

/* add_ring!, src/rings/states.jl:173-187 */
static int64_t add_ring(OrSystem *s, const double *pts, double pol) {
  const int64_t nm = s->rp.n_max;
  for (int64_t i = 0; i < s->rp.num_rings; i++) {
    if (s->ring_mask[i]) continue;
    memcpy(s->pos + 2 * i * nm, pts, sizeof(double) * 2 * (size_t)nm);
    s->second[i] = pol;
    s->ring_mask[i] = 1;
    int64_t mx = s->ring_uids[0];
    for (int64_t r = 1; r < s->rp.num_rings; r++)
      if (s->ring_uids[r] > mx) mx = s->ring_uids[r];
    s->ring_uids[i] = mx + 1;
    s->num_active += 1;
    return i;
  }
  return -1;
}

/* update_sources!, src/rings/integration.jl:353-358; process_source! / process_sink!, src/rings/sources.jl:229-263 */
static int32_t update_sources(OrSystem *s) {
  for (int k = 0; k < s->nsrc; k++) {
    OrSource *o = &s->src[k];
    if (o->kind == MAVI_SRC_SINK) {
      /* for ring_id in get_rings_ids(state): the view ids[1:num_active] is taken once, remove_ring! only lowers num_active */
      const int64_t nview = s->num_active;
      for (int64_t q = 0; q < nview; q++) {
        const int64_t ring = s->ring_ids[q];
        const double *cm = s->cms + 2 * ring;
        int in;
        if (o->sink_geom == MAVI_GEOM_RECT) in = inside_rect(cm, o->sink, 0.0);
        else {
          double ex = cm[0] - o->sink[0], ey = cm[1] - o->sink[1];
          in = ex * ex + ey * ey <= o->sink[2] * o->sink[2];
        }
        if (in) { /* remove_ring!, src/rings/states.jl:189-193 */
          s->ring_mask[ring] = 0;
          s->num_active -= 1;
        }
      }
      continue;
    }
    update_area_empty(s, o);
    for (int a = 0; a < o->nspawn; a++) {
      if (!o->is_empty[a]) continue;
      double pol = o->spawn_pol;
      if (isnan(pol)) { /* get_spawn_pol(::RandomPol): rand(rng, T) * 2 * pi — drawn BEFORE add_ring! looks for a slot */
        if (s->draw_pos >= s->n_draws) {
          snprintf(s->err, sizeof s->err, "spawn_draws exhausted after %lld draws", (long long)s->n_draws);
          return MAVI_ERR_BAD_PARAMS;
        }
        pol = s->draws[s->draw_pos++] * 2 * M_PI;
      }
      int64_t ring = add_ring(s, o->spawn + 2 * (size_t)a * o->nsp, pol);
      if (ring >= 0) { /* system.info.cms[ring_id] = sum(state.rings_pos[1:num_p, ring_id]) / num_p */
        int32_t np = ring_np(s, ring);
        const double *pts = s->pos + 2 * ring * s->rp.n_max;
        double sx = pts[0], sy = pts[1];
        for (int32_t i = 1; i < np; i++) { sx += pts[2 * i]; sy += pts[2 * i + 1]; }
        s->cms[2 * ring] = sx / np;
        s->cms[2 * ring + 1] = sy / np;
      }
    }
  }
  return MAVI_OK;
}


/* ------------------------------------------------------------------ invasions */

/* point_line_intersect, src/rings/integration.jl:379-420: does the ray from p towards +x cross the segment? */
static int point_line_intersect(const double *p, const double *l1, const double *l2) {
  const double dx = l2[0] - l1[0], dy = l2[1] - l1[1];
  double x_inter;
  if (dx == 0) {
    x_inter = l1[0];
    int check1 = x_inter > p[0];
    const double *ya = dy < 0 ? l2 : l1, *yb = dy < 0 ? l1 : l2;
    int check2 = ya[1] < p[1] && p[1] < yb[1];
    return check1 & check2;
  }
  if (dy == 0) return 0;
  const double c = dy * l1[0] - dx * l1[1];
  x_inter = (c + dx * p[1]) / dy;
  int check1 = x_inter > p[0];
  const double *ya = dy < 0 ? l2 : l1, *yb = dy < 0 ? l1 : l2;
  const double *xa = dx < 0 ? l2 : l1, *xb = dx < 0 ? l1 : l2;
  int check2 = (xa[0] < x_inter && x_inter < xb[0]) & (ya[1] < p[1] && p[1] < yb[1]);
  return check1 & check2;
}

static void inv_push(OrSystem *s, int64_t invasor, int64_t invaded, int64_t pid) {
  if (s->inv_n == s->inv_cap) {
    s->inv_cap = s->inv_cap ? 2 * s->inv_cap : 256;
    s->inv_list = (int32_t *)realloc(s->inv_list, sizeof(int32_t) * 3 * (size_t)s->inv_cap);
  }
  int32_t *e = s->inv_list + 3 * s->inv_n++;
  e[0] = (int32_t)invasor; e[1] = (int32_t)invaded; e[2] = (int32_t)pid;
}

/* polygons_intersect + find_invasions!, :422-469: the points of r1 inside r2 first, then the points of r2 inside r1 */
static void find_invasions(OrSystem *s, int64_t r1, int64_t r2) {
  const int64_t ring[2] = {r1, r2};
  for (int d = 0; d < 2; d++) {
    const int64_t a = ring[d], b = ring[1 - d];
    const double *pa = ring_points(s, a), *pb = ring_points(s, b);
    const int32_t na = ring_np(s, a), nb = ring_np(s, b);
    for (int32_t i = 0; i < na; i++) {
      int count = 0;
      for (int32_t e = 0; e < nb; e++) {
        const int32_t e2 = e == nb - 1 ? 0 : e + 1;
        count += point_line_intersect(pa + 2 * i, pb + 2 * e, pb + 2 * e2);
      }
      if (count % 2 != 0) inv_push(s, a, b, a * s->rp.n_max + i);
    }
  }
}

/* update_chunks!(r_chunks) for RingsChunksInfo (src/rings/integration.jl:18-23): the cms of the active rings */
static int32_t update_ring_chunks(OrSystem *s) {
  if (s->r_cols <= 0) return MAVI_OK;
  memset(s->r_num, 0, sizeof(int64_t) * (size_t)(s->r_rows * s->r_cols));
  for (int64_t q = 0; q < n_ring_ids(s); q++) {
    const int64_t ring = ring_id_at(s, q);
    const double x = s->cms[2 * ring], y = s->cms[2 * ring + 1];
    double rowf = mor_julia_div(-y + s->p.grid_bl[1] + s->p.grid_h, s->r_ch);
    double colf = mor_julia_div(x - s->p.grid_bl[0], s->r_cl);
    if (!(fabs(rowf) < 9.0e15) || !(fabs(colf) < 9.0e15)) return MAVI_ERR_OUT_OF_GRID;
    int64_t row_id = (int64_t)rowf + 1, col_id = (int64_t)colf + 1;
    row_id -= row_id == (s->r_rows + 1) ? 1 : 0;
    col_id -= col_id == (s->r_cols + 1) ? 1 : 0;
    if (row_id < 1 || row_id > s->r_rows || col_id < 1 || col_id > s->r_cols) {
      snprintf(s->err, sizeof s->err, "ring %lld: centre of mass out of the ring chunks (BoundsError in the reference)", (long long)ring);
      return MAVI_ERR_OUT_OF_GRID;
    }
    const int64_t cell = (row_id - 1) + s->r_rows * (col_id - 1);
    if (s->r_num[cell] >= s->r_nc) {
      snprintf(s->err, sizeof s->err, "ring cell %lld over capacity %lld (BoundsError in the reference)", (long long)cell, (long long)s->r_nc);
      return MAVI_ERR_CAPACITY;
    }
    s->r_particles[cell * s->r_nc + s->r_num[cell]++] = ring;
  }
  return MAVI_OK;
}

/* update_invasions! + check_invasions!, :471-520 */
static void update_invasions(OrSystem *s) {
  if (s->inv_steps <= 0) return;
  if (s->num_steps - s->inv_last_check < s->inv_steps) return;
  s->inv_last_check = s->num_steps;
  s->inv_n = 0;
  if (s->r_cols > 0) {
    for (int64_t col = 0; col < s->r_cols; col++)
      for (int64_t row = 0; row < s->r_rows; row++) {
        const int64_t cell = row + s->r_rows * col;
        const int64_t np = s->r_num[cell];
        const int64_t *chunk = s->r_particles + cell * s->r_nc;
        for (int64_t i = 0; i < np; i++) {
          for (int64_t j = i + 1; j < np; j++) find_invasions(s, chunk[i], chunk[j]);
          for (int k = 0; k < s->r_neigh_n[cell]; k++) {
            const int64_t nc = s->r_neigh[4 * cell + k];
            const int64_t *nei = s->r_particles + nc * s->r_nc;
            for (int64_t j = 0; j < s->r_num[nc]; j++) find_invasions(s, chunk[i], nei[j]);
          }
        }
      }
  } else {
    /* check_invasions!(system, ::Nothing): ring ids 1:num_rings of the active COUNT (sic), every pair once */
    const int64_t nr = n_ring_ids(s);
    for (int64_t r1 = 0; r1 < nr; r1++)
      for (int64_t r2 = r1 + 1; r2 < nr; r2++) find_invasions(s, r1, r2);
  }
}

int32_t mor_rings_set_invasions(OrSystem *s, int32_t steps_to_update, int32_t r_cols, int32_t r_rows) {
  if (s->p.dynamics != MAVI_DYN_RINGS || steps_to_update < 0 || r_cols < 0 || r_rows < 0) return MAVI_ERR_BAD_PARAMS;
  s->inv_steps = steps_to_update;
  s->inv_last_check = 0;
  s->inv_n = 0;
  s->r_cols = r_cols; s->r_rows = r_rows;
  if (r_cols > 0) {
    /* Chunks(num_cols, num_rows, bounding box, cms, minimum(get_ring_radius(dynamic_cfg))), src/rings/rings.jl:187-203 */
    s->r_cl = s->p.grid_len / (double)r_cols;
    s->r_ch = s->p.grid_h / (double)r_rows;
    double rr = INFINITY;
    for (int t = 0; t < s->rp.num_types; t++) {
      const double pr = s->rp.interaction[4 * ((int64_t)t * s->rp.num_types + t) + 2] / 2.0;
      const double ring_r = (pr * 2) / sqrt(2 * (1 - cos(2 * M_PI / s->rp.num_particles[t]))); /* src/rings/utils.jl:5-7 */
      if (ring_r < rr) rr = ring_r;
    }
    const double ncf = (ceil(0.5 * s->r_cl / rr) + 1) * (ceil(0.5 * s->r_ch / rr) + 1);
    s->r_nc = (int64_t)ceil(ncf * 2);
    s->r_particles = (int64_t *)calloc((size_t)(s->r_nc * r_cols * r_rows) + 1, sizeof(int64_t));
    s->r_num = (int64_t *)calloc((size_t)(r_cols * r_rows) + 1, sizeof(int64_t));
    NeighTab t;
    build_neighbors_tab(&t, r_rows, r_cols, s->p.spaces[0].wall == MAVI_WALL_PERIODIC);
    s->r_neigh = t.neigh;
    s->r_neigh_n = t.neigh_n;
  }
  return MAVI_OK;
}

/* info.invasions.list of the last check: triples (invasor ring, invaded ring, scalar particle id), 0-based */
int32_t mor_rings_download_invasions(OrSystem *s, int64_t *n, int32_t *triples, int64_t cap) {
  if (n) *n = s->inv_n;
  if (triples) memcpy(triples, s->inv_list, sizeof(int32_t) * 3 * (size_t)(s->inv_n < cap ? s->inv_n : cap));
  return MAVI_OK;
}

/* ------------------------------------------------------------------ steps */

static int64_t noise_stride(const OrSystem *s) {
  switch (s->p.dynamics) {
    case MAVI_DYN_SZABO: return s->n;
    case MAVI_DYN_RTP: return 2 * s->n;
    case MAVI_DYN_RINGS: return s->rp.num_rings;
    default: return 0;
  }
}

/* newton_step! / szabo_step! / rtp_step!, src/integration.jl:507-535; Rings step!, src/rings/integration.jl:522-543 */
static int32_t step_once(OrSystem *s, const double *noise) {
  int32_t st;
  if (s->p.dynamics == MAVI_DYN_RINGS) {
    update_cms(s);
    if ((st = update_sources(s))) return st;
    calc_active_ids(s); /* update_ids! */
    if ((st = mor_update_chunks(s))) return st; /* update_chunks_all!: particles, then the ring chunks */
    if ((st = update_ring_chunks(s))) return st;
    update_continuos_pos(s);
    update_invasions(s);
    mor_clean_forces(s);
    rings_forces(s);
    mor_walls_forces(s);
    rings_update(s, noise);
    mor_walls(s);
    s->num_steps += 1;
    s->time += s->p.dt;
    return MAVI_OK;
  }
  mor_clean_forces(s);
  if ((st = mor_update_chunks(s))) return st;
  mor_pair_forces(s);
  mor_walls_forces(s);
  if (s->p.dynamics == MAVI_DYN_SZABO) update_szabo(s, noise);
  else if (s->p.dynamics == MAVI_DYN_RTP) update_rtp(s, noise);
  else mor_update_verlet(s);
  mor_walls(s);
  update_time(s);
  return MAVI_OK;
}

int32_t mor_step(OrSystem *s, int64_t nsteps, const double *noise) {
  int64_t stride = noise_stride(s);
  for (int64_t k = 0; k < nsteps; k++) {
    int32_t st = step_once(s, noise ? noise + k * stride : NULL);
    if (st) return st;
  }
  return MAVI_OK;
}

int32_t mor_calc_forces(OrSystem *s) {
  mor_clean_forces(s);
  int32_t st = mor_update_chunks(s);
  if (st) return st;
  if (s->p.dynamics == MAVI_DYN_RINGS) {
    update_continuos_pos(s);
    rings_forces(s);
  } else {
    mor_pair_forces(s);
  }
  mor_walls_forces(s);
  return MAVI_OK;
}

int32_t mor_bin(OrSystem *s) { return mor_update_chunks(s); }

/* ------------------------------------------------------------------ quantities */

/* kinetic_energy src/quantities.jl:12-18 (all slots, mass 1); potential_energy LJ :46-66 (all pairs over
 * 1:N_active_count, no cutoff).  pe_mode 1: 4eps*sum over the cell-stencil pair set (not in the reference). */
int32_t mor_energies(OrSystem *s, int32_t pe_mode, double *ke, double *pe) {
  if (ke) {
    if (s->p.dynamics == MAVI_DYN_LJ || s->p.dynamics == MAVI_DYN_HARMTRUNC) {
      double acc = 0;
      for (int64_t i = 0; i < s->n; i++) acc += s->second[2 * i] * s->second[2 * i] + s->second[2 * i + 1] * s->second[2 * i + 1];
      *ke = acc / 2;
    } else *ke = NAN;
  }
  if (pe) {
    if (s->p.dynamics != MAVI_DYN_LJ) { *pe = NAN; return MAVI_OK; }
    const double sigma = s->p.dyn[0], epsilon = s->p.dyn[1];
    double pot = 0.0;
    if (pe_mode == 0) {
      const int64_t N = s->n_ids;
      for (int64_t i = 0; i < N; i++)
        for (int64_t j = i + 1; j < N; j++) {
          double dr[2];
          mor_calc_diff(s, s->pos + 2 * i, s->pos + 2 * j, dr);
          double dist = sqrt(dr[0] * dr[0] + dr[1] * dr[1]);
          pot += (pow(sigma / dist, 12) - pow(sigma / dist, 6));
        }
    } else {
      if (!s->has_chunks) return MAVI_ERR_BAD_PARAMS;
      for (int64_t cell = 0; cell < s->num_rows * s->num_cols; cell++) {
        int64_t np = s->num_in_chunk[cell];
        const int64_t *chunk = s->chunk_particles + cell * s->nc;
        for (int64_t i = 0; i < np; i++) {
          for (int64_t j = i + 1; j < np; j++) {
            double dr[2];
            mor_calc_diff(s, s->pos + 2 * chunk[i], s->pos + 2 * chunk[j], dr);
            double dist = sqrt(dr[0] * dr[0] + dr[1] * dr[1]);
            pot += (pow(sigma / dist, 12) - pow(sigma / dist, 6));
          }
          for (int k = 0; k < s->neigh_n[cell]; k++) {
            int64_t ncell = s->neigh[4 * cell + k];
            const int64_t *nch = s->chunk_particles + ncell * s->nc;
            for (int64_t j = 0; j < s->num_in_chunk[ncell]; j++) {
              double dr[2];
              mor_calc_diff(s, s->pos + 2 * chunk[i], s->pos + 2 * nch[j], dr);
              double dist = sqrt(dr[0] * dr[0] + dr[1] * dr[1]);
              pot += (pow(sigma / dist, 12) - pow(sigma / dist, 6));
            }
          }
        }
      }
    }
    *pe = pot * (4 * epsilon);
  }
  return MAVI_OK;
}

/* ------------------------------------------------------------------ lifetime / IO */

static int32_t validate(OrSystem *s) {
  const MaviParams *p = &s->p;
  if (p->struct_size != sizeof(MaviParams)) {
    snprintf(s->err, sizeof s->err, "MaviParams.struct_size %u != %zu", p->struct_size, sizeof(MaviParams));
    return MAVI_ERR_BAD_PARAMS;
  }
  if (p->dtype != MAVI_F64) {
    snprintf(s->err, sizeof s->err, "oracle is Float64 only");
    return MAVI_ERR_UNSUPPORTED;
  }
  if (p->n < 0 || p->n_spaces < 1 || p->n_spaces > MAVI_MAX_SPACES) return MAVI_ERR_BAD_PARAMS;
  for (int k = 0; k < p->n_spaces; k++)
    if (p->spaces[k].wall == MAVI_WALL_POTENTIAL && p->spaces[k].geom == MAVI_GEOM_RECT) {
      snprintf(s->err, sizeof s->err, "PotentialWalls on RectangleCfg: no signed_pos method in the reference");
      return MAVI_ERR_UNSUPPORTED;
    }
  if (p->dynamics == MAVI_DYN_RINGS) {
    if (!p->rings) return MAVI_ERR_BAD_PARAMS;
    if (p->rings->num_rings * p->rings->n_max != p->n) {
      snprintf(s->err, sizeof s->err, "n != num_rings*n_max");
      return MAVI_ERR_BAD_PARAMS;
    }
  }
  return MAVI_OK;
}

int32_t mor_create(const MaviParams *p, OrSystem **out) {
  OrSystem *s = (OrSystem *)calloc(1, sizeof(OrSystem));
  s->p = *p;
  *out = s;
  int32_t st = validate(s);
  if (st) return st;
  for (int k = 0; k < p->n_spaces; k++) {
    int nl = p->spaces[k].n_lines;
    if (p->spaces[k].geom == MAVI_GEOM_LINES && nl > 0) {
      s->lines[k] = (MaviLine *)malloc(sizeof(MaviLine) * (size_t)nl);
      memcpy(s->lines[k], p->spaces[k].lines, sizeof(MaviLine) * (size_t)nl);
      s->p.spaces[k].lines = s->lines[k];
    }
  }
  s->n = p->n;
  int64_t n_second = 2 * s->n;
  if (p->dynamics == MAVI_DYN_RINGS) {
    const MaviRingsParams *r = p->rings;
    s->rp = *r;
    int nt = r->num_types;
    s->rp.p0 = dup_d(r->p0, nt); s->rp.relax_time = dup_d(r->relax_time, nt); s->rp.vo = dup_d(r->vo, nt);
    s->rp.mobility = dup_d(r->mobility, nt); s->rp.rot_diff = dup_d(r->rot_diff, nt); s->rp.k_area = dup_d(r->k_area, nt);
    s->rp.k_spring = dup_d(r->k_spring, nt); s->rp.l_spring = dup_d(r->l_spring, nt);
    int32_t *np = (int32_t *)malloc(sizeof(int32_t) * (size_t)nt);
    memcpy(np, r->num_particles, sizeof(int32_t) * (size_t)nt);
    s->rp.num_particles = np;
    s->rp.interaction = dup_d(r->interaction, 4 * nt * nt);
    if (r->types) {
      int32_t *ty = (int32_t *)malloc(sizeof(int32_t) * (size_t)r->num_rings);
      memcpy(ty, r->types, sizeof(int32_t) * (size_t)r->num_rings);
      s->rp.types = ty;
    }
    s->p.rings = &s->rp;
    n_second = r->num_rings;
    s->cont_pos = (double *)calloc((size_t)(2 * s->n + 2), sizeof(double));
    s->areas = (double *)calloc((size_t)r->num_rings + 1, sizeof(double));
    s->cms = (double *)calloc((size_t)(2 * r->num_rings) + 2, sizeof(double));
  } else if (p->dynamics == MAVI_DYN_SZABO || p->dynamics == MAVI_DYN_RTP) {
    n_second = s->n;
  }
  s->pos = (double *)calloc((size_t)(2 * s->n + 2), sizeof(double));
  s->second = (double *)calloc((size_t)n_second + 2, sizeof(double));
  s->mask = (uint8_t *)malloc((size_t)s->n + 1);
  s->ids = (int64_t *)malloc(sizeof(int64_t) * ((size_t)s->n + 1));
  s->nthreads = 1;
  s->forces = (double **)malloc(sizeof(double *));
  s->forces[0] = (double *)calloc((size_t)(2 * s->n + 2), sizeof(double));
  return chunks_init(s);
}

void mor_set_threads(OrSystem *s, int32_t nthreads) {
  if (nthreads < 1) nthreads = 1;
  for (int t = 1; t < s->nthreads; t++) free(s->forces[t]);
  s->forces = (double **)realloc(s->forces, sizeof(double *) * (size_t)nthreads);
  for (int t = 1; t < nthreads; t++) s->forces[t] = (double *)calloc((size_t)(2 * s->n + 2), sizeof(double));
  s->nthreads = nthreads;
}

void mor_destroy(OrSystem *s) {
  if (!s) return;
  for (int k = 0; k < MAVI_MAX_SPACES; k++) free(s->lines[k]);
  if (s->p.dynamics == MAVI_DYN_RINGS && s->p.rings == &s->rp) {
    free((void *)s->rp.p0); free((void *)s->rp.relax_time); free((void *)s->rp.vo); free((void *)s->rp.mobility);
    free((void *)s->rp.rot_diff); free((void *)s->rp.k_area); free((void *)s->rp.k_spring); free((void *)s->rp.l_spring);
    free((void *)s->rp.num_particles); free((void *)s->rp.interaction); free((void *)s->rp.types);
  }
  free(s->pos); free(s->second); free(s->mask); free(s->ids);
  if (s->forces) {
    for (int t = 0; t < s->nthreads; t++) free(s->forces[t]);
    free(s->forces);
  }
  free(s->chunk_particles); free(s->num_in_chunk); free(s->neigh); free(s->neigh_n);
  free(s->cont_pos); free(s->areas); free(s->cms);
  free(s->pn_count); free(s->pn_list);
  free(s->inv_list); free(s->r_particles); free(s->r_num); free(s->r_neigh); free(s->r_neigh_n);
  free(s->ring_mask); free(s->ring_ids); free(s->ring_uids); free(s->draws);
  for (int k = 0; k < s->nsrc; k++) {
    OrSource *o = &s->src[k];
    free(o->bbox); free(o->spawn); free(o->is_empty);
    if (o->cells) for (int a = 0; a < o->nspawn; a++) free(o->cells[a]);
    free(o->cells); free(o->ncells);
  }
  free(s->src);
  free(s);
}

const char *mor_last_error(OrSystem *s) { return s->err; }

/* check_inside, src/space_checks.jl:9-61 — only Rectangle / Circle geometries have a real check;
 * ManyGeometries falls to the generic method that returns [] (no check). */
static int32_t check_inside(OrSystem *s) {
  if (s->p.n_spaces != 1) return MAVI_OK;
  const MaviSpace *sp = &s->p.spaces[0];
  for (int64_t q = 0; q < s->n_ids; q++) {
    int64_t i = s->ids[q];
    double x = s->pos[2 * i], y = s->pos[2 * i + 1];
    int out = 0;
    if (sp->geom == MAVI_GEOM_RECT) {
      double trx = sp->rect_bl[0] + sp->rect_len, try_ = sp->rect_bl[1] + sp->rect_h;
      out = (x < sp->rect_bl[0]) || (y < sp->rect_bl[1]) || (x > trx) || (y > try_);
    } else if (sp->geom == MAVI_GEOM_CIRCLE) {
      out = (x * x + y * y) > sp->circ_radius * sp->circ_radius; /* ignores the centre (sic) */
    }
    if (out) {
      snprintf(s->err, sizeof s->err, "Particles with ids=[%lld, ...] outside space.", (long long)(i + 1));
      return MAVI_ERR_OUTSIDE_SPACE;
    }
  }
  return MAVI_OK;
}

/* System ctor tail (src/systems.jl:73-114): ids, inside check, first update_chunks!;
 * RingsSystem ctor tail (src/rings/rings.jl:280-288): prime continuos_pos, cms, chunks, forces!. */
int32_t mor_upload_state(OrSystem *s, const double *pos, const double *second, const uint8_t *mask, int64_t n) {
  if (n != s->n) return MAVI_ERR_BAD_PARAMS;
  memcpy(s->pos, pos, sizeof(double) * 2 * (size_t)n);
  int64_t n_second = 2 * n;
  if (s->p.dynamics == MAVI_DYN_RINGS) n_second = s->rp.num_rings;
  else if (s->p.dynamics == MAVI_DYN_SZABO || s->p.dynamics == MAVI_DYN_RTP) n_second = n;
  if (second) memcpy(s->second, second, sizeof(double) * (size_t)n_second);
  s->n_ids = 0;
  if (s->p.dynamics == MAVI_DYN_RINGS) {
    if (s->var_rings) {
      calc_active_ids(s); /* RingsState ctor: update_ids!(state), src/rings/states.jl:121 */
    } else {
      /* FixRingsIds, src/rings/states.jl:45-61 */
      for (int64_t ring = 0; ring < s->rp.num_rings; ring++) {
        int32_t np = ring_np(s, ring);
        for (int64_t q = 0; q < s->rp.n_max; q++) {
          s->mask[ring * s->rp.n_max + q] = q < np;
          if (q < np) s->ids[s->n_ids++] = ring * s->rp.n_max + q;
        }
      }
    }
  } else {
    /* ParticleIds / eachindex(pos), src/states.jl:27-70 */
    for (int64_t i = 0; i < n; i++) {
      s->mask[i] = mask ? (mask[i] != 0) : 1;
      if (s->mask[i]) s->ids[s->n_ids++] = i;
    }
  }
  int32_t st = check_inside(s);
  if (st) return st;
  if (s->p.dynamics == MAVI_DYN_RINGS) {
    update_continuos_pos(s);
    update_cms(s);
    if ((st = mor_update_chunks(s))) return st;
    mor_clean_forces(s);
    rings_forces(s);
    return MAVI_OK;
  }
  return mor_update_chunks(s);
}

int32_t mor_download_state(OrSystem *s, double *pos, double *second) {
  if (pos) memcpy(pos, s->pos, sizeof(double) * 2 * (size_t)s->n);
  if (second) {
    int64_t n_second = 2 * s->n;
    if (s->p.dynamics == MAVI_DYN_RINGS) n_second = s->rp.num_rings;
    else if (s->p.dynamics == MAVI_DYN_SZABO || s->p.dynamics == MAVI_DYN_RTP) n_second = s->n;
    memcpy(second, s->second, sizeof(double) * (size_t)n_second);
  }
  return MAVI_OK;
}

int32_t mor_download_forces(OrSystem *s, double *f) {
  memcpy(f, s->forces[0], sizeof(double) * 2 * (size_t)s->n);
  return MAVI_OK;
}

int32_t mor_download_cells(OrSystem *s, int32_t *cell_of_particle, int32_t *counts) {
  if (!s->has_chunks) return MAVI_ERR_BAD_PARAMS;
  int64_t cells = s->num_rows * s->num_cols;
  if (counts)
    for (int64_t c = 0; c < cells; c++) counts[c] = (int32_t)s->num_in_chunk[c];
  if (cell_of_particle) {
    for (int64_t i = 0; i < s->n; i++) cell_of_particle[i] = -1;
    for (int64_t c = 0; c < cells; c++)
      for (int64_t k = 0; k < s->num_in_chunk[c]; k++) cell_of_particle[s->chunk_particles[c * s->nc + k]] = (int32_t)c;
  }
  return MAVI_OK;
}

int32_t mor_download_cell_lists(OrSystem *s, int32_t *start, int32_t *ids) {
  if (!s->has_chunks) return MAVI_ERR_BAD_PARAMS;
  int64_t cells = s->num_rows * s->num_cols, acc = 0;
  for (int64_t c = 0; c < cells; c++) {
    if (start) start[c] = (int32_t)acc;
    if (ids)
      for (int64_t k = 0; k < s->num_in_chunk[c]; k++) ids[acc + k] = (int32_t)s->chunk_particles[c * s->nc + k];
    acc += s->num_in_chunk[c];
  }
  if (start) start[cells] = (int32_t)acc;
  return MAVI_OK;
}

int32_t mor_cell_neighbors(OrSystem *s, int32_t cell, int32_t *out4, int32_t *n) {
  if (!s->has_chunks || cell < 0 || cell >= s->num_rows * s->num_cols) return MAVI_ERR_BAD_PARAMS;
  *n = s->neigh_n[cell];
  for (int k = 0; k < *n; k++) out4[k] = s->neigh[4 * cell + k];
  return MAVI_OK;
}

int64_t mor_chunk_capacity(OrSystem *s) { return s->nc; }

/* NeighborsCfg for the particle contact lists (RingsSystem p_neighbors_cfg); call before mor_upload_state to have the
 * constructor's forces! fill them (src/rings/rings.jl:280-288) */
int32_t mor_rings_set_neighbors(OrSystem *s, int32_t mode, int32_t type_all, double tol) {
  if (s->p.dynamics != MAVI_DYN_RINGS || mode < MAVI_NEIGH_OFF || mode > MAVI_NEIGH_LIST) return MAVI_ERR_BAD_PARAMS;
  free(s->pn_count); free(s->pn_list);
  s->pn_count = NULL; s->pn_list = NULL;
  s->pn_mode = mode; s->pn_all = type_all ? 1 : 0; s->pn_tol = tol; s->pn_overflow = 0;
  if (mode) s->pn_count = (int32_t *)calloc((size_t)s->n + 1, sizeof(int32_t));
  if (mode == MAVI_NEIGH_LIST) {
    s->pn_list = (int32_t *)malloc(sizeof(int32_t) * ((size_t)s->n * MAVI_NEIGH_MAX + 1));
    memset(s->pn_list, 0xff, sizeof(int32_t) * ((size_t)s->n * MAVI_NEIGH_MAX + 1));
  }
  return MAVI_OK;
}

/* get_neigh_count / get_neigh_list, src/rings/neighbors.jl:58-62 (lists in the reference's append order, -1 padded) */
int32_t mor_rings_download_neighbors(OrSystem *s, int32_t *count, int32_t *list) {
  if (!s->pn_mode) return MAVI_ERR_BAD_PARAMS;
  if (list && s->pn_mode != MAVI_NEIGH_LIST) return MAVI_ERR_BAD_PARAMS;
  if (s->pn_overflow) {
    snprintf(s->err, sizeof s->err, "a particle has more than %d contacts: BoundsError in the reference", MAVI_NEIGH_MAX);
    return MAVI_ERR_CAPACITY;
  }
  if (count) memcpy(count, s->pn_count, sizeof(int32_t) * (size_t)s->n);
  if (list) memcpy(list, s->pn_list, sizeof(int32_t) * (size_t)s->n * MAVI_NEIGH_MAX);
  return MAVI_OK;
}

int32_t mor_rings_download_info(OrSystem *s, double *areas, double *cms, double *cont_pos) {
  if (s->p.dynamics != MAVI_DYN_RINGS) return MAVI_ERR_BAD_PARAMS;
  if (areas) memcpy(areas, s->areas, sizeof(double) * (size_t)s->rp.num_rings);
  if (cms) memcpy(cms, s->cms, sizeof(double) * 2 * (size_t)s->rp.num_rings);
  if (cont_pos) memcpy(cont_pos, s->cont_pos, sizeof(double) * 2 * (size_t)s->n);
  return MAVI_OK;
}

int32_t mor_get_time(OrSystem *s, int64_t *num_steps, double *time) {
  if (num_steps) *num_steps = s->num_steps;
  if (time) *time = s->time;
  return MAVI_OK;
}
