// api.cu — the C ABI of libmavi_cuda.so (include/mavi.h): handle lifetime, state movement, step orchestration.
// Reference citations are relative to /root/reference/.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <utility>
#include <vector>

#include "handle.cuh"
#include "api_decl.inc"

using namespace MAVI_NS;

namespace MAVI_NS {

void Handle::set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err, sizeof err, fmt, ap);
  va_end(ap);
}

#define CUDA_TRY(h, expr)                                                                       \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      (h)->set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #expr); \
      return MAVI_ERR_CUDA;                                                                     \
    }                                                                                           \
  } while (0)

// largest x with fl(sqrt(x)) <= d  ->  (sqrt(r2) > d)  <=>  (r2 > x), exactly
static real sqrt_le_threshold(real d) {
  if (!(d > 0)) return d < 0 ? real(-1) : real(0);  // sqrt(r2) > d for every r2 > 0 (d = 0) or for every r2 at all (d < 0); NaN -> 0
  if (std::isinf(d)) return d;                      // "no cutoff"
  real x = d * d;
  while (std::sqrt(x) <= d) x = std::nextafter(x, INFINITY);
  while (std::sqrt(x) > d) x = std::nextafter(x, -INFINITY);
  return x;
}
// smallest x with fl(sqrt(x)) >= d  ->  (sqrt(r2) < d)  <=>  (r2 < x), exactly
static real sqrt_ge_threshold(real d) {
  if (!(d > 0)) return real(0);  // sqrt(r2) < d never holds
  if (std::isinf(d)) return d;
  real x = d * d;
  while (std::sqrt(x) >= d && x > 0) x = std::nextafter(x, -INFINITY);
  while (std::sqrt(x) < d) x = std::nextafter(x, INFINITY);
  return x;
}

template <typename T>
static int dev_alloc(Handle *h, T **ptr, size_t count) {
  *ptr = nullptr;
  if (count == 0) count = 1;
  CUDA_TRY(h, cudaMalloc((void **)ptr, count * sizeof(T)));
  h->allocs.push_back((void *)*ptr);
  return MAVI_OK;
}

template <typename T>
static int dev_upload(Handle *h, const T **dst, const T *src, size_t count) {
  T *d;
  int st = dev_alloc(h, &d, count);
  if (st) return st;
  CUDA_TRY(h, cudaMemcpy(d, src, count * sizeof(T), cudaMemcpyHostToDevice));
  *dst = d;
  return MAVI_OK;
}

static int validate_and_lower(Handle *h, const MaviParams *mp) {
  DevParams &p = h->p;
  if (mp->struct_size != sizeof(MaviParams)) {
    h->set_error("MaviParams.struct_size %u != %zu (ABI mismatch)", mp->struct_size, sizeof(MaviParams));
    return MAVI_ERR_BAD_PARAMS;
  }
  if (mp->dtype != (MAVI_REAL_IS_F32 ? MAVI_F32 : MAVI_F64)) {  // capi.cu routes by dtype
    h->set_error("MaviParams.dtype %d reached the %s build", mp->dtype, MAVI_REAL_IS_F32 ? "Float32" : "Float64");
    return MAVI_ERR_BAD_PARAMS;
  }
  if (mp->n < 0 || mp->n > 0x7fffffff || mp->n_spaces < 1 || mp->n_spaces > MAVI_MAX_SPACES) {
    h->set_error("bad n / n_spaces");
    return MAVI_ERR_BAD_PARAMS;
  }
  memset(&p, 0, sizeof p);
  p.n = (int)mp->n;
  p.n_count = p.n;
  p.dynamics = mp->dynamics;
  for (int i = 0; i < 8; i++) p.dyn[i] = mp->dyn[i];
  p.particle_radius = mp->particle_radius;
  p.dt = mp->dt;
  h->dt_host = mp->dt;
  p.term = mp->dt * mp->dt / 2;  // dt^2/2
  p.hdt = mp->dt / 2;
  p.rng_mode = mp->rng_mode;
  p.seed = mp->seed;

  // spaces
  p.n_spaces = mp->n_spaces;
  for (int k = 0; k < mp->n_spaces; k++) {
    const MaviSpace &s = mp->spaces[k];
    DevSpace &d = p.spaces[k];
    d.wall = s.wall;
    d.geom = s.geom;
    d.rect_bl[0] = s.rect_bl[0]; d.rect_bl[1] = s.rect_bl[1];
    d.rect_sz[0] = s.rect_len; d.rect_sz[1] = s.rect_h;
    d.cc[0] = s.circ_center[0]; d.cc[1] = s.circ_center[1];
    d.cr = s.circ_radius;
    d.pot_kind = s.pot_kind;
    for (int i = 0; i < 4; i++) d.pot[i] = s.pot[i];
    d.pot_mode = s.pot_mode;
    d.n_pot_types = 0;
    if (s.wall == MAVI_WALL_POTENTIAL && s.n_pot_types != 0) {
      // PotentialVector: get_particle_type exists for RingsState only (src/rings/states.jl:148); one entry per ring type
      if (mp->dynamics != MAVI_DYN_RINGS || !mp->rings || s.n_pot_types != mp->rings->num_types ||
          s.n_pot_types < 0 || s.n_pot_types > MAVI_MAX_POT_TYPES) {
        h->set_error("PotentialVector needs a Mavi.Rings state with types and one potential per ring type (<= %d), got %d",
                     MAVI_MAX_POT_TYPES, s.n_pot_types);
        return MAVI_ERR_BAD_PARAMS;
      }
      d.n_pot_types = s.n_pot_types;
      for (int t = 0; t < s.n_pot_types; t++)
        for (int i = 0; i < 4; i++) d.pot_t[t][i] = (real)s.pot_types[t][i];
    }
    d.n_lines = 0;
    d.lines = nullptr;
    if (s.wall == MAVI_WALL_POTENTIAL) {
      if (s.geom == MAVI_GEOM_RECT) {
        h->set_error("PotentialWalls on a RectangleCfg: the reference has no signed_pos method for it");
        return MAVI_ERR_UNSUPPORTED;
      }
      p.has_force_walls = 1;
    }
    if (s.geom == MAVI_GEOM_LINES && s.n_lines > 0) {
      std::vector<DevLine> lines(s.n_lines);
      for (int l = 0; l < s.n_lines; l++) {
        // Line2D ctor, src/configs.jl:102-117
        const MaviLine &ml = s.lines[l];
        DevLine &dl = lines[l];
        double d0 = ml.p2[0] - ml.p1[0], d1 = ml.p2[1] - ml.p1[1];
        double norm = std::sqrt(d0 * d0 + d1 * d1);
        dl.p1[0] = ml.p1[0]; dl.p1[1] = ml.p1[1];
        dl.p2[0] = ml.p2[0]; dl.p2[1] = ml.p2[1];
        dl.normal[0] = -d1 / norm; dl.normal[1] = d0 / norm;
        dl.tangent[0] = d0 / norm; dl.tangent[1] = d1 / norm;
        dl.length = norm;
      }
      int st = dev_upload(h, &d.lines, lines.data(), lines.size());
      if (st) return st;
      d.n_lines = s.n_lines;
    }
  }
  const MaviSpace &m0 = mp->spaces[0];
  p.periodic = (m0.wall == MAVI_WALL_PERIODIC && m0.geom == MAVI_GEOM_RECT) ? 1 : 0;
  p.size[0] = m0.rect_len; p.size[1] = m0.rect_h;
  p.half[0] = m0.rect_len / 2; p.half[1] = m0.rect_h / 2;
  p.wall_fast = (mp->n_spaces == 1 && p.periodic) ? 1 : 0;
  p.wall_ctr[0] = m0.rect_bl[0] + p.half[0]; p.wall_ctr[1] = m0.rect_bl[1] + p.half[1];

  // bounding box of the geometry (get_chunks, src/systems.jl:14-28): also the frame of the ring-level chunks
  p.grid_bl[0] = mp->grid_bl[0]; p.grid_bl[1] = mp->grid_bl[1];
  p.grid_h = mp->grid_h;
  h->grid_len_host = mp->grid_len;
  // Chunks ctor, src/chunks.jl:26-40
  if (mp->num_cols > 0) {
    if (mp->num_rows < 1) {
      h->set_error("num_rows must be >= 1");
      return MAVI_ERR_BAD_PARAMS;
    }
    const bool per_wall = m0.wall == MAVI_WALL_PERIODIC;
    if (per_wall && (mp->num_cols < 2 || mp->num_rows < 2)) {
      h->set_error("periodic cell grid needs >= 2 rows and columns (a 1-wide periodic grid lists a cell as its own "
                   "neighbour in the reference and yields NaN self-pairs)");
      return MAVI_ERR_BAD_PARAMS;
    }
    if (!per_wall && mp->num_cols < 2) {
      h->set_error("walled cell grid needs num_cols >= 2 (the reference indexes column 0 otherwise)");
      return MAVI_ERR_BAD_PARAMS;
    }
    long long cells = (long long)mp->num_cols * mp->num_rows;
    if (cells > 0x7ffffff0LL) {
      h->set_error("too many cells");
      return MAVI_ERR_BAD_PARAMS;
    }
    p.num_cols = mp->num_cols;
    p.num_rows = mp->num_rows;
    p.num_cells = (int)cells;
    p.wrap_cols = p.wrap_rows = per_wall ? 1 : 0;
    p.grid_bl[0] = mp->grid_bl[0]; p.grid_bl[1] = mp->grid_bl[1];
    p.grid_h = mp->grid_h;
    p.cl = mp->grid_len / (double)mp->num_cols;
    p.ch = mp->grid_h / (double)mp->num_rows;
    // interior cells may skip the minimum image when two cell widths + two guarded drifts stay below size/2
    p.fast_interior = (p.periodic && mp->num_cols >= 8 && mp->num_rows >= 8 && 4.0 * p.cl <= p.half[0] &&
                       4.0 * p.ch <= p.half[1]) ? 1 : 0;
  }

  if (mp->world > 1 || (mp->flags & MAVI_FLAG_SLAB_SELF)) {
    int st = slab_configure(h, mp);
    if (st) return st;
  }

  switch (mp->dynamics) {
    case MAVI_DYN_LJ:
      p.lj_sig2 = mp->dyn[0] * mp->dyn[0];
      p.lj_24eps = 24.0 * mp->dyn[1];
      p.lj_c48 = 48.0 * mp->dyn[1] / p.lj_sig2;
      p.lj_c24 = 24.0 * mp->dyn[1] / p.lj_sig2;
      break;
    case MAVI_DYN_HARMTRUNC:
      p.cut2 = sqrt_le_threshold(mp->dyn[3]);
      p.eq2_lo = sqrt_ge_threshold(mp->dyn[2]);
      p.harm_inv_deq = 1.0 / mp->dyn[2];
      break;
    case MAVI_DYN_SZABO:
      p.cut2 = sqrt_le_threshold(mp->dyn[6]);
      p.szabo_eq2_hi = sqrt_le_threshold(mp->dyn[5]);
      p.szabo_fadh = mp->dyn[4] / mp->dyn[5];
      p.szabo_frep = mp->dyn[3] / (mp->dyn[6] - mp->dyn[5]);
      p.szabo_inv_tau = 1.0 / mp->dyn[2];
      p.szabo_namp = std::sqrt(2.0 * mp->dyn[7] * mp->dt);
      break;
    case MAVI_DYN_RTP:
      p.lj_sig2 = mp->dyn[1] * mp->dyn[1];
      p.lj_24eps = 24.0 * mp->dyn[2];
      p.lj_c48 = 48.0 * mp->dyn[2] / p.lj_sig2;
      p.lj_c24 = 24.0 * mp->dyn[2] / p.lj_sig2;
      p.cut2 = sqrt_le_threshold(std::pow(2.0, 1.0 / 6.0) * mp->dyn[1]);
      break;
    case MAVI_DYN_RINGS:
      return rings_lower(h, mp);
    default:
      h->set_error("unknown dynamics %d", mp->dynamics);
      return MAVI_ERR_BAD_PARAMS;
  }
  return MAVI_OK;
}

// magic number for x / d, valid for 0 <= x < 2^31 (round-up method with a 32-bit multiplier)
static void magic_div(unsigned int d, unsigned int *mul, unsigned int *shr);
void mavi_magic_div(unsigned int d, unsigned int *mul, unsigned int *shr) { magic_div(d, mul, shr); }
static void magic_div(unsigned int d, unsigned int *mul, unsigned int *shr) {
  if (d <= 1) { *mul = 0; *shr = 0; return; }  // d == 1: identity (fastdiv special-cases mul == 0)
  unsigned int l = 0;
  while ((1ull << l) < d) ++l;  // l = ceil(log2 d)
  unsigned long long m = ((1ull << (31 + l)) + d - 1) / d;  // ceil(2^(31+l) / d) fits in 32 bits
  *mul = (unsigned int)m;
  *shr = l - 1;
}

static void dev_free(Handle *h, void *ptr) {
  if (!ptr) return;
  for (auto &q : h->allocs)
    if (q == ptr) q = nullptr;
  cudaFree(ptr);
}

// handle-lifetime arrays: staging (dense, n entries), control words, scratch that does not depend on the tile capacity
static int allocate(Handle *h) {
  const DevParams &p = h->p;
  DevArrays &a = h->a;
  // slab mode: the owned count changes with migration -> head room in the dense arrays
  h->n_cap = p.slab ? (int)(p.n * 1.25) + 4096 : p.n;
  const size_t n = (size_t)h->n_cap;
  int st;
  memset(&a, 0, sizeof a);
  if ((st = dev_alloc(h, &a.st_pos, n))) return st;
  if ((st = dev_alloc(h, &a.st_force, n))) return st;
  if ((st = dev_alloc(h, &a.st_id, n))) return st;
  if ((st = dev_alloc(h, &a.st_cell, n))) return st;
  if (h->second_kind == SECOND_VEL) {
    if ((st = dev_alloc(h, &a.st_vel, n))) return st;
  } else {
    if ((st = dev_alloc(h, &a.st_ang, h->second_kind == SECOND_RING_POL ? (size_t)p.rings.num_rings : n))) return st;
  }
  const size_t nc = (size_t)p.num_cells + 2;
  if ((st = dev_alloc(h, &a.count, nc))) return st;
  if ((st = dev_alloc(h, &a.flags, FLAG_COUNT))) return st;
  if ((st = dev_alloc(h, &a.fix_idx, n))) return st;
  if ((st = dev_alloc(h, &a.fix_pos, n))) return st;
  if ((st = dev_alloc(h, &a.reduce_buf, 4096))) return st;
  if ((st = dev_alloc(h, &a.cta_first, n / RPB + 2))) return st;
  if (p.num_cells > 0) {  // cell-edge tables (filled by alloc_state, once the tile geometry is known)
    if ((st = dev_alloc(h, &a.edge_x, (size_t)p.num_cols + 2))) return st;
    if ((st = dev_alloc(h, &a.edge_y, (size_t)((p.num_rows + MAVI_TR - 1) / MAVI_TR) * MAVI_TR))) return st;
  }
  if (p.slab) {  // emigrant / immigrant records (slab.cu)
    a.em_cap = p.num_rows / 2 > 1024 ? p.num_rows / 2 : 1024;
    for (int d = 0; d < 2; d++) {
      if ((st = dev_alloc(h, &a.em_send[d], (size_t)a.em_cap))) return st;
      if ((st = dev_alloc(h, &a.em_recv[d], (size_t)a.em_cap))) return st;
    }
  }
  CUDA_TRY(h, cudaMallocHost((void **)&h->flags_host, FLAG_COUNT * sizeof(int)));
  CUDA_TRY(h, cudaMemsetAsync(a.flags, 0, FLAG_COUNT * sizeof(int), h->stream));
  CUDA_TRY(h, cudaMemsetAsync(a.count, 0, nc * sizeof(int), h->stream));
  CUDA_TRY(h, cudaMemsetAsync(a.st_force, 0, n * sizeof(real2), h->stream));
  CUDA_TRY(h, cudaMemsetAsync(a.cta_first, 0, (n / RPB + 2) * sizeof(int), h->stream));
  if (h->second_kind == SECOND_RING_POL) return rings_allocate(h);
  return MAVI_OK;
}

// slot-indexed state for a given tile capacity (re-done when a tile overflows)
int Handle::alloc_state(int n_active, int cap) {
  DevArrays &A = a;
  void *old[] = {A.pos[0], A.pos[1], A.vel, A.ang, A.idflag, A.cell, A.force, A.force_old, A.tstart, A.tile_prefix,
                 A.perm, A.scan_partials, A.tile_dirty, A.dirty_list, A.inbox_cnt, A.inbox, A.mv_src, A.mv_pos,
                 A.mv_second, A.mv_force, A.mv_id, A.mv_cell, A.chg};
  // slab mode: the owned count changes with every migration, but every array size depends only on the local grid, the
  // tile capacity and the staging capacity -> a re-upload with an unchanged capacity keeps the allocations
  const bool reuse = p.slab && p.num_cells > 0 && A.pos[0] != nullptr && cap == p.cap && cap > 0;
  if (!reuse)
    for (void *q : old) dev_free(this, q);
  p.n_active = p.num_cells > 0 ? n_active : 0;
  p.n_count = p.slab ? slab.n_global : n_active;
  if (p.num_cells > 0) {
    p.tpc = (p.num_rows + MAVI_TR - 1) / MAVI_TR;
    p.nt = p.num_cols * p.tpc;
    p.cap = cap;
    long long slots = (long long)p.nt * cap + (p.slab ? 0 : (p.n - n_active));
    if (slots > 0x7ffffff0LL) {
      set_error("tile layout needs %lld slots (> 2^31)", slots);
      return MAVI_ERR_BAD_PARAMS;
    }
    p.tail_base = p.nt * cap;
    if (!p.slab) {
      p.gcols = p.num_cols;
      p.ord_cols = p.num_cols;
      p.ord_col0 = 0;
    }
    p.nt_ord = p.ord_cols * p.tpc;
    magic_div((unsigned int)p.num_rows, &p.rows_mul, &p.rows_shr);
    magic_div((unsigned int)p.tpc, &p.tpc_mul, &p.tpc_shr);
    magic_div((unsigned int)p.ord_cols, &p.cols_mul, &p.cols_shr);
    p.inbox_cap = cap;  // a tile can never receive more than it can hold: inbox overflow implies tile overflow
    p.mv_cap = n_cap / 4 > 4096 ? n_cap / 4 : 4096;
    p.chg_cap = 2 * p.mv_cap;
    {  // CTA = blk_cols tiles of one tile row, about 1000 particles (kernels.cu, tile-block force kernels)
      const double mean_tile = (double)(n_active > 0 ? n_active : 1) / (double)((long long)p.ord_cols * p.tpc);
      int g = (int)(1000.0 / (mean_tile > 1.0 ? mean_tile : 1.0));
      p.blk_cols = g < 1 ? 1 : (g > 30 ? 30 : g);
      if (flags_cfg & MAVI_FLAG_SMALL_BLOCKS) p.blk_cols = 3;
      p.blk_per_row = (p.ord_cols + p.blk_cols - 1) / p.blk_cols;
      p.blk_mode = 0;
    }
  } else {
    p.tpc = p.nt = p.cap = p.nt_ord = 0;
    p.tail_base = 0;
    p.inbox_cap = p.mv_cap = p.chg_cap = 0;
    p.blk_cols = p.blk_per_row = p.blk_mode = 0;
  }
  ns = (size_t)p.tail_base + (size_t)(p.slab ? 0 : (p.n - p.n_active));
  int st;
  launch_cell_edges(ctx(), p, A);
  if (reuse) {
    carry_valid = false;
    CUDA_TRY(this, cudaMemsetAsync(A.force_old, 0, ns * sizeof(real2), stream));
    CUDA_TRY(this, cudaMemsetAsync(A.pos[1], 0, ns * sizeof(real2), stream));
    return MAVI_OK;
  }
  for (int b = 0; b < 2; b++)
    if ((st = dev_alloc(this, &A.pos[b], ns))) return st;
  if (second_kind == SECOND_VEL) {
    if ((st = dev_alloc(this, &A.vel, ns))) return st;
  } else if (second_kind == SECOND_ANGLE) {
    if ((st = dev_alloc(this, &A.ang, ns))) return st;
  }
  if ((st = dev_alloc(this, &A.idflag, ns))) return st;
  if ((st = dev_alloc(this, &A.cell, ns))) return st;
  if ((st = dev_alloc(this, &A.force, ns))) return st;
  if ((st = dev_alloc(this, &A.force_old, ns))) return st;
  const size_t nt = (size_t)p.nt;
  if ((st = dev_alloc(this, &A.tstart, nt * (MAVI_TR + 1) + 1))) return st;
  if ((st = dev_alloc(this, &A.tile_prefix, nt + 2))) return st;
  if ((st = dev_alloc(this, &A.perm, ns + nt + 2))) return st;
  if ((st = dev_alloc(this, &A.scan_partials, (nt + 2) / 4096 + 2))) return st;
  if ((st = dev_alloc(this, &A.tile_dirty, nt + 1))) return st;
  if ((st = dev_alloc(this, &A.dirty_list, nt + 1))) return st;
  if ((st = dev_alloc(this, &A.inbox_cnt, nt + 1))) return st;
  if ((st = dev_alloc(this, &A.inbox, nt * (size_t)p.inbox_cap + 1))) return st;
  const size_t mv = (size_t)p.mv_cap + 1;
  if ((st = dev_alloc(this, &A.mv_src, mv))) return st;
  if ((st = dev_alloc(this, &A.mv_pos, mv))) return st;
  if ((st = dev_alloc(this, &A.mv_second, mv))) return st;
  if ((st = dev_alloc(this, &A.mv_force, mv))) return st;
  if ((st = dev_alloc(this, &A.mv_id, mv))) return st;
  if ((st = dev_alloc(this, &A.mv_cell, mv))) return st;
  if ((st = dev_alloc(this, &A.chg, (size_t)p.chg_cap + 2))) return st;
  carry_valid = false;
  CUDA_TRY(this, cudaMemsetAsync(A.force_old, 0, ns * sizeof(real2), stream));
  CUDA_TRY(this, cudaMemsetAsync(A.pos[1], 0, ns * sizeof(real2), stream));
  return MAVI_OK;
}

static int round_up16(double x) { return ((int)std::ceil(x) + 15) / 16 * 16; }

// update_chunks! from scratch: staging arrays -> tile layout.  Grows the tile capacity until every tile fits.
int Handle::rebuild_from_staging(int n_active) {
  const bool second_is_vel = second_kind == SECOND_VEL;
  carry_valid = false;  // new layout / new state: the next Newton step starts with the full first pass
  int cap = p.cap;
  if (p.num_cells > 0 && (a.pos[0] == nullptr || n_active != p.n_active || cap <= 0)) {
    const long long ntiles = (long long)(p.slab ? p.num_cols - 2 : p.num_cols) * ((p.num_rows + MAVI_TR - 1) / MAVI_TR);
    if (cap <= 0 || !p.slab) cap = (flags_cfg & MAVI_FLAG_TIGHT_TILES) ? 1 : round_up16(2.0 * (double)n_active / (double)ntiles + 16.0);
    if (p.slab) {  // all ranks must agree on the tile capacity
      int st = slab_allreduce_max(this, &cap);
      if (st) return st;
    }
    int st = alloc_state(n_active, cap);
    if (st) return st;
  } else if (a.pos[0] == nullptr || n_active != p.n_count) {
    int st = alloc_state(n_active, 0);
    if (st) return st;
  } else if (p.slab) {
    int st = slab_allreduce_max(this, &cap);  // keep the collective sequence identical on every rank
    if (st) return st;
    if (cap != p.cap && (st = alloc_state(n_active, cap))) return st;
  }
  for (int attempt = 0; attempt < 8; attempt++) {
    if (p.num_cells == 0) {
      // chunks === nothing: slots are the original order; plain copies
      const size_t n = (size_t)p.n;
      CUDA_TRY(this, cudaMemcpyAsync(a.pos[0], a.st_pos, n * sizeof(real2), cudaMemcpyDeviceToDevice, stream));
      if (second_is_vel) CUDA_TRY(this, cudaMemcpyAsync(a.vel, a.st_vel, n * sizeof(real2), cudaMemcpyDeviceToDevice, stream));
      else if (second_kind == SECOND_ANGLE) CUDA_TRY(this, cudaMemcpyAsync(a.ang, a.st_ang, n * sizeof(real), cudaMemcpyDeviceToDevice, stream));
      CUDA_TRY(this, cudaMemcpyAsync(a.force, a.st_force, n * sizeof(real2), cudaMemcpyDeviceToDevice, stream));
      CUDA_TRY(this, cudaMemcpyAsync(a.idflag, a.st_id, n * sizeof(unsigned int), cudaMemcpyDeviceToDevice, stream));
      return MAVI_OK;
    }
    CUDA_TRY(this, cudaMemsetAsync(a.count, 0, ((size_t)p.num_cells + 2) * sizeof(int), stream));
    CUDA_TRY(this, cudaMemsetAsync(a.flags + 1, 0, (FLAG_STEPS - 1) * sizeof(int), stream));
    CUDA_TRY(this, cudaMemsetAsync(a.flags + FLAG_NCHG, 0, 2 * sizeof(int), stream));
    CUDA_TRY(this, cudaMemsetAsync(a.tile_dirty, 0, ((size_t)p.nt + 1) * sizeof(int), stream));
    CUDA_TRY(this, cudaMemsetAsync(a.inbox_cnt, 0, ((size_t)p.nt + 1) * sizeof(int), stream));
    launch_build_tiles(ctx(), p, a, second_is_vel);
    int st = check_device_flags();
    if (st) return st;
    int need = flags_host[FLAG_OVERFLOW] ? ((flags_cfg & MAVI_FLAG_TIGHT_TILES) ? flags_host[FLAG_MAXCOUNT] : round_up16(flags_host[FLAG_MAXCOUNT] * 1.25 + 8.0)) : 0;
    if (p.slab && (st = slab_allreduce_max(this, &need))) return st;  // if ANY rank overflowed, everybody grows
    if (need == 0) return p.slab ? slab_after_build(this) : MAVI_OK;
    cap = need > p.cap ? need : ((flags_cfg & MAVI_FLAG_TIGHT_TILES) ? p.cap + 1 : round_up16(p.cap * 1.25 + 8.0));
    if ((st = alloc_state(n_active, cap))) return st;
  }
  set_error("tile capacity did not converge");
  return MAVI_ERR_CAPACITY;
}

// update_chunks! of the current device state (mavi_bin, overflow fallback)
int Handle::rebuild_from_current() {
  launch_compact_to_staging(ctx(), p, a, second_kind == SECOND_VEL);
  return rebuild_from_staging(p.num_cells > 0 ? p.n_active : p.n_count);
}

// A particle left the chunk grid during the previous step: the reference notices at its next update_chunks!
// (BoundsError, src/chunks.jl:144-146), i.e. when the next step / calc_forces! / update_chunks! is requested.
int Handle::pending_out_of_grid() {
  if (flags_host && (flags_host[FLAG_ERR] & ERRBIT_OOG_PENDING)) {
    set_error("a particle left the chunk grid (BoundsError in the reference, src/chunks.jl:144-146)");
    return MAVI_ERR_OUT_OF_GRID;
  }
  return MAVI_OK;
}

// Reads the device error word; called at every synchronisation point.
int Handle::check_device_flags() {
  int *f = flags_host;
  cudaError_t e = cudaMemcpyAsync(f, a.flags, FLAG_COUNT * sizeof(int), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) {
    set_error("CUDA error %s while reading the device error word", cudaGetErrorString(e));
    return MAVI_ERR_CUDA;
  }
  if (f[0] & ERRBIT_OUTSIDE_SPACE) {
    set_error("Particles outside space (check_inside, src/space_checks.jl)");
    return MAVI_ERR_OUTSIDE_SPACE;
  }
  if (f[0] & ERRBIT_OUT_OF_GRID) {
    set_error("a particle left the chunk grid (BoundsError in the reference, src/chunks.jl:144-146)");
    return MAVI_ERR_OUT_OF_GRID;
  }
  if (f[0] & ERRBIT_NAN) {
    set_error("non-finite state");
    return MAVI_ERR_NAN;
  }
  return MAVI_OK;
}

// Enqueue one step (no host synchronisation).  newton_step! / szabo_step! / rtp_step!, src/integration.jl:507-535.
int Handle::ensure_id64(size_t n) {
  if (n <= id64_cap) return MAVI_OK;
  if (id64_dev) cudaFree(id64_dev);
  id64_dev = nullptr;
  id64_cap = 0;
  const size_t cap = n > (size_t)n_cap ? n : (size_t)n_cap;
  CUDA_TRY(this, cudaMalloc((void **)&id64_dev, cap * sizeof(long long)));
  id64_cap = cap;
  return MAVI_OK;
}

int Handle::enqueue_step(const real *noise_dev) {
  int st;
  LaunchCtx c = ctx();
  const bool second_is_vel = second_kind == SECOND_VEL;
  if (prof) cudaEventRecord(ev[0], stream);
  // per-step control words: #dirty tiles, big-drift guard, #position fix-ups, #inter-tile movers, "step ran"
  launch_step_begin(c, a);
  if (prof) cudaEventRecord(ev[1], stream);
  // Force carry (chunked Newton runs; MAVI_FLAG_NO_FORCE_CARRY switches it off): the second pass of the previous step
  // already produced this step's F1 and drift, see k_newton_b.
  const bool carry = second_is_vel && p.num_cells > 0 && !(flags_cfg & (MAVI_FLAG_NO_FORCE_CARRY | MAVI_FLAG_RESORT_EVERY_STEP));
  if (second_is_vel) {
    if (!carry || !carry_valid) launch_newton_a(c, p, a);  // pos[0] -> pos[1] (drift), F1 -> force_old
    if (prof) cudaEventRecord(ev[2], stream);
    launch_newton_b(c, p, a, carry);  // F2 from pos[1]; vel, force; sparse wall fix-ups applied to pos[1]
  } else {
    if (prof) cudaEventRecord(ev[2], stream);
    launch_self_propelled(c, p, a, noise_dev, (unsigned long long)num_steps);
  }
  std::swap(a.pos[0], a.pos[1]);
  if (prof) cudaEventRecord(ev[3], stream);
  // update_chunks! for the NEXT step, incrementally: only tiles a particle left or entered are rewritten
  if (!(flags_cfg & MAVI_FLAG_RESORT_EVERY_STEP)) launch_repair_tiles(c, p, a, second_is_vel);
  else if (p.num_cells > 0 && (st = rebuild_from_current())) return st;  // A/B switch: global rebuild instead of the repair
  if (carry) {
    launch_carry_fixups(c, p, a);
    carry_valid = true;
  }
  if (prof) cudaEventRecord(ev[4], stream);
  time += dt_host;  // update_time!, src/integration.jl:500-503
  num_steps += 1;
  return MAVI_OK;
}

// nsteps steps with the control words read back only every SYNC_EVERY steps.  A step whose repair overflowed (or that
// pushed a particle out of the grid) latches a flag that turns every later kernel into a no-op; the device-side step
// counter tells how many steps really ran, the host rolls its clock and the ping-pong parity back, grows the tiles and
// resumes.
int Handle::run_steps(long long nsteps, const real *noise_dev, size_t stride) {
  constexpr int SYNC_EVERY = 32;
  long long done_total = 0;
  int st;
  while (done_total < nsteps) {
    if ((st = pending_out_of_grid())) return st;
    if (p.dynamics == MAVI_DYN_RINGS) {
      // Rings: up to SYNC_EVERY steps enqueued back to back; an index-tile overflow latches, the device step counter
      // tells how many steps ran, the host clock is rolled back and the rest is re-run with larger tiles
      const int batch = (int)((nsteps - done_total) < SYNC_EVERY ? (nsteps - done_total) : SYNC_EVERY);
      long long snap_steps[SYNC_EVERY + 1], snap_check[SYNC_EVERY + 1];
      double snap_time[SYNC_EVERY + 1];
      snap_steps[0] = num_steps;
      snap_time[0] = time;
      snap_check[0] = r.inv_last_check;
      const int c0 = steps_seen;
      for (int s = 0; s < batch; s++) {
        if ((st = rings_step(this, noise_dev ? noise_dev + (size_t)(done_total + s) * stride : nullptr))) return st;
        snap_steps[s + 1] = num_steps;
        snap_time[s + 1] = time;
        snap_check[s + 1] = r.inv_last_check;
      }
      st = check_device_flags();
      int done = flags_host[FLAG_STEPS] - c0;
      if (done < 0 || done > batch) done = batch;
      steps_seen = flags_host[FLAG_STEPS];
      num_steps = snap_steps[done];
      time = snap_time[done];
      r.inv_last_check = snap_check[done];
      done_total += done;
      if (st) return st;
      if (flags_host[FLAG_OVERFLOW] && (st = rings_grow_tiles(this))) return st;
      continue;
    }
    if (p.slab) {  // this path synchronises by itself (slab_sync_counts)
      const real *nz = noise_dev ? noise_dev + (size_t)done_total * stride : nullptr;
      st = slab_step_once(this, nz);
      if (st) return st;
      done_total += 1;
      // slab steps are enqueued without host synchronisation; counts / overflow word every SYNC_EVERY steps and at the end
      if (p.slab && (done_total == nsteps || done_total % SYNC_EVERY == 0) && (st = slab_sync_counts(this))) return st;
      continue;
    }
    const int batch = (int)((nsteps - done_total) < SYNC_EVERY ? (nsteps - done_total) : SYNC_EVERY);
    long long snap_steps[SYNC_EVERY + 1];
    double snap_time[SYNC_EVERY + 1];
    snap_steps[0] = num_steps;
    snap_time[0] = time;
    const int c0 = steps_seen;
    for (int s = 0; s < batch; s++) {
      if ((st = enqueue_step(noise_dev ? noise_dev + (size_t)(done_total + s) * stride : nullptr))) return st;
      snap_steps[s + 1] = num_steps;
      snap_time[s + 1] = time;
    }
    st = check_device_flags();
    int done = flags_host[FLAG_STEPS] - c0;
    if (flags_cfg & MAVI_FLAG_RESORT_EVERY_STEP) done = batch;
    if (done < 0 || done > batch) done = batch;
    steps_seen = flags_host[FLAG_STEPS];
    if (done < batch) {  // some enqueued steps were no-ops: undo their host-side bookkeeping
      num_steps = snap_steps[done];
      time = snap_time[done];
      if ((batch - done) & 1) std::swap(a.pos[0], a.pos[1]);
    }
    done_total += done;
    if (st) return st;
    if (flags_host[FLAG_OVERFLOW]) {
      // a tile ran out of slots (or the mover list overflowed): nothing was modified by the repair of that step;
      // rebuild, with a larger capacity when a tile was really full
      const bool second_is_vel = second_kind == SECOND_VEL;
      int cap = p.cap;
      if (flags_host[FLAG_OVERFLOW] & 1) {
        const int mx = flags_host[FLAG_MAXCOUNT] > p.cap ? flags_host[FLAG_MAXCOUNT] : p.cap + 1;
        cap = (flags_cfg & MAVI_FLAG_TIGHT_TILES) ? mx : round_up16(mx * 1.25 + 8.0);
      }
      n_rebuilds++;
      launch_compact_to_staging(ctx(), p, a, second_is_vel);
      const int n_active = p.n_active;
      if (cap != p.cap && (st = alloc_state(n_active, cap))) return st;
      if ((st = rebuild_from_staging(n_active))) return st;
    }
  }
  return MAVI_OK;
}

}  // namespace MAVI_NS

// =========================================================================================================
// C ABI
// =========================================================================================================
// The entry points of include/mavi.h for THIS arithmetic type (api_<name> == mavi_<name>); capi.cu holds the extern "C"
// symbols and dispatches on MaviParams.dtype.
namespace MAVI_NS {


int32_t api_abi_version(void) { return MAVI_ABI_VERSION; }

int32_t api_create(const MaviParams *params, void **out) {
  if (!params || !out) return MAVI_ERR_BAD_PARAMS;
  Handle *h = new Handle();
  *out = h;
  h->device = params->device;
  h->flags_cfg = params->flags;
  cudaError_t e = cudaSetDevice(h->device);
  if (e != cudaSuccess) {
    h->set_error("cudaSetDevice(%d): %s — libmavi_cuda.so needs a CUDA device; there is no CPU fallback", h->device,
                 cudaGetErrorString(e));
    return MAVI_ERR_CUDA;
  }
  h->stream = (cudaStream_t)params->stream;
  switch (params->dynamics) {
    case MAVI_DYN_LJ:
    case MAVI_DYN_HARMTRUNC: h->second_kind = SECOND_VEL; break;
    case MAVI_DYN_SZABO:
    case MAVI_DYN_RTP: h->second_kind = SECOND_ANGLE; break;
    default: h->second_kind = SECOND_RING_POL;
  }
  int st = validate_and_lower(h, params);
  if (st) return st;
  if ((st = allocate(h))) return st;
  for (int i = 0; i < 5; i++) cudaEventCreate(&h->ev[i]);
  cudaEventCreate(&h->ev_call[0]);
  cudaEventCreate(&h->ev_call[1]);
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return MAVI_OK;
}

int32_t api_destroy(void *hh) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return MAVI_OK;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  slab_destroy(h);
  for (void *ptr : h->allocs)
    if (ptr) cudaFree(ptr);
  if (h->flags_host) cudaFreeHost(h->flags_host);
  if (h->noise_dev) cudaFree(h->noise_dev);
  if (h->id64_dev) cudaFree(h->id64_dev);
  for (int i = 0; i < 5; i++)
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  for (int i = 0; i < 2; i++)
    if (h->ev_call[i]) cudaEventDestroy(h->ev_call[i]);
  delete h;
  return MAVI_OK;
}

int32_t api_last_error(void *hh, char *buf, int32_t n) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || !buf || n <= 0) return MAVI_ERR_BAD_PARAMS;
  snprintf(buf, (size_t)n, "%s", h->err);
  return MAVI_OK;
}

// System ctor tail (src/systems.jl:73-114): ids, inside check, first update_chunks!.
int32_t api_upload_state(void *hh, const void *pos, const void *second, const uint8_t *active_mask, int64_t n) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || !pos || n != h->p.n) return MAVI_ERR_BAD_PARAMS;
  if (h->p.slab) {
    h->set_error("slab mode: use mavi_upload_local (every rank uploads the particles of its own cell columns)");
    return MAVI_ERR_BAD_PARAMS;
  }
  cudaSetDevice(h->device);
  DevArrays &a = h->a;
  LaunchCtx c = h->ctx();
  const size_t sn = (size_t)n;
  CUDA_TRY(h, cudaMemcpyAsync(a.st_pos, pos, sn * sizeof(real2), cudaMemcpyHostToDevice, h->stream));
  if (second) {
    if (h->second_kind == SECOND_VEL)
      CUDA_TRY(h, cudaMemcpyAsync(a.st_vel, second, sn * sizeof(real2), cudaMemcpyHostToDevice, h->stream));
    else if (h->second_kind == SECOND_ANGLE)
      CUDA_TRY(h, cudaMemcpyAsync(a.st_ang, second, sn * sizeof(real), cudaMemcpyHostToDevice, h->stream));
    else
      CUDA_TRY(h, cudaMemcpyAsync(a.st_ang, second, (size_t)h->p.rings.num_rings * sizeof(real), cudaMemcpyHostToDevice, h->stream));
  }
  CUDA_TRY(h, cudaMemsetAsync(a.flags, 0, FLAG_COUNT * sizeof(int), h->stream));
  h->steps_seen = 0;
  if (h->second_kind == SECOND_RING_POL) return rings_upload_finish(h);
  unsigned char *mask_dev = nullptr;
  int n_active = h->p.n;
  if (active_mask) {
    // ParticleIds (src/states.jl:27-52): count = number of active ids
    mask_dev = reinterpret_cast<unsigned char *>(a.st_cell);
    CUDA_TRY(h, cudaMemcpyAsync(mask_dev, active_mask, sn, cudaMemcpyHostToDevice, h->stream));
    n_active = 0;
    for (size_t i = 0; i < sn; i++) n_active += active_mask[i] != 0;
  }
  launch_init_staging_ids(c, h->p.n, mask_dev, a.st_id);
  CUDA_TRY(h, cudaMemsetAsync(a.st_force, 0, sn * sizeof(real2), h->stream));
  if (h->p.n_spaces == 1) launch_check_inside(c, h->p, a);
  int st = h->check_device_flags();
  if (st) return st;
  if ((st = h->rebuild_from_staging(n_active))) return st;
  return h->check_device_flags();
}

int32_t api_download_state(void *hh, void *pos, void *second) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return MAVI_ERR_BAD_PARAMS;
  cudaSetDevice(h->device);
  DevArrays &a = h->a;
  LaunchCtx c = h->ctx();
  const size_t sn = (size_t)h->p.n;
  if (h->second_kind == SECOND_RING_POL) return rings_download_state(h, pos, second);
  if (h->p.slab) {
    h->set_error("slab mode: use mavi_download_local");
    return MAVI_ERR_BAD_PARAMS;
  }
  if (pos) {
    launch_unpermute2(c, h->p, a, a.pos[0], a.st_pos);
    CUDA_TRY(h, cudaMemcpyAsync(pos, a.st_pos, sn * sizeof(real2), cudaMemcpyDeviceToHost, h->stream));
  }
  if (second) {
    if (h->second_kind == SECOND_VEL) {
      launch_unpermute2(c, h->p, a, a.vel, a.st_vel);
      CUDA_TRY(h, cudaMemcpyAsync(second, a.st_vel, sn * sizeof(real2), cudaMemcpyDeviceToHost, h->stream));
    } else {
      launch_unpermute1(c, h->p, a, a.ang, a.st_ang);
      CUDA_TRY(h, cudaMemcpyAsync(second, a.st_ang, sn * sizeof(real), cudaMemcpyDeviceToHost, h->stream));
    }
  }
  return h->check_device_flags();
}

int32_t api_download_forces(void *hh, void *forces) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || !forces) return MAVI_ERR_BAD_PARAMS;
  cudaSetDevice(h->device);
  DevArrays &a = h->a;
  if (h->second_kind == SECOND_RING_POL) return rings_download_forces(h, forces);
  if (h->p.slab) {
    h->set_error("slab mode: use mavi_download_local (ids are global; this entry point un-permutes into a dense local array)");
    return MAVI_ERR_BAD_PARAMS;
  }
  launch_unpermute2(h->ctx(), h->p, a, a.force, a.st_force);
  CUDA_TRY(h, cudaMemcpyAsync(forces, a.st_force, (size_t)h->p.n * sizeof(real2), cudaMemcpyDeviceToHost, h->stream));
  return h->check_device_flags();
}

int32_t api_local_count(void *hh, int64_t *n_local) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || !n_local) return MAVI_ERR_BAD_PARAMS;
  *n_local = h->p.n;
  return MAVI_OK;
}

// ids / state / forces of the particles this rank currently owns (single GPU: everything, in original-id order)
int32_t api_download_local(void *hh, int64_t *ids, void *pos, void *second, void *forces) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return MAVI_ERR_BAD_PARAMS;
  if (!h->p.slab) {
    if (ids)
      for (int64_t i = 0; i < h->p.n; i++) ids[i] = i;
    int st = api_download_state(hh, pos, second);
    if (st) return st;
    if (forces) return api_download_forces(hh, forces);
    return MAVI_OK;
  }
  cudaSetDevice(h->device);
  DevArrays &a = h->a;
  const size_t n = (size_t)h->p.n;
  if ((int)n > h->n_cap) {
    h->set_error("owned particle count %zu exceeds the staging capacity %d", n, h->n_cap);
    return MAVI_ERR_CAPACITY;
  }
  const bool vel = h->second_kind == SECOND_VEL;
  launch_compact_to_staging(h->ctx(), h->p, a, vel);
  if (pos) CUDA_TRY(h, cudaMemcpyAsync(pos, a.st_pos, n * sizeof(real2), cudaMemcpyDeviceToHost, h->stream));
  if (second) {
    if (vel) CUDA_TRY(h, cudaMemcpyAsync(second, a.st_vel, n * sizeof(real2), cudaMemcpyDeviceToHost, h->stream));
    else CUDA_TRY(h, cudaMemcpyAsync(second, a.st_ang, n * sizeof(real), cudaMemcpyDeviceToHost, h->stream));
  }
  if (forces) CUDA_TRY(h, cudaMemcpyAsync(forces, a.st_force, n * sizeof(real2), cudaMemcpyDeviceToHost, h->stream));
  if (ids && n > 0) {  // u32 -> int64 on the device, one copy straight into the caller's buffer
    int st = h->ensure_id64(n);
    if (st) return st;
    launch_ids_to_i64(h->ctx(), (int)n, a.st_id, h->id64_dev);
    CUDA_TRY(h, cudaMemcpyAsync(ids, h->id64_dev, n * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  }
  return h->check_device_flags();
}

// internal (multi.cu): GLOBAL cell ids of the owned particles, in the order mavi_download_local lists them
int32_t api_download_local_cells(void *hh, int32_t *cells) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || !cells || h->p.num_cells == 0) return MAVI_ERR_BAD_PARAMS;
  cudaSetDevice(h->device);
  if (h->p.slab) slab_join(h);
  const size_t n = (size_t)h->p.n;
  if ((int)n > h->n_cap) return MAVI_ERR_CAPACITY;
  launch_compact_cells(h->ctx(), h->p, h->a, h->a.st_cell);
  CUDA_TRY(h, cudaMemcpyAsync(cells, h->a.st_cell, n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  return h->check_device_flags();
}

// slab mode upload: the particles whose cell column this rank owns, with their global original ids
int32_t api_upload_local(void *hh, const int64_t *ids, const void *pos, const void *second, int64_t n_local) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || !pos || !ids || n_local < 0) return MAVI_ERR_BAD_PARAMS;
  if (!h->p.slab) {
    h->set_error("mavi_upload_local is for slab mode (MaviParams.world > 1)");
    return MAVI_ERR_BAD_PARAMS;
  }
  if (n_local > h->n_cap) {
    h->set_error("n_local %lld exceeds the capacity %d derived from MaviParams.n", (long long)n_local, h->n_cap);
    return MAVI_ERR_CAPACITY;
  }
  cudaSetDevice(h->device);
  DevArrays &a = h->a;
  const size_t sn = (size_t)n_local;
  long long id_min = 0, id_max = 0;
  for (size_t i = 0; i < sn; i++) {
    id_min = ids[i] < id_min ? ids[i] : id_min;
    id_max = ids[i] > id_max ? ids[i] : id_max;
  }
  if (id_min < 0 || id_max >= 0x7fffffffLL) {
    h->set_error("particle ids must fit in 31 bits");
    return MAVI_ERR_BAD_PARAMS;
  }
  if (sn > 0) {  // int64 -> u32 on the device
    int st = h->ensure_id64(sn);
    if (st) return st;
    CUDA_TRY(h, cudaMemcpyAsync(h->id64_dev, ids, sn * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    launch_ids_from_i64(h->ctx(), (int)sn, h->id64_dev, a.st_id);
  }
  CUDA_TRY(h, cudaMemcpyAsync(a.st_pos, pos, sn * sizeof(real2), cudaMemcpyHostToDevice, h->stream));
  if (second) {
    if (h->second_kind == SECOND_VEL)
      CUDA_TRY(h, cudaMemcpyAsync(a.st_vel, second, sn * sizeof(real2), cudaMemcpyHostToDevice, h->stream));
    else
      CUDA_TRY(h, cudaMemcpyAsync(a.st_ang, second, sn * sizeof(real), cudaMemcpyHostToDevice, h->stream));
  }
  CUDA_TRY(h, cudaMemsetAsync(a.st_force, 0, sn * sizeof(real2), h->stream));
  CUDA_TRY(h, cudaMemsetAsync(a.flags, 0, FLAG_COUNT * sizeof(int), h->stream));
  h->steps_seen = 0;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  h->p.n = (int)n_local;
  int st = h->rebuild_from_staging((int)n_local);
  if (st) return st;
  if ((st = slab_sync_counts(h))) return st;  // tiles that start out nearly full grow before the first step (collective)
  return h->check_device_flags();
}

int32_t api_nccl_unique_id(void *out128) {
  if (!out128) return MAVI_ERR_BAD_PARAMS;
  return slab_unique_id(out128);
}

int32_t api_step(void *hh, int64_t nsteps, const void *host_noise) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || nsteps < 0) return MAVI_ERR_BAD_PARAMS;
  cudaSetDevice(h->device);
  const real *noise_dev = nullptr;
  size_t stride = 0;
  if (host_noise && h->p.rng_mode == MAVI_RNG_HOST_NOISE) {
    // the kernels index the noise row with the ORIGINAL id: in slab mode that is the global id, so every rank passes the
    // full global row (n_global entries per step), not one entry per owned particle
    const size_t n_ids = h->p.slab ? (size_t)h->slab.n_global : (size_t)h->p.n;
    switch (h->p.dynamics) {
      case MAVI_DYN_SZABO: stride = n_ids; break;
      case MAVI_DYN_RTP: stride = 2 * n_ids; break;
      case MAVI_DYN_RINGS: stride = (size_t)h->p.rings.num_rings; break;
      default: stride = 0;
    }
    size_t total = stride * (size_t)nsteps;
    if (total > 0) {
      if (total > h->noise_cap) {
        if (h->noise_dev) cudaFree(h->noise_dev);
        h->noise_dev = nullptr;
        h->noise_cap = 0;
        CUDA_TRY(h, cudaMalloc((void **)&h->noise_dev, total * sizeof(real)));
        h->noise_cap = total;
      }
      CUDA_TRY(h, cudaMemcpyAsync(h->noise_dev, host_noise, total * sizeof(real), cudaMemcpyHostToDevice, h->stream));
      noise_dev = h->noise_dev;
    }
  }
  if (h->prof) cudaEventRecord(h->ev_call[0], h->stream);
  int st = h->run_steps(nsteps, noise_dev, stride);
  if (h->prof) cudaEventRecord(h->ev_call[1], h->stream);
  if (st) return st;
  CUDA_TRY(h, cudaGetLastError());
  return h->check_device_flags();
}

int32_t api_calc_forces(void *hh) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return MAVI_ERR_BAD_PARAMS;
  cudaSetDevice(h->device);
  if (h->p.dynamics == MAVI_DYN_RINGS) return rings_calc_forces(h);
  if (int st0 = h->pending_out_of_grid()) return st0;
  // the tile layout always equals the fresh binning of the current positions (repaired at the end of every step)
  launch_force_only(h->ctx(), h->p, h->a, true);
  return h->check_device_flags();
}

int32_t api_bin(void *hh) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return MAVI_ERR_BAD_PARAMS;
  cudaSetDevice(h->device);
  if (h->p.dynamics == MAVI_DYN_RINGS) return rings_bin(h);
  if (h->p.num_cells == 0) return MAVI_OK;  // update_chunks!(::Nothing)
  if (int st0 = h->pending_out_of_grid()) return st0;
  int st = h->rebuild_from_current();
  if (st) return st;
  return h->check_device_flags();
}

int32_t api_download_cells(void *hh, int32_t *cell_of_particle, int32_t *counts) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || h->p.num_cells == 0) return MAVI_ERR_BAD_PARAMS;
  cudaSetDevice(h->device);
  DevArrays &a = h->a;
  if (h->p.dynamics == MAVI_DYN_RINGS) return rings_download_cells(h, cell_of_particle, counts, nullptr, nullptr);
  if (h->p.slab) {
    h->set_error("slab mode: use mavi_download_local (ids are global; this entry point un-permutes into a dense local array)");
    return MAVI_ERR_BAD_PARAMS;
  }
  if (cell_of_particle) {
    launch_unpermute_cells(h->ctx(), h->p, a, a.st_cell);
    CUDA_TRY(h, cudaMemcpyAsync(cell_of_particle, a.st_cell, (size_t)h->p.n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  }
  if (counts) {
    launch_cell_counts(h->ctx(), h->p, a, a.perm);
    CUDA_TRY(h, cudaMemcpyAsync(counts, a.perm, (size_t)h->p.num_cells * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  }
  return h->check_device_flags();
}

int32_t api_download_cell_lists(void *hh, int32_t *start, int32_t *ids) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || h->p.num_cells == 0) return MAVI_ERR_BAD_PARAMS;
  cudaSetDevice(h->device);
  DevArrays &a = h->a;
  if (h->p.dynamics == MAVI_DYN_RINGS) return rings_download_cells(h, nullptr, nullptr, start, ids);
  if (h->p.slab) {
    h->set_error("slab mode: use mavi_download_local (ids are global; this entry point un-permutes into a dense local array)");
    return MAVI_ERR_BAD_PARAMS;
  }
  if (start) {
    std::vector<int> counts((size_t)h->p.num_cells);
    launch_cell_counts(h->ctx(), h->p, a, a.perm);
    CUDA_TRY(h, cudaMemcpyAsync(counts.data(), a.perm, counts.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    int acc = 0;
    for (int cix = 0; cix < h->p.num_cells; cix++) {
      start[cix] = acc;
      acc += counts[cix];
    }
    start[h->p.num_cells] = acc;
  }
  if (ids) {
    launch_ids_in_cell_order(h->ctx(), h->p, a, a.st_cell);
    CUDA_TRY(h, cudaMemcpyAsync(ids, a.st_cell, (size_t)h->p.n_active * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  }
  return h->check_device_flags();
}

// The neighbour cells the stencil walker visits for `cell`, in visiting order (no de-duplication: 2-wide periodic
// grids list a cell twice exactly where the reference double counts).  Mirrors for_each_neighbor (common.cuh).
int32_t api_cell_neighbors(void *hh, int32_t cell, int32_t *out8, int32_t *n) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || !out8 || !n || h->p.num_cells == 0 || cell < 0 || cell >= h->p.num_cells) return MAVI_ERR_BAD_PARAMS;
  const DevParams &p = h->p;
  const int R = p.num_rows, Cn = p.num_cols;
  const int col = cell / R, row = cell - col * R;
  int cnt = 0;
  for (int dc = -1; dc <= 1; dc++) {
    int c2 = col + dc;
    if (c2 < 0) { if (!p.wrap_cols) continue; c2 = Cn - 1; }
    else if (c2 >= Cn) { if (!p.wrap_cols) continue; c2 = 0; }
    for (int dr = -1; dr <= 1; dr++) {
      int r2 = row + dr;
      if (r2 < 0) { if (!p.wrap_rows) continue; r2 = R - 1; }
      else if (r2 >= R) { if (!p.wrap_rows) continue; r2 = 0; }
      if (dc == 0 && dr == 0) continue;
      out8[cnt++] = c2 * R + r2;
    }
  }
  *n = cnt;
  return MAVI_OK;
}

int32_t api_energies(void *hh, int32_t pe_mode, double *ke, double *pe) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return MAVI_ERR_BAD_PARAMS;
  cudaSetDevice(h->device);
  DevArrays &a = h->a;
  LaunchCtx c = h->ctx();
  double *out = a.reduce_buf + 2048;
  double host[2] = {NAN, NAN};
  if (h->p.slab && pe && pe_mode == 0) {
    h->set_error("slab mode: the exact O(N^2) potential energy is single-GPU only (SURVEY.md 8e); use pe_mode 1");
    return MAVI_ERR_BAD_PARAMS;
  }
  if (ke && h->second_kind == SECOND_VEL) {
    launch_kinetic_energy(c, h->p, a, out);
    CUDA_TRY(h, cudaMemcpyAsync(&host[0], out, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  if (pe && h->p.dynamics == MAVI_DYN_LJ) {
    if (pe_mode == 1 && h->p.num_cells == 0) {
      h->set_error("pe_mode 1 needs chunks");
      return MAVI_ERR_BAD_PARAMS;
    }
    if (pe_mode == 0) launch_compact_to_staging(c, h->p, a, true);
    launch_potential_energy(c, h->p, a, pe_mode, out + 1);
    CUDA_TRY(h, cudaMemcpyAsync(&host[1], out + 1, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  int st = h->check_device_flags();
  if (ke) *ke = host[0];
  if (pe) *pe = host[1];
  return st;
}

int32_t api_rings_download_info(void *hh, void *areas, void *cms, void *cont_pos) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || h->p.dynamics != MAVI_DYN_RINGS) return MAVI_ERR_BAD_PARAMS;
  cudaSetDevice(h->device);
  return rings_download_info(h, areas, cms, cont_pos);
}

int32_t api_rings_set_neighbors(void *hh, int32_t mode, int32_t type_all, double tol) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return MAVI_ERR_BAD_PARAMS;
  cudaSetDevice(h->device);
  return rings_set_neighbors(h, mode, type_all, tol);
}

int32_t api_rings_download_neighbors(void *hh, int32_t *count, int32_t *list) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return MAVI_ERR_BAD_PARAMS;
  cudaSetDevice(h->device);
  return rings_download_neighbors(h, count, list);
}

int32_t api_rings_set_sources(void *hh, const MaviSourceSink *list, int32_t n, const uint8_t *ring_active, const double *spawn_draws,
                              int64_t n_draws) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return MAVI_ERR_BAD_PARAMS;
  cudaSetDevice(h->device);
  return rings_set_sources(h, list, n, ring_active, spawn_draws, n_draws);
}

int32_t api_rings_download_active(void *hh, uint8_t *ring_active, int64_t *uids, int64_t *num_active) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return MAVI_ERR_BAD_PARAMS;
  long long na = 0;
  int st = rings_download_active(h, ring_active, reinterpret_cast<long long *>(uids), &na);
  if (num_active) *num_active = na;
  return st;
}

int32_t api_rings_set_invasions(void *hh, int32_t steps_to_update, int32_t r_cols, int32_t r_rows) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return MAVI_ERR_BAD_PARAMS;
  cudaSetDevice(h->device);
  return rings_set_invasions(h, steps_to_update, r_cols, r_rows);
}

int32_t api_rings_download_invasions(void *hh, int64_t *n, int32_t *triples, int64_t cap) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return MAVI_ERR_BAD_PARAMS;
  cudaSetDevice(h->device);
  long long cnt = 0;
  int st = rings_download_invasions(h, &cnt, triples, cap);
  if (n) *n = cnt;
  return st;
}

int32_t api_get_time(void *hh, int64_t *num_steps, double *time) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return MAVI_ERR_BAD_PARAMS;
  if (num_steps) *num_steps = h->num_steps;
  if (time) *time = h->time;
  return MAVI_OK;
}

int32_t api_set_time(void *hh, int64_t num_steps, double time) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return MAVI_ERR_BAD_PARAMS;
  h->num_steps = num_steps;
  h->time = time;
  return MAVI_OK;
}

int32_t api_sync(void *hh) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return MAVI_ERR_BAD_PARAMS;
  cudaSetDevice(h->device);
  return h->check_device_flags();
}

int32_t api_launch_count(void *hh, int64_t *n) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || !n) return MAVI_ERR_BAD_PARAMS;
  *n = h->launches;
  return MAVI_OK;
}

int32_t api_rebuild_count(void *hh, int64_t *n) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || !n) return MAVI_ERR_BAD_PARAMS;
  *n = h->n_rebuilds;
  return MAVI_OK;
}

// instrumentation: out[0] = steps run since the last upload, [1] = particles re-binned (cell changes), [2] = inter-tile
// movers, [3] = tiles repaired, [4] = emigrants (slab mode), [5] = tile-overflow rebuilds, [6] = tile capacity, [7] = tiles
int32_t api_counters(void *hh, int64_t *out8) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || !out8) return MAVI_ERR_BAD_PARAMS;
  cudaSetDevice(h->device);
  if (h->p.slab) slab_join(h);
  int st = h->check_device_flags();
  const int *f = h->flags_host;
  // the totals are folded in by the NEXT step's first kernel: add the words of the last step that ran
  out8[0] = f[FLAG_STEPS];
  out8[1] = (int64_t)f[FLAG_CUM_CHG] + f[FLAG_NMOVED];
  out8[2] = (int64_t)f[FLAG_CUM_MV] + f[FLAG_NMV];
  out8[3] = (int64_t)f[FLAG_CUM_DIRTY] + f[FLAG_CHANGED];
  out8[4] = (int64_t)f[FLAG_CUM_EM] + f[FLAG_NEM0] + f[FLAG_NEM1];
  out8[5] = h->n_rebuilds;
  out8[6] = h->p.cap;
  out8[7] = h->p.nt;
  return st;
}

int32_t api_set_profiling(void *hh, int32_t on) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h) return MAVI_ERR_BAD_PARAMS;
  h->prof = on != 0;
  return MAVI_OK;
}

int32_t api_last_step_ms(void *hh, float *ms5) {
  Handle *h = reinterpret_cast<Handle *>(hh);
  if (!h || !ms5) return MAVI_ERR_BAD_PARAMS;
  cudaSetDevice(h->device);
  for (int i = 0; i < 5; i++) ms5[i] = 0.f;
  if (!h->prof) return MAVI_OK;
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  cudaEventElapsedTime(&ms5[0], h->ev[0], h->ev[1]);
  cudaEventElapsedTime(&ms5[1], h->ev[1], h->ev[2]);
  cudaEventElapsedTime(&ms5[2], h->ev[2], h->ev[3]);
  cudaEventElapsedTime(&ms5[3], h->ev[3], h->ev[4]);
  cudaEventElapsedTime(&ms5[4], h->ev_call[0], h->ev_call[1]);
  cudaGetLastError();
  return MAVI_OK;
}

}  // namespace MAVI_NS
