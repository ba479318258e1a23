// kernels.cu — padded-tile cell lists (full build + O(#movers) incremental repair), fused force+integrate passes and
// energy reductions.  sm_100a.  No tensor cores: the path is FP64 vector math over HBM-resident SoA arrays (real2
// loads), not a dense contraction.  Reference citations are relative to /root/reference/.
#include <cuda_pipeline.h>

#include <cstdlib>

#include "kernels.cuh"
#include "walls.cuh"

namespace MAVI_NS {

static inline int nblk(long long n, int tpb = TPB) { return (int)((n + tpb - 1) / tpb); }

#define MAVI_LAUNCH(ctx, kernel, grid, block, smem, ...)             \
  do {                                                               \
    kernel<<<(grid), (block), (smem), (ctx).stream>>>(__VA_ARGS__);  \
    (*(ctx).launches)++;                                             \
  } while (0)

// rank (dense particle index) -> slot.  Ranks < n_active enumerate the tile populations in TILE-ROW-MAJOR order
// (tile_prefix is the exclusive scan in that order); cta_first[b] is the order index of the tile holding rank b*RPB.
// Ranks >= n_active are the inactive tail.  Used by the download / reduction kernels only: the force kernels are
// tile-block kernels that need no rank map.
__device__ __forceinline__ int slot_of_rank(const DevParams &p, const int *__restrict__ tile_prefix,
                                            const int *__restrict__ cta_first, int rank) {
  if (rank >= p.n_active) return p.tail_base + (rank - p.n_active);
  int o = __ldg(cta_first + rank / RPB);
  while (rank >= __ldg(tile_prefix + o + 1)) ++o;
  return tile_of_order(p, o) * p.cap + (rank - __ldg(tile_prefix + o));
}

// =========================================================================================================
// Full build: update_chunks! (src/chunks.jl:150-163, src/integration.jl:54-59) from the dense staging arrays.
// =========================================================================================================

// check_inside, src/space_checks.jl:9-61 (Rectangle: any coordinate < bottom_left or > top_right; Circle: |pos|^2 > R^2,
// centre ignored (sic)); only single-geometry spaces are checked (ManyGeometries hits the generic no-op method).
__global__ void k_check_inside(const __grid_constant__ DevParams p, const real2 *__restrict__ pos,
                               const unsigned int *__restrict__ idflag, int *__restrict__ flags) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.n || (idflag[k] & MAVI_INACTIVE_BIT)) return;
  const DevSpace &sp = p.spaces[0];
  real2 r = pos[k];
  bool out = false;
  if (sp.geom == MAVI_GEOM_RECT) {
    real trx = sp.rect_bl[0] + sp.rect_sz[0], try_ = sp.rect_bl[1] + sp.rect_sz[1];
    out = (r.x < sp.rect_bl[0]) || (r.y < sp.rect_bl[1]) || (r.x > trx) || (r.y > try_);
  } else if (sp.geom == MAVI_GEOM_CIRCLE) {
    out = (r.x * r.x + r.y * r.y) > sp.cr * sp.cr;
  }
  if (out) atomicOr(&flags[FLAG_ERR], ERRBIT_OUTSIDE_SPACE);
}

void launch_check_inside(const LaunchCtx &c, const DevParams &p, const DevArrays &a) {
  MAVI_LAUNCH(c, k_check_inside, nblk(p.n), TPB, 0, p, a.st_pos, a.st_id, a.flags);
}

__global__ void k_init_staging_ids(int n, const unsigned char *__restrict__ mask, unsigned int *__restrict__ st_id) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) st_id[k] = (unsigned int)k | ((mask && mask[k] == 0) ? MAVI_INACTIVE_BIT : 0u);
}

void launch_init_staging_ids(const LaunchCtx &c, int n, const unsigned char *mask, unsigned int *st_id) {
  MAVI_LAUNCH(c, k_init_staging_ids, nblk(n), TPB, 0, n, mask, st_id);
}

// slab-mode state movement: original ids cross the ABI as int64 and live as u32 (+ inactive bit) on the device
__global__ void k_ids_to_i64(int n, const unsigned int *__restrict__ st_id, long long *__restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = (long long)(st_id[k] & ~MAVI_INACTIVE_BIT);
}
__global__ void k_ids_from_i64(int n, const long long *__restrict__ in, unsigned int *__restrict__ st_id) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) st_id[k] = (unsigned int)in[k];
}
void launch_ids_to_i64(const LaunchCtx &c, int n, const unsigned int *st_id, long long *out) {
  MAVI_LAUNCH(c, k_ids_to_i64, nblk(n), TPB, 0, n, st_id, out);
}
void launch_ids_from_i64(const LaunchCtx &c, int n, const long long *in, unsigned int *st_id) {
  MAVI_LAUNCH(c, k_ids_from_i64, nblk(n), TPB, 0, n, in, st_id);
}

// cell id of every staged particle (update_particle_chunk!, src/chunks.jl:120-147) + per-cell histogram
__global__ void k_build_cell_index(const __grid_constant__ DevParams p, const real2 *__restrict__ st_pos,
                                   const unsigned int *__restrict__ st_id, int *__restrict__ st_cell,
                                   int *__restrict__ count, int *__restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  int c = -1;  // inactive: never binned (active ids only)
  if (!(st_id[i] & MAVI_INACTIVE_BIT)) {
    real2 r = st_pos[i];
    c = cell_of_point(p, r.x, r.y);
    if (c >= 0 && p.slab) {  // a full build only ever sees particles of the owned columns
      const int lcol = div_rows(p, c);
      if (lcol < 1 || lcol > p.num_cols - 2) c = -1;
    }
    if (c < 0) {  // BoundsError in the reference (src/chunks.jl:144-146)
      atomicOr(&flags[FLAG_ERR], ERRBIT_OUT_OF_GRID);
      c = p.slab ? p.num_rows : 0;
    }
    atomicAdd(&count[c], 1);
  }
  st_cell[i] = c;
}

// per tile: exclusive scan of its cells' populations -> tstart; capacity check
__global__ void k_build_layout(const __grid_constant__ DevParams p, const int *__restrict__ count,
                               int *__restrict__ tstart, int *__restrict__ flags) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.nt) return;
  const int col = t / p.tpc, tr = t - col * p.tpc;
  const int base = t * p.cap;
  int run = base;
  int *ts = tstart + (size_t)t * (MAVI_TR + 1);
  for (int lr = 0; lr < MAVI_TR; lr++) {
    ts[lr] = run;
    const int row = tr * MAVI_TR + lr;
    if (row < p.num_rows) run += count[col * p.num_rows + row];
  }
  ts[MAVI_TR] = run;
  const int cnt = run - base;
  atomicMax(&flags[FLAG_MAXCOUNT], cnt);
  if (cnt > p.cap) flags[FLAG_OVERFLOW] = 1;
}

// scatter staged particle indices into their cell range (cursor = count, consumed down to zero)
__global__ void k_build_scatter(const __grid_constant__ DevParams p, const int *__restrict__ st_cell,
                                const int *__restrict__ tstart, int *__restrict__ count, int *__restrict__ perm,
                                const int *__restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n || flags[FLAG_OVERFLOW]) return;
  int c = st_cell[i];
  if (c < 0) return;
  const int col = c / p.num_rows, row = c - col * p.num_rows;
  int slot = tstart[tq_of(p, col, row)] + atomicSub(&count[c], 1) - 1;
  perm[slot] = i;
}

// Stable placement: staged particle i goes to (start of its cell) + (rank of its original id inside the cell), i.e.
// ascending ids per cell like the reference's fill loop (src/chunks.jl:153-155).  Makes the layout (and every force
// summation order) independent of atomic scheduling -> bit-reproducible runs.  Inactive particles go to the tail.
__global__ void k_build_place(const __grid_constant__ DevParams p, const int *__restrict__ st_cell,
                              const int *__restrict__ tstart, const int *__restrict__ perm,
                              const real2 *__restrict__ st_pos, const real2 *__restrict__ st_vel,
                              const real *__restrict__ st_ang, const real2 *__restrict__ st_force,
                              const unsigned int *__restrict__ st_id, real2 *__restrict__ pos,
                              real2 *__restrict__ vel, real *__restrict__ ang, real2 *__restrict__ force,
                              unsigned int *__restrict__ idflag, int *__restrict__ cell, int *__restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n || flags[FLAG_OVERFLOW]) return;
  const int c = st_cell[i];
  const unsigned int idf = st_id[i];
  int d;
  if (c < 0) {
    d = p.tail_base + atomicAdd(&flags[FLAG_TAIL], 1);
  } else {
    const int col = c / p.num_rows, row = c - col * p.num_rows;
    const int q = tq_of(p, col, row);
    const int b = tstart[q], e = tstart[q + 1];
    const unsigned int myid = idf & ~MAVI_INACTIVE_BIT;
    int rank = 0;
    for (int t = b; t < e; t++) {
      const int o = perm[t];
      if (o != i) rank += (st_id[o] & ~MAVI_INACTIVE_BIT) < myid;
    }
    d = b + rank;
  }
  pos[d] = st_pos[i];
  if (vel) vel[d] = st_vel[i];
  if (ang) ang[d] = st_ang[i];
  force[d] = st_force[i];
  idflag[d] = idf;
  cell[d] = c;
}

// tile populations -> (scan) -> tile_prefix, then the first tile of every 256-rank block
__global__ void k_tile_counts(const __grid_constant__ DevParams p, const int *__restrict__ tstart, int *__restrict__ out) {
  int o = blockIdx.x * blockDim.x + threadIdx.x;  // tile-row-major order index
  if (o > p.nt_ord) return;
  int cnt = 0;
  if (o < p.nt_ord) {
    const int t = tile_of_order(p, o);
    cnt = tstart[(size_t)t * (MAVI_TR + 1) + MAVI_TR] - t * p.cap;
  }
  out[o] = cnt;
}

__global__ void k_cta_first(const __grid_constant__ DevParams p, const int *__restrict__ tile_prefix,
                            int *__restrict__ cta_first) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.nt_ord) return;
  const int lo = tile_prefix[t], hi = tile_prefix[t + 1];
  for (int b = (lo + RPB - 1) / RPB; b * RPB < hi; b++) cta_first[b] = t;
}

// ---- exclusive prefix scan (reduce / top / final), 4096 items per block ------------------------------------
constexpr int SCAN_TPB = 256, SCAN_ITEMS = 16, SCAN_BLOCK = SCAN_TPB * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
  __shared__ int warp_sums[SCAN_TPB / 32];
  __shared__ int s_total;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[w] = incl;
  __syncthreads();
  if (w == 0) {
    int ws = lane < SCAN_TPB / 32 ? warp_sums[lane] : 0;
    int wi = ws;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < SCAN_TPB / 32) warp_sums[lane] = wi - ws;
    if (lane == SCAN_TPB / 32 - 1) s_total = wi;
  }
  __syncthreads();
  if (total) *total = s_total;
  return incl - v + warp_sums[w];
}

__global__ void k_scan_reduce(const int *__restrict__ in, int *__restrict__ partials, int n) {
  int base = blockIdx.x * SCAN_BLOCK;
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    int idx = base + i * SCAN_TPB + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  int tot;
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) partials[blockIdx.x] = tot;
}

__global__ void k_scan_top(int *__restrict__ partials, int nb) {
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += SCAN_TPB) {
    int idx = base + threadIdx.x;
    int v = idx < nb ? partials[idx] : 0;
    int tot;
    int ex = block_exclusive_scan(v, &tot);
    int carry = carry_s;
    if (idx < nb) partials[idx] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
}

__global__ void k_scan_final(const int *__restrict__ in, int *__restrict__ out, const int *__restrict__ partials, int n) {
  int base = blockIdx.x * SCAN_BLOCK + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    int idx = base + i;
    v[i] = idx < n ? in[idx] : 0;
    s += v[i];
  }
  int ex = block_exclusive_scan(s, nullptr) + partials[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) {
    int idx = base + i;
    if (idx < n) out[idx] = ex;
    ex += v[i];
  }
}

void launch_exclusive_scan(const LaunchCtx &c, const int *in, int *out, int *partials, int n) {
  int nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  MAVI_LAUNCH(c, k_scan_reduce, nb, SCAN_TPB, 0, in, partials, n);
  MAVI_LAUNCH(c, k_scan_top, 1, SCAN_TPB, 0, partials, nb);
  MAVI_LAUNCH(c, k_scan_final, nb, SCAN_TPB, 0, in, out, partials, n);
}

void refresh_rank_maps(const LaunchCtx &c, const DevParams &p, const DevArrays &a) {
  if (p.nt == 0) return;
  int *tmp = a.perm;  // nt+1 ints of scratch
  MAVI_LAUNCH(c, k_tile_counts, nblk(p.nt_ord + 1), TPB, 0, p, a.tstart, tmp);
  launch_exclusive_scan(c, tmp, a.tile_prefix, a.scan_partials, p.nt_ord + 1);
  MAVI_LAUNCH(c, k_cta_first, nblk(p.nt_ord), TPB, 0, p, a.tile_prefix, a.cta_first);
  if (c.maps_valid) *c.maps_valid = true;
}

void ensure_rank_maps(const LaunchCtx &c, const DevParams &p, const DevArrays &a) {
  if (!c.maps_valid || !*c.maps_valid) refresh_rank_maps(c, p, a);
}

void launch_build_tiles(const LaunchCtx &c, const DevParams &p, const DevArrays &a, bool second_is_vel) {
  // caller has zeroed count[], flags[FLAG_OVERFLOW], flags[FLAG_MAXCOUNT], flags[7]
  MAVI_LAUNCH(c, k_build_cell_index, nblk(p.n), TPB, 0, p, a.st_pos, a.st_id, a.st_cell, a.count, a.flags);
  MAVI_LAUNCH(c, k_build_layout, nblk(p.nt), TPB, 0, p, a.count, a.tstart, a.flags);
  MAVI_LAUNCH(c, k_build_scatter, nblk(p.n), TPB, 0, p, a.st_cell, a.tstart, a.count, a.perm, a.flags);
  MAVI_LAUNCH(c, k_build_place, nblk(p.n), TPB, 0, p, a.st_cell, a.tstart, a.perm, a.st_pos,
              second_is_vel ? a.st_vel : nullptr, second_is_vel ? nullptr : a.st_ang, a.st_force, a.st_id, a.pos[0],
              second_is_vel ? a.vel : nullptr, second_is_vel ? nullptr : a.ang, a.force, a.idflag, a.cell, a.flags);
  refresh_rank_maps(c, p, a);
}


// ---- index-only tiles (Mavi.Rings): the state stays ring-ordered; only particle INDICES are binned -----------------
// after the scatter every cell's slot range of perm[] is sorted ascending (= ascending ids, src/chunks.jl:153-155)
// ... and spos[] (optional) receives the positions in slot order for the pair kernel
__global__ void k_sort_perm_cells(const __grid_constant__ DevParams p, const int *__restrict__ tstart,
                                  int *__restrict__ perm, const int *__restrict__ flags,
                                  const real2 *__restrict__ pos, real2 *__restrict__ spos) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.num_cells || flags[FLAG_OVERFLOW]) return;
  const int col = div_rows(p, c), row = c - col * p.num_rows;
  const int q = tq_of(p, col, row);
  const int b = tstart[q], e = tstart[q + 1];
  for (int i = b + 1; i < e; i++) {
    const int v = perm[i];
    int j = i - 1;
    while (j >= b && perm[j] > v) {
      perm[j + 1] = perm[j];
      --j;
    }
    perm[j + 1] = v;
  }
  if (spos)
    for (int i = b; i < e; i++) spos[i] = pos[perm[i]];
}

void launch_build_index_tiles(const LaunchCtx &c, const DevParams &p, const real2 *pos, const unsigned int *idflag,
                              int *cell_out, int *count, int *tstart, int *perm, int *flags, real2 *spos) {
  MAVI_LAUNCH(c, k_build_cell_index, nblk(p.n), TPB, 0, p, pos, idflag, cell_out, count, flags);
  MAVI_LAUNCH(c, k_build_layout, nblk(p.nt), TPB, 0, p, count, tstart, flags);
  MAVI_LAUNCH(c, k_build_scatter, nblk(p.n), TPB, 0, p, cell_out, tstart, count, perm, flags);
  MAVI_LAUNCH(c, k_sort_perm_cells, nblk(p.num_cells), TPB, 0, p, tstart, perm, flags, pos, spos);
}

// dense copy of the current state into the staging arrays, in rank order
__global__ void k_compact(const __grid_constant__ DevParams p, const int *__restrict__ tile_prefix,
                          const int *__restrict__ cta_first, const real2 *__restrict__ pos,
                          const real2 *__restrict__ vel, const real *__restrict__ ang,
                          const real2 *__restrict__ force, const unsigned int *__restrict__ idflag,
                          real2 *__restrict__ st_pos, real2 *__restrict__ st_vel, real *__restrict__ st_ang,
                          real2 *__restrict__ st_force, unsigned int *__restrict__ st_id) {
  int rank = blockIdx.x * blockDim.x + threadIdx.x;
  if (rank >= p.n) return;
  const int k = slot_of_rank(p, tile_prefix, cta_first, rank);
  st_pos[rank] = pos[k];
  if (vel) st_vel[rank] = vel[k];
  if (ang) st_ang[rank] = ang[k];
  st_force[rank] = force[k];
  st_id[rank] = idflag[k];
}

// GLOBAL cell id (the reference's, src/chunks.jl:129-130) of every owned particle, in the rank order of k_compact
__global__ void k_compact_cells(const __grid_constant__ DevParams p, const int *__restrict__ tile_prefix,
                                const int *__restrict__ cta_first, const int *__restrict__ cell, int *__restrict__ out) {
  int rank = blockIdx.x * blockDim.x + threadIdx.x;
  if (rank >= p.n) return;
  int c = -1;
  if (rank < p.n_active) {
    c = cell[slot_of_rank(p, tile_prefix, cta_first, rank)];
    if (p.slab) {  // local frame [0 = left halo, 1..m owned, m+1 = right halo] -> global column
      const int lcol = div_rows(p, c), row = c - lcol * p.num_rows;
      int g = lcol - 1 + p.col_lo;
      if (g < 0) g += p.gcols;
      else if (g >= p.gcols) g -= p.gcols;
      c = g * p.num_rows + row;
    }
  }
  out[rank] = c;
}

void launch_compact_cells(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int *out) {
  ensure_rank_maps(c, p, a);
  MAVI_LAUNCH(c, k_compact_cells, nblk(p.n), TPB, 0, p, a.tile_prefix, a.cta_first, a.cell, out);
}

void launch_compact_to_staging(const LaunchCtx &c, const DevParams &p, const DevArrays &a, bool second_is_vel) {
  ensure_rank_maps(c, p, a);
  MAVI_LAUNCH(c, k_compact, nblk(p.n), TPB, 0, p, a.tile_prefix, a.cta_first, a.pos[0], second_is_vel ? a.vel : nullptr,
              second_is_vel ? nullptr : a.ang, a.force, a.idflag, a.st_pos, second_is_vel ? a.st_vel : nullptr,
              second_is_vel ? nullptr : a.st_ang, a.st_force, a.st_id);
}

// =========================================================================================================
// Incremental update_chunks!: the integrate kernels have already written the fresh cell of every particle that
// left its cell (cell[k]), marked its tile dirty and queued inter-tile movers in the destination tile's inbox.
// =========================================================================================================

// copy the records of inter-tile movers aside so that source tiles can be rewritten independently
__global__ void k_repair_collect(const int *__restrict__ flags, const int *__restrict__ mv_src,
                                 const real2 *__restrict__ pos, const real2 *__restrict__ vel,
                                 const real *__restrict__ ang, const real2 *__restrict__ force,
                                 const unsigned int *__restrict__ idflag, const int *__restrict__ cell,
                                 real2 *__restrict__ mv_pos, real2 *__restrict__ mv_second,
                                 real2 *__restrict__ mv_force, unsigned int *__restrict__ mv_id,
                                 int *__restrict__ mv_cell, int mv_cap) {
  if (!flags[FLAG_RAN]) return;
  const int n = min(flags[FLAG_NMV], mv_cap);
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) {
    const int k = mv_src[m];
    if (k < 0) continue;  // immigrant from another GPU: its record is already in the mover arrays (slab.cu)
    mv_pos[m] = pos[k];
    mv_second[m] = vel ? vel[k] : make_real2(ang[k], 0.0);
    mv_force[m] = force[k];
    mv_id[m] = idflag[k];
    mv_cell[m] = cell[k];
  }
}

// One warp per dirty tile.  CHECK pass: population after the repair must fit (else FLAG_OVERFLOW -> the host rebuilds
// with a larger capacity; nothing has been modified).  WRITE pass: stayers + arrivals are sorted by (cell, id) and
// the tile and its tstart[] row are rewritten in place (all reads are staged in shared memory first).
constexpr int REPAIR_WARPS = 4;
struct RepairRec {
  real2 pos, second, force;
  unsigned long long key;  // (cell << 32) | id  -> ascending cells, ascending ids inside a cell
  unsigned int idflag;
  int cell;
};

template <bool WRITE>
__global__ void __launch_bounds__(REPAIR_WARPS * 32) k_repair_tiles(
    const __grid_constant__ DevParams p, int *__restrict__ flags, const int *__restrict__ dirty_list,
    int *__restrict__ tile_dirty, int *__restrict__ inbox_cnt, const int *__restrict__ inbox, int *__restrict__ tstart,
    real2 *__restrict__ pos, real2 *__restrict__ vel, real *__restrict__ ang, real2 *__restrict__ force,
    unsigned int *__restrict__ idflag, int *__restrict__ cell, const real2 *__restrict__ mv_pos,
    const real2 *__restrict__ mv_second, const real2 *__restrict__ mv_force, const unsigned int *__restrict__ mv_id,
    const int *__restrict__ mv_cell) {
  extern __shared__ unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  RepairRec *rec = reinterpret_cast<RepairRec *>(smem_raw) + (size_t)w * p.cap;
  const int ndirty = flags[FLAG_CHANGED];
  if (!flags[FLAG_RAN]) return;  // the step did not run (latched overflow / out-of-grid of an earlier step)
  if (WRITE && flags[FLAG_OVERFLOW]) return;
  // CHECK pass: FLAG_MAXCOUNT is a high-water mark of every tile's population (full builds and the WRITE pass raise it) and
  // FLAG_INBOX_STEP the largest inbox of this step — if even their sum fits, no tile can overflow and the pass has nothing
  // to find (Szabo C3: 63 us of the 0.86 ms step)
  if (!WRITE && flags[FLAG_MAXCOUNT] + flags[FLAG_INBOX_STEP] <= p.cap && flags[FLAG_INBOX_STEP] <= p.inbox_cap) return;
  for (int d = blockIdx.x * REPAIR_WARPS + w; d < ndirty; d += gridDim.x * REPAIR_WARPS) {
    const int t = dirty_list[d];
    const int base = t * p.cap;
    int *ts = tstart + (size_t)t * (MAVI_TR + 1);
    const int cnt_old = ts[MAVI_TR] - base;
    const int nin = inbox_cnt[t];
    // ---- gather: stayers of this tile ...
    int m = 0;
    for (int l0 = 0; l0 < cnt_old; l0 += 32) {
      const int l = l0 + lane;
      bool stay = false;
      int c = 0;
      if (l < cnt_old) {
        c = cell[base + l];
        stay = tile_of_cell(p, c) == t;
      }
      const unsigned int bal = __ballot_sync(0xffffffffu, stay);
      if (WRITE && stay) {
        const int e = m + __popc(bal & ((1u << lane) - 1));
        RepairRec &r = rec[e];
        const int k = base + l;
        r.pos = pos[k];
        r.second = vel ? vel[k] : make_real2(ang[k], 0.0);
        r.force = force[k];
        r.idflag = idflag[k];
        r.cell = c;
        r.key = ((unsigned long long)(unsigned int)c << 32) | (r.idflag & ~MAVI_INACTIVE_BIT);
      }
      m += __popc(bal);
    }
    const int total = m + nin;
    if (!WRITE) {
      if (lane == 0) {
        atomicMax(&flags[FLAG_MAXCOUNT], total);
        if (total > p.cap) atomicOr(&flags[FLAG_OVERFLOW], 1);
        if (nin > p.inbox_cap) {
          atomicOr(&flags[FLAG_OVERFLOW], 2);
          if (atomicMax(&flags[FLAG_MAXINBOX], nin) < nin) flags[FLAG_MAXINBOX_TILE] = t;
        }
      }
      continue;
    }
    if (!WRITE) continue;  // (keeps the compiler from warning about the unreachable tail in the CHECK instantiation)
    if (lane == 0) raise_mark(&flags[FLAG_MAXCOUNT], total);
    // ---- ... plus arrivals from other tiles
    for (int i = lane; i < nin; i += 32) {
      const int mi = inbox[(size_t)t * p.inbox_cap + i];
      RepairRec &r = rec[m + i];
      r.pos = mv_pos[mi];
      r.second = mv_second[mi];
      r.force = mv_force[mi];
      r.idflag = mv_id[mi];
      r.cell = mv_cell[mi];
      r.key = ((unsigned long long)(unsigned int)r.cell << 32) | (r.idflag & ~MAVI_INACTIVE_BIT);
    }
    __syncwarp();
    // ---- rank sort (keys are unique) and in-place rewrite
    for (int e = lane; e < total; e += 32) {
      const unsigned long long key = rec[e].key;
      int rank = 0;
      for (int o = 0; o < total; o++) rank += rec[o].key < key;
      const int k = base + rank;
      const RepairRec &r = rec[e];
      pos[k] = r.pos;
      if (vel) vel[k] = r.second;
      else ang[k] = r.second.x;
      force[k] = r.force;
      idflag[k] = r.idflag;
      cell[k] = r.cell;
    }
    // ---- tstart row of the tile: lane lr counts the particles of local row lr, warp scan
    {
      const int col = t / p.tpc, tr = t - col * p.tpc;
      const int mycell = col * p.num_rows + tr * MAVI_TR + lane;
      int cnt = 0;
      for (int o = 0; o < total; o++) cnt += rec[o].cell == mycell;
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      ts[lane] = base + incl - cnt;
      if (lane == 31) ts[MAVI_TR] = base + incl;
    }
    if (lane == 0) {
      tile_dirty[t] = 0;
      inbox_cnt[t] = 0;
    }
    __syncwarp();
  }
}

void launch_repair_tiles(const LaunchCtx &c, const DevParams &p, const DevArrays &a, bool second_is_vel) {
  if (p.nt == 0) return;
  real2 *vel = second_is_vel ? a.vel : nullptr;
  real *ang = second_is_vel ? nullptr : a.ang;
  const size_t smem = (size_t)REPAIR_WARPS * p.cap * sizeof(RepairRec);
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_repair_tiles<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int grid = 148 * 8;  // one warp per dirty tile at a time: latency-bound, so as many warps as fit (8 CTAs of 4 warps per SM)
  MAVI_LAUNCH(c, (k_repair_tiles<false>), grid, REPAIR_WARPS * 32, 0, p, a.flags, a.dirty_list, a.tile_dirty, a.inbox_cnt,
              a.inbox, a.tstart, a.pos[0], vel, ang, a.force, a.idflag, a.cell, a.mv_pos, a.mv_second, a.mv_force, a.mv_id,
              a.mv_cell);
  MAVI_LAUNCH(c, k_repair_collect, 64, TPB, 0, a.flags, a.mv_src, a.pos[0], vel, ang, a.force, a.idflag, a.cell, a.mv_pos,
              a.mv_second, a.mv_force, a.mv_id, a.mv_cell, p.mv_cap);
  MAVI_LAUNCH(c, (k_repair_tiles<true>), grid, REPAIR_WARPS * 32, smem, p, a.flags, a.dirty_list, a.tile_dirty, a.inbox_cnt,
              a.inbox, a.tstart, a.pos[0], vel, ang, a.force, a.idflag, a.cell, a.mv_pos, a.mv_second, a.mv_force, a.mv_id,
              a.mv_cell);
  if (c.maps_valid) *c.maps_valid = false;  // tile populations changed; nothing in the step loop needs the rank maps
  else refresh_rank_maps(c, p, a);
}

// =========================================================================================================
// Pair forces (calc_forces!, src/integration.jl:112-224) as a per-particle gather, fused with the integrators.
// =========================================================================================================

template <int DYN, bool MINIMG>
__device__ __forceinline__ void accumulate_pair(const DevParams &p, real2 ri, real2 rj, real &fx, real &fy) {
  real dx = min_image<MINIMG>(ri.x - rj.x, p.half[0], p.size[0]);
  real dy = min_image<MINIMG>(ri.y - rj.y, p.half[1], p.size[1]);
  // laws with a cutoff need r2 rounded exactly like the reference (bit-exact neighbour decisions); LJ has no cutoff
  real r2 = (DYN == MAVI_DYN_LJ) ? fma(dx, dx, dy * dy) : dist2_exact(dx, dy);
  real c = pair_coef<DYN>(p, r2);
  fx = fma(c, dx, fx);
  fy = fma(c, dy, fy);
}

// All-pairs mode (chunks === nothing, src/integration.jl:197-224): every active particle interacts with every other
// active one in ascending-id order; slots are the original ids.  O(N^2): the reference's own small-N path (README
// quick start, C1), kept as simple thread-per-particle kernels.
template <int DYN, bool PER>
__device__ __forceinline__ real2 allpairs_force(const DevParams &p, const real2 *__restrict__ pos,
                                                const unsigned int *__restrict__ idflag, int k, real2 r) {
  real fx = 0.0, fy = 0.0;
  for (int j = 0; j < p.n; j++) {
    if (j == k || (idflag[j] & MAVI_INACTIVE_BIT)) continue;
    accumulate_pair<DYN, PER>(p, r, __ldg(pos + j), fx, fy);
  }
  return make_real2(fx, fy);
}

// The particle in slot k (sorted under cell c_old) now sits at (x, y).  If update_particle_chunk! would bin it
// elsewhere, record its fresh cell, mark the tiles involved for the incremental repair and queue it in the
// destination tile's inbox when it changes tile.
struct MoverSink {
  int *cell, *tile_dirty, *dirty_list, *inbox_cnt, *inbox, *mv_src, *flags;
  int *chg;  // changed-cell list of the force carry (nullptr: not recorded)
  // slab mode: emigrant records towards the left [0] / right [1] neighbour
  EmRec *em0, *em1;  // (two scalars, not an array: a dynamically indexed member would force the whole kernel parameter
                     // struct into local memory, one copy per thread)
  int em_cap;
  const unsigned int *idflag;
  // cell-edge tables of the grid (k_cell_edges); read by the pipelined kernels' producer warp
  const double2 *edge_x, *edge_y;
};

__device__ __forceinline__ void mark_dirty(const MoverSink &ms, int t) {
  if (atomicExch(&ms.tile_dirty[t], 1) == 0) ms.dirty_list[atomicAdd(&ms.flags[FLAG_CHANGED], 1)] = t;
}

// Force carry: cells whose membership (or a member's position) changed behind the back of the pair pass of this step.
// Every particle that has such a cell in its stencil gets its next F1 recomputed (k_recompute_changed).
__device__ __forceinline__ void note_changed_cells(const DevParams &p, const MoverSink &ms, int c0, int c1) {
  if (!ms.chg) return;
  const int q = atomicAdd(&ms.flags[FLAG_NCHG], 2);
  if (q + 1 < p.chg_cap) {
    ms.chg[q] = c0;
    ms.chg[q + 1] = c1;
  } else {
    atomicOr(&ms.flags[FLAG_OVERFLOW], 8);
  }
}

// The particle in slot k (binned under c_old) is no longer inside that cell: the rare path of note_if_moved, kept out
// of line so that it costs the hot kernels no registers.
__device__ __noinline__ void note_moved_slow(const DevParams &p, const MoverSink &ms, int k, int c_old, real x, real y,
                                            bool fixed, real2 force, real2 second) {
  int c_new = cell_of_point(p, x, y);
  if (c_new < 0) {  // left the grid: the reference throws BoundsError at its NEXT update_chunks! -> reported then
    atomicOr(&ms.flags[FLAG_ERR], ERRBIT_OOG_PENDING);
    return;
  }
  if (c_new == c_old) {
    if (fixed) note_changed_cells(p, ms, c_old, c_old);
    return;
  }
  note_changed_cells(p, ms, c_old, c_new);
  atomicAdd(&ms.flags[FLAG_NMOVED], 1);
  ms.cell[k] = c_new;
  const int t_old = tile_of_cell(p, c_old), t_new = tile_of_cell(p, c_new);
  mark_dirty(ms, t_old);
  if (p.slab) {
    const int lc = div_rows(p, c_new);
    if (lc == 0 || lc == p.num_cols - 1) {  // crossed into a halo column: goes to the neighbour rank, not into a tile
      if (p.blk_mode == 1) {  // interior block of the pipelined slab step: the emigrant buffers are already on their way
        atomicOr(&ms.flags[FLAG_ERR], ERRBIT_OUT_OF_GRID);  // >= 3 cell columns in ONE step: the run has blown up
        return;
      }
      const int d = lc == 0 ? 0 : 1;
      const int i = atomicAdd(&ms.flags[FLAG_NEM0 + d], 1);
      if (i < ms.em_cap) {
        EmRec &e = (d == 0 ? ms.em0 : ms.em1)[i];
        e.pos = make_real2(x, y);
        e.second = second;
        e.force = force;
        e.idflag = ms.idflag[k];
      } else {
        atomicOr(&ms.flags[FLAG_OVERFLOW], 16);
      }
      return;
    }
  }
  if (t_new != t_old) {
    mark_dirty(ms, t_new);
    const int m = atomicAdd(&ms.flags[FLAG_NMV], 1);
    const int i = atomicAdd(&ms.inbox_cnt[t_new], 1);
    raise_mark(&ms.flags[FLAG_INBOX_STEP], i + 1);
    if (m < p.mv_cap && i < p.inbox_cap) {
      ms.mv_src[m] = k;
      ms.inbox[(size_t)t_new * p.inbox_cap + i] = m;
    } else {
      atomicOr(&ms.flags[FLAG_OVERFLOW], m < p.mv_cap ? 2 : 4);
    }
  }
}

// The particle in slot k (sorted under cell c_old) now sits at (x, y).  If update_particle_chunk! would bin it
// elsewhere, record its fresh cell, mark the tiles involved for the incremental repair and queue it in the destination
// tile's inbox (or, in slab mode, in the emigrant records of the neighbour rank it moves to).
// `fixed`: walls! moved the particle (periodic wrap, slippery projection) after the pair pass read its position.
// second_of(): the particle's second state record (velocity / angle) after this step — only evaluated for a particle
// that left its cell, to fill the emigrant record together with `force`.
// in_cell(x, y): the exact "update_particle_chunk! would bin it into c_old again" test — still_in_cell() itself (InCellExact)
// or the same comparisons against per-chunk tables of the cell edges (InCellTab, pipelined kernels).
struct InCellExact {
  const DevParams &p;
  int c;
  __device__ __forceinline__ bool operator()(double x, double y) const { return still_in_cell(p, x, y, c); }
};
struct InCellTab {  // edges of the particle's cell as axis_in_cell() compares them: lo <= t < hi on both axes
  const DevParams &p;
  double xlo, xhi, ylo, yhi;
  __device__ __forceinline__ bool operator()(double x, double y) const {
    const double tx = x - p.grid_bl[0], ty = -y + p.grid_bl[1] + p.grid_h;
    return tx >= xlo && tx < xhi && ty >= ylo && ty < yhi;
  }
};
struct InCellNone {
  __device__ __forceinline__ bool operator()(double, double) const { return true; }
};

template <typename SecondF, typename InCell>
__device__ __forceinline__ void note_if_moved(const DevParams &p, const MoverSink &ms, int k, int c_old, real x,
                                              real y, bool fixed, real2 force, SecondF &&second_of, const InCell &in_cell) {
  if (in_cell(x, y)) {
    if (fixed) note_changed_cells(p, ms, c_old, c_old);
    return;
  }
  note_moved_slow(p, ms, k, c_old, x, y, fixed, force, second_of());
}

// Drift of update_verlet! (src/integration.jl:424): pos + vel dt + F dt^2/2, written with explicit FMAs so that every
// kernel that produces a drifted position (k_newton_a, the carry in k_newton_b, the sparse fix-up kernels) rounds alike.
__device__ __forceinline__ real2 verlet_drift(const DevParams &p, real2 r, real2 v, real2 F, bool &big) {
  const real mx = fma(v.x, p.dt, F.x * p.term), my = fma(v.y, p.dt, F.y * p.term);
  // displacement guard of the min-image shortcut: a drift beyond one cell in one step makes the pair pass that reads
  // the drifted positions on stale cells take the exact (minimum-image everywhere) path.
  big = !(fabs(mx) <= p.cl && fabs(my) <= p.ch);
  return make_real2(r.x + mx, r.y + my);
}

// pull a line towards L1 without tying up a register across the pair loop (the value is loaded after the loop)
__device__ __forceinline__ void prefetch_l1(const void *ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }

// each thread of the all-pairs kernels handles RPB/TPB slots of its block
#define MAVI_FOR_EACH_SLOT                                           \
  for (int it = 0; it < RPB / TPB; ++it) {                           \
    const int k = blockIdx.x * RPB + it * TPB + threadIdx.x;         \
    if (k >= p.n) break;

// clean_forces! + calc_forces! (+ calc_walls_forces!): the force state after src/integration.jl:508-511.
template <int DYN, bool PER>
__global__ void __launch_bounds__(TPB) k_force_only(const __grid_constant__ DevParams p,
                                                    const unsigned int *__restrict__ idflag,
                                                    const real2 *__restrict__ pos, real2 *__restrict__ force,
                                                    int with_walls) {
  MAVI_FOR_EACH_SLOT
    const real2 r = pos[k];
    real2 F = make_real2(0.0, 0.0);
    if (!(idflag[k] & MAVI_INACTIVE_BIT)) {
      F = allpairs_force<DYN, PER>(p, pos, idflag, k, r);
      if (with_walls && p.has_force_walls) wall_forces(p, r.x, r.y, F.x, F.y);
    }
    force[k] = F;
  }
}

// newton_step! first half (src/integration.jl:507-512 + update_verlet! :418-424):
//   F1 = pair forces + wall forces;  pos' = pos + vel dt + F1 dt^2/2  (every slot, active or not).
template <int DYN, bool PER>
__global__ void __launch_bounds__(TPB) k_newton_a(const __grid_constant__ DevParams p,
                                                  const unsigned int *__restrict__ idflag,
                                                  const real2 *__restrict__ pos_in, const real2 *__restrict__ vel,
                                                  real2 *__restrict__ pos_out, real2 *__restrict__ f1,
                                                  int *__restrict__ flags) {
  if (!flags[FLAG_RAN]) return;
  MAVI_FOR_EACH_SLOT
    real2 r = pos_in[k];
    real2 F = make_real2(0.0, 0.0);
    if (!(idflag[k] & MAVI_INACTIVE_BIT)) {
      F = allpairs_force<DYN, PER>(p, pos_in, idflag, k, r);
      if (p.has_force_walls) wall_forces(p, r.x, r.y, F.x, F.y);
    }
    bool big;
    r = verlet_drift(p, r, vel[k], F, big);
    pos_out[k] = r;
    f1[k] = F;
  }
}

// newton_step! second half (update_verlet! :426-430, walls! :513): F2 on the drifted positions with the STALE cell
// lists and WITHOUT wall forces; vel += dt/2 (F2 + F1); walls!(active ids).  Position changes made by walls!
// (periodic wrap, slippery projection) are deferred to a sparse fix-up list because neighbours still read the
// unmodified drifted positions in this launch; the fresh cell of the FINAL position feeds the incremental repair.
//
// CARRY (force carry, default for chunked Newton runs): the first half of the NEXT newton_step! recomputes the pair
// forces at exactly the positions this kernel just used — F1(n+1) differs from F2(n) only for particles that have, in
// their stencil, a cell whose membership changed (the re-binning between the two calls) or a particle that walls!
// moved.  So this kernel also writes the next drift  pos'' = pos' + vel' dt + (F2 + wall forces) dt^2/2  and records the
// changed cells; after the tile repair k_redrift_tiles / k_recompute_changed redo F1 and the drift for the few affected
// particles with the fresh cell lists, and the next step starts directly with this kernel.  Results are bit-identical
// to running k_newton_a every step (tests/test_gpu_core.py::test_force_carry_bitwise).
//   f1 / f1_next: F1 of this step / of the next one (the same array, updated in place; f2 = get_forces stays F2)
// All-pairs version (no cell lists, hence no re-binning and no carry); the tile-block version is k_newton_b2.
template <int DYN, bool PER>
__global__ void __launch_bounds__(TPB) k_newton_b(const __grid_constant__ DevParams p,
                                                  const unsigned int *__restrict__ idflag,
                                                  const real2 *__restrict__ pos_in, real2 *__restrict__ vel,
                                                  const real2 *__restrict__ f1, real2 *__restrict__ f2,
                                                  int *__restrict__ fix_idx, real2 *__restrict__ fix_pos,
                                                  int *__restrict__ flags) {
  if (!flags[FLAG_RAN]) return;
  MAVI_FOR_EACH_SLOT
    real2 r = pos_in[k];
    real2 F = make_real2(0.0, 0.0);
    const bool active = !(idflag[k] & MAVI_INACTIVE_BIT);
    if (active) F = allpairs_force<DYN, PER>(p, pos_in, idflag, k, r);
    real2 v = vel[k];
    const real2 Fo = f1[k];
    v.x = v.x + p.hdt * (F.x + Fo.x);
    v.y = v.y + p.hdt * (F.y + Fo.y);
    if (active) {
      const real x0 = r.x, y0 = r.y;
      apply_walls<true>(p, r.x, r.y, v.x, v.y, p.particle_radius);
      if (r.x != x0 || r.y != y0) {  // neighbours still read the unmodified drifted positions in this launch
        int m = atomicAdd(&flags[FLAG_NFIX], 1);
        fix_idx[m] = k;
        fix_pos[m] = r;
      }
    }
    vel[k] = v;
    f2[k] = F;
  }
}

// Force carry, after the tile repair (fresh cell lists, current positions in `pos`):
// (1) the repair moved pos / vel / force of every particle of a dirty tile to new slots but not the carried drift:
//     redo it (and F1 = F2 + wall forces) for those tiles, one warp per dirty tile.
__global__ void k_redrift_tiles(const __grid_constant__ DevParams p, int *__restrict__ flags,
                                const int *__restrict__ dirty_list, const int *__restrict__ tstart,
                                const real2 *__restrict__ pos, const real2 *__restrict__ vel,
                                const real2 *__restrict__ force, real2 *__restrict__ f1_next,
                                real2 *__restrict__ pos_next) {
  if (!flags[FLAG_RAN] || flags[FLAG_OVERFLOW]) return;
  const int lane = threadIdx.x & 31, w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int nd = flags[FLAG_CHANGED];
  for (int d = w; d < nd; d += nw) {
    const int t = dirty_list[d];
    const int b = t * p.cap, e = tstart[(size_t)t * (MAVI_TR + 1) + MAVI_TR];
    for (int k = b + lane; k < e; k += 32) {
      const real2 r = pos[k];
      real2 F = force[k];
      if (p.has_force_walls) wall_forces(p, r.x, r.y, F.x, F.y);
      f1_next[k] = F;
      bool big;
      pos_next[k] = verlet_drift(p, r, vel[k], F, big);
      if (p.periodic && big) flags[FLAG_BIGMOVE_NEXT] = 1;
    }
  }
}

// F1 (fresh cell lists, + wall forces) and the drift of the particle in slot k.  Same neighbour order (column-1 rows
// r-1..r+1, own column, column+1) and the same pair arithmetic as the staged walk -> bit-identical to k_newton_a.
template <int DYN, bool PER>
__device__ __forceinline__ void recompute_particle(const DevParams &p, int *__restrict__ flags,
                                                   const int *__restrict__ tstart, const real2 *__restrict__ pos,
                                                   const real2 *__restrict__ vel, real2 *__restrict__ f1_next,
                                                   real2 *__restrict__ pos_next, int k, int cell,
                                                   bool report_big = true) {
  const real2 r = pos[k];
  real fx = 0.0, fy = 0.0;
  for_each_neighbor(p, tstart, cell, k, [&](int j) { accumulate_pair<DYN, PER>(p, r, __ldg(pos + j), fx, fy); });
  if (p.has_force_walls) wall_forces(p, r.x, r.y, fx, fy);
  const real2 F = make_real2(fx, fy);
  f1_next[k] = F;
  bool big;
  pos_next[k] = verlet_drift(p, r, vel[k], F, big);
  if (PER && big && report_big) flags[FLAG_BIGMOVE_NEXT] = 1;
}

// (2) F1 and the drift of every particle that has a changed cell in its stencil, with the fresh cell lists: one warp
//     per list entry; lanes 0..8 look up the 9 cells of the entry's stencil (the stencil relation is symmetric), then
//     the particles of those cells are dealt out one per lane.  Entries may repeat and neighbourhoods overlap: the
//     recomputation is idempotent (reads pos / vel, writes f1_next / pos_next), so concurrent duplicates store
//     identical values.
template <int DYN, bool PER>
__global__ void k_recompute_changed(const __grid_constant__ DevParams p, int *__restrict__ flags,
                                    const int *__restrict__ chg, const int *__restrict__ tstart,
                                    const real2 *__restrict__ pos, const real2 *__restrict__ vel,
                                    real2 *__restrict__ f1_next, real2 *__restrict__ pos_next, int skip_edge) {
  // slab mode: halo columns hold no state of this rank (skip_edge = 1); with skip_edge = 3 the two owned columns next
  // to each halo are left to k_recompute_columns as well (it runs on the side stream, after the halo exchange)
  if (!flags[FLAG_RAN] || flags[FLAG_OVERFLOW]) return;
  const int lane = threadIdx.x & 31, w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  const int n = min(flags[FLAG_NCHG], p.chg_cap);
  const int R = p.num_rows, Cn = p.num_cols;
  for (int e = w; e < n; e += nw) {  // warp-uniform
    const int c0 = chg[e];
    if ((e & 1) && c0 == chg[e - 1]) continue;  // (c, c) pair of a particle that was moved inside its cell
    const int col0 = div_rows(p, c0), row0 = c0 - col0 * R;
    int cnt = 0, kb = 0, cellv = 0;
    if (lane < 9) {
      int c2 = col0 + lane / 3 - 1, r2 = row0 + lane % 3 - 1;
      bool ok = true;
      if (c2 < 0) { ok = p.wrap_cols; c2 = Cn - 1; }
      else if (c2 >= Cn) { ok = p.wrap_cols; c2 = 0; }
      if (r2 < 0) { ok = ok && p.wrap_rows; r2 = R - 1; }
      else if (r2 >= R) { ok = ok && p.wrap_rows; r2 = 0; }
      if (skip_edge && (c2 < skip_edge || c2 >= Cn - skip_edge)) ok = false;
      if (ok) {
        const int q = tq_of(p, c2, r2);
        kb = tstart[q];
        cnt = tstart[q + 1] - kb;
        cellv = c2 * R + r2;
      }
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 8);
    for (int base = 0; base < total; base += 32) {  // warp-uniform
      const int idx = base + lane;
      int k = -1, cell = 0;
#pragma unroll
      for (int l = 0; l < 9; l++) {
        const int il = __shfl_sync(0xffffffffu, incl, l), cl = __shfl_sync(0xffffffffu, cnt, l);
        const int kl = __shfl_sync(0xffffffffu, kb, l), ce = __shfl_sync(0xffffffffu, cellv, l);
        if (idx < il && idx >= il - cl) { k = kl + (idx - (il - cl)); cell = ce; }
      }
      if (k >= 0) recompute_particle<DYN, PER>(p, flags, tstart, pos, vel, f1_next, pos_next, k, cell);
    }
  }
}

// (3) slab mode: the neighbour rank's boundary column re-bins behind this rank's back, so every particle of the owned
//     boundary columns is recomputed every step: `depth` columns on each side (1: local columns 1 and num_cols-2;
//     2: also the next ones, when the list-driven kernel above leaves them out).  report_big = 0: no big-drift flag (the
//     blocks that read these particles take the exact minimum-image path anyway, see slab.cu).
template <int DYN, bool PER>
__global__ void k_recompute_columns(const __grid_constant__ DevParams p, int *__restrict__ flags,
                                    const int *__restrict__ tstart, const int *__restrict__ cell,
                                    const real2 *__restrict__ pos, const real2 *__restrict__ vel,
                                    real2 *__restrict__ f1_next, real2 *__restrict__ pos_next, int depth,
                                    int report_big) {
  if (!flags[FLAG_RAN] || flags[FLAG_OVERFLOW]) return;
  const int cs = p.tpc * p.cap;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * depth * cs) return;
  const int ci = i / cs;                                                    // 0 .. 2*depth-1
  const int col = ci < depth ? 1 + ci : p.num_cols - 2 - (ci - depth);      // left side, then right side
  if (ci >= depth && col <= depth) return;                                   // narrow slab: column already covered
  const int k = col * cs + (i - ci * cs);
  const int t = k / p.cap;
  if (k >= tstart[(size_t)t * (MAVI_TR + 1) + MAVI_TR]) return;  // slack slot
  recompute_particle<DYN, PER>(p, flags, tstart, pos, vel, f1_next, pos_next, k, cell[k], report_big != 0);
}

// First kernel of every step: decides ONCE whether the step runs (no overflow / out-of-grid latched by an earlier
// step) and clears the per-step control words.  Every other kernel of the step only looks at FLAG_RAN, so flags raised
// DURING the step cannot stop it half way.
__global__ void k_step_begin(int *__restrict__ flags) {
  if (threadIdx.x == 0) {
    const int run = step_poisoned(flags) ? 0 : 1;
    // instrumentation: totals of the previous step (wrap-around after 2^31 records is harmless, the host takes differences)
    flags[FLAG_CUM_CHG] += flags[FLAG_NMOVED];
    flags[FLAG_NMOVED] = 0;
    flags[FLAG_CUM_MV] += flags[FLAG_NMV];
    flags[FLAG_CUM_DIRTY] += flags[FLAG_CHANGED];
    flags[FLAG_CUM_EM] += flags[FLAG_NEM0] + flags[FLAG_NEM1];
    flags[FLAG_CHANGED] = 0;
    flags[FLAG_BIGMOVE] = flags[FLAG_BIGMOVE_NEXT];  // raised by the carried drift of the previous step
    flags[FLAG_BIGMOVE_NEXT] = 0;
    flags[FLAG_NCHG] = 0;
    flags[FLAG_NEM0] = 0;
    flags[FLAG_NEM1] = 0;
    flags[FLAG_NFIX] = 0;
    flags[FLAG_NMV] = 0;
    flags[FLAG_WORK0] = 0;
    flags[FLAG_WORK1] = 0;
    flags[FLAG_INBOX_STEP] = 0;
    flags[FLAG_RAN] = run;
  }
}

void launch_step_begin(const LaunchCtx &c, const DevArrays &a) { MAVI_LAUNCH(c, k_step_begin, 1, 32, 0, a.flags); }

// applies the deferred wall position fix-ups and counts the step as done (runs after the integrate kernels of EVERY step)
__global__ void k_apply_pos_fixes(int *__restrict__ flags, const int *__restrict__ fix_idx,
                                  const real2 *__restrict__ fix_pos, real2 *__restrict__ pos) {
  if (!flags[FLAG_RAN]) return;  // poisoned by an EARLIER step: this step does not run
  if (blockIdx.x == 0 && threadIdx.x == 0) flags[FLAG_STEPS] += 1;
  const int n = flags[FLAG_NFIX];
  for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < n; m += gridDim.x * blockDim.x) pos[fix_idx[m]] = fix_pos[m];
}

// update_szabo! (src/integration.jl:433-465) / update_rtp! (:467-498) for the particle in slot k (original id `id`)
template <int DYN>
__device__ __forceinline__ void self_propelled_update(const DevParams &p, real2 &r, const real2 F,
                                                      real *__restrict__ ang, int k, unsigned int id,
                                                      const real *__restrict__ noise, unsigned long long step, real theta) {
  real sn, cs;
  sincos(theta, &sn, &cs);
  if (DYN == MAVI_DYN_SZABO) {
    const real vo = p.dyn[0], mu = p.dyn[1], drot = p.dyn[7];
    real velx = vo * cs + mu * F.x, vely = vo * sn + mu * F.y;
    real speed = sqrt(fabs(velx) + fabs(vely));  // sqrt(sum(abs, vel)) (sic), :448
    real cross_prod = speed > 0.0 ? (cs * vely - sn * velx) / speed : 0.0;
    if (fabs(cross_prod) > 1.0) cross_prod = sign_d(cross_prod);
    real nz = 0.0;
    if (drot != 0.0) nz = (p.rng_mode == MAVI_RNG_HOST_NOISE) ? (noise ? noise[id] : 0.0) : philox_normal(p.seed, id, step);
    real d_theta = p.szabo_inv_tau * asin(cross_prod) * p.dt + p.szabo_namp * nz;  // 1/relax_time * asin(c) * dt + sqrt(2 D_r dt) * randn()
    r.x += velx * p.dt;
    r.y += vely * p.dt;
    ang[k] = theta + d_theta;
  } else {
    const real vo = p.dyn[0], tumble_rate = p.dyn[3];
    real velx = vo * cs + F.x, vely = vo * sn + F.y;
    r.x += velx * p.dt;
    r.y += vely * p.dt;
    double u, u2;  // uniforms are drawn (or read) and compared in double in both builds
    if (p.rng_mode == MAVI_RNG_HOST_NOISE) {
      u = noise ? noise[2 * (size_t)id] : 1.0;
      u2 = noise ? noise[2 * (size_t)id + 1] : 0.0;
    } else {
      philox_uniform2(p.seed, id, step, u, u2);
    }
    if (u < tumble_rate * p.dt) ang[k] = 6.283185307179586 * u2;  // 2*pi*rand(), :495
  }
}

// szabo_step! / rtp_step! (src/integration.jl:517-535): forces + update_szabo! (:433-465) / update_rtp! (:467-498)
// + walls! in ONE pass.  The update loops slots 1:count (not ids) like the reference.
template <int DYN, bool PER>
__global__ void __launch_bounds__(TPB) k_self_propelled(const __grid_constant__ DevParams p,
                                                        const unsigned int *__restrict__ idflag,
                                                        const real2 *__restrict__ pos_in, real *__restrict__ ang,
                                                        real2 *__restrict__ pos_out, real2 *__restrict__ force,
                                                        const real *__restrict__ noise, unsigned long long step,
                                                        int *__restrict__ flags) {
  if (!flags[FLAG_RAN]) return;
  MAVI_FOR_EACH_SLOT
    const unsigned int idf = idflag[k];
    const unsigned int id = idf & ~MAVI_INACTIVE_BIT;
    const bool active = !(idf & MAVI_INACTIVE_BIT);
    real2 r = pos_in[k];
    real2 F = make_real2(0.0, 0.0);
    if (active) {
      F = allpairs_force<DYN, PER>(p, pos_in, idflag, k, r);
      if (p.has_force_walls) wall_forces(p, r.x, r.y, F.x, F.y);
    }
    force[k] = F;
    if ((int)id < p.n_count) self_propelled_update<DYN>(p, r, F, ang, k, id, noise, step, ang[k]);
    if (active) {
      real vx = 0.0, vy = 0.0;
      apply_walls<false>(p, r.x, r.y, vx, vy, p.particle_radius);
    }
    pos_out[k] = r;
  }
}

// =========================================================================================================
// Tile-block force kernels (chunked runs).  A CTA owns `blk_cols` consecutive tiles of ONE tile row (no rank map, no
// search: the block geometry is arithmetic on blockIdx).  It stages, column by column, [cell row above | tile | cell
// row below] of its columns and of the two side columns into shared memory (async 16-byte copies), together with
//   cwin[j][lr]  = (first staged index of cell row lr-1, end of cell row lr+1) of staged column j  -> the three
//                  neighbour runs of a particle are three LDS.64 away,
//   list[q]      = staged index | column << 16 | row << 24 of the q-th OWN particle of the block,
// and then every thread takes own particles q = tid, tid + 256, ...  Columns are staged in chunks of as many as
// fit, so dense regions only cost more chunks; a single column that does not fit falls back to a per-thread walk of
// the global arrays.  Trailing blocks handle the inactive tail (slots of masked particles).
// =========================================================================================================
constexpr int G2MAX = 30;        // own columns per chunk (plus two side columns: one lane each in the scan)
constexpr int SPOS2_CAP = 1536;  // staged positions per chunk
constexpr int OWN2_CAP = 1280;   // own particles per chunk

struct Chunk2 {
  int nc, use_mi, nown, next, ok;
  int wr[G2MAX + 2];  // column reached through a periodic wrap / the slab seam
  int src_a[G2MAX + 2], src_t[G2MAX + 2], src_b[G2MAX + 2];
  int la[G2MAX + 2], lt[G2MAX + 2], lb[G2MAX + 2];
  int off[G2MAX + 3], ownoff[G2MAX + 3], gbase[G2MAX + 2];
  int2 cwin[G2MAX + 2][MAVI_TR];  // [j][lr-1]
  __device__ __forceinline__ int2 win(int j, int r) const { return cwin[j][r]; }
};
constexpr int C2_BYTES = (sizeof(Chunk2) + 15) / 16 * 16;
constexpr int PASS2_SMEM = C2_BYTES + SPOS2_CAP * (int)sizeof(real2) + OWN2_CAP * (int)sizeof(unsigned int);

// Stage the chunk of own columns starting at local column cs (at most `rem` columns) of tile row tr.
template <bool PER>
__device__ __forceinline__ void chunk_stage(const DevParams &p, const int *__restrict__ tstart,
                                            const real2 *__restrict__ pos, int tr, int cs, int rem, bool exact_minimg,
                                            Chunk2 *ck, real2 *s_pos, unsigned int *s_list) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int R = p.num_rows, Cn = p.num_cols;
  const int r0 = tr * MAVI_TR;
  const int rows = min(MAVI_TR, R - r0);
  const int ncand = min(rem, G2MAX);
  __syncthreads();  // previous chunk fully consumed
  // ---- column descriptors (one thread per candidate column; independent loads)
  if (threadIdx.x < ncand + 2) {
    const int j = threadIdx.x;
    int c = cs - 1 + j;
    bool exists = true, wrapped = false;
    if (c < 0) { if (p.wrap_cols) { c = Cn - 1; wrapped = true; } else exists = false; }
    else if (c >= Cn) { if (p.wrap_cols) { c = 0; wrapped = true; } else exists = false; }
    if (exists && p.slab && ((c == 0 && p.seam_left) || (c == Cn - 1 && p.seam_right))) wrapped = true;
    int ra = r0 - 1, rb = r0 + MAVI_TR;
    bool has_a = exists, has_b = exists;
    if (ra < 0) { if (p.wrap_rows) { ra = R - 1; wrapped = wrapped || exists; } else has_a = false; }
    if (rb >= R) { if (p.wrap_rows) { rb = 0; wrapped = wrapped || exists; } else has_b = false; }
    int la = 0, lt = 0, lb = 0, sa = 0, st = 0, sb = 0;
    if (exists) {
      const int *tt = tstart + (size_t)(c * p.tpc + tr) * (MAVI_TR + 1);
      const int qa = has_a ? tq_of(p, c, ra) : 0, qb = has_b ? tq_of(p, c, rb) : 0;
      st = __ldg(tt);
      const int et = __ldg(tt + MAVI_TR);
      const int a0 = has_a ? __ldg(tstart + qa) : 0, a1 = has_a ? __ldg(tstart + qa + 1) : 0;
      const int b0 = has_b ? __ldg(tstart + qb) : 0, b1 = has_b ? __ldg(tstart + qb + 1) : 0;
      lt = et - st; sa = a0; la = a1 - a0; sb = b0; lb = b1 - b0;
    }
    ck->src_t[j] = st; ck->src_a[j] = sa; ck->src_b[j] = sb;
    ck->la[j] = la; ck->lt[j] = lt; ck->lb[j] = lb;
    ck->wr[j] = wrapped ? 1 : 0;
  }
  __syncthreads();
  // ---- warp 0: prefix sums, greedy number of own columns that fit
  if (w == 0) {
    const int j = lane;
    const bool in = j < ncand + 2;
    const int sz = in ? ck->la[j] + ck->lt[j] + ck->lb[j] : 0;
    const int own = (in && j >= 1 && j <= ncand) ? ck->lt[j] : 0;
    int incl = sz, oincl = own;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o), u = __shfl_up_sync(0xffffffffu, oincl, o);
      if (lane >= o) { incl += v; oincl += u; }
    }
    const int incl_next = __shfl_down_sync(0xffffffffu, incl, 1);  // staged total if this lane were the last own column
    const bool fits = j >= 1 && j <= ncand && incl_next <= SPOS2_CAP && oincl <= OWN2_CAP;
    const unsigned int bal = __ballot_sync(0xffffffffu, fits);
    const int nc = bal ? 31 - __clz(bal) : 1;  // fits is monotone in j
    const unsigned int wbal = __ballot_sync(0xffffffffu, in && j <= nc + 1 && ck->wr[j]);
    ck->off[j] = incl - sz;
    ck->ownoff[j] = oincl - own;
    if (lane == 0) {
      ck->nc = nc;
      ck->ok = bal ? 1 : 0;
      ck->next = 0;
      ck->use_mi = (PER && (exact_minimg || !p.fast_interior || wbal)) ? 1 : 0;
    }
    if (j == nc) ck->nown = oincl;
  }
  __syncthreads();
  if (!ck->ok) return;
  const int nc = ck->nc;
  // ---- copy: one warp per column (coalesced), cell-row windows and the own-particle list
  for (int j = w; j < nc + 2; j += TPB / 32) {
    const int off = ck->off[j], la = ck->la[j], lt = ck->lt[j], lb = ck->lb[j];
    const int src_t = ck->src_t[j], src_a = ck->src_a[j], src_b = ck->src_b[j];
    const int tot = la + lt + lb;
    int cj = cs - 1 + j;
    if (cj < 0) cj = Cn - 1;
    else if (cj >= Cn) cj = 0;
    const int *tt = tstart + (size_t)(cj * p.tpc + tr) * (MAVI_TR + 1);
    // staged start of tile row lane+1 (rows beyond the grid start where the tile ends)
    const int tsl = tot ? __ldg(tt + lane) - src_t : 0, tsn = tot ? __ldg(tt + lane + 1) - src_t : 0;
    const int rs = off + la + tsl;
    const int up = __shfl_up_sync(0xffffffffu, rs, 1), dn = __shfl_down_sync(0xffffffffu, rs, 2);
    const int end = off + tot;
    const int wa = lane == 0 ? off : up;                                              // start of row lr-1
    const int wb = (lane + 3 >= rows + 2) ? end : (lane <= 29 ? dn : off + la + lt);  // end of row lr+1 = start of row lr+2
    ck->cwin[j][lane] = make_int2(wa, wb);
    if (lane == 0) ck->gbase[j] = src_t - (off + la);
    if (j >= 1 && j <= nc && lane < rows) {
      const int qb = ck->ownoff[j];
      for (int i = tsl; i < tsn; i++)
        s_list[qb + i] = (unsigned int)(off + la + i) | ((unsigned int)j << 16) | ((unsigned int)(lane + 1) << 24);
    }
    // 16-byte asynchronous copies (LDGSTS): every piece of every column is in flight at once, no register staging
    for (int i = lane; i < tot; i += 32) {
      const int src = i < la ? src_a + i : (i < la + lt ? src_t + (i - la) : src_b + (i - la - lt));
      __pipeline_memcpy_async(s_pos + off + i, pos + src, sizeof(real2));
    }
  }
  __pipeline_commit();
  __pipeline_wait_prior(0);
  __syncthreads();
}

#ifndef MAVI_WALK_UNROLL
#define MAVI_WALK_UNROLL 4  // pairs per trip of the neighbour loop (build-time A/B)
#endif
// pair force on the staged particle (staged column jj, tile row lr, staged index self) from the staged positions
template <int DYN, bool MINIMG, typename CK>
__device__ __forceinline__ void chunk_walk(const DevParams &p, const CK *ck, const real2 *s_pos, int jj, int lr,
                                           int self, real2 ri, real &fx, real &fy) {
  const int2 w0 = ck->win(jj - 1, lr - 1), w1 = ck->win(jj, lr - 1), w2 = ck->win(jj + 1, lr - 1);
  // byte offsets into s_pos; neighbours t < c1 -> column jj-1, c1 <= t < c2 -> own column before self,
  // c2 <= t < c3 -> own column after self, c3 <= t -> column jj+1
  constexpr int B = (int)sizeof(real2);  // bytes per staged position
  const int c1 = (w0.y - w0.x) * B;
  const int c2 = c1 + (self - w1.x) * B;
  const int c3 = c2 + (w1.y - self - 1) * B;
  const int total = c3 + (w2.y - w2.x) * B;
  const int o0 = w0.x * B;
  const int d1 = w1.x * B - c1 - o0, d3 = w2.x * B - c3 - (w1.x * B - c1) - B;
  const char *base = reinterpret_cast<const char *>(s_pos) + o0;
  constexpr int UNROLL = MAVI_WALK_UNROLL;
#pragma unroll UNROLL
  for (int t = 0; t < total; t += B) {
    const int o = t + (t >= c1 ? d1 : 0) + (t >= c2 ? B : 0) + (t >= c3 ? d3 : 0);
    accumulate_pair<DYN, MINIMG>(p, ri, *reinterpret_cast<const real2 *>(base + o), fx, fy);
  }
}

// Drives a force kernel: calls pre(k) before and body(k, r, cell, active, F) after the pair force F of every particle
// slot of this block (F = 0 for inactive slots).
template <int DYN, bool PER, typename Pre, typename Body>
__device__ __forceinline__ void for_each_block_particle(const DevParams &p, const int *__restrict__ tstart,
                                                        const real2 *__restrict__ pos, const int *__restrict__ cell,
                                                        bool exact_minimg, Pre &&pre, Body &&body) {
  extern __shared__ __align__(16) unsigned char dsm[];
  Chunk2 *ck = reinterpret_cast<Chunk2 *>(dsm);
  real2 *s_pos = reinterpret_cast<real2 *>(dsm + C2_BYTES);
  unsigned int *s_list = reinterpret_cast<unsigned int *>(dsm + C2_BYTES + SPOS2_CAP * sizeof(real2));
  // blk_mode 1 / 2 split a launch into the blocks that never read a halo column and the first / last block of every
  // tile row (slab mode: the halo exchange overlaps with the former)
  const int per_row = blk_items_per_row(p);
  const int nblk_tiles = per_row * p.tpc;
  if ((int)blockIdx.x >= nblk_tiles) {  // inactive tail: no pair forces
    const int i = ((int)blockIdx.x - nblk_tiles) * TPB + threadIdx.x;
    if (i < p.n - p.n_active) {
      const int k = p.tail_base + i;
      auto pv = pre(k);
      body(k, pos[k], 0, false, make_real2(0.0, 0.0), InCellNone{}, pv);
    }
    return;
  }
  const int tr = (int)blockIdx.x / per_row;
  int c_begin, c_end;
  blk_columns(p, (int)blockIdx.x - tr * per_row, c_begin, c_end);
  const int r0 = tr * MAVI_TR;
  for (int cs = c_begin; cs < c_end;) {
    chunk_stage<PER>(p, tstart, pos, tr, cs, c_end - cs, exact_minimg, ck, s_pos, s_list);
    const int nc = ck->nc;
    if (ck->ok) {
      const int nown = ck->nown;
      const bool mi = PER && ck->use_mi;
      // blk_cols is chosen so that a block holds about 1000 own particles: the last round of 256 is nearly full
      for (int q = threadIdx.x; q < nown; q += TPB) {
        const unsigned int u = s_list[q];
        const int self = u & 0xffffu, jj = (u >> 16) & 0xffu, lr = u >> 24;
        const int k = self + ck->gbase[jj];
        auto pv = pre(k);
        const real2 r = s_pos[self];
        real fx = 0.0, fy = 0.0;
        if (mi) chunk_walk<DYN, true>(p, ck, s_pos, jj, lr, self, r, fx, fy);
        else chunk_walk<DYN, false>(p, ck, s_pos, jj, lr, self, r, fx, fy);
        const int cc = (cs - 1 + jj) * p.num_rows + r0 + lr - 1;
        body(k, r, cc, true, make_real2(fx, fy), InCellExact{p, cc}, pv);
      }
    } else {
      // a single column too dense for the staging area: per-thread walk over the global arrays
      const int b = ck->src_t[1], e = b + ck->lt[1];
      for (int k = b + threadIdx.x; k < e; k += TPB) {
        auto pv = pre(k);
        const real2 r = pos[k];
        const int c = cell[k];
        real fx = 0.0, fy = 0.0;
        for_each_neighbor(p, tstart, c, k, [&](int j) { accumulate_pair<DYN, PER>(p, r, __ldg(pos + j), fx, fy); });
        body(k, r, c, true, make_real2(fx, fy), InCellExact{p, c}, pv);
      }
    }
    cs += nc;
  }
}

// =========================================================================================================
// Pipelined tile-block kernels (the hot passes of chunked runs: k_newton_p, k_self_propelled_p).
//
// Same decomposition and the same per-particle arithmetic as for_each_block_particle above (so results are bit-identical
// to it), but the staging is taken off the compute warps' critical path — profiles/r01_ncu_newton_summary.md: 46 % of the
// warp-stall samples of the round-1 kernel sat in chunk_stage (three CTA barriers, dependent tstart loads, LDGSTS landing):
//   * PERSISTENT CTAs (3 per SM) fetch tile blocks from a device-side work counter (zeroed by k_step_begin);
//   * ONE PRODUCER WARP per CTA stages chunk k+1 into the second of two shared-memory buffers while the EIGHT CONSUMER
//     WARPS run the pair loops of chunk k: column descriptors live in the producer's registers (lane = column), the
//     prefix sums are warp scans, and every contiguous run of positions ([cell row above | tile | cell row below] of each
//     staged column) is ONE bulk-async copy (cp.async.bulk.shared::cluster.global, UBLKCP) that completes on the
//     buffer's `full` mbarrier — no per-thread address arithmetic, no register staging, no CTA barrier anywhere;
//   * consumers wait on `full` (mbarrier parity), walk their particles, and release the buffer through `empty`; a warp
//     that finishes its share of a chunk early simply starts on the next one.
// Float32 build: a float2 run is only 8-byte aligned, below the 16 bytes bulk copies need, so there the producer warp
// copies the runs itself (plain loads / stores) before it arrives on `full`.
// =========================================================================================================
// warps per CTA: CW consumer warps + NP producer warps (template parameters; see pipe_cfg()).  7 + 1 = 256 threads, 80
// registers, no spills, 21 consumer and 3 producer warps per SM (default, fastest measured); 8 + 1: 24 consumer warps at
// 72 registers with a few spills; a second producer warp halves the consumers' waits but costs more issue slots than it wins.
constexpr int PIPE_CTAS_PER_SM = 3;
constexpr int PG_MAX = 30;                           // own columns per chunk (+ two side columns: one producer lane each)
constexpr int PSPOS_CAP = 1280;                      // staged positions per chunk
constexpr int POWN_CAP = 1024;                       // own particles per chunk
constexpr bool PIPE_BULK = sizeof(real2) == 16;      // bulk-async copies need 16-byte aligned runs (Float64 build)

struct PChunk {
  int state;        // 1: a staged chunk, 0: no more work for this CTA
  int nc, use_mi, nown, ok, tr, cs;
  int src_t1, lt1;  // !ok: the one column that does not fit the staging area (walked in global memory)
  int next;         // consumers: first own particle of the next round of 32 (atomicAdd; reset by the producer)
  int pad_[2];
  // per staged column j.  Own particle q (0 <= q < nown, counted over the own columns 1 .. nc in order) of column j:
  //   staged index self = q + col[j].x, global slot k = self + col[j].y, its cell = s_cell[q + col[j].z],
  //   cell row inside the tile (1-based) lr = cell - col[j].w
  int4 col[PG_MAX + 2];
  int oend[32];     // oend[j]: end of the own-particle range of column j (oend[0] = 0; j > nc: INT_MAX)
  // edges of the cells as axis_in_cell() compares them (lo <= t < hi): staged column j / cell row lr-1 of the tile;
  // bulk copies of the handle's edge tables (DevArrays.edge_x / edge_y)
  double2 xb[PG_MAX + 2];
  double2 yb[MAVI_TR];
  // ext[j][k]: staged index where cell row k-1 of staged column j starts — k = 0 is the cell row ABOVE the tile, k = 1 ..
  // 32 the tile's own rows, then the row below and, past it, the end of the column.  The window of a particle in cell row
  // r (0-based) is [ext[j][r], ext[j][r + 3]) = rows r-1 .. r+1.  Row stride 37 words: the producer's lanes (one per
  // column) and the consumers' lanes (consecutive rows of a column) both access it without bank conflicts.
  int ext[PG_MAX + 2][MAVI_TR + 5];
  __device__ __forceinline__ int2 win(int j, int r) const { return make_int2(ext[j][r], ext[j][r + 3]); }
};
constexpr int PCH_BYTES = (sizeof(PChunk) + 15) / 16 * 16;
// staged cells of the own particles: every own column's run of cell[] is copied from the 16-byte boundary below its first
// slot to the one above its last (<= 6 extra entries), to a 16-byte aligned place -> 12 entries of slack per column
constexpr int PCELL_CAP = (POWN_CAP + 4 + 12 * PG_MAX + 3) / 4 * 4;
constexpr int PBUF_BYTES = PCH_BYTES + PSPOS_CAP * (int)sizeof(real2) + PCELL_CAP * (int)sizeof(int);
constexpr int PIPE_SMEM = 64 + 2 * PBUF_BYTES;   // [4 mbarriers | buffer 0 | buffer 1]

__device__ __forceinline__ unsigned int smem_u32(const void *ptr) { return (unsigned int)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned int bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.release.cta.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait suspends the warp for a while, but on this part it comes back every few hundred cycles: in the capture of
// profiles/r02_ncu_newton_summary.md (r2n) the retry loop was 20 % of ALL issued instructions, taken from the warps the
// waiters wait for.  A failed try is therefore followed by a short sleep (a hand-off costs at most that much; a chunk
// takes ~15 us).
#ifndef MAVI_WAIT_SLEEP_NS
#define MAVI_WAIT_SLEEP_NS 96
#endif
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned int parity) {
  asm volatile(
      "{\n\t.reg .pred ok;\n"
      "W%=:\n\t"
      "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 ok, [%0], %1, %2;\n\t"
      "@ok bra D%=;\n\t"
      "nanosleep.u32 %3;\n\t"
      "bra W%=;\n"
      "D%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(100000u), "r"((unsigned int)MAVI_WAIT_SLEEP_NS)
      : "memory");
}
// one contiguous run global -> shared, completing `bytes` on the mbarrier (UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned int bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// PRODUCER WARP: stage the chunk of own columns starting at local column cs (at most `rem` columns) of tile row tr into
// one buffer and arrive on its `full` barrier.  Returns the number of own columns taken.  Same layout as chunk_stage.
// Lane j owns staged column j.  The warp is a serial resource (one chunk of the CTA per call), so it does as little as
// possible: one round trip for the column's tstart row, a warp scan, bulk-async copies for everything that is a contiguous
// run in global memory (positions, cells, cell edges), and the table of row starts straight from its registers.  The
// consumers find a particle's column / cell row themselves (PChunk::oend, col, s_cell).
template <bool PER>
__device__ __forceinline__ int pipe_stage(const DevParams &p, const int *__restrict__ tstart, const real2 *__restrict__ pos,
                                          const int *__restrict__ cell, const double2 *__restrict__ edge_x,
                                          const double2 *__restrict__ edge_y, int tr, int cs, int rem, bool exact_minimg,
                                          PChunk *ck, real2 *s_pos, int *s_cell, unsigned long long *full_bar) {
  const int lane = threadIdx.x & 31;
  const int R = p.num_rows, Cn = p.num_cols;
  const int r0 = tr * MAVI_TR;
  const int rows = min(MAVI_TR, R - r0);
  const int ncand = min(rem, PG_MAX);
  // ---- column descriptors: lane j describes staged column j (0 .. ncand+1); independent loads
  const int j = lane;
  const bool in = j < ncand + 2;
  int la = 0, lt = 0, lb = 0, sa = 0, st = 0, sb = 0;
  bool wrapped = false;
  int v[MAVI_TR + 1];  // the column's tstart row
#pragma unroll
  for (int r = 0; r <= MAVI_TR; r++) v[r] = 0;
  if (in) {
    int c = cs - 1 + j;
    bool exists = true;
    if (c < 0) { if (p.wrap_cols) { c = Cn - 1; wrapped = true; } else exists = false; }
    else if (c >= Cn) { if (p.wrap_cols) { c = 0; wrapped = true; } else exists = false; }
    if (exists && p.slab && ((c == 0 && p.seam_left) || (c == Cn - 1 && p.seam_right))) wrapped = true;
    int ra = r0 - 1, rb = r0 + MAVI_TR;
    bool has_a = exists, has_b = exists;
    if (ra < 0) { if (p.wrap_rows) { ra = R - 1; wrapped = wrapped || exists; } else has_a = false; }
    if (rb >= R) { if (p.wrap_rows) { rb = 0; wrapped = wrapped || exists; } else has_b = false; }
    if (exists) {
      const int *tt = tstart + (size_t)(c * p.tpc + tr) * (MAVI_TR + 1);
      const int qa = has_a ? tq_of(p, c, ra) : 0, qb = has_b ? tq_of(p, c, rb) : 0;
      // the whole tstart row of the column in ONE round trip (33 + 4 independent loads per lane)
#pragma unroll
      for (int r = 0; r <= MAVI_TR; r++) v[r] = __ldg(tt + r);
      const int a0 = has_a ? __ldg(tstart + qa) : 0, a1 = has_a ? __ldg(tstart + qa + 1) : 0;
      const int b0 = has_b ? __ldg(tstart + qb) : 0, b1 = has_b ? __ldg(tstart + qb + 1) : 0;
      st = v[0];
      lt = v[MAVI_TR] - st; sa = a0; la = a1 - a0; sb = b0; lb = b1 - b0;
    }
  }
  // ---- prefix sums over the lanes, greedy number of own columns that fit
  const int sz = la + lt + lb;
  const int own = (in && j >= 1 && j <= ncand) ? lt : 0;
  int incl = sz, oincl = own;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int vv = __shfl_up_sync(0xffffffffu, incl, o), u = __shfl_up_sync(0xffffffffu, oincl, o);
    if (lane >= o) { incl += vv; oincl += u; }
  }
  const int incl_next = __shfl_down_sync(0xffffffffu, incl, 1);  // staged total if this lane were the last own column
  const bool fits = j >= 1 && j <= ncand && incl_next <= PSPOS_CAP && oincl <= POWN_CAP;
  const unsigned int bal = __ballot_sync(0xffffffffu, fits);
  const int nc = bal ? 31 - __clz(bal) : 1;  // fits is monotone in j
  const unsigned int wbal = __ballot_sync(0xffffffffu, in && j <= nc + 1 && wrapped);
  const int off = incl - sz, ownoff = oincl - own;
  const bool staged = j <= nc + 1, is_own = j >= 1 && j <= nc;
  // cells of the own particles: the run [st, st + lt) of cell[], widened to 16-byte boundaries
  const int cg0 = st & ~3, cg1 = (st + lt + 3) & ~3;
  const int coff = ((ownoff + 3) & ~3) + 12 * (j - 1);
  // bytes that arrive through bulk copies
  unsigned int tx = 0;
  if (bal && staged) {
    if (PIPE_BULK) tx += (unsigned int)lt * (unsigned int)sizeof(real2);
    if (is_own && lt) tx += (unsigned int)(cg1 - cg0) * 4u;
    if (lane == 0) tx += (unsigned int)(nc + 2 + MAVI_TR) * (unsigned int)sizeof(double2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tx += __shfl_xor_sync(0xffffffffu, tx, o);
  const int nown = __shfl_sync(0xffffffffu, oincl, nc);
  const int st1 = __shfl_sync(0xffffffffu, st, 1), lt1 = __shfl_sync(0xffffffffu, lt, 1);
  if (lane == 0) {
    ck->state = 1;
    ck->nc = nc;
    ck->ok = bal ? 1 : 0;
    ck->use_mi = (PER && (exact_minimg || !p.fast_interior || wbal)) ? 1 : 0;
    ck->nown = nown;
    ck->tr = tr;
    ck->cs = cs;
    ck->src_t1 = st1;
    ck->lt1 = lt1;
    ck->next = 0;
    if (tx) mbar_expect_tx(full_bar, tx);
  }
  __syncwarp();
  if (bal) {
    const int base = off + la;             // staged index of the first particle of the tile
    const int end = off + la + lt + lb;    // end of the staged column
    real2 ea[2], eb[2];                    // the cell row above / below the tile: a couple of positions each
    if (staged) {
      // ---- the long-latency part first.  The tile run of the column ([first .. last cell row of the tile], ~40 positions)
      // is ONE bulk-async copy that completes on the buffer's `full` barrier, and so are its cells; the cell row above and
      // below are loaded by the lane itself (three bulk copies per column cost more issue slots than they save)
      if (PIPE_BULK && lt) bulk_g2s(s_pos + base, pos + st, (unsigned int)lt * (unsigned int)sizeof(real2), full_bar);
      if (is_own && lt) bulk_g2s(s_cell + coff, cell + cg0, (unsigned int)(cg1 - cg0) * 4u, full_bar);
      if (lane == 0) {
        bulk_g2s(ck->xb, edge_x + cs, (unsigned int)(nc + 2) * (unsigned int)sizeof(double2), full_bar);  // entry = column + 1
        bulk_g2s(ck->yb, edge_y + r0, (unsigned int)MAVI_TR * (unsigned int)sizeof(double2), full_bar);
      }
#pragma unroll
      for (int i = 0; i < 2; i++) {
        if (i < la) ea[i] = __ldg(pos + sa + i);
        if (i < lb) eb[i] = __ldg(pos + sb + i);
      }
    }
    // ---- per-column words and the row-start table, straight from the registers
    ck->oend[j] = (j == 0) ? 0 : (is_own ? oincl : 0x7fffffff);
    if (staged) {
      ck->col[j] = make_int4(base - ownoff, st - base, coff + (st - cg0) - ownoff, (cs - 1 + j) * R + r0 - 1);
      int *ext = ck->ext[j];
      const int shift = base - st;  // global slot -> staged index (tile rows)
      ext[0] = off;
      if (rows == MAVI_TR) {
#pragma unroll
        for (int r = 0; r <= MAVI_TR; r++) ext[1 + r] = v[r] + shift;
      } else {  // last tile row of the grid: rows beyond it are empty and the cell row "below" is the wrapped one
#pragma unroll
        for (int r = 0; r <= MAVI_TR; r++) ext[1 + r] = r > rows ? end : v[r] + shift;
      }
      ext[MAVI_TR + 2] = end;
      // ---- positions the lane copies itself
#pragma unroll
      for (int i = 0; i < 2; i++) {
        if (i < la) s_pos[off + i] = ea[i];
        if (i < lb) s_pos[base + lt + i] = eb[i];
      }
      for (int i = 2; i < la; i++) s_pos[off + i] = __ldg(pos + sa + i);
      for (int i = 2; i < lb; i++) s_pos[base + lt + i] = __ldg(pos + sb + i);
      if (!PIPE_BULK)  // Float32 build: a float2 run is only 8-byte aligned
        for (int i = 0; i < lt; i++) s_pos[base + i] = __ldg(pos + st + i);
    }
  }
  __syncwarp();  // every lane's descriptor / table / position stores are ordered before the arrival below
  if (lane == 0) mbar_arrive(full_bar);
  return nc;
}

// Drives a pipelined force kernel: pre(k) before and body(k, r, cell, active, F) after the pair force F of every
// particle slot (F = 0 for the inactive tail).  work: device counter of the next tile block (see FLAG_WORK*).
template <int DYN, bool PER, int PIPE_CW, int NPROD, typename Pre, typename Body>
__device__ __forceinline__ void pipe_for_each_particle(const DevParams &p, const int *__restrict__ tstart,
                                                       const real2 *__restrict__ pos, const int *__restrict__ cell,
                                                       const double2 *__restrict__ edge_x, const double2 *__restrict__ edge_y,
                                                       bool exact_minimg, int *__restrict__ work, Pre &&pre, Body &&body) {
  constexpr int PIPE_CT = PIPE_CW * 32;  // consumer threads; the producers are the LAST warp(s) of the CTA
  if (p.pipe_reserved > 0) {  // interior launch of the slab step: leave the reserved SMs to the side stream (grid_pipe())
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (smid < (unsigned int)p.pipe_reserved) return;
  }
  extern __shared__ __align__(16) unsigned char dsm[];
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(dsm);  // [0,1] full, [2,3] empty
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int per_row = blk_items_per_row(p);
  const int nitems = per_row * p.tpc;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], PIPE_CW);
    mbar_init(&bars[3], PIPE_CW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the async proxy (bulk copies) sees the initialised barriers
  }
  __syncthreads();  // the only CTA barrier of the kernel
  if (w >= PIPE_CW) {
    // ================================ producer warp(s) ================================
    // With NPROD = 2 each producer owns ONE of the two buffers and the consumers alternate between them, so two chunks are
    // staged concurrently (A/B switch MAVI_PIPE_CFG; one producer is enough since the staging became cheap).
    const int pw = w - PIPE_CW;
    int k = 0;  // chunks staged by this producer
    for (int taken = 0; p.pipe_items == 0 || taken < p.pipe_items; taken++) {
      int item = 0;
      if (lane == 0) item = atomicAdd(work, 1);
      item = __shfl_sync(0xffffffffu, item, 0);
      if (item >= nitems) break;
      const int tr = item / per_row;
      int c_begin, c_end;
      blk_columns(p, item - tr * per_row, c_begin, c_end);
      for (int cs = c_begin; cs < c_end; k++) {
        const int b = NPROD == 2 ? pw : (k & 1);
        const int u = NPROD == 2 ? k : (k >> 1);  // how often this buffer has been filled before
        if (u >= 1) mbar_wait(&bars[2 + b], (unsigned int)((u - 1) & 1));  // the consumers released its previous contents
        unsigned char *buf = dsm + 64 + (size_t)b * PBUF_BYTES;
        cs += pipe_stage<PER>(p, tstart, pos, cell, edge_x, edge_y, tr, cs, c_end - cs, exact_minimg,
                              reinterpret_cast<PChunk *>(buf), reinterpret_cast<real2 *>(buf + PCH_BYTES),
                              reinterpret_cast<int *>(buf + PCH_BYTES + PSPOS_CAP * sizeof(real2)), &bars[b]);
      }
    }
    // end marker(s): one per buffer this producer feeds
    for (int e = 0; e < (NPROD == 2 ? 1 : 2); e++, k++) {
      const int b = NPROD == 2 ? pw : (k & 1);
      const int u = NPROD == 2 ? k : (k >> 1);
      if (u >= 1) mbar_wait(&bars[2 + b], (unsigned int)((u - 1) & 1));
      if (lane == 0) {
        reinterpret_cast<PChunk *>(dsm + 64 + (size_t)b * PBUF_BYTES)->state = 0;
        mbar_arrive(&bars[b]);
      }
    }
    return;
  }
  // ================================ consumer warps ================================
  if (p.blk_mode != 2) {  // inactive tail: no pair forces
    const int ntail = p.n - p.n_active;
    for (int i = (int)blockIdx.x * PIPE_CT + (int)threadIdx.x; i < ntail; i += (int)gridDim.x * PIPE_CT) {
      const int k = p.tail_base + i;
      auto pv = pre(k);
      body(k, pos[k], 0, false, make_real2(0.0, 0.0), InCellNone{}, pv);
    }
  }
  // The two buffers are consumed alternately.  Per warp: the barrier phase of each buffer (bits 0, 1) and whether its end
  // marker was seen (bits 2, 3) — kept in shared memory: as registers the compiler spilled them to LOCAL memory, and the
  // reload at every chunk boundary was 3.4 % of the stall samples (capture r2u of profiles/r02_ncu_newton_summary.md)
  __shared__ int s_wstate[32];
  volatile int *wst = s_wstate + w;  // written by lane 0 only, read by the warp after a __syncwarp()
  if (lane == 0) *wst = 0;
  __syncwarp();
  for (int b = 0;; b ^= 1) {
    const int st = *wst;
    if ((st >> 2) == 3) break;
    if (st & (4 << b)) continue;
    unsigned char *buf = dsm + 64 + (size_t)b * PBUF_BYTES;
    PChunk *ck = reinterpret_cast<PChunk *>(buf);
    const real2 *s_pos = reinterpret_cast<const real2 *>(buf + PCH_BYTES);
    const int *s_cell = reinterpret_cast<const int *>(buf + PCH_BYTES + PSPOS_CAP * sizeof(real2));
    mbar_wait(&bars[b], (unsigned int)((st >> b) & 1));
    if (!ck->state) {
      __syncwarp();
      if (lane == 0) *wst = st | (4 << b);
      __syncwarp();
      continue;
    }
    if (ck->ok) {
      const int nown = ck->nown;
      const bool mi = PER && ck->use_mi;
      const int oe = ck->oend[lane];
      // Rounds of 32 consecutive own particles are handed out dynamically: with a fixed round-robin the same warps get the
      // extra round of every chunk, and with only two buffers the others cannot run ahead — in the capture r2p of
      // profiles/r02_ncu_newton_summary.md the consumers slept ~40 % of the time on `full` although the producer was idle
      for (;;) {
        int q0 = 0;
        if (lane == 0) q0 = atomicAdd(&ck->next, 32);
        q0 = __shfl_sync(0xffffffffu, q0, 0);
        if (q0 >= nown) break;
        // column of the warp's first particle (oend is non-decreasing), then a lane-local walk: 32 consecutive own
        // particles rarely span more than two columns
        int jj = __popc(__ballot_sync(0xffffffffu, oe <= q0));
        const int q = q0 + lane;
        if (q < nown) {
          while (q >= ck->oend[jj]) jj++;
          const int4 cd = ck->col[jj];
          const int self = q + cd.x, k = self + cd.y;
          auto pv = pre(k);
          const int cc = s_cell[q + cd.z];
          const int lr = cc - cd.w;
          const real2 r = s_pos[self];
          real fx = 0.0, fy = 0.0;
          if (mi) chunk_walk<DYN, true>(p, ck, s_pos, jj, lr, self, r, fx, fy);
          else chunk_walk<DYN, false>(p, ck, s_pos, jj, lr, self, r, fx, fy);
          const double2 xe = ck->xb[jj], ye = ck->yb[lr - 1];
          body(k, r, cc, true, make_real2(fx, fy), InCellTab{p, xe.x, xe.y, ye.x, ye.y}, pv);
        }
      }
    } else {
      // a single column too dense for the staging area: per-thread walk over the global arrays
      const int b0 = ck->src_t1, e0 = b0 + ck->lt1;
      for (int k = b0 + (int)threadIdx.x; k < e0; k += PIPE_CT) {
        auto pv = pre(k);
        const real2 r = pos[k];
        const int c = cell[k];
        real fx = 0.0, fy = 0.0;
        for_each_neighbor(p, tstart, c, k, [&](int jn) { accumulate_pair<DYN, PER>(p, r, __ldg(pos + jn), fx, fy); });
        body(k, r, c, true, make_real2(fx, fy), InCellExact{p, c}, pv);
      }
    }
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&bars[2 + b]);  // this warp is done with the buffer
      *wst = st ^ (1 << b);
    }
    __syncwarp();
  }
}

// more than the 48 KB of dynamic shared memory a kernel gets by default: opt in once per (kernel instantiation, DEVICE) — the
// attribute belongs to the device's context, and a single-process multi-GPU handle launches from one thread per device
#define MAVI_OPT_IN_SMEM(kernel, bytes)                                                                  \
  do {                                                                                                   \
    static bool done_[64] = {};                                                                          \
    int dev_ = 0;                                                                                        \
    cudaGetDevice(&dev_);                                                                                \
    if (dev_ >= 0 && dev_ < 64 && !done_[dev_]) {                                                        \
      cudaFuncSetAttribute((const void *)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (bytes)); \
      done_[dev_] = true;                                                                                \
    }                                                                                                    \
  } while (0)

// (consumer warps) * 10 + (producer warps) per CTA of the pipelined kernels: 71 (default), 81, 72, 82 — MAVI_PIPE_CFG is the
// A/B switch of the measurements in profiles/r02_ncu_newton_summary.md (LJ 16 M: 71 0.551 ms, 81 0.562, 82 0.562, 72 0.598)
static inline int pipe_cfg() {
  static const int cfg = [] {
    const char *e = getenv("MAVI_PIPE_CFG");
    const int v = e ? atoi(e) : 71;
    return (v == 81 || v == 72 || v == 82) ? v : 71;
  }();
  return cfg;
}

// Float32 build: without bulk copies (8-byte aligned runs) the producer warp moves every position through its registers and
// the pipelined kernels LOSE to the cp.async kernels (LJ 16 M, gpurun_out r2n: 0.684 vs 0.501 ms/step), so the Float32 build
// keeps the latter; MAVI_F32_PIPELINED=1 is the A/B switch.
static inline bool pipe_default() {
  static const bool v = PIPE_BULK || getenv("MAVI_F32_PIPELINED") != nullptr;
  return v;
}

// Grids of the pipelined kernels.  Single launches (blk_mode 0) are PERSISTENT: 3 CTAs per SM until the work counter runs out.
// The two-stream slab step needs room on the device WHILE its interior launch runs — for the boundary launch and above all
// for NCCL's send/recv kernel (96 registers x ~544 threads: it only fits an SM that holds none of these CTAs) — so there a
// CTA exits after `pipe_items` tile blocks (MAVI_SLAB_PIPE_ITEMS, default 8, ~100 us; measured at 2 GPUs: 5 -> 0.615, 8 -> 0.601,
// 12 -> 0.622, 16 -> 0.637 ms/step, cp.async kernels 0.696): the block scheduler drains an SM for
// the high-priority side stream within one CTA lifetime.  With persistent CTAs the NCCL kernel of the migration exchange
// started only when the interior launch had finished (2 GPUs: 0.86 ms/step, profiles/r02_slab_timeline.md).
static inline int slab_pipe_items() {
  static const int v = [] {
    const char *e = getenv("MAVI_SLAB_PIPE_ITEMS");
    const int k = e ? atoi(e) : 8;
    return k < 0 ? 0 : k;
  }();
  return v;
}

// ... or, better (default), the interior launch stays persistent and LEAVES `reserved` SMs ALONE: it launches a few CTAs more
// than the device holds, and every CTA that finds itself on an SM with %smid < reserved exits at once.  Those SMs then stay
// empty for the whole pass (persistent CTAs never move), and the boundary launch (reserved * 3 persistent CTAs over the
// narrow edge blocks, MAVI_EDGE_COLS), NCCL's kernels (<= 2 channels, slab.cu) and the small kernels of the side stream run
// there without waiting for anything to drain.  MAVI_SLAB_RESERVED_SMS (default 4; 0 = CTAs of bounded lifetime instead).
// Measured at 2 GPUs (LJ, 16 M per GPU, ms/step): bounded lifetime 0.592; reserved 2 / 3 / 4 SMs 0.633 / 0.574 / 0.564
// (one GPU, no exchange: 0.538).  Checked once per device that the SM ids 0 .. reserved-1 exist.
__global__ void k_smid_probe(int *__restrict__ seen) {
  unsigned int smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  if (threadIdx.x == 0 && smid < 64) seen[smid] = 1;
}
static inline int slab_reserved_sms(cudaStream_t stream) {
  static int cache[64];
  static bool done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 0;
  if (!done[dev]) {
    const char *e = getenv("MAVI_SLAB_RESERVED_SMS");
    int want = e ? atoi(e) : 4;
    if (want < 0) want = 0;
    if (want > 32) want = 32;
    int ok = 0;
    if (want > 0) {
      int *seen = nullptr;
      int host[64] = {};
      if (cudaMalloc((void **)&seen, 64 * sizeof(int)) == cudaSuccess) {
        cudaMemsetAsync(seen, 0, 64 * sizeof(int), stream);
        k_smid_probe<<<148 * 16, 64, 0, stream>>>(seen);  // far more CTAs than SMs: every SM gets some
        cudaMemcpyAsync(host, seen, 64 * sizeof(int), cudaMemcpyDeviceToHost, stream);
        cudaStreamSynchronize(stream);
        cudaFree(seen);
        ok = 1;
        for (int i = 0; i < want; i++) ok = ok && host[i];
      }
    }
    cache[dev] = ok ? want : 0;
    done[dev] = true;
  }
  return cache[dev];
}
constexpr int PIPE_EXTRA_CTAS = 128;  // interior launch with reserved SMs: CTAs beyond one device-full (see above)

static inline int grid_pipe(DevParams &p, cudaStream_t stream) {
  const int items = blk_items_per_row(p) * p.tpc;
  static const int items_all = getenv("MAVI_PIPE_ITEMS_ALL") ? atoi(getenv("MAVI_PIPE_ITEMS_ALL")) : 0;  // diagnosis: bounded CTAs everywhere
  const int want = 148 * PIPE_CTAS_PER_SM;
  p.pipe_reserved = 0;
  const int reserved = p.blk_mode == 0 ? 0 : slab_reserved_sms(stream);
  if (reserved > 0) {
    p.pipe_items = 0;
    if (p.blk_mode == 1) {
      p.pipe_reserved = reserved;
      return want + PIPE_EXTRA_CTAS;
    }
    const int g = reserved * PIPE_CTAS_PER_SM;
    return items < g ? (items > 0 ? items : 1) : g;
  }
  p.pipe_items = p.blk_mode == 0 ? items_all : slab_pipe_items();
  if (p.pipe_items > 0) {
    // one device-full of CTAs more than items / pipe_items: the last wave is then as wide as the others and the work counter
    // balances it (with exactly items / pipe_items CTAs the last, partial wave ran at a fraction of the device for a whole
    // CTA lifetime: 16 items per CTA were SLOWER than 8); CTAs that find the counter exhausted exit at once
    const int g = (items + p.pipe_items - 1) / p.pipe_items + want;
    return items > 0 ? (g < items ? g : items) : 1;
  }
  return items < want ? (items > 0 ? items : 1) : want;
}

static inline int grid2(const DevParams &p) {
  return blk_items_per_row(p) * p.tpc + (p.blk_mode == 2 ? 0 : nblk(p.n - p.n_active));
}

template <int DYN, bool PER>
__global__ void __launch_bounds__(TPB) k_force_only2(const __grid_constant__ DevParams p, const int *__restrict__ tstart,
                              const int *__restrict__ cell, const real2 *__restrict__ pos,
                              real2 *__restrict__ force, int with_walls) {
  for_each_block_particle<DYN, PER>(p, tstart, pos, cell, false, [](int) { return 0; },
    [&](int k, real2 r, int, bool active, real2 F, auto, auto) {
      if (active && with_walls && p.has_force_walls) wall_forces(p, r.x, r.y, F.x, F.y);
      force[k] = F;
    });
}

template <int DYN, bool PER>
__global__ void __launch_bounds__(TPB) k_newton_a2(const __grid_constant__ DevParams p, const int *__restrict__ tstart,
                            const int *__restrict__ cell, const real2 *__restrict__ pos_in,
                            const real2 *__restrict__ vel, real2 *__restrict__ pos_out, real2 *__restrict__ f1,
                            int *__restrict__ flags) {
  if (!flags[FLAG_RAN]) return;
  for_each_block_particle<DYN, PER>(p, tstart, pos_in, cell, false, [&](int k) { prefetch_l1(vel + k); return 0; },
    [&](int k, real2 r, int, bool active, real2 F, auto, auto) {
      if (active && p.has_force_walls) wall_forces(p, r.x, r.y, F.x, F.y);
      bool big;
      pos_out[k] = verlet_drift(p, r, vel[k], F, big);
      if (PER && big) flags[FLAG_BIGMOVE] = 1;
      f1[k] = F;
    });
}

// pre(k) / body(k, r, cell, active, F) of the second Newton pass
// vel / F1 are loaded into registers before the pair loop.  A prefetch to L1 (MAVI_LATE_LOAD, build-time A/B) leaves the loads
// after the loop exposed: 16 % of the stall samples of capture r2q in profiles/r02_ncu_newton_summary.md, 0.544 -> 0.518 ms/step
struct VelF1 { real2 v, f; };
#define MAVI_NB_PRE_EARLY [&](int k) { return VelF1{vel[k], f1[k]}; }
#define MAVI_NB_GET_EARLY real2 v = pv.v; const real2 Fo = pv.f;
#define MAVI_NB_PRE_LATE [&](int k) { prefetch_l1(vel + k); prefetch_l1(f1 + k); return 0; }
#define MAVI_NB_GET_LATE real2 v = vel[k]; const real2 Fo = f1[k];
// the cp.async kernels (64 registers, 4 CTAs per SM) keep the prefetch: no room for eight more live registers
#define MAVI_NEWTON_B_LAMBDAS_LEGACY MAVI_NEWTON_B_LAMBDAS_(MAVI_NB_PRE_LATE, MAVI_NB_GET_LATE)
#ifndef MAVI_LATE_LOAD
#define MAVI_NEWTON_B_LAMBDAS MAVI_NEWTON_B_LAMBDAS_(MAVI_NB_PRE_EARLY, MAVI_NB_GET_EARLY)
#else
#define MAVI_NEWTON_B_LAMBDAS MAVI_NEWTON_B_LAMBDAS_(MAVI_NB_PRE_LATE, MAVI_NB_GET_LATE)
#endif
#define MAVI_NEWTON_B_LAMBDAS_(PRE_, GET_) \
    PRE_, \
    [&](int k, real2 r, int c, bool active, real2 F, auto in_cell, auto pv) { \
      GET_ \
      v.x = v.x + p.hdt * (F.x + Fo.x); \
      v.y = v.y + p.hdt * (F.y + Fo.y); \
      if (active) { \
        const real x0 = r.x, y0 = r.y; \
        apply_walls<true>(p, r.x, r.y, v.x, v.y, p.particle_radius); \
        const bool fixed = (r.x != x0 || r.y != y0); \
        if (fixed) { \
          int m = atomicAdd(&ms.flags[FLAG_NFIX], 1); \
          fix_idx[m] = k; \
          fix_pos[m] = r; \
        } \
        note_if_moved(p, ms, k, c, r.x, r.y, fixed, F, [&] { return v; }, in_cell); \
      } \
      vel[k] = v; \
      f2[k] = F; \
      if (CARRY) { \
        real2 Fn = F; \
        if (p.has_force_walls && active) wall_forces(p, r.x, r.y, Fn.x, Fn.y); \
        f1_next[k] = Fn; \
        bool big; \
        pos_next[k] = verlet_drift(p, r, v, Fn, big); \
        if (PER && big) ms.flags[FLAG_BIGMOVE_NEXT] = 1; \
      } \
    }

template <int DYN, bool PER, bool CARRY, int MINB>
__global__ void __launch_bounds__(TPB, MINB) k_newton_b2(const __grid_constant__ DevParams p, const int *__restrict__ tstart,
                            const real2 *__restrict__ pos_in, real2 *__restrict__ vel, const real2 *f1,
                            real2 *f2, real2 *f1_next, real2 *__restrict__ pos_next,
                            int *__restrict__ fix_idx, real2 *__restrict__ fix_pos,
                            const __grid_constant__ MoverSink ms) {
  if (!ms.flags[FLAG_RAN]) return;
  // slab mode, blocks next to a halo column (blk_mode 2): their boundary particles are re-drifted on the side stream
  // without a big-drift report -> always the exact minimum-image path there
  const bool exact = ms.flags[FLAG_BIGMOVE] != 0 || (p.slab && p.blk_mode == 2);
  for_each_block_particle<DYN, PER>(p, tstart, pos_in, ms.cell, exact, MAVI_NEWTON_B_LAMBDAS_LEGACY);
}

// the pipelined version of k_newton_b2 (default); work item counter: FLAG_WORK0 / FLAG_WORK1 (boundary-block launch)
template <int DYN, bool PER, bool CARRY, int CW, int NP>
__global__ void __launch_bounds__((CW + NP) * 32, PIPE_CTAS_PER_SM) k_newton_p(
    const __grid_constant__ DevParams p, const int *__restrict__ tstart, const real2 *__restrict__ pos_in,
    real2 *__restrict__ vel, const real2 *f1, real2 *f2, real2 *f1_next, real2 *__restrict__ pos_next,
    int *__restrict__ fix_idx, real2 *__restrict__ fix_pos, const __grid_constant__ MoverSink ms) {
  if (!ms.flags[FLAG_RAN]) return;
  const bool exact = ms.flags[FLAG_BIGMOVE] != 0 || (p.slab && p.blk_mode == 2);
  pipe_for_each_particle<DYN, PER, CW, NP>(p, tstart, pos_in, ms.cell, ms.edge_x, ms.edge_y, exact, ms.flags + (p.blk_mode == 2 ? FLAG_WORK1 : FLAG_WORK0),
                                       MAVI_NEWTON_B_LAMBDAS);
}

struct AngId { real theta; unsigned int idf; };
template <int DYN, bool PER, int CW, int NP>
__global__ void __launch_bounds__((CW + NP) * 32, PIPE_CTAS_PER_SM) k_self_propelled_p(
    const __grid_constant__ DevParams p, const int *__restrict__ tstart, const unsigned int *__restrict__ idflag,
    const real2 *__restrict__ pos_in, real *__restrict__ ang, real2 *__restrict__ pos_out, real2 *__restrict__ force,
    const real *__restrict__ noise, unsigned long long step, const __grid_constant__ MoverSink ms) {
  if (!ms.flags[FLAG_RAN]) return;
  pipe_for_each_particle<DYN, PER, CW, NP>(p, tstart, pos_in, ms.cell, ms.edge_x, ms.edge_y, false, ms.flags + FLAG_WORK0,
    // angle and id are loaded BEFORE the pair loop (3 registers): a prefetch to L1 left the loads exposed after it
    [&](int k) { return AngId{ang[k], idflag[k]}; },
    [&](int k, real2 r, int c, bool active, real2 F, auto in_cell, auto pv) {
      const unsigned int id = pv.idf & ~MAVI_INACTIVE_BIT;
      if (active && p.has_force_walls) wall_forces(p, r.x, r.y, F.x, F.y);
      force[k] = F;
      if ((int)id < p.n_count) self_propelled_update<DYN>(p, r, F, ang, k, id, noise, step, pv.theta);
      if (active) {
        real vx = 0.0, vy = 0.0;
        apply_walls<false>(p, r.x, r.y, vx, vy, p.particle_radius);
        note_if_moved(p, ms, k, c, r.x, r.y, false, F, [&] { return make_real2(ang[k], 0.0); }, in_cell);
      }
      pos_out[k] = r;
    });
}

template <int DYN, bool PER>
__global__ void __launch_bounds__(TPB) k_self_propelled2(const __grid_constant__ DevParams p, const int *__restrict__ tstart,
                                  const unsigned int *__restrict__ idflag, const real2 *__restrict__ pos_in,
                                  real *__restrict__ ang, real2 *__restrict__ pos_out, real2 *__restrict__ force,
                                  const real *__restrict__ noise, unsigned long long step,
                                  const __grid_constant__ MoverSink ms) {
  if (!ms.flags[FLAG_RAN]) return;
  for_each_block_particle<DYN, PER>(p, tstart, pos_in, ms.cell, false,
    [&](int k) { prefetch_l1(ang + k); prefetch_l1(idflag + k); return 0; },
    [&](int k, real2 r, int c, bool active, real2 F, auto in_cell, auto) {
      const unsigned int id = idflag[k] & ~MAVI_INACTIVE_BIT;
      if (active && p.has_force_walls) wall_forces(p, r.x, r.y, F.x, F.y);
      force[k] = F;
      if ((int)id < p.n_count) self_propelled_update<DYN>(p, r, F, ang, k, id, noise, step, ang[k]);
      if (active) {
        real vx = 0.0, vy = 0.0;
        apply_walls<false>(p, r.x, r.y, vx, vy, p.particle_radius);
        note_if_moved(p, ms, k, c, r.x, r.y, false, F, [&] { return make_real2(ang[k], 0.0); }, in_cell);
      }
      pos_out[k] = r;
    });
}

// dispatch over the periodic flag (the dynamics is switched on by the callers)
#define MAVI_DISPATCH2(DYNV, PERV, CALL) \
  do {                                   \
    if (PERV) { CALL(DYNV, true); } else { CALL(DYNV, false); } \
  } while (0)

// Edges of the cells exactly as axis_in_cell() / still_in_cell() compare them (lo <= t < hi with t = x - grid_bl[0] or
// -y + grid_bl[1] + grid_h): edge_x[lc + 1] for the LOCAL column lc = -1 .. num_cols (the two extra entries are the columns
// reached through the periodic wrap), edge_y[row] for row = 0 .. tpc * MAVI_TR - 1.  Depends on the grid only.
__global__ void k_cell_edges(const __grid_constant__ DevParams p, double2 *__restrict__ edge_x, double2 *__restrict__ edge_y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int Cn = p.num_cols, R = p.num_rows;
  if (i < Cn + 2) {
    int g = i - 1;  // local column -> global column
    if (g < 0) g += Cn;
    else if (g >= Cn) g -= Cn;
    if (p.slab) {
      g = g + p.col_lo - 1;
      if (g < 0) g += p.gcols;
      else if (g >= p.gcols) g -= p.gcols;
    }
    const double lo = g == 0 ? nextafter(-p.cl, 1.0) : __dmul_ru((double)g, p.cl);  // t > -c  <=>  t >= nextafter(-c, +inf)
    const double hi = __dmul_ru((double)((g == p.gcols - 1) ? p.gcols + 1 : g + 1), p.cl);
    edge_x[i] = make_double2(lo, hi);
  }
  if (i < p.tpc * MAVI_TR) {
    const int row = i;
    const double lo = row == 0 ? nextafter(-p.ch, 1.0) : __dmul_ru((double)row, p.ch);
    const double hi = __dmul_ru((double)((row == R - 1) ? R + 1 : row + 1), p.ch);
    edge_y[i] = make_double2(lo, hi);
  }
}

void launch_cell_edges(const LaunchCtx &c, const DevParams &p, const DevArrays &a) {
  if (p.num_cells == 0 || !a.edge_x) return;
  const int n = max(p.num_cols + 2, p.tpc * MAVI_TR);
  MAVI_LAUNCH(c, k_cell_edges, nblk(n), TPB, 0, p, a.edge_x, a.edge_y);
}

static MoverSink mover_sink(const DevArrays &a) {
  return MoverSink{a.cell, a.tile_dirty, a.dirty_list, a.inbox_cnt, a.inbox, a.mv_src, a.flags, nullptr, a.em_send[0], a.em_send[1], a.em_cap, a.idflag, a.edge_x, a.edge_y};
}

void launch_force_only(const LaunchCtx &c, const DevParams &p, const DevArrays &a, bool with_wall_forces) {
  const bool allp = p.num_cells == 0;
  if (!allp) {
#define CALL2(D, P) MAVI_LAUNCH(c, (k_force_only2<D, P>), grid2(p), TPB, PASS2_SMEM, p, a.tstart, a.cell, a.pos[0], a.force, (int)with_wall_forces)
    switch (p.dynamics) {
      case MAVI_DYN_LJ: MAVI_DISPATCH2(MAVI_DYN_LJ, p.periodic, CALL2); break;
      case MAVI_DYN_HARMTRUNC: MAVI_DISPATCH2(MAVI_DYN_HARMTRUNC, p.periodic, CALL2); break;
      case MAVI_DYN_SZABO: MAVI_DISPATCH2(MAVI_DYN_SZABO, p.periodic, CALL2); break;
      case MAVI_DYN_RTP: MAVI_DISPATCH2(MAVI_DYN_RTP, p.periodic, CALL2); break;
    }
#undef CALL2
    return;
  }
#define CALL(D, P) MAVI_LAUNCH(c, (k_force_only<D, P>), nblk(p.n, RPB), TPB, 0, p, a.idflag, a.pos[0], a.force, (int)with_wall_forces)
  switch (p.dynamics) {
    case MAVI_DYN_LJ: MAVI_DISPATCH2(MAVI_DYN_LJ, p.periodic, CALL); break;
    case MAVI_DYN_HARMTRUNC: MAVI_DISPATCH2(MAVI_DYN_HARMTRUNC, p.periodic, CALL); break;
    case MAVI_DYN_SZABO: MAVI_DISPATCH2(MAVI_DYN_SZABO, p.periodic, CALL); break;
    case MAVI_DYN_RTP: MAVI_DISPATCH2(MAVI_DYN_RTP, p.periodic, CALL); break;
  }
#undef CALL
}

void launch_newton_a(const LaunchCtx &c, const DevParams &p, const DevArrays &a) {
  const bool allp = p.num_cells == 0;
  if (!allp) {
#define CALL2(D, P) MAVI_LAUNCH(c, (k_newton_a2<D, P>), grid2(p), TPB, PASS2_SMEM, p, a.tstart, a.cell, a.pos[0], a.vel, a.pos[1], a.force_old, a.flags)
    if (p.dynamics == MAVI_DYN_LJ) MAVI_DISPATCH2(MAVI_DYN_LJ, p.periodic, CALL2);
    else MAVI_DISPATCH2(MAVI_DYN_HARMTRUNC, p.periodic, CALL2);
#undef CALL2
    return;
  }
#define CALL(D, P) MAVI_LAUNCH(c, (k_newton_a<D, P>), nblk(p.n, RPB), TPB, 0, p, a.idflag, a.pos[0], a.vel, a.pos[1], a.force_old, a.flags)
  if (p.dynamics == MAVI_DYN_LJ) MAVI_DISPATCH2(MAVI_DYN_LJ, p.periodic, CALL);
  else MAVI_DISPATCH2(MAVI_DYN_HARMTRUNC, p.periodic, CALL);
#undef CALL
}

void launch_apply_pos_fixes(const LaunchCtx &c, const DevArrays &a) {
  MAVI_LAUNCH(c, k_apply_pos_fixes, 64, TPB, 0, a.flags, a.fix_idx, a.fix_pos, a.pos[1]);
}

// carry: also write the next step's F1 / drift (into force_old / pos[0]) and record the changed cells
void launch_newton_b(const LaunchCtx &c, const DevParams &p_in, const DevArrays &a, bool carry, int blk_mode) {
  DevParams p = p_in;
  p.blk_mode = blk_mode;
  const bool allp = p.num_cells == 0;
  MoverSink ms = mover_sink(a);
  // reads the drifted positions pos[1] (which become the current positions after the sparse wall fix-ups)
#define ARGS p, a.idflag, a.pos[1], a.vel, a.force_old, a.force, a.fix_idx, a.fix_pos, a.flags
#define CALL(D, P) MAVI_LAUNCH(c, (k_newton_b<D, P>), nblk(p.n, RPB), TPB, 0, ARGS)
  if (!allp) {
#define ARGS2 p, a.tstart, a.pos[1], a.vel, a.force_old, a.force, a.force_old, a.pos[0], a.fix_idx, a.fix_pos, ms
#define CALL2(D, P) MAVI_LAUNCH(c, (k_newton_b2<D, P, false, 4>), grid2(p), TPB, PASS2_SMEM, ARGS2)
#define CALL2C(D, P) MAVI_LAUNCH(c, (k_newton_b2<D, P, true, 4>), grid2(p), TPB, PASS2_SMEM, ARGS2)
    // pipelined persistent kernels (default); more than the 48 KB a kernel gets by default: opt in once per instantiation
#define CALLP__(D, P, CARRYV, CWV, NPV)                                                                             \
  do {                                                                                                              \
    MAVI_OPT_IN_SMEM((k_newton_p<D, P, CARRYV, CWV, NPV>), PIPE_SMEM);                                              \
    const int grid_ = grid_pipe(p, c.stream);                                                                                \
    MAVI_LAUNCH(c, (k_newton_p<D, P, CARRYV, CWV, NPV>), grid_, (CWV + NPV) * 32, PIPE_SMEM, ARGS2);                \
  } while (0)
#define CALLP_(D, P, CARRYV)                               \
  do {                                                     \
    switch (pipe_cfg()) {                                  \
      case 81: CALLP__(D, P, CARRYV, 8, 1); break;         \
      case 72: CALLP__(D, P, CARRYV, 7, 2); break;         \
      case 82: CALLP__(D, P, CARRYV, 8, 2); break;         \
      default: CALLP__(D, P, CARRYV, 7, 1); break;         \
    }                                                      \
  } while (0)
#define CALLP(D, P) CALLP_(D, P, false)
#define CALLPC(D, P) CALLP_(D, P, true)
    // The two-stream slab step splits the pass into an interior and a boundary launch that must overlap: the boundary blocks
    // (side stream, high priority) slip into the slots the short-lived CTAs of the non-persistent kernels free all the time,
    // but cannot get onto a device that persistent CTAs occupy for the whole pass (measured at 2 GPUs, profiles/
    // r02_slab_timeline.md: 0.70 ms/step with the non-persistent kernels, 0.86 with persistent ones + 16 reserved slots):
    // the split launches use the pipelined kernels with CTAs of bounded lifetime (grid_pipe).  MAVI_SLAB_LEGACY=1 is the
    // A/B switch back to the cp.async kernels.
    static const bool slab_pipelined = getenv("MAVI_SLAB_LEGACY") == nullptr;
    if (pipe_default() && !(c.flags & MAVI_FLAG_LEGACY_STAGING) && (blk_mode == 0 || slab_pipelined)) {
      if (carry) {
        ms.chg = a.chg;
        if (p.dynamics == MAVI_DYN_LJ) MAVI_DISPATCH2(MAVI_DYN_LJ, p.periodic, CALLPC);
        else MAVI_DISPATCH2(MAVI_DYN_HARMTRUNC, p.periodic, CALLPC);
      } else {
        if (p.dynamics == MAVI_DYN_LJ) MAVI_DISPATCH2(MAVI_DYN_LJ, p.periodic, CALLP);
        else MAVI_DISPATCH2(MAVI_DYN_HARMTRUNC, p.periodic, CALLP);
      }
    } else if (carry) {
      ms.chg = a.chg;
      if (p.dynamics == MAVI_DYN_LJ) MAVI_DISPATCH2(MAVI_DYN_LJ, p.periodic, CALL2C);
      else MAVI_DISPATCH2(MAVI_DYN_HARMTRUNC, p.periodic, CALL2C);
    } else {
      if (p.dynamics == MAVI_DYN_LJ) MAVI_DISPATCH2(MAVI_DYN_LJ, p.periodic, CALL2);
      else MAVI_DISPATCH2(MAVI_DYN_HARMTRUNC, p.periodic, CALL2);
    }
#undef CALL2
#undef CALL2C
#undef CALLP__
#undef CALLP_
#undef CALLP
#undef CALLPC
#undef ARGS2
  } else {
    if (p.dynamics == MAVI_DYN_LJ) MAVI_DISPATCH2(MAVI_DYN_LJ, p.periodic, CALL);
    else MAVI_DISPATCH2(MAVI_DYN_HARMTRUNC, p.periodic, CALL);
  }
#undef CALL
#undef ARGS
  if (blk_mode == 0) launch_apply_pos_fixes(c, a);
}

// force carry, after the swap and the tile repair: pos[0] = current positions, pos[1] = carried drift
void launch_carry_redrift(const LaunchCtx &c, const DevParams &p, const DevArrays &a) {
  MAVI_LAUNCH(c, k_redrift_tiles, 148 * 2, TPB, 0, p, a.flags, a.dirty_list, a.tstart, a.pos[0], a.vel, a.force,
              a.force_old, a.pos[1]);
}

// skip_edge / depth / report_big: see k_recompute_changed / k_recompute_columns
void launch_carry_recompute_list(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int skip_edge) {
#define CALL(D, P) \
  MAVI_LAUNCH(c, (k_recompute_changed<D, P>), 148 * 4, TPB, 0, p, a.flags, a.chg, a.tstart, a.pos[0], a.vel, a.force_old, a.pos[1], skip_edge)
  if (p.dynamics == MAVI_DYN_LJ) MAVI_DISPATCH2(MAVI_DYN_LJ, p.periodic, CALL);
  else MAVI_DISPATCH2(MAVI_DYN_HARMTRUNC, p.periodic, CALL);
#undef CALL
}

void launch_carry_recompute_columns(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int depth, bool report_big) {
  const int nthreads = 2 * depth * p.tpc * p.cap;
#define CALL(D, P) \
  MAVI_LAUNCH(c, (k_recompute_columns<D, P>), nblk(nthreads), TPB, 0, p, a.flags, a.tstart, a.cell, a.pos[0], a.vel, a.force_old, a.pos[1], depth, (int)report_big)
  if (p.dynamics == MAVI_DYN_LJ) MAVI_DISPATCH2(MAVI_DYN_LJ, p.periodic, CALL);
  else MAVI_DISPATCH2(MAVI_DYN_HARMTRUNC, p.periodic, CALL);
#undef CALL
}

void launch_carry_recompute(const LaunchCtx &c, const DevParams &p, const DevArrays &a) {
  launch_carry_recompute_list(c, p, a, p.slab ? 1 : 0);
  if (p.slab) launch_carry_recompute_columns(c, p, a, 1, true);
}

void launch_carry_fixups(const LaunchCtx &c, const DevParams &p, const DevArrays &a) {
  launch_carry_redrift(c, p, a);
  launch_carry_recompute(c, p, a);
}

void launch_self_propelled(const LaunchCtx &c, const DevParams &p, const DevArrays &a, const real *noise,
                           unsigned long long step) {
  const bool allp = p.num_cells == 0;
  const MoverSink ms = mover_sink(a);
  DevParams pp = p;  // grid_pipe() sets pipe_items
  if (!allp) {
#define CALL2(D, P) MAVI_LAUNCH(c, (k_self_propelled2<D, P>), grid2(p), TPB, PASS2_SMEM, p, a.tstart, a.idflag, a.pos[0], a.ang, a.pos[1], a.force, noise, step, ms)
#define CALLP_(D, P, CWV, NPV)                                                                                      \
  do {                                                                                                              \
    MAVI_OPT_IN_SMEM((k_self_propelled_p<D, P, CWV, NPV>), PIPE_SMEM);                                              \
    const int grid_ = grid_pipe(pp, c.stream);                                                                                \
    MAVI_LAUNCH(c, (k_self_propelled_p<D, P, CWV, NPV>), grid_, (CWV + NPV) * 32, PIPE_SMEM, pp, a.tstart, a.idflag, a.pos[0], a.ang, a.pos[1], a.force, noise, step, ms); \
  } while (0)
#define CALLP(D, P)                                  \
  do {                                               \
    switch (pipe_cfg()) {                            \
      case 81: CALLP_(D, P, 8, 1); break;            \
      case 72: CALLP_(D, P, 7, 2); break;            \
      case 82: CALLP_(D, P, 8, 2); break;            \
      default: CALLP_(D, P, 7, 1); break;            \
    }                                                \
  } while (0)
    if (pipe_default() && !(c.flags & MAVI_FLAG_LEGACY_STAGING)) {
      if (p.dynamics == MAVI_DYN_SZABO) MAVI_DISPATCH2(MAVI_DYN_SZABO, p.periodic, CALLP);
      else MAVI_DISPATCH2(MAVI_DYN_RTP, p.periodic, CALLP);
    } else {
      if (p.dynamics == MAVI_DYN_SZABO) MAVI_DISPATCH2(MAVI_DYN_SZABO, p.periodic, CALL2);
      else MAVI_DISPATCH2(MAVI_DYN_RTP, p.periodic, CALL2);
    }
#undef CALLP
#undef CALLP_
#undef CALL2
    MAVI_LAUNCH(c, k_apply_pos_fixes, 1, 32, 0, a.flags, a.fix_idx, a.fix_pos, a.pos[1]);  // no fix-ups here: step counter only
    return;
  }
#define CALL(D, P) MAVI_LAUNCH(c, (k_self_propelled<D, P>), nblk(p.n, RPB), TPB, 0, p, a.idflag, a.pos[0], a.ang, a.pos[1], a.force, noise, step, a.flags)
  if (p.dynamics == MAVI_DYN_SZABO) MAVI_DISPATCH2(MAVI_DYN_SZABO, p.periodic, CALL);
  else MAVI_DISPATCH2(MAVI_DYN_RTP, p.periodic, CALL);
#undef CALL
  MAVI_LAUNCH(c, k_apply_pos_fixes, 1, 32, 0, a.flags, a.fix_idx, a.fix_pos, a.pos[1]);  // no fix-ups here: step counter only
}

// =========================================================================================================
// Quantities (src/quantities.jl) as deterministic two-stage block reductions.
// =========================================================================================================
constexpr int RED_TPB = 256, RED_MAX_BLOCKS = 1184;  // 8 CTAs x 148 SMs

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double ws[RED_TPB / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < RED_TPB / 32) t = ws[threadIdx.x];
  if (threadIdx.x < 32) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;  // valid in thread 0
}

__global__ void k_reduce_final(const double *__restrict__ partials, int nb, double scale, double *__restrict__ out) {
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) s += partials[i];
  s = block_sum(s);
  if (threadIdx.x == 0) *out = s * scale;
}

// kinetic_energy, src/quantities.jl:12-18: sum over ALL slots of |v|^2, /2 (mass 1)
__global__ void k_kinetic(const __grid_constant__ DevParams p, const int *__restrict__ tile_prefix,
                          const int *__restrict__ cta_first, const real2 *__restrict__ vel,
                          double *__restrict__ partials) {
  double s = 0.0;
  for (int base = blockIdx.x * blockDim.x; base < p.n; base += gridDim.x * blockDim.x) {
    int rank = base + threadIdx.x;
    if (rank < p.n) {
      real2 v = vel[slot_of_rank(p, tile_prefix, cta_first, rank)];
      s += (double)v.x * v.x + (double)v.y * v.y;
    }
  }
  s = block_sum(s);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

void launch_kinetic_energy(const LaunchCtx &c, const DevParams &p, const DevArrays &a, double *out) {
  int nb = min(nblk(p.n, RED_TPB), RED_MAX_BLOCKS);
  if (nb < 1) nb = 1;
  if (p.num_cells > 0) ensure_rank_maps(c, p, a);
  MAVI_LAUNCH(c, k_kinetic, nb, RED_TPB, 0, p, a.tile_prefix, a.cta_first, a.vel, a.reduce_buf);
  MAVI_LAUNCH(c, k_reduce_final, 1, RED_TPB, 0, a.reduce_buf, nb, 0.5, out);
}

// potential_energy(::LenJonesCfg), src/quantities.jl:46-66.
// exact mode: every pair i<j of slots 1:count, O(N^2), on the dense staging copy (no cutoff, min image)
template <bool PER>
__global__ void k_potential_exact(const __grid_constant__ DevParams p, const unsigned int *__restrict__ st_id,
                                  const real2 *__restrict__ st_pos, double *__restrict__ partials) {
  double s = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
    if ((int)(st_id[i] & ~MAVI_INACTIVE_BIT) >= p.n_count) continue;
    const real2 ri = st_pos[i];
    for (int j = i + 1; j < p.n; j++) {
      if ((int)(st_id[j] & ~MAVI_INACTIVE_BIT) >= p.n_count) continue;
      const real2 rj = __ldg(st_pos + j);
      real dx = min_image<PER>(ri.x - rj.x, p.half[0], p.size[0]);
      real dy = min_image<PER>(ri.y - rj.y, p.half[1], p.size[1]);
      double s2 = (double)p.lj_sig2 / ((double)dx * dx + (double)dy * dy);
      double s6 = s2 * s2 * s2;
      s += s6 * s6 - s6;
    }
  }
  s = block_sum(s);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

// stencil mode: the cell-stencil pair set (each pair seen from both ends -> halved by the caller)
template <bool PER>
__global__ void k_potential_stencil(const __grid_constant__ DevParams p, const int *__restrict__ tstart,
                                    const int *__restrict__ tile_prefix, const int *__restrict__ cta_first,
                                    const int *__restrict__ cell, const unsigned int *__restrict__ idflag,
                                    const real2 *__restrict__ pos, double *__restrict__ partials) {
  double s = 0.0;
  for (int base = blockIdx.x * blockDim.x; base < p.n; base += gridDim.x * blockDim.x) {
    int rank = base + threadIdx.x;
    if (rank >= p.n) continue;
    const int k = slot_of_rank(p, tile_prefix, cta_first, rank);
    if (idflag[k] & MAVI_INACTIVE_BIT) continue;
    const real2 ri = pos[k];
    for_each_neighbor(p, tstart, cell[k], k, [&](int j) {
      const real2 rj = __ldg(pos + j);
      real dx = min_image<PER>(ri.x - rj.x, p.half[0], p.size[0]);
      real dy = min_image<PER>(ri.y - rj.y, p.half[1], p.size[1]);
      double s2 = (double)p.lj_sig2 / ((double)dx * dx + (double)dy * dy);
      double s6 = s2 * s2 * s2;
      s += s6 * s6 - s6;
    });
  }
  s = block_sum(s);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

void launch_potential_energy(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int mode, double *out) {
  int nb = min(nblk(p.n, RED_TPB), RED_MAX_BLOCKS);
  if (nb < 1) nb = 1;
  const double eps4 = 4.0 * (double)p.dyn[1];
  if (mode == 0) {  // caller has refreshed the staging copy
    if (p.periodic) MAVI_LAUNCH(c, (k_potential_exact<true>), nb, RED_TPB, 0, p, a.st_id, a.st_pos, a.reduce_buf);
    else MAVI_LAUNCH(c, (k_potential_exact<false>), nb, RED_TPB, 0, p, a.st_id, a.st_pos, a.reduce_buf);
    MAVI_LAUNCH(c, k_reduce_final, 1, RED_TPB, 0, a.reduce_buf, nb, eps4, out);
  } else {
    ensure_rank_maps(c, p, a);
    if (p.periodic) MAVI_LAUNCH(c, (k_potential_stencil<true>), nb, RED_TPB, 0, p, a.tstart, a.tile_prefix, a.cta_first, a.cell, a.idflag, a.pos[0], a.reduce_buf);
    else MAVI_LAUNCH(c, (k_potential_stencil<false>), nb, RED_TPB, 0, p, a.tstart, a.tile_prefix, a.cta_first, a.cell, a.idflag, a.pos[0], a.reduce_buf);
    MAVI_LAUNCH(c, k_reduce_final, 1, RED_TPB, 0, a.reduce_buf, nb, 0.5 * eps4, out);
  }
}

// =========================================================================================================
// Download helpers: un-permute slot order to original ids
// =========================================================================================================
__global__ void k_unpermute2(const __grid_constant__ DevParams p, const int *__restrict__ tile_prefix,
                             const int *__restrict__ cta_first, const unsigned int *__restrict__ idflag,
                             const real2 *__restrict__ in, real2 *__restrict__ out) {
  int rank = blockIdx.x * blockDim.x + threadIdx.x;
  if (rank >= p.n) return;
  const int k = slot_of_rank(p, tile_prefix, cta_first, rank);
  out[idflag[k] & ~MAVI_INACTIVE_BIT] = in[k];
}
__global__ void k_unpermute1(const __grid_constant__ DevParams p, const int *__restrict__ tile_prefix,
                             const int *__restrict__ cta_first, const unsigned int *__restrict__ idflag,
                             const real *__restrict__ in, real *__restrict__ out) {
  int rank = blockIdx.x * blockDim.x + threadIdx.x;
  if (rank >= p.n) return;
  const int k = slot_of_rank(p, tile_prefix, cta_first, rank);
  out[idflag[k] & ~MAVI_INACTIVE_BIT] = in[k];
}
__global__ void k_unpermute_cells(const __grid_constant__ DevParams p, const int *__restrict__ tile_prefix,
                                  const int *__restrict__ cta_first, const unsigned int *__restrict__ idflag,
                                  const int *__restrict__ cell, int *__restrict__ out) {
  int rank = blockIdx.x * blockDim.x + threadIdx.x;
  if (rank >= p.n) return;
  const int k = slot_of_rank(p, tile_prefix, cta_first, rank);
  out[idflag[k] & ~MAVI_INACTIVE_BIT] = rank < p.n_active ? cell[k] : -1;
}
// num_particles_in_chunk (row fastest) from tstart
__global__ void k_cell_counts(const __grid_constant__ DevParams p, const int *__restrict__ tstart, int *__restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.num_cells) return;
  const int col = c / p.num_rows, row = c - col * p.num_rows;
  const int q = tq_of(p, col, row);
  out[c] = tstart[q + 1] - tstart[q];
}
// ids of the active particles in (cell, id) order, i.e. as the reference's chunk_particles lists them:
// out[cell_start[c] + i] = id of the i-th particle of cell c   (cell_start = exclusive scan of the cell populations)
__global__ void k_ids_in_cell_order(const __grid_constant__ DevParams p, const int *__restrict__ tstart,
                                    const int *__restrict__ cell_start, const unsigned int *__restrict__ idflag,
                                    int *__restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.num_cells) return;
  const int col = div_rows(p, c), row = c - col * p.num_rows;
  const int q = tq_of(p, col, row);
  const int b = tstart[q], e = tstart[q + 1], d = cell_start[c];
  for (int j = b; j < e; j++) out[d + (j - b)] = (int)(idflag[j] & ~MAVI_INACTIVE_BIT);
}

void launch_unpermute2(const LaunchCtx &c, const DevParams &p, const DevArrays &a, const real2 *in, real2 *out) {
  if (p.num_cells > 0) ensure_rank_maps(c, p, a);
  MAVI_LAUNCH(c, k_unpermute2, nblk(p.n), TPB, 0, p, a.tile_prefix, a.cta_first, a.idflag, in, out);
}
void launch_unpermute1(const LaunchCtx &c, const DevParams &p, const DevArrays &a, const real *in, real *out) {
  if (p.num_cells > 0) ensure_rank_maps(c, p, a);
  MAVI_LAUNCH(c, k_unpermute1, nblk(p.n), TPB, 0, p, a.tile_prefix, a.cta_first, a.idflag, in, out);
}
void launch_unpermute_cells(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int *out) {
  if (p.num_cells > 0) ensure_rank_maps(c, p, a);
  MAVI_LAUNCH(c, k_unpermute_cells, nblk(p.n), TPB, 0, p, a.tile_prefix, a.cta_first, a.idflag, a.cell, out);
}
void launch_cell_counts(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int *out) {
  MAVI_LAUNCH(c, k_cell_counts, nblk(p.num_cells), TPB, 0, p, a.tstart, out);
}
void launch_ids_in_cell_order(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int *out) {
  // cell populations -> exclusive scan (count[] is free between builds; re-zeroed by the next build)
  MAVI_LAUNCH(c, k_cell_counts, nblk(p.num_cells), TPB, 0, p, a.tstart, a.perm);
  launch_exclusive_scan(c, a.perm, a.count, a.scan_partials, p.num_cells);
  MAVI_LAUNCH(c, k_ids_in_cell_order, nblk(p.num_cells), TPB, 0, p, a.tstart, a.count, a.idflag, out);
}

}  // namespace MAVI_NS
