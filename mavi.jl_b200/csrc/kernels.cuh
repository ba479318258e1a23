// kernels.cuh — host-callable launchers of the device kernels (definitions in kernels.cu / rings.cu).
#pragma once
#include "common.cuh"

namespace MAVI_NS {

constexpr int TPB = 256;   // threads per block
constexpr int RPB = 1024;  // ranks per block of every rank-mapped kernel: each thread handles RPB/TPB particles
                           // (cta_first[] is indexed by rank / RPB); amortises the CTA-level staging

// Device arrays of one handle.
//
// Particle state lives in SLOTS.  With chunks, slots are grouped in padded tiles (see DevParams): tile t owns slots
// [t*cap, (t+1)*cap) and keeps its particles sorted by (cell, original id) in the prefix — the reference's chunk fill
// order (src/chunks.jl:153-155) — so update_chunks! is an O(#movers) repair of the few tiles a particle left or
// entered instead of a global re-sort.  Inactive particles (never binned by the reference) live in a tail region.
// Kernels run over RANKS 0..n-1 (dense) and map rank -> slot through tile_prefix[] / cta_first[].
struct DevArrays {
  // slot-indexed state
  real2 *pos[2];      // ping-pong: pos[0] is the current state
  real2 *vel;         // SecondLawState velocities
  real *ang;          // SelfPropelledState pol_angle
  unsigned int *idflag; // original id of the particle in the slot (bit 31: inactive)
  int *cell;            // cell the particle is binned in
  real2 *force;       // F (get_forces)
  real2 *force_old;   // F1 of the Verlet step
  // staging in dense rank order (upload, downloads, full rebuilds)
  real2 *st_pos, *st_vel, *st_force;
  real *st_ang;
  unsigned int *st_id;
  int *st_cell;
  // tile bookkeeping
  int *tstart;          // [nt*(MAVI_TR+1)]
  int *tile_prefix;     // [nt+2] exclusive scan of the tile populations
  int *cta_first;       // [n/TPB+2] tile holding rank b*TPB
  int *count;           // [num_cells+2] histogram scratch of full builds
  int *perm;            // [ns+nt+2] scatter scratch / int scratch
  int *scan_partials;
  // incremental repair
  int *tile_dirty;      // [nt] 0/1
  int *dirty_list;      // [nt] tiles to repair this step
  int *inbox_cnt;       // [nt] particles arriving from other tiles
  int *inbox;           // [nt*inbox_cap] -> index into the mover list
  int *mv_src;          // [mv_cap] source slot of an inter-tile mover
  real2 *mv_pos, *mv_second, *mv_force;
  unsigned int *mv_id;
  int *mv_cell;
  int *chg;             // [chg_cap] cells whose membership changed in this step (force carry)
  // slab mode: emigrant records to / immigrant records from the left [0] and right [1] neighbour
  EmRec *em_send[2], *em_recv[2];
  int em_cap;
  // control
  int *flags;           // see FLAG_* in common.cuh
  int *fix_idx;         // sparse list of slots whose position walls! changed in pass B
  real2 *fix_pos;
  double *reduce_buf;
  double2 *edge_x, *edge_y;  // cell edges of the grid (k_cell_edges): [num_cols + 2], [tpc * MAVI_TR]
};

struct LaunchCtx {
  cudaStream_t stream;
  long long *launches;  // incremented per kernel launch
  bool *maps_valid;     // rank maps (tile_prefix / cta_first) match tstart; the step loop only invalidates them
  int flags;            // MaviParams.flags (kernel-variant switches)
};

// full build of the tile layout from the staging arrays (upload, mavi_bin, overflow fallback)
void launch_check_inside(const LaunchCtx &c, const DevParams &p, const DevArrays &a);
void launch_build_tiles(const LaunchCtx &c, const DevParams &p, const DevArrays &a, bool second_is_vel);
// Mavi.Rings: bin particle indices only (perm[slot] = particle index, ascending inside every cell)
void launch_build_index_tiles(const LaunchCtx &c, const DevParams &p, const real2 *pos, const unsigned int *idflag,
                              int *cell_out, int *count, int *tstart, int *perm, int *flags, real2 *spos = nullptr);
// dense copy of the current state into staging (rank order)
void launch_compact_to_staging(const LaunchCtx &c, const DevParams &p, const DevArrays &a, bool second_is_vel);
void launch_compact_cells(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int *out);
void launch_init_staging_ids(const LaunchCtx &c, int n, const unsigned char *mask, unsigned int *st_id);
void launch_ids_to_i64(const LaunchCtx &c, int n, const unsigned int *st_id, long long *out);
void launch_ids_from_i64(const LaunchCtx &c, int n, const long long *in, unsigned int *st_id);
// incremental update_chunks!: repair the tiles touched by this step's movers, then refresh the rank maps
void launch_repair_tiles(const LaunchCtx &c, const DevParams &p, const DevArrays &a, bool second_is_vel);
void launch_exclusive_scan(const LaunchCtx &c, const int *in, int *out, int *partials, int n);
void refresh_rank_maps(const LaunchCtx &c, const DevParams &p, const DevArrays &a);
// the rank -> slot maps are only needed by downloads, reductions and rebuilds: refreshed on demand
void ensure_rank_maps(const LaunchCtx &c, const DevParams &p, const DevArrays &a);

// force + integrate passes
void launch_step_begin(const LaunchCtx &c, const DevArrays &a);
void launch_cell_edges(const LaunchCtx &c, const DevParams &p, const DevArrays &a);  // after the grid / slab geometry is final
void launch_force_only(const LaunchCtx &c, const DevParams &p, const DevArrays &a, bool with_wall_forces);
void launch_newton_a(const LaunchCtx &c, const DevParams &p, const DevArrays &a);
// blk_mode: 0 = every block, then the wall position fix-ups; 1 = blocks that read no halo column; 2 = the first / last
// block of every tile row (1 and 2: the caller applies the fix-ups after BOTH launches)
void launch_apply_pos_fixes(const LaunchCtx &c, const DevArrays &a);
void launch_newton_b(const LaunchCtx &c, const DevParams &p, const DevArrays &a, bool carry = false, int blk_mode = 0);
// force carry (see k_newton_b): redo the carried drift of repaired tiles / recompute F1 around re-binned cells
void launch_carry_redrift(const LaunchCtx &c, const DevParams &p, const DevArrays &a);
void launch_carry_recompute(const LaunchCtx &c, const DevParams &p, const DevArrays &a);
void launch_carry_recompute_list(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int skip_edge);
void launch_carry_recompute_columns(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int depth, bool report_big);
void launch_carry_fixups(const LaunchCtx &c, const DevParams &p, const DevArrays &a);
void launch_self_propelled(const LaunchCtx &c, const DevParams &p, const DevArrays &a, const real *noise,
                           unsigned long long step);

// quantities
void launch_kinetic_energy(const LaunchCtx &c, const DevParams &p, const DevArrays &a, double *out);
void launch_potential_energy(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int mode, double *out);

// downloads (un-permute to original ids)
void launch_unpermute2(const LaunchCtx &c, const DevParams &p, const DevArrays &a, const real2 *in, real2 *out);
void launch_unpermute1(const LaunchCtx &c, const DevParams &p, const DevArrays &a, const real *in, real *out);
void launch_unpermute_cells(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int *out);
void launch_cell_counts(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int *out);
void launch_ids_in_cell_order(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int *out);

}  // namespace MAVI_NS
