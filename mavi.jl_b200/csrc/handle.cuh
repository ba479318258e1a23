// handle.cuh — host-side state of one MaviHandle (one device, one stream).
#pragma once
#include <vector>

#include "kernels.cuh"

namespace MAVI_NS {

enum SecondKind { SECOND_VEL = 0, SECOND_ANGLE = 1, SECOND_RING_POL = 2 };

// Rings scratch (RingsInfo, src/rings/rings.jl:118-128)
struct RingsArrays {
  real2 *cont_pos = nullptr;  // continuos_pos as of the last unwrap
  real *areas = nullptr;
  real2 *cms = nullptr;       // info.cms (lags one step, src/rings/integration.jl:523)
  real *pol = nullptr;        // state.pol, one angle per ring
  real2 *spos = nullptr;      // positions in index-tile slot order (spos[s] = pos[perm[s]]), rewritten by every binning
  // ParticleNeighbors (src/rings/neighbors.jl, src/rings/rings.jl:143-158): contact counts / lists of the last forces!
  int *neigh_count = nullptr;   // [n]
  int *neigh_list = nullptr;    // [n][MAVI_NEIGH_MAX], ascending ids, -1 padded (list mode only)
  int neigh_mode = 0;           // MAVI_NEIGH_*
  int neigh_all = 0;            // 1: type = :all, 0: type = :rings (other rings only)
  double neigh_tol = 1.1;       // NeighborsCfg.tol
  // VarRingsIds + sources / sinks (src/rings/states.jl:24-43,173-227, src/rings/sources.jl): the ring mask lives on the
  // device (binning, pair and ring kernels honour it); add_ring! / remove_ring! / calc_active_ids! are sequential by
  // definition (first free slot, uid = max + 1) and run on the host once per step from two tiny downloads
  struct Source {
    int kind = 0, nspawn = 0, nsp = 0, sink_geom = 0;
    double pad = 0.0, spawn_pol = 0.0, sink[5] = {0, 0, 0, 0, 0};
    std::vector<double> bbox;   // [nspawn][4]
    std::vector<double> spawn;  // [nspawn][nsp][2]
    int first_area = 0;         // index of its first spawn area in the device list
  };
  bool var_rings = false;
  std::vector<int> np_h, types_h;      // host copies of num_particles[type] and the 0-based ring types (empty: one type)
  std::vector<Source> sources;
  std::vector<unsigned char> mask_h;   // rings_ids.mask
  std::vector<long long> uids_h, ids_h;  // rings_ids.uids, rings_ids.ids[1:num_active] as of the last calc_active_ids!
  long long num_active = 0;
  std::vector<double> draws;           // rand(rng) stand-ins for spawn_pol = :random
  size_t draw_pos = 0;
  unsigned long long spawn_count = 0;  // production mode: counter of the host-side generator
  unsigned char *mask_dev = nullptr;   // [num_rings]
  double *areas_dev = nullptr;         // spawn areas: [n_areas][5] = bl.x, bl.y, length, height, pad
  int *empty_dev = nullptr;            // [n_areas] 1 = no active particle inside
  int n_areas = 0;
  bool has_sinks = false;
  // invasions (src/rings/integration.jl:379-520): InvasionsCfg.steps_to_update, ring-level chunks on the centres of mass
  int inv_steps = 0;              // 0 = off
  long long inv_last_check = 0;   // InvasionsInfo.last_check
  int r_cols = 0, r_rows = 0;     // r_chunks_cfg (0: check every pair of rings)
  int *rcell = nullptr, *rcount = nullptr, *rstart = nullptr, *rperm = nullptr, *rpart = nullptr;
  int *inv_list = nullptr;        // [inv_cap][3] = invasor ring, invaded ring, scalar particle id
  int *inv_n = nullptr;
  int inv_cap = 0;
};

// x-slab decomposition state (slab.cu)
struct SlabState {
  int world = 1, rank = 0, left = 0, right = 0;
  int m = 0, m_left = 0, m_right = 0, col_lo = 0;
  int n_global = 0;
  struct ncclComm *comm = nullptr;
  cudaStream_t side = nullptr;                  // the drifted-halo exchange overlaps with the interior blocks
  cudaEvent_t ev_ready = nullptr, ev_halo = nullptr, ev_begin = nullptr, ev_side = nullptr;
  bool side_busy = false;  // work of the step pipeline is in flight on the side stream (joined by slab_join)
};

struct Handle {
  DevParams p;
  DevArrays a;
  RingsArrays r;
  SlabState slab;
  int n_cap = 0;  // capacity of the dense staging arrays (slab mode: the owned count changes with migration)
  int device = 0;
  int flags_cfg = 0;
  cudaStream_t stream = nullptr;
  SecondKind second_kind = SECOND_VEL;
  int rings_n_active = 0;  // active particle slots of a RingsState
  size_t ns = 0;  // number of particle slots (tiles * cap + inactive tail)
  int *flags_host = nullptr;  // pinned mirror of a.flags[0..3]
  bool prof = false;
  bool carry_valid = false;  // pos[1] / force_old hold the drift and F1 of the next Newton step (force carry)
  long long launches = 0;
  long long n_rebuilds = 0;  // overflow -> rebuild events (instrumentation)
  int steps_seen = 0;  // host mirror of the device-side step counter flags[FLAG_STEPS]
  long long num_steps = 0;
  double time = 0.0;
  double grid_len_host = 0.0;  // length of the bounding box of the geometry (MaviParams.grid_len)
  double dt_host = 0.0;  // IntCfg.dt as the host passed it (Float64): TimeInfo accumulates it in Float64 in both builds
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_call[2] = {nullptr, nullptr};
  std::vector<void *> allocs;
  real *noise_dev = nullptr;
  size_t noise_cap = 0;
  long long *id64_dev = nullptr;  // slab mode: staging for the int64 ids of mavi_upload_local / mavi_download_local
  size_t id64_cap = 0;
  int ensure_id64(size_t n);
  char err[512] = {0};

  bool maps_valid = false;
  LaunchCtx ctx() { return LaunchCtx{stream, &launches, &maps_valid, flags_cfg}; }
  void set_error(const char *fmt, ...);
  int check_device_flags();
  int pending_out_of_grid();
  int alloc_state(int n_active, int cap);
  int rebuild_from_staging(int n_active);
  int rebuild_from_current();
  int enqueue_step(const real *noise_dev);
  int run_steps(long long nsteps, const real *noise_dev, size_t stride);
};

// api.cu
void mavi_magic_div(unsigned int d, unsigned int *mul, unsigned int *shr);

// slab.cu
int slab_unique_id(void *out128);
int slab_configure(Handle *h, const MaviParams *mp);
void slab_destroy(Handle *h);
int slab_after_build(Handle *h);
int slab_allreduce_max(Handle *h, int *value);
int slab_step_once(Handle *h, const real *noise_dev);
int slab_sync_counts(Handle *h);
int slab_join(Handle *h);

// rings.cu
int rings_lower(Handle *h, const MaviParams *mp);
int rings_allocate(Handle *h);
int rings_upload_finish(Handle *h);
int rings_step(Handle *h, const real *noise_dev);
int rings_grow_tiles(Handle *h);
int rings_calc_forces(Handle *h);
int rings_download_info(Handle *h, void *areas, void *cms, void *cont_pos);
int rings_download_state(Handle *h, void *pos, void *second);
int rings_download_forces(Handle *h, void *forces);
int rings_bin(Handle *h);
int rings_download_cells(Handle *h, int *cell_of_particle, int *counts, int *start, int *ids);
int rings_set_neighbors(Handle *h, int mode, int type_all, double tol);
int rings_download_neighbors(Handle *h, int *count, int *list);
int rings_set_sources(Handle *h, const MaviSourceSink *list, int n, const unsigned char *ring_active, const double *draws, long long n_draws);
int rings_download_active(Handle *h, unsigned char *mask, long long *uids, long long *num_active);
int rings_set_invasions(Handle *h, int steps_to_update, int r_cols, int r_rows);
int rings_download_invasions(Handle *h, long long *n, int *triples, long long cap);

}  // namespace MAVI_NS
