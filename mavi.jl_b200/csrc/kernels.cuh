// kernels.cuh — host-callable launchers of the device kernels (definitions in kernels.cu / rings.cu).
#pragma once
#include "common.cuh"

namespace mavi {

// Device arrays of one handle.  Physical order is "sorted by cell, ascending original id inside a cell"
// (the reference's chunk fill order, src/chunks.jl:153-155); `idflag[k]` is the original id of slot k
// (bit 31 set for inactive particles, which live in the pseudo-cell `num_cells` at the tail).
struct DevArrays {
  double2 *pos[2];      // ping-pong: pos[cur] is the current state
  double2 *vel[2];      // SecondLawState velocities (ping-pong for the re-sort)
  double *ang[2];       // SelfPropelledState pol_angle (ping-pong for the re-sort)
  unsigned int *idflag[2];
  int *cell[2];         // cell of each sorted slot
  double2 *force;       // F (get_forces)
  double2 *force_old;   // F1 of the Verlet step / re-sort scratch for forces
  int *cell_new;        // fresh cell ids before the re-sort
  int *perm;            // scatter result: slot -> source slot
  int *count;           // [num_cells+2] histogram / scatter cursors
  int *start;           // [num_cells+2] exclusive scan of count
  int *scan_partials;   // block sums of the scan
  int *flags;           // [0] error bits, [1] #particles that left their sorted cell, [2] big-drift guard, [3] #position fix-ups
  int *fix_idx;         // sparse list of slots whose position walls! changed in pass B
  double2 *fix_pos;
  double *reduce_buf;   // block partials of the energy reductions
};

struct LaunchCtx {
  cudaStream_t stream;
  long long *launches;  // incremented per kernel launch
};

// binning / counting sort
void launch_cell_index(const LaunchCtx &c, const DevParams &p, const double2 *pos, const unsigned int *idflag,
                       const int *cell_old, int *cell_new, int *count, int *flags);
void launch_exclusive_scan(const LaunchCtx &c, const int *in, int *out, int *partials, int n);
void launch_scatter(const LaunchCtx &c, const DevParams &p, const int *cell_new, const int *start, int *count, int *perm);
void launch_gather(const LaunchCtx &c, const DevParams &p, const int *perm, const int *cell_new, const int *start,
                   const DevArrays &a, int src, int dst, bool second_is_vel, bool has_second, bool with_forces);

// force + integrate passes (cell list)
void launch_force_only(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int cur, bool with_wall_forces);
void launch_newton_a(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int cur);
void launch_newton_b(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int cur);
void launch_self_propelled(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int cur, const double *noise,
                           unsigned long long step);

// quantities
void launch_kinetic_energy(const LaunchCtx &c, const DevParams &p, const double2 *vel, double *partials, double *out);
void launch_potential_energy(const LaunchCtx &c, const DevParams &p, const DevArrays &a, int cur, int mode, double *out);

// un-permute helpers for downloads: out[id[k]] = in[k]
void launch_unpermute2(const LaunchCtx &c, int n, const unsigned int *idflag, const double2 *in, double2 *out);
void launch_unpermute1(const LaunchCtx &c, int n, const unsigned int *idflag, const double *in, double *out);
void launch_unpermute_cells(const LaunchCtx &c, int n, int num_cells, const unsigned int *idflag, const int *cell, int *out);
void launch_ids(const LaunchCtx &c, int n, const unsigned int *idflag, int *out);
void launch_init_ids(const LaunchCtx &c, int n, const unsigned char *mask, unsigned int *idflag, int *cell, int num_cells);
void launch_check_inside(const LaunchCtx &c, const DevParams &p, const double2 *pos, const unsigned int *idflag, int *flags);

}  // namespace mavi
