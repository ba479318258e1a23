// capi.cu — the extern "C" entry points of include/mavi.h.  Every device source is compiled twice (real.cuh): namespace
// mavi_f64 (Float64, the default) and mavi_f32 (Float32 mode); a MaviHandle remembers which build owns it and every call
// is forwarded to that build's api_<name>.  There is no other logic here.
#include <stdint.h>

#include <new>

#include "../../include/mavi.h"

#define MAVI_NS mavi_f64
#include "api_decl.inc"
#undef MAVI_NS
#define MAVI_NS mavi_f32
#include "api_decl.inc"
#undef MAVI_NS

#include "hostcell.h"
#include "multi.h"

struct MaviHandle {
  int32_t dtype;
  void *impl;
  bool multi;  // impl is a mavi_multi handle (MaviParams.n_gpus > 1): one x-slab sub-handle per device inside
};

#define MAVI_TABLE(NS)                                                                                                  \
  {NS::api_create, NS::api_destroy, NS::api_last_error, NS::api_local_count, NS::api_download_local, NS::api_upload_local, \
   NS::api_download_local_cells, NS::api_nccl_unique_id, NS::api_step, NS::api_calc_forces, NS::api_bin, NS::api_energies, \
   NS::api_get_time, NS::api_set_time, NS::api_sync, NS::api_launch_count, NS::api_rebuild_count, NS::api_last_step_ms,    \
   NS::api_set_profiling, NS::api_counters}
static const mavi_multi::ApiTable g_table_f64 = MAVI_TABLE(mavi_f64);
static const mavi_multi::ApiTable g_table_f32 = MAVI_TABLE(mavi_f32);

#define MAVI_FWD(h, call)                                     \
  do {                                                        \
    if (!(h)) return MAVI_ERR_BAD_PARAMS;                     \
    void *impl = (h)->impl;                                   \
    if ((h)->dtype == MAVI_F32) return mavi_f32::call;        \
    return mavi_f64::call;                                    \
  } while (0)
// entry points the multi-GPU handle implements itself
#define MAVI_FWD_M(h, mcall, call)                            \
  do {                                                        \
    if (!(h)) return MAVI_ERR_BAD_PARAMS;                     \
    void *impl = (h)->impl;                                   \
    if ((h)->multi) return mavi_multi::mcall;                 \
    if ((h)->dtype == MAVI_F32) return mavi_f32::call;        \
    return mavi_f64::call;                                    \
  } while (0)
// entry points that make no sense on it (Mavi.Rings is single-GPU)
#define MAVI_FWD_S(h, call)                                   \
  do {                                                        \
    if (!(h) || (h)->multi) return MAVI_ERR_BAD_PARAMS;       \
    void *impl = (h)->impl;                                   \
    if ((h)->dtype == MAVI_F32) return mavi_f32::call;        \
    return mavi_f64::call;                                    \
  } while (0)

extern "C" {

int32_t mavi_abi_version(void) { return mavi_f64::api_abi_version(); }

int32_t mavi_create(const MaviParams *params, MaviHandle **out) {
  if (!params || !out) return MAVI_ERR_BAD_PARAMS;
  *out = nullptr;
  if (params->dtype != MAVI_F64 && params->dtype != MAVI_F32) return MAVI_ERR_BAD_PARAMS;
  MaviHandle *h = new (std::nothrow) MaviHandle{params->dtype, nullptr, false};
  if (!h) return MAVI_ERR_BAD_PARAMS;
  // on failure the handle is still returned so that mavi_last_error can be read; the caller destroys it
  *out = h;
  if (params->n_gpus > 1) {
    h->multi = true;
    return mavi_multi::create(params->dtype == MAVI_F32 ? &g_table_f32 : &g_table_f64, params, &h->impl);
  }
  return params->dtype == MAVI_F32 ? mavi_f32::api_create(params, &h->impl) : mavi_f64::api_create(params, &h->impl);
}

int32_t mavi_destroy(MaviHandle *h) {
  if (!h) return MAVI_OK;
  int32_t st = h->multi ? mavi_multi::destroy(h->impl)
                        : (h->dtype == MAVI_F32 ? mavi_f32::api_destroy(h->impl) : mavi_f64::api_destroy(h->impl));
  delete h;
  return st;
}

int32_t mavi_last_error(MaviHandle *h, char *buf, int32_t n) { MAVI_FWD_M(h, last_error(impl, buf, n), api_last_error(impl, buf, n)); }
int32_t mavi_upload_state(MaviHandle *h, const void *pos, const void *second, const uint8_t *active_mask, int64_t n) {
  MAVI_FWD_M(h, upload_state(impl, pos, second, active_mask, n), api_upload_state(impl, pos, second, active_mask, n));
}
int32_t mavi_download_state(MaviHandle *h, void *pos, void *second) {
  MAVI_FWD_M(h, download_state(impl, pos, second), api_download_state(impl, pos, second));
}
int32_t mavi_download_forces(MaviHandle *h, void *forces) {
  MAVI_FWD_M(h, download_forces(impl, forces), api_download_forces(impl, forces));
}
int32_t mavi_local_count(MaviHandle *h, int64_t *n_local) { MAVI_FWD_M(h, local_count(impl, n_local), api_local_count(impl, n_local)); }
int32_t mavi_download_local(MaviHandle *h, int64_t *ids, void *pos, void *second, void *forces) {
  MAVI_FWD_M(h, download_local(impl, ids, pos, second, forces), api_download_local(impl, ids, pos, second, forces));
}
int32_t mavi_nccl_unique_id(void *out128) { return mavi_f64::api_nccl_unique_id(out128); }
int32_t mavi_upload_local(MaviHandle *h, const int64_t *ids, const void *pos, const void *second, int64_t n_local) {
  MAVI_FWD_M(h, upload_local(impl, ids, pos, second, n_local), api_upload_local(impl, ids, pos, second, n_local));
}
int32_t mavi_step(MaviHandle *h, int64_t nsteps, const void *host_noise) {
  MAVI_FWD_M(h, step(impl, nsteps, host_noise), api_step(impl, nsteps, host_noise));
}
int32_t mavi_calc_forces(MaviHandle *h) { MAVI_FWD_M(h, calc_forces(impl), api_calc_forces(impl)); }
int32_t mavi_bin(MaviHandle *h) { MAVI_FWD_M(h, bin(impl), api_bin(impl)); }
int32_t mavi_download_cells(MaviHandle *h, int32_t *cell_of_particle, int32_t *counts) {
  MAVI_FWD_M(h, download_cells(impl, cell_of_particle, counts), api_download_cells(impl, cell_of_particle, counts));
}
int32_t mavi_download_cell_lists(MaviHandle *h, int32_t *start, int32_t *ids) {
  MAVI_FWD_M(h, download_cell_lists(impl, start, ids), api_download_cell_lists(impl, start, ids));
}
int32_t mavi_cell_neighbors(MaviHandle *h, int32_t cell, int32_t *out8, int32_t *n) {
  MAVI_FWD_M(h, cell_neighbors(impl, cell, out8, n), api_cell_neighbors(impl, cell, out8, n));
}
int32_t mavi_energies(MaviHandle *h, int32_t pe_mode, double *ke, double *pe) {
  MAVI_FWD_M(h, energies(impl, pe_mode, ke, pe), api_energies(impl, pe_mode, ke, pe));
}
int32_t mavi_rings_download_info(MaviHandle *h, void *areas, void *cms, void *cont_pos) {
  MAVI_FWD_S(h, api_rings_download_info(impl, areas, cms, cont_pos));
}
int32_t mavi_rings_set_neighbors(MaviHandle *h, int32_t mode, int32_t type_all, double tol) {
  MAVI_FWD_S(h, api_rings_set_neighbors(impl, mode, type_all, tol));
}
int32_t mavi_rings_download_neighbors(MaviHandle *h, int32_t *count, int32_t *list) {
  MAVI_FWD_S(h, api_rings_download_neighbors(impl, count, list));
}
int32_t mavi_rings_set_sources(MaviHandle *h, const MaviSourceSink *list, int32_t n, const uint8_t *ring_active,
                               const double *spawn_draws, int64_t n_draws) {
  MAVI_FWD_S(h, api_rings_set_sources(impl, list, n, ring_active, spawn_draws, n_draws));
}
int32_t mavi_rings_download_active(MaviHandle *h, uint8_t *ring_active, int64_t *uids, int64_t *num_active) {
  MAVI_FWD_S(h, api_rings_download_active(impl, ring_active, uids, num_active));
}
int32_t mavi_rings_set_invasions(MaviHandle *h, int32_t steps_to_update, int32_t r_cols, int32_t r_rows) {
  MAVI_FWD_S(h, api_rings_set_invasions(impl, steps_to_update, r_cols, r_rows));
}
int32_t mavi_rings_download_invasions(MaviHandle *h, int64_t *n, int32_t *triples, int64_t cap) {
  MAVI_FWD_S(h, api_rings_download_invasions(impl, n, triples, cap));
}
int32_t mavi_get_time(MaviHandle *h, int64_t *num_steps, double *time) {
  MAVI_FWD_M(h, get_time(impl, num_steps, time), api_get_time(impl, num_steps, time));
}
int32_t mavi_set_time(MaviHandle *h, int64_t num_steps, double time) {
  MAVI_FWD_M(h, set_time(impl, num_steps, time), api_set_time(impl, num_steps, time));
}
int32_t mavi_sync(MaviHandle *h) { MAVI_FWD_M(h, sync(impl), api_sync(impl)); }
int32_t mavi_launch_count(MaviHandle *h, int64_t *n) { MAVI_FWD_M(h, launch_count(impl, n), api_launch_count(impl, n)); }
int32_t mavi_rebuild_count(MaviHandle *h, int64_t *n) { MAVI_FWD_M(h, rebuild_count(impl, n), api_rebuild_count(impl, n)); }
int32_t mavi_last_step_ms(MaviHandle *h, float *ms5) { MAVI_FWD_M(h, last_step_ms(impl, ms5), api_last_step_ms(impl, ms5)); }
int32_t mavi_set_profiling(MaviHandle *h, int32_t on) { MAVI_FWD_M(h, set_profiling(impl, on), api_set_profiling(impl, on)); }
int32_t mavi_counters(MaviHandle *h, int64_t *out8) { MAVI_FWD_M(h, counters(impl, out8), api_counters(impl, out8)); }

// update_particle_chunk! (src/chunks.jl:120-147) on the host with the device's exact arithmetic: 0-based linear cell id
// (col * num_rows + row) of every point, -1 when it lies outside the grid.  Needs no GPU.
int32_t mavi_cells_of_points(const MaviParams *params, const void *pos, int64_t n, int32_t *cell_out) {
  if (!params || !pos || !cell_out || n < 0 || params->num_cols <= 0 || params->num_rows <= 0) return MAVI_ERR_BAD_PARAMS;
  mavi_host::Grid g;
  g.bl[0] = params->grid_bl[0];
  g.bl[1] = params->grid_bl[1];
  g.h = params->grid_h;
  g.cl = params->grid_len / (double)params->num_cols;
  g.ch = params->grid_h / (double)params->num_rows;
  g.cols = params->num_cols;
  g.rows = params->num_rows;
  const bool f32 = params->dtype == MAVI_F32;
  for (int64_t i = 0; i < n; i++) {
    const double x = f32 ? (double)((const float *)pos)[2 * i] : ((const double *)pos)[2 * i];
    const double y = f32 ? (double)((const float *)pos)[2 * i + 1] : ((const double *)pos)[2 * i + 1];
    int col, row;
    cell_out[i] = mavi_host::cell_of_point(g, x, y, &col, &row) ? col * g.rows + row : -1;
  }
  return MAVI_OK;
}

}  // extern "C"
