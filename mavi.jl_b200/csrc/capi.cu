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

struct MaviHandle {
  int32_t dtype;
  void *impl;
};

#define MAVI_FWD(h, call)                                     \
  do {                                                        \
    if (!(h)) return MAVI_ERR_BAD_PARAMS;                     \
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
  MaviHandle *h = new (std::nothrow) MaviHandle{params->dtype, nullptr};
  if (!h) return MAVI_ERR_BAD_PARAMS;
  // on failure the handle is still returned so that mavi_last_error can be read; the caller destroys it
  *out = h;
  return params->dtype == MAVI_F32 ? mavi_f32::api_create(params, &h->impl) : mavi_f64::api_create(params, &h->impl);
}

int32_t mavi_destroy(MaviHandle *h) {
  if (!h) return MAVI_OK;
  int32_t st = h->dtype == MAVI_F32 ? mavi_f32::api_destroy(h->impl) : mavi_f64::api_destroy(h->impl);
  delete h;
  return st;
}

int32_t mavi_last_error(MaviHandle *h, char *buf, int32_t n) { MAVI_FWD(h, api_last_error(impl, buf, n)); }
int32_t mavi_upload_state(MaviHandle *h, const void *pos, const void *second, const uint8_t *active_mask, int64_t n) {
  MAVI_FWD(h, api_upload_state(impl, pos, second, active_mask, n));
}
int32_t mavi_download_state(MaviHandle *h, void *pos, void *second) { MAVI_FWD(h, api_download_state(impl, pos, second)); }
int32_t mavi_download_forces(MaviHandle *h, void *forces) { MAVI_FWD(h, api_download_forces(impl, forces)); }
int32_t mavi_local_count(MaviHandle *h, int64_t *n_local) { MAVI_FWD(h, api_local_count(impl, n_local)); }
int32_t mavi_download_local(MaviHandle *h, int64_t *ids, void *pos, void *second, void *forces) {
  MAVI_FWD(h, api_download_local(impl, ids, pos, second, forces));
}
int32_t mavi_nccl_unique_id(void *out128) { return mavi_f64::api_nccl_unique_id(out128); }
int32_t mavi_upload_local(MaviHandle *h, const int64_t *ids, const void *pos, const void *second, int64_t n_local) {
  MAVI_FWD(h, api_upload_local(impl, ids, pos, second, n_local));
}
int32_t mavi_step(MaviHandle *h, int64_t nsteps, const void *host_noise) { MAVI_FWD(h, api_step(impl, nsteps, host_noise)); }
int32_t mavi_calc_forces(MaviHandle *h) { MAVI_FWD(h, api_calc_forces(impl)); }
int32_t mavi_bin(MaviHandle *h) { MAVI_FWD(h, api_bin(impl)); }
int32_t mavi_download_cells(MaviHandle *h, int32_t *cell_of_particle, int32_t *counts) {
  MAVI_FWD(h, api_download_cells(impl, cell_of_particle, counts));
}
int32_t mavi_download_cell_lists(MaviHandle *h, int32_t *start, int32_t *ids) {
  MAVI_FWD(h, api_download_cell_lists(impl, start, ids));
}
int32_t mavi_cell_neighbors(MaviHandle *h, int32_t cell, int32_t *out8, int32_t *n) {
  MAVI_FWD(h, api_cell_neighbors(impl, cell, out8, n));
}
int32_t mavi_energies(MaviHandle *h, int32_t pe_mode, double *ke, double *pe) { MAVI_FWD(h, api_energies(impl, pe_mode, ke, pe)); }
int32_t mavi_rings_download_info(MaviHandle *h, void *areas, void *cms, void *cont_pos) {
  MAVI_FWD(h, api_rings_download_info(impl, areas, cms, cont_pos));
}
int32_t mavi_rings_set_neighbors(MaviHandle *h, int32_t mode, int32_t type_all, double tol) {
  MAVI_FWD(h, api_rings_set_neighbors(impl, mode, type_all, tol));
}
int32_t mavi_rings_download_neighbors(MaviHandle *h, int32_t *count, int32_t *list) {
  MAVI_FWD(h, api_rings_download_neighbors(impl, count, list));
}
int32_t mavi_get_time(MaviHandle *h, int64_t *num_steps, double *time) { MAVI_FWD(h, api_get_time(impl, num_steps, time)); }
int32_t mavi_set_time(MaviHandle *h, int64_t num_steps, double time) { MAVI_FWD(h, api_set_time(impl, num_steps, time)); }
int32_t mavi_sync(MaviHandle *h) { MAVI_FWD(h, api_sync(impl)); }
int32_t mavi_launch_count(MaviHandle *h, int64_t *n) { MAVI_FWD(h, api_launch_count(impl, n)); }
int32_t mavi_rebuild_count(MaviHandle *h, int64_t *n) { MAVI_FWD(h, api_rebuild_count(impl, n)); }
int32_t mavi_last_step_ms(MaviHandle *h, float *ms5) { MAVI_FWD(h, api_last_step_ms(impl, ms5)); }
int32_t mavi_set_profiling(MaviHandle *h, int32_t on) { MAVI_FWD(h, api_set_profiling(impl, on)); }
int32_t mavi_counters(MaviHandle *h, int64_t *out8) { MAVI_FWD(h, api_counters(impl, out8)); }

}  // extern "C"
