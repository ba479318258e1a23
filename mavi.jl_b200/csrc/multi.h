// multi.h — single-process multi-GPU handle (multi.cu): entry points mirroring include/mavi.h, and the table of
// per-arithmetic-type implementations (api_decl.inc) it drives its sub-handles through.
#pragma once
#include <stdint.h>

#include "../../include/mavi.h"

namespace mavi_multi {

struct ApiTable {
  int32_t (*create)(const MaviParams *, void **);
  int32_t (*destroy)(void *);
  int32_t (*last_error)(void *, char *, int32_t);
  int32_t (*local_count)(void *, int64_t *);
  int32_t (*download_local)(void *, int64_t *, void *, void *, void *);
  int32_t (*upload_local)(void *, const int64_t *, const void *, const void *, int64_t);
  int32_t (*download_local_cells)(void *, int32_t *);
  int32_t (*nccl_unique_id)(void *);
  int32_t (*step)(void *, int64_t, const void *);
  int32_t (*calc_forces)(void *);
  int32_t (*bin)(void *);
  int32_t (*energies)(void *, int32_t, double *, double *);
  int32_t (*get_time)(void *, int64_t *, double *);
  int32_t (*set_time)(void *, int64_t, double);
  int32_t (*sync)(void *);
  int32_t (*launch_count)(void *, int64_t *);
  int32_t (*rebuild_count)(void *, int64_t *);
  int32_t (*last_step_ms)(void *, float *);
  int32_t (*set_profiling)(void *, int32_t);
  int32_t (*counters)(void *, int64_t *);
};

int create(const ApiTable *api, const MaviParams *p, void **out);
int destroy(void *m);
int last_error(void *m, char *buf, int32_t n);
int upload_state(void *m, const void *pos, const void *second, const uint8_t *mask, int64_t n);
int download_state(void *m, void *pos, void *second);
int download_forces(void *m, void *forces);
int local_count(void *m, int64_t *n);
int download_local(void *m, int64_t *ids, void *pos, void *second, void *forces);
int upload_local(void *m, const int64_t *ids, const void *pos, const void *second, int64_t n);
int step(void *m, int64_t nsteps, const void *host_noise);
int calc_forces(void *m);
int bin(void *m);
int sync(void *m);
int set_profiling(void *m, int32_t on);
int download_cells(void *m, int32_t *cell_of_particle, int32_t *counts);
int download_cell_lists(void *m, int32_t *start, int32_t *ids);
int cell_neighbors(void *m, int32_t cell, int32_t *out8, int32_t *n);
int energies(void *m, int32_t pe_mode, double *ke, double *pe);
int get_time(void *m, int64_t *num_steps, double *time);
int set_time(void *m, int64_t num_steps, double time);
int launch_count(void *m, int64_t *n);
int rebuild_count(void *m, int64_t *n);
int last_step_ms(void *m, float *ms5);
int counters(void *m, int64_t *out8);

}  // namespace mavi_multi
