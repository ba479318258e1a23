/* mavi_oracle.h — CPU oracle for the Mavi.jl hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference algorithm (file:line citations in mavi_oracle.c), with the
 * reference's data structures (AoS doubles, Int64 cell table with the reference capacity formula,
 * half-stencil neighbour tables, Newton-3 scatter, per-thread force slices) and loop orders.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (libmavi_cuda.so) never does.
 *
 * PARITY UNPINNED: the reference is Julia and Julia is not installed in this image, the reference's
 * tests hold no golden vectors for the core path, and its Rings goldens depend on Julia's
 * MersenneTwister streams.  The oracle is pinned only by known-answer tests derived
 * from the reference source and by the reference's own portable invariants (chunks == all-pairs,
 * Threaded == Sequencial).
 *
 * The oracle is parameterised by the same flat MaviParams POD as the product (include/mavi.h), so
 * a test builds ONE parameter block and hands it to both sides.
 */
#ifndef MAVI_ORACLE_H
#define MAVI_ORACLE_H

#include "../include/mavi.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrSystem OrSystem;

int32_t mor_create(const MaviParams *p, OrSystem **out);
void mor_destroy(OrSystem *s);
/* 0 / 1 -> Sequencial (src/integration.jl:112-157); t>1 -> Threaded with t force slices (:159-194) */
void mor_set_threads(OrSystem *s, int32_t nthreads);
const char *mor_last_error(OrSystem *s);

int32_t mor_upload_state(OrSystem *s, const double *pos, const double *second, const uint8_t *mask, int64_t n);
int32_t mor_download_state(OrSystem *s, double *pos, double *second);
int32_t mor_download_forces(OrSystem *s, double *f);

int32_t mor_step(OrSystem *s, int64_t nsteps, const double *noise);
int32_t mor_calc_forces(OrSystem *s); /* clean + update_chunks + calc_forces (+rings) + walls forces */
int32_t mor_bin(OrSystem *s);
int32_t mor_download_cells(OrSystem *s, int32_t *cell_of_particle, int32_t *counts);
int32_t mor_download_cell_lists(OrSystem *s, int32_t *start, int32_t *ids);
/* the reference's half stencil of a cell, in the reference's order (src/chunks.jl:61-118) */
int32_t mor_cell_neighbors(OrSystem *s, int32_t cell, int32_t *out4, int32_t *n);
int64_t mor_chunk_capacity(OrSystem *s); /* nc of src/chunks.jl:32-35 */

int32_t mor_energies(OrSystem *s, int32_t pe_mode, double *ke, double *pe);
int32_t mor_rings_download_info(OrSystem *s, double *areas, double *cms, double *cont_pos);
int32_t mor_rings_set_neighbors(OrSystem *s, int32_t mode, int32_t type_all, double tol);
int32_t mor_rings_download_neighbors(OrSystem *s, int32_t *count, int32_t *list);
/* sources / sinks / variable ring count (src/rings/sources.jl, src/rings/states.jl:173-227); before mor_upload_state */
int32_t mor_rings_set_sources(OrSystem *s, const MaviSourceSink *list, int32_t n, const uint8_t *ring_active,
                              const double *spawn_draws, int64_t n_draws);
int32_t mor_rings_download_active(OrSystem *s, uint8_t *ring_active, int64_t *uids, int64_t *num_active);
/* InvasionsCfg(steps_to_update) + RingsIntCfg(r_chunks_cfg) (src/rings/configs.jl:334-352); r_cols = 0: no ring chunks */
int32_t mor_rings_set_invasions(OrSystem *s, int32_t steps_to_update, int32_t r_cols, int32_t r_rows);
int32_t mor_rings_download_invasions(OrSystem *s, int64_t *n, int32_t *triples, int64_t cap);
int32_t mor_get_time(OrSystem *s, int64_t *num_steps, double *time);

/* fine-grained operators for unit tests (each is one reference function) */
void mor_clean_forces(OrSystem *s);          /* src/systems.jl:119-123 */
int32_t mor_update_chunks(OrSystem *s);      /* src/integration.jl:54-59 */
void mor_pair_forces(OrSystem *s);           /* calc_forces!(system) only */
void mor_walls_forces(OrSystem *s);          /* src/integration.jl:228-266 */
void mor_walls(OrSystem *s);                 /* src/integration.jl:268-412 */
void mor_update_verlet(OrSystem *s);         /* src/integration.jl:415-431 */

/* scalar helpers */
double mor_julia_div(double x, double y);    /* Base.div(::Float64, ::Float64) */
void mor_calc_diff(const OrSystem *s, const double *r1, const double *r2, double *dr);
void mor_potential_force(int32_t kind, const double *par, const double *dr, double dist, double *f);
void mor_szabo_interaction(const double *par, const double *dr, double *f);
void mor_rtp_interaction(const double *par, const double *dr, double *f);

#ifdef __cplusplus
}
#endif
#endif
