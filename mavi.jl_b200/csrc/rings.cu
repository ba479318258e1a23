// rings.cu — Mavi.Rings on the device (placeholder until the ring kernels land).
#include "handle.cuh"

namespace mavi {
static int unsupported(Handle *h) {
  h->set_error("Mavi.Rings kernels are not built in this version");
  return MAVI_ERR_UNSUPPORTED;
}
int rings_lower(Handle *h, const MaviParams *) { return unsupported(h); }
int rings_allocate(Handle *h) { return unsupported(h); }
int rings_upload_finish(Handle *h) { return unsupported(h); }
int rings_step(Handle *h, const double *) { return unsupported(h); }
int rings_calc_forces(Handle *h) { return unsupported(h); }
int rings_download_info(Handle *h, void *, void *, void *) { return unsupported(h); }
int rings_download_state(Handle *h, void *, void *) { return unsupported(h); }
int rings_download_forces(Handle *h, void *) { return unsupported(h); }
int rings_bin(Handle *h) { return unsupported(h); }
int rings_download_cells(Handle *h, int *, int *, int *, int *) { return unsupported(h); }
}  // namespace mavi
