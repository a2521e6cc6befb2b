// Sampling front end (S1/S2/afterpulse photon generation) -- placeholder until the Philox
// kernels land; the deterministic entry wfs_simulate_photons does not need it.
#include "handle.cuh"

namespace wfs {
struct Frontend {};
void Handle::frontend_init(const wfs_tables &) { frontend = nullptr; }
void Handle::frontend_release() {}
}  // namespace wfs

using namespace wfs;

extern "C" {
int wfs_simulate(void *handle, const uint8_t *, int64_t, const wfs_instr_maps *, uint64_t, uint8_t *,
                 int64_t, uint8_t *, int64_t, wfs_group_info *, int64_t, wfs_counts *) {
    if (handle) reinterpret_cast<Handle *>(handle)->last_error = "wfs_simulate: not built yet";
    return WFS_E_ARG;
}
int wfs_stage_instructions(void *handle, const uint8_t *, int64_t, const wfs_instr_maps *) {
    if (handle) reinterpret_cast<Handle *>(handle)->last_error = "not built yet";
    return WFS_E_ARG;
}
int wfs_run_staged(void *handle, uint64_t, wfs_counts *) {
    if (handle) reinterpret_cast<Handle *>(handle)->last_error = "not built yet";
    return WFS_E_ARG;
}
int wfs_sample_stage(void *handle, int, const uint8_t *, int64_t, const wfs_instr_maps *, uint64_t,
                     void *, int64_t, int64_t *) {
    if (handle) reinterpret_cast<Handle *>(handle)->last_error = "not built yet";
    return WFS_E_ARG;
}
}
