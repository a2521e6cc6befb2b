// Group-resident fused back end (fused.cu): one CTA per digitisation group, photons -> raw_records.
#pragma once
#include "backend.cuh"

namespace wfs {

constexpr int kFusedThreads = 256;
constexpr int kFusedMaxPhotons = 8192;      // photons of one group (13 index bits in the key)
constexpr int kFusedTile = 512;             // samples of a warp's tile: 4 records (440 samples) fit
constexpr int kFusedRecCap = 4096;          // records of one group ordered in shared memory
constexpr int kFusedItvCap = 2048;          // ZLE intervals of one group

enum FusedScalar { FS_NVALID = 0, FS_NPULSES, FS_NWIN, FS_NITV, FS_NSAMPLES, FS_NREC, FS_ERR, FS_OVERFLOW, FS_COUNT };

struct FusedArgs {
    PhotonBatch b;
    DeviceConfig c;
    int n_cap;                      // photon capacity of the shared-memory key array
    int relpc_bits;
    const int64_t *group_t0;        // [n_groups] lower bound of the photon times of the group [ns]
    const int32_t *group_run0;      // [n_groups] first Pulse call (run) of the group
    uint64_t *status;               // [n_groups] look-back words, zeroed
    uint32_t *ticket;               // group counter, zeroed
    int64_t *scalars;               // [FS_COUNT], zeroed
    uint32_t *group_nitv;           // unused (kept for layout stability)
    uint8_t *records_out;
    int64_t cap_records;
    wfs_group_info *group_info;
};

}  // namespace wfs
