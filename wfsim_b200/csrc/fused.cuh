// Group-resident fused back end (fused.cu): one CTA per digitisation group, photons -> raw_records.
#pragma once
#include "backend.cuh"

namespace wfs {

constexpr int kFusedThreads = 1024;
constexpr int kFusedSmallThreads = 512;     // classes up to this many threads per CTA run the 3-CTAs-per-SM build
constexpr int kFusedMaxPhotons = 8192;      // photons of one group (13 index bits in the key)
constexpr int kFusedMaxRecCap = 8192;       // records of one group ordered in shared memory (13 bits of record slot)
constexpr int kFusedTrigSlots = 64;         // (pulse call, total / bottom) trigger counters kept in shared memory

constexpr int kFusedMaxClasses = 6;         // groups are binned by photon count
constexpr int kFusedRecordThreads = 256;

enum FusedScalar { FS_NVALID = 0, FS_NPULSES, FS_NWIN, FS_NITV, FS_NSAMPLES, FS_NREC, FS_ERR, FS_OVERFLOW, FS_COUNT };

struct FusedArgs {
    PhotonBatch b;
    DeviceConfig c;
    int relpc_bits;
    const int64_t *group_t0;        // [n_groups] lower bound of the photon times of the group [ns]
    const int32_t *group_run0;      // [n_groups] first Pulse call (run) of the group
    int64_t *scalars;               // [FS_COUNT], zeroed; FS_NREC doubles as the bump allocator of the descriptor pool
    uint32_t *over_count;           // groups that outgrew the lists of their class
    uint32_t *group_nrec;           // [n_groups + 1] records of every group (zeroed)
    uint32_t *group_desc;           // [n_groups] first descriptor of the group in the pool
    const uint32_t *rec_base;       // [n_groups + 1] exclusive scan of group_nrec
    uint32_t *tkey;                 // [photons] channel-ordered photons of every group: time | samples owned
    uint4 *adc_slots;               // [photons x 4] ADC values of the samples every photon owns (64-byte slot per photon)
    uint4 *desc;                    // [cap_records] record descriptors
    uint8_t *records_out;
    int64_t cap_records;
    wfs_group_info *group_info;
};

struct FusedClass {                 // one launch of k_group_analyse: the groups of one size class
    int n_cap, itv_cap, rec_cap;    // photons / intervals / records of a group held in shared memory
    int bin_bits;                   // time bins of the record order: 2^bin_bits per octave
    const uint32_t *list;           // group ids
    uint32_t n_list;
    uint32_t *ticket;
    uint32_t *overflow_list;        // nullptr: an overflow fails the batch (multi-pass back end)
};

}  // namespace wfs
