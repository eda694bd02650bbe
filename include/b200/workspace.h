/* workspace.h -- the scratch arena a b200_ctx lends to the kernel launchers.
 *
 * Replaces the per-call mem_t scratch of the reference (count: advance.hxx:30,
 * scan partials: kernel_scan.hxx:30, merge-path partitions: search.hxx:19,
 * compact bits/cta_offsets + pinned scalar: kernel_compact.hxx:33-34,75): all of
 * it is allocated once here, so an operator call performs no allocation.
 * Plain C so it can cross the C ABI (b200_ctx_workspace). */
#ifndef B200_WORKSPACE_H
#define B200_WORKSPACE_H
#include <stdint.h>

enum {
    B200_CNT_TOTAL = 0,     /* scan total (m_F) */
    B200_CNT_OUT = 1,       /* items appended to the output frontier */
    B200_CNT_OVERFLOW = 2,  /* nonzero if an append ran past the capacity */
    B200_CNT_ARCS = 3,      /* arcs inspected (pull) */
    B200_CNT_AUX = 4,       /* next frontier's degree sum (fused), etc. */
    B200_CNT_AUX2 = 5,
    B200_CNT_WORK = 6,      /* dynamic work claim of a kernel (pull: next bitmap chunk); zero at launch */
    B200_NUM_COUNTERS = 8
};

typedef struct b200_workspace {
    void *stream;                       /* cudaStream_t every launch goes to */
    int32_t device;
    int32_t num_sms;
    unsigned long long *d_status;       /* decoupled look-back tile status words */
    int64_t status_tiles;               /* capacity of d_status in tiles */
    unsigned int *d_tile_counter;       /* dynamic tile id (self-resetting) */
    unsigned int epoch;                 /* tags d_status entries; bumped per scan launch */
    unsigned int reserved;
    unsigned long long *d_counters;     /* [B200_NUM_COUNTERS] device counters */
    unsigned long long *h_counters;     /* pinned host mirror, filled by b200::read_counters */
    uint32_t *d_scanned;                /* exclusive scan of frontier degrees: the role of
                                           graph_device_t::d_scanned_row_offsets (graph.hxx:52) */
    int64_t scanned_capacity;           /* in items */
    int64_t launches;                   /* kernels launched through this workspace (bench `gpu_launches`) */
    void *d_rows;                       /* uint2[scanned_capacity]: (row begin, row end) of every frontier
                                           vertex, written by the quad scan beside d_scanned */
    uint32_t *d_scanned2;               /* second (scan, rows) pair: the work-creating advance writes the NEXT level's */
    void *d_rows2;                      /* scan and row bounds while it reads this level's (quad_advance.cuh) */
} b200_workspace;

#ifdef __cplusplus
extern "C" {
#endif
struct b200_ctx;
/* The ctx-owned workspace (mutable: launchers bump epoch / launches). */
b200_workspace *b200_ctx_workspace(struct b200_ctx *ctx);
#ifdef __cplusplus
}
#endif
#endif
