/*
 * bt_kernel_cfg.cuh — launch configuration and shared-memory layout of the search kernels, shared by bt_lib.cu (bt_search_kernel,
 * the main / restart / overflow passes) and bt_tail.cu (bt_tail_kernel, the round-robin tail).
 */
#pragma once
#include "bt_ctxq.cuh"

#ifndef BT_THREADS
#define BT_THREADS 128
#endif
#ifndef BT_MIN_BLOCKS
#define BT_MIN_BLOCKS 4            /* register cap = 65536 / (128 * BT_MIN_BLOCKS): 128 registers with the cold lane state in shared memory */
#endif
#ifndef BT_RARE_PERIOD
#define BT_RARE_PERIOD 8           /* rare transitions run at least every BT_RARE_PERIOD-th iteration ... */
#endif
#ifndef BT_RARE_THRESH
#define BT_RARE_THRESH 16          /* ... or as soon as this many lanes of the warp wait for one (8 / 16: sweep on the hg19-sized index, profiles/) */
#endif
#ifndef BT_Q_NCTX
#define BT_Q_NCTX 1024              /* read contexts per block of the queue-driven kernel                    */
#endif
#ifndef BT_Q_THREADS
#define BT_Q_THREADS 384            /* worker threads per block                                               */
#endif
#define BT_SMEM_LEN 128            /* reads up to this length are staged in shared memory                   */
#ifndef BT_COLD_SMEM
#define BT_COLD_SMEM 1             /* the rare transitions' part of the lane state (BtLaneCold) lives in shared memory, not in registers */
#endif
/* A lane's shared-memory area: its writable copy of the read's bases (seedling mutations are applied to it), the 6 operation counters
 * snapped when its current read began, and — BT_COLD_SMEM — its BtLaneCold.  Qualities are never written, so they are read in place
 * from the batch (L1-resident: 100 bytes per read, fetched a position ahead of their use).  The stride is an odd number of words:
 * lanes' equal offsets fall in different banks. */
#define BT_SMEM_SNAP BT_SMEM_LEN
#define BT_SMEM_COLD (BT_SMEM_LEN + 24)
#if BT_COLD_SMEM
#define BT_SMEM_STRIDE ((BT_SMEM_LEN + 24 + (uint32_t)sizeof(BtLaneCold) + 4) | 4u)
#else
#define BT_SMEM_STRIDE (BT_SMEM_LEN + 28)
#endif
static_assert((BT_SMEM_STRIDE / 4) % 2 == 1 && BT_SMEM_STRIDE % 4 == 0, "odd word stride");

struct BtWorkCtl { unsigned long long next; unsigned long long nwork; };

/* The round-robin tail (bt_tail.cu): the reads the main pass suspended (checkpoint slots, bt_ctxq.cuh) circulate through one ring of
 * slot ids.  `tail` runs ahead of the item stores — a consumer that reserved an index waits until its cell stops holding BT_TAILQ_EMPTY. */
#define BT_TAILQ_EMPTY 0xffffffffu
struct BtTailQ {
	unsigned long long head, tail;     /* items [head, tail) are queued                                          */
	long long live;                    /* unfinished reads: queued, or held by a lane                            */
	uint32_t cap_mask, quantum;        /* ring capacity - 1 (power of two >= 2 x slots); transitions per turn    */
	uint32_t wtarget, mincap;          /* warps the live reads are spread over; fewest reads a warp is filled to  */
	uint32_t *items;
};
int bt_tail_launch(const BtKParams &P, BtTailQ *q, const unsigned long long *count, uint32_t nslot, uint32_t cap, uint32_t *items, uint32_t quantum,
                   uint32_t wtarget, uint32_t mincap, uint32_t blocks, cudaStream_t st);
