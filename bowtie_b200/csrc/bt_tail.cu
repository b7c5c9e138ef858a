/*
 * bt_tail.cu — the round-robin tail of the DFS search path (bt_tail_kernel).
 *
 * Search cost per read is heavy-tailed (hg19-sized index, `-n 2`: mean 880 transitions, 0.7 % of the reads above 8 000, the
 * longest around 10^6, each one a strictly sequential chain — every backtrack consumes the read's random stream).  The main
 * pass (bt_search_kernel, bt_lib.cu) therefore suspends a read that exceeds its transition budget into a checkpoint slot
 * (bt_ctxq.cuh: 39 packed state words + the slot's own scratch).  This kernel finishes those reads:
 *
 *   - the suspended reads circulate through ONE ring of slot ids (BtTailQ).  A lane pops a slot, resumes the read, runs it for a
 *     quantum of transitions, and — if it is still unfinished — stores its 39 words and pushes the slot back: round robin, so a
 *     warp's lanes always hold reads as long as the ring has any, whatever the individual reads' lengths;
 *   - a warp may hold at most min(cap, live - cap x its index) reads, where `live` counts the unfinished ones and cap = live / wtarget
 *     clamped to [mincap, 32]: as the reads run out the warps with the highest indices hand their reads back and exit, and what is
 *     left always sits in about min(wtarget, live / mincap) warps.  Blocks are single warps so that an exiting warp returns its
 *     registers and shared memory at once.
 *
 * The straggler problem this replaces (one kernel for the whole tail, one read per lane until it ends): the last reads of a
 * batch run alone in their warps for ~0.7 s — measured 47 G warp instructions at 3.6 active threads for the tail of a
 * 1 M-read batch against 16 G at 7.4 for its main pass (profiles/README.md, call 7).
 *
 * A read migrates between SMs here within one kernel, and an SM's L1 is not coherent with another SM's stores: this translation
 * unit is compiled with -Xptxas -dlcm=cg (every plain global load goes to L2).  The index is read through ld.global.nc as
 * everywhere else (read-only data).  Suspending: state words, then __threadfence(), then the ring cell.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "bt_kernel_cfg.cuh"

#ifndef BT_TAIL_MIN_BLOCKS
#define BT_TAIL_MIN_BLOCKS 16      /* single-warp blocks per SM: 128 registers per thread */
#endif

__global__ void bt_tail_init_kernel(BtTailQ *q, const unsigned long long *count, uint32_t nslot, uint32_t cap, uint32_t *items, uint32_t quantum, uint32_t wtarget, uint32_t mincap) {
	unsigned long long n = *count;                                  /* the main pass counts past the last slot (those reads are re-run) */
	if (n > nslot) n = nslot;
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < cap; k += stride) items[k] = k < n ? k : BT_TAILQ_EMPTY;
	if (blockIdx.x == 0 && threadIdx.x == 0) { q->head = 0; q->tail = n; q->live = (long long)n; q->cap_mask = cap - 1; q->quantum = quantum; q->wtarget = wtarget; q->mincap = mincap; q->items = items; }
}

__global__ void __launch_bounds__(32, BT_TAIL_MIN_BLOCKS)
bt_tail_kernel(BtKParams P, BtTailQ *Q) {
	extern __shared__ __align__(16) uint8_t bt_smem[];
	const uint32_t lane = threadIdx.x, wid = blockIdx.x;
	uint8_t *const my_stage = bt_smem + (size_t)lane * BT_SMEM_STRIDE;
	BtScratch S;
	S.rows = nullptr; S.elims = nullptr; S.frames = nullptr; S.partials = nullptr;          /* a lane works in its slot's scratch */
	uint32_t my_slot = 0, my_budget = 0;
	BtLane L;
#if BT_COLD_SMEM
	L.K = reinterpret_cast<BtLaneCold *>(my_stage + BT_SMEM_COLD);
#else
	BtLaneCold cold_regs; L.K = &cold_regs;
#endif
	L.pc = PC_NEXT_READ;
	L.s_lfex = L.s_lf = L.s_chase = L.K->s_ftab = L.K->s_offs = L.K->s_bt = L.s_iter = L.s_blk = 0;
	L.K->nmuts = 0; L.K->mut0 = L.K->mut1 = L.K->mut2 = 0; L.ebwtSel = 0; L.lfk = 0; L.ltop = L.lbot = L.crow = 0; L.flags = 0; L.d = 0; L.qlen = 0;
	L.rlen = 0; L.rseq = my_stage; L.rqual = my_stage; L.K->hasN = 1; L.K->step = 0; L.nit = 0;
	volatile uint32_t *const items = Q->items;
	const uint32_t cap_mask = Q->cap_mask, quantum = Q->quantum, wtarget = Q->wtarget ? Q->wtarget : 1u, mincap = Q->mincap ? Q->mincap : 1u;
	uint32_t *const snap = reinterpret_cast<uint32_t *>(my_stage + BT_SMEM_SNAP);
	for (;;) {
		if (__ballot_sync(0xffffffffu, L.pc != PC_EXIT) == 0) break;
		if (BT_IS_RARE_STEP(L.pc)) bt_rare_iter(L, P, S, my_budget);
		/* reads that ended */
		const bool fin = (L.pc == PC_FINISH_READ);
		if (fin) {
			if (L.flags & BT_FLAG_RETRY) {                              /* its slot's scratch overflowed: re-run by the overflow pass, this turn's operations are not counted */
				L.s_lfex = snap[0]; L.s_lf = snap[1]; L.s_chase = snap[2]; L.K->s_ftab = snap[3]; L.K->s_offs = snap[4]; L.s_blk = snap[5];
			}
			bt_finish_read(L, P); L.pc = PC_NEXT_READ;
		}
		const unsigned finm = __ballot_sync(0xffffffffu, fin);
		/* reads whose turn is over: back into the ring */
		const bool pre = (L.flags & BT_FLAG_PREEMPT) != 0;
		if (pre) { L.flags &= ~BT_FLAG_PREEMPT; bt_ctx_store(L, P.slot_ctx, P.nslot, my_slot); __threadfence(); }
		const unsigned prem = __ballot_sync(0xffffffffu, pre);
		if (prem) {
			unsigned long long base = 0;
			if (lane == 0) base = atomicAdd(&Q->tail, (unsigned long long)__popc(prem));
			base = __shfl_sync(0xffffffffu, base, 0);
			if (pre) { items[(uint32_t)(base + __popc(prem & ((1u << lane) - 1u))) & cap_mask] = my_slot; L.pc = PC_NEXT_READ; }
		}
		if (finm && lane == 0) atomicAdd(reinterpret_cast<unsigned long long *>(&Q->live), (unsigned long long)(-(long long)__popc(finm)));
		/* lanes without a read */
		const bool want = (L.pc == PC_NEXT_READ);
		const unsigned wm = __ballot_sync(0xffffffffu, want);
		if (wm) {
			long long live = 0;
			if (lane == 0) live = *reinterpret_cast<volatile long long *>(&Q->live);
			live = __shfl_sync(0xffffffffu, live, 0);
			const uint32_t hold = (uint32_t)__popc(__ballot_sync(0xffffffffu, L.pc != PC_NEXT_READ && L.pc != PC_EXIT));
			/* reads per warp: as many as spread the live reads over `wtarget` warps, between `mincap` and 32 — packed while there are many
			 * (throughput), thinning out towards the end, where the longest reads set the batch's latency and a full warp advances
			 * each of its reads several times more slowly than a nearly empty one */
			long long cap = (live + (long long)wtarget - 1) / (long long)wtarget;
			cap = cap < (long long)mincap ? (long long)mincap : (cap > 32 ? 32 : cap);
			const long long al = live - cap * (long long)wid;
			const uint32_t allowed = al <= 0 ? 0u : (al > cap ? (uint32_t)cap : (uint32_t)al);
			uint32_t take = 0;
			unsigned long long h = 0;
			if (live <= 0 || (allowed == 0 && hold == 0)) { if (want) L.pc = PC_EXIT; }        /* nothing left, or this warp's share is gone: leave */
			else if (allowed > hold) {
				const uint32_t room = allowed - hold, nw = (uint32_t)__popc(wm), need = nw < room ? nw : room;
				if (lane == 0) {
					for (int tries = 0; tries < 8; tries++) {
						h = *reinterpret_cast<volatile unsigned long long *>(&Q->head);
						const unsigned long long t = *reinterpret_cast<volatile unsigned long long *>(&Q->tail);
						if (t <= h) break;
						const unsigned long long avail = t - h, k = avail < need ? avail : need;
						if (atomicCAS(&Q->head, h, h + k) == h) { take = (uint32_t)k; break; }
					}
				}
				take = __shfl_sync(0xffffffffu, take, 0);
				h = __shfl_sync(0xffffffffu, h, 0);
				const uint32_t rank = (uint32_t)__popc(wm & ((1u << lane) - 1u));
				if (want && rank < take) {
					const uint32_t idx = (uint32_t)(h + rank) & cap_mask;
					uint32_t v;
					while ((v = items[idx]) == BT_TAILQ_EMPTY) { }                     /* its producer has reserved the cell and is about to fill it */
					items[idx] = BT_TAILQ_EMPTY;
					__threadfence();
					my_slot = v;
					bt_slot_resume(L, P, S, v);
					my_budget = L.nit + quantum;
					snap[0] = L.s_lfex; snap[1] = L.s_lf; snap[2] = L.s_chase; snap[3] = L.K->s_ftab; snap[4] = L.K->s_offs; snap[5] = L.s_blk;
				}
			}
			if (hold == 0 && take == 0 && live > 0) __nanosleep(2000);                   /* an idle warp that may still be needed: poll gently */
		}
#pragma unroll 1
		for (uint32_t k = 0; k < P.rare_period; k++) {
			const bool fast = BT_IS_FAST(L.pc);
			const unsigned fmask = __ballot_sync(0xffffffffu, fast);
			const unsigned rmask = __ballot_sync(0xffffffffu, !fast && L.pc != PC_EXIT && L.pc != PC_NEXT_READ);
			if (fmask == 0 || (uint32_t)__popc(rmask) >= P.rare_thresh) break;
			if (fast) bt_fast_iter(L, P, S);
		}
	}
	unsigned long long v[8] = { L.s_lfex, L.s_lf, L.s_chase, L.K->s_ftab, L.K->s_offs, L.K->s_bt, L.s_iter, L.s_blk };
#pragma unroll
	for (int k = 0; k < 8; k++) {
		unsigned long long x = v[k];
		for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
		if (lane == 0 && x) atomicAdd(&P.stats[k], x);
	}
}

/* Enqueues the tail of one batch on `st`: ring set-up from the main pass's slot count, then the kernel.  Returns the CUDA error code. */
int bt_tail_launch(const BtKParams &P, BtTailQ *q, const unsigned long long *count, uint32_t nslot, uint32_t cap, uint32_t *items, uint32_t quantum,
                   uint32_t wtarget, uint32_t mincap, uint32_t blocks, cudaStream_t st) {
	bt_tail_init_kernel<<<64, 256, 0, st>>>(q, count, nslot, cap, items, quantum, wtarget, mincap);
	bt_tail_kernel<<<blocks, 32, 32 * BT_SMEM_STRIDE, st>>>(P, q);
	return (int)cudaGetLastError();
}
