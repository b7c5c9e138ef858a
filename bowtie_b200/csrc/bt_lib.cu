/*
 * bt_lib.cu — kernels and C ABI of libbowtie_b200.so (sm_100a).
 *
 *   bt_relayout_kernel   .ebwt sides -> 32-byte rank blocks (once per index load)
 *   bt_search_kernel     persistent lanes, one read per thread, dynamic work queue
 *   bt_collect_kernel    device-side list of reads whose scratch overflowed (for the retry pass)
 *   bt_best_kernel       the best-first ("stateful") path: one read per thread, per-read arena (bt_best.cuh)
 *
 * There is no host search path in this library: without a CUDA device every entry point fails.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <mutex>

#include "bt_native.cuh"
#include "bt_ctxq.cuh"
#include "bt_best_prog.h"
#include "bt_ref_load.h"
#include "../../include/bowtie_b200.h"

static_assert(sizeof(bt_policy_t) == sizeof(BtPolicy), "bt_policy_t and BtPolicy must share a layout");
static_assert(BT_HIT_HDR_WORDS == BT_HIT_HDR, "hit header size");
static_assert(sizeof(BtFrame) == 64, "BtFrame layout");

/* ------------------------------------------------------------------------------------------- */
/* kernels                                                                                      */
/* ------------------------------------------------------------------------------------------- */

__global__ void bt_relayout_kernel(BtNativeIndex n, uint32_t nblocks, uint4 *out) {
	uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= nblocks) return;
	uint4 b[2];
	bt_relayout_block(n, k, b);
	out[2 * (size_t)k] = b[0];
	out[2 * (size_t)k + 1] = b[1];
}

__global__ void bt_debug_lf_kernel(BtDevIndex ix, const uint32_t *rows, uint32_t n, uint32_t *out) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t row = rows[i], ex[4];
	BtBlock b = bt_load_block(ix, row);
	bt_lf_ex(ix, b, row, ex);
	for (int c = 0; c < 4; c++) {
		out[5 * (size_t)i + c] = ex[c];
		/* the single-character path must agree with the quartet */
		if (bt_lf(ix, b, row, c) != ex[c]) out[5 * (size_t)i + c] = 0xdeadbeefu;
	}
	out[5 * (size_t)i + 4] = bt_row_l(b, row);
}

#include "bt_kernel_cfg.cuh"

/* Persistent search kernel: every thread is a lane that pulls read ids from a global cursor until
 * the batch is exhausted.  `ctl->nwork` is read from device memory so that the retry pass can be
 * enqueued before its size is known on the host.
 *
 * Divergence control: the LF step and the row-chase step ("fast" transitions, > 85 % of all transitions)
 * run every iteration; every other transition ("rare": phase changes, backtrack selection, frame push/pop,
 * hit resolution, fetching the next read) is deferred until enough lanes of the warp wait for one, so that
 * their long, serialised code paths are paid once for many lanes instead of once per iteration. */
__global__ void __launch_bounds__(BT_THREADS, BT_MIN_BLOCKS)
bt_search_kernel(BtKParams P, BtWorkCtl *ctl) {
	extern __shared__ __align__(16) uint8_t bt_smem[];
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t lane = threadIdx.x & 31;
	uint8_t *const my_stage = bt_smem + (size_t)threadIdx.x * BT_SMEM_STRIDE;
	BtScratch S;
	S.rows = P.rows + (size_t)tid * P.R * 2;
	S.elims = P.elims + (size_t)tid * P.R;
	S.frames = P.frames + (size_t)tid * P.FCAP;
	S.partials = P.partials + (size_t)tid * P.PCAP;
	BtLane L;
#if BT_COLD_SMEM
	L.K = reinterpret_cast<BtLaneCold *>(my_stage + BT_SMEM_COLD);
#else
	BtLaneCold cold_regs; L.K = &cold_regs;
#endif
	L.pc = PC_NEXT_READ;
	L.s_lfex = L.s_lf = L.s_chase = L.K->s_ftab = L.K->s_offs = L.K->s_bt = L.s_iter = L.s_blk = 0;
	L.K->nmuts = 0; L.K->mut0 = L.K->mut1 = L.K->mut2 = 0; L.ebwtSel = 0; L.lfk = 0; L.ltop = L.lbot = L.crow = 0; L.flags = 0; L.d = 0; L.qlen = 0;
	L.rlen = 0; L.rseq = my_stage; L.rqual = my_stage; L.K->hasN = 1; L.K->step = 0;
	const unsigned long long nwork = ctl->nwork;
	/* Once the work queue is empty a pass only waits for its slowest reads while most lanes idle — with a per-read budget of 8000
	 * transitions that drain is as long as everything a lane did before it at a million reads per pass.  From the moment a warp
	 * finds the queue empty its lanes therefore run on the (smaller) drain budget: what exceeds it moves to the tail pass, where it
	 * overlaps the next batch instead of holding this one's blocks. */
	uint32_t budget = P.budget;
#ifndef BT_INNER_FAST
#define BT_INNER_FAST 1
#endif
#if BT_INNER_FAST
	/* Two nested loops: the outer one runs a rare pass, the inner one up to `rare_period` fast iterations — left early when no lane
	 * has fast work or `rare_thresh` lanes wait.  The inner loop's body is the fast transition alone, so its back edge merges only
	 * the fast path's register assignment (the single-loop form paid ~100 register moves per iteration at the merge with the rare code). */
	for (;;) {
		{
			const bool rare0 = !BT_IS_FAST(L.pc) && L.pc != PC_EXIT;
			if (__ballot_sync(0xffffffffu, BT_IS_FAST(L.pc) || rare0) == 0) break;         /* every lane has exited */
		}
		{
#else
	uint32_t it = 0;
	for (;; it++) {
		const bool fast = BT_IS_FAST(L.pc);
		const bool rare = !fast && L.pc != PC_EXIT;
		const unsigned fmask = __ballot_sync(0xffffffffu, fast);
		const unsigned rmask = __ballot_sync(0xffffffffu, rare);
		if ((fmask | rmask) == 0) break;                                  /* every lane has exited */
		const bool run_rare = (fmask == 0) || ((uint32_t)__popc(rmask) >= P.rare_thresh) || ((it % P.rare_period) == 0);
		if (run_rare) {
#endif
			if (L.pc == PC_FINISH_READ) {
				if (L.flags & BT_FLAG_RETRY) {
					/* this read is re-run from scratch by a later pass: its operations so far are not part of the algorithm's
					 * count (SURVEY.md §8d counts each read's side fetches once) */
					const uint32_t *snap = reinterpret_cast<const uint32_t *>(my_stage + BT_SMEM_SNAP);
					L.s_lfex = snap[0]; L.s_lf = snap[1]; L.s_chase = snap[2]; L.K->s_ftab = snap[3]; L.K->s_offs = snap[4]; L.s_blk = snap[5];
				}
				bt_finish_read(L, P); L.pc = PC_NEXT_READ;
			}
			/* work distribution: warp-aggregated grab from the global cursor, then the warp copies each new
			 * read into the owning lane's shared-memory stage with coalesced loads */
			const bool want = (L.pc == PC_NEXT_READ);
			const unsigned wmask = __ballot_sync(0xffffffffu, want);
			if (wmask) {
				const int leader = __ffs(wmask) - 1;
				unsigned long long base = 0;
				if ((int)lane == leader) base = atomicAdd(&ctl->next, (unsigned long long)__popc(wmask));
				base = __shfl_sync(0xffffffffu, base, leader);
				bool got = false, took = false;
				unsigned long long ro = 0;
				if (want) {
					unsigned long long w = base + (unsigned long long)__popc(wmask & ((1u << lane) - 1u));
					if (w < nwork) {
						const uint32_t rid = P.sel ? P.sel[w] : (uint32_t)w;
						took = true;
						bt_begin_read(L, P, rid);
						ro = P.roff[rid];
						got = true;
						uint32_t *snap = reinterpret_cast<uint32_t *>(my_stage + BT_SMEM_SNAP);
						snap[0] = L.s_lfex; snap[1] = L.s_lf; snap[2] = L.s_chase; snap[3] = L.K->s_ftab; snap[4] = L.K->s_offs; snap[5] = L.s_blk;
					} else L.pc = PC_EXIT;
				}
				if (P.drain_budget && budget > P.drain_budget && __ballot_sync(0xffffffffu, want && !took)) budget = P.drain_budget;
				unsigned gmask = __ballot_sync(0xffffffffu, got && L.rlen <= BT_SMEM_LEN);
				while (gmask) {
					const int j = __ffs(gmask) - 1;
					gmask &= gmask - 1;
					const uint32_t rl = __shfl_sync(0xffffffffu, L.rlen, j);
					const unsigned long long rj = __shfl_sync(0xffffffffu, ro, j);
					uint8_t *dst = bt_smem + (size_t)((threadIdx.x & ~31u) + j) * BT_SMEM_STRIDE;
					bool sawN = false;
					for (uint32_t k = 0; k < rl; k += 32) {
						const uint32_t idx = k + lane;
						if (idx < rl) {
							const uint8_t b = __ldg(P.seq + rj + idx);
							dst[idx] = b;
							sawN |= (b == 4);
						}
					}
					const unsigned nm = __ballot_sync(0xffffffffu, sawN);
					if ((int)lane == j) { L.rseq = my_stage; L.rqual = const_cast<uint8_t *>(P.qual + rj); L.K->hasN = (nm != 0); }
				}
				if (got && L.rlen > BT_SMEM_LEN) {
					/* long read: private copy in global scratch */
					uint8_t *dst = P.stage + (size_t)tid * 2 * P.stage_len;
					bool sawN = false;
					for (uint32_t k = 0; k < L.rlen; k++) {
						const uint8_t b = __ldg(P.seq + ro + k);
						dst[k] = b; dst[P.stage_len + k] = __ldg(P.qual + ro + k);
						sawN |= (b == 4);
					}
					L.rseq = dst; L.rqual = dst + P.stage_len; L.K->hasN = sawN;
				}
				__syncwarp();
			}
			if (BT_IS_RARE_STEP(L.pc)) bt_rare_iter(L, P, S, budget);
			if (L.flags & BT_FLAG_PREEMPT) {
				/* over this pass's budget (or out of seedling space): suspend the read into a checkpoint slot; the round-robin tail resumes it */
				L.flags &= ~BT_FLAG_PREEMPT;
				const unsigned long long p = atomicAdd(P.slot_count, 1ull);
				if (p < P.nslot) { bt_slot_save_new(L, P, S, (uint32_t)p); P.flags[L.K->rid] = 0; P.found[L.K->rid] = 0; L.pc = PC_NEXT_READ; }
				else { L.flags |= BT_FLAG_BUDGET; L.pc = PC_FINISH_READ; }      /* no slot left: re-run from scratch by the overflow pass */
			}
		}
#if BT_INNER_FAST
#pragma unroll 1
		for (uint32_t k = 0; k < P.rare_period; k++) {
			const bool fast = BT_IS_FAST(L.pc);
			const unsigned fmask = __ballot_sync(0xffffffffu, fast);
			const unsigned rmask = __ballot_sync(0xffffffffu, !fast && L.pc != PC_EXIT);
			if (fmask == 0 || (uint32_t)__popc(rmask) >= P.rare_thresh) break;
			if (fast) bt_fast_iter(L, P, S);
		}
#else
		if (fast) bt_fast_iter(L, P, S);
#endif
	}
	/* statistics: warp-reduce, one atomic per warp and counter */
	unsigned long long v[8] = { L.s_lfex, L.s_lf, L.s_chase, L.K->s_ftab, L.K->s_offs, L.K->s_bt, L.s_iter, L.s_blk };
#pragma unroll
	for (int k = 0; k < 8; k++) {
		unsigned long long x = v[k];
		for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
		if (lane == 0 && x) atomicAdd(&P.stats[k], x);
	}
}

/* ------------------------------------------------------------------------------------------- */
/* Queue-driven search kernel (v5): contexts in shared memory, state-homogeneous warps.            */
/* ------------------------------------------------------------------------------------------- */
#define BT_QCAP 1024                /* ring capacity (power of two) >= contexts per block           */
#define BT_NQ 8                     /* work queues, one per class of state                          */
enum { QF = 0, QC = 1, QPH = 2, QPOS = 3, QBT = 4, QRET = 5, QREP = 6, QIO = 7 };

struct BtQueues { uint32_t head[BT_NQ], tail[BT_NQ], live, pad; uint32_t item[BT_NQ][BT_QCAP]; };

/* Classes group the states whose code a warp can execute together:
 *   QF   one query position (LF step)            QC   one locate step
 *   QPH  phase program, backtrack() entry/exit   QPOS frame entry, position prologue with its own loads
 *   QBT  backtrack-target selection + push       QRET frame pop, bookkeeping after a recursive call
 *   QREP reporting (alignment, rows, resolve)    QIO  finishing a read, fetching the next one          */
__device__ __forceinline__ uint32_t bt_class_of(uint32_t pc) {
	switch (pc) {
	case PC_LF: return QF;
	case PC_CHASE: return QC;
	case PC_PHASE: case PC_BT_BEGIN: case PC_BT_END: return QPH;
	case PC_FRAME_ENTER: case PC_POS: case PC_POS_END: return QPOS;
	case PC_BTLOOP: return QBT;
	case PC_FRAME_RET: case PC_CHILD_RET: return QRET;
	case PC_REPORT: case PC_REPORT_ROW: case PC_RESOLVE: case PC_REPORT_RET: return QREP;
	default: return QIO;
	}
}

/* One block = `nctx` read contexts (packed in shared memory) served by blockDim.x/32 worker warps.  A warp
 * repeatedly takes up to 32 contexts of ONE class from that class's queue, loads them into registers,
 * advances each by one fast transition (or one chain of rare transitions), stores them back and routes
 * them to the queue of their new class.  Lanes of a warp therefore execute the same code, and no context
 * ever waits for another one.  Reads come from the global cursor as before (ctl->next / ctl->nwork). */
__global__ void __launch_bounds__(BT_Q_THREADS, 1)
bt_search_kernel_q(BtKParams P, BtWorkCtl *ctl, uint32_t nctx) {
	extern __shared__ __align__(16) uint8_t bt_smem[];
	BtQueues *Q = reinterpret_cast<BtQueues *>(bt_smem);
	uint32_t *ctx = reinterpret_cast<uint32_t *>(bt_smem + sizeof(BtQueues));
	const uint32_t lane = threadIdx.x & 31;
	const unsigned long long nwork = ctl->nwork;
	if (threadIdx.x == 0) {
		for (int k = 0; k < BT_NQ; k++) { Q->head[k] = 0; Q->tail[k] = 0; }
		Q->tail[QIO] = nctx; Q->live = nctx;
	}
	for (uint32_t i = threadIdx.x; i < BT_NQ * BT_QCAP; i += blockDim.x) (&Q->item[0][0])[i] = 0;
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < nctx; i += blockDim.x) {
		Q->item[QIO][i] = i + 1;
		ctx[(size_t)26 * nctx + i] = PC_NEXT_READ;             /* word 26 holds pc in its low bits */
		ctx[(size_t)25 * nctx + i] = 0;
	}
	__syncthreads();
	BtLane L; BtLaneCold cold;
	memset(&L, 0, sizeof L); memset(&cold, 0, sizeof cold);
	L.K = &cold;
	L.qualThresh = P.pol.mode == 0 ? 0xffffffffu : P.pol.qualThresh;
	L.K->maxBts = P.pol.mode == 0 ? 0xffffffffu : P.pol.maxBts;
	L.maqPenalty = P.pol.mode == 0 ? 1u : (uint32_t)P.pol.maqRound;
	volatile uint32_t *vhead = Q->head, *vtail = Q->tail;
	volatile uint32_t *vlive = &Q->live;
	for (;;) {
		/* pick the fullest queue (lane 0), reserve up to 32 entries */
		int k = -1; uint32_t h = 0, n = 0;
		if (lane == 0) {
			for (;;) {
				uint32_t best = 0; k = -1;
				for (int q = 0; q < BT_NQ; q++) { uint32_t sz = vtail[q] - vhead[q]; if ((int32_t)sz > (int32_t)best) { best = sz; k = q; } }
				if (k < 0) { if (*vlive == 0) { k = -2; } break; }
				h = vhead[k]; uint32_t t = vtail[k];
				n = t - h; if ((int32_t)n <= 0) continue;
				if (n > 32) n = 32;
				if (atomicCAS(&Q->head[k], h, h + n) == h) break;
			}
		}
		k = __shfl_sync(0xffffffffu, k, 0); h = __shfl_sync(0xffffffffu, h, 0); n = __shfl_sync(0xffffffffu, n, 0);
		if (k == -2) break;
		if (k < 0) { __nanosleep(100); continue; }
		const bool active = lane < n;
		uint32_t id = 0;
		if (active) {
			volatile uint32_t *slot = &Q->item[k][(h + lane) & (BT_QCAP - 1)];
			uint32_t v;
			while ((v = *slot) == 0) { }                         /* the producer bumps tail before it writes the slot */
			*slot = 0;
			id = v - 1;
		}
		__threadfence_block();
		uint32_t ncls = 0xffffffffu;                             /* class after this visit; none = context retired */
		if (active) {
			const uint32_t gid = blockIdx.x * nctx + id;
			BtScratch S;
			S.rows = P.rows + (size_t)gid * P.R * 2; S.elims = P.elims + (size_t)gid * P.R;
			S.frames = P.frames + (size_t)gid * P.FCAP; S.partials = P.partials + (size_t)gid * P.PCAP;
			bt_ctx_load(L, ctx, nctx, id);
			L.rseq = P.stage + (size_t)gid * 2 * P.stage_len; L.rqual = L.rseq + P.stage_len;   /* the context's writable copy of its read */
			if (k == QF || k == QC) bt_fast_iter(L, P, S);
			else if (k == QIO) {
				if (L.pc == PC_FINISH_READ) { bt_finish_read(L, P); L.pc = PC_NEXT_READ; }
				if (L.pc == PC_NEXT_READ) {
					const unsigned long long w = atomicAdd(&ctl->next, 1ull);
					if (w < nwork) {
						const uint32_t rid = P.sel ? P.sel[w] : (uint32_t)w;
						bt_begin_read(L, P, rid);
						const unsigned long long ro = P.roff[rid];
						uint8_t *dst = P.stage + (size_t)gid * 2 * P.stage_len; const uint32_t qoff = P.stage_len;
						bool sawN = false;
						for (uint32_t i = 0; i < L.rlen; i++) {
							const uint8_t b = __ldg(P.seq + ro + i);
							dst[i] = b; dst[qoff + i] = __ldg(P.qual + ro + i);
							sawN |= (b == 4);
						}
						L.rseq = dst; L.rqual = dst + qoff; L.K->hasN = sawN;
					} else L.pc = PC_EXIT;
				}
			} else {
				/* rare transitions, chained while the context stays in this class */
#pragma unroll 1
				for (int c = 0; c < BT_RARE_CHAIN && bt_class_of(L.pc) == (uint32_t)k; c++) {
					if (L.flags & BT_FLAG_SCRATCH_OVF) { L.pc = PC_FINISH_READ; break; }
					if (P.budget && L.nit > P.budget) { L.flags |= BT_FLAG_BUDGET; L.pc = PC_FINISH_READ; break; }
					L.s_iter++; L.nit++;
					bt_rare_step(L, P, S);
				}
			}
			if (L.pc != PC_EXIT) { ncls = bt_class_of(L.pc); bt_ctx_store(L, ctx, nctx, id); }
		}
		__threadfence_block();                                   /* context words visible before the id is published */
		unsigned retired = __ballot_sync(0xffffffffu, active && ncls == 0xffffffffu);
		if (retired && lane == 0) atomicSub(&Q->live, (uint32_t)__popc(retired));
#pragma unroll
		for (int q = 0; q < BT_NQ; q++) {
			const unsigned m = __ballot_sync(0xffffffffu, ncls == (uint32_t)q);
			if (m) {
				uint32_t base = 0;
				if (lane == 0) base = atomicAdd(&Q->tail[q], (uint32_t)__popc(m));
				base = __shfl_sync(0xffffffffu, base, 0);
				if (ncls == (uint32_t)q) Q->item[q][(base + (uint32_t)__popc(m & ((1u << lane) - 1u))) & (BT_QCAP - 1)] = id + 1;
			}
		}
	}
	/* statistics: warp-reduce, one atomic per warp and counter */
	unsigned long long v[8] = { L.s_lfex, L.s_lf, L.s_chase, L.K->s_ftab, L.K->s_offs, L.K->s_bt, L.s_iter, L.s_blk };
#pragma unroll
	for (int kk = 0; kk < 8; kk++) {
		unsigned long long x = v[kk];
		for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
		if (lane == 0 && x) atomicAdd(&P.stats[kk], x);
	}
}

static size_t bt_q_smem(uint32_t nctx) { return sizeof(BtQueues) + (size_t)nctx * BT_CTX_WORDS * 4; }

/* Best-first path (--best / --strata / -M / -v 3): every thread takes reads from the global cursor and runs the whole
 * aligner loop (UnpairedAlignerV2 / PairedBWAlignerV1 / V2) for each on its own arena of P.arenaWords words.  `lanes` threads of each block are active. */
#define BF_THREADS 64
template <bool PAIRED>
__global__ void __launch_bounds__(BF_THREADS)
bt_best_kernel(const __grid_constant__ BfKParams P, BtWorkCtl *ctl, uint32_t lanes) {
	const unsigned long long nwork = ctl->nwork;
	/* claim one entry of the tier's arena pool for this block: first clear bit of the mask, starting at a block-dependent word */
	__shared__ uint32_t s_blk;
	if (threadIdx.x == 0) {
		const uint32_t nw = (P.poolBlocks + 31) / 32;
		uint32_t w = blockIdx.x % nw, got = 0xffffffffu, sweeps = 0, seen = 0;
		if (*(volatile unsigned long long *)&ctl->next >= nwork) got = 0xfffffffeu;    /* nothing left (the later tiers are usually empty): no arena needed */
		while (got == 0xffffffffu) {
			const unsigned valid = (w == nw - 1 && (P.poolBlocks & 31)) ? ((1u << (P.poolBlocks & 31)) - 1u) : 0xffffffffu;
			const unsigned fr = ~atomicOr(P.poolMask + w, 0u) & valid;
			if (fr) {
				const unsigned b = (unsigned)__ffs(fr) - 1u;
				if (!(atomicOr(P.poolMask + w, 1u << b) & (1u << b))) got = w * 32 + b;
			} else {
				w = (w + 1) % nw;
				if (++seen == nw) {                                                 /* every entry is held by a resident block: wait for one to finish */
					seen = 0; __nanosleep(20000);
					if (++sweeps > 3000000u) __trap();                                 /* a minute: something is wrong — fail loudly rather than hang */
				}
			}
		}
		s_blk = got;
	}
	__syncthreads();
	const uint32_t blk = s_blk;
	if (blk == 0xfffffffeu) return;
	if (threadIdx.x < lanes) {
	BfCtx X;
	X.P = &P;
	X.A = P.arena + ((size_t)blk * lanes + threadIdx.x) * P.arenaWords; X.acap = P.arenaWords;
	X.s_lfex = X.s_lf = X.s_chase = X.s_ftab = X.s_offs = X.s_bt = 0;
	for (;;) {
		const unsigned long long w = atomicAdd(&ctl->next, 1ull);
		if (w >= nwork) break;
		const uint32_t rid = P.sel ? P.sel[w] : (uint32_t)w;                 /* read id, or pair id (mates are reads 2p, 2p+1) */
		const uint32_t r0 = PAIRED ? 2 * rid : rid;
		const unsigned long long ro = P.roff[r0];
		X.rid = rid; X.rlenM[0] = (uint32_t)(P.roff[r0 + 1] - ro); X.seedM[0] = P.seeds[r0];
		X.seqM[0] = P.seq + ro; X.qualM[0] = P.qual + ro;
		X.rlenM[1] = 0; X.seedM[1] = 0; X.seqM[1] = X.seqM[0]; X.qualM[1] = X.qualM[0];
		if (PAIRED) {
			const unsigned long long ro1 = P.roff[r0 + 1];
			X.rlenM[1] = (uint32_t)(P.roff[r0 + 2] - ro1); X.seedM[1] = P.seeds[r0 + 1];
			X.seqM[1] = P.seq + ro1; X.qualM[1] = P.qual + ro1;
		}
		X.atop = 1; X.amax = 1; X.steps = 0; X.flags = 0; X.found = 0;
		X.top.rssOff = X.top.rssCap = X.top.nRss = X.top.actOff = X.top.actCap = X.top.nAct = 0;
		X.top.lastRange = X.top.delayedRange = 0; X.top.minCost = 0; X.top.done = 0; X.top.foundRange = 0; X.top.rnd = 0; X.top.paired = 0;
		if (PAIRED) { if (P.prog.pairedV2) bf_align_pair_v2(X); else bf_align_pair(X); } else bf_align_read(X);
		if (X.flags & (BT_FLAG_STACK_OVF | BT_FLAG_FRAME_OVF)) X.found = 0;     /* STACK_OVF: re-run by a pass with a larger arena */
		P.found[rid] = X.found; P.flags[rid] = X.flags;
	}
	if (X.s_lfex) atomicAdd(&P.stats[0], (unsigned long long)X.s_lfex);
	if (X.s_lf) atomicAdd(&P.stats[1], (unsigned long long)X.s_lf);
	if (X.s_chase) atomicAdd(&P.stats[2], (unsigned long long)X.s_chase);
	if (X.s_ftab) atomicAdd(&P.stats[3], (unsigned long long)X.s_ftab);
	if (X.s_offs) atomicAdd(&P.stats[4], (unsigned long long)X.s_offs);
	if (X.s_bt) atomicAdd(&P.stats[5], (unsigned long long)X.s_bt);
	}
	__syncthreads();                                                         /* every lane is done with its arena */
	if (threadIdx.x == 0) atomicAnd(P.poolMask + (blk >> 5), ~(1u << (blk & 31)));
}

/* Appends to sel_out the reads (of the first n work items of sel_in / the identity) whose flags intersect `mask`;
 * ctl->nwork is the list length.  If `count_ctl` is set, the number of work items is read from it (device-sized lists). */
__global__ void bt_collect_kernel(const uint32_t *flags, const uint32_t *sel_in, uint32_t n, const BtWorkCtl *count_ctl, uint32_t mask, uint32_t *sel_out, BtWorkCtl *ctl) {
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (count_ctl) n = (uint32_t)count_ctl->nwork;
	if (i >= n) return;
	uint32_t rid = sel_in ? sel_in[i] : i;
	if (flags[rid] & mask) {
		unsigned long long p = atomicAdd(&ctl->nwork, 1ull);
		sel_out[p] = rid;
	}
}
__global__ void bt_ctl_set_kernel(BtWorkCtl *ctl, unsigned long long nwork) { ctl->next = 0; ctl->nwork = nwork; }
/* resets a context's whole array of work lists: list 0 holds `nwork` items, the others are empty */
#define BT_CTL_WORDS 12
__global__ void bt_ctl_set_all_kernel(BtWorkCtl *ctl, uint32_t n, unsigned long long nwork) {
	if (threadIdx.x < n) { ctl[threadIdx.x].next = 0; ctl[threadIdx.x].nwork = threadIdx.x == 0 ? nwork : 0; }
}

/* ------------------------------------------------------------------------------------------- */
/* host side                                                                                    */
/* ------------------------------------------------------------------------------------------- */

static thread_local std::string g_err;
static int fail(const std::string &m) { g_err = m; return 1; }
int bt_internal_fail(const std::string &m) { return fail(m); }       /* for bt_build.cu, bt_io.cu */
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

struct HostEbwt {     /* the parsed contents of X.1.ebwt / X.2.ebwt */
	uint32_t len = 0; int32_t lineRate = 0, linesPerSide = 0, offRate = 0, ftabChars = 0, flags = 0;
	uint32_t nPat = 0, nFrag = 0, zOff = 0; uint32_t fchr[5] = {0,0,0,0,0};
	std::vector<uint32_t> plen, rstarts, ftab, eftab, offs;
	std::vector<uint8_t> ebwt;
	std::vector<std::string> refnames;
};

struct DevEbwt {
	uint4 *blocks = nullptr; uint32_t *offs = nullptr, *ftab = nullptr, *eftab = nullptr, *rstarts = nullptr, *plen = nullptr;
	uint32_t nblocks = 0;
	BtDevIndex dev;
	uint64_t bytes = 0;
};

struct Workspace {
	uint32_t nthreads = 0, R = 0, FCAP = 0, PCAP = 0;
	uint32_t stage_len = 0;
	uint4 *rows = nullptr; uint8_t *elims = nullptr; BtFrame *frames = nullptr; uint64_t *partials = nullptr; uint8_t *stage = nullptr;
	void release() { cudaFree(rows); cudaFree(elims); cudaFree(frames); cudaFree(partials); cudaFree(stage); rows = nullptr; elims = nullptr; frames = nullptr; partials = nullptr; stage = nullptr; nthreads = 0; }
};

struct ArenaPool { uint32_t *mem = nullptr; unsigned *mask = nullptr; uint32_t blocks = 0, lanes = 0, words = 0; };
struct bt_context;
struct bt_index {
	int device = 0;
	bool has_mirror = false;
	HostEbwt host[2];            /* small arrays + names kept; ebwt bytes dropped after upload */
	DevEbwt dev[2];
	unsigned long long *stats = nullptr;
	int sms = 0, blocks_per_sm = 0;
	bt_context *def = nullptr;   /* context behind bt_align_batch / bt_align_batch_device */
	std::string base;            /* index basename: the bit-pair reference (X.3.ebwt / X.4.ebwt) is loaded on the first paired-end call */
	ArenaPool bf_pool[4];        /* best-first path: the arena tiers, shared by all contexts (entries are claimed per resident block) */
	bool ref_loaded = false; BtDevRef dref; uint32_t *d_refwords[4] = { nullptr, nullptr, nullptr, nullptr }; uint8_t *d_refbuf = nullptr;
	std::mutex mu;
};

/* Everything one in-flight batch needs besides the (shared, immutable) index: scratch for both passes,
 * the work-queue words and the device staging of the host-buffer entry points.  One call at a time per
 * context; different contexts may be in flight concurrently on different streams. */
struct bt_context {
	bt_index *ix = nullptr;
	Workspace ws1, wsh, ws2;     /* main pass / heavy-read pass / scratch-overflow pass */
	BtWorkCtl *ctl = nullptr;    /* [BT_CTL_WORDS]: main pass, tail, overflow pass (best-first path: its four tiers) */
	Workspace wsl; uint32_t *slot_ctx = nullptr, *tailq_items = nullptr; BtTailQ *tailq = nullptr; uint32_t slot_cap = 0, tailq_cap = 0;   /* checkpoint slots (bt_ctxq.cuh) and the tail's ring (bt_tail.cu) */
	uint32_t *heavy_sel = nullptr, *ultra_sel = nullptr, *retry_sel = nullptr; uint32_t retry_cap = 0;
	cudaStream_t side = nullptr; /* the heavy and overflow passes run here, overlapping the next batch's main pass */
	cudaEvent_t ev_main = nullptr, ev_tail = nullptr;
	uint8_t *d_seq = nullptr, *d_qual = nullptr; uint64_t *d_offs = nullptr; uint32_t *d_seeds = nullptr, *d_sel = nullptr;
	uint32_t *d_found = nullptr, *d_flags = nullptr, *d_hits = nullptr;
	size_t cap_seq = 0, cap_qual = 0, cap_offs = 0, cap_seeds = 0, cap_found = 0, cap_flags = 0, cap_hitwords = 0, cap_sel = 0;
	std::mutex mu;
};

int bt_internal_context_device(bt_context_t *cx) { return cx->ix->device; }
bt_index_t *bt_internal_context_index(bt_context_t *cx) { return cx->ix; }

static bool read_exact(FILE *f, void *p, size_t n) { return fread(p, 1, n, f) == n; }

/* Ebwt::readIntoMemory (ebwt.h:2835-3445), small-index format */
static int parse_ebwt(const std::string &base, bool mirror, HostEbwt &h) {
	std::string p1 = base + (mirror ? ".rev" : "") + ".1.ebwt", p2 = base + (mirror ? ".rev" : "") + ".2.ebwt";
	FILE *f1 = fopen(p1.c_str(), "rb");
	if (!f1) return fail("cannot open " + p1);
	FILE *f2 = fopen(p2.c_str(), "rb");
	if (!f2) { fclose(f1); return fail("cannot open " + p2); }
	int rc = 1;
	do {
		uint32_t one = 0, hdr[6];
		if (!read_exact(f1, &one, 4) || one != 1) { fail(p1 + ": bad endianness sentinel (big-endian indexes are not supported)"); break; }
		if (!read_exact(f2, &one, 4) || one != 1) { fail(p2 + ": bad endianness sentinel"); break; }
		if (!read_exact(f1, hdr, sizeof hdr)) { fail(p1 + ": truncated header"); break; }
		h.len = hdr[0]; h.lineRate = (int32_t)hdr[1]; h.linesPerSide = (int32_t)hdr[2]; h.offRate = (int32_t)hdr[3];
		h.ftabChars = (int32_t)hdr[4]; h.flags = (int32_t)hdr[5];
		if (h.lineRate != 6 || h.linesPerSide != 1) { fail(p1 + ": unsupported side geometry (need lineRate 6, linesPerSide 1)"); break; }
		if (h.ftabChars < 1 || h.ftabChars > 16 || h.offRate < 0 || h.offRate > 31) { fail(p1 + ": implausible header"); break; }
		if (!read_exact(f1, &h.nPat, 4)) { fail(p1 + ": truncated"); break; }
		h.plen.resize(h.nPat);
		if (!read_exact(f1, h.plen.data(), 4 * (size_t)h.nPat)) { fail(p1 + ": truncated plen"); break; }
		if (!read_exact(f1, &h.nFrag, 4)) { fail(p1 + ": truncated"); break; }
		h.rstarts.resize(3 * (size_t)h.nFrag);
		if (!read_exact(f1, h.rstarts.data(), 12 * (size_t)h.nFrag)) { fail(p1 + ": truncated rstarts"); break; }
		uint32_t bwtSz = h.len / 4 + 1;
		size_t ebwtTotLen = (size_t)((bwtSz + 111) / 112) * 128;          /* EbwtParams::init ebwt.h:168-171 */
		h.ebwt.resize(ebwtTotLen + 128);                                  /* + slack: bw side reads side+120 */
		if (!read_exact(f1, h.ebwt.data(), ebwtTotLen)) { fail(p1 + ": truncated ebwt[]"); break; }
		if (!read_exact(f1, &h.zOff, 4) || !read_exact(f1, h.fchr, 20)) { fail(p1 + ": truncated"); break; }
		size_t ftabLen = ((size_t)1 << (2 * h.ftabChars)) + 1, eftabLen = 2 * (size_t)h.ftabChars;
		h.ftab.resize(ftabLen); h.eftab.resize(eftabLen);
		if (!read_exact(f1, h.ftab.data(), 4 * ftabLen) || !read_exact(f1, h.eftab.data(), 4 * eftabLen)) { fail(p1 + ": truncated ftab"); break; }
		int c;
		while ((c = fgetc(f1)) != EOF) {                                   /* ebwt.h:3258-3272 */
			if (c == '\0') break;
			if (c == '\n') h.refnames.push_back("");
			else { if (h.refnames.empty()) h.refnames.push_back(""); h.refnames.back().push_back((char)c); }
		}
		size_t offsLen = ((size_t)h.len + 1 + ((size_t)1 << h.offRate) - 1) >> h.offRate;
		h.offs.resize(offsLen);
		if (!read_exact(f2, h.offs.data(), 4 * offsLen)) { fail(p2 + ": truncated offs[]"); break; }
		rc = 0;
	} while (0);
	fclose(f1); fclose(f2);
	return rc;
}

template <typename T> static int upload(const std::vector<T> &v, T **d, uint64_t &bytes) {
	size_t n = v.size() ? v.size() : 1;
	CUDA_TRY(cudaMalloc((void **)d, n * sizeof(T)));
	if (v.size()) CUDA_TRY(cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
	bytes += n * sizeof(T);
	return 0;
}

static int upload_index(HostEbwt &h, bool fw, DevEbwt &d) {
	uint8_t *d_raw = nullptr;
	CUDA_TRY(cudaMalloc((void **)&d_raw, h.ebwt.size()));
	CUDA_TRY(cudaMemcpy(d_raw, h.ebwt.data(), h.ebwt.size(), cudaMemcpyHostToDevice));
	BtNativeIndex n;
	n.ebwt = d_raw; n.len = h.len; n.zOff = h.zOff;
	{   /* Ebwt::postReadInit (ebwt.h:1043-1059) */
		uint32_t sideNum = h.zOff / 224, sideCharOff = h.zOff % 224, by = sideCharOff >> 2, bp = sideCharOff & 3;
		if ((sideNum & 1) == 0) { by = 56 - by - 1; bp = 3 - bp; }
		n.zEbwtByteOff = by + sideNum * 64; n.zEbwtBpOff = bp;
	}
	memcpy(n.fchr, h.fchr, sizeof n.fchr);
	d.nblocks = (h.len >> 6) + 1;
	CUDA_TRY(cudaMalloc((void **)&d.blocks, (size_t)d.nblocks * 32));
	d.bytes += (uint64_t)d.nblocks * 32;
	bt_relayout_kernel<<<(d.nblocks + 127) / 128, 128>>>(n, d.nblocks, d.blocks);
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaDeviceSynchronize());
	CUDA_TRY(cudaFree(d_raw));
	if (upload(h.offs, &d.offs, d.bytes) || upload(h.ftab, &d.ftab, d.bytes) || upload(h.eftab, &d.eftab, d.bytes) ||
	    upload(h.rstarts, &d.rstarts, d.bytes) || upload(h.plen, &d.plen, d.bytes)) return 1;
	BtDevIndex &x = d.dev;
	x.blocks = d.blocks; x.offs = d.offs; x.ftab = d.ftab; x.eftab = d.eftab; x.rstarts = d.rstarts; x.plen = d.plen;
	x.len = h.len; x.zOff = h.zOff; x.nFrag = h.nFrag; x.nPat = h.nPat; x.offMask = 0xffffffffu << h.offRate;
	x.offRate = h.offRate; x.ftabChars = h.ftabChars; memcpy(x.fchr, h.fchr, sizeof x.fchr); x.fw = fw ? 1u : 0u;
	std::vector<uint8_t>().swap(h.ebwt);   /* the native image is not needed after the re-layout */
	std::vector<uint32_t>().swap(h.offs);
	std::vector<uint32_t>().swap(h.ftab);
	return 0;
}

extern "C" int bt_abi_version(void) { return BT_ABI_VERSION; }
extern "C" const char *bt_last_error(void) { return g_err.c_str(); }

extern "C" void bt_policy_init(bt_policy_t *p) {
	memset(p, 0, sizeof *p);
	p->mode = 1; p->mms = 2; p->seed_len = 28; p->qual_thresh = 70; p->max_bts = 125; p->khits = 1; p->mhits = 0xffffffffu; p->maq_round = 1; p->max_bts_best = 800; p->max_ins = 250; p->mate1fw = 1; p->pair_tries = 100;
}

extern "C" void bt_context_free(bt_context_t *cx) {
	if (!cx) return;
	cudaSetDevice(cx->ix->device);
	if (cx->side) { cudaStreamSynchronize(cx->side); cudaStreamDestroy(cx->side); }
	if (cx->ev_main) cudaEventDestroy(cx->ev_main);
	if (cx->ev_tail) cudaEventDestroy(cx->ev_tail);
	cx->ws1.release(); cx->wsh.release(); cx->ws2.release(); cx->wsl.release();
	cudaFree(cx->slot_ctx); cudaFree(cx->tailq_items); cudaFree(cx->tailq);
	cudaFree(cx->ctl); cudaFree(cx->retry_sel); cudaFree(cx->heavy_sel); cudaFree(cx->ultra_sel);
	cudaFree(cx->d_seq); cudaFree(cx->d_qual); cudaFree(cx->d_offs); cudaFree(cx->d_seeds); cudaFree(cx->d_sel);
	cudaFree(cx->d_found); cudaFree(cx->d_flags); cudaFree(cx->d_hits);
	delete cx;
}

extern "C" int bt_context_create(bt_index_t *ix, bt_context_t **out) {
	if (!ix || !out) return fail("bt_context_create: null argument");
	*out = nullptr;
	CUDA_TRY(cudaSetDevice(ix->device));
	bt_context *cx = new bt_context();
	cx->ix = ix;
	if (cudaMalloc((void **)&cx->ctl, BT_CTL_WORDS * sizeof(BtWorkCtl)) != cudaSuccess || cudaStreamCreateWithFlags(&cx->side, cudaStreamNonBlocking) != cudaSuccess ||
	    cudaEventCreateWithFlags(&cx->ev_main, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&cx->ev_tail, cudaEventDisableTiming) != cudaSuccess) {
		bt_context_free(cx); return fail("bt_context_create: CUDA resource allocation failed");
	}
	*out = cx;
	return 0;
}

static bt_context *default_ctx(bt_index_t *ix) {
	std::lock_guard<std::mutex> g(ix->mu);
	if (!ix->def) { bt_context *cx = nullptr; if (bt_context_create(ix, &cx)) return nullptr; ix->def = cx; }
	return ix->def;
}

extern "C" void bt_index_free(bt_index_t *ix) {
	if (!ix) return;
	cudaSetDevice(ix->device);
	for (int k = 0; k < 2; k++) {
		DevEbwt &d = ix->dev[k];
		cudaFree(d.blocks); cudaFree(d.offs); cudaFree(d.ftab); cudaFree(d.eftab); cudaFree(d.rstarts); cudaFree(d.plen);
	}
	if (ix->def) bt_context_free(ix->def);
	for (int k = 0; k < 4; k++) cudaFree(ix->d_refwords[k]);
	for (int k = 0; k < 4; k++) { cudaFree(ix->bf_pool[k].mem); cudaFree(ix->bf_pool[k].mask); }
	cudaFree(ix->d_refbuf);
	cudaFree(ix->stats);
	delete ix;
}

extern "C" int bt_index_load(const char *basename, int need_mirror, int device, bt_index_t **out) {
	if (!basename || !out) return fail("bt_index_load: null argument");
	*out = nullptr;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
		return fail("bt_index_load: no CUDA device available (this library has no CPU search path)");
	if (device < 0 || device >= ndev) return fail("bt_index_load: bad device ordinal");
	CUDA_TRY(cudaSetDevice(device));
	bt_index *ix = new bt_index();
	ix->device = device;
	ix->base = basename;
	ix->has_mirror = need_mirror != 0;
	int rc = parse_ebwt(basename, false, ix->host[0]);
	if (!rc) rc = upload_index(ix->host[0], true, ix->dev[0]);
	if (!rc && need_mirror) {
		rc = parse_ebwt(basename, true, ix->host[1]);
		if (!rc) rc = upload_index(ix->host[1], false, ix->dev[1]);
	}
	if (!rc) {
		cudaDeviceProp prop;
		if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) rc = fail("cudaGetDeviceProperties failed");
		else {
			ix->sms = prop.multiProcessorCount;
			int bps = 0;
			if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, bt_search_kernel, BT_THREADS, BT_THREADS * BT_SMEM_STRIDE) != cudaSuccess || bps < 1) bps = 1;
			{ const char *e = getenv("BT_MAIN_BLOCKS"); if (e && atoi(e) >= 1 && atoi(e) < bps) bps = atoi(e); }   /* occupancy experiments (profiles/) */
			ix->blocks_per_sm = bps;
			if (cudaFuncSetAttribute(bt_search_kernel_q, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bt_q_smem(BT_Q_NCTX)) != cudaSuccess) rc = fail("cudaFuncSetAttribute(shared memory) failed");
		}
	}
	if (!rc && (cudaMalloc((void **)&ix->stats, 8 * sizeof(unsigned long long)) != cudaSuccess ||
	            cudaMemset(ix->stats, 0, 8 * sizeof(unsigned long long)) != cudaSuccess)) rc = fail("cudaMalloc failed");
	if (rc) { std::string keep = g_err; bt_index_free(ix); g_err = keep; return rc; }
	*out = ix;
	return 0;
}

extern "C" int bt_index_info(const bt_index_t *ix, bt_index_info_t *info) {
	if (!ix || !info) return fail("bt_index_info: null argument");
	info->len = ix->host[0].len; info->n_refs = ix->host[0].nPat; info->off_rate = ix->host[0].offRate; info->ftab_chars = ix->host[0].ftabChars;
	info->has_mirror = ix->has_mirror; info->device_bytes = ix->dev[0].bytes + ix->dev[1].bytes;
	return 0;
}
extern "C" const char *bt_index_refname(const bt_index_t *ix, uint32_t i) {
	if (!ix || i >= ix->host[0].refnames.size()) return nullptr;
	return ix->host[0].refnames[i].c_str();
}
extern "C" uint32_t bt_index_reflen(const bt_index_t *ix, uint32_t i) {
	if (!ix || i >= ix->host[0].plen.size()) return 0;
	return ix->host[0].plen[i];
}

static int ensure_ws(Workspace &w, uint32_t nthreads, uint32_t R, uint32_t FCAP, uint32_t PCAP, uint32_t stage_len) {
	if (w.nthreads >= nthreads && w.R >= R && w.FCAP >= FCAP && w.PCAP >= PCAP && w.stage_len >= stage_len) return 0;
	w.release();
	CUDA_TRY(cudaMalloc((void **)&w.rows, (size_t)nthreads * R * 32));
	CUDA_TRY(cudaMalloc((void **)&w.elims, (size_t)nthreads * R));
	CUDA_TRY(cudaMalloc((void **)&w.frames, (size_t)nthreads * FCAP * sizeof(BtFrame)));
	CUDA_TRY(cudaMalloc((void **)&w.partials, (size_t)nthreads * PCAP * 8));
	CUDA_TRY(cudaMalloc((void **)&w.stage, (size_t)nthreads * 2 * (stage_len ? stage_len : 1)));
	w.nthreads = nthreads; w.R = R; w.FCAP = FCAP; w.PCAP = PCAP; w.stage_len = stage_len;
	return 0;
}

static bool policy_is_best(const bt_policy_t *pol) { return pol->best || pol->strata || pol->sample_max || pol->paired || (pol->mode == 0 && pol->mms == 3); }
/* BitPairReference (reference.h) onto the device, once per index */
static int ensure_ref(bt_index_t *ix) {
	std::lock_guard<std::mutex> g(ix->mu);
	if (ix->ref_loaded) return 0;
	BtHostRef h; std::string err;
	if (!bt_load_ref(ix->base, h, err)) return fail("bt_align (paired-end): " + err);
	uint64_t bytes = 0;
	const std::vector<uint32_t> *v[4] = { &h.recs, &h.refRecOffs, &h.refOffs, &h.approxLen };
	for (int k = 0; k < 4; k++) if (upload(*v[k], &ix->d_refwords[k], bytes)) return 1;
	if (upload(h.buf, &ix->d_refbuf, bytes)) return 1;
	ix->dref.recs = ix->d_refwords[0]; ix->dref.refRecOffs = ix->d_refwords[1]; ix->dref.refOffs = ix->d_refwords[2]; ix->dref.approxLen = ix->d_refwords[3];
	ix->dref.buf = ix->d_refbuf; ix->dref.nRefs = h.nRefs;
	ix->ref_loaded = true;
	return 0;
}
static int check_policy(const bt_index_t *ix, const bt_policy_t *pol) {
	if (pol->mode == 0) { if (pol->mms < 0 || pol->mms > 3) return fail("bt_align: -v must be 0..3"); }
	else if (pol->mode == 1) { if (pol->mms < 0 || pol->mms > 3) return fail("bt_align: -n must be 0..3"); if (pol->seed_len < 5) return fail("bt_align: -l must be >= 5"); }
	else return fail("bt_align: bad mode");
	if ((pol->mode == 1 || pol->mms > 0) && !ix->has_mirror) return fail("bt_align: this policy needs the mirror index (load with need_mirror=1)");
	if (!pol->all_hits && pol->khits == 0) return fail("bt_align: -k must be >= 1");
	return 0;
}

/* Enqueue first pass + collect + retry pass.  All pointers are device pointers. `maxlen` bounds the read length. */
#ifndef BT_MAIN_BUDGET
#define BT_MAIN_BUDGET 8000u       /* transitions a read may take in the main pass before it is moved to the heavy pass */
#endif
/* Main / heavy pass kernels: thread-per-lane (default) or the experimental queue-driven kernel
 * (BT_MAIN_KERNEL=q; see DESIGN.md §4.2 for the measurements that decided the default). */
static uint32_t env_u32(const char *name, uint32_t dflt) { const char *e = getenv(name); return e ? (uint32_t)atol(e) : dflt; }
static uint32_t main_budget() {
	static long v = -1;
	if (v < 0) { const char *e = getenv("BT_MAIN_BUDGET"); v = e ? atol(e) : (long)BT_MAIN_BUDGET; }
	return (uint32_t)v;
}
static bool main_kernel_is_queue() {
	static int v = -1;
	if (v < 0) { const char *e = getenv("BT_MAIN_KERNEL"); v = (e && e[0] == 'q') ? 1 : 0; }
	return v == 1;
}


/* Grows tier k's arena pool of the index to `blocks` entries of `lanes` x `words` (caller holds ix->mu).  Entries are claimed by resident
 * blocks (bt_best_kernel), so a pool may be smaller than what is in flight — blocks then wait for an entry — but it must not be
 * freed while any kernel may hold one: growing synchronises the device first. */
static int ensure_pool(bt_index *ix, int k, uint32_t blocks, uint32_t lanes, uint32_t words) {
	ArenaPool &p = ix->bf_pool[k];
	if (p.mem && p.blocks >= blocks && p.lanes == lanes && p.words == words) return 0;
	if (p.mem) { CUDA_TRY(cudaDeviceSynchronize()); if (blocks < p.blocks) blocks = p.blocks; }
	cudaFree(p.mem); cudaFree(p.mask); p.mem = nullptr; p.mask = nullptr; p.blocks = 0;
	CUDA_TRY(cudaMalloc((void **)&p.mem, (size_t)blocks * lanes * words * 4));
	const size_t mw = (blocks + 31) / 32;
	CUDA_TRY(cudaMalloc((void **)&p.mask, mw * 4));
	CUDA_TRY(cudaMemset(p.mask, 0, mw * 4));
	p.blocks = blocks; p.lanes = lanes; p.words = words;
	return 0;
}

/* The best-first path (bt_best.cuh).  Four passes with growing per-read arenas: every read with 128 KB (pairs: 192 KB) on the caller's
 * stream; the reads that exhausted it with 1 MB (32 lanes per block), then 16 MB (one lane per block), then 256 MB (8 blocks; the
 * reference's own ceiling is 64 MB of chunked pools per thread) on the side stream.  The arenas of a tier are ONE pool per index, sized
 * by the number of blocks that can be resident (occupancy x SMs for the first tier), whatever the number of batches in flight:
 * about 25 + 19 + 2.5 + 2 GB for full-size batches, a few MB for small ones. */
static int enqueue_best(bt_context *cx, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out, uint32_t maxlen, cudaStream_t st) {
	bt_index_t *ix = cx->ix;
	if (pol->paired && (in->nreads & 1)) return fail("bt_align (paired-end): nreads must be even (mates are adjacent reads)");
	const uint32_t nwork = in->sel ? in->nsel : (pol->paired ? in->nreads / 2 : in->nreads);
	if (nwork == 0) return 0;
	if (pol->paired && ensure_ref(ix)) return 1;
	static const uint32_t kw0 = env_u32("BT_BEST_ARENA_KW", 32);
	enum { NT = 4 };
	/* 192 KB (sized for pairs; unpaired reads use the same pool), 1 MB, 16 MB, 256 MB per read.  Measured high-water marks on the bench
	 * workloads (host emulation, profiles/README.md): -n 2 --best p99 35 KB, p99.9 282 KB; paired -n 3 p95 60 KB, p99 113 KB, p99.9 326 KB */
	const uint32_t tierWords[NT] = { (kw0 + kw0 / 2) << 10, 256u << 10, 4096u << 10, 65536u << 10 };
	const uint32_t tierLanes[NT] = { BF_THREADS, 32, 1, 1 };                                /* active threads per block */
	/* first tier: 12 blocks of 64 lanes per SM = 768 resident threads (72 / 80 registers per thread: the register file allows 910 / 819) */
	static const uint32_t bps0 = env_u32("BT_BEST_BLOCKS", 12);
	static const uint32_t t1b = env_u32("BT_BEST_T1_BLOCKS", 4);
	static int occ = 0;
	if (!occ) {
		int a = 0, b = 0;
		CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, bt_best_kernel<false>, BF_THREADS, 0));
		CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, bt_best_kernel<true>, BF_THREADS, 0));
		occ = a > b ? a : b; if (occ < 1) occ = 1;
	}
	uint32_t tierBlocks[NT] = { (uint32_t)ix->sms * bps0, (uint32_t)ix->sms * t1b, (uint32_t)ix->sms, 8 };
	uint32_t poolBlocks[NT] = { (uint32_t)ix->sms * (uint32_t)occ, (uint32_t)ix->sms * t1b, (uint32_t)ix->sms, 8 };
	for (int k = 0; k < NT; k++) {
		const uint32_t need_blocks = (nwork + tierLanes[k] - 1) / tierLanes[k];   /* small batches do not need a full machine of arenas */
		if (tierBlocks[k] > need_blocks) tierBlocks[k] = need_blocks;
		if (poolBlocks[k] > need_blocks) poolBlocks[k] = need_blocks;
	}
	if (cx->retry_cap < nwork) {
		cudaFree(cx->retry_sel); cudaFree(cx->heavy_sel); cudaFree(cx->ultra_sel); cx->retry_sel = cx->heavy_sel = cx->ultra_sel = nullptr; cx->retry_cap = 0;
		CUDA_TRY(cudaMalloc((void **)&cx->retry_sel, (size_t)nwork * 4));
		CUDA_TRY(cudaMalloc((void **)&cx->heavy_sel, (size_t)nwork * 4));
		CUDA_TRY(cudaMalloc((void **)&cx->ultra_sel, (size_t)nwork * 4));
		cx->retry_cap = nwork;
	}
	BfKParams P; memset(&P, 0, sizeof P);
	P.ix[0] = ix->dev[0].dev; P.ix[1] = ix->dev[1].dev;
	memcpy(&P.pol, pol, sizeof(BtPolicy));
	bf_build_prog(pol->mode, pol->mms, pol->seed_len, pol->qual_thresh, pol->nofw, pol->norc, &P.prog,
	              pol->paired, pol->mate1fw, pol->mate2fw, pol->min_ins, pol->max_ins, pol->pair_tries, pol->mhits, 0, pol->best);
	if (pol->paired) P.ref = ix->dref;
	P.seq = in->seq; P.qual = in->qual; P.roff = in->offs; P.seeds = in->seeds; P.sel = in->sel; P.nwork = nwork;
	P.found = out->found; P.flags = out->flags; P.hits = out->hits; P.slots = out->slots; P.mm_cap = out->mm_cap; P.rec_words = BT_HIT_HDR + out->mm_cap;
	P.stats = ix->stats;
	(void)maxlen;
	const uint32_t cblocks = (nwork + 255) / 256;
	/* pools and launches under the index lock: a pool that another thread grows must not be handed to a kernel here meanwhile */
	std::lock_guard<std::mutex> g(ix->mu);
	for (int k = 0; k < NT; k++) if (ensure_pool(ix, k, poolBlocks[k], tierLanes[k], tierWords[k])) return 1;
	CUDA_TRY(cudaStreamWaitEvent(st, cx->ev_tail, 0));
	bt_ctl_set_all_kernel<<<1, 32, 0, st>>>(cx->ctl, BT_CTL_WORDS, nwork);
	auto kernel = pol->paired ? bt_best_kernel<true> : bt_best_kernel<false>;
	uint32_t *sels[2] = { cx->heavy_sel, cx->retry_sel };                 /* tier k reads the list tier k-1 wrote; two buffers alternate */
	for (int k = 0; k < NT; k++) {
		cudaStream_t s = k == 0 ? st : cx->side;
		const uint32_t *sel_in = k == 0 ? in->sel : sels[(k - 1) & 1];
		P.sel = sel_in; P.arena = ix->bf_pool[k].mem; P.arenaWords = tierWords[k]; P.poolMask = ix->bf_pool[k].mask; P.poolBlocks = ix->bf_pool[k].blocks;
		kernel<<<tierBlocks[k], BF_THREADS, 0, s>>>(P, cx->ctl + k, tierLanes[k]);
		if (k + 1 < NT) bt_collect_kernel<<<cblocks, 256, 0, s>>>(out->flags, sel_in, nwork, k == 0 ? nullptr : cx->ctl + k, BT_FLAG_STACK_OVF, sels[k & 1], cx->ctl + k + 1);
		if (k == 0) { CUDA_TRY(cudaEventRecord(cx->ev_main, st)); CUDA_TRY(cudaStreamWaitEvent(cx->side, cx->ev_main, 0)); }
	}
	CUDA_TRY(cudaGetLastError());
	return 0;
}

static void set_ws(BtKParams &P, const Workspace &w) {
	P.rows = w.rows; P.elims = w.elims; P.frames = w.frames; P.partials = w.partials;
	P.R = w.R; P.FCAP = w.FCAP; P.PCAP = w.PCAP; P.stage = w.stage; P.stage_len = w.stage_len;
}

/* Enqueues one batch.  All pointers are device pointers; `maxlen` bounds the read length.
 *   main pass      on `st`:        every read, with a per-read transition budget and first-tier scratch (64 seedlings).  A read that
 *                                  exceeds the budget or fills its seedling list is SUSPENDED into a checkpoint slot (bt_ctxq.cuh).
 *   tail           on cx->side:    bt_tail_kernel (bt_tail.cu): the suspended reads circulate through a ring, a quantum of transitions
 *                                  per turn, always packed into as few single-warp blocks as there are reads left.  (BT_TAIL=restart:
 *                                  instead, over-budget reads are flagged and re-run from scratch by one unbudgeted pass.)
 *   overflow pass  on cx->side:    reads whose scratch overflowed even in a slot (or that found no free slot): re-run from scratch with
 *                                  worst-case scratch (normally empty)
 * The side stream lets the long tail of batch k overlap the main pass of batch k+1 (another context); the
 * batch is complete when cx->ev_tail has fired (bt_context_join / bt_context_sync). */
static int enqueue_align(bt_context *cx, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out, uint32_t maxlen, cudaStream_t st) {
	bt_index_t *ix = cx->ix;
	const uint32_t nwork = in->sel ? in->nsel : in->nreads;
	if (nwork == 0) return 0;
	if (maxlen < 1) maxlen = 1;
	if (policy_is_best(pol)) {
		if (maxlen > 1023) return fail("bt_align: reads longer than 1023 bases are not supported (the reference's Hit::mms is a FixedBitset<1024>)");
		return enqueue_best(cx, pol, in, out, maxlen, st);
	}
	if (maxlen > 1023) return fail("bt_align: reads longer than 1023 bases are not supported (the reference's Hit::mms is a FixedBitset<1024>)");
	const uint32_t nthreads = main_kernel_is_queue() ? (uint32_t)ix->sms * BT_Q_NCTX : (uint32_t)ix->sms * (uint32_t)ix->blocks_per_sm * BT_THREADS;
	const uint32_t stage_len = (maxlen + 15) & ~15u;                     /* every context keeps a writable copy of its read */
	const uint32_t mask_rows = (maxlen + 255) >> 8;                      /* per-frame live-position mask (bt_live_mask): 256 positions per row */
	if (ensure_ws(cx->ws1, nthreads, 6 * maxlen + 8 + 9 * mask_rows, 8, 64, stage_len)) return 1;
	/* checkpoint slots: one per suspended read.  On the bench workload 0.7 % of the reads exceed the main budget and 0.3 % fill the
	 * 64-seedling list; slots for 1/64 of the batch (BT_SLOT_DIV), at least 4096, each with room for 4096 seedlings.  A read that finds none is re-run by the overflow pass. */
	static const uint32_t slot_div = env_u32("BT_SLOT_DIV", 64), slot_pcap = env_u32("BT_SLOT_PCAP", 4096);
	/* The tail — the reads over the main budget.  Default ("restart"): they are flagged and re-run from scratch by one unbudgeted pass, one
	 * read per lane until it ends.  BT_TAIL=rr: they are suspended into checkpoint slots and finished by the round-robin tail (bt_tail.cu).
	 * Measured on the hg19-sized index (profiles/README.md, call 8): the round-robin tail executes 4.8 x fewer warp instructions (9.9 G
	 * against 47 G per million reads, 8.9 against 3.6 active threads per instruction) but a packed warp advances each of its reads ~3.5 x
	 * more slowly than a lone lane does, and a read is sequential: 2.4 s against 0.7 s for the tail of a batch, and with the contexts that
	 * fit in flight 2-4 M reads/s against 5.7-9.3 M. */
	static const bool use_slots = getenv("BT_TAIL") && getenv("BT_TAIL")[0] == 'r' && getenv("BT_TAIL")[1] == 'r' && !main_kernel_is_queue();
	uint32_t nslot = nwork / (slot_div ? slot_div : 64); if (nslot < 4096) nslot = 4096; if (nslot > nwork) nslot = nwork;
	if (use_slots) {
		if (ensure_ws(cx->wsl, nslot, 6 * maxlen + 8 + 17 * mask_rows, 16, slot_pcap, stage_len)) return 1;
		if (cx->slot_cap < nslot) {
			cudaFree(cx->slot_ctx); cudaFree(cx->tailq_items); cudaFree(cx->tailq); cx->slot_ctx = cx->tailq_items = nullptr; cx->tailq = nullptr; cx->slot_cap = 0;
			uint32_t cap = 8192; while (cap < 2 * cx->wsl.nthreads) cap <<= 1;                 /* ring: a power of two >= 2 x slots */
			CUDA_TRY(cudaMalloc((void **)&cx->slot_ctx, (size_t)cx->wsl.nthreads * BT_CTX_WORDS * 4));
			CUDA_TRY(cudaMalloc((void **)&cx->tailq_items, (size_t)cap * 4));
			CUDA_TRY(cudaMalloc((void **)&cx->tailq, sizeof(BtTailQ)));
			cx->tailq_cap = cap;
			cx->slot_cap = cx->wsl.nthreads;
		}
		nslot = cx->wsl.nthreads >= nslot ? nslot : cx->wsl.nthreads;
	} else {
		/* no slots (the default, or the queue kernel): the tail pass re-runs heavy reads from scratch on full-size scratch */
		static const uint32_t tail_bps = env_u32("BT_TAIL_BLOCKS", 2);
		if (ensure_ws(cx->wsh, (uint32_t)ix->sms * tail_bps * BT_THREADS, 6 * maxlen + 8 + 17 * mask_rows, 16, 4096, stage_len)) return 1;
	}
	const uint32_t nthreads2 = (uint32_t)ix->sms * 32;
	uint32_t R2 = maxlen * maxlen + 8 + (maxlen + 3) * mask_rows; if (R2 > 65000) R2 = 65000;   /* BtFrame::rowbase is 16 bits */
	if (ensure_ws(cx->ws2, nthreads2, R2, maxlen + 2, 4096, stage_len)) return 1;
	if (cx->retry_cap < nwork) {
		cudaFree(cx->retry_sel); cudaFree(cx->heavy_sel); cudaFree(cx->ultra_sel); cx->retry_sel = cx->heavy_sel = cx->ultra_sel = nullptr; cx->retry_cap = 0;
		CUDA_TRY(cudaMalloc((void **)&cx->retry_sel, (size_t)nwork * 4));
		CUDA_TRY(cudaMalloc((void **)&cx->heavy_sel, (size_t)nwork * 4));
		CUDA_TRY(cudaMalloc((void **)&cx->ultra_sel, (size_t)nwork * 4));
		cx->retry_cap = nwork;
	}
	BtKParams P; memset(&P, 0, sizeof P);
	P.ix[0] = ix->dev[0].dev; P.ix[1] = ix->dev[1].dev;
	memcpy(&P.pol, pol, sizeof(BtPolicy));
	bt_build_prog(pol->mode, pol->mms, pol->nofw, pol->norc, P.prog);
	P.seq = in->seq; P.qual = in->qual; P.roff = in->offs; P.seeds = in->seeds; P.sel = in->sel; P.nwork = nwork;
	P.found = out->found; P.flags = out->flags; P.hits = out->hits; P.slots = out->slots; P.mm_cap = out->mm_cap; P.rec_words = BT_HIT_HDR + out->mm_cap;
	P.stats = ix->stats; P.mask_rows = mask_rows;
	const uint32_t cblocks = (nwork + 255) / 256;
	BtWorkCtl *const ctl_main = cx->ctl, *const ctl_tail = cx->ctl + 1, *const ctl_ovf = cx->ctl + 2;   /* ctl_tail: the suspended reads' slot count (round-robin tail), or the restart pass's work list */
	/* this context's previous batch must have finished with the scratch and the lists */
	CUDA_TRY(cudaStreamWaitEvent(st, cx->ev_tail, 0));
	/* main pass */
	set_ws(P, cx->ws1);
	P.budget = main_budget();
	{ static const uint32_t db = env_u32("BT_DRAIN_BUDGET", 1500); P.drain_budget = db; }
	{ static uint32_t p = env_u32("BT_RARE_PERIOD", BT_RARE_PERIOD), t = env_u32("BT_RARE_THRESH", BT_RARE_THRESH); P.rare_period = p ? p : 1; P.rare_thresh = t; }
	bt_ctl_set_all_kernel<<<1, 32, 0, st>>>(cx->ctl, BT_CTL_WORDS, nwork);
	if (use_slots) {
		P.slot_ctx = cx->slot_ctx; P.slot_rows = cx->wsl.rows; P.slot_elims = cx->wsl.elims; P.slot_frames = cx->wsl.frames; P.slot_partials = cx->wsl.partials;
		P.slot_stage = cx->wsl.stage; P.nslot = nslot; P.slot_R = cx->wsl.R; P.slot_FCAP = cx->wsl.FCAP; P.slot_PCAP = cx->wsl.PCAP; P.slot_stage_len = cx->wsl.stage_len;
		P.resume = 0; P.slot_count = &ctl_tail[0].nwork;
	}
	if (main_kernel_is_queue()) {
		uint32_t grid = (uint32_t)ix->sms;
		const uint32_t need = (nwork + BT_Q_NCTX - 1) / BT_Q_NCTX;
		if (grid > need) grid = need;
		bt_search_kernel_q<<<grid, BT_Q_THREADS, bt_q_smem(BT_Q_NCTX), st>>>(P, ctl_main, BT_Q_NCTX);
	} else {
		uint32_t grid = (uint32_t)ix->sms * (uint32_t)ix->blocks_per_sm;
		const uint32_t need = (nwork + BT_THREADS - 1) / BT_THREADS;
		if (grid > need) grid = need;
		bt_search_kernel<<<grid, BT_THREADS, BT_THREADS * BT_SMEM_STRIDE, st>>>(P, ctl_main);
	}
	CUDA_TRY(cudaEventRecord(cx->ev_main, st));
	CUDA_TRY(cudaStreamWaitEvent(cx->side, cx->ev_main, 0));
	{ static uint32_t p = env_u32("BT_HEAVY_PERIOD", BT_RARE_PERIOD), t = env_u32("BT_HEAVY_THRESH", BT_RARE_THRESH); P.rare_period = p ? p : 1; P.rare_thresh = t; }
	if (use_slots) {
		/* the round-robin tail: single-warp blocks, as many as the suspended reads can fill (bt_tail.cu) */
		static const uint32_t quantum = env_u32("BT_TAIL_QUANTUM", 4096), wps = env_u32("BT_TAIL_WARPS", 16), wtarget = env_u32("BT_TAIL_WTARGET", 256), mincap = env_u32("BT_TAIL_MINCAP", 4);
		uint32_t blocks = (uint32_t)ix->sms * (wps ? wps : 16);
		{ const uint32_t need = (nslot + 31) / 32; if (blocks > need) blocks = need; }
		P.resume = 1; P.drain_budget = 0; P.budget = 0; P.sel = nullptr;
		P.R = cx->wsl.R; P.FCAP = cx->wsl.FCAP; P.PCAP = cx->wsl.PCAP; P.stage = nullptr; P.stage_len = cx->wsl.stage_len;
		P.rows = nullptr; P.elims = nullptr; P.frames = nullptr; P.partials = nullptr;
		if (bt_tail_launch(P, cx->tailq, &ctl_tail[0].nwork, nslot, cx->tailq_cap, cx->tailq_items, quantum ? quantum : 4096, wtarget, mincap, blocks, cx->side) != 0) return fail("bt_tail_launch failed");
		P.resume = 0; P.slot_ctx = nullptr;
		/* what is flagged now: scratch overflow in a slot, or no free slot */
		bt_collect_kernel<<<cblocks, 256, 0, cx->side>>>(out->flags, in->sel, nwork, nullptr, BT_FLAG_RETRY, cx->retry_sel, ctl_ovf);
	} else {
		bt_collect_kernel<<<cblocks, 256, 0, cx->side>>>(out->flags, in->sel, nwork, nullptr, BT_FLAG_RETRY, cx->heavy_sel, ctl_tail);
		set_ws(P, cx->wsh);
		P.sel = cx->heavy_sel; P.budget = 0; P.drain_budget = 0;
		{
			/* the tail's blocks may be smaller than the main pass's: a block lives as long as its slowest read, and what a straggler pins
			 * (registers and shared memory of its whole block) is what the other batches' main passes cannot use meanwhile */
			static const uint32_t tt = env_u32("BT_TAIL_THREADS", BT_THREADS);
			const uint32_t threads = (tt >= 32 && tt <= BT_THREADS && tt % 32 == 0) ? tt : BT_THREADS;
			bt_search_kernel<<<cx->wsh.nthreads / threads, threads, threads * BT_SMEM_STRIDE, cx->side>>>(P, ctl_tail);
		}
		bt_collect_kernel<<<cblocks, 256, 0, cx->side>>>(out->flags, cx->heavy_sel, nwork, ctl_tail, BT_FLAG_SCRATCH_OVF, cx->retry_sel, ctl_ovf);
	}
	P.sel = cx->retry_sel; P.budget = 0; P.drain_budget = 0;
	set_ws(P, cx->ws2);
	bt_search_kernel<<<ix->sms, 32, 32 * BT_SMEM_STRIDE, cx->side>>>(P, ctl_ovf);
	CUDA_TRY(cudaGetLastError());
	return 0;
}

/* Marks the batch complete on the side stream (after optional D2H copies enqueued there by the caller). */
static int finish_tail(bt_context *cx) { CUDA_TRY(cudaEventRecord(cx->ev_tail, cx->side)); return 0; }

extern "C" int bt_context_align_device(bt_context_t *cx, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out, void *stream) {
	if (!cx || !pol || !in || !out) return fail("bt_context_align_device: null argument");
	if (check_policy(cx->ix, pol)) return 1;
	if (in->max_len == 0) return fail("bt_context_align_device: max_len must be set (the offsets live on the device)");
	std::lock_guard<std::mutex> g(cx->mu);
	CUDA_TRY(cudaSetDevice(cx->ix->device));
	if (enqueue_align(cx, pol, in, out, in->max_len, (cudaStream_t)stream)) return 1;
	return finish_tail(cx);
}

/* Makes `stream` wait until the context's outstanding batch (including its heavy / overflow passes) is complete. */
extern "C" int bt_context_join(bt_context_t *cx, void *stream) {
	if (!cx) return fail("bt_context_join: null argument");
	CUDA_TRY(cudaSetDevice(cx->ix->device));
	CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, cx->ev_tail, 0));
	return 0;
}

extern "C" int bt_align_batch_device(bt_index_t *ix, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out, void *stream) {
	if (!ix) return fail("bt_align_batch_device: null argument");
	bt_context *cx = default_ctx(ix);
	if (!cx) return 1;
	if (bt_context_align_device(cx, pol, in, out, stream)) return 1;
	return bt_context_join(cx, stream);      /* simple entry point: complete when `stream` is */
}

template <typename T> static int grow(T **p, size_t &cap, size_t need) {
	if (cap >= need) return 0;
	cudaFree(*p); *p = nullptr; cap = 0;
	size_t n = need + need / 4 + 16;
	CUDA_TRY(cudaMalloc((void **)p, n * sizeof(T)));
	cap = n;
	return 0;
}

static int align_host(bt_context *cx, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out, void *stream, bool sync) {
	if (!cx || !pol || !in || !out) return fail("bt_align_batch: null argument");
	bt_index_t *ix = cx->ix;
	if (check_policy(ix, pol)) return 1;
	if (in->nreads == 0) return 0;
	if (!in->offs || !in->seeds || !out->found || !out->flags || !out->hits) return fail("bt_align_batch: null buffer");
	if ((!in->seq || !in->qual) && in->offs[in->nreads] != 0) return fail("bt_align_batch: null sequence buffer");
	if (out->slots == 0) return fail("bt_align_batch: slots must be >= 1");
	cudaStream_t st = (cudaStream_t)stream;
	uint32_t maxlen = 0;
	{
		std::lock_guard<std::mutex> g(cx->mu);
		CUDA_TRY(cudaSetDevice(ix->device));
		const uint32_t n = in->nreads;
		const size_t nb = (size_t)in->offs[n];
		for (uint32_t i = 0; i < n; i++) { uint64_t l = in->offs[i + 1] - in->offs[i]; if (l > maxlen) maxlen = (uint32_t)l; }
		const size_t rec_words = BT_HIT_HDR + out->mm_cap;
		const uint32_t nout = pol->paired ? n / 2 : n;                        /* results are per read, or per pair */
		const size_t hitwords = (size_t)nout * out->slots * rec_words;
		/* this context's previous batch (its heavy / overflow passes and D2H copies run on cx->side) must be finished before its
		 * staging buffers are overwritten: the wait comes before the first copy, not only before the kernels */
		CUDA_TRY(cudaStreamWaitEvent(st, cx->ev_tail, 0));
		if (grow(&cx->d_seq, cx->cap_seq, nb + 1) || grow(&cx->d_qual, cx->cap_qual, nb + 1)) return 1;
		if (grow(&cx->d_offs, cx->cap_offs, (size_t)n + 1) || grow(&cx->d_seeds, cx->cap_seeds, n) ||
		    grow(&cx->d_found, cx->cap_found, nout) || grow(&cx->d_flags, cx->cap_flags, nout)) return 1;
		if (grow(&cx->d_hits, cx->cap_hitwords, hitwords)) return 1;
		if (in->sel && grow(&cx->d_sel, cx->cap_sel, in->nsel)) return 1;
		if (nb) {
			CUDA_TRY(cudaMemcpyAsync(cx->d_seq, in->seq, nb, cudaMemcpyHostToDevice, st));
			CUDA_TRY(cudaMemcpyAsync(cx->d_qual, in->qual, nb, cudaMemcpyHostToDevice, st));
		}
		CUDA_TRY(cudaMemcpyAsync(cx->d_offs, in->offs, ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, st));
		CUDA_TRY(cudaMemcpyAsync(cx->d_seeds, in->seeds, (size_t)n * 4, cudaMemcpyHostToDevice, st));
		if (in->sel) CUDA_TRY(cudaMemcpyAsync(cx->d_sel, in->sel, (size_t)in->nsel * 4, cudaMemcpyHostToDevice, st));
		else { CUDA_TRY(cudaMemsetAsync(cx->d_found, 0, (size_t)nout * 4, st)); CUDA_TRY(cudaMemsetAsync(cx->d_flags, 0, (size_t)nout * 4, st)); }
		bt_read_batch_t din = *in; bt_hit_batch_t dout = *out;
		din.seq = cx->d_seq; din.qual = cx->d_qual; din.offs = cx->d_offs; din.seeds = cx->d_seeds; din.sel = in->sel ? cx->d_sel : nullptr;
		dout.found = cx->d_found; dout.flags = cx->d_flags; dout.hits = cx->d_hits;
		if (enqueue_align(cx, pol, &din, &dout, maxlen, st)) return 1;
		if (!in->sel) {
			/* results leave on the side stream, behind the heavy / overflow passes */
			CUDA_TRY(cudaMemcpyAsync(out->found, cx->d_found, (size_t)nout * 4, cudaMemcpyDeviceToHost, cx->side));
			CUDA_TRY(cudaMemcpyAsync(out->flags, cx->d_flags, (size_t)nout * 4, cudaMemcpyDeviceToHost, cx->side));
			CUDA_TRY(cudaMemcpyAsync(out->hits, cx->d_hits, hitwords * 4, cudaMemcpyDeviceToHost, cx->side));
			if (finish_tail(cx)) return 1;
			if (sync) CUDA_TRY(cudaStreamSynchronize(cx->side));
		} else {
			/* selection call: only the selected reads' entries are defined; copy them back one by one */
			if (finish_tail(cx)) return 1;
			CUDA_TRY(cudaStreamSynchronize(cx->side));
			for (uint32_t k = 0; k < in->nsel; k++) {
				uint32_t r = in->sel[k];
				CUDA_TRY(cudaMemcpy(out->found + r, cx->d_found + r, 4, cudaMemcpyDeviceToHost));
				CUDA_TRY(cudaMemcpy(out->flags + r, cx->d_flags + r, 4, cudaMemcpyDeviceToHost));
				CUDA_TRY(cudaMemcpy(out->hits + (size_t)r * out->slots * rec_words, cx->d_hits + (size_t)r * out->slots * rec_words, out->slots * rec_words * 4, cudaMemcpyDeviceToHost));
			}
		}
	}
	return 0;
}

extern "C" int bt_context_align(bt_context_t *cx, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out, void *stream) {
	return align_host(cx, pol, in, out, stream, true);
}
/* Enqueue-only variant: host buffers must be pinned and stay untouched until bt_context_sync(). Not for `sel` calls. */
extern "C" int bt_context_align_async(bt_context_t *cx, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out, void *stream) {
	if (in && in->sel) return fail("bt_context_align_async: selection calls are synchronous (use bt_context_align)");
	return align_host(cx, pol, in, out, stream, false);
}
extern "C" int bt_context_sync(bt_context_t *cx, void *stream) {
	if (!cx) return fail("bt_context_sync: null argument");
	CUDA_TRY(cudaSetDevice(cx->ix->device));
	CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
	CUDA_TRY(cudaStreamSynchronize(cx->side));
	return 0;
}
extern "C" int bt_align_batch(bt_index_t *ix, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out, void *stream) {
	if (!ix) return fail("bt_align_batch: null argument");
	bt_context *cx = default_ctx(ix);
	if (!cx) return 1;
	return align_host(cx, pol, in, out, stream, true);
}

extern "C" int bt_stats_get(bt_index_t *ix, bt_stats_t *out, int reset) {
	if (!ix || !out) return fail("bt_stats_get: null argument");
	CUDA_TRY(cudaSetDevice(ix->device));
	CUDA_TRY(cudaDeviceSynchronize());
	unsigned long long v[8];
	CUDA_TRY(cudaMemcpy(v, ix->stats, sizeof v, cudaMemcpyDeviceToHost));
	out->lfex = v[0]; out->lf = v[1]; out->chase = v[2]; out->ftab = v[3]; out->offs = v[4]; out->backtracks = v[5]; out->iters = v[6]; out->block_loads = v[7];
	if (reset) CUDA_TRY(cudaMemset(ix->stats, 0, sizeof v));
	return 0;
}

extern "C" void *bt_host_alloc(size_t bytes) {
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); fail("bt_host_alloc: cudaHostAlloc failed"); return nullptr; }
	return p;
}
extern "C" void bt_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" int bt_debug_lf(bt_index_t *ix, int mirror, const uint32_t *rows, uint32_t n, uint32_t *out) {
	if (!ix || !rows || !out) return fail("bt_debug_lf: null argument");
	if (mirror && !ix->has_mirror) return fail("bt_debug_lf: mirror index not loaded");
	if (n == 0) return 0;
	CUDA_TRY(cudaSetDevice(ix->device));
	uint32_t *d_rows = nullptr, *d_out = nullptr;
	CUDA_TRY(cudaMalloc((void **)&d_rows, (size_t)n * 4));
	CUDA_TRY(cudaMalloc((void **)&d_out, (size_t)n * 20));
	CUDA_TRY(cudaMemcpy(d_rows, rows, (size_t)n * 4, cudaMemcpyHostToDevice));
	bt_debug_lf_kernel<<<(n + 127) / 128, 128>>>(ix->dev[mirror ? 1 : 0].dev, d_rows, n, d_out);
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaMemcpy(out, d_out, (size_t)n * 20, cudaMemcpyDeviceToHost));
	cudaFree(d_rows); cudaFree(d_out);
	return 0;
}
