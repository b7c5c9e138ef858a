/*
 * bt_core.cuh — the per-read search state machine of the B200 kernels.
 *
 * One GPU thread owns one read at a time ("lane").  The depth-first backtracking search of
 * Bowtie 1 (GreedyDFSRangeSource, reference ebwt_search_backtrack.h:23-1779, driven by the
 * search_*.c phase fragments) is re-expressed as an explicit state machine with a frame stack
 * in global scratch, so that the 32 lanes of a warp share one instruction stream:
 *
 *     fast transitions   PC_LF (one query position: LF step + bookkeeping + next prologue)
 *                        PC_CHASE (one step of the row walk that resolves a hit)
 *                        -> every iteration; their rank-block loads are issued together (converged)
 *     rare transitions   everything else (phase program, backtrack target selection, frame
 *                        push/pop, hit reporting, ...) -> deferred and executed in batches
 *
 * Recursion in the reference becomes PUSH/POP of BtFrame records; "return into the caller" is a
 * continuation code.  All pseudo-random draws happen in exactly the reference's order, so the
 * output is bit-identical (see tests/).
 *
 * Code size matters as much as instruction count here: lanes of a warp sit in different states, so
 * the kernel jumps around its own code; every block of logic therefore exists exactly ONCE
 * (reporting, frame entry, the phase interpreter), reached through the state variable instead of
 * being inlined at each call site, to keep the kernel inside the SM's instruction cache.
 *
 * The rank structure is NOT the reference's side layout.  The .ebwt sides are re-laid-out at load
 * time (bt_relayout kernel) into 32-byte blocks covering 64 BWT rows each:
 *     u32 occ[4]   fchr[c] + #c in BWT[0, 64k)   ('$' excluded)
 *     u64 hi, lo   bit-planes of the 2-bit codes of rows 64k .. 64k+63
 * so one LF step is one aligned 32-byte sector and 2..6 POPCs.  LF(row, c) results are identical
 * to Ebwt::mapLF / mapLFEx / mapLF1 (reference ebwt.h:2334-2560).
 *
 * This header compiles for the device (nvcc) and, for the test-only logic emulation under
 * tests/host_emu/, for the host (g++ -DBT_HOST_EMU).  The product never runs the host build.
 */
#pragma once
#include <stdint.h>
#include "bt_prog.h"

#if defined(__CUDACC__)
#define BT_FN __device__ __forceinline__
#define BT_NOINLINE static __device__ __noinline__
#define BT_LDG(p) __ldg(p)
#define BT_POPC64(x) __popcll(x)
#define BT_POPC32(x) ((uint32_t)__popc(x))
#else
#define BT_FN static inline
#define BT_NOINLINE static
#define BT_LDG(p) (*(p))
#define BT_POPC64(x) __builtin_popcountll(x)
#define BT_POPC32(x) ((uint32_t)__builtin_popcount(x))
#ifndef BT_HOST_EMU
#error "bt_core.cuh is device code; the host build exists only for tests/host_emu (define BT_HOST_EMU)"
#endif
struct uint4 { uint32_t x, y, z, w; };
#endif

#define BT_OFF_MASK 0xffffffffu
#ifndef BT_ALT_FLAT
#define BT_ALT_FLAT 1              /* branch-free bookkeeping of a position's alternatives (bt_position) */
#endif

/* development aid (tools/pc_hist.py): trip counts of the per-lane loops inside the rare blocks, host emulation only */
#if defined(BT_HOST_EMU) && defined(BT_EMU_PROFILE)
extern unsigned long long bt_emu_prof[16];
#define BT_PROF(i, n) (bt_emu_prof[i] += (n))
#else
#define BT_PROF(i, n) ((void)0)
#endif

/* ---- device-resident index (one per orientation) -------------------------------------------- */
struct BtDevIndex {
	const uint4 *blocks;          /* 2 x uint4 per 64-row block                              */
	const uint32_t *offs;         /* SA sample (X.2.ebwt)                                      */
	const uint32_t *ftab;
	const uint32_t *eftab;
	const uint32_t *rstarts;      /* 3 words per fragment                                      */
	const uint32_t *plen;
	uint32_t len, zOff, nFrag, nPat, offMask;
	int32_t offRate, ftabChars;
	uint32_t fchr[5];
	uint32_t fw;                  /* 1 forward index, 0 mirror                                 */
};

/* ---- policy (the option globals that reach the hot path, ebwt_search.cpp:153-253) ------------ */
struct BtPolicy {
	int32_t mode;        /* 0: -v, 1: -n                                */
	int32_t mms;
	int32_t seedLen;
	uint32_t qualThresh; /* -e ; -v modes use 0xffffffff                */
	uint32_t maxBts;
	uint32_t khits, mhits;
	int32_t allHits, nofw, norc, maqRound;
	int32_t best, strata;   /* best-first ("stateful") path, bt_best.cuh    */
	uint32_t maxBtsBest;    /* its backtrack budget (maxBts, ebwt_search.cpp:186) */
	int32_t sampleMax;      /* -M: keep every hit up to the -m ceiling      */
	int32_t paired;         /* reads 2p, 2p+1 of the batch are mates 1, 2   */
	uint32_t minIns, maxIns;/* -I / -X (after trimming adjustments)          */
	int32_t mate1fw, mate2fw;   /* --fr: 1,0  --rf: 0,1  --ff: 1,1            */
	uint32_t pairTries;     /* --pairtries (mixedAttemptLim)                 */
};

/* flags written per read */
#define BT_FLAG_STACK_OVF 1u   /* frame-row scratch exhausted          */
#define BT_FLAG_FRAME_OVF 2u   /* more nested frames than FCAP         */
#define BT_FLAG_PART_OVF  4u   /* more seedlings than PCAP             */
#define BT_FLAG_HITS_OVF  8u   /* more reportable hits than slots      */
#define BT_FLAG_MM_OVF   16u   /* more mismatches than the record holds */
#define BT_FLAG_BUDGET  32u   /* (internal) iteration budget of the main pass exceeded: moved to the heavy pass */
#define BT_FLAG_PREEMPT 64u   /* (internal, transient) the read leaves this pass with its state: checkpointed into a slot (or back into the tail's ring), resumed later */
#define BT_FLAG_SCRATCH_OVF 7u
#define BT_FLAG_RETRY (BT_FLAG_SCRATCH_OVF | BT_FLAG_BUDGET)

/* hit record: BT_HIT_HDR header words followed by mm_cap mismatch words (pos | refc << 16) */
#define BT_HIT_HDR 5

struct BtFrame {               /* a suspended parent frame (64 bytes)                         */
	uint32_t top, bot, eligibleSz, eltop, elbot, btspread;
	uint16_t depth, d, unrevOff, oneRevOff, twoRevOff, threeRevOff, ham, altNum, eligibleNum, eli, rowbase, rowd0;
	uint16_t bt_i, mm_pos;
	uint8_t lowAltQual, elham, elcint, flags, bt_j, mm_refc, pad0, pad1;
	uint32_t pad2;
};
/* BtFrame.flags */
#define FF_ELIGNORE 1
#define FF_BDM 2
#define FF_MUST 4
#define FF_INVHH 8
#define FF_INVEXACT 16
#define FF_DISABLEFTAB 32

struct BtKParams {
	BtDevIndex ix[2];             /* [0] forward, [1] mirror                                  */
	BtPolicy pol;
	uint32_t prog[BT_PROG_MAX];   /* phase program for this policy (bt_prog.h)                */
	/* reads */
	const uint8_t *seq;           /* codes 0..4, concatenated                                 */
	const uint8_t *qual;          /* phred+33 chars, concatenated                             */
	const uint64_t *roff;         /* nreads+1 offsets                                         */
	const uint32_t *seeds;        /* Read::seed per read                                      */
	const uint32_t *sel;          /* optional list of read ids to process (re-runs), or NULL  */
	uint32_t nwork;               /* number of work items (len of sel, or nreads)             */
	/* outputs */
	uint32_t *found;              /* hitsForThisRead at finishRead                            */
	uint32_t *flags;
	uint32_t *hits;               /* nreads x slots x rec_words                               */
	uint32_t slots, mm_cap, rec_words;
	/* scratch (per thread) */
	uint4 *rows;                  /* R rows x 2 uint4                                         */
	uint8_t *elims;               /* R bytes                                                  */
	BtFrame *frames;              /* FCAP                                                     */
	uint64_t *partials;           /* PCAP                                                     */
	uint8_t *stage;               /* 2 * stage_len bytes: writable copy of the read (long reads) */
	uint32_t R, FCAP, PCAP, stage_len;
	uint32_t mask_rows;           /* rows (of 256 bits) that every frame reserves below its rowbase for its live-position mask (bt_live_mask) */
	uint32_t budget;              /* per-read transition budget of this pass (0 = unlimited)  */
	uint32_t drain_budget;        /* ... once the pass's work queue is empty (0 = the same)    */
	uint32_t rare_period, rare_thresh;   /* deferral of rare transitions in the thread-per-lane kernel */
	/* Checkpoint slots (bt_ctxq.cuh; BT_TAIL=rr): a read that exceeds the main pass's budget (or fills its seedling list) is not
	 * re-run, it is suspended — packed lane state, its copy of the read and its live scratch move into a slot — and the round-robin
	 * tail (bt_tail.cu) resumes it.  slot_ctx == NULL (the default): no slots, such reads are flagged and re-run from scratch by the tail pass. */
	uint32_t *slot_ctx;           /* BT_CTX_WORDS x nslot (word-major)                        */
	uint4 *slot_rows; uint8_t *slot_elims; BtFrame *slot_frames; uint64_t *slot_partials; uint8_t *slot_stage;
	uint32_t nslot, slot_R, slot_FCAP, slot_PCAP, slot_stage_len;
	uint32_t resume;              /* 1: the work items are slot ids to resume; 0: read ids     */
	unsigned long long *slot_count;    /* main pass: slots handed out so far (may run past nslot: those reads are flagged for a re-run) */
	unsigned long long *stats;    /* [8]: lfex, lf, chase, ftab, offs, backtracks, iters, blockloads */
};

/* ---- program counters ------------------------------------------------------------------------ */
enum {
	PC_LF = 0, PC_CHASE = 1,                      /* fast transitions                          */
	PC_PHASE, PC_BT_BEGIN, PC_FRAME_ENTER, PC_POS, PC_BTLOOP, PC_CHILD_RET, PC_POS_END, PC_FRAME_RET,
	PC_REPORT, PC_REPORT_ROW, PC_RESOLVE, PC_REPORT_RET, PC_BT_END,
	PC_FINISH_READ, PC_NEXT_READ, PC_EXIT
};
#define BT_IS_FAST(pc) ((pc) <= PC_CHASE)
#define BT_IS_RARE_STEP(pc) ((pc) > PC_CHASE && (pc) < PC_FINISH_READ)
enum { SITE_MAIN = 0, SITE_BT, SITE_END, SITE_FTABFULL };
enum { LFK_EX = 0, LFK_ONE, LFK_PAIR, LFK_FCHR, LFK_NONE };

/* The part of a lane's state that only the rare transitions touch (phase program, frame push / pop, reporting, seedlings).  It is
 * reached through BtLane::K so that the kernel can keep it out of the register file: bt_search_kernel points K into the lane's
 * shared-memory area (BT_COLD_SMEM), which takes the lane from 168 to <= 128 registers — one more resident block per SM — while the
 * fast transition (bt_fast_iter) never dereferences it on its common path.  Elsewhere K points at an ordinary local object. */
struct BtLaneCold {
	uint32_t rid, seed, found, hasN;                  /* read */
	uint32_t ph, done, step;                          /* control */
	uint32_t fw, reportExacts, maxBts, unrev0, rev1_0, iham;   /* backtracker object state */
	uint32_t nmuts, mut0, mut1, mut2;                 /* pos | newBase << 16 | oldBase << 24                 */
	uint32_t rnd, numBts, bailed;
	uint32_t depth, oneRevOff, twoRevOff, threeRevOff, disableFtab;   /* current frame */
	uint32_t bt_i, bt_j, bttop, btbot, btham;
	uint32_t rep_site, rep_sd, rep_cost, rep_stratum, rep_top, rep_bot, rep_r, rep_i;   /* report */
	uint32_t npart, pal_i;                            /* seedlings */
	uint32_t s_ftab, s_offs, s_bt;                    /* statistics */
};
#define BT_COLD_WORDS ((uint32_t)(sizeof(BtLaneCold) / 4))

struct BtLane {
	BtLaneCold *K;
	/* read */
	uint32_t rlen, flags;
	uint8_t *rseq, *rqual;         /* the lane's writable copy of the read (shared memory, or scratch for long reads); qualities are never written */
	/* control */
	uint32_t pc, ret, lfk, nit;    /* nit: transitions taken by the current read */
	/* backtracker object state */
	uint32_t ebwtSel, considerQuals, halfAndHalf, reportPartials, maqPenalty;
	uint32_t qualThresh;
	uint32_t qlen, depth5, depth3, rev2_0, rev3_0;
	uint32_t viewRev, viewComp;
	/* current frame */
	uint32_t stackDepth, d, unrevOff, ham;
	uint32_t altNum, eligibleNum, eligibleSz, eli, elignore, eltop, elbot, elham, elcint, lowAltQual;
	uint32_t rowbase, rowd0;
	uint32_t top, bot, ltop, lbot;   /* ltop/lbot: rows of the SideLocus pair (ebwt_search_backtrack.h:419-426,569-574) */
	uint32_t c, q, curIsAlt, curIsElig, curOverrides;
	uint32_t f_bdm, f_must, f_invHH, f_invExact;
	/* chase */
	uint32_t crow, cjumps;
	/* statistics */
	uint32_t s_lfex, s_lf, s_chase, s_iter, s_blk;
};

/* ---- small helpers --------------------------------------------------------------------------- */
BT_FN uint32_t bt_qual_round(uint32_t q) {                                                          /* qual.cpp:4-32: 0 / 10 / 20 / 30 by Phred bucket */
	const uint32_t r = ((q + 5u) / 10u) * 10u;                     /* branch-free: < 5 -> 0, 5..14 -> 10, 15..24 -> 20, >= 25 -> 30 */
	return r < 30u ? r : 30u;
}
BT_FN uint32_t bt_mm_penalty(uint32_t maq, uint32_t q) { return maq ? bt_qual_round(q) : q; }          /* qual.h:55-61 */
BT_FN uint32_t bt_rand_next(uint32_t &last) {                                                          /* random_source.h:45-54 */
	last = 1664525u * last + 1013904223u;
	uint32_t ret = last >> 16;
	last = 1664525u * last + 1013904223u;
	return ret ^ last;
}

/* query character / quality at offset `cur` of the current view (_qry / _qual of
 * GreedyDFSRangeSource::setQuery, ebwt_search_backtrack.h:90-99).  Seedling mutations are applied
 * to the lane's copy of the read itself (bt_apply_muts), as the reference does to the Read. */
BT_FN uint32_t bt_view_idx(const BtLane &L, uint32_t cur) { return L.viewRev ? (L.rlen - 1 - cur) : cur; }
BT_FN uint32_t bt_qry(const BtLane &L, uint32_t cur) {
	uint32_t c = L.rseq[bt_view_idx(L, cur)];
	if (L.viewComp && c < 4) c ^= 3;
	return c;
}
BT_FN uint32_t bt_qual_at(const BtLane &L, uint32_t cur) {
	uint32_t ch = L.rqual[bt_view_idx(L, cur)];
	return ch >= 33 ? ch - 33 : 0;                     /* phredCharToPhredQual qual.h:15-17 */
}
/* applyPartialMutations / undoPartialMutations (ebwt_search_backtrack.h:1368-1382, 1432-1446) on the
 * lane's copy; mut = viewpos | newBase << 16 | oldBase << 24, bases in view space. */
BT_FN void bt_put_base(BtLane &L, uint32_t cur, uint32_t base) {
	if (L.viewComp && base < 4) base ^= 3;
	L.rseq[bt_view_idx(L, cur)] = (uint8_t)base;
}
BT_FN void bt_apply_muts(BtLane &L, bool undo) {
	if (L.K->nmuts > 0) bt_put_base(L, L.K->mut0 & 0xffffu, undo ? (L.K->mut0 >> 24) : ((L.K->mut0 >> 16) & 0xff));
	if (L.K->nmuts > 1) bt_put_base(L, L.K->mut1 & 0xffffu, undo ? (L.K->mut1 >> 24) : ((L.K->mut1 >> 16) & 0xff));
	if (L.K->nmuts > 2) bt_put_base(L, L.K->mut2 & 0xffffu, undo ? (L.K->mut2 >> 24) : ((L.K->mut2 >> 16) & 0xff));
}

/* ---- rank blocks ----------------------------------------------------------------------------- */
struct BtBlock { uint4 occ; uint64_t hi, lo; };

BT_FN BtBlock bt_load_block(const BtDevIndex &ix, uint32_t row) {
	const uint4 *p = ix.blocks + 2 * (size_t)(row >> 6);
	BtBlock b;
	b.occ = BT_LDG(p);
	uint4 w = BT_LDG(p + 1);
	b.hi = (uint64_t)w.x | ((uint64_t)w.y << 32);
	b.lo = (uint64_t)w.z | ((uint64_t)w.w << 32);
	return b;
}
/* LF for one character: Ebwt::mapLF(l, c) */
BT_FN uint32_t bt_lf(const BtDevIndex &ix, const BtBlock &b, uint32_t row, uint32_t c) {
	uint32_t o = row & 63;
	uint64_t mask = (o == 0) ? 0ull : (~0ull >> (64 - o));
	uint64_t h = (c & 2) ? b.hi : ~b.hi, l = (c & 1) ? b.lo : ~b.lo;
	uint32_t n = (uint32_t)BT_POPC64(h & l & mask);
	uint32_t base = c == 0 ? b.occ.x : c == 1 ? b.occ.y : c == 2 ? b.occ.z : b.occ.w;
	if (c == 0) { uint32_t zo = ix.zOff - (row & ~63u); if (zo < o) n--; }     /* '$' is stored as A but not counted */
	return base + n;
}
/* LF for all four characters: one half of Ebwt::mapLFEx */
BT_FN void bt_lf_ex(const BtDevIndex &ix, const BtBlock &b, uint32_t row, uint32_t out[4]) {
	uint32_t o = row & 63;
	uint64_t mask = (o == 0) ? 0ull : (~0ull >> (64 - o));
	uint64_t h = b.hi & mask, l = b.lo & mask;
	uint32_t t = (uint32_t)BT_POPC64(h & l);
	uint32_t g = (uint32_t)BT_POPC64(h) - t;
	uint32_t cc = (uint32_t)BT_POPC64(l) - t;
	uint32_t a = o - t - g - cc;
	uint32_t zo = ix.zOff - (row & ~63u); if (zo < o) a--;
	out[0] = b.occ.x + a; out[1] = b.occ.y + cc; out[2] = b.occ.z + g; out[3] = b.occ.w + t;
}
BT_FN uint32_t bt_row_l(const BtBlock &b, uint32_t row) {
	uint32_t o = row & 63;
	return (uint32_t)(((b.hi >> o) & 1) << 1) | (uint32_t)((b.lo >> o) & 1);
}
/* ftabHi / ftabLo (ebwt.h:985-1034) */
BT_FN uint32_t bt_ftab_hi(const BtDevIndex &ix, uint32_t i) {
	uint32_t v = BT_LDG(ix.ftab + i);
	return v <= ix.len ? v : BT_LDG(ix.eftab + (v ^ BT_OFF_MASK) * 2 + 1);
}
BT_FN uint32_t bt_ftab_lo(const BtDevIndex &ix, uint32_t i) {
	uint32_t v = BT_LDG(ix.ftab + i);
	return v <= ix.len ? v : BT_LDG(ix.eftab + (v ^ BT_OFF_MASK) * 2);
}
/* Ebwt::joinedToTextOff (ebwt.h:2569-2629); returns false if the hit straddles a fragment boundary */
BT_FN bool bt_joined_to_text(const BtDevIndex &ix, uint32_t qlen, uint32_t off, uint32_t &tidx, uint32_t &textoff) {
	uint32_t top = 0, bot = ix.nFrag;
	for (;;) {
		uint32_t elt = top + ((bot - top) >> 1);
		uint32_t lower = BT_LDG(ix.rstarts + elt * 3);
		uint32_t upper = (elt == ix.nFrag - 1) ? ix.len : BT_LDG(ix.rstarts + (elt + 1) * 3);
		if (lower <= off) {
			if (upper > off) {
				if (off + qlen > upper) return false;
				tidx = BT_LDG(ix.rstarts + elt * 3 + 1);
				uint32_t fragoff = off - lower;
				if (!ix.fw) { fragoff = (upper - lower) - fragoff - 1; fragoff -= (qlen - 1); }
				textoff = fragoff + BT_LDG(ix.rstarts + elt * 3 + 2);
				return true;
			}
			top = elt;
		} else bot = elt;
	}
}

/* ---- frame scratch --------------------------------------------------------------------------- */
struct BtScratch { uint4 *rows; uint8_t *elims; BtFrame *frames; uint64_t *partials; };

BT_FN uint32_t bt_row_idx(const BtLane &L, uint32_t d) { return L.rowbase + (d - L.rowd0); }
BT_FN uint32_t bt_pair_top(const BtScratch &S, uint32_t ri, uint32_t c) { const uint32_t *p = (const uint32_t *)(S.rows + 2 * (size_t)ri); return p[c]; }
BT_FN uint32_t bt_pair_bot(const BtScratch &S, uint32_t ri, uint32_t c) { const uint32_t *p = (const uint32_t *)(S.rows + 2 * (size_t)ri); return p[4 + c]; }
BT_FN uint32_t bt_mm_pos(const BtScratch &S, uint32_t k) { return S.frames[k].mm_pos; }
/* The frame's live-position mask: bit (k - rowd0) is set for the positions k of the current frame that were recorded with at least one
 * alternative of non-zero width (the only positions the backtrack-target scan and the next-lowest-quality re-scan can ever select —
 * ebwt_search_backtrack.h:760-800, 1004-1058 walk ALL positions of the frame and skip the others one by one).  It lives in the
 * P.mask_rows rows below the frame's rowbase, so pushing / popping a frame, suspending a read and the lane state need nothing new. */
BT_FN uint32_t *bt_live_mask(const BtLane &L, const BtKParams &P, const BtScratch &S) { return (uint32_t *)(S.rows + 2 * (size_t)(L.rowbase - P.mask_rows)); }
BT_FN uint32_t bt_clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
	return (uint32_t)__clz((int)x);
#else
	return x ? (uint32_t)__builtin_clz(x) : 32u;
#endif
}

/* ---- the phase interpreter (program built by bt_build_prog): GET_READ .. search_*.c .. ----------
 * Leaves L.pc = PC_BT_BEGIN with a configured backtracker, or PC_FINISH_READ. */
BT_FN void bt_phase(BtLane &L, const BtKParams &P, const BtScratch &S) {
	const uint32_t len = L.rlen;
	const uint32_t s = P.pol.mode == 0 ? len : (uint32_t)P.pol.seedLen;
	const uint32_t SS = len < s ? len : s, S3 = SS >> 1, S5 = S3 + (SS & 1);
	for (;;) {
		if (L.K->done) { L.pc = PC_FINISH_READ; return; }
		const uint32_t st = P.prog[L.K->ph];
		const uint32_t kind = BTS_KIND(st);
		if (kind == BTK_END) { L.pc = PC_FINISH_READ; return; }
		if (kind == BTK_FILTER) {
			/* search_seeded_phase1.c:17-43: too short, or more Ns in the seed than seedMms */
			L.K->ph++;
			bool skip = len < 4;
			if (!skip && L.K->hasN) {
				uint32_t ns = 0;
#pragma unroll 1
				for (uint32_t i = 0; i < SS; i++) if (L.rseq[i] == 4) { if (++ns > (uint32_t)P.pol.mms) { skip = true; break; } }
			}
			if (skip) { L.pc = PC_FINISH_READ; return; }
			continue;
		}
		bool first = true;
		if (kind == BTK_SEEDLOOP) {
			/* search_seeded_phase3.c:25-55 / phase4.c:24-52: extend each seedling; setQuery() is called once
			 * before the loop, so the RNG state carries over from one seedling to the next */
			if (L.K->pal_i >= L.K->npart) { L.K->npart = 0; L.K->pal_i = 0; L.K->nmuts = 0; L.K->ph++; continue; }
			first = (L.K->pal_i == 0);
		} else L.K->ph++;
		/* constructor arguments + setQuery + setOffs + setQlen + setReportExacts of one GreedyDFSRangeSource */
		L.K->step = st;
		L.ebwtSel = BTS_EBWT(st); L.K->fw = BTS_FW(st); L.considerQuals = BTS_CQ(st); L.halfAndHalf = BTS_HH(st);
		L.reportPartials = BTS_RP(st) ? (uint32_t)P.pol.mms : 0u; L.K->reportExacts = BTS_RE(st);
		L.qlen = BTS_SEEDQ(st) ? SS : len;
		uint32_t v[6];
#pragma unroll
		for (int k = 0; k < 6; k++) { uint32_t sel = BTS_SEL(st, k); v[k] = sel == BTV_0 ? 0u : sel == BTV_LEN ? len : sel == BTV_S ? SS : sel == BTV_S3 ? S3 : S5; }
		L.depth5 = v[0]; L.depth3 = v[1]; L.K->unrev0 = v[2]; L.K->rev1_0 = v[3]; L.rev2_0 = v[4]; L.rev3_0 = v[5];
		L.K->iham = 0; L.K->nmuts = 0;
		if (first) L.K->rnd = L.K->seed;                         /* setQuery: _rand.init(r.seed) */
		L.viewRev = (L.ebwtSel == 0) ? !L.K->fw : L.K->fw;
		L.viewComp = !L.K->fw;
		if (BTS_CLEARP(st)) L.K->npart = 0;
		if (kind == BTK_SEEDLOOP) {
			/* PartialAlignmentManager::toMutsString (ebwt_search_util.h:299-357) */
			const uint64_t pal = S.partials[L.K->pal_i++];
			uint32_t oldQuals = 0;
#pragma unroll 1
			for (uint32_t k = 0; k < 3; k++) {
				uint32_t pos = (uint32_t)(pal >> (16 * k)) & 0xffffu;
				if (pos == 0xffffu) break;
				uint32_t chr = (uint32_t)(pal >> (48 + 2 * k)) & 3u;
				uint32_t tpos = (L.rlen - 1 - pos) & 0xffffu;
				oldQuals = (oldQuals + bt_mm_penalty(L.maqPenalty, bt_qual_at(L, tpos))) & 0xff;
				uint32_t mv = tpos | (chr << 16) | (bt_qry(L, tpos) << 24);
				if (k == 0) L.K->mut0 = mv; else if (k == 1) L.K->mut1 = mv; else L.K->mut2 = mv;
				L.K->nmuts = k + 1;
			}
			L.K->iham = oldQuals;
			bt_apply_muts(L, false);                       /* setMuts(&muts) */
		}
		L.pc = PC_BT_BEGIN;
		return;
	}
}

/* Mismatches so far per seed half (the loops of ebwt_search_backtrack.h:691-699, 1245-1252): lo | hi << 16.
 * Out of line on purpose: rare, and it must not be replicated into the hot position code. */
BT_NOINLINE uint32_t bt_half_counts(const BtFrame *frames, uint32_t stackDepth, uint32_t qlen, uint32_t depth5, uint32_t depth3) {
	uint32_t lo = 0, hi = 0;
#pragma unroll 1
	for (uint32_t i = 0; i < stackDepth; i++) {
		const uint32_t dd = qlen - frames[i].mm_pos - 1;
		if (dd < depth5) hi++; else if (dd < depth3) lo++;
	}
	return lo | (hi << 16);
}

/* hhCheckTop (ebwt_search_backtrack.h:1200-1275) */
BT_FN bool bt_hh_check_top(const BtLane &L, const BtScratch &S) {
	if (L.d == L.depth5) {
		if (L.stackDepth == 0) return false;
	} else if (L.d == L.depth3) {
		if (L.rev3_0 == L.rev2_0) { if (L.stackDepth < 2) return false; }
		else if ((bt_half_counts(S.frames, L.stackDepth, L.qlen, L.depth5, L.depth3) & 0xffffu) == 0) return false;
	}
	return true;
}

/* reportPartial (ebwt_search_backtrack.h:1571-1655): seedling = up to 3 (pos, char) pairs.  Returns the
 * overflow flag (0 or BT_FLAG_PART_OVF). */
BT_NOINLINE uint32_t bt_store_partial(const BtFrame *frames, uint64_t *partials, uint32_t npart, uint32_t pcap, uint32_t sd) {
	uint64_t al = 0xffffffffffffull;                 /* pos0..2 = 0xffff, chars 0 */
#pragma unroll 1
	for (uint32_t k = 0; k < sd && k < 3; k++) {
		al &= ~(0xffffull << (16 * k));
		al |= (uint64_t)(frames[k].mm_pos & 0xffffu) << (16 * k);
		al |= (uint64_t)(frames[k].mm_refc & 3u) << (48 + 2 * k);
	}
	if (npart < pcap) { partials[npart] = al; return 0; }
	return BT_FLAG_PART_OVF;
}
BT_FN void bt_report_partial(BtLane &L, const BtKParams &P, const BtScratch &S, uint32_t sd) {
	L.flags |= bt_store_partial(S.frames, S.partials, L.K->npart, P.PCAP, sd);
	L.K->npart++;
}

/* Position prologue: the part of the while-loop body before the LF step (ebwt_search_backtrack.h:472-529),
 * given the query character and quality of position L.d.  Leaves L.pc = PC_LF when rank blocks are needed. */
BT_FN void bt_prologue(BtLane &L, uint32_t c, uint32_t q) {
	L.c = c; L.q = q;
	L.curIsElig = 0; L.curOverrides = 0;
	L.curIsAlt = (L.d >= L.unrevOff) && (!L.considerQuals || (L.ham + bt_mm_penalty(L.maqPenalty, q) <= L.qualThresh));
	if (L.curIsAlt) {
		if (L.considerQuals) {
			if (q < L.lowAltQual) { L.curIsElig = 1; L.curOverrides = 1; }
			else if (q == L.lowAltQual) L.curIsElig = 1;
		} else L.curIsElig = 1;
	}
	if (c == 4 && L.d > 0) L.top = L.bot = 1;
	L.pc = PC_LF;
	if (L.top == 0 && L.bot == 0) L.lfk = LFK_FCHR;
	else if (L.curIsAlt) L.lfk = LFK_EX;
	else if (c < 4) L.lfk = (L.top + 1 == L.bot) ? LFK_ONE : LFK_PAIR;
	else L.lfk = LFK_NONE;
}

/* One query position: the bookkeeping after the LF step (ebwt_search_backtrack.h:569-739) and, on a
 * plain match, the tail of the loop body (1066-1078) plus the next position's prologue.
 * tops/bots hold the quartet for LFK_EX / LFK_FCHR. */
BT_FN void bt_position(BtLane &L, const BtKParams &P, const BtScratch &S,
                       const uint32_t tops[4], const uint32_t bots[4], uint32_t nc, uint32_t nq) {
	const uint32_t c = L.c, q = L.q, d = L.d;
	const uint32_t cur = L.qlen - d - 1;
	if (L.top != L.bot) { L.ltop = L.top; L.lbot = L.bot; }   /* SideLocus::initFromTopBot */
	if (d >= L.rowd0) {
		const uint32_t ri = bt_row_idx(L, d);
		uint32_t el = (c < 4) ? (1u << c) : 0u;                   /* eliminate() */
#if BT_ALT_FLAT
		if (L.curIsAlt) {
			/* the per-character loop of ebwt_search_backtrack.h:603-653 without its branches: the live alternatives as a bit mask, their
			 * number and total width; the first live one becomes the recorded eligible edit when this position overrides */
			uint4 tv = { tops[0], tops[1], tops[2], tops[3] }, bv = { bots[0], bots[1], bots[2], bots[3] };
			S.rows[2 * (size_t)ri] = tv; S.rows[2 * (size_t)ri + 1] = bv;
			const uint32_t s0 = bots[0] - tops[0], s1 = bots[1] - tops[1], s2 = bots[2] - tops[2], s3 = bots[3] - tops[3];
			const uint32_t nz = ((uint32_t)(s0 != 0) | ((uint32_t)(s1 != 0) << 1) | ((uint32_t)(s2 != 0) << 2) | ((uint32_t)(s3 != 0) << 3)) & ~el;
			el = ~nz & 15u;
			const uint32_t n = BT_POPC32(nz);
			L.altNum += n;
			if (n) { uint32_t *lm = bt_live_mask(L, P, S) + ((d - L.rowd0) >> 5); *lm |= 1u << ((d - L.rowd0) & 31u); }
			if (L.curIsElig && n) {
				if (L.curOverrides) {
					const uint32_t f = (nz & 1u) ? 0u : (nz & 2u) ? 1u : (nz & 4u) ? 2u : 3u;
					L.lowAltQual = q; L.eligibleNum = 0; L.eligibleSz = 0; L.curOverrides = 0;
					L.eli = d; L.eltop = f == 0 ? tops[0] : f == 1 ? tops[1] : f == 2 ? tops[2] : tops[3];
					L.elbot = f == 0 ? bots[0] : f == 1 ? bots[1] : f == 2 ? bots[2] : bots[3];
					L.elham = bt_mm_penalty(L.maqPenalty, q); L.elcint = f; L.elignore = 0;
				}
				L.eligibleSz += ((nz & 1u) ? s0 : 0u) + ((nz & 2u) ? s1 : 0u) + ((nz & 4u) ? s2 : 0u) + ((nz & 8u) ? s3 : 0u);
				L.eligibleNum += n;
			}
		}
#else
		if (L.curIsAlt) {
			uint4 tv = { tops[0], tops[1], tops[2], tops[3] }, bv = { bots[0], bots[1], bots[2], bots[3] };
			S.rows[2 * (size_t)ri] = tv; S.rows[2 * (size_t)ri + 1] = bv;
#pragma unroll
			for (uint32_t i = 0; i < 4; i++) {
				if (i == c) continue;
				const uint32_t ptop = tops[i], pbot = bots[i];
				const uint32_t spread = pbot - ptop;
				if (spread == 0) el |= (1u << i);
				else {
					if (L.curIsElig) {
						if (L.curOverrides) {
							L.lowAltQual = q; L.eligibleNum = 0; L.eligibleSz = 0; L.curOverrides = 0;
							L.eli = d; L.eltop = ptop; L.elbot = pbot; L.elham = bt_mm_penalty(L.maqPenalty, q);
							L.elcint = i; L.elignore = 0;
						}
						L.eligibleSz += spread; L.eligibleNum++;
					}
					L.altNum++;
				}
			}
			if (el != 15u) { uint32_t *lm = bt_live_mask(L, P, S) + ((d - L.rowd0) >> 5); *lm |= 1u << ((d - L.rowd0) & 31u); }
		}
#endif
		S.elims[ri] = (uint8_t)el;
	}
	L.f_bdm = 0; L.f_must = 0; L.f_invHH = 0; L.f_invExact = 0;
	uint32_t reportedPartial = 0;
	if (cur == 0) {
		if (L.top < L.bot && L.stackDepth < L.reportPartials && L.reportPartials > 0) {
			if (L.altNum > 0) L.f_bdm = 1;
			if (L.stackDepth > 0) { bt_report_partial(L, P, S, L.stackDepth); reportedPartial = 1; }
		}
		if (L.stackDepth == 0 && L.bot > L.top && !L.K->reportExacts) { L.f_invExact = 1; L.f_bdm = 1; }
	}
#ifndef BT_MULTI_EXIT
	/* one exit: every outcome sets `npc`, the lane state is written once at the end (the form with a `return` per outcome made the
	 * compiler merge ~15 differently allocated copies of the lane into the loop header, a third of the fast path's instructions) */
	uint32_t npc = PC_LF;
	if (L.halfAndHalf) {
		if ((d == (L.depth5 - 1)) && L.top < L.bot) {
			L.f_invHH = (L.stackDepth == 0);
			if (L.stackDepth == 0 && L.altNum > 0) { L.f_bdm = 1; L.f_must = 1; }
			else if (L.stackDepth == 0) npc = PC_FRAME_RET;
		} else if ((d == (L.depth3 - 1)) && L.top < L.bot) {
			const uint32_t lh = bt_half_counts(S.frames, L.stackDepth, L.qlen, L.depth5, L.depth3);
			L.f_invHH = ((lh & 0xffffu) == 0 || (lh >> 16) == 0);
			if ((L.stackDepth < 2 || L.f_invHH) && L.altNum > 0) { L.f_must = 1; L.f_bdm = 1; }
			else if (L.stackDepth < 2) npc = PC_FRAME_RET;
		}
	}
	if (npc == PC_LF) {
		if (cur == 0 && L.bot > L.top && !L.f_invHH && !L.f_invExact && !reportedPartial) {
			/* reportAlignment(stackDepth, top, bot, ham) */
			L.K->rep_sd = L.stackDepth; L.K->rep_top = L.top; L.K->rep_bot = L.bot; L.K->rep_cost = L.ham; L.K->rep_site = SITE_MAIN;
			npc = PC_REPORT;
		} else if ((L.top == L.bot || L.f_bdm) && L.altNum > 0) npc = PC_BTLOOP;          /* mismatch with alternatives */
		else if (L.f_must || L.f_invHH || L.f_invExact || L.top == L.bot) npc = PC_FRAME_RET;
		else {
			L.d = d + 1;
			if (L.d >= L.qlen) npc = PC_POS;
			else if (L.halfAndHalf && !bt_hh_check_top(L, S)) npc = PC_FRAME_RET;
		}
	}
	if (npc == PC_FRAME_RET) L.ret = 0;
	L.pc = npc;
	if (npc == PC_LF) bt_prologue(L, nc, nq);
#else
	if (L.halfAndHalf) {
		if ((d == (L.depth5 - 1)) && L.top < L.bot) {
			L.f_invHH = (L.stackDepth == 0);
			if (L.stackDepth == 0 && L.altNum > 0) { L.f_bdm = 1; L.f_must = 1; }
			else if (L.stackDepth == 0) { L.ret = 0; L.pc = PC_FRAME_RET; return; }
		} else if ((d == (L.depth3 - 1)) && L.top < L.bot) {
			const uint32_t lh = bt_half_counts(S.frames, L.stackDepth, L.qlen, L.depth5, L.depth3);
			L.f_invHH = ((lh & 0xffffu) == 0 || (lh >> 16) == 0);
			if ((L.stackDepth < 2 || L.f_invHH) && L.altNum > 0) { L.f_must = 1; L.f_bdm = 1; }
			else if (L.stackDepth < 2) { L.ret = 0; L.pc = PC_FRAME_RET; return; }
		}
	}
	if (cur == 0 && L.bot > L.top && !L.f_invHH && !L.f_invExact && !reportedPartial) {
		/* reportAlignment(stackDepth, top, bot, ham) */
		L.K->rep_sd = L.stackDepth; L.K->rep_top = L.top; L.K->rep_bot = L.bot; L.K->rep_cost = L.ham; L.K->rep_site = SITE_MAIN;
		L.pc = PC_REPORT;
		return;
	}
	if ((L.top == L.bot || L.f_bdm) && L.altNum > 0) { L.pc = PC_BTLOOP; return; }   /* mismatch with alternatives */
	if (L.f_must || L.f_invHH || L.f_invExact) { L.ret = 0; L.pc = PC_FRAME_RET; return; }
	if (L.top == L.bot) { L.ret = 0; L.pc = PC_FRAME_RET; return; }
	L.d = d + 1;
	if (L.d >= L.qlen) { L.pc = PC_POS; return; }
	if (L.halfAndHalf && !bt_hh_check_top(L, S)) { L.ret = 0; L.pc = PC_FRAME_RET; return; }
	bt_prologue(L, nc, nq);
#endif
}

/* The rare transitions, one block each.  Every block exists once; call sites communicate through lane fields.  A block leaves
 * L.pc at the next state; `break` inside a block ends it. */
BT_FN void bt_blk_phase(BtLane &L, const BtKParams &P, const BtScratch &S) {
	const BtDevIndex &ix = P.ix[L.ebwtSel]; (void)ix;
	do {
		bt_phase(L, P, S);
		break;
	} while (0);
}

BT_FN void bt_blk_bt_begin(BtLane &L, const BtKParams &P, const BtScratch &S) {
	const BtDevIndex &ix = P.ix[L.ebwtSel]; (void)ix;
	do { {
		/* backtrack(ham) (ebwt_search_backtrack.h:237-297) */
		const uint32_t ftabChars = (uint32_t)ix.ftabChars;
		L.K->numBts = 0; L.K->bailed = 0;
		/* tallyNs (1308-1341) */
		uint32_t nsInSeed = 0, nsInFtab = 0; bool ok = true;
		if (L.K->hasN) {
#pragma unroll 1
			for (uint32_t i = 0; i < L.rev3_0 && ok; i++) {
				if (bt_qry(L, L.qlen - i - 1) == 4) {
					nsInSeed++;
					if (nsInSeed == 1) { if (i < L.K->unrev0) ok = false; }
					else if (nsInSeed == 2) { if (i < L.K->rev1_0) ok = false; }
					else if (nsInSeed == 3) { if (i < L.rev2_0) ok = false; }
					else ok = false;
				}
			}
#pragma unroll 1
			for (uint32_t i = 0; ok && i < ftabChars && i < L.qlen; i++) if (bt_qry(L, L.qlen - i - 1) == 4) nsInFtab++;
		}
		if (!ok) { L.ret = 0; L.pc = PC_BT_END; break; }
		/* the new root frame: backtrack(0, depth, _unrevOff, _1revOff, _2revOff, _3revOff, top, bot, iham, iham, ...) */
		L.stackDepth = 0; L.K->depth = 0; L.unrevOff = L.K->unrev0; L.K->oneRevOff = L.K->rev1_0; L.K->twoRevOff = L.rev2_0; L.K->threeRevOff = L.rev3_0;
		L.top = 0; L.bot = 0; L.ham = L.K->iham; L.rowbase = P.mask_rows; L.K->disableFtab = nsInFtab > 0;
		L.pc = PC_FRAME_ENTER;
		const uint32_t mlim = L.K->unrev0 < L.qlen ? L.K->unrev0 : L.qlen;
		if (nsInFtab == 0 && mlim >= ftabChars) {
			uint32_t ftabOff = 0;                                             /* calcFtabOff (1348-1362) */
#pragma unroll 1
			for (uint32_t i = ftabChars; i > 0; i--) ftabOff = (ftabOff << 2) | bt_qry(L, L.qlen - i);
			const uint32_t top = bt_ftab_hi(ix, ftabOff), bot = bt_ftab_lo(ix, ftabOff + 1);
			L.K->s_ftab++;
			if (L.qlen == ftabChars && bot > top) {
				if (L.reportPartials == 0) {
					L.K->rep_sd = 0; L.K->rep_top = top; L.K->rep_bot = bot; L.K->rep_cost = L.K->iham; L.K->rep_site = SITE_FTABFULL;
					L.pc = PC_REPORT;
				}
			} else if (bot > top) { L.K->depth = ftabChars; L.top = top; L.bot = bot; }
			else { L.ret = 0; L.pc = PC_BT_END; }
		}
		break; }
	} while (0);
}

BT_FN void bt_blk_frame_enter(BtLane &L, const BtKParams &P, const BtScratch &S) {
	const BtDevIndex &ix = P.ix[L.ebwtSel]; (void)ix;
	do { {
		/* the head of backtrack(stackDepth, depth, ...) up to the while loop (ebwt_search_backtrack.h:363-455);
		 * the caller has filled stackDepth, depth, the rev offsets, top, bot, ham, rowbase, disableFtab */
		L.rowd0 = L.K->depth > L.unrevOff ? L.K->depth : L.unrevOff;
		if (L.top != 0 || L.bot != 0) { L.ltop = L.top; L.lbot = L.bot; }
		if (L.stackDepth > 0) L.K->s_bt++;
		if (L.rowd0 < L.qlen && L.rowbase + (L.qlen - L.rowd0) > P.R) { L.flags |= BT_FLAG_STACK_OVF; break; }
		if (L.rowd0 < L.qlen) {
			const uint4 z = { 0, 0, 0, 0 };
			uint4 *m = S.rows + 2 * (size_t)(L.rowbase - P.mask_rows);
			for (uint32_t k = 0; k < 2 * P.mask_rows; k++) m[k] = z;
		}
		if (L.halfAndHalf) {
			if (L.K->maxBts > 0 && L.K->numBts == L.K->maxBts) { L.K->bailed = 1; L.ret = 0; L.pc = PC_FRAME_RET; break; }
			L.K->numBts++;
		}
		L.altNum = 0; L.eligibleNum = 0; L.eligibleSz = 0; L.eli = 0; L.elignore = 1; L.eltop = 0; L.elbot = 0;
		L.elham = L.ham; L.elcint = 0; L.lowAltQual = 0xff; L.d = L.K->depth;
		L.pc = PC_POS;
	}
	} while (0);
}

BT_FN void bt_blk_pos(BtLane &L, const BtKParams &P, const BtScratch &S) {
	const BtDevIndex &ix = P.ix[L.ebwtSel]; (void)ix;
	do { {
		/* top of while(cur < _qlen) (ebwt_search_backtrack.h:456-529) with its own query loads */
		if (L.d >= L.qlen) {
			if (L.stackDepth >= L.reportPartials) {
				L.K->rep_sd = L.stackDepth; L.K->rep_top = L.top; L.K->rep_bot = L.bot; L.K->rep_cost = L.ham; L.K->rep_site = SITE_END;
				L.pc = PC_REPORT;
			} else { L.ret = 0; L.pc = PC_FRAME_RET; }
			break;
		}
		if (L.halfAndHalf && !bt_hh_check_top(L, S)) { L.ret = 0; L.pc = PC_FRAME_RET; break; }
		const uint32_t cur = L.qlen - L.d - 1;
		bt_prologue(L, bt_qry(L, cur), bt_qual_at(L, cur));
		break; }
	} while (0);
}

BT_FN void bt_blk_btloop(BtLane &L, const BtKParams &P, const BtScratch &S) {
	const BtDevIndex &ix = P.ix[L.ebwtSel]; (void)ix;
	do { {
		/* while((top == bot || backtrackDespiteMatch) && altNum > 0) (ebwt_search_backtrack.h:743-971) */
		if (!((L.top == L.bot || L.f_bdm) && L.altNum > 0)) { L.pc = PC_POS_END; break; }
		uint32_t i = L.d, j = 0, bttop = 0, btbot = 0, btham = L.ham, btcint = 0;
		BT_PROF(0, 1);
		if (L.eligibleNum > 1 || L.elignore) {
			BT_PROF(1, 1);
			/* the highest position <= d of this frame whose quality is the lowest eligible one and that still has an alternative: the reference
			 * steps down from d one position at a time (760-800); here only the positions of the live mask are visited, highest first */
			const uint32_t *lm = bt_live_mask(L, P, S);
			int32_t w = (int32_t)((L.d - L.rowd0) >> 5);
			uint32_t mw = lm[w];
#pragma unroll 1
			for (;;) {
				while (mw == 0 && w > 0) mw = lm[--w];
				if (mw == 0) break;                     /* cannot happen while eligibleNum > 0 */
				const uint32_t b = 31u - bt_clz32(mw);
				mw &= ~(1u << b);
				i = L.rowd0 + ((uint32_t)w << 5) + b;
				BT_PROF(2, 1);
				const uint32_t qi = bt_qual_at(L, L.qlen - i - 1);
				const uint32_t ri = bt_row_idx(L, i);
				const uint32_t el = S.elims[ri];
				if (el != 15) BT_PROF(6, 1);
				if ((qi == L.lowAltQual || !L.considerQuals) && el != 15) {
					uint32_t posSz = 0;
					for (j = 0; j < 4; j++) if ((el & (1u << j)) == 0) posSz += bt_pair_bot(S, ri, j) - bt_pair_top(S, ri, j);
					uint32_t r = bt_rand_next(L.K->rnd) % posSz;
					for (j = 0; j < 4; j++) {
						if ((el & (1u << j)) == 0) {
							const uint32_t ptop = bt_pair_top(S, ri, j), pbot = bt_pair_bot(S, ri, j);
							const uint32_t spread = pbot - ptop;
							if (r < spread) { bttop = ptop; btbot = pbot; btham += bt_mm_penalty(L.maqPenalty, qi); btcint = j; break; }
							r -= spread;
						}
					}
					break;
				}
			}
		} else {
			i = L.eli; bttop = L.eltop; btbot = L.elbot; btham += L.elham; j = L.elcint; btcint = L.elcint;
		}
		const uint32_t icur = L.qlen - i - 1;
		uint32_t btUnrevOff = L.unrevOff, btOneRevOff = L.K->oneRevOff, btTwoRevOff = L.K->twoRevOff;
		const uint32_t btThreeRevOff = L.K->threeRevOff;
		if (i < L.K->oneRevOff) { btUnrevOff = L.K->oneRevOff; btOneRevOff = L.K->twoRevOff; btTwoRevOff = L.K->threeRevOff; }
		else if (i < L.K->twoRevOff) { btOneRevOff = L.K->twoRevOff; btTwoRevOff = L.K->threeRevOff; }
		else if (i < L.K->threeRevOff) { btTwoRevOff = L.K->threeRevOff; }
		if (L.stackDepth >= P.FCAP) { L.flags |= BT_FLAG_FRAME_OVF; break; }
		BtFrame &F = S.frames[L.stackDepth];
		F.mm_pos = (uint16_t)icur; F.mm_refc = (uint8_t)btcint;       /* _mms[stackDepth], _refcs[stackDepth] */
		L.K->bt_i = i; L.K->bt_j = j; L.K->bttop = bttop; L.K->btbot = btbot; L.K->btham = btham;
		if (i + 1 == L.qlen) {
			L.K->rep_sd = L.stackDepth + 1; L.K->rep_top = bttop; L.K->rep_bot = btbot; L.K->rep_cost = btham; L.K->rep_site = SITE_BT;
			L.pc = PC_REPORT;
			break;
		}
		const bool rejump = L.halfAndHalf && !L.K->disableFtab && L.rev2_0 == L.rev3_0 && i + 1 < (uint32_t)ix.ftabChars && (uint32_t)ix.ftabChars <= L.depth5;
		uint32_t ndepth = i + 1, ntop = bttop, nbot = btbot;
		if (rejump) {
			/* ftab re-jump with the substituted character (ebwt_search_backtrack.h:908-952) */
			const uint32_t ftabChars = (uint32_t)ix.ftabChars;
			uint32_t ftabOff = 0;
#pragma unroll 1
			for (uint32_t jj = ftabChars; jj > 0; jj--) ftabOff = (ftabOff << 2) | ((L.qlen - jj == icur) ? btcint : bt_qry(L, L.qlen - jj));
			ntop = bt_ftab_hi(ix, ftabOff); nbot = bt_ftab_lo(ix, ftabOff + 1);
			L.K->s_ftab++;
			ndepth = ftabChars;
			if (ntop == nbot) { L.ret = 0; L.pc = PC_CHILD_RET; break; }
		}
		/* PUSH: suspend this frame, then set up the callee */
		F.top = L.top; F.bot = L.bot; F.eligibleSz = L.eligibleSz; F.eltop = L.eltop; F.elbot = L.elbot; F.btspread = btbot - bttop;
		F.depth = (uint16_t)L.K->depth; F.d = (uint16_t)L.d; F.unrevOff = (uint16_t)L.unrevOff; F.oneRevOff = (uint16_t)L.K->oneRevOff;
		F.twoRevOff = (uint16_t)L.K->twoRevOff; F.threeRevOff = (uint16_t)L.K->threeRevOff; F.ham = (uint16_t)L.ham; F.altNum = (uint16_t)L.altNum;
		F.eligibleNum = (uint16_t)L.eligibleNum; F.eli = (uint16_t)L.eli; F.rowbase = (uint16_t)L.rowbase; F.rowd0 = (uint16_t)L.rowd0;
		F.bt_i = (uint16_t)i; F.lowAltQual = (uint8_t)L.lowAltQual; F.elham = (uint8_t)L.elham; F.elcint = (uint8_t)L.elcint; F.bt_j = (uint8_t)j;
		F.flags = (uint8_t)((L.elignore ? FF_ELIGNORE : 0) | (L.f_bdm ? FF_BDM : 0) | (L.f_must ? FF_MUST : 0) | (L.f_invHH ? FF_INVHH : 0) |
		                    (L.f_invExact ? FF_INVEXACT : 0) | (L.K->disableFtab ? FF_DISABLEFTAB : 0));
		L.rowbase = L.rowbase + ((L.d >= L.rowd0) ? (L.d - L.rowd0 + 1) : 0) + P.mask_rows;
		L.stackDepth++; L.K->depth = ndepth; L.unrevOff = btUnrevOff; L.K->oneRevOff = btOneRevOff; L.K->twoRevOff = btTwoRevOff; L.K->threeRevOff = btThreeRevOff;
		L.top = ntop; L.bot = nbot; L.ham = btham; L.K->disableFtab = 0;
		L.pc = PC_FRAME_ENTER;
		break; }
	} while (0);
}

BT_FN void bt_blk_frame_ret(BtLane &L, const BtKParams &P, const BtScratch &S) {
	const BtDevIndex &ix = P.ix[L.ebwtSel]; (void)ix;
	do { {
		if (L.stackDepth == 0) { L.pc = PC_BT_END; break; }
		/* POP: resume the parent after its recursive call returned L.ret */
		const BtFrame &F = S.frames[L.stackDepth - 1];
		L.stackDepth--;
		L.top = F.top; L.bot = F.bot; L.eligibleSz = F.eligibleSz; L.eltop = F.eltop; L.elbot = F.elbot;
		L.K->depth = F.depth; L.d = F.d; L.unrevOff = F.unrevOff; L.K->oneRevOff = F.oneRevOff; L.K->twoRevOff = F.twoRevOff; L.K->threeRevOff = F.threeRevOff;
		L.ham = F.ham; L.altNum = F.altNum; L.eligibleNum = F.eligibleNum; L.eli = F.eli; L.rowbase = F.rowbase; L.rowd0 = F.rowd0;
		L.K->bt_i = F.bt_i; L.K->bt_j = F.bt_j; L.lowAltQual = F.lowAltQual; L.elham = F.elham; L.elcint = F.elcint;
		L.elignore = (F.flags & FF_ELIGNORE) != 0; L.f_bdm = (F.flags & FF_BDM) != 0; L.f_must = (F.flags & FF_MUST) != 0;
		L.f_invHH = (F.flags & FF_INVHH) != 0; L.f_invExact = (F.flags & FF_INVEXACT) != 0; L.K->disableFtab = (F.flags & FF_DISABLEFTAB) != 0;
		L.K->bttop = 0; L.K->btbot = F.btspread;
		L.pc = PC_CHILD_RET;
	}
	} while (0);
}

BT_FN void bt_blk_child_ret(BtLane &L, const BtKParams &P, const BtScratch &S) {
	const BtDevIndex &ix = P.ix[L.ebwtSel]; (void)ix;
	do { {
		/* after the recursive call (ebwt_search_backtrack.h:972-1064) */
		if (L.ret) { L.pc = PC_FRAME_RET; break; }
		if (L.K->bailed || (L.halfAndHalf && L.K->maxBts > 0 && L.K->numBts >= L.K->maxBts)) { L.K->bailed = 1; L.ret = 0; L.pc = PC_FRAME_RET; break; }
		{
			const uint32_t ri = bt_row_idx(L, L.K->bt_i);
			const uint32_t el = S.elims[ri] | (1u << L.K->bt_j);
			S.elims[ri] = (uint8_t)el;
			if (el == 15u) { uint32_t *lm = bt_live_mask(L, P, S) + ((L.K->bt_i - L.rowd0) >> 5); *lm &= ~(1u << ((L.K->bt_i - L.rowd0) & 31u)); }   /* no alternative left at that position */
		}
		L.eligibleSz -= (L.K->btbot - L.K->bttop);
		L.eligibleNum--;
		L.elignore = 1;
		L.altNum--;
		if (L.altNum == 0) { L.ret = 0; L.pc = PC_FRAME_RET; break; }
		BT_PROF(3, 1);
		if (L.eligibleNum == 0 && L.considerQuals) {
			/* re-scan the frame for the next-lowest quality (1004-1058) */
			L.lowAltQual = 0xff;
			BT_PROF(4, 1);
			/* (the reference walks every position of the frame from d down; only those of the live mask can contribute) */
			const uint32_t *lm = bt_live_mask(L, P, S);
			int32_t w = (int32_t)((L.d - L.rowd0) >> 5);
			uint32_t mw = lm[w];
#pragma unroll 1
			for (;;) {
				while (mw == 0 && w > 0) mw = lm[--w];
				if (mw == 0) break;
				const uint32_t b = 31u - bt_clz32(mw);
				mw &= ~(1u << b);
				const uint32_t k = L.rowd0 + ((uint32_t)w << 5) + b;
				BT_PROF(5, 1);
				const uint32_t kq = bt_qual_at(L, L.qlen - k - 1);
				const bool kAlt = (L.ham + bt_mm_penalty(L.maqPenalty, kq) <= L.qualThresh);
				bool kOverrides = false;
				if (kAlt) {
					if (kq < L.lowAltQual) kOverrides = true;
					if (kq <= L.lowAltQual) {
						const uint32_t ri = bt_row_idx(L, k);
						const uint32_t el = S.elims[ri];
						BT_PROF(7, 1); if (el != 15) BT_PROF(8, 1);
						for (uint32_t l = 0; l < 4; l++) {
							if ((el & (1u << l)) == 0) {
								const uint32_t ptop = bt_pair_top(S, ri, l), pbot = bt_pair_bot(S, ri, l);
								if (kOverrides) {
									L.lowAltQual = kq; kOverrides = false; L.eligibleNum = 0; L.eligibleSz = 0;
									L.eli = k; L.eltop = ptop; L.elbot = pbot; L.elham = bt_mm_penalty(L.maqPenalty, kq); L.elcint = l; L.elignore = 0;
								}
								L.eligibleNum++;
								L.eligibleSz += pbot - ptop;
							}
						}
					}
				}
			}
		}
		L.pc = PC_BTLOOP;
		break; }
	} while (0);
}

BT_FN void bt_blk_pos_end(BtLane &L, const BtKParams &P, const BtScratch &S) {
	const BtDevIndex &ix = P.ix[L.ebwtSel]; (void)ix;
	do { {
		/* (ebwt_search_backtrack.h:1066-1078) reached when the backtrack loop gives up on a position */
		if (L.f_must || L.f_invHH || L.f_invExact) { L.ret = 0; L.pc = PC_FRAME_RET; break; }
		if (L.top == L.bot && L.altNum == 0) { L.ret = 0; L.pc = PC_FRAME_RET; break; }
		L.d++;
		L.pc = PC_POS;
		break; }
	} while (0);
}

BT_FN void bt_blk_report(BtLane &L, const BtKParams &P, const BtScratch &S) {
	const BtDevIndex &ix = P.ix[L.ebwtSel]; (void)ix;
	do { {
		/* reportAlignment(rep_sd, rep_top, rep_bot, rep_cost) (ebwt_search_backtrack.h:1455-1513) and the
		 * prologue of reportFullAlignment (1522-1538) */
		uint32_t sd = L.K->rep_sd;
		if (L.reportPartials) {
			if (sd > 0) bt_report_partial(L, P, S, sd);
			L.ret = 0; L.pc = PC_REPORT_RET; break;
		}
		uint32_t stratum = 0;
#pragma unroll 1
		for (uint32_t i = 0; i < sd; i++) if (bt_mm_pos(S, i) >= (L.qlen - L.rev3_0)) stratum++;    /* calcStratum */
		stratum += L.K->nmuts;
		sd += L.K->nmuts;
		if (sd == 0 && !L.K->reportExacts) { L.ret = 0; L.pc = PC_REPORT_RET; break; }
		L.K->rep_sd = sd; L.K->rep_cost = (L.K->rep_cost & 0xffffu) | ((stratum << 14) & 0xffffu); L.K->rep_stratum = stratum;
		L.K->rep_r = L.K->rep_top + (bt_rand_next(L.K->rnd) % (L.K->rep_bot - L.K->rep_top));
		L.K->rep_i = 0;
		L.pc = PC_REPORT_ROW;
	}
	} while (0);
}

BT_FN void bt_blk_report_row(BtLane &L, const BtKParams &P, const BtScratch &S) {
	const BtDevIndex &ix = P.ix[L.ebwtSel]; (void)ix;
	do { {
		/* loop of reportFullAlignment (ebwt_search_backtrack.h:1539-1564) */
		const uint32_t spread = L.K->rep_bot - L.K->rep_top;
		if (L.K->rep_i >= spread) { L.ret = 0; L.pc = PC_REPORT_RET; break; }
		uint32_t ri = L.K->rep_r + L.K->rep_i;
		if (ri >= L.K->rep_bot) ri -= spread;
		L.crow = ri; L.cjumps = 0;
		if (((ri & ix.offMask) != ri) && ri != ix.zOff) { L.pc = PC_CHASE; break; }
		L.pc = PC_RESOLVE;
	}
	} while (0);
}

BT_FN void bt_blk_resolve(BtLane &L, const BtKParams &P, const BtScratch &S) {
	const BtDevIndex &ix = P.ix[L.ebwtSel]; (void)ix;
	do { {
		/* marked row reached (ebwt.h:2735-2755), then Ebwt::report (2635-2682) and the sink:
		 * NGoodHitSinkPerThread::reportHit (hit.h:969-985) / AllHitSinkPerThread::reportHit (hit.h:1201-1209)
		 * fused with the Hit construction of EbwtSearchParams::reportHit (ebwt.h:1288-1405) */
		uint32_t off;
		if (L.crow == ix.zOff) off = L.cjumps;
		else { off = BT_LDG(ix.offs + (L.crow >> ix.offRate)) + L.cjumps; L.K->s_offs++; }
		uint32_t tidx = 0, toff = 0;
		bool stop = false;
		if (bt_joined_to_text(ix, L.qlen, off, tidx, toff)) {
			const BtPolicy &pol = P.pol;
			const uint32_t n = pol.allHits ? 0xffffffffu : pol.khits;
			L.K->found++;
			if (L.K->found > pol.mhits) stop = true;
			else {
				if (L.K->found <= n) {
					if (L.K->found <= P.slots) {
						uint32_t *rec = P.hits + ((size_t)L.K->rid * P.slots + (L.K->found - 1)) * P.rec_words;
						const uint32_t nmm = L.K->rep_sd, nsearch = nmm - L.K->nmuts;      /* frame-stack mismatches, then promoted seedling muts */
						rec[0] = tidx; rec[1] = toff; rec[2] = L.K->rep_bot - L.K->rep_top - 1;
						rec[3] = (L.K->rep_cost & 0xffffu) | (L.K->rep_stratum << 16) | (L.K->fw << 24);
						rec[4] = nmm;
						const bool flip = (ix.fw != L.K->fw);                          /* ebwt.h:1339-1350 */
#pragma unroll 1
						for (uint32_t i = 0; i < nmm; i++) {
							uint32_t pos, refc;
							if (i < nsearch) { pos = S.frames[i].mm_pos; refc = S.frames[i].mm_refc; }
							else { const uint32_t km = i - nsearch; const uint32_t mu = km == 0 ? L.K->mut0 : km == 1 ? L.K->mut1 : L.K->mut2; pos = mu & 0xffffu; refc = (mu >> 16) & 0xff; }
							if (flip) pos = L.qlen - pos - 1;
							if (i < P.mm_cap) rec[BT_HIT_HDR + i] = pos | (refc << 16); else L.flags |= BT_FLAG_MM_OVF;
						}
					} else L.flags |= BT_FLAG_HITS_OVF;
				}
				if (!pol.allHits && L.K->found == n && (pol.mhits == 0xffffffffu || pol.mhits < n)) stop = true;
			}
		}
		if (stop) { L.ret = 1; L.pc = PC_REPORT_RET; }
		else { L.K->rep_i++; L.pc = PC_REPORT_ROW; }
		break; }
	} while (0);
}

BT_FN void bt_blk_report_ret(BtLane &L, const BtKParams &P, const BtScratch &S) {
	const BtDevIndex &ix = P.ix[L.ebwtSel]; (void)ix;
	do { {
		switch (L.K->rep_site) {
		case SITE_MAIN:
			if (!L.ret) { L.top = L.bot; L.pc = PC_BTLOOP; }
			else L.pc = PC_FRAME_RET;
			break;
		case SITE_BT: L.pc = PC_CHILD_RET; break;
		case SITE_END: L.pc = PC_FRAME_RET; break;
		default: L.pc = PC_BT_END; break;
		}
		break; }
	} while (0);
}

BT_FN void bt_blk_bt_end(BtLane &L, const BtKParams &P, const BtScratch &S) {
	const BtDevIndex &ix = P.ix[L.ebwtSel]; (void)ix;
	do { {
		/* tail of backtrack(depth, top, bot, ...) and finalize() (ebwt_search_backtrack.h:348-352, 303-324);
		 * setMuts(NULL) of the seedling loops */
		L.K->numBts = 0; L.K->bailed = 0;
		if (L.K->nmuts > 0) bt_apply_muts(L, true);
		if (L.reportPartials > 0 && L.K->npart > 0) L.ret = 1;
		L.K->done = BTS_IGNORE(L.K->step) ? 0u : L.ret;
		L.pc = PC_PHASE;
		break; }
	} while (0);
}

/* One rare transition (the queue-driven kernel and the chain below): dispatch on the state; FRAME_ENTER runs into POS, FRAME_RET
 * into CHILD_RET, REPORT into REPORT_ROW into RESOLVE when the block left the lane there. */
BT_FN void bt_rare_step(BtLane &L, const BtKParams &P, const BtScratch &S) {
	switch (L.pc) {
	case PC_PHASE: bt_blk_phase(L, P, S); break;
	case PC_BT_BEGIN: bt_blk_bt_begin(L, P, S); break;
	case PC_FRAME_ENTER: bt_blk_frame_enter(L, P, S); if (L.pc != PC_POS) break; /* fallthrough */
	case PC_POS: bt_blk_pos(L, P, S); break;
	case PC_BTLOOP: bt_blk_btloop(L, P, S); break;
	case PC_FRAME_RET: bt_blk_frame_ret(L, P, S); if (L.pc != PC_CHILD_RET) break; /* fallthrough */
	case PC_CHILD_RET: bt_blk_child_ret(L, P, S); break;
	case PC_POS_END: bt_blk_pos_end(L, P, S); break;
	case PC_REPORT: bt_blk_report(L, P, S); if (L.pc != PC_REPORT_ROW) break; /* fallthrough */
	case PC_REPORT_ROW: bt_blk_report_row(L, P, S); if (L.pc != PC_RESOLVE) break; /* fallthrough */
	case PC_RESOLVE: bt_blk_resolve(L, P, S); break;
	case PC_REPORT_RET: bt_blk_report_ret(L, P, S); break;
	case PC_BT_END: bt_blk_bt_end(L, P, S); break;
	default: break;
	}
}

/* Begin a read: GET_READ (ebwt_search.cpp:923-961).  The caller points rseq/rqual at a writable copy. */
BT_FN void bt_begin_read(BtLane &L, const BtKParams &P, uint32_t rid) {
	L.K->rid = rid;
	L.rlen = (uint32_t)(P.roff[rid + 1] - P.roff[rid]);
	L.K->seed = P.seeds[rid];
	L.K->found = 0; L.flags = 0; L.K->ph = 0; L.K->done = 0; L.K->npart = 0; L.K->nmuts = 0; L.K->pal_i = 0; L.K->step = 0; L.nit = 0;
	L.qualThresh = P.pol.mode == 0 ? 0xffffffffu : P.pol.qualThresh;
	L.K->maxBts = P.pol.mode == 0 ? 0xffffffffu : P.pol.maxBts;
	L.maqPenalty = P.pol.mode == 0 ? 1u : (uint32_t)P.pol.maqRound;
	L.pc = L.rlen > 0 ? PC_PHASE : PC_FINISH_READ;
}

/* HitSinkPerThread::finishRead (hit.h:741-786): the host applies -m suppression / -k truncation
 * from `found`; the kernel stored the first min(found, n, slots) hits. */
BT_FN void bt_finish_read(BtLane &L, const BtKParams &P) {
	if (L.flags & BT_FLAG_RETRY) L.K->found = 0;          /* re-run by a later pass */
	P.found[L.K->rid] = L.K->found;
	P.flags[L.K->rid] = L.flags;
}

/* A fast transition with its fetch stage: the rank block(s) of the pending LF / chase step plus the
 * NEXT position's query character and quality — independent loads, all in flight together. */
BT_FN void bt_fast_iter(BtLane &L, const BtKParams &P, const BtScratch &S) {
	const BtDevIndex &ix = P.ix[L.ebwtSel];
	const bool isChase = (L.pc == PC_CHASE);
	BtBlock bA, bB;
	bA.occ.x = bA.occ.y = bA.occ.z = bA.occ.w = 0; bA.hi = bA.lo = 0;
	if (isChase || L.lfk <= LFK_PAIR) { bA = bt_load_block(ix, isChase ? L.crow : L.ltop); L.s_blk++; }
	bB = bA;
	if (!isChase && (L.lfk == LFK_EX || L.lfk == LFK_PAIR) && (L.lbot >> 6) != (L.ltop >> 6)) { bB = bt_load_block(ix, L.lbot); L.s_blk++; }
	uint32_t nc = 4, nq = 0;
	if (!isChase && L.d + 1 < L.qlen) { nc = bt_qry(L, L.qlen - L.d - 2); nq = bt_qual_at(L, L.qlen - L.d - 2); }
	L.s_iter++; L.nit++;
#if defined(BT_SPLIT_LF) || defined(BT_SPLIT_CHASE)
	if (isChase) {
		/* one step of the row walk of Ebwt::reportChaseOne (ebwt.h:2727-2734): mapLF(l) */
		const uint32_t c = bt_row_l(bA, L.crow);
		const uint32_t nr = bt_lf(ix, bA, L.crow, c);
		L.crow = nr; L.cjumps++;
		L.s_lf++; L.s_chase++;
		if (!(((nr & ix.offMask) != nr) && nr != ix.zOff)) L.pc = PC_RESOLVE;
		return;
	}
#endif
	const uint32_t c = L.c;
	uint32_t tops[4] = { 0, 0, 0, 0 }, bots[4] = { 0, 0, 0, 0 };
#ifndef BT_SPLIT_LF
	/* One code path for the three LF kinds (profiles/README.md, "What a lane does" / "Warp-level replay"): they are about equally
	 * frequent, so a warp that branches on the kind usually runs all three paths back to back.  Here: the quartet on both rows,
	 * from which every kind's result follows — mapLF1's is bots[c] - tops[c] == 1 exactly when rowL(top) == c and top is not the
	 * '$' row (the block's A count skips '$').  Counters keep the reference's meaning.  -DBT_SPLIT_LF restores the three paths
	 * (`make experiments`) for the A/B that this change still owes: it was made without a GPU, on the replay's prediction. */
	/* The locate step shares the path too: mapLF(row) = the quartet's entry of the row's own character (ebwt.h:2727-2734), so a warp
	 * whose lanes are split between matching and locating still executes the block fetch and the rank arithmetic once. */
	if (isChase || L.lfk <= LFK_PAIR) {
		const uint32_t rowA = isChase ? L.crow : L.ltop;
		bt_lf_ex(ix, bA, rowA, tops);
		const uint32_t rlA = bt_row_l(bA, rowA);
#ifndef BT_SPLIT_CHASE
		if (isChase) {
			const uint32_t nr = rlA == 0 ? tops[0] : rlA == 1 ? tops[1] : rlA == 2 ? tops[2] : tops[3];
			L.crow = nr; L.cjumps++;
			L.s_lf++; L.s_chase++;
			if (!(((nr & ix.offMask) != nr) && nr != ix.zOff)) L.pc = PC_RESOLVE;
			return;
		}
#endif
		if (L.lfk == LFK_ONE) {                                          /* bot = top + 1: the quartet of the next row differs by the one character at top */
			const uint32_t rl = (L.top != ix.zOff) ? rlA : 4u;
			bots[0] = tops[0] + (rl == 0); bots[1] = tops[1] + (rl == 1); bots[2] = tops[2] + (rl == 2); bots[3] = tops[3] + (rl == 3);
		} else bt_lf_ex(ix, bB, L.lbot, bots);
		if (L.lfk == LFK_EX) { L.s_lfex++; if (c < 4) { L.top = tops[c]; L.bot = bots[c]; } }
		else if (L.lfk == LFK_ONE) {
			const bool hit = c < 4 && bots[c] != tops[c];
			L.top = hit ? tops[c] : BT_OFF_MASK; L.bot = hit ? tops[c] + 1 : BT_OFF_MASK;
			L.s_lf++;
		} else { L.top = tops[c & 3]; L.bot = bots[c & 3]; L.s_lf += 2; }   /* (bt_position reads tops/bots only for LFK_EX / LFK_FCHR) */
	} else
#else
	if (L.lfk == LFK_EX) {
		/* mapLFEx(ltop, lbot, tops, bots) (ebwt.h:2334-2380) */
		bt_lf_ex(ix, bA, L.ltop, tops);
		bt_lf_ex(ix, bB, L.lbot, bots);
		L.s_lfex++;
		if (c < 4) { L.top = tops[c]; L.bot = bots[c]; }
	} else if (L.lfk == LFK_ONE) {
		/* mapLF1(top, ltop, c) (ebwt.h:2494-2524) */
		uint32_t t;
		if (bt_row_l(bA, L.ltop) != c || L.top == ix.zOff) t = BT_OFF_MASK;
		else t = bt_lf(ix, bA, L.ltop, c);
		L.top = t; L.bot = t;
		if (t != BT_OFF_MASK) L.bot++;
		L.s_lf++;
	} else if (L.lfk == LFK_PAIR) {
		const uint32_t t = bt_lf(ix, bA, L.ltop, c), b = bt_lf(ix, bB, L.lbot, c);
		L.top = t; L.bot = b;
		L.s_lf += 2;
	} else
#endif
	if (L.lfk == LFK_FCHR) {
		/* first quartet from fchr[] (ebwt_search_backtrack.h:531-543) */
		tops[0] = ix.fchr[0]; tops[1] = ix.fchr[1]; tops[2] = ix.fchr[2]; tops[3] = ix.fchr[3];
		bots[0] = ix.fchr[1]; bots[1] = ix.fchr[2]; bots[2] = ix.fchr[3]; bots[3] = ix.fchr[4];
		if (c < 4) { L.top = tops[c]; L.bot = bots[c]; }
	}
	bt_position(L, P, S, tops, bots, nc, nq);
}

/* A batch of rare transitions: keep going until the lane needs a rank block again (or the read ends).
 *
 * BT_RARE_SWEEP (default): the blocks in the order a search flows through them — report -> (pop frame / bookkeeping after a child)* ->
 * end of a backtracker -> phase program -> new backtracker -> backtrack target + push -> frame entry -> position prologue — so that a
 * lane passes through its whole chain in ONE pass over the code and every block is executed at most once per round for all the lanes
 * of the warp that need it (reconverging in between), instead of one `switch` per chain link in which the warp executes every distinct
 * state of its lanes serially, link after link.  -DBT_RARE_SWEEP=0 restores the chain of single transitions. */
#ifndef BT_RARE_CHAIN
#define BT_RARE_CHAIN 8
#endif
#ifndef BT_RARE_SWEEP
#define BT_RARE_SWEEP 1
#endif
#ifndef BT_SWEEP_ROUNDS
#define BT_SWEEP_ROUNDS 3
#endif
BT_FN void bt_rare_iter(BtLane &L, const BtKParams &P, const BtScratch &S, const uint32_t budget) {
#if BT_RARE_SWEEP
#define BT_STEP(state, blk) if (L.pc == (state) && !(L.flags & BT_FLAG_SCRATCH_OVF)) { L.s_iter++; L.nit++; blk(L, P, S); }
#pragma unroll 1
	for (int r = 0; r < BT_SWEEP_ROUNDS && BT_IS_RARE_STEP(L.pc); r++) {
		if (L.flags & BT_FLAG_SCRATCH_OVF) { L.pc = PC_FINISH_READ; break; }
		if ((budget && L.nit > budget) || (P.slot_ctx && !P.resume && L.K->npart >= P.PCAP)) {                  /* heavy read, or its seedling list is full */
			if (P.slot_ctx) L.flags |= BT_FLAG_PREEMPT;                                                    /* suspended as it is: the state stays put */
			else { L.flags |= BT_FLAG_BUDGET; L.pc = PC_FINISH_READ; }
			break;
		}
		BT_STEP(PC_REPORT, bt_blk_report)
		BT_STEP(PC_REPORT_ROW, bt_blk_report_row)
		BT_STEP(PC_RESOLVE, bt_blk_resolve)
		BT_STEP(PC_REPORT_RET, bt_blk_report_ret)
#pragma unroll 1
		for (int g = 0; g < 4 && (L.pc == PC_FRAME_RET || L.pc == PC_CHILD_RET); g++) {
			BT_STEP(PC_FRAME_RET, bt_blk_frame_ret)
			BT_STEP(PC_CHILD_RET, bt_blk_child_ret)
		}
		BT_STEP(PC_BT_END, bt_blk_bt_end)
		BT_STEP(PC_PHASE, bt_blk_phase)
		BT_STEP(PC_BT_BEGIN, bt_blk_bt_begin)
		BT_STEP(PC_BTLOOP, bt_blk_btloop)
		BT_STEP(PC_POS_END, bt_blk_pos_end)
		BT_STEP(PC_FRAME_ENTER, bt_blk_frame_enter)
		BT_STEP(PC_POS, bt_blk_pos)
	}
#undef BT_STEP
	/* a report in the last round may have filled the seedling list and the lane left for a fast state: the next position could report again */
	if (P.slot_ctx && !P.resume && L.K->npart >= P.PCAP && L.pc != PC_FINISH_READ) L.flags |= BT_FLAG_PREEMPT;
#else
#pragma unroll 1
	for (int k = 0; k < BT_RARE_CHAIN && BT_IS_RARE_STEP(L.pc); k++) {
		if (L.flags & BT_FLAG_SCRATCH_OVF) { L.pc = PC_FINISH_READ; break; }
		if ((budget && L.nit > budget) || (P.slot_ctx && !P.resume && L.K->npart >= P.PCAP)) {
			if (P.slot_ctx) L.flags |= BT_FLAG_PREEMPT;
			else { L.flags |= BT_FLAG_BUDGET; L.pc = PC_FINISH_READ; }
			break;
		}
		L.s_iter++; L.nit++;
		bt_rare_step(L, P, S);
	}
	if (P.slot_ctx && !P.resume && L.K->npart >= P.PCAP && L.pc != PC_FINISH_READ) L.flags |= BT_FLAG_PREEMPT;
#endif
}
BT_FN void bt_rare_iter(BtLane &L, const BtKParams &P, const BtScratch &S) { bt_rare_iter(L, P, S, P.budget); }
