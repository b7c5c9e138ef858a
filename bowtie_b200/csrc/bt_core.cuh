/*
 * bt_core.cuh — the per-read search state machine of the B200 kernels.
 *
 * One GPU thread owns one read at a time ("lane").  The depth-first backtracking search of
 * Bowtie 1 (GreedyDFSRangeSource, reference ebwt_search_backtrack.h:23-1779, driven by the
 * search_*.c phase fragments) is re-expressed as an explicit state machine with a frame stack
 * in global scratch, so that the 32 lanes of a warp stay in one instruction stream:
 *
 *     loop:  [fetch]   every lane that needs an LF step loads its rank block(s)   (converged)
 *            [switch]  lanes advance their own state by one transition            (by state)
 *
 * Recursion in the reference becomes PUSH/POP of BtFrame records; "return into the caller" is a
 * continuation code.  All pseudo-random draws happen in exactly the reference's order, so the
 * output is bit-identical (see tests/).
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

#if defined(__CUDACC__)
#define BT_FN __device__ __forceinline__
#define BT_LDG(p) __ldg(p)
#define BT_POPC64(x) __popcll(x)
#else
#define BT_FN static inline
#define BT_LDG(p) (*(p))
#define BT_POPC64(x) __builtin_popcountll(x)
#ifndef BT_HOST_EMU
#error "bt_core.cuh is device code; the host build exists only for tests/host_emu (define BT_HOST_EMU)"
#endif
struct uint4 { uint32_t x, y, z, w; };
#endif

#define BT_OFF_MASK 0xffffffffu

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
};

/* flags written per read */
#define BT_FLAG_STACK_OVF 1u   /* frame-row scratch exhausted          */
#define BT_FLAG_FRAME_OVF 2u   /* more nested frames than FCAP         */
#define BT_FLAG_PART_OVF  4u   /* more seedlings than PCAP             */
#define BT_FLAG_HITS_OVF  8u   /* more reportable hits than slots      */
#define BT_FLAG_MM_OVF   16u   /* more mismatches than the record holds */
#define BT_FLAG_ANY_OVF  31u

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
	uint32_t R, FCAP, PCAP;
	unsigned long long *work;     /* work-queue cursor                                        */
	unsigned long long *stats;    /* [8]: lfex, lf, chase, ftab, offs, backtracks, iters, blockloads */
};

/* ---- program counters ------------------------------------------------------------------------ */
enum {
	PC_NEXT_READ = 0, PC_PHASE, PC_BT_BEGIN, PC_FRAME_ENTER, PC_POS, PC_LF, PC_POS2, PC_BTLOOP,
	PC_CHILD_RET, PC_POS_END, PC_FRAME_RET, PC_REPORT, PC_REPORT_ROW, PC_CHASE, PC_RESOLVE,
	PC_REPORT_RET, PC_BT_END, PC_FINISH_READ, PC_EXIT
};
enum { SITE_MAIN = 0, SITE_BT, SITE_END, SITE_FTABFULL };
enum { LFK_EX = 0, LFK_ONE, LFK_PAIR, LFK_FCHR, LFK_NONE };

struct BtLane {
	/* read */
	uint32_t rid, rlen, seed, found, flags, hasN;
	uint64_t roff;
	const uint8_t *rseq, *rqual;     /* this read's bases / qualities (shared-memory staging copy, or global) */
	/* control */
	uint32_t pc, ph, done, ret, lfk;
	/* backtracker object state */
	uint32_t ebwtSel, fw, considerQuals, halfAndHalf, reportPartials, reportExacts, maqPenalty;
	uint32_t qualThresh, maxBts;
	uint32_t qlen, depth5, depth3, unrev0, rev1_0, rev2_0, rev3_0, iham;
	uint32_t nmuts, mut0, mut1, mut2;   /* pos | newBase << 16 | oldBase << 24                 */
	uint32_t rnd, numBts, bailed, viewRev, viewComp;
	/* current frame */
	uint32_t stackDepth, depth, d, unrevOff, oneRevOff, twoRevOff, threeRevOff, ham;
	uint32_t altNum, eligibleNum, eligibleSz, eli, elignore, eltop, elbot, elham, elcint, lowAltQual;
	uint32_t rowbase, rowd0, disableFtab;
	uint32_t top, bot, ltop, lbot;   /* ltop/lbot: rows of the SideLocus pair (ebwt_search_backtrack.h:419-426,569-574) */
	uint32_t c, q, curIsAlt, curIsElig, curOverrides;
	uint32_t f_bdm, f_must, f_invHH, f_invExact;
	uint32_t bt_i, bt_j, bttop, btbot, btham;
	/* report / chase */
	uint32_t rep_site, rep_sd, rep_cost, rep_stratum, rep_top, rep_bot, rep_r, rep_i;
	uint32_t crow, cjumps;
	/* seedlings */
	uint32_t npart, pal_i;
	/* statistics */
	uint32_t s_lfex, s_lf, s_chase, s_ftab, s_offs, s_bt, s_iter, s_blk;
};

/* ---- small helpers --------------------------------------------------------------------------- */
BT_FN uint32_t bt_qual_round(uint32_t q) { return q < 5 ? 0u : q < 15 ? 10u : q < 25 ? 20u : 30u; }  /* qual.cpp:4-32 */
BT_FN uint32_t bt_mm_penalty(uint32_t maq, uint32_t q) { return maq ? bt_qual_round(q) : q; }          /* qual.h:55-61 */
BT_FN uint32_t bt_rand_next(uint32_t &last) {                                                          /* random_source.h:45-54 */
	last = 1664525u * last + 1013904223u;
	uint32_t ret = last >> 16;
	last = 1664525u * last + 1013904223u;
	return ret ^ last;
}

/* query character / quality at offset `cur` of the current view (_qry / _qual of
 * GreedyDFSRangeSource::setQuery, ebwt_search_backtrack.h:90-99), with seedling mutations applied */
BT_FN uint32_t bt_qry_raw(const BtKParams &P, const BtLane &L, uint32_t cur) {
	uint32_t idx = L.viewRev ? (L.rlen - 1 - cur) : cur;
	uint32_t c = L.rseq[idx];
	if (L.viewComp && c < 4) c ^= 3;
	return c;
}
BT_FN uint32_t bt_qry(const BtKParams &P, const BtLane &L, uint32_t cur) {
	uint32_t c = bt_qry_raw(P, L, cur);
	if (L.nmuts > 0) {
		if ((L.mut0 & 0xffffu) == cur) c = (L.mut0 >> 16) & 0xff;
		if (L.nmuts > 1 && (L.mut1 & 0xffffu) == cur) c = (L.mut1 >> 16) & 0xff;
		if (L.nmuts > 2 && (L.mut2 & 0xffffu) == cur) c = (L.mut2 >> 16) & 0xff;
	}
	return c;
}
BT_FN uint32_t bt_qual_at(const BtKParams &P, const BtLane &L, uint32_t cur) {
	uint32_t idx = L.viewRev ? (L.rlen - 1 - cur) : cur;
	uint32_t ch = L.rqual[idx];
	return ch >= 33 ? ch - 33 : 0;                     /* phredCharToPhredQual qual.h:15-17 */
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

/* Sets up a backtracker invocation: the union of the constructor arguments, setQuery, setOffs,
 * setQlen, setReportExacts of one GreedyDFSRangeSource object. */
BT_FN void bt_cfg(BtLane &L, const BtKParams &P, uint32_t ebwtSel, uint32_t fw, uint32_t considerQuals, uint32_t halfAndHalf,
                  uint32_t reportPartials, uint32_t reportExacts, uint32_t qlen,
                  uint32_t depth5, uint32_t depth3, uint32_t unrev, uint32_t r1, uint32_t r2, uint32_t r3) {
	L.ebwtSel = ebwtSel; L.fw = fw; L.considerQuals = considerQuals; L.halfAndHalf = halfAndHalf;
	L.reportPartials = reportPartials; L.reportExacts = reportExacts;
	L.qlen = qlen; L.depth5 = depth5; L.depth3 = depth3; L.unrev0 = unrev; L.rev1_0 = r1; L.rev2_0 = r2; L.rev3_0 = r3;
	L.iham = 0; L.nmuts = 0;
	L.rnd = L.seed;                                   /* setQuery: _rand.init(r.seed) */
	uint32_t ebwtFw = (ebwtSel == 0);
	L.viewRev = ebwtFw ? !fw : fw;
	L.viewComp = !fw;
	L.pc = PC_BT_BEGIN;
}

/* PartialAlignmentManager::toMutsString (ebwt_search_util.h:299-357) for seedling `pal` */
BT_FN void bt_set_muts(BtLane &L, const BtKParams &P, uint64_t pal) {
	uint32_t oldQuals = 0; L.nmuts = 0;
	for (uint32_t k = 0; k < 3; k++) {
		uint32_t pos = (uint32_t)(pal >> (16 * k)) & 0xffffu;
		if (pos == 0xffffu) break;
		uint32_t chr = (uint32_t)(pal >> (48 + 2 * k)) & 3u;
		uint32_t tpos = (L.rlen - 1 - pos) & 0xffffu;
		oldQuals = (oldQuals + bt_mm_penalty(L.maqPenalty, bt_qual_at(P, L, tpos))) & 0xff;
		uint32_t oldc = bt_qry_raw(P, L, tpos);
		uint32_t mv = tpos | (chr << 16) | (oldc << 24);
		if (k == 0) L.mut0 = mv; else if (k == 1) L.mut1 = mv; else L.mut2 = mv;
		L.nmuts = k + 1;
	}
	L.iham = oldQuals;
	L.pc = PC_BT_BEGIN;
}

/* ---- phase programs (search_exact.c, search_1mm_phase*.c, search_23mm_phase*.c,
 *      search_seeded_phase*.c).  Sets L.pc = PC_BT_BEGIN to launch a backtracker, or
 *      PC_FINISH_READ. ---------------------------------------------------------------------------- */
BT_FN void bt_phase(BtLane &L, const BtKParams &P, const BtScratch &S) {
	const BtPolicy &pol = P.pol;
	const uint32_t len = L.rlen;
	const uint32_t nofw = pol.nofw, norc = pol.norc;
	if (pol.mode == 0 && pol.mms == 0) {
		/* search_exact.c:7-27 */
		for (;;) switch (L.ph) {
		case 0: L.ph = 1; if (!nofw) { bt_cfg(L, P, 0, 1, 0, 0, 0, 1, len, 0, 0, len, len, len, len); return; } break;
		case 1: if (L.done) { L.pc = PC_FINISH_READ; return; }
		        L.ph = 2; if (!norc) { bt_cfg(L, P, 0, 0, 0, 0, 0, 1, len, 0, 0, len, len, len, len); return; } break;
		default: L.pc = PC_FINISH_READ; return;
		}
	} else if (pol.mode == 0 && pol.mms == 1) {
		/* search_1mm_phase1.c, search_1mm_phase2.c */
		const uint32_t s = len, s3 = s >> 1, s5 = (s >> 1) + (s & 1);
		for (;;) {
			if (L.done) { L.pc = PC_FINISH_READ; return; }
			switch (L.ph++) {
			case 0: if (!nofw) { bt_cfg(L, P, 0, 1, 0, 0, 0, 1, len, 0, 0, s, s, s, s); return; } break;
			case 1: if (!norc) { bt_cfg(L, P, 0, 0, 0, 0, 0, 1, len, 0, 0, s, s, s, s); return; } break;
			case 2: if (!norc) { bt_cfg(L, P, 0, 0, 0, 0, 0, 0, len, 0, 0, s5, s, s, s); return; } break;
			case 3: if (!nofw) { bt_cfg(L, P, 0, 1, 0, 0, 0, 0, len, 0, 0, s5, s, s, s); return; } break;
			case 4: if (!norc) { bt_cfg(L, P, 1, 0, 0, 0, 0, 0, len, 0, 0, s3, s, s, s); return; } break;
			case 5: if (!nofw) { bt_cfg(L, P, 1, 1, 0, 0, 0, 0, len, 0, 0, s3, s, s, s); return; } break;
			default: L.pc = PC_FINISH_READ; return;
			}
		}
	} else if (pol.mode == 0) {
		/* search_23mm_phase1.c, _phase2.c, _phase3.c with two = true */
		const uint32_t s = len, s3 = s >> 1, s5 = (s >> 1) + (s & 1);
		for (;;) {
			if (L.done) { L.pc = PC_FINISH_READ; return; }
			switch (L.ph++) {
			case 0: if (!nofw) { bt_cfg(L, P, 0, 1, 0, 0, 0, 1, len, 0, 0, len, len, len, len); return; } break;
			case 1: if (!norc) { bt_cfg(L, P, 0, 0, 0, 0, 0, 1, len, 0, 0, s5, s5, s, s); return; } break;
			case 2: if (!nofw) { bt_cfg(L, P, 1, 1, 0, 0, 0, 0, len, 0, 0, s5, s5, s, s); return; } break;
			case 3: if (!norc) { bt_cfg(L, P, 1, 0, 0, 0, 0, 0, len, 0, 0, s3, s3, s, s); return; } break;
			case 4: if (!nofw) { bt_cfg(L, P, 0, 1, 0, 0, 0, 0, len, 0, 0, s3, s3, s, s); return; } break;
			case 5: if (!nofw) { bt_cfg(L, P, 0, 1, 0, 1, 0, 1, len, s3, s, 0, s3, s, s); return; } break;
			case 6: if (!norc) { bt_cfg(L, P, 0, 0, 0, 1, 0, 1, len, s5, s, 0, s5, s, s); return; } break;
			default: L.pc = PC_FINISH_READ; return;
			}
		}
	} else {
		/* search_seeded_phase1.c .. phase4.c */
		const uint32_t m = (uint32_t)pol.mms;
		const uint32_t s = (uint32_t)pol.seedLen, s3 = s >> 1, s5 = (s >> 1) + (s & 1);
		const uint32_t qs = len < s ? len : s, qs3 = qs >> 1, qs5 = (qs >> 1) + (qs & 1);
		const uint32_t SS = (qs < s) ? qs : s, S3 = (qs < s) ? qs3 : s3, S5 = (qs < s) ? qs5 : s5;
		for (;;) {
			switch (L.ph) {
			case 0: {
				L.ph = 1;
				bool skip = false;
				if (len < 4) skip = true;
				else {
					uint32_t ns = 0;
					if (L.hasN) for (uint32_t i = 0; i < qs; i++) if (L.rseq[i] == 4) { if (++ns > m) { skip = true; break; } }
				}
				if (skip) { L.pc = PC_FINISH_READ; return; }
				if (!nofw) { bt_cfg(L, P, 0, 1, 0, 0, 0, 1, len, 0, len, len, len, len, len); return; }
				break; }
			case 1:
				if (L.done) { L.pc = PC_FINISH_READ; return; }
				L.ph = 2;
				if (!norc) { bt_cfg(L, P, 0, 0, 1, 0, 0, 1, len, 0, 0, m > 0 ? S5 : SS, m > 1 ? S5 : SS, m > 2 ? S5 : SS, m > 3 ? S5 : SS); return; }
				break;
			case 2:
				if (L.done) { L.pc = PC_FINISH_READ; return; }
				L.ph = 3;
				if (!nofw) { bt_cfg(L, P, 1, 1, 1, 0, 0, 0, len, 0, 0, m > 0 ? S5 : SS, m > 1 ? S5 : SS, m > 2 ? S5 : SS, m > 3 ? S5 : SS); return; }
				break;
			case 3:
				if (L.done) { L.pc = PC_FINISH_READ; return; }
				if (m == 0) { L.pc = PC_FINISH_READ; return; }
				L.ph = 4;
				L.npart = 0;
				if (!norc) {
					bt_cfg(L, P, 1, 0, 1, 0, m, nofw ? 1u : 0u, len < s ? len : s, 0, 0, S3, m > 1 ? S3 : SS, m > 2 ? S3 : SS, m > 3 ? S3 : SS);
					return;
				}
				break;
			case 4:
				/* phase 3: extend the 4R seedlings on the forward index */
				L.done = 0; L.ph = 5; L.pal_i = 0;
				if (norc) { L.ph = 7; break; }
				if (L.npart > 0) {
					bt_cfg(L, P, 0, 0, 1, 0, 0, 1, len, 0, 0, SS, SS, SS, SS);   /* btr3.setQuery + setOffs */
					bt_set_muts(L, P, S.partials[0]);
					L.pal_i = 1;
					return;
				}
				break;
			case 5:
				if (L.done) { L.pc = PC_FINISH_READ; return; }
				if (L.pal_i < L.npart) {
					/* next seedling: same object, RNG state carries over (no setQuery in the loop) */
					bt_set_muts(L, P, S.partials[L.pal_i]);
					L.pal_i++;
					return;
				}
				L.npart = 0; L.nmuts = 0; L.ph = 6;
				if (m >= 2) { bt_cfg(L, P, 0, 0, 1, 1, 0, 1, len, S5, SS, 0, m <= 2 ? S5 : 0, m < 3 ? SS : S5, SS); return; }
				break;
			case 6:
				if (L.done) { L.pc = PC_FINISH_READ; return; }
				L.ph = 7;
				break;
			case 7:
				if (nofw) { L.pc = PC_FINISH_READ; return; }
				L.ph = 8; L.npart = 0;
				bt_cfg(L, P, 0, 1, 1, 0, m, 1, len < s ? len : s, 0, 0, S3, m > 1 ? S3 : SS, m > 2 ? S3 : SS, m > 3 ? S3 : SS);
				return;
			case 8:
				/* phase 4: extend the 4F seedlings on the mirror index */
				L.done = 0; L.ph = 9; L.pal_i = 0;
				if (L.npart > 0) {
					bt_cfg(L, P, 1, 1, 1, 0, 0, 1, len, 0, 0, SS, SS, SS, SS);
					bt_set_muts(L, P, S.partials[0]);
					L.pal_i = 1;
					return;
				}
				break;
			case 9:
				if (L.done) { L.pc = PC_FINISH_READ; return; }
				if (L.pal_i < L.npart) { bt_set_muts(L, P, S.partials[L.pal_i]); L.pal_i++; return; }
				L.npart = 0; L.nmuts = 0; L.ph = 10;
				if (m >= 2) { bt_cfg(L, P, 1, 1, 1, 1, 0, 1, len, S5, SS, 0, m <= 2 ? S5 : 0, m < 3 ? SS : S5, SS); return; }
				break;
			default: L.pc = PC_FINISH_READ; return;
			}
		}
	}
}

/* hhCheckTop (ebwt_search_backtrack.h:1200-1275) */
BT_FN bool bt_hh_check_top(const BtLane &L, const BtScratch &S) {
	if (L.d == L.depth5) {
		if (L.stackDepth == 0) return false;
	} else if (L.d == L.depth3) {
		if (L.rev3_0 == L.rev2_0) { if (L.stackDepth < 2) return false; }
		else {
			uint32_t lo = 0;
			for (uint32_t i = 0; i < L.stackDepth; i++) {
				uint32_t dd = L.qlen - bt_mm_pos(S, i) - 1;
				if (dd >= L.depth5 && dd < L.depth3) lo++;
			}
			if (lo == 0) return false;
		}
	}
	return true;
}

/* reportPartial (ebwt_search_backtrack.h:1571-1655): seedling = up to 3 (pos, char) pairs */
BT_FN void bt_report_partial(BtLane &L, const BtKParams &P, const BtScratch &S, uint32_t sd) {
	uint64_t al = 0xffffffffffffull;                 /* pos0..2 = 0xffff, chars 0 */
	for (uint32_t k = 0; k < sd && k < 3; k++) {
		al &= ~(0xffffull << (16 * k));
		al |= (uint64_t)(S.frames[k].mm_pos & 0xffffu) << (16 * k);
		al |= (uint64_t)(S.frames[k].mm_refc & 3u) << (48 + 2 * k);
	}
	if (L.npart < P.PCAP) S.partials[L.npart] = al; else L.flags |= BT_FLAG_PART_OVF;
	L.npart++;
}

/* NGoodHitSinkPerThread::reportHit (hit.h:969-985) / AllHitSinkPerThread::reportHit (hit.h:1201-1209)
 * fused with the Hit construction of EbwtSearchParams::reportHit (ebwt.h:1288-1405). */
BT_FN bool bt_sink_report(BtLane &L, const BtKParams &P, const BtScratch &S, uint32_t tidx, uint32_t toff) {
	const BtPolicy &pol = P.pol;
	L.found++;
	if (L.found > pol.mhits) return true;
	uint32_t n = pol.allHits ? 0xffffffffu : pol.khits;
	if (L.found <= n) {
		if (L.found <= P.slots) {
			uint32_t *rec = P.hits + ((size_t)L.rid * P.slots + (L.found - 1)) * P.rec_words;
			const BtDevIndex &ix = P.ix[L.ebwtSel];
			uint32_t nmm = L.rep_sd;
			rec[0] = tidx; rec[1] = toff; rec[2] = L.rep_bot - L.rep_top - 1;
			rec[3] = (L.rep_cost & 0xffffu) | (L.rep_stratum << 16) | (L.fw << 24);
			rec[4] = nmm;
			uint32_t nsearch = nmm - L.nmuts;              /* mismatches from the frame stack, then promoted seedling muts */
			bool flip = (ix.fw != L.fw);                   /* ebwt.h:1339-1350 */
			for (uint32_t i = 0; i < nmm; i++) {
				uint32_t pos, refc;
				if (i < nsearch) { pos = S.frames[i].mm_pos; refc = S.frames[i].mm_refc; }
				else { uint32_t km = i - nsearch; uint32_t mu = km == 0 ? L.mut0 : km == 1 ? L.mut1 : L.mut2; pos = mu & 0xffffu; refc = (mu >> 16) & 0xff; }
				if (flip) pos = L.qlen - pos - 1;
				if (i < P.mm_cap) rec[BT_HIT_HDR + i] = pos | (refc << 16); else L.flags |= BT_FLAG_MM_OVF;
			}
		} else L.flags |= BT_FLAG_HITS_OVF;
	}
	if (!pol.allHits && L.found == n && (pol.mhits == 0xffffffffu || pol.mhits < n)) return true;
	return false;
}

/* Enter a frame: backtrack(stackDepth, depth, unrevOff, ..., top, bot, ham, iham, pairs, elims, disableFtab)
 * up to the while loop (ebwt_search_backtrack.h:363-455). */
BT_FN void bt_frame_enter(BtLane &L, const BtKParams &P, uint32_t stackDepth, uint32_t depth, uint32_t unrevOff, uint32_t oneRevOff,
                          uint32_t twoRevOff, uint32_t threeRevOff, uint32_t top, uint32_t bot, uint32_t ham, uint32_t rowbase, uint32_t disableFtab) {
	L.stackDepth = stackDepth; L.depth = depth; L.unrevOff = unrevOff; L.oneRevOff = oneRevOff; L.twoRevOff = twoRevOff;
	L.threeRevOff = threeRevOff; L.top = top; L.bot = bot; L.ham = ham; L.rowbase = rowbase; L.disableFtab = disableFtab;
	L.rowd0 = depth > unrevOff ? depth : unrevOff;
	if (top != 0 || bot != 0) { L.ltop = top; L.lbot = bot; }
	if (stackDepth > 0) L.s_bt++;
	if (L.rowd0 < L.qlen && L.rowbase + (L.qlen - L.rowd0) > P.R) { L.flags |= BT_FLAG_STACK_OVF; L.found = 0; L.pc = PC_FINISH_READ; return; }
	if (L.halfAndHalf) {
		if (L.maxBts > 0 && L.numBts == L.maxBts) { L.bailed = 1; L.ret = 0; L.pc = PC_FRAME_RET; return; }
		L.numBts++;
	}
	L.altNum = 0; L.eligibleNum = 0; L.eligibleSz = 0; L.eli = 0; L.elignore = 1; L.eltop = 0; L.elbot = 0;
	L.elham = ham; L.elcint = 0; L.lowAltQual = 0xff; L.d = depth;
	L.pc = PC_POS;
}

/* reportAlignment entry (ebwt_search_backtrack.h:1455-1513) + reportFullAlignment prologue (1522-1538) */
BT_FN void bt_report_begin(BtLane &L, const BtKParams &P, const BtScratch &S, uint32_t sd, uint32_t top, uint32_t bot, uint32_t cost, uint32_t site) {
	L.rep_site = site;
	if (L.reportPartials) {
		if (sd > 0) bt_report_partial(L, P, S, sd);
		L.ret = 0; L.pc = PC_REPORT_RET; return;
	}
	uint32_t stratum = 0;
	for (uint32_t i = 0; i < sd; i++) if (bt_mm_pos(S, i) >= (L.qlen - L.rev3_0)) stratum++;    /* calcStratum */
	stratum += L.nmuts;
	cost = (cost & 0xffffu) | ((stratum << 14) & 0xffffu);
	sd += L.nmuts;
	if (sd == 0 && !L.reportExacts) { L.ret = 0; L.pc = PC_REPORT_RET; return; }
	L.rep_sd = sd; L.rep_cost = cost & 0xffffu; L.rep_stratum = stratum; L.rep_top = top; L.rep_bot = bot;
	uint32_t spread = bot - top;
	L.rep_r = top + (bt_rand_next(L.rnd) % spread);
	L.rep_i = 0;
	L.pc = PC_REPORT_ROW;
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
	if (L.top == 0 && L.bot == 0) { L.lfk = LFK_FCHR; L.pc = PC_POS2; }
	else if (L.curIsAlt) { L.lfk = LFK_EX; L.pc = PC_LF; }
	else if (c < 4) { L.lfk = (L.top + 1 == L.bot) ? LFK_ONE : LFK_PAIR; L.pc = PC_LF; }
	else { L.lfk = LFK_NONE; L.pc = PC_POS2; }
}

/* One transition of the lane's state machine.  bA/bB are the rank blocks of (top, bot) resp. the
 * chase row, valid when the lane was in PC_LF / PC_CHASE at the fetch stage. */
BT_FN void bt_step(BtLane &L, const BtKParams &P, const BtScratch &S, const BtBlock &bA, const BtBlock &bB, uint32_t nc, uint32_t nq) {
	const BtDevIndex &ix = P.ix[L.ebwtSel];
	switch (L.pc) {
	case PC_PHASE:
		bt_phase(L, P, S);
		break;

	case PC_BT_BEGIN: {
		/* backtrack(ham) (ebwt_search_backtrack.h:237-297) */
		const uint32_t ftabChars = (uint32_t)ix.ftabChars;
		L.numBts = 0; L.bailed = 0;
		/* tallyNs (1308-1341) */
		uint32_t nsInSeed = 0, nsInFtab = 0; bool ok = true;
		for (uint32_t i = 0; L.hasN && i < L.rev3_0 && ok; i++) {
			if (bt_qry(P, L, L.qlen - i - 1) == 4) {
				nsInSeed++;
				if (nsInSeed == 1) { if (i < L.unrev0) ok = false; }
				else if (nsInSeed == 2) { if (i < L.rev1_0) ok = false; }
				else if (nsInSeed == 3) { if (i < L.rev2_0) ok = false; }
				else ok = false;
			}
		}
		if (!ok) { L.done = 0; L.pc = PC_PHASE; break; }
		for (uint32_t i = 0; L.hasN && i < ftabChars && i < L.qlen; i++) if (bt_qry(P, L, L.qlen - i - 1) == 4) nsInFtab++;
		uint32_t mlim = L.unrev0 < L.qlen ? L.unrev0 : L.qlen;
		if (nsInFtab == 0 && mlim >= ftabChars) {
			uint32_t ftabOff = bt_qry(P, L, L.qlen - ftabChars);                 /* calcFtabOff (1348-1362) */
			for (uint32_t i = ftabChars - 1; i > 0; i--) ftabOff = (ftabOff << 2) | bt_qry(P, L, L.qlen - i);
			uint32_t top = bt_ftab_hi(ix, ftabOff), bot = bt_ftab_lo(ix, ftabOff + 1);
			L.s_ftab++;
			if (L.qlen == ftabChars && bot > top) {
				if (L.reportPartials > 0) bt_frame_enter(L, P, 0, 0, L.unrev0, L.rev1_0, L.rev2_0, L.rev3_0, 0, 0, L.iham, 0, 0);
				else bt_report_begin(L, P, S, 0, top, bot, L.iham, SITE_FTABFULL);
			} else if (bot > top) {
				bt_frame_enter(L, P, 0, ftabChars, L.unrev0, L.rev1_0, L.rev2_0, L.rev3_0, top, bot, L.iham, 0, 0);
			} else { L.ret = 0; L.pc = PC_BT_END; }
		} else {
			bt_frame_enter(L, P, 0, 0, L.unrev0, L.rev1_0, L.rev2_0, L.rev3_0, 0, 0, L.iham, 0, nsInFtab > 0);
		}
		break; }

	case PC_POS: {
		/* top of while(cur < _qlen) (ebwt_search_backtrack.h:456-568), entered with its own query loads
		 * (frame entry and rare paths; the common path chains positions inside PC_LF below) */
		if (L.d >= L.qlen) {
			if (L.stackDepth >= L.reportPartials) bt_report_begin(L, P, S, L.stackDepth, L.top, L.bot, L.ham, SITE_END);
			else { L.ret = 0; L.pc = PC_FRAME_RET; }
			break;
		}
		if (L.halfAndHalf && !bt_hh_check_top(L, S)) { L.ret = 0; L.pc = PC_FRAME_RET; break; }
		const uint32_t cur = L.qlen - L.d - 1;
		bt_prologue(L, bt_qry(P, L, cur), bt_qual_at(P, L, cur));
		break; }

	case PC_LF:
	case PC_POS2: {
		/* the LF step (ebwt_search_backtrack.h:530-568) and what follows it (569-739) */
		const uint32_t c = L.c, q = L.q, d = L.d;
		const uint32_t cur = L.qlen - d - 1;
		uint32_t tops[4] = { 0, 0, 0, 0 }, bots[4] = { 0, 0, 0, 0 };
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
			uint32_t t = bt_lf(ix, bA, L.ltop, c), b = bt_lf(ix, bB, L.lbot, c);
			L.top = t; L.bot = b;
			L.s_lf += 2;
		} else if (L.lfk == LFK_FCHR) {
			/* first quartet from fchr[] (ebwt_search_backtrack.h:531-543) */
			tops[0] = ix.fchr[0]; tops[1] = ix.fchr[1]; tops[2] = ix.fchr[2]; tops[3] = ix.fchr[3];
			bots[0] = ix.fchr[1]; bots[1] = ix.fchr[2]; bots[2] = ix.fchr[3]; bots[3] = ix.fchr[4];
			if (c < 4) { L.top = tops[c]; L.bot = bots[c]; }
		}
		if (L.top != L.bot) { L.ltop = L.top; L.lbot = L.bot; }   /* SideLocus::initFromTopBot */
		if (d >= L.rowd0) {
			const uint32_t ri = bt_row_idx(L, d);
			uint32_t el = (c < 4) ? (1u << c) : 0u;                   /* eliminate() */
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
			}
			S.elims[ri] = (uint8_t)el;
		}
		L.f_bdm = 0; L.f_must = 0; L.f_invHH = 0; L.f_invExact = 0;
		uint32_t reportedPartial = 0;
		if (cur == 0 && L.top < L.bot && L.stackDepth < L.reportPartials && L.reportPartials > 0) {
			if (L.altNum > 0) L.f_bdm = 1;
			if (L.stackDepth > 0) { bt_report_partial(L, P, S, L.stackDepth); reportedPartial = 1; }
		}
		if (cur == 0 && L.stackDepth == 0 && L.bot > L.top && !L.reportExacts) { L.f_invExact = 1; L.f_bdm = 1; }
		if (L.halfAndHalf) {
			if ((d == (L.depth5 - 1)) && L.top < L.bot) {
				L.f_invHH = (L.stackDepth == 0);
				if (L.stackDepth == 0 && L.altNum > 0) { L.f_bdm = 1; L.f_must = 1; }
				else if (L.stackDepth == 0) { L.ret = 0; L.pc = PC_FRAME_RET; break; }
			} else if ((d == (L.depth3 - 1)) && L.top < L.bot) {
				uint32_t lo = 0, hi = 0;
				for (uint32_t i = 0; i < L.stackDepth; i++) {
					uint32_t dd = L.qlen - bt_mm_pos(S, i) - 1;
					if (dd < L.depth5) hi++; else if (dd < L.depth3) lo++;
				}
				L.f_invHH = (lo == 0 || hi == 0);
				if ((L.stackDepth < 2 || L.f_invHH) && L.altNum > 0) { L.f_must = 1; L.f_bdm = 1; }
				else if (L.stackDepth < 2) { L.ret = 0; L.pc = PC_FRAME_RET; break; }
			}
		}
		if (cur == 0 && L.bot > L.top && !L.f_invHH && !L.f_invExact && !reportedPartial) {
			bt_report_begin(L, P, S, L.stackDepth, L.top, L.bot, L.ham, SITE_MAIN);
			break;
		}
		if ((L.top == L.bot || L.f_bdm) && L.altNum > 0) { L.pc = PC_BTLOOP; break; }   /* mismatch with alternatives: next transition */
		/* match (or dead end): the tail of the loop body (1066-1078) and, on a match, the next position's
		 * prologue with the query character prefetched in the fetch stage — one transition per position */
		if (L.f_must || L.f_invHH || L.f_invExact) { L.ret = 0; L.pc = PC_FRAME_RET; break; }
		if (L.top == L.bot) { L.ret = 0; L.pc = PC_FRAME_RET; break; }
		L.d = d + 1;
		if (L.d >= L.qlen) { L.pc = PC_POS; break; }
		if (L.halfAndHalf && !bt_hh_check_top(L, S)) { L.ret = 0; L.pc = PC_FRAME_RET; break; }
		bt_prologue(L, nc, nq);
		break; }

	case PC_BTLOOP: {
		/* while((top == bot || backtrackDespiteMatch) && altNum > 0) (ebwt_search_backtrack.h:743-971) */
		if (!((L.top == L.bot || L.f_bdm) && L.altNum > 0)) { L.pc = PC_POS_END; break; }
		uint32_t i = L.d, j = 0, bttop = 0, btbot = 0, btham = L.ham, btcint = 0;
		if (L.eligibleNum > 1 || L.elignore) {
			for (;; i--) {
				uint32_t icur = L.qlen - i - 1;
				uint32_t qi = bt_qual_at(P, L, icur);
				uint32_t ri = bt_row_idx(L, i);
				uint32_t el = (i >= L.rowd0) ? S.elims[ri] : 15u;
				if ((qi == L.lowAltQual || !L.considerQuals) && el != 15) {
					uint32_t posSz = 0;
					for (j = 0; j < 4; j++) if ((el & (1u << j)) == 0) posSz += bt_pair_bot(S, ri, j) - bt_pair_top(S, ri, j);
					uint32_t r = bt_rand_next(L.rnd) % posSz;
					for (j = 0; j < 4; j++) {
						if ((el & (1u << j)) == 0) {
							uint32_t ptop = bt_pair_top(S, ri, j), pbot = bt_pair_bot(S, ri, j);
							uint32_t spread = pbot - ptop;
							if (r < spread) { bttop = ptop; btbot = pbot; btham += bt_mm_penalty(L.maqPenalty, qi); btcint = j; break; }
							r -= spread;
						}
					}
					break;
				}
				if (i == L.depth) break;   /* cannot happen while eligibleNum > 0 */
			}
		} else {
			i = L.eli; bttop = L.eltop; btbot = L.elbot; btham += L.elham; j = L.elcint; btcint = L.elcint;
		}
		const uint32_t icur = L.qlen - i - 1;
		uint32_t btUnrevOff = L.unrevOff, btOneRevOff = L.oneRevOff, btTwoRevOff = L.twoRevOff, btThreeRevOff = L.threeRevOff;
		if (i < L.oneRevOff) { btUnrevOff = L.oneRevOff; btOneRevOff = L.twoRevOff; btTwoRevOff = L.threeRevOff; }
		else if (i < L.twoRevOff) { btOneRevOff = L.twoRevOff; btTwoRevOff = L.threeRevOff; }
		else if (i < L.threeRevOff) { btTwoRevOff = L.threeRevOff; }
		if (L.stackDepth >= P.FCAP) { L.flags |= BT_FLAG_FRAME_OVF; L.found = 0; L.pc = PC_FINISH_READ; break; }
		BtFrame &F = S.frames[L.stackDepth];
		F.mm_pos = (uint16_t)icur; F.mm_refc = (uint8_t)btcint;       /* _mms[stackDepth], _refcs[stackDepth] */
		L.bt_i = i; L.bt_j = j; L.bttop = bttop; L.btbot = btbot; L.btham = btham;
		if (i + 1 == L.qlen) {
			bt_report_begin(L, P, S, L.stackDepth + 1, bttop, btbot, btham, SITE_BT);
			break;
		}
		bool rejump = L.halfAndHalf && !L.disableFtab && L.rev2_0 == L.rev3_0 && i + 1 < (uint32_t)ix.ftabChars && (uint32_t)ix.ftabChars <= L.depth5;
		uint32_t ndepth = i + 1, ntop = bttop, nbot = btbot;
		if (rejump) {
			/* ftab re-jump with the substituted character (ebwt_search_backtrack.h:908-952) */
			const uint32_t ftabChars = (uint32_t)ix.ftabChars;
			uint32_t ftabOff = bt_qry(P, L, L.qlen - ftabChars);
			for (uint32_t jj = ftabChars - 1; jj > 0; jj--) {
				ftabOff <<= 2;
				if (L.qlen - jj == icur) ftabOff |= btcint; else ftabOff |= bt_qry(P, L, L.qlen - jj);
			}
			ntop = bt_ftab_hi(ix, ftabOff); nbot = bt_ftab_lo(ix, ftabOff + 1);
			L.s_ftab++;
			ndepth = ftabChars;
			if (ntop == nbot) { L.ret = 0; L.pc = PC_CHILD_RET; break; }
		}
		/* PUSH: suspend this frame */
		F.top = L.top; F.bot = L.bot; F.eligibleSz = L.eligibleSz; F.eltop = L.eltop; F.elbot = L.elbot; F.btspread = btbot - bttop;
		F.depth = (uint16_t)L.depth; F.d = (uint16_t)L.d; F.unrevOff = (uint16_t)L.unrevOff; F.oneRevOff = (uint16_t)L.oneRevOff;
		F.twoRevOff = (uint16_t)L.twoRevOff; F.threeRevOff = (uint16_t)L.threeRevOff; F.ham = (uint16_t)L.ham; F.altNum = (uint16_t)L.altNum;
		F.eligibleNum = (uint16_t)L.eligibleNum; F.eli = (uint16_t)L.eli; F.rowbase = (uint16_t)L.rowbase; F.rowd0 = (uint16_t)L.rowd0;
		F.bt_i = (uint16_t)i; F.lowAltQual = (uint8_t)L.lowAltQual; F.elham = (uint8_t)L.elham; F.elcint = (uint8_t)L.elcint; F.bt_j = (uint8_t)j;
		F.flags = (uint8_t)((L.elignore ? FF_ELIGNORE : 0) | (L.f_bdm ? FF_BDM : 0) | (L.f_must ? FF_MUST : 0) | (L.f_invHH ? FF_INVHH : 0) |
		                    (L.f_invExact ? FF_INVEXACT : 0) | (L.disableFtab ? FF_DISABLEFTAB : 0));
		uint32_t nrowbase = L.rowbase + ((L.d >= L.rowd0) ? (L.d - L.rowd0 + 1) : 0);
		bt_frame_enter(L, P, L.stackDepth + 1, ndepth, btUnrevOff, btOneRevOff, btTwoRevOff, btThreeRevOff, ntop, nbot, btham, nrowbase, 0);
		break; }

	case PC_FRAME_RET: {
		if (L.stackDepth == 0) { L.pc = PC_BT_END; break; }
		/* POP: resume the parent after its recursive call returned L.ret */
		const BtFrame &F = S.frames[L.stackDepth - 1];
		L.stackDepth--;
		L.top = F.top; L.bot = F.bot; L.eligibleSz = F.eligibleSz; L.eltop = F.eltop; L.elbot = F.elbot;
		L.depth = F.depth; L.d = F.d; L.unrevOff = F.unrevOff; L.oneRevOff = F.oneRevOff; L.twoRevOff = F.twoRevOff; L.threeRevOff = F.threeRevOff;
		L.ham = F.ham; L.altNum = F.altNum; L.eligibleNum = F.eligibleNum; L.eli = F.eli; L.rowbase = F.rowbase; L.rowd0 = F.rowd0;
		L.bt_i = F.bt_i; L.bt_j = F.bt_j; L.lowAltQual = F.lowAltQual; L.elham = F.elham; L.elcint = F.elcint;
		L.elignore = (F.flags & FF_ELIGNORE) != 0; L.f_bdm = (F.flags & FF_BDM) != 0; L.f_must = (F.flags & FF_MUST) != 0;
		L.f_invHH = (F.flags & FF_INVHH) != 0; L.f_invExact = (F.flags & FF_INVEXACT) != 0; L.disableFtab = (F.flags & FF_DISABLEFTAB) != 0;
		L.bttop = 0; L.btbot = F.btspread;
		L.pc = PC_CHILD_RET;
	} /* fallthrough */

	case PC_CHILD_RET: {
		/* after the recursive call (ebwt_search_backtrack.h:972-1064) */
		if (L.ret) { L.pc = PC_FRAME_RET; break; }
		if (L.bailed || (L.halfAndHalf && L.maxBts > 0 && L.numBts >= L.maxBts)) { L.bailed = 1; L.ret = 0; L.pc = PC_FRAME_RET; break; }
		const uint32_t i = L.bt_i, j = L.bt_j;
		{
			uint32_t ri = bt_row_idx(L, i);
			S.elims[ri] = (uint8_t)(S.elims[ri] | (1u << j));
		}
		L.eligibleSz -= (L.btbot - L.bttop);
		L.eligibleNum--;
		L.elignore = 1;
		L.altNum--;
		if (L.altNum == 0) { L.ret = 0; L.pc = PC_FRAME_RET; break; }
		if (L.eligibleNum == 0 && L.considerQuals) {
			/* re-scan the frame for the next-lowest quality (1004-1058) */
			L.lowAltQual = 0xff;
			for (uint32_t k = L.d;; k--) {
				if (k < L.unrevOff) break;
				uint32_t kq = bt_qual_at(P, L, L.qlen - k - 1);
				bool kAlt = (L.ham + bt_mm_penalty(L.maqPenalty, kq) <= L.qualThresh);
				bool kOverrides = false;
				if (kAlt) {
					if (kq < L.lowAltQual) kOverrides = true;
					if (kq <= L.lowAltQual) {
						uint32_t ri = bt_row_idx(L, k);
						uint32_t el = S.elims[ri];
						for (uint32_t l = 0; l < 4; l++) {
							if ((el & (1u << l)) == 0) {
								uint32_t ptop = bt_pair_top(S, ri, l), pbot = bt_pair_bot(S, ri, l);
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
				if (k == L.depth || k == 0) break;
			}
		}
		L.pc = PC_BTLOOP;
		break; }

	case PC_POS_END: {
		/* (ebwt_search_backtrack.h:1066-1078) */
		if (L.f_must || L.f_invHH || L.f_invExact) { L.ret = 0; L.pc = PC_FRAME_RET; break; }
		if (L.top == L.bot && L.altNum == 0) { L.ret = 0; L.pc = PC_FRAME_RET; break; }
		L.d++;
		L.pc = PC_POS;
		break; }

	case PC_REPORT_ROW: {
		/* loop of reportFullAlignment (ebwt_search_backtrack.h:1539-1564) */
		uint32_t spread = L.rep_bot - L.rep_top;
		if (L.rep_i >= spread) { L.ret = 0; L.pc = PC_REPORT_RET; break; }
		uint32_t ri = L.rep_r + L.rep_i;
		if (ri >= L.rep_bot) ri -= spread;
		L.crow = ri; L.cjumps = 0;
		L.pc = (((ri & ix.offMask) != ri) && ri != ix.zOff) ? PC_CHASE : PC_RESOLVE;
		break; }

	case PC_CHASE: {
		/* one step of the row walk of Ebwt::reportChaseOne (ebwt.h:2727-2734): mapLF(l) */
		uint32_t c = bt_row_l(bA, L.crow);
		uint32_t nr = bt_lf(ix, bA, L.crow, c);
		L.crow = nr; L.cjumps++;
		L.s_lf++; L.s_chase++;
		if (((nr & ix.offMask) != nr) && nr != ix.zOff) break;   /* stay in PC_CHASE */
		L.pc = PC_RESOLVE;
	} /* fallthrough */

	case PC_RESOLVE: {
		/* marked row reached (ebwt.h:2735-2755), then Ebwt::report (2635-2682) */
		uint32_t off;
		if (L.crow == ix.zOff) off = L.cjumps;
		else { off = BT_LDG(ix.offs + (L.crow >> ix.offRate)) + L.cjumps; L.s_offs++; }
		uint32_t tidx = 0, toff = 0;
		bool stop = false;
		if (bt_joined_to_text(ix, L.qlen, off, tidx, toff)) stop = bt_sink_report(L, P, S, tidx, toff);
		if (stop) { L.ret = 1; L.pc = PC_REPORT_RET; }
		else { L.rep_i++; L.pc = PC_REPORT_ROW; }
		break; }

	case PC_REPORT_RET: {
		switch (L.rep_site) {
		case SITE_MAIN:
			if (!L.ret) { L.top = L.bot; L.pc = PC_BTLOOP; }
			else L.pc = PC_FRAME_RET;
			break;
		case SITE_BT: L.pc = PC_CHILD_RET; break;
		case SITE_END: L.pc = PC_FRAME_RET; break;
		default: L.pc = PC_BT_END; break;
		}
		break; }

	case PC_BT_END: {
		/* tail of backtrack(depth, top, bot, ...) and finalize() (ebwt_search_backtrack.h:348-352, 303-324) */
		L.numBts = 0; L.bailed = 0;
		if (L.reportPartials > 0 && L.npart > 0) L.ret = 1;
		L.done = L.ret;
		L.pc = PC_PHASE;
		break; }

	default: break;
	}
}

/* Begin a read: GET_READ (ebwt_search.cpp:923-961) */
BT_FN void bt_begin_read(BtLane &L, const BtKParams &P, uint32_t rid) {
	L.rid = rid;
	L.roff = P.roff[rid];
	L.rlen = (uint32_t)(P.roff[rid + 1] - L.roff);
	L.seed = P.seeds[rid];
	L.rseq = P.seq + L.roff; L.rqual = P.qual + L.roff; L.hasN = 1;   /* the kernel may re-point these at its staging copy */
	L.found = 0; L.flags = 0; L.ph = 0; L.done = 0; L.npart = 0; L.nmuts = 0; L.pal_i = 0;
	L.qualThresh = P.pol.mode == 0 ? 0xffffffffu : P.pol.qualThresh;
	L.maxBts = P.pol.mode == 0 ? 0xffffffffu : P.pol.maxBts;
	L.maqPenalty = P.pol.mode == 0 ? 1u : (uint32_t)P.pol.maqRound;
	L.pc = L.rlen > 0 ? PC_PHASE : PC_FINISH_READ;
}

/* HitSinkPerThread::finishRead (hit.h:741-786): the host applies -m suppression / -k truncation
 * from `found`; the kernel stored the first min(found, n, slots) hits. */
BT_FN void bt_finish_read(BtLane &L, const BtKParams &P) {
	P.found[L.rid] = L.found;
	P.flags[L.rid] = L.flags;
}

/* One iteration of the lane loop: the (warp-converged) fetch stage followed by one transition.
 * Fetch stage = the rank block(s) of the pending LF / chase step plus the NEXT position's query
 * character and quality, all independent loads in flight together. */
BT_FN void bt_iter(BtLane &L, const BtKParams &P, const BtScratch &S) {
	if (L.flags & BT_FLAG_PART_OVF) { L.found = 0; L.pc = PC_FINISH_READ; return; }   /* retried with a larger workspace */
	BtBlock bA, bB;
	bA.occ.x = bA.occ.y = bA.occ.z = bA.occ.w = 0; bA.hi = bA.lo = 0;
	const bool isLF = (L.pc == PC_LF), isChase = (L.pc == PC_CHASE);
	uint32_t nc = 4, nq = 0;
	if (isLF || isChase) {
		const BtDevIndex &ix = P.ix[L.ebwtSel];
		const uint32_t rowA = isChase ? L.crow : L.ltop;
		bA = bt_load_block(ix, rowA);
		L.s_blk++;
	}
	bB = bA;
	if (isLF && L.lfk != LFK_ONE && (L.lbot >> 6) != (L.ltop >> 6)) {
		bB = bt_load_block(P.ix[L.ebwtSel], L.lbot);
		L.s_blk++;
	}
	if ((isLF || L.pc == PC_POS2) && L.d + 1 < L.qlen) {
		const uint32_t ncur = L.qlen - L.d - 2;
		nc = bt_qry(P, L, ncur); nq = bt_qual_at(P, L, ncur);
	}
	L.s_iter++;
	bt_step(L, P, S, bA, bB, nc, nq);
}
