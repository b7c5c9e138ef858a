/*
 * bt_best.cuh — the reference's "stateful" best-first search path, restated for one CUDA thread per read.
 *
 * Covers SURVEY.md §8 rows a14-a16: EbwtRangeSource (ebwt_search_backtrack.h:1788-2599), Branch / RangeState /
 * PathManager / BranchQueue+CostCompare (range_source.h:240-1580), the Single / CostAware range-source drivers
 * (range_source.h:1614-2463), EbwtRangeSourceDriver::initRangeSource and EbwtSeededRangeSourceDriver
 * (ebwt_search_backtrack.h:2670-3141), UnpairedAlignerV2 (aligner.h:381-599), RangeChaser / RowChaser
 * (range_chaser.h:22-268, row_chaser.h:20-182) and the sinks' stop rules (hit.h:969-985, 1070-1129, 1201-1209).
 * Selected by --best, --strata, -M and -v 3 (ebwt_search.cpp:776,851-854,877-882).
 *
 * Design (not the reference's): the reference keeps one object graph per worker thread — a tree of driver
 * objects, each with its own PathManager, three chunked AllocOnlyPools and a std::priority_queue.  Here every
 * read owns one bump ARENA of 32-bit words in HBM; driver nodes, branches, per-position range states, heap
 * arrays and driver lists are carved from it and die with the read.  Branch ids are a per-PathManager counter
 * that mimics AllocOnlyPool::lastId()/free() (pool.h:239-277,329-331) because CostCompare breaks ties on them.
 * The binary heap reproduces libstdc++'s push_heap/pop_heap moves exactly: Branch keys mutate while a branch
 * sits at the top of the queue (extend(), curtail() without a cost change), so the pop order of the reference
 * depends on the heap's physical layout, not only on the comparison.
 * Range states are stored only for positions where an edit can ever be considered (i >= depth0 - rdepth,
 * range_source.h:671-672,890-891), in chained blocks of 4 that a branch acquires as it extends (most branches die young).  Arena exhaustion sets BT_FLAG_STACK_OVF and the read is re-run by a pass
 * with a larger arena (the reference's own limit is --chunkmbs; it skips the read instead).
 *
 * Compiles for the device and, for tests/host_emu only, for the host (BT_HOST_EMU).
 */
#pragma once
#include "bt_core.cuh"

#define BF_MAX_TOP 8
/* No legitimate search comes near this many branch-extension steps for one read (the longest seen, -a -y on short reads, is
 * ~10^7); a read that does is abandoned with BT_FLAG_FRAME_OVF instead of occupying a GPU lane without bound. */
#define BF_MAX_STEPS 0x40000000u

enum { BF_PIN_BEGINNING = 1, BF_PIN_LEN, BF_PIN_HI_HALF, BF_PIN_SEED };   /* SearchConstraintExtent, ebwt_search_backtrack.h:2658-2663 */
enum { BF_KIND_SRC = 0, BF_KIND_SEEDED = 1 };

struct BfSrcCfg {            /* constructor arguments of one EbwtRangeSource + EbwtRangeSourceDriver */
	uint8_t ebwtSel, fw, reportExacts, hh, seed, nudgeLeft, useBtCnt, mate;   /* mate: 0 = the read / mate 1, 1 = mate 2 */
	uint8_t rev[4];
};
struct BfTopCfg { uint32_t kind; BfSrcCfg a, b; };   /* a: the driver (or the seedling generator); b: the per-seedling extension driver */
struct BfProg {
	uint32_t ntop, seedLen, qualLim, strandFix; BfTopCfg top[BF_MAX_TOP];
	BfTopCfg top2[BF_MAX_TOP];   /* paired-end: mate 2's drivers (the -v factories configure the two mates differently) */
	/* paired-end (PairedBWAlignerV1): which of the four per-mate, per-strand driver lists exist (do1Fw, do1Rc, do2Fw, do2Rc),
	 * the reference-scan policy of the opposite mate (RefAligner family) and the insert-size window */
	uint32_t paired, pairedV2, doList[4], refMms, refSeedLen, refQualMax, minIns, maxIns, fw1, fw2, mixedAttemptLim, symCeiling;
};

/* BitPairReference (reference.h:20-723): the 2-bit reference of X.4.ebwt with the stretch records of X.3.ebwt */
struct BtDevRef {
	const uint32_t *recs;          /* 2 words per record: off (ambiguous characters before the stretch), len */
	const uint32_t *refRecOffs;    /* nRefs + 1: first record of every reference that has unambiguous characters */
	const uint32_t *refOffs;       /* nRefs + 1: offset of its first character in buf */
	const uint32_t *approxLen;     /* nRefs */
	const uint8_t *buf;            /* 4 characters per byte */
	uint32_t nRefs;
};

/* an edit is one arena word: pos | chr << 16 (pos = depth of the edit in a Branch, query offset in a Range; chr = reference base) */
#define BF_EDIT(pos, chr) ((uint32_t)(pos) | ((uint32_t)(chr) << 16))
#define BF_EDIT_POS(e) ((e) & 0xffffu)
#define BF_EDIT_CHR(e) (((e) >> 16) & 0xffu)
struct BfRS { uint32_t tops[4], bots[4], eq; };       /* RangeState: eq bits 0-3 = mm{A,C,G,T} eliminated, 8-14 = quallo, 31 = eliminated_ */
#define BF_RS_WORDS 9
#define BF_RS_ELIM 0x80000000u
#define BF_RS_BLK 4              /* range states live in chained blocks of 4 positions (word 0 = next block), carved when a branch first reaches them */
#define BF_RS_BLK_WORDS (1 + BF_RS_BLK * BF_RS_WORDS)

struct BfBranch {
	uint32_t id, top, bot, ltop, lbot, ranges, edits;      /* ranges: first block of the chain (0 = none yet); edits: nedits arena words */
	uint32_t tipBlk, curBlk;                               /* last block of the chain; the block of the latest access (scans are sequential) */
	uint16_t depth0, depth1, depth2, depth3, rdepth, len, cost, ham, rangesSz, i0, delayedCost, nedits;
	uint8_t curtailed, exhausted, delayedIncrease, lbotValid;
	uint16_t nblk, curK;                                   /* blocks in the chain; index of curBlk */
};
#define BF_BRANCH_WORDS ((uint32_t)(sizeof(BfBranch) / 4))

struct BfHdr { uint8_t kind, done, foundRange, fw; uint16_t minCost, minCostAdj; uint8_t mate, pad[3]; };   /* mate: RangeSourceDriver::mate1() ? 0 : 1 */

struct BfSrc {               /* EbwtRangeSourceDriver + its EbwtRangeSource + its PathManager */
	BfHdr h;
	BfSrcCfg cfg;
	uint8_t rsDone, rsFound, skipping, seedValid, seedNmm, viewRev, viewComp, pad3;
	uint16_t seedCost, rCost, pmMinCost, pad0;
	uint32_t rNmm, rEdits, rEditsCap;                      /* the current Range's mismatches: arena words */
	uint32_t qlen, depth5, depth3, off0, off1, off2, off3, rnd;
	uint16_t seedMms[3]; uint8_t seedRefc[3], pad1;
	uint16_t ovPos[3], pad2;
	uint32_t rTop, rBot;
	uint32_t heapOff, heapCap, heapSz, bcur;
};
#define BF_SRC_WORDS ((uint32_t)(sizeof(BfSrc) / 4))

struct BfCA {                /* CostAwareRangeSourceDriver */
	uint32_t rssOff, rssCap, nRss, actOff, actCap, nAct, rnd, lastRange, delayedRange;
	uint16_t minCost; uint8_t done, foundRange;
	uint32_t paired;             /* calcPaired(): drivers of both mates in one list (PairedBWAlignerV2) */
};
struct BfSeeded { BfHdr h; BfSrcCfg fact; uint32_t seedgen; BfCA full; };
#define BF_SEEDED_WORDS ((uint32_t)(sizeof(BfSeeded) / 4))

struct BfKParams {
	BtDevIndex ix[2];
	BtPolicy pol;
	BfProg prog;
	BtDevRef ref;
	const uint8_t *seq, *qual; const uint64_t *roff; const uint32_t *seeds; const uint32_t *sel; uint32_t nwork;
	uint32_t *found, *flags, *hits; uint32_t slots, mm_cap, rec_words;
	uint32_t *arena; uint32_t arenaWords;     /* per lane */
	/* The arenas of a tier are a pool shared by every context of the index: a block claims one pool entry (`lanes` arenas) for its
	 * lifetime by setting a bit of poolMask, so the memory is sized by how many blocks can be resident, not by how many batches are
	 * in flight. */
	unsigned *poolMask; uint32_t poolBlocks;
	unsigned long long *stats;
};

struct BfCtx {
	const BfKParams *P;
	const uint8_t *seqM[2], *qualM[2];                    /* [0]: the read (mate 1), [1]: mate 2 */
	uint32_t rid, rlenM[2], seedM[2];
	uint32_t *A; uint32_t acap, atop, amax;               /* amax: high-water mark of the arena (diagnostics) */
	uint32_t steps;                                        /* branch-extension steps of this read: a watchdog, see BF_MAX_STEPS */
	uint32_t flags, found, randA;
	int32_t bestStratum, btCnt;
	BfRS spare;                                            /* where range-state writes land once the arena is exhausted */
	BfCA top;
	uint32_t s_lfex, s_lf, s_chase, s_ftab, s_offs, s_bt;
};

/* ---- arena ----------------------------------------------------------------------------------- */
BT_FN uint32_t bf_alloc(BfCtx &X, uint32_t words) {
	if (X.atop + words > X.acap) { X.flags |= BT_FLAG_STACK_OVF; return 0; }
	const uint32_t off = X.atop; X.atop += words;
	if (X.atop > X.amax) X.amax = X.atop;
	return off;
}
BT_FN uint32_t bf_alloc_zero(BfCtx &X, uint32_t words) {
	const uint32_t off = bf_alloc(X, words);
	if (off) for (uint32_t i = 0; i < words; i++) X.A[off + i] = 0;
	return off;
}
BT_FN void bf_free_top(BfCtx &X, uint32_t off, uint32_t words) { if (off && off + words == X.atop) X.atop = off; }
#define BF_AT(T, X, off) ((T *)((X).A + (off)))

/* a growable list of node refs */
BT_FN void bf_vec_push(BfCtx &X, uint32_t &off, uint32_t &cap, uint32_t &n, uint32_t v) {
	if (n == cap) {
		const uint32_t ncap = cap ? cap * 2 : 4;
		const uint32_t noff = bf_alloc(X, ncap);
		if (!noff) return;
		for (uint32_t i = 0; i < n; i++) X.A[noff + i] = X.A[off + i];
		bf_free_top(X, off, cap);
		off = noff; cap = ncap;
	}
	X.A[off + n++] = v;
}

/* ---- read views (EbwtRangeSource::setQuery, ebwt_search_backtrack.h:1831-1866) ---------------- */
BT_FN uint32_t bf_qry(const BfCtx &X, const BfSrc &s, uint32_t cur) {
	for (uint32_t k = 0; k < s.seedNmm; k++) if (s.ovPos[k] == cur) return s.seedRefc[k];   /* qryBuf_ */
	uint32_t c = X.seqM[s.cfg.mate][s.viewRev ? (X.rlenM[s.cfg.mate] - 1 - cur) : cur];
	if (s.viewComp && c < 4) c ^= 3;
	return c;
}
BT_FN uint32_t bf_qualch(const BfCtx &X, const BfSrc &s, uint32_t cur) { return X.qualM[s.cfg.mate][s.viewRev ? (X.rlenM[s.cfg.mate] - 1 - cur) : cur]; }
BT_FN uint32_t bf_phred(uint32_t ch) { return ch >= 33 ? ch - 33 : 0; }

/* ---- PathManager / BranchQueue ------------------------------------------------------------------ */
/* CostCompare (range_source.h:1103-1135): true when b goes before a */
BT_FN bool bf_before(const BfBranch &a, const BfBranch &b) {
	if (a.cost == b.cost) {
		const bool aUn = a.curtailed || a.exhausted, bUn = b.curtailed || b.exhausted;
		if (bUn && !aUn) return false;
		if (aUn && !bUn) return true;
		const uint32_t ta = (uint32_t)a.rdepth + a.len, tb = (uint32_t)b.rdepth + b.len;
		if (ta != tb) return ta < tb;
		return b.id < a.id;
	}
	return b.cost < a.cost;
}
#define BF_BR(X, ref) BF_AT(BfBranch, X, ref)

BT_FN void bf_heap_sift_up(BfCtx &X, uint32_t *h, uint32_t hole, uint32_t topIdx, uint32_t value) {   /* std::__push_heap */
	while (hole > topIdx) {
		const uint32_t parent = (hole - 1) / 2;
		if (!bf_before(*BF_BR(X, h[parent]), *BF_BR(X, value))) break;
		h[hole] = h[parent]; hole = parent;
	}
	h[hole] = value;
}
BT_NOINLINE void bf_pm_push(BfCtx &X, BfSrc &s, uint32_t br) {          /* PathManager::push, range_source.h:1352-1358 */
	bf_vec_push(X, s.heapOff, s.heapCap, s.heapSz, br);
	if (X.flags & BT_FLAG_STACK_OVF) return;
	uint32_t *h = X.A + s.heapOff;
	bf_heap_sift_up(X, h, s.heapSz - 1, 0, br);
	s.pmMinCost = BF_BR(X, h[0])->cost;
}
BT_NOINLINE uint32_t bf_pm_pop(BfCtx &X, BfSrc &s) {                    /* PathManager::pop, range_source.h:1334-1347 */
	uint32_t *h = X.A + s.heapOff;
	const uint32_t popped = h[0];
	const uint32_t len = s.heapSz;
	if (len > 1) {                                                       /* std::pop_heap = __adjust_heap(0, len-1, last) */
		const uint32_t value = h[len - 1];
		h[len - 1] = h[0];
		const uint32_t n = len - 1;
		uint32_t hole = 0, child = 0;
		while (child < (n - 1) / 2) {
			child = 2 * (child + 1);
			if (bf_before(*BF_BR(X, h[child]), *BF_BR(X, h[child - 1]))) child--;
			h[hole] = h[child]; hole = child;
		}
		if ((n & 1) == 0 && child == (n - 2) / 2) {
			child = 2 * (child + 1);
			h[hole] = h[child - 1]; hole = child - 1;
		}
		bf_heap_sift_up(X, h, hole, 0, value);
	}
	s.heapSz = len - 1;
	/* minCost = branchQ_.front()->cost_ : on an empty queue this reads the slot the popped pointer still occupies */
	s.pmMinCost = BF_BR(X, s.heapSz ? h[0] : popped)->cost;
	return popped;
}
BT_FN void bf_pm_reset(BfSrc &s) { s.heapSz = 0; s.bcur = 0; s.pmMinCost = 0; }   /* PathManager::reset */

/* ---- Branch ------------------------------------------------------------------------------------- */
BT_FN void bf_branch_prep(BfBranch &b) {                                 /* Branch::prep / the locus part of init */
	if (b.bot > b.top + 1) { b.ltop = b.top; b.lbot = b.bot; b.lbotValid = 1; }
	else if (b.bot > b.top) { b.ltop = b.top; b.lbotValid = 0; }
}
/* The reference reserves qlen - rdepth RangeStates per Branch (range_source.h:570-578); most branches die within a few
 * positions, so here a branch owns a chain of blocks of BF_RS_BLK zeroed states and a block is carved on first touch
 * (positions are first touched in increasing order: init marks [i0, len), extension then reaches one position at a time). */
BT_NOINLINE BfRS *bf_rs(BfCtx &X, BfBranch &b, uint32_t i) {
	const uint32_t idx = i - b.i0, k = idx / BF_RS_BLK;
	uint32_t blk;
	if (b.curBlk && k == b.curK) blk = b.curBlk;
	else if (b.curBlk && k == (uint32_t)b.curK + 1 && X.A[b.curBlk]) blk = X.A[b.curBlk];
	else if (k + 1 == b.nblk) blk = b.tipBlk;
	else if (k < b.nblk) { blk = b.ranges; for (uint32_t j = 0; j < k; j++) blk = X.A[blk]; }
	else {
		blk = b.tipBlk;
		while (b.nblk <= k) {                                             /* normally one step */
			const uint32_t nb = bf_alloc_zero(X, BF_RS_BLK_WORDS);
			if (!nb) { X.spare.eq = 0; return &X.spare; }
			if (b.nblk) X.A[b.tipBlk] = nb; else b.ranges = nb;
			b.tipBlk = nb; b.nblk++;
			blk = nb;
		}
	}
	b.curBlk = blk; b.curK = (uint16_t)k;
	return BF_AT(BfRS, X, blk + 1 + (idx % BF_RS_BLK) * BF_RS_WORDS);
}
BT_FN bool bf_eliminated(BfCtx &X, BfBranch &b, uint32_t i) {      /* Branch::eliminated, range_source.h:619-634 */
	if (i <= b.len && i < b.rangesSz) return (bf_rs(X, b, i)->eq & BF_RS_ELIM) != 0;
	return true;
}
/* Branch::init (range_source.h:531-607).  Returns the branch ref or 0 when the arena is exhausted. */
BT_NOINLINE uint32_t bf_branch_new(BfCtx &X, BfSrc &s, uint32_t qlen, uint32_t d0, uint32_t d1, uint32_t d2, uint32_t d3,
                                   uint32_t rdepth, uint32_t len, uint32_t cost, uint32_t ham, uint32_t top, uint32_t bot,
                                   const BfBranch *parent, uint32_t newEdit) {
	const uint32_t ref = bf_alloc(X, BF_BRANCH_WORDS);
	if (!ref) return 0;
	BfBranch &b = *BF_BR(X, ref);
	b.nedits = 0; b.edits = 0;
	if (parent) {                                                        /* edits_ of the parent plus the edit that starts this branch */
		b.nedits = (uint16_t)(parent->nedits + 1);
		b.edits = bf_alloc(X, b.nedits);
		if (!b.edits) return 0;
		for (uint32_t i = 0; i < parent->nedits; i++) X.A[b.edits + i] = X.A[parent->edits + i];
		X.A[b.edits + parent->nedits] = newEdit;
	}
	s.bcur++;                                                            /* bpool.alloc(); id = bpool.lastId() */
	b.id = s.bcur;
	b.delayedCost = 0;
	b.depth0 = (uint16_t)d0; b.depth1 = (uint16_t)d1; b.depth2 = (uint16_t)d2; b.depth3 = (uint16_t)d3;
	b.rdepth = (uint16_t)rdepth; b.len = (uint16_t)len; b.cost = (uint16_t)cost; b.ham = (uint16_t)ham;
	b.top = top; b.bot = bot; b.ltop = 0; b.lbot = 0; b.lbotValid = 0;
	bf_branch_prep(b);
	b.rangesSz = (uint16_t)(qlen - rdepth);
	const uint32_t i0 = d0 > rdepth ? d0 - rdepth : 0;
	b.i0 = (uint16_t)(i0 < b.rangesSz ? i0 : b.rangesSz);
	b.ranges = 0; b.tipBlk = 0; b.curBlk = 0; b.nblk = 0; b.curK = 0;
	b.curtailed = 0; b.exhausted = 0; b.delayedIncrease = 0;
	for (uint32_t i = b.i0; i < len && i < b.rangesSz; i++) bf_rs(X, b, i)->eq |= BF_RS_ELIM;
	return ref;
}
BT_FN void bf_branch_free(BfCtx &X, BfSrc &s, uint32_t ref) {            /* Branch::free: only the newest slot is really reclaimed */
	BfBranch &b = *BF_BR(X, ref);
	if (b.nblk == 1) bf_free_top(X, b.tipBlk, BF_RS_BLK_WORDS);           /* the common early death: struct, edits and one block sit on top */
	if (b.id == s.bcur && s.bcur > 0) s.bcur--;
	if (b.edits) bf_free_top(X, b.edits, b.nedits);
	bf_free_top(X, ref, BF_BRANCH_WORDS);
}
/* Branch::curtail (range_source.h:876-923); the trimming of ranges_ only returns memory */
BT_NOINLINE void bf_branch_curtail(BfCtx &X, BfBranch &b, uint32_t seedLen) {
	if (b.rangesSz == 0) { b.exhausted = 1; b.curtailed = 1; return; }
	uint32_t lowest = 0xffff;
	for (uint32_t i = b.i0; i <= b.len; i++) {
		if (!bf_eliminated(X, b, i)) {
			const uint32_t cost = ((bf_rs(X, b, i)->eq >> 8) & 0x7f) | (((uint32_t)b.rdepth + i < seedLen) ? (1u << 14) : 0u);
			if (cost < lowest) lowest = cost;
		}
	}
	if (lowest > 0 && lowest != 0xffff) b.cost = (uint16_t)(b.cost + lowest);
	else if (lowest == 0xffff) b.exhausted = 1;
	b.curtailed = 1;
}
/* RangeState::pickEdit (range_source.h:321-431) */
BT_FN uint32_t bf_pick_edit(BfCtx &X, BfRS &r, uint32_t &rnd, uint32_t &top, uint32_t &bot, bool &last) {
	const uint32_t el = r.eq & 15u;
	const uint32_t num = 4 - (uint32_t)__builtin_popcount(el);
	if (num > 1) {
		last = false;
		uint32_t tot = 0;
		for (uint32_t c = 0; c < 4; c++) if (!((el >> c) & 1)) tot += r.bots[c] - r.tops[c];
		if (tot == 0) { X.flags |= BT_FLAG_FRAME_OVF; tot = 1; }        /* the reference would divide by zero here */
		uint32_t dart = bt_rand_next(rnd) % tot;
		for (uint32_t c = 0; c < 3; c++) {
			if (!((el >> c) & 1)) {
				const uint32_t w = r.bots[c] - r.tops[c];
				if (dart < w) { top = r.tops[c]; bot = r.bots[c]; r.eq |= (1u << c); return c; }
				dart -= w;
			}
		}
		if (!((el >> 3) & 1)) { top = r.tops[3]; bot = r.bots[3]; r.eq |= 8u; return 3; }
		return 0;                                                        /* chr stays 0: not reachable with consistent ranges */
	}
	last = true;
	const uint32_t c = !(el & 1) ? 0u : !(el & 2) ? 1u : !(el & 4) ? 2u : 3u;
	top = r.tops[c]; bot = r.bots[c];
	r.eq |= BF_RS_ELIM;
	return c;
}
/* Branch::splitBranch (range_source.h:640-759) */
BT_NOINLINE uint32_t bf_branch_split(BfCtx &X, BfSrc &s, uint32_t srcRef, uint32_t qlen, uint32_t seedLen) {
	BfBranch &b = *BF_BR(X, srcRef);
	uint32_t tied[3] = { 0, 0, 0 }; uint32_t numTied = 0, best = 0xffff, next = 0xffff, notElim = 0;
	for (uint32_t i = b.i0; i <= b.len; i++) {
		if (bf_eliminated(X, b, i)) continue;
		notElim++;
		const uint32_t cost = (((uint32_t)b.rdepth + i < seedLen) ? (1u << 14) : 0u) | ((bf_rs(X, b, i)->eq >> 8) & 0x7f);
		if (cost < best) { next = best; best = cost; numTied = 1; tied[0] = i; }
		else if (cost == best) {
			if (numTied < 3) tied[numTied++] = i;
			else { tied[0] = tied[1]; tied[1] = tied[2]; tied[2] = i; }
		} else if (cost < next) next = cost;
	}
	uint32_t r = 0;
	if (numTied > 1) r = bt_rand_next(s.rnd) % numTied;
	const uint32_t pos = tied[r];
	bool last = false; uint32_t top = 0, bot = 0;
	const uint32_t chr = bf_pick_edit(X, *bf_rs(X, b, pos), s.rnd, top, bot, last);
	const uint32_t depth = pos + b.rdepth;
	const uint32_t nd0 = depth < b.depth1 ? b.depth1 : b.depth0, nd1 = depth < b.depth2 ? b.depth2 : b.depth1,
	               nd2 = depth < b.depth3 ? b.depth3 : b.depth2, nd3 = b.depth3;
	const uint32_t nref = bf_branch_new(X, s, qlen, nd0, nd1, nd2, nd3, depth + 1, 0, b.cost, (uint32_t)b.ham + (best & ~0xc000u), top, bot, &b, BF_EDIT(depth, chr));
	if (!nref) return 0;
	if (notElim == 1 && last) b.exhausted = 1;
	else if (numTied == 1 && last) {
		if (best != next) { b.delayedCost = (uint16_t)(b.cost - best + next); b.delayedIncrease = 1; }
	}
	return nref;
}
/* Branch::installRanges (range_source.h:955-996) */
BT_FN void bf_install_ranges(BfRS &r, uint32_t c, uint32_t qAllow, uint32_t q) {
	uint32_t eq = BF_RS_ELIM | 15u | ((q & 0x7f) << 8);
	if (q <= qAllow) {
		for (uint32_t k = 0; k < 4; k++) if (c != k && r.bots[k] > r.tops[k]) { eq &= ~BF_RS_ELIM; eq &= ~(1u << k); }
	}
	r.eq = eq;
}

/* PathManager::curtail (range_source.h:1409-1420) */
BT_NOINLINE void bf_pm_curtail(BfCtx &X, BfSrc &s, uint32_t brRef, uint32_t seedLen) {
	BfBranch &b = *BF_BR(X, brRef);
	const uint32_t orig = b.cost;
	bf_branch_curtail(X, b, seedLen);
	if (b.exhausted) { bf_pm_pop(X, s); bf_branch_free(X, s, brRef); }
	else if (b.cost != orig) { const uint32_t p = bf_pm_pop(X, s); bf_pm_push(X, s, p); }
}
/* PathManager::splitAndPrep (range_source.h:1426-1474); false = backtrack budget or memory exhausted */
BT_NOINLINE bool bf_pm_split_and_prep(BfCtx &X, BfSrc &s, uint32_t qlen, uint32_t seedLen) {
	if (s.heapSz == 0) return true;
	if (s.cfg.useBtCnt && X.btCnt == 0) return false;
	uint32_t f = X.A[s.heapOff];
	while (BF_BR(X, f)->delayedIncrease) {
		BfBranch &fb = *BF_BR(X, f);
		bf_pm_pop(X, s);
		fb.cost = fb.delayedCost; fb.delayedIncrease = 0; fb.delayedCost = 0;
		bf_pm_push(X, s, f);
		f = X.A[s.heapOff];
	}
	if (BF_BR(X, f)->curtailed) {
		if (s.cfg.useBtCnt) { if (--X.btCnt == 0) return false; }
		X.s_bt++;
		const uint32_t nb = bf_branch_split(X, s, f, qlen, seedLen);
		if (!nb) return false;
		if (BF_BR(X, f)->exhausted) { bf_pm_pop(X, s); bf_branch_free(X, s, f); }
		bf_pm_push(X, s, nb);
		if (X.flags & BT_FLAG_STACK_OVF) return false;
	}
	if (s.heapSz) bf_branch_prep(*BF_BR(X, X.A[s.heapOff]));
	return true;
}

/* ---- EbwtRangeSource ---------------------------------------------------------------------------- */
/* curRange_.mms / refcs: the branch's edits as query offsets, then the seedling's (addPartialEdits, ebwt_search_backtrack.h:2372-2381) */
BT_NOINLINE void bf_set_range_edits(BfCtx &X, BfSrc &s, const BfBranch *br) {
	const uint32_t nb = br ? br->nedits : 0u, need = nb + (s.seedValid ? s.seedNmm : 0u);
	if (need > s.rEditsCap) {
		const uint32_t cap = need < 8 ? 8 : need;
		const uint32_t off = bf_alloc(X, cap);
		if (!off) { s.rNmm = 0; return; }
		s.rEdits = off; s.rEditsCap = cap;
	}
	for (uint32_t i = 0; i < nb; i++) { const uint32_t e = X.A[br->edits + i]; X.A[s.rEdits + i] = BF_EDIT(s.qlen - BF_EDIT_POS(e) - 1, BF_EDIT_CHR(e)); }
	if (s.seedValid) for (uint32_t i = 0; i < s.seedNmm; i++) X.A[s.rEdits + nb + i] = BF_EDIT(s.qlen - s.seedMms[i] - 1, s.seedRefc[i]);
	s.rNmm = need;
}
BT_FN bool bf_hh_check_top(const BfSrc &s, const BfBranch &b, uint32_t d) {       /* ebwt_search_backtrack.h:2420-2445 */
	if (d == s.depth5) { if (b.nedits == 0) return false; }
	else if (d == s.depth3) { if (b.nedits < s.cfg.hh) return false; }
	return true;
}
BT_FN bool bf_hh_check(const BfCtx &X, const BfSrc &s, const BfBranch &b, uint32_t depth, bool empty) {   /* ebwt_search_backtrack.h:2383-2414 */
	if (depth == s.depth5 - 1 && !empty) return b.nedits > 0;
	if (depth == s.depth3 - 1 && !empty) {
		uint32_t lo = 0, hi = 0;
		for (uint32_t i = 0; i < b.nedits; i++) { const uint32_t pos = BF_EDIT_POS(X.A[b.edits + i]); if (pos < s.depth5) hi++; else if (pos < s.depth3) lo++; }
		return b.nedits >= s.cfg.hh && !(lo == 0 || hi == 0);
	}
	return true;
}
/* EbwtRangeSource::initBranch (ebwt_search_backtrack.h:1907-2013) */
BT_NOINLINE void bf_init_branch(BfCtx &X, BfSrc &s) {
	const BtDevIndex &ix = X.P->ix[s.cfg.ebwtSel];
	const uint32_t ftabChars = (uint32_t)ix.ftabChars;
	s.rsFound = 0;
	if (s.skipping) { s.rsDone = 1; return; }
	if (s.qlen < 4) {
		uint32_t maxmms = 0;
		if (s.off0 != s.off1) maxmms = 1;
		if (s.off1 != s.off2) maxmms = 2;
		if (s.off2 != s.off3) maxmms = 3;
		if (s.qlen <= maxmms) { s.rsDone = 1; s.skipping = 1; return; }
	}
	/* tallyNs (2452-2484) */
	uint32_t nsInSeed = 0, nsInFtab = 0;
	for (uint32_t i = 0; i < s.off3; i++) {
		if (bf_qry(X, s, s.qlen - i - 1) == 4) {
			nsInSeed++;
			if (nsInSeed == 1) { if (i < s.off0) return; }
			else if (nsInSeed == 2) { if (i < s.off1) return; }
			else if (nsInSeed == 3) { if (i < s.off2) return; }
			else return;
		}
	}
	for (uint32_t i = 0; i < ftabChars && i < s.qlen; i++) if (bf_qry(X, s, s.qlen - i - 1) == 4) nsInFtab++;
	const uint32_t icost = s.seedValid ? s.seedCost : 0u;
	const uint32_t iham = s.seedValid ? (s.seedCost & ~0xc000u) : 0u;      /* qualOrder_ is always set on this path */
	const uint32_t m = s.off0 < s.qlen ? s.off0 : s.qlen;
	const bool skipInvalidExact = !s.cfg.reportExacts && s.qlen == ftabChars;
	if (nsInFtab == 0 && m >= ftabChars && !skipInvalidExact) {
		uint32_t off = bf_qry(X, s, s.qlen - ftabChars);                  /* calcFtabOff (2486-2495) */
		for (uint32_t i = ftabChars - 1; i > 0; i--) off = (off << 2) | bf_qry(X, s, s.qlen - i);
		const uint32_t top = bt_ftab_hi(ix, off), bot = bt_ftab_lo(ix, off + 1);
		X.s_ftab++;
		if (s.qlen == ftabChars && bot > top) {
			s.rTop = top; s.rBot = bot; s.rCost = (uint16_t)icost;
			bf_set_range_edits(X, s, 0);
			s.rsFound = 1;
		} else if (bot > top) {
			const uint32_t b = bf_branch_new(X, s, s.qlen, s.off0, s.off1, s.off2, s.off3, 0, ftabChars, icost, iham, top, bot, 0, 0);
			if (b) bf_pm_push(X, s, b);
		}
	} else {
		const uint32_t b = bf_branch_new(X, s, s.qlen, s.off0, s.off1, s.off2, s.off3, 0, 0, icost, iham, 0, 0, 0, 0);
		if (b) bf_pm_push(X, s, b);
	}
}

/* EbwtRangeSource::advanceBranch with until = ADV_COST_CHANGES (ebwt_search_backtrack.h:2060-2361) */
BT_NOINLINE void bf_advance_branch(BfCtx &X, BfSrc &s) {
	const BtDevIndex &ix = X.P->ix[s.cfg.ebwtSel];
	const uint32_t qualLim = X.P->prog.qualLim, maq = (uint32_t)X.P->pol.maqRound;
	s.rsFound = 0;
	do {
		if (++X.steps > BF_MAX_STEPS) { X.flags |= BT_FLAG_FRAME_OVF; bf_pm_reset(s); return; }
		const uint32_t brRef = X.A[s.heapOff];
		BfBranch &br = *BF_BR(X, brRef);
		const uint32_t depth = (uint32_t)br.rdepth + br.len;
		const uint32_t cost = br.cost;
		uint32_t cur = 0;
		bool bail = false;
		if (s.cfg.hh && !bf_hh_check_top(s, br, depth)) { bf_pm_curtail(X, s, brRef, s.depth3); bail = true; }
		if (!bail) {
			cur = s.qlen - depth - 1;
			if (depth < s.qlen) {
				const uint32_t c = bf_qry(X, s, cur);
				const uint32_t q = bt_mm_penalty(maq, bf_phred(bf_qualch(X, s, cur)));
				const bool curIsAlt = depth >= br.depth0 && (uint32_t)br.ham + q <= qualLim;
				uint32_t otop = br.top;
				if (c == 4 && depth > 0) br.top = br.bot = 1;
				BfRS dummy; dummy.eq = 0;
				BfRS *rs = (br.len >= br.i0 && br.len < br.rangesSz) ? bf_rs(X, br, br.len) : &dummy;
				if (br.top == 0 && br.bot == 0) {
					for (uint32_t k = 0; k < 4; k++) { rs->tops[k] = ix.fchr[k]; rs->bots[k] = ix.fchr[k + 1]; }
					bf_install_ranges(*rs, c, qualLim - br.ham, q);
					if (c < 4) { br.top = rs->tops[c]; br.bot = rs->bots[c]; }
				} else if (curIsAlt && (br.bot > br.top || c == 4)) {
					for (uint32_t k = 0; k < 4; k++) rs->tops[k] = rs->bots[k] = 0;
					if (br.lbotValid) {
						BtBlock bA = bt_load_block(ix, br.ltop), bB = bA;
						if ((br.lbot >> 6) != (br.ltop >> 6)) bB = bt_load_block(ix, br.lbot);
						bt_lf_ex(ix, bA, br.ltop, rs->tops); bt_lf_ex(ix, bB, br.lbot, rs->bots);
						X.s_lfex++;
					} else {
						/* mapLF1(otop, ltop): follow the single row whatever its character is (ebwt.h:2530-2560) */
						const BtBlock bA = bt_load_block(ix, br.ltop);
						X.s_lf++;
						int cc = -1;
						if (otop != ix.zOff) { cc = (int)bt_row_l(bA, br.ltop); otop = bt_lf(ix, bA, br.ltop, (uint32_t)cc); }
						br.top = otop;
						if (cc >= 0) { rs->tops[cc] = br.top; rs->bots[cc] = br.top + 1; }
					}
					bf_install_ranges(*rs, c, qualLim - br.ham, q);
					if (c < 4) { br.top = rs->tops[c]; br.bot = rs->bots[c]; }
					else br.top = br.bot = 1;
				} else if (br.bot > br.top) {
					rs->eq |= BF_RS_ELIM;
					if (c < 4) {
						const BtBlock bA = bt_load_block(ix, br.ltop);
						if (br.top + 1 == br.bot) {
							uint32_t t;                                       /* mapLF1(top, ltop, c) (ebwt.h:2494-2524) */
							if (bt_row_l(bA, br.ltop) != c || br.top == ix.zOff) t = BT_OFF_MASK;
							else t = bt_lf(ix, bA, br.ltop, c);
							br.top = br.bot = t;
							if (t != BT_OFF_MASK) br.bot++;
							X.s_lf++;
						} else {
							BtBlock bB = bA;
							if ((br.lbot >> 6) != (br.ltop >> 6)) bB = bt_load_block(ix, br.lbot);
							br.top = bt_lf(ix, bA, br.ltop, c);
							br.bot = bt_lf(ix, bB, br.lbot, c);
							X.s_lf += 2;
						}
					}
				} else rs->eq |= BF_RS_ELIM;
			} else cur = 0;
			const bool empty = br.top == br.bot;
			const bool hit = cur == 0 && !empty;
			const uint32_t nedits = br.nedits;
			const bool invalidExact = hit && nedits == 0 && !s.cfg.reportExacts;
			if (s.cfg.hh && !bf_hh_check(X, s, br, depth, empty)) bf_pm_curtail(X, s, brRef, s.depth3);
			else if (hit && !invalidExact) {
				s.rTop = br.top; s.rBot = br.bot; s.rCost = br.cost;
				bf_set_range_edits(X, s, &br);
				s.rsFound = 1;
				bf_pm_curtail(X, s, brRef, s.depth3);
			} else if (empty || cur == 0) bf_pm_curtail(X, s, brRef, s.depth3);
			else br.len++;                                                /* Branch::extend */
		}
		if (!bf_pm_split_and_prep(X, s, s.qlen, s.depth3)) bf_pm_reset(s);
		if (X.flags & (BT_FLAG_STACK_OVF | BT_FLAG_FRAME_OVF)) { bf_pm_reset(s); return; }
		if (s.heapSz == 0) break;
		if (BF_BR(X, X.A[s.heapOff])->cost != cost) break;
	} while (!s.rsFound);
}

/* ---- EbwtRangeSourceDriver / SingleRangeSourceDriver ------------------------------------------- */
BT_FN uint32_t bf_cext(uint32_t cext, uint32_t sRight, uint32_t s, uint32_t len) {
	return cext == BF_PIN_SEED ? s : cext == BF_PIN_HI_HALF ? sRight : cext == BF_PIN_BEGINNING ? 0u : len;
}
/* SingleRangeSourceDriver::setQueryImpl (range_source.h:1725-1747) with EbwtRangeSource::setQuery (1831-1866) and
 * EbwtRangeSourceDriver::initRangeSource (2724-2797).  `seed` = the seedling (a BfSrc whose current range it is) or NULL. */
BT_NOINLINE void bf_src_set_query(BfCtx &X, BfSrc &s, const BfSrc *seed) {
	const uint32_t len = X.rlenM[s.cfg.mate], maq = (uint32_t)X.P->pol.maqRound;
	s.h.done = 0;
	bf_pm_reset(s);
	s.heapOff = 0; s.heapCap = 0;
	const bool ebwtFw = s.cfg.ebwtSel == 0;
	s.viewRev = ebwtFw ? !s.cfg.fw : s.cfg.fw; s.viewComp = !s.cfg.fw;
	s.seedValid = 0; s.seedNmm = 0; s.seedCost = 0;
	s.qlen = len; s.skipping = 0; s.rsDone = 0; s.rsFound = 0; s.rnd = X.seedM[s.cfg.mate];
	if (seed) {
		s.seedValid = 1; s.seedCost = seed->rCost;
		uint32_t n = seed->rNmm; if (n > 3) { n = 3; X.flags |= BT_FLAG_MM_OVF; }
		s.seedNmm = (uint8_t)n;
		for (uint32_t i = 0; i < n; i++) {
			const uint32_t e = X.A[seed->rEdits + i];
			s.seedMms[i] = (uint16_t)BF_EDIT_POS(e); s.seedRefc[i] = (uint8_t)BF_EDIT_CHR(e);
			s.ovPos[i] = (uint16_t)(len - BF_EDIT_POS(e) - 1);
		}
	}
	/* initRangeSource */
	const uint32_t seedLen = X.P->prog.seedLen;
	const uint32_t sl = seedLen > 0 ? (seedLen < len ? seedLen : len) : len;
	uint32_t sRight = sl >> 1;                                          /* the odd base goes left (nudgeLeft) or right */
	if ((sl & 1) && !s.cfg.nudgeLeft) sRight++;
	const uint32_t r0 = bf_cext(s.cfg.rev[0], sRight, sl, len), r1 = bf_cext(s.cfg.rev[1], sRight, sl, len),
	               r2 = bf_cext(s.cfg.rev[2], sRight, sl, len), r3 = bf_cext(s.cfg.rev[3], sRight, sl, len);
	uint32_t qlen = len;
	if (s.cfg.seed && len > sl) { s.qlen = sl; qlen = sl; }
	uint32_t minCost = 0;
	if (s.cfg.reportExacts) { }
	else if (!s.cfg.hh && r0 < sl) {
		minCost = 1u << 14;
		uint32_t low = 0xff;
		for (uint32_t d = r0; d < sl; d++) { const uint32_t ch = bf_qualch(X, s, qlen - d - 1); if (ch < low) low = ch; }
		minCost += bt_mm_penalty(maq, bf_phred(low));
	} else if (s.cfg.hh && sRight > 0 && sRight < sl - 1) {
		minCost = (s.cfg.seed ? 3u : 2u) << 14;
		uint32_t low1 = 0xff;
		for (uint32_t d = 0; d < sRight; d++) { const uint32_t ch = bf_qualch(X, s, qlen - d - 1); if (ch < low1) low1 = ch; }
		minCost += bt_mm_penalty(maq, bf_phred(low1));
		uint32_t l21 = 0xff, l22 = 0xff;
		for (uint32_t d = sRight; d < sl; d++) {
			const uint32_t ch = bf_qualch(X, s, qlen - d - 1);
			if (ch < l21) { if (l21 != 0xff) l22 = l21; l21 = ch; }
			else if (ch < l22) l22 = ch;
		}
		minCost += bt_mm_penalty(maq, bf_phred(l21));
		if (s.cfg.hh > 2 && l22 != 0xff) minCost += bt_mm_penalty(maq, bf_phred(l22));
	}
	s.h.minCostAdj = (uint16_t)minCost;
	s.depth5 = sRight; s.depth3 = sl; s.off0 = r0; s.off1 = r1; s.off2 = r2; s.off3 = r3;
	bf_init_branch(X, s);
	const uint32_t icost = seed ? seed->rCost : 0u;
	s.h.minCost = (uint16_t)(icost > s.h.minCostAdj ? icost : s.h.minCostAdj);
	s.h.done = s.rsDone; s.h.foundRange = s.rsFound;
}
/* SingleRangeSourceDriver::advanceImpl (range_source.h:1753-1802) */
BT_FN void bf_src_advance(BfCtx &X, BfSrc &s) {
	if (s.h.done || s.heapSz == 0) { s.h.done = 1; return; }
	bf_advance_branch(X, s);
	s.h.done = s.heapSz == 0;
	if (s.pmMinCost != 0) s.h.minCost = s.pmMinCost > s.h.minCostAdj ? s.pmMinCost : s.h.minCostAdj;
	s.h.foundRange = s.rsFound;
}
BT_NOINLINE uint32_t bf_src_new(BfCtx &X, const BfSrcCfg &cfg) {
	const uint32_t ref = bf_alloc_zero(X, BF_SRC_WORDS);
	if (!ref) return 0;
	BfSrc &s = *BF_AT(BfSrc, X, ref);
	s.h.kind = BF_KIND_SRC; s.h.done = 1; s.h.fw = cfg.fw; s.h.mate = cfg.mate; s.cfg = cfg;
	return ref;
}

/* ---- CostAwareRangeSourceDriver (range_source.h:2023-2461) and EbwtSeededRangeSourceDriver ------- */
BT_NOINLINE void bf_node_advance(BfCtx &X, uint32_t node);
BT_NOINLINE void bf_node_set_query(BfCtx &X, uint32_t node);
BT_FN BfHdr &bf_hdr(BfCtx &X, uint32_t node) { return *BF_AT(BfHdr, X, node); }
BT_FN uint32_t bf_node_range(BfCtx &X, uint32_t node) {                  /* &p->range() : the BfSrc that owns the Range */
	if (bf_hdr(X, node).kind == BF_KIND_SRC) return node;
	return BF_AT(BfSeeded, X, node)->full.lastRange;
}
/* sortActives (range_source.h:2382-2424): a selection sort whose ties are broken by the driver's RNG */
BT_NOINLINE void bf_ca_sort(BfCtx &X, BfCA &ca) {
	uint32_t *v = X.A + ca.actOff;
	uint32_t sz = ca.nAct;
	for (uint32_t i = 0; i < sz;) {
		if (bf_hdr(X, v[i]).done && !bf_hdr(X, v[i]).foundRange) {
			for (uint32_t k = i + 1; k < ca.nAct; k++) v[k - 1] = v[k];
			ca.nAct--;
			if (sz == 0) break; else sz--;
			continue;
		}
		uint32_t minCost = bf_hdr(X, v[i]).minCost, minOff = i;
		for (uint32_t j = i + 1; j < sz; j++) {
			const BfHdr &hj = bf_hdr(X, v[j]);
			if (hj.done && !hj.foundRange) continue;
			if (hj.minCost < minCost) { minCost = hj.minCost; minOff = j; }
			else if (hj.minCost == minCost) { if (bt_rand_next(ca.rnd) & 0x1000) minOff = j; }
		}
		if (i != minOff) { const uint32_t t = v[i]; v[i] = v[minOff]; v[minOff] = t; }
		i++;
	}
	if (ca.delayedRange == 0 && sz > 0) ca.minCost = bf_hdr(X, v[0]).minCost;
}
/* mateEliminated (range_source.h:2302-2315): in a list that mixes both mates, no live driver is left for one of them */
BT_FN bool bf_ca_mate_eliminated(BfCtx &X, const BfCA &ca) {
	if (!ca.paired) return false;
	bool left[2] = { false, false };
	for (uint32_t i = 0; i < ca.nAct; i++) { const BfHdr &h = bf_hdr(X, X.A[ca.actOff + i]); if (!h.done) left[h.mate & 1] = true; }
	return !left[0] || !left[1];
}
/* foundFirstRange (range_source.h:2339-2377); strandFix is only ever set on the top-level driver */
BT_NOINLINE bool bf_ca_found_first(BfCtx &X, BfCA &ca, uint32_t r, bool strandFix) {
	ca.foundRange = 1;
	ca.lastRange = r;
	if (strandFix) {
		const uint32_t rfw = BF_AT(BfSrc, X, r)->cfg.fw, rmate = BF_AT(BfSrc, X, r)->cfg.mate;
		const uint32_t *rss = X.A + ca.rssOff; const uint32_t *act = X.A + ca.actOff;
		for (uint32_t i = 1; i < ca.nAct; i++) {
			if (bf_hdr(X, rss[i]).mate == rmate && bf_hdr(X, rss[i]).fw != rfw) {   /* sic: tests rss_[i], then uses active_[i] */
				const uint32_t p = act[i];
				const uint32_t minCost = ca.minCost > bf_hdr(X, p).minCost ? ca.minCost : bf_hdr(X, p).minCost;
				if (minCost > BF_AT(BfSrc, X, r)->rCost) break;
				while (!bf_hdr(X, p).done && !bf_hdr(X, p).foundRange) {
					bf_node_advance(X, p);
					if (X.flags & (BT_FLAG_STACK_OVF | BT_FLAG_FRAME_OVF)) return true;
					if (bf_hdr(X, p).minCost > minCost) break;
				}
				if (bf_hdr(X, p).foundRange) {
					ca.delayedRange = bf_node_range(X, p);
					const BfSrc &d = *BF_AT(BfSrc, X, ca.delayedRange), &l = *BF_AT(BfSrc, X, ca.lastRange);
					const uint32_t tot = (d.rBot - d.rTop) + (l.rBot - l.rTop);
					const uint32_t rq = bt_rand_next(ca.rnd) % tot;
					if (rq < d.rBot - d.rTop) { const uint32_t t = ca.lastRange; ca.lastRange = ca.delayedRange; ca.delayedRange = t; }
					bf_hdr(X, p).foundRange = 0;
				}
				return true;
			}
		}
	}
	return false;
}
/* CostAwareRangeSourceDriver::advanceImpl (range_source.h:2131-2180).  TOP: the aligner's driver list (plain and seeded
 * drivers, strand fix); otherwise the per-seedling list inside a seeded driver, whose elements are always plain
 * drivers — spelled out so that the device call graph has no cycle. */
template <bool TOP>
BT_NOINLINE void bf_ca_advance(BfCtx &X, BfCA &ca, bool strandFix) {
	ca.lastRange = 0;
	if (ca.delayedRange != 0) {
		ca.lastRange = ca.delayedRange; ca.delayedRange = 0; ca.foundRange = 1;
		if (ca.nAct) { const uint32_t c0 = bf_hdr(X, X.A[ca.actOff]).minCost; if (c0 > ca.minCost) ca.minCost = (uint16_t)c0; }
		else ca.done = 1;
		return;
	}
	if ((TOP && bf_ca_mate_eliminated(X, ca)) || ca.nAct == 0) { ca.nAct = 0; ca.done = 1; return; }
	const uint32_t p = X.A[ca.actOff];
	const uint32_t precost = bf_hdr(X, p).minCost;
	if (!bf_hdr(X, p).foundRange) { if (TOP) bf_node_advance(X, p); else bf_src_advance(X, *BF_AT(BfSrc, X, p)); }
	if (X.flags & (BT_FLAG_STACK_OVF | BT_FLAG_FRAME_OVF)) return;
	bool needsSort = false;
	if (bf_hdr(X, p).foundRange) {
		if (TOP) needsSort = bf_ca_found_first(X, ca, bf_node_range(X, p), strandFix);
		else { ca.foundRange = 1; ca.lastRange = p; }
		bf_hdr(X, p).foundRange = 0;
	}
	if (bf_hdr(X, p).done || precost != bf_hdr(X, p).minCost || needsSort) {
		bf_ca_sort(X, ca);
		if ((TOP && bf_ca_mate_eliminated(X, ca)) || ca.nAct == 0) { ca.nAct = 0; ca.done = (ca.delayedRange == 0); }
	}
}
BT_FN void bf_ca_copy_active(BfCtx &X, BfCA &ca) {                       /* active_ = rss_ */
	if (ca.actCap < ca.nRss) { bf_free_top(X, ca.actOff, ca.actCap); ca.actOff = bf_alloc(X, ca.rssCap); ca.actCap = ca.actOff ? ca.rssCap : 0; }
	ca.nAct = 0;
	if (ca.actCap >= ca.nRss) { for (uint32_t i = 0; i < ca.nRss; i++) X.A[ca.actOff + i] = X.A[ca.rssOff + i]; ca.nAct = ca.nRss; }
}
/* CostAwareRangeSourceDriver::setQueryImpl (range_source.h:2072-2088) */
BT_FN void bf_ca_set_query_empty(BfCtx &X, BfCA &ca) {                    /* rss_ is empty: the seeded driver's list after clearSources */
	ca.done = 0; ca.foundRange = 0; ca.lastRange = 0; ca.delayedRange = 0;
	ca.rnd = X.seedM[0];                                                  /* rand_.init(patsrc->bufa().seed) */
}
BT_NOINLINE void bf_ca_set_query(BfCtx &X, BfCA &ca) {
	bf_ca_set_query_empty(X, ca);
	if (ca.nRss == 0) return;
	for (uint32_t i = 0; i < ca.nRss; i++) { bf_node_set_query(X, X.A[ca.rssOff + i]); if (X.flags & BT_FLAG_STACK_OVF) return; }
	bf_ca_copy_active(X, ca);
	ca.minCost = 0;
	bf_ca_sort(X, ca);
}
/* EbwtSeededRangeSourceDriver::setQueryImpl (ebwt_search_backtrack.h:2965-2977) */
BT_NOINLINE void bf_seeded_set_query(BfCtx &X, BfSeeded &sd) {
	sd.h.done = 0;
	BfSrc &gen = *BF_AT(BfSrc, X, sd.seedgen);
	bf_src_set_query(X, gen, 0);
	sd.h.minCostAdj = gen.h.minCostAdj > gen.h.minCost ? gen.h.minCostAdj : gen.h.minCost;
	sd.h.minCost = sd.h.minCostAdj;
	sd.full.nRss = 0; sd.full.nAct = 0;                                   /* clearSources */
	bf_ca_set_query_empty(X, sd.full);
	sd.full.minCost = sd.h.minCost;
	sd.h.foundRange = 0;
}
/* EbwtSeededRangeSourceDriver::advanceImpl (ebwt_search_backtrack.h:3003-3090) */
BT_NOINLINE void bf_seeded_advance(BfCtx &X, uint32_t node) {
	BfSeeded &sd = *BF_AT(BfSeeded, X, node);
	BfSrc &gen = *BF_AT(BfSrc, X, sd.seedgen);
	BfCA &full = sd.full;
	if (gen.h.done && full.done && !gen.h.foundRange && !full.foundRange) { sd.h.done = 1; return; }
	if (gen.h.done && !gen.h.foundRange) {
		gen.h.minCost = 0xffff;
		if (full.minCost > sd.h.minCost) { sd.h.minCost = full.minCost; return; }
	}
	if (full.done && !full.foundRange) {
		full.minCost = 0xffff;
		if (gen.h.minCost > sd.h.minCost) { sd.h.minCost = gen.h.minCost; return; }
	}
	const bool doFull = full.minCost <= gen.h.minCost;
	if (!doFull) {
		if (!gen.h.foundRange) bf_src_advance(X, gen);
		if (X.flags & (BT_FLAG_STACK_OVF | BT_FLAG_FRAME_OVF)) return;
		if (gen.h.foundRange) {
			gen.h.foundRange = 0;
			sd.h.minCostAdj = gen.rCost;
			const uint32_t pref = bf_src_new(X, sd.fact);                  /* rsFact_->create() */
			if (!pref) return;
			/* re-derive references: the arena does not move, but keep the code honest about aliasing */
			BfSrc &partial = *BF_AT(BfSrc, X, pref);
			partial.h.minCost = gen.rCost;
			full.minCost = gen.rCost;
			/* addSource (range_source.h:2094-2106) */
			full.lastRange = 0; full.delayedRange = 0; full.done = 0;
			bf_src_set_query(X, partial, &gen);
			if (X.flags & BT_FLAG_STACK_OVF) return;
			bf_vec_push(X, full.rssOff, full.rssCap, full.nRss, pref);
			bf_vec_push(X, full.actOff, full.actCap, full.nAct, pref);
			if (X.flags & BT_FLAG_STACK_OVF) return;
			full.minCost = 0;
			bf_ca_sort(X, full);
			if (full.foundRange) { sd.h.foundRange = 1; full.foundRange = 0; }
		}
		if (gen.h.minCost > sd.h.minCost) {
			sd.h.minCost = gen.h.minCost;
			if (!full.done) sd.h.minCost = sd.h.minCost < full.minCost ? sd.h.minCost : full.minCost;
		}
	} else {
		const uint32_t oldFull = full.minCost;
		if (!full.foundRange) bf_ca_advance<false>(X, full, false);
		if (full.foundRange) { sd.h.foundRange = 1; full.foundRange = 0; }
		if (full.minCost > oldFull) sd.h.minCost = full.minCost < gen.h.minCost ? full.minCost : gen.h.minCost;
	}
}
BT_NOINLINE void bf_node_advance(BfCtx &X, uint32_t node) {
	if (bf_hdr(X, node).kind == BF_KIND_SRC) bf_src_advance(X, *BF_AT(BfSrc, X, node));
	else bf_seeded_advance(X, node);
}
BT_NOINLINE void bf_node_set_query(BfCtx &X, uint32_t node) {
	if (bf_hdr(X, node).kind == BF_KIND_SRC) bf_src_set_query(X, *BF_AT(BfSrc, X, node), 0);
	else bf_seeded_set_query(X, *BF_AT(BfSeeded, X, node));
}

/* ---- sink + hit construction --------------------------------------------------------------------- */
/* HitSinkPerThread::irrelevantCost: only the stratified sink ever says yes (hit.h:1113-1118) */
BT_FN bool bf_irrelevant_cost(const BfCtx &X, uint32_t cost) {
	if (!X.P->pol.strata) return false;
	if (X.found) return (int32_t)(cost >> 14) > X.bestStratum;
	return false;
}
/* A Range as the aligners hand it to EbwtSearchParams::reportHit: found by a range source (index) or by the reference scan */
struct BfRangeView { uint32_t top, bot, cost, nmm, edits; uint8_t fw, ebwtFw, mate, pad; };
BT_FN BfRangeView bf_view_of(const BfSrc &s) {
	BfRangeView v; v.top = s.rTop; v.bot = s.rBot; v.cost = s.rCost; v.nmm = s.rNmm; v.edits = s.rEdits;
	v.fw = s.cfg.fw; v.ebwtFw = s.cfg.ebwtSel == 0; v.mate = s.cfg.mate; v.pad = 0;
	return v;
}
/* EbwtSearchParams::reportHit (ebwt.h:1288-1405) → the per-thread sink's reportHit (NGood hit.h:969-985, NBestFirstStrat
 * 1070-1094, AllHit 1201-1209).  `mult` = 2 for paired sinks (createMult), `mateNo` = Hit::mate (0 unpaired, 1, 2).
 * Returns true when the read (pair) is finished. */
BT_NOINLINE bool bf_report_hit(BfCtx &X, const BfRangeView &ra, uint32_t tidx, uint32_t toff, uint32_t oms, uint32_t mateNo, uint32_t mult) {
	const BfKParams &P = *X.P;
	const BtPolicy &pol = P.pol;
	uint32_t n = pol.allHits ? (pol.strata ? 0x7fffffffu : 0xffffffffu) : pol.khits;
	uint32_t mx = pol.mhits;
	if (mult > 1) { if (n != 0xffffffffu) n *= mult; if (mx != 0xffffffffu) mx *= mult; }
	const uint32_t stratum = ra.cost >> 14;
	X.found++;
	if ((int32_t)stratum < X.bestStratum) X.bestStratum = (int32_t)stratum;
	if (X.found > mx) return true;
	const uint32_t keep = (pol.sampleMax && mx != 0xffffffffu && mx > n) ? mx : n;   /* bufferHit precedes the n test */
	if (X.found <= keep) {
		if (X.found <= P.slots) {
			uint32_t *rec = P.hits + ((size_t)X.rid * P.slots + (X.found - 1)) * P.rec_words;
			rec[0] = tidx; rec[1] = toff; rec[2] = oms;
			rec[3] = (ra.cost & 0xffffu) | (stratum << 16) | ((uint32_t)ra.fw << 24) | (mateNo << 25);
			rec[4] = ra.nmm;
			const bool flip = (ra.ebwtFw != 0) != (ra.fw != 0);                 /* ebwt.h:1339-1350 */
			const uint32_t qlen = X.rlenM[ra.mate];
			for (uint32_t i = 0; i < ra.nmm; i++) {
				const uint32_t e = X.A[ra.edits + i];
				uint32_t pos = BF_EDIT_POS(e);
				if (flip) pos = qlen - pos - 1;
				if (i < P.mm_cap) rec[BT_HIT_HDR + i] = BF_EDIT(pos, BF_EDIT_CHR(e)); else X.flags |= BT_FLAG_MM_OVF;
			}
		} else X.flags |= BT_FLAG_HITS_OVF;
	}
	if (!(pol.allHits && !pol.strata) && X.found == n && (mx == 0xffffffffu || mx < n)) return true;
	return false;
}
/* UnpairedAlignerV2::report (aligner.h:462-492) */
BT_FN bool bf_report(BfCtx &X, const BfSrc &ra, uint32_t tidx, uint32_t toff) {
	return bf_report_hit(X, bf_view_of(ra), tidx, toff, ra.rBot - ra.rTop - 1, 0, 1);
}

/* RowChaser (row_chaser.h:60-110): resolve one BW row to a joined-text offset */
BT_FN uint32_t bf_resolve_row(BfCtx &X, const BtDevIndex &ix, uint32_t row) {
	uint32_t jumps = 0;
	while (row != ix.zOff && (row & ix.offMask) != row) {
		const BtBlock b = bt_load_block(ix, row);
		row = bt_lf(ix, b, row, bt_row_l(b, row));
		jumps++; X.s_lf++; X.s_chase++;
	}
	if (row == ix.zOff) return jumps;
	X.s_offs++;
	return BT_LDG(ix.offs + (row >> ix.offRate)) + jumps;
}

/* ---- UnpairedAlignerV2 (aligner.h:420-560) ------------------------------------------------------ */
BT_NOINLINE void bf_align_read(BfCtx &X) {
	const BfKParams &P = *X.P;
	X.randA = X.seedM[0];                                                 /* Aligner::setQuery: rand_.init(seed) */
	X.found = 0; X.bestStratum = 999; X.btCnt = (int32_t)P.pol.maxBtsBest;
	if (X.rlenM[0] < 4) return;                                           /* "Skipping read ... less than 4 characters long" */
	/* build this read's driver tree (the reference builds it once per thread and re-targets it with setQuery) */
	BfCA &top = X.top;
	top.rssOff = bf_alloc(X, BF_MAX_TOP); top.rssCap = BF_MAX_TOP; top.nRss = 0;
	top.actOff = bf_alloc(X, BF_MAX_TOP); top.actCap = BF_MAX_TOP; top.nAct = 0;
	top.minCost = 0;
	for (uint32_t i = 0; i < P.prog.ntop; i++) {
		const BfTopCfg &tc = P.prog.top[i];
		uint32_t node;
		if (tc.kind == BF_KIND_SRC) node = bf_src_new(X, tc.a);
		else {
			node = bf_alloc_zero(X, BF_SEEDED_WORDS);
			if (node) {
				BfSeeded &sd = *BF_AT(BfSeeded, X, node);
				sd.h.kind = BF_KIND_SEEDED; sd.h.done = 1; sd.h.fw = tc.a.fw; sd.h.mate = tc.a.mate; sd.fact = tc.b;
				sd.seedgen = bf_src_new(X, tc.a);
			}
		}
		if (X.flags & BT_FLAG_STACK_OVF) return;
		X.A[top.rssOff + top.nRss++] = node;
	}
	bf_ca_set_query(X, top);
	bool done = top.done, chase = false;
	/* range chaser state (range_chaser.h) */
	uint32_t cTop = 0, cBot = 0, cIrow = 0, cRow = 0; bool cDone = false, cPending = false;
	const bool strandFix = P.prog.strandFix != 0;
	while (!done) {
		if (X.flags & (BT_FLAG_STACK_OVF | BT_FLAG_FRAME_OVF)) return;
		if (chase) {
			/* RangeChaser::advance until an offset is found or the range is exhausted */
			const BfSrc &ra = *BF_AT(BfSrc, X, top.lastRange);
			const BtDevIndex &ix = P.ix[ra.cfg.ebwtSel];
			bool foundOff = false; uint32_t tidx = 0, toff = 0;
			while (!foundOff && !cDone) {
				if (!cPending) {
					cRow++; if (cRow == cBot) cRow = cTop;
					if (cRow == cIrow) { cDone = true; break; }
				}
				cPending = false;
				const uint32_t off = bf_resolve_row(X, ix, cRow);
				foundOff = bt_joined_to_text(ix, X.rlenM[0], off, tidx, toff);
			}
			if (foundOff) done = bf_report(X, ra, tidx, toff);
			else { chase = false; top.foundRange = 0; done = top.done; }
		}
		if (!done && !chase) {
			if (top.foundRange) {
				const BfSrc &ra = *BF_AT(BfSrc, X, top.lastRange);
				const BtDevIndex &ix = P.ix[ra.cfg.ebwtSel];
				/* setTopBot + setRow: rows that sit on a sampled position (or on zOff) resolve at once */
				cTop = ra.rTop; cBot = ra.rBot;
				cIrow = cTop + bt_rand_next(X.randA) % (cBot - cTop);
				cDone = false; cPending = false; cRow = cIrow;
				bool foundOff = false; uint32_t tidx = 0, toff = 0;
				for (;;) {
					if (cRow != ix.zOff && (cRow & ix.offMask) != cRow) { cPending = true; break; }
					const uint32_t off = bf_resolve_row(X, ix, cRow);
					if (bt_joined_to_text(ix, X.rlenM[0], off, tidx, toff)) { foundOff = true; break; }
					cRow++; if (cRow == cBot) cRow = cTop;
					if (cRow == cIrow) { cDone = true; break; }
				}
				if (foundOff) done = bf_report(X, ra, tidx, toff);
				if (!cDone && !bf_irrelevant_cost(X, ra.rCost)) chase = true;
				else top.foundRange = 0;
			} else {
				done = bf_irrelevant_cost(X, top.minCost);
				if (!done) bf_ca_advance<true>(X, top, strandFix);
			}
			if (top.done && !top.foundRange && !chase) done = true;
		}
	}
}

/* ================================================================================================= */
/* Paired-end: PairedBWAlignerV1 (aligner.h:606-1468) with dontReconcileMates (the default,          */
/* ebwt_search.cpp:219): every resolved offset of one mate anchors a scan of the reference window in  */
/* which the opposite mate may lie (RefAligner family, ref_aligner.h; BitPairReference, reference.h). */
/* ================================================================================================= */

/* BitPairReference::getStretch (reference.h:455-640; the naive form 417-453 defines it): `count` characters of reference
 * `tidx` from `toff`, 4 for ambiguous stretches and past the end.  Staged as bytes in the arena; returns the word offset. */
BT_NOINLINE uint32_t bf_ref_stretch(BfCtx &X, uint32_t tidx, uint32_t toff, uint32_t count) {
	const BtDevRef &R = X.P->ref;
	const uint32_t words = (count + 3) / 4 + 1;
	const uint32_t aoff = bf_alloc(X, words);
	if (!aoff) return 0;
	uint8_t *dest = (uint8_t *)(X.A + aoff);
	uint32_t cur = 0, off = 0;
	const uint32_t reci = BT_LDG(R.refRecOffs + tidx), recf = BT_LDG(R.refRecOffs + tidx + 1);
	uint32_t bufOff = BT_LDG(R.refOffs + tidx);
	for (uint32_t i = reci; i < recf && count > 0; i++) {
		const uint32_t roff = BT_LDG(R.recs + 2 * (size_t)i), rlen = BT_LDG(R.recs + 2 * (size_t)i + 1);
		off += roff;
		for (; toff < off && count > 0; toff++) { dest[cur++] = 4; count--; }
		if (count == 0) break;
		if (toff < off + rlen) bufOff += toff - off; else bufOff += rlen;
		off += rlen;
		for (; toff < off && count > 0; toff++) {
			dest[cur++] = (uint8_t)((BT_LDG(R.buf + (bufOff >> 2)) >> ((bufOff & 3) << 1)) & 3);
			bufOff++; count--;
		}
	}
	while (count > 0) { dest[cur++] = 4; count--; }
	return aoff;
}

struct BfPairSet { uint32_t off, cap, n; };       /* TSetPairs pairs_fw_ / pairs_rc_: (tidx, lo, hi) triples */

/* RefAligner::find(1, ...) (ref_aligner.h:63-97).  The eight concrete aligners (Exact/OneMM/TwoMM/ThreeMM for -v,
 * Seed0..3 for -n) are hand-unrolled 64-bit-anchor scans of one definition — their naiveFind members, which the
 * reference's debug build checks them against: candidates in zig-zag order from the middle of the window, a reference N
 * anywhere kills the candidate, a query N is a mismatch, at most `refMms` mismatches in the seed (the whole read for -v),
 * quality-weighted distance <= qualMax, first survivor that is not already in `pairs` wins.
 * qry = the mate in the orientation it must have on the forward reference strand; the seed is at its 5' end, which is the
 * left end iff `seedOnLeft` (= fw).  Returns true and fills `out` / `result` (leftmost reference offset). */
BT_NOINLINE bool bf_ref_find(BfCtx &X, uint32_t mate, bool fw, uint32_t tidx, uint32_t begin, uint32_t end, uint32_t aoff,
                             BfPairSet &pairs, BfRangeView &out, uint32_t &result) {
	const BfProg &g = X.P->prog;
	const uint32_t qlen = X.rlenM[mate], maq = (uint32_t)X.P->pol.maqRound;
	const uint8_t *seq = X.seqM[mate], *qual = X.qualM[mate];
	const bool seedOnLeft = fw;
	const uint32_t slen = g.refSeedLen ? (qlen < g.refSeedLen ? qlen : g.refSeedLen) : qlen;
	const uint32_t spread = end - begin;
	const uint32_t refOff = bf_ref_stretch(X, tidx, begin, spread);
	if (!refOff) return false;
	const uint8_t *ref = (const uint8_t *)(X.A + refOff);
	const uint32_t editsOff = bf_alloc(X, qlen ? qlen : 1);                    /* worst case: every position mismatches */
	if (!editsOff) return false;
	const uint32_t lim = end - qlen - begin;
	const uint32_t halfway = begin + (lim >> 1);
	bool hi = false, found = false;
	for (uint32_t i = 1; i <= lim + 1 && !found; i++) {
		const uint32_t L = hi ? halfway + (i >> 1) : halfway - (i >> 1);
		hi = !hi;
		const uint32_t rir = L - begin;
		bool match = true;
		uint32_t mms = 0, seedMms = 0, ham = 0;
		for (uint32_t jj = 0; jj < qlen; jj++) {
			const uint32_t j = seedOnLeft ? jj : qlen - 1 - jj;
			const uint32_t r = ref[rir + j];
			if (r & 4) { match = false; break; }
			uint32_t q = fw ? seq[j] : seq[qlen - 1 - j];
			if (!fw && q < 4) q ^= 3;
			if (q != r) {
				if (mms + 1 > g.refMms && jj < slen) { match = false; break; }
				const uint32_t qc = fw ? qual[j] : qual[qlen - 1 - j];
				ham += bt_mm_penalty(maq, bf_phred(qc));
				if (ham > g.refQualMax) { match = false; break; }
				X.A[editsOff + mms] = BF_EDIT(j, r);
				mms++;
				if (jj < slen) seedMms++;
			}
		}
		if (!match) continue;
		const uint32_t lo = L < aoff ? L : aoff, hi2 = L < aoff ? aoff : L;
		bool dup = false;
		for (uint32_t k = 0; k < pairs.n; k++) { const uint32_t *t = X.A + pairs.off + 3 * k; if (t[0] == tidx && t[1] == lo && t[2] == hi2) { dup = true; break; } }
		if (dup) continue;
		{   /* pairs->insert(p) */
			if (pairs.n == pairs.cap) {
				const uint32_t ncap = pairs.cap ? pairs.cap * 2 : 8;
				const uint32_t noff = bf_alloc(X, 3 * ncap);
				if (!noff) return false;
				for (uint32_t k = 0; k < 3 * pairs.n; k++) X.A[noff + k] = X.A[pairs.off + k];
				pairs.off = noff; pairs.cap = ncap;
			}
			uint32_t *t = X.A + pairs.off + 3 * pairs.n++;
			t[0] = tidx; t[1] = lo; t[2] = hi2;
		}
		out.nmm = mms; out.edits = editsOff; out.cost = seedMms << 14;           /* r.cost |= (r.stratum << 14), aligner.h:1064 */
		out.fw = fw; out.ebwtFw = 1; out.mate = (uint8_t)mate; out.pad = 0;
		result = L;
		found = true;
	}
	return found;
}

/* RangeChaser + RowChaser, step for step (range_chaser.h:40-210, row_chaser.h:60-110) */
struct BfChaser { uint32_t top, bot, irow, row, qlen, ebwtSel, tidx, toff; bool done, rowDone, hasOff; };
BT_FN void bf_chaser_set_row(BfCtx &X, BfChaser &c, uint32_t row) {
	const BtDevIndex &ix = X.P->ix[c.ebwtSel];
	c.row = row;
	for (;;) {
		if (c.row != ix.zOff && (c.row & ix.offMask) != c.row) { c.rowDone = false; return; }   /* needs a walk: RangeChaser::advance does it */
		c.rowDone = true;
		const uint32_t off = bf_resolve_row(X, ix, c.row);
		if (bt_joined_to_text(ix, c.qlen, off, c.tidx, c.toff)) { c.hasOff = true; return; }
		c.row++;
		if (c.row == c.bot) c.row = c.top;
		if (c.row == c.irow) { c.done = true; return; }
	}
}
BT_FN void bf_chaser_set_top_bot(BfCtx &X, BfChaser &c, uint32_t top, uint32_t bot, uint32_t qlen, uint32_t ebwtSel) {
	c.ebwtSel = ebwtSel; c.qlen = qlen; c.top = top; c.bot = bot;
	c.irow = top + bt_rand_next(X.randA) % (bot - top);
	c.done = false; c.hasOff = false;
	bf_chaser_set_row(X, c, c.irow);
}
BT_FN void bf_chaser_advance(BfCtx &X, BfChaser &c) {
	c.hasOff = false;
	if (c.rowDone) {
		c.row++;
		if (c.row == c.bot) c.row = c.top;
		if (c.row == c.irow) { c.done = true; return; }
		bf_chaser_set_row(X, c, c.row);
	} else {
		const BtDevIndex &ix = X.P->ix[c.ebwtSel];
		const uint32_t off = bf_resolve_row(X, ix, c.row);
		c.rowDone = true;
		if (bt_joined_to_text(ix, c.qlen, off, c.tidx, c.toff)) c.hasOff = true;
	}
}

struct BfPairState {
	BfCA dr[4];                       /* 0: mate 1 fw, 1: mate 1 rc, 2: mate 2 fw, 3: mate 2 rc                   */
	bool chase[4], delayed[4]; uint32_t offsSz[4];
	BfPairSet pairs[2];               /* [0] fw orientation of the pair, [1] rc                                     */
	BfChaser rc;
	uint32_t L, R, mixedAttempts;
	bool doneFw, doneFwFirst, done;
};

/* PairedBWAlignerV1::report (aligner.h:855-945): the upstream mate first, then the downstream one */
BT_NOINLINE bool bf_pair_report(BfCtx &X, const BfRangeView &rL, const BfRangeView &rR, uint32_t tidx, uint32_t upOff, uint32_t dnOff, bool pairFw) {
	const uint32_t spreadL = rL.bot - rL.top, spreadR = rR.bot - rR.top;
	const uint32_t oms = (spreadL < spreadR ? spreadL : spreadR) - 1;
	if (bf_report_hit(X, rL, tidx, upOff, oms, pairFw ? 1u : 2u, 2)) return true;    /* "can happen when -m is set" */
	return bf_report_hit(X, rR, tidx, dnOff, oms, pairFw ? 2u : 1u, 2);
}

/* PairedBWAlignerV1::resolveOutstandingInRef (aligner.h:951-1087).  off1: the anchor is mate 1. */
BT_NOINLINE bool bf_pair_resolve_in_ref(BfCtx &X, BfPairState &S, bool off1, uint32_t tidx, uint32_t toff, const BfSrc &range) {
	const BfProg &g = X.P->prog;
	const bool matchRight = off1 ? !S.doneFw : S.doneFw;
	bool fw = off1 ? (g.fw2 != 0) : (g.fw1 != 0);                       /* orientation of the outstanding mate */
	if (S.doneFw) fw = !fw;
	const uint32_t omate = off1 ? 1u : 0u;
	const uint32_t qlen = X.rlenM[omate], alen = X.rlenM[omate ^ 1];
	const uint32_t minins = g.minIns, maxins = g.maxIns;                    /* trimming adjustments are applied by the host (policy) */
	if (maxins <= (qlen > alen ? qlen : alen)) return false;
	uint32_t begin, end;
	const uint32_t insDiff = maxins - minins;
	const uint32_t approx = BT_LDG(X.P->ref.approxLen + tidx);
	if (matchRight) {
		end = toff + maxins;
		begin = toff + 1;
		if (qlen < alen) begin += alen - qlen;
		if (end > insDiff + qlen) { const uint32_t b2 = end - insDiff - qlen; if (b2 > begin) begin = b2; }
		if (end > approx) end = approx;
		if (begin > approx) begin = approx;
	} else {
		if (toff + alen < maxins) begin = 0; else begin = toff + alen - maxins;
		const uint32_t mi = alen < qlen ? alen : qlen;
		end = toff + mi - 1;
		const uint32_t e2 = toff + alen - minins + qlen - 1;
		if (e2 < end) end = e2;
		if (toff + alen + qlen < minins + 1) end = 0;
	}
	if (end < begin || end - begin < qlen) return false;
	const uint32_t mark = X.atop;
	BfRangeView r; uint32_t result = 0;
	const bool got = bf_ref_find(X, omate, fw, tidx, begin, end, toff, S.pairs[S.doneFw ? 1 : 0], r, result);
	bool ret = false;
	if (got) {
		r.top = range.rTop; r.bot = range.rBot;
		const BfRangeView a = bf_view_of(range);
		BfRangeView av = a;
		/* the found mate is reported as if aligned on the forward index; the anchor keeps its own index direction */
		ret = bf_pair_report(X, matchRight ? av : r, matchRight ? r : av, tidx, matchRight ? toff : result, matchRight ? result : toff, !S.doneFw);
	}
	/* scratch of this attempt (window, edits) is dead; the pair set may have been re-allocated above `mark` */
	if (X.atop > mark) {
		BfPairSet &ps = S.pairs[S.doneFw ? 1 : 0];
		if (!(ps.off >= mark)) X.atop = mark;
	}
	return ret;
}

/* Debug aid of the host emulation (g++ -DBF_TRACE_EVENTS): print the reference's own --verbose event messages so that the two
 * event streams can be diffed (`bowtie-align-s --verbose ... | grep -E "Chasing|Delaying|Resuming|Done with chase|Making an attempt"`). */
#if defined(BT_HOST_EMU) && defined(BF_TRACE_EVENTS)
#include <stdio.h>
#define BF_TRACE(msg) fprintf(stderr, "%s\n", msg)
#define BF_TRACE_RANGE(r) fprintf(stderr, "   range top=%u bot=%u idx=%u fw=%u mate=%u cost=%u nmm=%u e0=%x\n", (r).rTop, (r).rBot, (r).cfg.ebwtSel, (r).cfg.fw, (r).cfg.mate, (r).rCost, (r).rNmm, (r).rNmm ? X.A[(r).rEdits] : 0)
#else
#define BF_TRACE(msg) ((void)0)
#define BF_TRACE_RANGE(r) ((void)0)
#endif
/* PairedBWAlignerV1::advanceOrientation (aligner.h:1092-1326) */
BT_NOINLINE void bf_pair_advance_orientation(BfCtx &X, BfPairState &S, bool pairFw) {
	const BfProg &g = X.P->prog;
	const uint32_t L = S.L, R = S.R;
	BfCA &drL = S.dr[L], &drR = S.dr[R];
	bool &donePair = S.doneFw ? S.done : S.doneFw;
	const uint32_t qlenL = S.doneFw ? X.rlenM[1] : X.rlenM[0], qlenR = S.doneFw ? X.rlenM[0] : X.rlenM[1];
	if (S.chase[L]) {
		if (S.rc.hasOff) {
			if (!S.done) {                                                  /* overThresh || dontReconcile_ */
				BF_TRACE("Making an attempt to find the outstanding mate");
				const BfSrc &r = *BF_AT(BfSrc, X, drL.lastRange);
				S.done = bf_pair_resolve_in_ref(X, S, pairFw, S.rc.tidx, S.rc.toff, r);
				if (++S.mixedAttempts > g.mixedAttemptLim) { donePair = true; return; }
			}
			S.rc.hasOff = false;
		} else {
			S.chase[L] = false; drL.foundRange = 0;
			BF_TRACE("Done with chase for first mate");
			if (S.delayed[R]) {
				BF_TRACE("Resuming delayed chase for second mate");
				const BfSrc &r = *BF_AT(BfSrc, X, drR.lastRange);
				bf_chaser_set_top_bot(X, S.rc, r.rTop, r.rBot, qlenR, r.cfg.ebwtSel);
				S.chase[R] = true; S.delayed[R] = false;
			}
		}
	} else if (S.chase[R]) {
		if (S.rc.hasOff) {
			if (!S.done) {
				BF_TRACE("Making an attempt to find the outstanding mate");
				const BfSrc &r = *BF_AT(BfSrc, X, drR.lastRange);
				S.done = bf_pair_resolve_in_ref(X, S, !pairFw, S.rc.tidx, S.rc.toff, r);
				if (++S.mixedAttempts > g.mixedAttemptLim) { donePair = true; return; }
			}
			S.rc.hasOff = false;
		} else {
			S.chase[R] = false; drR.foundRange = 0;
			BF_TRACE("Done with chase for second mate");
			if (S.delayed[L]) {
				BF_TRACE("Resuming delayed chase for first mate");
				const BfSrc &r = *BF_AT(BfSrc, X, drL.lastRange);
				bf_chaser_set_top_bot(X, S.rc, r.rTop, r.rBot, qlenL, r.cfg.ebwtSel);
				S.chase[L] = true; S.delayed[L] = false;
			}
		}
	}
	if (!S.done && !donePair && !S.chase[L] && !S.chase[R]) {
		if ((S.offsSz[L] < S.offsSz[R] || drR.done) && !drL.done) {
			if (drR.done && S.offsSz[R] == 0) { BF_TRACE("Giving up on paired orientation in mate 1"); donePair = true; return; }
			if (!drL.foundRange) bf_ca_advance<true>(X, drL, g.strandFix != 0);
			if (X.flags & (BT_FLAG_STACK_OVF | BT_FLAG_FRAME_OVF)) return;
			if (drL.foundRange) {
				const BfSrc &rl = *BF_AT(BfSrc, X, drL.lastRange);
				S.offsSz[L] += rl.rBot - rl.rTop;
				if (S.offsSz[R] == 0 && S.offsSz[L] > 3) { BF_TRACE("Delaying a chase for first mate"); S.delayed[L] = true; }   /* !dontReconcile_ || offsLsz > 3 */
				else {
					BF_TRACE("Chasing a range for first mate");
					if (S.offsSz[L] > g.symCeiling && S.offsSz[R] > g.symCeiling) { donePair = true; return; }
					if (S.delayed[R] && S.offsSz[R] < S.offsSz[L]) {
						S.delayed[R] = false; S.delayed[L] = true; S.chase[R] = true;
						const BfSrc &r = *BF_AT(BfSrc, X, drR.lastRange);
						bf_chaser_set_top_bot(X, S.rc, r.rTop, r.rBot, qlenR, r.cfg.ebwtSel);
					} else {
						S.chase[L] = true;
						BF_TRACE_RANGE(rl);
						bf_chaser_set_top_bot(X, S.rc, rl.rTop, rl.rBot, qlenL, rl.cfg.ebwtSel);
					}
				}
			}
		} else if (!drR.done) {
			if (drL.done && S.offsSz[L] == 0) { BF_TRACE("Giving up on paired orientation in mate 2"); donePair = true; return; }
			if (!drR.foundRange) bf_ca_advance<true>(X, drR, g.strandFix != 0);
			if (X.flags & (BT_FLAG_STACK_OVF | BT_FLAG_FRAME_OVF)) return;
			if (drR.foundRange) {
				const BfSrc &rr = *BF_AT(BfSrc, X, drR.lastRange);
				S.offsSz[R] += rr.rBot - rr.rTop;
				if (S.offsSz[L] == 0 && S.offsSz[R] > 3) { BF_TRACE("Delaying a chase for second mate"); S.delayed[R] = true; }
				else {
					BF_TRACE("Chasing a range for second mate");
					if (S.offsSz[L] > g.symCeiling && S.offsSz[R] > g.symCeiling) { donePair = true; return; }
					if (S.delayed[L] && S.offsSz[L] < S.offsSz[R]) {
						S.delayed[L] = false; S.delayed[R] = true; S.chase[L] = true;
						const BfSrc &r = *BF_AT(BfSrc, X, drL.lastRange);
						bf_chaser_set_top_bot(X, S.rc, r.rTop, r.rBot, qlenL, r.cfg.ebwtSel);
					} else {
						S.chase[R] = true;
						bf_chaser_set_top_bot(X, S.rc, rr.rTop, rr.rBot, qlenR, rr.cfg.ebwtSel);
					}
				}
			}
		} else donePair = true;
	}
}

/* PairedBWAlignerV1::setQuery + the advance() loop of MixedMultiAligner (aligner.h:733-848, 244-304) */
BT_NOINLINE void bf_align_pair(BfCtx &X) {
	const BfKParams &P = *X.P;
	const BfProg &g = P.prog;
	X.randA = X.seedM[0];
	X.found = 0; X.bestStratum = 999; X.btCnt = (int32_t)P.pol.maxBtsBest;
	if (X.rlenM[0] < 4 || X.rlenM[1] < 4) return;                           /* "Skipping pair ... a mate is less than 4 characters long" */
	BfPairState S;
	for (uint32_t k = 0; k < 4; k++) {
		BfCA &ca = S.dr[k];
		ca.rssOff = bf_alloc(X, BF_MAX_TOP); ca.rssCap = BF_MAX_TOP; ca.nRss = 0;
		ca.actOff = bf_alloc(X, BF_MAX_TOP); ca.actCap = BF_MAX_TOP; ca.nAct = 0;
		ca.minCost = 0; ca.lastRange = ca.delayedRange = 0; ca.done = 0; ca.foundRange = 0; ca.rnd = 0; ca.paired = 0;
		S.chase[k] = S.delayed[k] = false; S.offsSz[k] = 0;
		if (X.flags & BT_FLAG_STACK_OVF) return;
		if (g.doList[k]) {
			const uint32_t wantFw = (k & 1) ? 0u : 1u, mate = k >> 1;
			for (uint32_t i = 0; i < g.ntop; i++) {
				BfTopCfg tc = mate ? g.top2[i] : g.top[i];
				if (tc.a.fw != wantFw) continue;
				tc.a.mate = (uint8_t)mate; tc.b.mate = (uint8_t)mate;
				uint32_t node;
				if (tc.kind == BF_KIND_SRC) node = bf_src_new(X, tc.a);
				else {
					node = bf_alloc_zero(X, BF_SEEDED_WORDS);
					if (node) {
						BfSeeded &sd = *BF_AT(BfSeeded, X, node);
						sd.h.kind = BF_KIND_SEEDED; sd.h.done = 1; sd.h.fw = tc.a.fw; sd.h.mate = tc.a.mate; sd.fact = tc.b;
						sd.seedgen = bf_src_new(X, tc.a);
					}
				}
				if (X.flags & BT_FLAG_STACK_OVF) return;
				X.A[ca.rssOff + ca.nRss++] = node;
			}
		}
	}
	for (uint32_t k = 0; k < 4; k++) {
		/* CostAwareRangeSourceDriver's constructor leaves an empty list "not done, nothing found"; setQuery on it returns early */
		bf_ca_set_query(X, S.dr[k]);
		if (X.flags & BT_FLAG_STACK_OVF) return;
	}
	S.pairs[0].off = S.pairs[0].cap = S.pairs[0].n = 0; S.pairs[1] = S.pairs[0];
	S.rc.done = false; S.rc.hasOff = false; S.rc.rowDone = true;
	S.doneFw = false; S.doneFwFirst = true; S.done = false; S.mixedAttempts = 0;
	S.L = g.fw1 ? 0u : 1u; S.R = g.fw2 ? 2u : 3u;
	while (!S.done) {
		if (X.flags & (BT_FLAG_STACK_OVF | BT_FLAG_FRAME_OVF)) return;
		if (S.doneFw && S.doneFwFirst) {
			S.L = g.fw2 ? 3u : 2u; S.R = g.fw1 ? 1u : 0u;
			S.doneFwFirst = false; S.mixedAttempts = 0;
		}
		const bool chasing = S.chase[S.L] || S.chase[S.R];
		if (chasing && !S.rc.hasOff && !S.rc.done) { bf_chaser_advance(X, S.rc); continue; }
		bf_pair_advance_orientation(X, S, !S.doneFw);
	}
}


/* ================================================================================================= */
/* Paired-end with --best / --strata / -M: PairedBWAlignerV2 (aligner.h:1483-2053).  One cost-aware  */
/* list holds the drivers of both mates and both strands; every located row of every range, in cost   */
/* order, anchors a reference scan for the opposite mate.  (reportSe, i.e. the per-mate sinks, is off  */
/* by default and not provided.)                                                                       */
/* ================================================================================================= */

/* PairedBWAlignerV2::resolveOutstandingInRef (aligner.h:1851-1963) */
BT_NOINLINE bool bf_pair2_resolve_in_ref(BfCtx &X, BfPairSet pairs[2], uint32_t tidx, uint32_t toff, const BfSrc &range) {
	const BfProg &g = X.P->prog;
	const bool mate1 = range.cfg.mate == 0;
	const bool pairFw = mate1 ? ((range.cfg.fw != 0) == (g.fw1 != 0)) : ((range.cfg.fw != 0) == (g.fw2 != 0));
	const bool matchRight = pairFw ? mate1 : !mate1;
	bool fw = mate1 ? (g.fw2 != 0) : (g.fw1 != 0);
	if (!pairFw) fw = !fw;
	const uint32_t omate = mate1 ? 1u : 0u;
	const uint32_t qlen = X.rlenM[omate], alen = X.rlenM[omate ^ 1];
	const uint32_t minins = g.minIns, maxins = g.maxIns;
	if (maxins <= (qlen > alen ? qlen : alen)) return false;
	uint32_t begin, end;
	const uint32_t insDiff = maxins - minins;
	const uint32_t approx = BT_LDG(X.P->ref.approxLen + tidx);
	if (matchRight) {
		end = toff + maxins;
		begin = toff + 1;
		if (qlen < alen) begin += alen - qlen;
		if (end > insDiff + qlen) { const uint32_t b2 = end - insDiff - qlen; if (b2 > begin) begin = b2; }
		if (end > approx) end = approx;
		if (begin > approx) begin = approx;
	} else {
		if (toff + alen < maxins) begin = 0; else begin = toff + alen - maxins;
		const uint32_t mi = alen < qlen ? alen : qlen;
		end = toff + mi - 1;
		const uint32_t e2 = toff + alen - minins + qlen - 1;
		if (e2 < end) end = e2;
		if (toff + alen + qlen < minins + 1) end = 0;
	}
	if (end < begin || end - begin < qlen) return false;
	const uint32_t mark = X.atop;
	BfRangeView r; uint32_t result = 0;
	BfPairSet &ps = pairs[pairFw ? 0 : 1];
	const bool got = bf_ref_find(X, omate, fw, tidx, begin, end, toff, ps, r, result);
	bool ret = false;
	if (got) {
		r.top = range.rTop; r.bot = range.rBot;
		const BfRangeView av = bf_view_of(range);
		ret = bf_pair_report(X, matchRight ? av : r, matchRight ? r : av, tidx, matchRight ? toff : result, matchRight ? result : toff, pairFw);
	}
	if (X.atop > mark && !(ps.off >= mark)) X.atop = mark;
	return ret;
}

/* PairedBWAlignerV2::setQuery + advance (aligner.h:1566-1700) without single-end reporting */
BT_NOINLINE void bf_align_pair_v2(BfCtx &X) {
	const BfKParams &P = *X.P;
	const BfProg &g = P.prog;
	X.randA = X.seedM[0];
	X.found = 0; X.bestStratum = 999; X.btCnt = (int32_t)P.pol.maxBtsBest;
	if (X.rlenM[0] < 4 || X.rlenM[1] < 4) return;
	BfCA &top = X.top;
	const uint32_t cap = 4 * BF_MAX_TOP;
	top.rssOff = bf_alloc(X, cap); top.rssCap = cap; top.nRss = 0;
	top.actOff = bf_alloc(X, cap); top.actCap = cap; top.nAct = 0;
	top.minCost = 0; top.lastRange = top.delayedRange = 0; top.done = 0; top.foundRange = 0;
	if (X.flags & BT_FLAG_STACK_OVF) return;
	/* the factories push the per-mate, per-strand lists in this order: -v: 1Fw 1Rc 2Fw 2Rc (aligner_0mm.h:322-325,
	 * aligner_1mm.h:284-420, aligner_23mm.h:352-610); -n: 1Fw 2Fw 1Rc 2Rc (aligner_seed_mm.h:707-1313) */
	const uint32_t orderV[4] = { 0, 1, 2, 3 }, orderN[4] = { 0, 2, 1, 3 };
	bool saw[2] = { false, false };
	for (uint32_t kk = 0; kk < 4; kk++) {
		const uint32_t k = (P.pol.mode == 0 ? orderV : orderN)[kk];
		if (!g.doList[k]) continue;
		const uint32_t wantFw = (k & 1) ? 0u : 1u, mate = k >> 1;
		for (uint32_t i = 0; i < g.ntop; i++) {
			BfTopCfg tc = mate ? g.top2[i] : g.top[i];
			if (tc.a.fw != wantFw) continue;
			tc.a.mate = (uint8_t)mate; tc.b.mate = (uint8_t)mate;
			uint32_t node;
			if (tc.kind == BF_KIND_SRC) node = bf_src_new(X, tc.a);
			else {
				node = bf_alloc_zero(X, BF_SEEDED_WORDS);
				if (node) {
					BfSeeded &sd = *BF_AT(BfSeeded, X, node);
					sd.h.kind = BF_KIND_SEEDED; sd.h.done = 1; sd.h.fw = tc.a.fw; sd.h.mate = (uint8_t)mate; sd.fact = tc.b;
					sd.seedgen = bf_src_new(X, tc.a);
				}
			}
			if (X.flags & BT_FLAG_STACK_OVF) return;
			X.A[top.rssOff + top.nRss++] = node;
			saw[mate] = true;
		}
	}
	top.paired = saw[0] && saw[1];
	bf_ca_set_query(X, top);
	BfPairSet pairs[2]; pairs[0].off = pairs[0].cap = pairs[0].n = 0; pairs[1] = pairs[0];
	BfChaser rc; rc.done = false; rc.hasOff = false; rc.rowDone = true;
	bool done = false, chase = false, donePe = false;
	uint32_t mixedAttempts = 0;
	const bool strandFix = g.strandFix != 0;
	while (!done) {
		if (X.flags & (BT_FLAG_STACK_OVF | BT_FLAG_FRAME_OVF)) return;
		if (chase) {
			if (!rc.hasOff && !rc.done) { bf_chaser_advance(X, rc); continue; }
			if (rc.hasOff) {
				/* resolveOutstanding (aligner.h:1827-1849) */
				if (!donePe) {
					const BfSrc &r = *BF_AT(BfSrc, X, top.lastRange);
					const bool ret = bf_pair2_resolve_in_ref(X, pairs, rc.tidx, rc.toff, r);
					if (++mixedAttempts > g.mixedAttemptLim || ret) donePe = true;
					done = donePe;
				}
				rc.hasOff = false;
			} else { chase = false; done = top.done; }
		}
		if (!done && !chase) {
			if (!top.done) {
				if (!donePe) { donePe = bf_irrelevant_cost(X, top.minCost); if (donePe) done = true; }
				if (!done) bf_ca_advance<true>(X, top, strandFix);
				if (X.flags & (BT_FLAG_STACK_OVF | BT_FLAG_FRAME_OVF)) return;
				if (top.foundRange) {
					chase = true; top.foundRange = 0;
					const BfSrc &r = *BF_AT(BfSrc, X, top.lastRange);
					bf_chaser_set_top_bot(X, rc, r.rTop, r.rBot, X.rlenM[r.cfg.mate], r.cfg.ebwtSel);
				}
			} else done = true;
		}
	}
}
