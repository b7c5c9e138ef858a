/*
 * bt_best_prog.h — the driver tree of the best-first ("stateful") path for one policy.
 *
 * The reference builds these trees in the create() methods of its aligner factories:
 *   -v 0        UnpairedExactAlignerV1Factory   aligner_0mm.h:66-114
 *   -v 1        Unpaired1mmAlignerV1Factory     aligner_1mm.h:70-152
 *   -v 2 / -v 3 Unpaired23mmAlignerV1Factory    aligner_23mm.h:70-222
 *   -n 0..3     UnpairedSeedAlignerFactory      aligner_seed_mm.h:84-541
 * Everything there that depends only on the policy is resolved here, on the host, into a table of at most
 * eight top-level drivers (a plain EbwtRangeSourceDriver, or an EbwtSeededRangeSourceDriver = seedling
 * generator + per-seedling extension driver); the kernel instantiates the table per read (bt_best.cuh).
 *
 * Plain C++ (host); shared by bt_lib.cu and the test-only emulation.
 */
#pragma once
#include <string.h>
#include "bt_best.cuh"

static inline BfSrcCfg bf_cfg(int mirror, int fw, int reportExacts, int hh, int seed, int nudgeLeft, int useBtCnt, int r0, int r1, int r2, int r3) {
	BfSrcCfg c; memset(&c, 0, sizeof c);
	c.ebwtSel = (uint8_t)mirror; c.fw = (uint8_t)fw; c.reportExacts = (uint8_t)reportExacts; c.hh = (uint8_t)hh; c.seed = (uint8_t)seed;
	c.nudgeLeft = (uint8_t)nudgeLeft; c.useBtCnt = (uint8_t)useBtCnt;
	c.rev[0] = (uint8_t)r0; c.rev[1] = (uint8_t)r1; c.rev[2] = (uint8_t)r2; c.rev[3] = (uint8_t)r3;
	return c;
}

/* Paired-end additions: PairedExact/1mm/23mm/SeedAlignerFactory::create (aligner_0mm.h:213-345, aligner_1mm.h:255-470,
 * aligner_23mm.h:319-650, aligner_seed_mm.h:650-1351) build the same driver lists once per mate and strand and pick the
 * RefAligner that matches the policy. */
static inline void bf_build_prog(int mode, int mms, int seedLen, uint32_t qualThresh, int nofw, int norc, BfProg *out,
                                 int paired = 0, int mate1fw = 1, int mate2fw = 0, uint32_t minIns = 0, uint32_t maxIns = 250,
                                 uint32_t pairTries = 100, uint32_t mhits = 0xffffffffu, int forMate2 = 0, int best = 0) {
	BfProg &g = *out; memset(&g, 0, sizeof g);
	if (paired) {
		g.paired = 1; g.pairedV2 = best ? 1 : 0; g.fw1 = mate1fw ? 1 : 0; g.fw2 = mate2fw ? 1 : 0; g.minIns = minIns; g.maxIns = maxIns;
		g.mixedAttemptLim = pairTries; g.symCeiling = mhits;
		g.refMms = (uint32_t)mms; g.refSeedLen = mode == 0 ? 0u : (uint32_t)seedLen; g.refQualMax = mode == 0 ? 0xffffffffu : qualThresh;
		bool d1f = true, d1r = true, d2f = true, d2r = true;
		if (nofw) { if (mate1fw) d1f = false; else d1r = false; if (mate2fw) d2f = false; else d2r = false; }
		if (norc) { if (mate1fw) d1r = false; else d1f = false; if (mate2fw) d2r = false; else d2f = false; }
		g.doList[0] = d1f; g.doList[1] = d1r; g.doList[2] = d2f; g.doList[3] = d2r;
		nofw = norc = 0;                                                 /* the table below holds both strands; doList selects */
	}
	const int B = BF_PIN_BEGINNING, L = BF_PIN_LEN, H = BF_PIN_HI_HALF, S = BF_PIN_SEED;
	g.strandFix = 1;                                                     /* ebwt_search.cpp:227 */
#define SRC(cfg) do { g.top[g.ntop].kind = BF_KIND_SRC; g.top[g.ntop].a = (cfg); g.ntop++; } while (0)
#define SEEDED(gen, fact) do { g.top[g.ntop].kind = BF_KIND_SEEDED; g.top[g.ntop].a = (gen); g.top[g.ntop].b = (fact); g.ntop++; } while (0)
	if (mode == 0) {
		g.seedLen = 0; g.qualLim = 0xffffffffu;                          /* "0 = whole read is seed"; qualLim OFF_MASK */
		for (int fw = 1; fw >= 0; fw--) {
			if (fw ? nofw : norc) continue;
			/* the first driver of a strand searches the index whose direction puts the unrevisitable half first:
			 * forward read -> mirror index, reverse complement -> forward index (aligner_1mm.h:82-135) */
			const int ia = fw ? 1 : 0, ib = fw ? 0 : 1;
			if (mms == 0) SRC(bf_cfg(0, fw, 1, 0, 0, 1, 0, L, L, L, L));             /* both strands on the forward index */
			else if (mms == 1) {
				/* nudgeLeft: "true for Fw index, false for Bw" unpaired (aligner_1mm.h:87-135); the paired factory gives
				 * true to the first and false to the second driver of every list (aligner_1mm.h:292-420) */
				SRC(bf_cfg(ia, fw, 1, 0, 0, paired ? 1 : (ia ? 0 : 1), 0, H, L, L, L));
				SRC(bf_cfg(ib, fw, 0, 0, 0, paired ? 0 : (ib ? 0 : 1), 0, H, L, L, L));
			} else {
				const int two = mms == 2, r2 = two ? L : H;
				SRC(bf_cfg(ia, fw, 1, 0, 0, 1, 0, H, H, r2, L));
				SRC(bf_cfg(ib, fw, 0, 0, 0, 0, 0, H, H, r2, L));
				SRC(bf_cfg(ia, fw, 0, 2, 0, 1, 0, B, H, r2, L));
				/* the 3-mismatch half-and-half driver: (B,H,H,L) unpaired and for mate 1 rc; (B,B,H,L) for mate 1 fw and both
				 * strands of mate 2 (aligner_23mm.h:126-136,191-201 vs 403-413,468-478,532-542,597-607) */
				if (!two) SRC(bf_cfg(ib, fw, 0, 3, 0, 0, 0, B, (paired && (fw || forMate2)) ? B : H, H, L));
			}
		}
	} else {
		g.seedLen = (uint32_t)seedLen; g.qualLim = qualThresh;
		const int bc = mms >= 2;                                         /* "no backtrack limit for -n 1/2" */
		for (int fw = 1; fw >= 0; fw--) {
			if (fw ? nofw : norc) continue;
			const int ia = fw ? 1 : 0, ib = fw ? 0 : 1;                  /* ia: extension index, ib: seedling index */
			const BfSrcCfg fact = bf_cfg(ia, fw, 1, 0, 0, 1, bc, S, S, S, S);
			if (mms == 0) SRC(bf_cfg(ia, fw, 1, 0, 0, 1, 0, S, S, S, S));
			else if (mms == 1) {
				SRC(bf_cfg(ia, fw, 1, 0, 0, 1, 0, H, S, S, S));
				SEEDED(bf_cfg(ib, fw, 0, 0, 1, 0, 0, H, S, S, S), fact);
			} else if (mms == 2) {
				SRC(bf_cfg(ia, fw, 1, 0, 0, 1, 1, H, H, S, S));
				SEEDED(bf_cfg(ib, fw, 0, 0, 1, 0, 1, H, H, S, S), fact);
				SRC(bf_cfg(ia, fw, 0, 2, 0, 1, 1, B, H, S, S));
			} else {
				SRC(bf_cfg(ia, fw, 1, 0, 0, 1, 1, H, H, H, S));
				SEEDED(bf_cfg(ib, fw, 0, 0, 1, 0, 1, H, H, H, S), fact);
				SEEDED(bf_cfg(ib, fw, 0, 3, 1, 0, 1, B, H, H, S), fact);
				SRC(bf_cfg(ia, fw, 0, 2, 0, 1, 1, B, H, H, S));
			}
		}
	}
#undef SRC
#undef SEEDED
	if (paired && !forMate2) {
		BfProg m2;
		bf_build_prog(mode, mms, seedLen, qualThresh, 0, 0, &m2, 1, mate1fw, mate2fw, minIns, maxIns, pairTries, mhits, 1);
		memcpy(g.top2, m2.top, sizeof g.top2);
	}
}
