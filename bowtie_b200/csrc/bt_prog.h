/*
 * bt_prog.h — the per-policy "phase program": which backtracker invocations a read goes through.
 *
 * The reference hard-codes these sequences in the search_*.c fragments that are #included into the
 * worker loops (search_exact.c, search_1mm_phase{1,2}.c, search_23mm_phase{1,2,3}.c,
 * search_seeded_phase{1,2,3,4}.c) and in the constructor arguments of the GreedyDFSRangeSource objects
 * of each worker (ebwt_search.cpp:1155-1166, 1632-1644, 2082-2134, 2413-2539).  Here the host resolves
 * everything that depends only on the policy (-v/-n, mismatches, --nofw/--norc) into a short table of
 * packed steps; the kernel interprets it (bt_phase in bt_core.cuh).  Offsets that depend on the read
 * length are selected from a per-read value table V[] = { 0, len, S, S3, S5 } where S = min(len, s),
 * S3 = S >> 1, S5 = S3 + (S & 1) and s = len for -v modes, seedLen for -n modes.
 *
 * Plain C++ (host); shared by bt_lib.cu and the test-only emulation.
 */
#pragma once
#include <stdint.h>

#define BT_PROG_MAX 16

/* step kinds */
enum { BTK_END = 0, BTK_LAUNCH = 1, BTK_FILTER = 2, BTK_SEEDLOOP = 3 };
/* value-table selectors */
enum { BTV_0 = 0, BTV_LEN = 1, BTV_S = 2, BTV_S3 = 3, BTV_S5 = 4 };

/* bit layout of one step */
#define BTS_KIND(x)    ((x) & 7u)
#define BTS_EBWT(x)    (((x) >> 3) & 1u)    /* 0 forward index, 1 mirror                               */
#define BTS_FW(x)      (((x) >> 4) & 1u)    /* read orientation (EbwtSearchParams::_fw)                 */
#define BTS_CQ(x)      (((x) >> 5) & 1u)    /* considerQuals                                            */
#define BTS_HH(x)      (((x) >> 6) & 1u)    /* halfAndHalf                                              */
#define BTS_RP(x)      (((x) >> 7) & 1u)    /* reportPartials = policy.mms                              */
#define BTS_RE(x)      (((x) >> 8) & 1u)    /* reportExacts                                             */
#define BTS_SEEDQ(x)   (((x) >> 9) & 1u)    /* setQlen(seedLen): search the seed only                   */
#define BTS_IGNORE(x)  (((x) >> 10) & 1u)   /* the worker ignores backtrack()'s return value            */
#define BTS_CLEARP(x)  (((x) >> 11) & 1u)   /* start with an empty seedling list                        */
#define BTS_SEL(x, k)  (((x) >> (12 + 3 * (k))) & 7u)   /* k = 0..5: depth5, depth3, unrev, rev1, rev2, rev3 */

static inline uint32_t bts_make(uint32_t kind, uint32_t ebwt, uint32_t fw, uint32_t cq, uint32_t hh, uint32_t rp, uint32_t re,
                                uint32_t seedq, uint32_t ignore, uint32_t clearp,
                                uint32_t d5, uint32_t d3, uint32_t un, uint32_t r1, uint32_t r2, uint32_t r3) {
	return kind | (ebwt << 3) | (fw << 4) | (cq << 5) | (hh << 6) | (rp << 7) | (re << 8) | (seedq << 9) | (ignore << 10) | (clearp << 11) |
	       (d5 << 12) | (d3 << 15) | (un << 18) | (r1 << 21) | (r2 << 24) | (r3 << 27);
}

/* Builds the program for a policy; returns the number of steps (including the terminating END). */
static inline int bt_build_prog(int mode, int mms, int nofw, int norc, uint32_t prog[BT_PROG_MAX]) {
	int n = 0;
	const uint32_t O = BTV_0, L = BTV_LEN, S = BTV_S, S3 = BTV_S3, S5 = BTV_S5;
#define LAUNCH(...) prog[n++] = bts_make(BTK_LAUNCH, __VA_ARGS__)
	if (mode == 0 && mms == 0) {
		/* search_exact.c:7-27 */
		if (!nofw) LAUNCH(0, 1, 0, 0, 0, 1, 0, 0, 0, O, O, L, L, L, L);
		if (!norc) LAUNCH(0, 0, 0, 0, 0, 1, 0, 0, 0, O, O, L, L, L, L);
	} else if (mode == 0 && mms == 1) {
		/* search_1mm_phase1.c, search_1mm_phase2.c (s = len) */
		if (!nofw) LAUNCH(0, 1, 0, 0, 0, 1, 0, 0, 0, O, O, S, S, S, S);
		if (!norc) LAUNCH(0, 0, 0, 0, 0, 1, 0, 0, 0, O, O, S, S, S, S);
		if (!norc) LAUNCH(0, 0, 0, 0, 0, 0, 0, 0, 0, O, O, S5, S, S, S);
		if (!nofw) LAUNCH(0, 1, 0, 0, 0, 0, 0, 0, 0, O, O, S5, S, S, S);
		if (!norc) LAUNCH(1, 0, 0, 0, 0, 0, 0, 0, 0, O, O, S3, S, S, S);
		if (!nofw) LAUNCH(1, 1, 0, 0, 0, 0, 0, 0, 0, O, O, S3, S, S, S);
	} else if (mode == 0) {
		/* search_23mm_phase1.c, _phase2.c, _phase3.c with two = true */
		if (!nofw) LAUNCH(0, 1, 0, 0, 0, 1, 0, 0, 0, O, O, L, L, L, L);            /* btr1 fw exact              */
		if (!norc) LAUNCH(0, 0, 0, 0, 0, 1, 0, 0, 0, O, O, S5, S5, S, S);          /* btr1 rc                    */
		if (!nofw) LAUNCH(1, 1, 0, 0, 0, 0, 0, 0, 0, O, O, S5, S5, S, S);          /* bt2 fw                     */
		if (!norc) LAUNCH(1, 0, 0, 0, 0, 0, 0, 0, 0, O, O, S3, S3, S, S);          /* bt2 rc                     */
		if (!nofw) LAUNCH(0, 1, 0, 0, 0, 0, 0, 0, 0, O, O, S3, S3, S, S);          /* bt3 fw                     */
		if (!nofw) LAUNCH(0, 1, 0, 1, 0, 1, 0, 0, 0, S3, S, O, S3, S, S);          /* bthh3 fw                   */
		if (!norc) LAUNCH(0, 0, 0, 1, 0, 1, 0, 0, 0, S5, S, O, S5, S, S);          /* bthh3 rc                   */
	} else {
		/* search_seeded_phase1.c .. phase4.c; m = seedMms */
		const int m = mms;
		const uint32_t a0 = m > 0 ? S5 : S, a1 = m > 1 ? S5 : S, a2 = m > 2 ? S5 : S, a3 = m > 3 ? S5 : S;
		const uint32_t b1 = m > 1 ? S3 : S, b2 = m > 2 ? S3 : S, b3 = m > 3 ? S3 : S;
		prog[n++] = bts_make(BTK_FILTER, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
		if (!nofw) LAUNCH(0, 1, 0, 0, 0, 1, 0, 0, 0, O, L, L, L, L, L);            /* btf1                       */
		if (!norc) LAUNCH(0, 0, 1, 0, 0, 1, 0, 0, 0, O, O, a0, a1, a2, a3);        /* bt1   cases 1R 2R 3R       */
		if (!nofw) LAUNCH(1, 1, 1, 0, 0, 0, 0, 0, 0, O, O, a0, a1, a2, a3);        /* btf2  cases 1F 2F 3F       */
		if (m > 0) {
			if (!norc) {
				LAUNCH(1, 0, 1, 0, 1, nofw ? 1u : 0u, 1, 1, 1, O, O, S3, b1, b2, b3);    /* btr2: 4R seedlings  */
				prog[n++] = bts_make(BTK_SEEDLOOP, 0, 0, 1, 0, 0, 1, 0, 0, 0, O, O, S, S, S, S);   /* btr3        */
				if (m >= 2) LAUNCH(0, 0, 1, 1, 0, 1, 0, 0, 0, S5, S, O, m <= 2 ? S5 : O, m < 3 ? S : S5, S);   /* btr23 */
			}
			if (!nofw) {
				LAUNCH(0, 1, 1, 0, 1, 1, 1, 1, 1, O, O, S3, b1, b2, b3);             /* btf3: 4F seedlings         */
				prog[n++] = bts_make(BTK_SEEDLOOP, 1, 1, 1, 0, 0, 1, 0, 0, 0, O, O, S, S, S, S);   /* btf4        */
				if (m >= 2) LAUNCH(1, 1, 1, 1, 0, 1, 0, 0, 0, S5, S, O, m <= 2 ? S5 : O, m < 3 ? S : S5, S);   /* btf24 */
			}
		}
	}
#undef LAUNCH
	prog[n++] = BTK_END;
	return n;
}
