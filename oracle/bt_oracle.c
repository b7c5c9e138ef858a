/*
 * bt_oracle.c — TEST INFRASTRUCTURE ONLY (see bt_oracle.h).
 *
 * Plain-C restatement of the Bowtie 1.3.1 non-stateful search path:
 *   index parsing        ebwt.h:2835-3445 (readIntoMemory), 1043-1059 (postReadInit)
 *   LF arithmetic        ebwt.h:1438-1497 (SideLocus), 1897-2226 (count*), 2334-2560 (mapLF*)
 *   ftab                 ebwt.h:985-1034
 *   locate               ebwt.h:2693-2755 (reportChaseOne), 2569-2629 (joinedToTextOff)
 *   DFS backtracker      ebwt_search_backtrack.h:23-1779 (GreedyDFSRangeSource)
 *   phase loops          search_exact.c, search_1mm_phase{1,2}.c, search_23mm_phase{1,2,3}.c,
 *                        search_seeded_phase{1,2,3,4}.c ; workers ebwt_search.cpp:1130,1606,2056,2378
 *   seedlings            ebwt_search_util.h:16-370
 *   hit sinks            hit.h:937-985 (NGood), 1177-1222 (AllHit), 741-786 (finishRead)
 *   penalties / RNG      qual.h:15-67, qual.cpp:4-32, random_source.h:27-54, pat.cpp:21-57
 *
 * Written to be read side by side with the reference; sequential, recursion allowed.
 * The CUDA product re-derives the same semantics with its own data layout and an
 * explicit frame stack; it shares no code with this file.
 */
#include "bt_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* index loading                                                       */
/* ------------------------------------------------------------------ */

static int rd_u32(FILE *f, uint32_t *v) { return fread(v, 4, 1, f) == 1 ? 0 : -1; }

static void set_err(char *err, size_t n, const char *msg, const char *arg) {
	if (err && n) snprintf(err, n, "%s%s", msg, arg ? arg : "");
}

/* EbwtParams::init (ebwt.h:138-184) + Ebwt::readIntoMemory (ebwt.h:2926-3445) for the
 * small-index, non-bt2 format (Appendix A of SURVEY.md). */
bto_index *bto_index_load(const char *basename, int mirror, char *err, size_t errlen) {
	char path[4096];
	bto_index *ix = (bto_index *)calloc(1, sizeof(*ix));
	FILE *f1 = NULL, *f2 = NULL;
	uint32_t one, flags_u;
	int32_t flags;
	if (!ix) { set_err(err, errlen, "out of memory", NULL); return NULL; }
	snprintf(path, sizeof path, "%s%s.1.ebwt", basename, mirror ? ".rev" : "");
	f1 = fopen(path, "rb");
	if (!f1) { set_err(err, errlen, "cannot open ", path); goto fail; }
	snprintf(path, sizeof path, "%s%s.2.ebwt", basename, mirror ? ".rev" : "");
	f2 = fopen(path, "rb");
	if (!f2) { set_err(err, errlen, "cannot open ", path); goto fail; }
	ix->fw = !mirror;
	if (rd_u32(f1, &one) || one != 1) { set_err(err, errlen, "bad endianness sentinel in .1.ebwt", NULL); goto fail; }
	if (rd_u32(f2, &one) || one != 1) { set_err(err, errlen, "bad endianness sentinel in .2.ebwt", NULL); goto fail; }
	if (rd_u32(f1, &ix->len) || rd_u32(f1, (uint32_t *)&ix->lineRate) || rd_u32(f1, (uint32_t *)&ix->linesPerSide) ||
	    rd_u32(f1, (uint32_t *)&ix->offRate) || rd_u32(f1, (uint32_t *)&ix->ftabChars) || rd_u32(f1, &flags_u)) {
		set_err(err, errlen, "truncated header", NULL); goto fail;
	}
	flags = (int32_t)flags_u;
	ix->entireRev = !(flags < 0 && (((-flags) & 4) == 0)); /* ebwt.h:2965-2973 */
	if (ix->lineRate != 6 || ix->linesPerSide != 1) {
		set_err(err, errlen, "unsupported side geometry (need lineRate 6, linesPerSide 1)", NULL); goto fail;
	}
	ix->bwtLen = ix->len + 1;
	ix->offMask = BTO_OFF_MASK << ix->offRate;
	ix->sideSz = 64; ix->sideBwtSz = 56; ix->sideBwtLen = 224;
	{
		uint32_t bwtSz = ix->len / 4 + 1;
		uint32_t numSidePairs = (bwtSz + 2 * ix->sideBwtSz - 1) / (2 * ix->sideBwtSz);
		ix->numSides = numSidePairs * 2;
		ix->ebwtTotLen = numSidePairs * 2 * ix->sideSz;
	}
	ix->ftabLen = (1u << (ix->ftabChars * 2)) + 1;
	ix->eftabLen = (uint32_t)ix->ftabChars * 2;
	ix->offsLen = (ix->bwtLen + (1u << ix->offRate) - 1) >> ix->offRate;
	if (rd_u32(f1, &ix->nPat)) goto trunc;
	ix->plen = (uint32_t *)malloc(sizeof(uint32_t) * (ix->nPat ? ix->nPat : 1));
	if (fread(ix->plen, 4, ix->nPat, f1) != ix->nPat) goto trunc;
	if (rd_u32(f1, &ix->nFrag)) goto trunc;
	ix->rstarts = (uint32_t *)malloc(sizeof(uint32_t) * 3 * (ix->nFrag ? ix->nFrag : 1));
	if (fread(ix->rstarts, 4, 3 * (size_t)ix->nFrag, f1) != 3 * (size_t)ix->nFrag) goto trunc;
	ix->ebwt = (uint8_t *)malloc(ix->ebwtTotLen);
	if (fread(ix->ebwt, 1, ix->ebwtTotLen, f1) != ix->ebwtTotLen) goto trunc;
	if (rd_u32(f1, &ix->zOff)) goto trunc;
	if (fread(ix->fchr, 4, 5, f1) != 5) goto trunc;
	ix->ftab = (uint32_t *)malloc(sizeof(uint32_t) * ix->ftabLen);
	if (fread(ix->ftab, 4, ix->ftabLen, f1) != ix->ftabLen) goto trunc;
	ix->eftab = (uint32_t *)malloc(sizeof(uint32_t) * ix->eftabLen);
	if (fread(ix->eftab, 4, ix->eftabLen, f1) != ix->eftabLen) goto trunc;
	/* reference names: '\n'-separated, '\0'-terminated (ebwt.h:3258-3272) */
	{
		size_t cap = 16; int c;
		ix->refnames = (char **)calloc(cap, sizeof(char *));
		ix->nRefnames = 0;
		while ((c = fgetc(f1)) != EOF) {
			if (c == '\0') break;
			if (c == '\n') {
				if (ix->nRefnames == cap) { cap *= 2; ix->refnames = (char **)realloc(ix->refnames, cap * sizeof(char *)); }
				ix->refnames[ix->nRefnames++] = (char *)calloc(1, 1);
			} else {
				char *s; size_t l;
				if (ix->nRefnames == 0) ix->refnames[ix->nRefnames++] = (char *)calloc(1, 1);
				s = ix->refnames[ix->nRefnames - 1]; l = strlen(s);
				s = (char *)realloc(s, l + 2); s[l] = (char)c; s[l + 1] = 0;
				ix->refnames[ix->nRefnames - 1] = s;
			}
		}
	}
	ix->offs = (uint32_t *)malloc(sizeof(uint32_t) * ix->offsLen);
	if (fread(ix->offs, 4, ix->offsLen, f2) != ix->offsLen) goto trunc;
	/* postReadInit (ebwt.h:1043-1059) */
	{
		uint32_t sideNum = ix->zOff / ix->sideBwtLen;
		uint32_t sideCharOff = ix->zOff % ix->sideBwtLen;
		uint32_t sideByteOff = sideNum * ix->sideSz;
		ix->zEbwtByteOff = sideCharOff >> 2;
		ix->zEbwtBpOff = (int)(sideCharOff & 3);
		if ((sideNum & 1) == 0) {
			ix->zEbwtByteOff = ix->sideBwtSz - ix->zEbwtByteOff - 1;
			ix->zEbwtBpOff = 3 - ix->zEbwtBpOff;
		}
		ix->zEbwtByteOff += sideByteOff;
	}
	fclose(f1); fclose(f2);
	return ix;
trunc:
	set_err(err, errlen, "truncated index file for ", basename);
fail:
	if (f1) fclose(f1);
	if (f2) fclose(f2);
	bto_index_free(ix);
	return NULL;
}

void bto_index_free(bto_index *ix) {
	uint32_t i;
	if (!ix) return;
	free(ix->plen); free(ix->rstarts); free(ix->ebwt); free(ix->ftab); free(ix->eftab); free(ix->offs);
	if (ix->refnames) { for (i = 0; i < ix->nRefnames; i++) free(ix->refnames[i]); free(ix->refnames); }
	free(ix);
}

/* ------------------------------------------------------------------ */
/* LF arithmetic                                                       */
/* ------------------------------------------------------------------ */

/* SideLocus (ebwt.h:1418-1523) */
typedef struct { uint32_t sideByteOff, sideNum; uint32_t charOff; int fw; int by; int bp; } locus_t;

/* SideLocus::initFromRow (ebwt.h:1469-1497) */
static void locus_from_row(const bto_index *ix, uint32_t row, locus_t *l) {
	l->sideNum = row / 224;
	l->charOff = row % 224;
	l->sideByteOff = l->sideNum * ix->sideSz;
	l->fw = (l->sideNum & 1) != 0;
	l->by = (int)(l->charOff >> 2);
	l->bp = (int)(l->charOff & 3);
	if (!l->fw) { l->by = (int)ix->sideBwtSz - l->by - 1; l->bp ^= 3; }
}

/* SideLocus::initFromTopBot (ebwt.h:1438-1463) */
static void locus_from_topbot(const bto_index *ix, uint32_t top, uint32_t bot, locus_t *lt, locus_t *lb) {
	uint32_t spread = bot - top;
	locus_from_row(ix, top, lt);
	if (lt->charOff + spread < ix->sideBwtLen) {
		lb->charOff = lt->charOff + spread;
		lb->sideNum = lt->sideNum;
		lb->sideByteOff = lt->sideByteOff;
		lb->fw = lt->fw;
		lb->by = (int)(lb->charOff >> 2);
		if (!lb->fw) lb->by = (int)ix->sideBwtSz - lb->by - 1;
		lb->bp = (int)(lb->charOff & 3);
		if (!lb->fw) lb->bp ^= 3;
	} else {
		locus_from_row(ix, bot, lb);
	}
}

/* rowL (ebwt.h: Ebwt::rowL): bit-pair bp of byte by of the side */
static int row_l(const bto_index *ix, const locus_t *l) {
	return (ix->ebwt[l->sideByteOff + (uint32_t)l->by] >> (l->bp * 2)) & 3;
}

/* countUpToEx (ebwt.h:1963-2027) stated through the cCntLUT_4 semantics (ccnt_lut.cpp):
 * occurrences of each character in bytes [0,by) plus the low bp bit-pairs of byte by. */
static void count_up_to_ex(const bto_index *ix, const locus_t *l, uint32_t arrs[4]) {
	const uint8_t *side = ix->ebwt + l->sideByteOff;
	int i, k;
	for (i = 0; i < l->by; i++) {
		uint8_t b = side[i];
		arrs[b & 3]++; arrs[(b >> 2) & 3]++; arrs[(b >> 4) & 3]++; arrs[(b >> 6) & 3]++;
	}
	for (k = 0; k < l->bp; k++) arrs[(side[l->by] >> (2 * k)) & 3]++;
}

static int z_adjust(const bto_index *ix, const locus_t *l, int inclusive) {
	/* '$' is stored as an A but must not be counted (ebwt.h:2044-2052 fw '>' ; 2147-2155 bw '>=') */
	uint32_t B = l->sideByteOff + (uint32_t)l->by;
	if (l->sideByteOff <= ix->zEbwtByteOff && B >= ix->zEbwtByteOff) {
		if (B > ix->zEbwtByteOff) return 1;
		if (B == ix->zEbwtByteOff && (inclusive ? l->bp >= ix->zEbwtBpOff : l->bp > ix->zEbwtBpOff)) return 1;
	}
	return 0;
}

/* countFwSideEx / countBwSideEx (ebwt.h:2081-2129, 2184-2226) */
static void count_side_ex(const bto_index *ix, const locus_t *l, uint32_t arrs[4]) {
	const uint8_t *side = ix->ebwt + l->sideByteOff;
	arrs[0] = arrs[1] = arrs[2] = arrs[3] = 0;
	count_up_to_ex(ix, l, arrs);
	if (l->fw) {
		const uint32_t *ac = (const uint32_t *)(side - 8);
		const uint32_t *gt = (const uint32_t *)(side + ix->sideSz - 8);
		if (z_adjust(ix, l, 0)) arrs[0]--;
		arrs[0] += ac[0] + ix->fchr[0];
		arrs[1] += ac[1] + ix->fchr[1];
		arrs[2] += gt[0] + ix->fchr[2];
		arrs[3] += gt[1] + ix->fchr[3];
	} else {
		const uint32_t *ac = (const uint32_t *)(side + ix->sideSz - 8);
		const uint32_t *gt = (const uint32_t *)(side + 2 * ix->sideSz - 8);
		arrs[row_l(ix, l)]++;
		if (z_adjust(ix, l, 1)) arrs[0]--;
		arrs[0] = ac[0] - arrs[0] + ix->fchr[0];
		arrs[1] = ac[1] - arrs[1] + ix->fchr[1];
		arrs[2] = gt[0] - arrs[2] + ix->fchr[2];
		arrs[3] = gt[1] - arrs[3] + ix->fchr[3];
	}
}

/* countFwSide / countBwSide (ebwt.h:2034-2074, 2136-2177): single character */
static uint32_t count_side(const bto_index *ix, const locus_t *l, int c) {
	uint32_t arrs[4];
	count_side_ex(ix, l, arrs);
	return arrs[c];
}

uint32_t bto_map_lf(const bto_index *ix, uint32_t row, int c) {
	locus_t l; locus_from_row(ix, row, &l); return count_side(ix, &l, c);
}
uint32_t bto_map_lf1(const bto_index *ix, uint32_t row, int c) {
	locus_t l; locus_from_row(ix, row, &l);
	if (row_l(ix, &l) != c || row == ix->zOff) return BTO_OFF_MASK; /* ebwt.h:2501 */
	return count_side(ix, &l, c);
}
void bto_map_lf_ex(const bto_index *ix, uint32_t row, uint32_t out[4]) {
	locus_t l; locus_from_row(ix, row, &l); count_side_ex(ix, &l, out);
}
int bto_row_l(const bto_index *ix, uint32_t row) { locus_t l; locus_from_row(ix, row, &l); return row_l(ix, &l); }

/* ftabHi / ftabLo (ebwt.h:985-1034) */
uint32_t bto_ftab_hi(const bto_index *ix, uint32_t i) {
	if (ix->ftab[i] <= ix->len) return ix->ftab[i];
	return ix->eftab[(ix->ftab[i] ^ BTO_OFF_MASK) * 2 + 1];
}
uint32_t bto_ftab_lo(const bto_index *ix, uint32_t i) {
	if (ix->ftab[i] <= ix->len) return ix->ftab[i];
	return ix->eftab[(ix->ftab[i] ^ BTO_OFF_MASK) * 2];
}

/* Row chase of Ebwt::reportChaseOne (ebwt.h:2711-2746) */
static uint32_t chase_impl(const bto_index *ix, uint32_t i, uint32_t *jumps_out, uint32_t *endrow) {
	uint32_t jumps = 0, off;
	while (((i & ix->offMask) != i) && i != ix->zOff) {
		locus_t l; locus_from_row(ix, i, &l);
		i = count_side(ix, &l, row_l(ix, &l)); /* mapLF(l) ebwt.h:2420-2452 */
		jumps++;
	}
	if (i == ix->zOff) off = jumps;
	else off = ix->offs[i >> ix->offRate] + jumps;
	if (jumps_out) *jumps_out = jumps;
	if (endrow) *endrow = i;
	return off;
}
uint32_t bto_chase(const bto_index *ix, uint32_t i, uint32_t *jumps_out) { return chase_impl(ix, i, jumps_out, NULL); }

/* Ebwt::joinedToTextOff (ebwt.h:2569-2629) */
void bto_joined_to_text_off(const bto_index *ix, uint32_t qlen, uint32_t off,
                            uint32_t *tidx, uint32_t *textoff, uint32_t *tlen) {
	uint32_t top = 0, bot = ix->nFrag, elt;
	*tidx = BTO_OFF_MASK; *textoff = 0; *tlen = 0;
	for (;;) {
		uint32_t lower, upper, fraglen;
		elt = top + ((bot - top) >> 1);
		lower = ix->rstarts[elt * 3];
		upper = (elt == ix->nFrag - 1) ? ix->len : ix->rstarts[(elt + 1) * 3];
		fraglen = upper - lower;
		if (lower <= off) {
			if (upper > off) {
				uint32_t fragoff;
				if (off + qlen > upper) { *tidx = BTO_OFF_MASK; return; }
				*tidx = ix->rstarts[elt * 3 + 1];
				fragoff = off - ix->rstarts[elt * 3];
				if (!ix->fw) { fragoff = fraglen - fragoff - 1; fragoff -= (qlen - 1); }
				*textoff = fragoff + ix->rstarts[elt * 3 + 2];
				break;
			} else top = elt;
		} else bot = elt;
	}
	*tlen = ix->plen[*tidx];
}

/* ------------------------------------------------------------------ */
/* penalties, RNG, seed                                                */
/* ------------------------------------------------------------------ */

/* qualRounds[] (qual.cpp:4-32): 0-4 -> 0, 5-14 -> 10, 15-24 -> 20, >=25 -> 30 */
static uint8_t qual_round(uint8_t q) { return q < 5 ? 0 : q < 15 ? 10 : q < 25 ? 20 : 30; }
/* mmPenalty (qual.h:55-61) */
static uint8_t mm_penalty(int maq, uint8_t q) { return maq ? qual_round(q) : q; }
/* phredCharToPhredQual (qual.h:15-17) */
static uint8_t phred_of(uint8_t c) { return c >= 33 ? (uint8_t)(c - 33) : 0; }

uint32_t bto_rand_next(uint32_t *last) {
	uint32_t ret;
	*last = 1664525u * (*last) + 1013904223u;
	ret = *last >> 16;
	*last = 1664525u * (*last) + 1013904223u;
	ret ^= *last;
	return ret;
}

uint32_t bto_gen_rand_seed(const uint8_t *seq, const uint8_t *qual, uint32_t len,
                           const char *name, uint32_t namelen, uint32_t seed) {
	uint32_t rseed = (seed + 101u) * 59u * 61u * 67u * 71u * 73u * 79u * 83u;
	uint32_t i;
	for (i = 0; i < len; i++) rseed ^= ((uint32_t)seq[i] << ((i & 15) << 1));
	for (i = 0; i < len; i++) rseed ^= ((uint32_t)qual[i] << ((i & 3) << 3));
	for (i = 0; i < namelen; i++) rseed ^= ((uint32_t)(uint8_t)name[i] << ((i & 3) << 3));
	return rseed;
}

/* ------------------------------------------------------------------ */
/* hit sink (NGood / AllHit)                                           */
/* ------------------------------------------------------------------ */

typedef struct {
	uint32_t n, max;            /* _n (khits or 0xffffffff for -a), _max (mhits) */
	uint32_t hitsForThisRead;
	uint64_t numValidHits;
	/* buffered hits of the current read */
	bto_hit *buf; size_t nbuf, capbuf;
	bto_mm  *mmbuf; size_t nmm, capmm;
} sink_t;

/* NGoodHitSinkPerThread::reportHit (hit.h:969-985) / AllHitSinkPerThread::reportHit (hit.h:1201-1209) */
static int sink_report(sink_t *s, const bto_hit *h, const bto_mm *mms) {
	s->numValidHits++;
	s->hitsForThisRead++;
	if (s->hitsForThisRead > s->max) return 1;
	if (s->nbuf == s->capbuf) { s->capbuf = s->capbuf ? s->capbuf * 2 : 16; s->buf = (bto_hit *)realloc(s->buf, s->capbuf * sizeof(bto_hit)); }
	if (s->nmm + h->nmm > s->capmm) { s->capmm = (s->capmm + h->nmm) * 2 + 16; s->mmbuf = (bto_mm *)realloc(s->mmbuf, s->capmm * sizeof(bto_mm)); }
	s->buf[s->nbuf] = *h;
	s->buf[s->nbuf].mm_off = (uint32_t)s->nmm;
	memcpy(s->mmbuf + s->nmm, mms, h->nmm * sizeof(bto_mm));
	s->nmm += h->nmm; s->nbuf++;
	if (s->n != 0xffffffffu && s->hitsForThisRead == s->n && (s->max == 0xffffffffu || s->max < s->n)) return 1;
	return 0;
}

/* ------------------------------------------------------------------ */
/* seedlings                                                           */
/* ------------------------------------------------------------------ */

/* PartialAlignment (ebwt_search_util.h:34-88) */
typedef struct { uint16_t pos[3]; uint8_t chr[3]; } partial_t;
/* QueryMutation (ebwt_search_util.h:16-25) */
typedef struct { uint16_t pos; uint8_t oldBase, newBase; } qmut_t;
typedef struct { partial_t *v; size_t n, cap; } partial_list;

static void plist_push(partial_list *l, partial_t p) {
	if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 16; l->v = (partial_t *)realloc(l->v, l->cap * sizeof(partial_t)); }
	l->v[l->n++] = p;
}

/* ------------------------------------------------------------------ */
/* the read, in the five orientations of Read (read.h:42-277)          */
/* ------------------------------------------------------------------ */

typedef struct {
	uint32_t len;
	uint8_t *patFw, *patRc, *patFwRev, *patRcRev, *qual, *qualRev;
	uint32_t seed;
	uint32_t idx;
} read_t;

/* ------------------------------------------------------------------ */
/* GreedyDFSRangeSource                                                */
/* ------------------------------------------------------------------ */

typedef struct {
	const bto_index *ebwt;
	/* constructor parameters (ebwt_search_backtrack.h:28-79) */
	uint32_t qualThresh;
	uint32_t maxBts;
	uint32_t reportPartials;
	int reportExacts;
	int considerQuals;
	int halfAndHalf;
	int maqPenalty;
	/* per-query state */
	uint8_t *qry; const uint8_t *qual; uint32_t qlen, qryLen;
	int fw;                    /* EbwtSearchParams::_fw */
	uint32_t unrevOff, rev1Off, rev2Off, rev3Off, depth5, depth3;
	uint32_t *pairs; uint8_t *elims; uint8_t *chars; size_t scratchQlen;
	uint32_t *mms; uint8_t *refcs;    /* _mms/_refcs; refcs as codes 0..3 */
	const qmut_t *muts; uint32_t nmuts;
	uint32_t rnd;
	uint32_t numBts; int bailedOnBacktracks;
	partial_list partialsBuf;
	partial_list *partials;    /* destination manager (one read at a time) */
	sink_t *sink;
	uint32_t readIdx;
	bto_stats *st;
} dfs_t;

static uint8_t qual_at(const dfs_t *b, size_t off) { return phred_of(b->qual[off]); }
#define PAIR_TOP(p, d, c) ((p)[(size_t)(d) * 8 + (c)])
#define PAIR_BOT(p, d, c) ((p)[(size_t)(d) * 8 + (c) + 4])
#define PAIR_SPREAD(p, d, c) (PAIR_BOT(p, d, c) - PAIR_TOP(p, d, c))

/* setQuery (ebwt_search_backtrack.h:90-140) */
static void dfs_set_query(dfs_t *b, read_t *r, int fw) {
	b->fw = fw;
	if (b->ebwt->fw) { b->qry = fw ? r->patFw : r->patRc; b->qual = fw ? r->qual : r->qualRev; }
	else             { b->qry = fw ? r->patFwRev : r->patRcRev; b->qual = fw ? r->qualRev : r->qual; }
	b->qryLen = r->len;
	b->qlen = r->len;
	if ((size_t)r->len > b->scratchQlen) {
		size_t q = r->len;
		free(b->pairs); free(b->elims); free(b->chars); free(b->mms); free(b->refcs);
		b->pairs = (uint32_t *)malloc(q * q * 8 * sizeof(uint32_t));
		b->elims = (uint8_t *)calloc(q * q, 1);
		b->chars = (uint8_t *)calloc(q, 1);
		b->mms = (uint32_t *)calloc(q + 4, sizeof(uint32_t));
		b->refcs = (uint8_t *)calloc(q + 4, 1);
		b->scratchQlen = q;
	}
	b->rnd = r->seed;
	b->readIdx = r->idx;
}

/* setOffs (ebwt_search_backtrack.h:162-176) */
static void dfs_set_offs(dfs_t *b, uint32_t depth5, uint32_t depth3, uint32_t unrevOff,
                         uint32_t revOff1, uint32_t revOff2, uint32_t revOff3) {
	b->depth5 = depth5; b->depth3 = depth3;
	b->unrevOff = unrevOff; b->rev1Off = revOff1; b->rev2Off = revOff2; b->rev3Off = revOff3;
}

/* setQlen (ebwt_search_backtrack.h:217-220) */
static void dfs_set_qlen(dfs_t *b, uint32_t qlen) { b->qlen = b->qryLen < qlen ? b->qryLen : qlen; }

/* applyPartialMutations / undoPartialMutations (ebwt_search_backtrack.h:1368-1382, 1432-1446) */
static void dfs_apply_muts(dfs_t *b) { uint32_t i; for (i = 0; i < b->nmuts; i++) b->qry[b->muts[i].pos] = b->muts[i].newBase; }
static void dfs_undo_muts(dfs_t *b)  { uint32_t i; for (i = 0; i < b->nmuts; i++) b->qry[b->muts[i].pos] = b->muts[i].oldBase; }

/* setMuts (ebwt_search_backtrack.h:146-157) */
static void dfs_set_muts(dfs_t *b, const qmut_t *muts, uint32_t n) {
	if (b->muts != NULL) dfs_undo_muts(b);
	b->muts = muts; b->nmuts = muts ? n : 0;
	if (b->muts != NULL) dfs_apply_muts(b);
}

/* calcStratum (ebwt_search_backtrack.h:1164-1177) */
static int dfs_calc_stratum(const dfs_t *b, uint32_t stackDepth) {
	int stratum = 0; uint32_t i;
	for (i = 0; i < stackDepth; i++) if (b->mms[i] >= (b->qlen - b->rev3Off)) stratum++;
	return stratum;
}

/* Ebwt::reportChaseOne + Ebwt::report + EbwtSearchParams::reportHit
 * (ebwt.h:2693-2755, 2635-2682, 1288-1405) */
static int dfs_report_chase_one(dfs_t *b, uint32_t numMms, uint32_t row, uint32_t top, uint32_t bot,
                                int stratum, uint16_t cost) {
	const bto_index *ix = b->ebwt;
	uint32_t jumps, off, tidx, textoff, tlen, i, endrow;
	bto_hit h; bto_mm mm[1024];
	off = chase_impl(ix, row, &jumps, &endrow);
	b->st->lf += jumps; b->st->chase += jumps;
	if (endrow != ix->zOff) b->st->offs++;
	bto_joined_to_text_off(ix, b->qlen, off, &tidx, &textoff, &tlen);
	if (tidx == BTO_OFF_MASK) return 0;
	h.read = b->readIdx; h.tidx = tidx; h.toff = textoff; h.oms = bot - top - 1;
	h.cost = cost; h.fw = (uint8_t)b->fw; h.stratum = (uint8_t)stratum; h.nmm = numMms; h.mm_off = 0;
	for (i = 0; i < numMms && i < 1024; i++) {
		uint32_t p = b->mms[i];
		if (ix->fw != b->fw) p = b->qlen - p - 1; /* ebwt.h:1339-1350 */
		mm[i].pos = (uint16_t)p; mm[i].refc = b->refcs[i]; mm[i].pad = 0;
	}
	return sink_report(b->sink, &h, mm);
}

/* reportFullAlignment (ebwt_search_backtrack.h:1522-1565) */
static int dfs_report_full(dfs_t *b, uint32_t stackDepth, uint32_t top, uint32_t bot, int stratum, uint16_t cost) {
	uint32_t spread, r, i;
	if (stackDepth == 0 && !b->reportExacts) return 0;
	spread = bot - top;
	r = top + (bto_rand_next(&b->rnd) % spread);
	for (i = 0; i < spread; i++) {
		uint32_t ri = r + i;
		if (ri >= bot) ri -= spread;
		if (dfs_report_chase_one(b, stackDepth, ri, top, bot, stratum, cost)) return 1;
	}
	return 0;
}

/* reportPartial (ebwt_search_backtrack.h:1571-1655) */
static void dfs_report_partial(dfs_t *b, uint32_t stackDepth) {
	partial_t al; uint32_t k;
	al.pos[0] = al.pos[1] = al.pos[2] = 0xffff; al.chr[0] = al.chr[1] = al.chr[2] = 3;
	for (k = 0; k < stackDepth && k < 3; k++) {
		uint32_t ci = b->qlen - b->mms[k] - 1;
		al.pos[k] = (uint16_t)b->mms[k];
		al.chr[k] = b->chars[ci];
	}
	plist_push(&b->partialsBuf, al);
}

/* reportAlignment (ebwt_search_backtrack.h:1455-1513) */
static int dfs_report_alignment(dfs_t *b, uint32_t stackDepth, uint32_t top, uint32_t bot, uint16_t cost) {
	int stratum = 0;
	if (b->reportPartials) {
		if (stackDepth > 0) dfs_report_partial(b, stackDepth);
		return 0;
	}
	if (stackDepth > 0) stratum = dfs_calc_stratum(b, stackDepth);
	if (b->muts != NULL) {
		uint32_t i; int hit;
		dfs_undo_muts(b);
		/* promotePartialMutations (ebwt_search_backtrack.h:1389-1426) */
		for (i = 0; i < b->nmuts; i++) { b->mms[stackDepth + i] = b->muts[i].pos; b->refcs[stackDepth + i] = b->muts[i].newBase; }
		stratum += (int)b->nmuts;
		cost |= (uint16_t)(stratum << 14);
		hit = dfs_report_full(b, stackDepth + b->nmuts, top, bot, stratum, cost);
		dfs_apply_muts(b);
		return hit;
	}
	cost |= (uint16_t)(stratum << 14);
	return dfs_report_full(b, stackDepth, top, bot, stratum, cost);
}

/* hhCheckTop (ebwt_search_backtrack.h:1200-1275) */
static int dfs_hh_check_top(const dfs_t *b, uint32_t stackDepth, uint32_t d) {
	if (d == b->depth5) {
		if (b->rev3Off == b->rev2Off) { if (stackDepth == 0) return 0; }
		else { if (stackDepth < 1) return 0; }
	} else if (d == b->depth3) {
		if (b->rev3Off == b->rev2Off) { if (stackDepth < 2) return 0; }
		else {
			int lo = 0, hi = 0; uint32_t i;
			for (i = 0; i < stackDepth; i++) {
				uint32_t dd = b->qlen - b->mms[i] - 1;
				if (dd < b->depth5) hi++; else if (dd < b->depth3) lo++;
			}
			(void)hi;
			if (lo == 0) return 0;
		}
	}
	return 1;
}

/* the recursive frame: backtrack(stackDepth, depth, ...) (ebwt_search_backtrack.h:363-1091) */
static int dfs_frame(dfs_t *b, uint32_t stackDepth, uint32_t depth,
                     uint32_t unrevOff, uint32_t oneRevOff, uint32_t twoRevOff, uint32_t threeRevOff,
                     uint32_t top, uint32_t bot, uint32_t ham, uint32_t iham,
                     uint32_t *pairs, uint8_t *elims, int disableFtab) {
	const bto_index *ebwt = b->ebwt;
	locus_t ltop, lbot;
	uint32_t altNum = 0, eligibleNum = 0, eligibleSz = 0;
	uint32_t eli = 0; int elignore = 1; uint32_t eltop = 0, elbot = 0, elham = ham; int elcint = 0;
	uint8_t lowAltQual = 0xff;
	uint32_t d = depth;
	uint32_t cur = b->qlen - d - 1;
	memset(&ltop, 0, sizeof ltop); memset(&lbot, 0, sizeof lbot);
	if (stackDepth > 0) b->st->backtracks++;
	if (top != 0 || bot != 0) locus_from_topbot(ebwt, top, bot, &ltop, &lbot);
	if (b->halfAndHalf) {
		if (b->maxBts > 0 && b->numBts == b->maxBts) { b->bailedOnBacktracks = 1; return 0; }
		b->numBts++;
	}
	while (cur < b->qlen) {
		int curIsEligible = 0, curOverridesEligible = 0, curIsAlternative;
		int c; uint8_t q;
		int backtrackDespiteMatch = 0, reportedPartial = 0, invalidExact = 0, mustBacktrack = 0, invalidHalfAndHalf = 0;
		if (b->halfAndHalf && !dfs_hh_check_top(b, stackDepth, d)) return 0;
		c = (int)b->qry[cur];
		q = qual_at(b, cur);
		curIsAlternative = (d >= unrevOff) &&
			(!b->considerQuals || (ham + mm_penalty(b->maqPenalty, q) <= b->qualThresh));
		if (curIsAlternative) {
			if (b->considerQuals) {
				if (q < lowAltQual) { curIsEligible = 1; curOverridesEligible = 1; }
				else if (q == lowAltQual) curIsEligible = 1;
			} else curIsEligible = 1;
		}
		if (c == 4 && d > 0) top = bot = 1;
		if (top == 0 && bot == 0) {
			/* first quartet from fchr[] (ebwt_search_backtrack.h:531-543) */
			pairs[0 + 0] = ebwt->fchr[0];
			pairs[0 + 4] = pairs[1 + 0] = ebwt->fchr[1];
			pairs[1 + 4] = pairs[2 + 0] = ebwt->fchr[2];
			pairs[2 + 4] = pairs[3 + 0] = ebwt->fchr[3];
			pairs[3 + 4] = ebwt->fchr[4];
			if (c < 4) { top = PAIR_TOP(pairs, d, c); bot = PAIR_BOT(pairs, d, c); }
		} else if (curIsAlternative) {
			count_side_ex(ebwt, &ltop, &pairs[d * 8]);      /* mapLFEx (ebwt.h:2334-2380) */
			count_side_ex(ebwt, &lbot, &pairs[d * 8 + 4]);
			b->st->lfex++;
			if (c < 4) { top = PAIR_TOP(pairs, d, c); bot = PAIR_BOT(pairs, d, c); }
		} else {
			if (c < 4) {
				if (top + 1 == bot) {
					/* mapLF1 (ebwt.h:2494-2524) */
					if (row_l(ebwt, &ltop) != c || top == ebwt->zOff) top = BTO_OFF_MASK;
					else top = count_side(ebwt, &ltop, c);
					bot = top;
					if (bot != BTO_OFF_MASK) bot++;
					b->st->lf++;
				} else {
					top = count_side(ebwt, &ltop, c); bot = count_side(ebwt, &lbot, c);
					b->st->lf += 2;
				}
			}
		}
		if (top != bot) locus_from_topbot(ebwt, top, bot, &ltop, &lbot);
		/* eliminate() (ebwt_search_backtrack.h:1183-1192) */
		elims[d] = (c < 4) ? (uint8_t)(1 << c) : 0;
		if (curIsAlternative) {
			int i;
			for (i = 0; i < 4; i++) {
				uint32_t spread;
				if (i == c) continue;
				spread = PAIR_SPREAD(pairs, d, i);
				if (spread == 0) elims[d] |= (uint8_t)(1 << i);
				if (spread > 0 && ((elims[d] & (1 << i)) == 0)) {
					if (curIsEligible) {
						if (curOverridesEligible) {
							lowAltQual = q; eligibleNum = 0; eligibleSz = 0; curOverridesEligible = 0;
							eli = d; eltop = PAIR_TOP(pairs, d, i); elbot = PAIR_BOT(pairs, d, i);
							elham = mm_penalty(b->maqPenalty, q); elcint = i; elignore = 0;
						}
						eligibleSz += spread; eligibleNum++;
					}
					altNum++;
				}
			}
		}
		if (cur == 0 && top < bot && stackDepth < b->reportPartials && b->reportPartials > 0) {
			if (altNum > 0) backtrackDespiteMatch = 1;
			if (stackDepth > 0) { dfs_report_partial(b, stackDepth); reportedPartial = 1; }
		}
		if (cur == 0 && stackDepth == 0 && bot > top && !b->reportExacts) { invalidExact = 1; backtrackDespiteMatch = 1; }
		if (b->halfAndHalf) {
			if ((d == (b->depth5 - 1)) && top < bot) {
				invalidHalfAndHalf = (stackDepth == 0);
				if (stackDepth == 0 && altNum > 0) { backtrackDespiteMatch = 1; mustBacktrack = 1; }
				else if (stackDepth == 0) return 0;
			} else if ((d == (b->depth3 - 1)) && top < bot) {
				uint32_t lo = 0, hi = 0, i;
				for (i = 0; i < stackDepth; i++) {
					uint32_t dd = b->qlen - b->mms[i] - 1;
					if (dd < b->depth5) hi++; else if (dd < b->depth3) lo++;
				}
				invalidHalfAndHalf = (lo == 0 || hi == 0);
				if ((stackDepth < 2 || invalidHalfAndHalf) && altNum > 0) { mustBacktrack = 1; backtrackDespiteMatch = 1; }
				else if (stackDepth < 2) return 0;
			}
		}
		if (cur == 0 && bot > top && !invalidHalfAndHalf && !invalidExact && !reportedPartial) {
			if (!dfs_report_alignment(b, stackDepth, top, bot, (uint16_t)ham)) top = bot;
			else return 1;
		}
		/* mismatch with alternatives (ebwt_search_backtrack.h:743-1065) */
		while ((top == bot || backtrackDespiteMatch) && altNum > 0) {
			size_t i = d, j = 0;
			uint32_t bttop = 0, btbot = 0, btham = ham, icur;
			int btcint = 0, ret;
			uint32_t btUnrevOff, btOneRevOff, btTwoRevOff, btThreeRevOff;
			uint32_t *newPairs; uint8_t *newElims;
			if (eligibleNum > 1 || elignore) {
				for (; i >= depth; i--) {
					uint8_t qi;
					icur = (uint32_t)(b->qlen - i - 1);
					qi = qual_at(b, icur);
					if ((qi == lowAltQual || !b->considerQuals) && elims[i] != 15) {
						uint32_t posSz = 0, r;
						for (j = 0; j < 4; j++) if ((elims[i] & (1 << j)) == 0) posSz += PAIR_SPREAD(pairs, i, j);
						r = bto_rand_next(&b->rnd) % posSz;
						for (j = 0; j < 4; j++) {
							if ((elims[i] & (1 << j)) == 0) {
								uint32_t spread = PAIR_SPREAD(pairs, i, j);
								if (r < spread) {
									bttop = PAIR_TOP(pairs, i, j); btbot = PAIR_BOT(pairs, i, j);
									btham += mm_penalty(b->maqPenalty, qi);
									btcint = (int)j;
									break;
								}
								r -= spread;
							}
						}
						break;
					}
				}
			} else {
				i = eli; bttop = eltop; btbot = elbot; btham += elham; j = (size_t)elcint; btcint = elcint;
			}
			icur = (uint32_t)(b->qlen - i - 1);
			newPairs = pairs + ((size_t)b->qlen * 8);
			newElims = elims + b->qlen;
			btUnrevOff = unrevOff; btOneRevOff = oneRevOff; btTwoRevOff = twoRevOff; btThreeRevOff = threeRevOff;
			if (i < oneRevOff) { btUnrevOff = oneRevOff; btOneRevOff = twoRevOff; btTwoRevOff = threeRevOff; }
			else if (i < twoRevOff) { btOneRevOff = twoRevOff; btTwoRevOff = threeRevOff; }
			else if (i < threeRevOff) { btTwoRevOff = threeRevOff; }
			b->mms[stackDepth] = icur;
			b->refcs[stackDepth] = (uint8_t)btcint;
			b->chars[i] = (uint8_t)btcint;
			if (i + 1 == b->qlen) {
				ret = dfs_report_alignment(b, stackDepth + 1, bttop, btbot, (uint16_t)btham);
			} else if (b->halfAndHalf && !disableFtab && b->rev2Off == b->rev3Off &&
			           i + 1 < (uint32_t)ebwt->ftabChars && (uint32_t)ebwt->ftabChars <= b->depth5) {
				/* ftab re-jump with the substituted character (ebwt_search_backtrack.h:908-952) */
				int ftabChars = ebwt->ftabChars, jj;
				uint32_t ftabOff = b->qry[b->qlen - ftabChars], ftabTop, ftabBot;
				for (jj = ftabChars - 1; jj > 0; jj--) {
					ftabOff <<= 2;
					if (b->qlen - jj == icur) ftabOff |= (uint32_t)btcint;
					else ftabOff |= b->qry[b->qlen - jj];
				}
				ftabTop = bto_ftab_hi(ebwt, ftabOff);
				ftabBot = bto_ftab_lo(ebwt, ftabOff + 1);
				b->st->ftab++;
				if (ftabTop == ftabBot) ret = 0;
				else ret = dfs_frame(b, stackDepth + 1, (uint32_t)ebwt->ftabChars, btUnrevOff, btOneRevOff, btTwoRevOff, btThreeRevOff,
				                     ftabTop, ftabBot, btham, iham, newPairs, newElims, 0);
			} else {
				ret = dfs_frame(b, stackDepth + 1, (uint32_t)i + 1, btUnrevOff, btOneRevOff, btTwoRevOff, btThreeRevOff,
				                bttop, btbot, btham, iham, newPairs, newElims, 0);
			}
			if (ret) return 1;
			if (b->bailedOnBacktracks || (b->halfAndHalf && (b->maxBts > 0) && (b->numBts >= b->maxBts))) {
				b->bailedOnBacktracks = 1; return 0;
			}
			b->chars[i] = b->qry[icur];
			elims[i] |= (uint8_t)(1 << j);
			eligibleSz -= (btbot - bttop);
			eligibleNum--;
			elignore = 1;
			altNum--;
			if (altNum == 0) return 0;
			else if (eligibleNum == 0 && b->considerQuals) {
				/* rescan the frame for the next-lowest quality (ebwt_search_backtrack.h:1004-1058) */
				size_t k;
				lowAltQual = 0xff;
				for (k = d; k >= depth && k <= b->qlen; k--) {
					size_t kcur = b->qlen - k - 1;
					uint8_t kq = qual_at(b, kcur);
					int kAlt, kOverrides = 0;
					if (k < unrevOff) break;
					kAlt = (ham + mm_penalty(b->maqPenalty, kq) <= b->qualThresh);
					if (kAlt) {
						if (kq < lowAltQual) kOverrides = 1;
						if (kq <= lowAltQual) {
							int l;
							for (l = 0; l < 4; l++) {
								if ((elims[k] & (1 << l)) == 0) {
									uint32_t spread = PAIR_SPREAD(pairs, k, l);
									if (kOverrides) {
										lowAltQual = kq; kOverrides = 0;
										eligibleNum = 0; eligibleSz = 0;
										eli = (uint32_t)k; eltop = PAIR_TOP(pairs, k, l); elbot = PAIR_BOT(pairs, k, l);
										elham = mm_penalty(b->maqPenalty, kq); elcint = l; elignore = 0;
									}
									eligibleNum++;
									eligibleSz += spread;
								}
							}
						}
					}
				}
			}
		}
		if (mustBacktrack || invalidHalfAndHalf || invalidExact) return 0;
		if (top == bot && altNum == 0) return 0;
		b->chars[d] = b->qry[cur];
		d++; cur--;
	}
	if (stackDepth >= b->reportPartials) return dfs_report_alignment(b, stackDepth, top, bot, (uint16_t)ham);
	return 0;
}

/* tallyNs (ebwt_search_backtrack.h:1308-1341) */
static int dfs_tally_ns(const dfs_t *b, int *nsInSeed, int *nsInFtab) {
	size_t i; int ftabChars = b->ebwt->ftabChars;
	for (i = 0; i < b->rev3Off; i++) {
		if (b->qry[b->qlen - i - 1] == 4) {
			(*nsInSeed)++;
			if (*nsInSeed == 1) { if (i < b->unrevOff) return 0; }
			else if (*nsInSeed == 2) { if (i < b->rev1Off) return 0; }
			else if (*nsInSeed == 3) { if (i < b->rev2Off) return 0; }
			else return 0;
		}
	}
	for (i = 0; i < (size_t)ftabChars && i < b->qlen; i++) if (b->qry[b->qlen - i - 1] == 4) (*nsInFtab)++;
	return 1;
}

/* backtrack(depth, top, bot, iham, disableFtab) (ebwt_search_backtrack.h:333-353) */
static int dfs_backtrack_from(dfs_t *b, uint32_t depth, uint32_t top, uint32_t bot, uint32_t iham, int disableFtab) {
	int done;
	b->bailedOnBacktracks = 0;
	done = dfs_frame(b, 0, depth, b->unrevOff, b->rev1Off, b->rev2Off, b->rev3Off, top, bot, iham, iham,
	                 b->pairs, b->elims, disableFtab);
	b->numBts = 0;
	b->bailedOnBacktracks = 0;
	return done;
}

/* finalize (ebwt_search_backtrack.h:303-324) */
static int dfs_finalize(dfs_t *b) {
	int ret = 0;
	if (b->reportPartials > 0 && b->partialsBuf.n > 0) {
		size_t i;
		for (i = 0; i < b->partialsBuf.n; i++) plist_push(b->partials, b->partialsBuf.v[i]);
		b->partialsBuf.n = 0;
		ret = 1;
	}
	return ret;
}

/* backtrack(ham) (ebwt_search_backtrack.h:237-297) */
static int dfs_backtrack(dfs_t *b, uint32_t ham) {
	const bto_index *ebwt = b->ebwt;
	int ftabChars = ebwt->ftabChars, nsInSeed = 0, nsInFtab = 0, ret;
	uint32_t m;
	if (!dfs_tally_ns(b, &nsInSeed, &nsInFtab)) return 0;
	m = b->unrevOff < b->qlen ? b->unrevOff : b->qlen;
	if (nsInFtab == 0 && m >= (uint32_t)ftabChars) {
		/* calcFtabOff (ebwt_search_backtrack.h:1348-1362) */
		uint32_t ftabOff = b->qry[b->qlen - ftabChars], top, bot; int i;
		for (i = ftabChars - 1; i > 0; i--) { ftabOff <<= 2; ftabOff |= b->qry[b->qlen - i]; }
		top = bto_ftab_hi(ebwt, ftabOff);
		bot = bto_ftab_lo(ebwt, ftabOff + 1);
		b->st->ftab++;
		if (b->qlen == (uint32_t)ftabChars && bot > top) {
			if (b->reportPartials > 0) ret = dfs_backtrack_from(b, 0, 0, 0, ham, nsInFtab > 0);
			else ret = dfs_report_alignment(b, 0, top, bot, (uint16_t)ham);
		} else if (bot > top) {
			ret = dfs_backtrack_from(b, (uint32_t)ftabChars, top, bot, ham, nsInFtab > 0);
		} else ret = 0;
	} else {
		ret = dfs_backtrack_from(b, 0, 0, 0, ham, nsInFtab > 0);
	}
	if (dfs_finalize(b)) ret = 1;
	return ret;
}

static void dfs_init(dfs_t *b, const bto_index *ebwt, sink_t *sink, bto_stats *st, uint32_t qualThresh, uint32_t maxBts,
                     uint32_t reportPartials, partial_list *partials, int considerQuals, int halfAndHalf, int maqPenalty) {
	memset(b, 0, sizeof(*b));
	b->ebwt = ebwt; b->sink = sink; b->st = st;
	b->qualThresh = qualThresh; b->maxBts = maxBts; b->reportPartials = reportPartials; b->partials = partials;
	b->reportExacts = 1; b->considerQuals = considerQuals; b->halfAndHalf = halfAndHalf; b->maqPenalty = maqPenalty;
}
static void dfs_destroy(dfs_t *b) {
	free(b->pairs); free(b->elims); free(b->chars); free(b->mms); free(b->refcs); free(b->partialsBuf.v);
}

/* ------------------------------------------------------------------ */
/* per-policy phase loops                                              */
/* ------------------------------------------------------------------ */

typedef struct {
	const bto_index *fwix, *bwix;
	const bto_policy *pol;
	sink_t sink;
	bto_stats *st;
	dfs_t bt[9];
	partial_list pamRc, pamFw;
} worker_t;

/* search_exact.c (exactSearchWorker ebwt_search.cpp:1130-1219) */
static void search_v0(worker_t *w, read_t *r) {
	dfs_t *bt = &w->bt[0];
	uint32_t plen = r->len;
	if (!w->pol->nofw) {
		dfs_set_query(bt, r, 1);
		dfs_set_offs(bt, 0, 0, plen, plen, plen, plen);
		if (dfs_backtrack(bt, 0)) return;
	}
	if (!w->pol->norc) {
		dfs_set_query(bt, r, 0);
		dfs_set_offs(bt, 0, 0, plen, plen, plen, plen);
		dfs_backtrack(bt, 0);
	}
}

/* search_1mm_phase1.c + search_1mm_phase2.c (mismatchSearchWorkerFull ebwt_search.cpp:1606-1700) */
static void search_v1(worker_t *w, read_t *r) {
	dfs_t *bt = &w->bt[0];
	uint32_t s = r->len, s3 = s >> 1, s5 = (s >> 1) + (s & 1);
	int nofw = w->pol->nofw, norc = w->pol->norc;
	bt->ebwt = w->fwix; bt->reportExacts = 1;
	if (!nofw) { dfs_set_query(bt, r, 1); dfs_set_offs(bt, 0, 0, s, s, s, s); if (dfs_backtrack(bt, 0)) return; }
	if (!norc) { dfs_set_query(bt, r, 0); dfs_set_offs(bt, 0, 0, s, s, s, s); if (dfs_backtrack(bt, 0)) return; }
	bt->reportExacts = 0;
	if (!norc) { dfs_set_query(bt, r, 0); dfs_set_offs(bt, 0, 0, s5, s, s, s); if (dfs_backtrack(bt, 0)) return; }
	if (!nofw) { dfs_set_query(bt, r, 1); dfs_set_offs(bt, 0, 0, s5, s, s, s); if (dfs_backtrack(bt, 0)) return; }
	bt->ebwt = w->bwix; bt->reportExacts = 0;
	if (!norc) { dfs_set_query(bt, r, 0); dfs_set_offs(bt, 0, 0, s3, s, s, s); if (dfs_backtrack(bt, 0)) return; }
	if (!nofw) { dfs_set_query(bt, r, 1); dfs_set_offs(bt, 0, 0, s3, s, s, s); if (dfs_backtrack(bt, 0)) return; }
}

/* search_23mm_phase{1,2,3}.c with two=true (twoOrThreeMismatchSearchWorkerFull ebwt_search.cpp:2056-2195) */
static void search_v2(worker_t *w, read_t *r) {
	dfs_t *btr1 = &w->bt[0], *bt2 = &w->bt[1], *bt3 = &w->bt[2], *bthh3 = &w->bt[3];
	uint32_t plen = r->len, s = plen, s3 = s >> 1, s5 = (s >> 1) + (s & 1);
	int nofw = w->pol->nofw, norc = w->pol->norc;
	const int two = 1;
	btr1->reportExacts = 1;
	if (!nofw) { dfs_set_query(btr1, r, 1); dfs_set_offs(btr1, 0, 0, plen, plen, plen, plen); if (dfs_backtrack(btr1, 0)) return; }
	if (!norc) { dfs_set_query(btr1, r, 0); dfs_set_offs(btr1, 0, 0, s5, s5, two ? s : s5, s); if (dfs_backtrack(btr1, 0)) return; }
	bt2->reportExacts = 0;
	if (!nofw) { dfs_set_query(bt2, r, 1); dfs_set_offs(bt2, 0, 0, s5, s5, two ? s : s5, s); if (dfs_backtrack(bt2, 0)) return; }
	if (!norc) { dfs_set_query(bt2, r, 0); dfs_set_offs(bt2, 0, 0, s3, s3, two ? s : s3, s); if (dfs_backtrack(bt2, 0)) return; }
	if (!nofw) {
		bt3->reportExacts = 0;
		dfs_set_query(bt3, r, 1); dfs_set_offs(bt3, 0, 0, s3, s3, two ? s : s3, s);
		if (dfs_backtrack(bt3, 0)) return;
		dfs_set_query(bthh3, r, 1); dfs_set_offs(bthh3, s3, s, 0, two ? s3 : 0, two ? s : s3, s);
		if (dfs_backtrack(bthh3, 0)) return;
	}
	if (!norc) {
		dfs_set_query(bthh3, r, 0); dfs_set_offs(bthh3, s5, s, 0, two ? s5 : 0, two ? s : s5, s);
		if (dfs_backtrack(bthh3, 0)) return;
	}
}

/* PartialAlignmentManager::toMutsString (ebwt_search_util.h:299-357) */
static uint8_t to_muts(const partial_t *pal, const uint8_t *seq, const uint8_t *quals, uint32_t plen,
                       qmut_t *muts, uint32_t *nmuts, int maq) {
	uint8_t oldQuals = 0; int k; *nmuts = 0;
	for (k = 0; k < 3; k++) {
		uint16_t tpos;
		if (pal->pos[k] == 0xffff) break;
		tpos = (uint16_t)(plen - 1 - pal->pos[k]);
		oldQuals = (uint8_t)(oldQuals + mm_penalty(maq, phred_of(quals[tpos])));
		muts[*nmuts].pos = tpos; muts[*nmuts].oldBase = seq[tpos]; muts[*nmuts].newBase = pal->chr[k];
		(*nmuts)++;
	}
	return oldQuals;
}

/* search_seeded_phase{1,2,3,4}.c (seededQualSearchWorkerFull ebwt_search.cpp:2378-2603).
 * Returns 1 if the read was skipped by the phase-1 filter. */
static void search_n(worker_t *w, read_t *r) {
	dfs_t *btf1 = &w->bt[0], *bt1 = &w->bt[1], *btf2 = &w->bt[2], *btr2 = &w->bt[3], *btf3 = &w->bt[4],
	      *btr3 = &w->bt[5], *btr23 = &w->bt[6], *btf4 = &w->bt[7], *btf24 = &w->bt[8];
	const bto_policy *pol = w->pol;
	int seedMms = pol->mms, nofw = pol->nofw, norc = pol->norc, maq = pol->maqRound;
	uint32_t plen = r->len, s = (uint32_t)pol->seedLen, s3 = s >> 1, s5 = (s >> 1) + (s & 1);
	uint32_t qs = plen < s ? plen : s, qs3 = qs >> 1, qs5 = (qs >> 1) + (qs & 1);
	uint32_t S = (qs < s) ? qs : s, S3 = (qs < s) ? qs3 : s3, S5 = (qs < s) ? qs5 : s5;
	qmut_t muts[4]; uint32_t nmuts; size_t i;
	int done = 0;
	/* phase 1 */
	btf1->reportExacts = 1; bt1->reportExacts = 1;
	if (plen < 4) done = 1;
	else {
		uint32_t slen = plen < (uint32_t)pol->seedLen ? plen : (uint32_t)pol->seedLen; int ns = 0; uint32_t k;
		for (k = 0; k < slen; k++) if (r->patFw[k] == 4) { if (++ns > seedMms) { done = 1; break; } }
	}
	if (done) return;
	if (!nofw) {
		dfs_set_query(btf1, r, 1); dfs_set_offs(btf1, 0, plen, plen, plen, plen, plen);
		if (dfs_backtrack(btf1, 0)) return;
	}
	if (!norc) {
		dfs_set_offs(bt1, 0, 0, (seedMms > 0) ? S5 : S, (seedMms > 1) ? S5 : S, (seedMms > 2) ? S5 : S, (seedMms > 3) ? S5 : S);
		dfs_set_query(bt1, r, 0);
		if (dfs_backtrack(bt1, 0)) return;
	}
	/* phase 2 */
	if (!nofw) {
		btf2->reportExacts = 0; btr2->reportExacts = 0;
		dfs_set_query(btf2, r, 1);
		dfs_set_offs(btf2, 0, 0, (seedMms > 0) ? S5 : S, (seedMms > 1) ? S5 : S, (seedMms > 2) ? S5 : S, (seedMms > 3) ? S5 : S);
		if (dfs_backtrack(btf2, 0)) return;
	}
	if (seedMms == 0) return;
	if (!norc) {
		dfs_set_offs(btr2, 0, 0, S3, (seedMms > 1) ? S3 : S, (seedMms > 2) ? S3 : S, (seedMms > 3) ? S3 : S);
		dfs_set_query(btr2, r, 0);
		dfs_set_qlen(btr2, s);
		dfs_backtrack(btr2, 0);
	}
	/* phase 3 */
	if (!norc) {
		btr3->reportExacts = 1;
		dfs_set_query(btr3, r, 0);
		done = 0;
		if (w->pamRc.n > 0) {
			dfs_set_offs(btr3, 0, 0, S, S, S, S);
			for (i = 0; i < w->pamRc.n; i++) {
				uint8_t oldQuals = to_muts(&w->pamRc.v[i], r->patRc, r->qualRev, plen, muts, &nmuts, maq);
				dfs_set_muts(btr3, muts, nmuts);
				done = dfs_backtrack(btr3, oldQuals);
				dfs_set_muts(btr3, NULL, 0);
				if (done) break;
			}
		}
		w->pamRc.n = 0;
		if (done) return;
		if (seedMms >= 2) {
			dfs_set_query(btr23, r, 0);
			dfs_set_offs(btr23, S5, S, 0, (seedMms <= 2) ? S5 : 0, (seedMms < 3) ? S : S5, S);
			if (dfs_backtrack(btr23, 0)) return;
		}
	}
	if (nofw) return;
	dfs_set_query(btf3, r, 1);
	dfs_set_qlen(btf3, (uint32_t)pol->seedLen);
	dfs_set_offs(btf3, 0, 0, S3, (seedMms > 1) ? S3 : S, (seedMms > 2) ? S3 : S, (seedMms > 3) ? S3 : S);
	dfs_backtrack(btf3, 0);
	/* phase 4 */
	btf4->reportExacts = 1;
	dfs_set_query(btf4, r, 1);
	done = 0;
	if (w->pamFw.n > 0) {
		dfs_set_offs(btf4, 0, 0, S, S, S, S);
		for (i = 0; i < w->pamFw.n; i++) {
			uint8_t oldQuals = to_muts(&w->pamFw.v[i], r->patFwRev, r->qualRev, plen, muts, &nmuts, maq);
			dfs_set_muts(btf4, muts, nmuts);
			done = dfs_backtrack(btf4, oldQuals);
			dfs_set_muts(btf4, NULL, 0);
			if (done) break;
		}
	}
	w->pamFw.n = 0;
	if (done) return;
	if (seedMms >= 2) {
		dfs_set_query(btf24, r, 1);
		dfs_set_offs(btf24, S5, S, 0, (seedMms <= 2) ? S5 : 0, (seedMms < 3) ? S : S5, S);
		if (dfs_backtrack(btf24, 0)) return;
	}
}

/* ------------------------------------------------------------------ */
/* batch driver                                                        */
/* ------------------------------------------------------------------ */

bto_result *bto_result_new(size_t nreads) {
	bto_result *r = (bto_result *)calloc(1, sizeof(*r));
	r->nhits_per_read = (uint32_t *)calloc(nreads ? nreads : 1, sizeof(uint32_t));
	r->maxed = (uint8_t *)calloc(nreads ? nreads : 1, 1);
	return r;
}
void bto_result_free(bto_result *r) {
	if (!r) return;
	free(r->hits); free(r->mms); free(r->nhits_per_read); free(r->maxed); free(r);
}

static void result_push(bto_result *out, const bto_hit *h, const bto_mm *mms) {
	if (out->nhits == out->cap_hits) { out->cap_hits = out->cap_hits ? out->cap_hits * 2 : 1024; out->hits = (bto_hit *)realloc(out->hits, out->cap_hits * sizeof(bto_hit)); }
	if (out->nmms + h->nmm > out->cap_mms) { out->cap_mms = (out->cap_mms + h->nmm) * 2 + 1024; out->mms = (bto_mm *)realloc(out->mms, out->cap_mms * sizeof(bto_mm)); }
	out->hits[out->nhits] = *h;
	out->hits[out->nhits].mm_off = (uint32_t)out->nmms;
	memcpy(out->mms + out->nmms, mms, h->nmm * sizeof(bto_mm));
	out->nmms += h->nmm; out->nhits++;
}

int bto_align(const bto_index *fwix, const bto_index *bwix, const bto_policy *pol,
              size_t nreads, const uint8_t *seq, const uint8_t *qual,
              const uint64_t *offs, const uint32_t *seeds, bto_result *out) {
	worker_t w; size_t ri, maxlen = 0, k; read_t r;
	uint32_t qt, mb; int maq;
	if (pol->mode == 0 && (pol->mms < 0 || pol->mms > 2)) return 1;   /* -v 3 is the stateful path */
	if (pol->mode == 1 && (pol->mms < 0 || pol->mms > 3)) return 1;
	if ((pol->mode == 1 || pol->mms > 0) && bwix == NULL) return 2;
	memset(&w, 0, sizeof w);
	w.fwix = fwix; w.bwix = bwix; w.pol = pol; w.st = &out->stats;
	w.sink.n = pol->allHits ? 0xffffffffu : pol->khits;
	w.sink.max = pol->mhits;
	qt = (uint32_t)pol->qualThresh; mb = (uint32_t)pol->maxBts; maq = pol->maqRound;
	if (pol->mode == 0) {
		if (pol->mms == 0) {
			dfs_init(&w.bt[0], fwix, &w.sink, w.st, 0xffffffffu, 0xffffffffu, 0, NULL, 0, 0, 1);
		} else if (pol->mms == 1) {
			dfs_init(&w.bt[0], fwix, &w.sink, w.st, 0xffffffffu, 0xffffffffu, 0, NULL, 0, 0, 1);
		} else {
			dfs_init(&w.bt[0], fwix, &w.sink, w.st, 0xffffffffu, 0xffffffffu, 0, NULL, 0, 0, 1);
			dfs_init(&w.bt[1], bwix, &w.sink, w.st, 0xffffffffu, 0xffffffffu, 0, NULL, 0, 0, 1);
			dfs_init(&w.bt[2], fwix, &w.sink, w.st, 0xffffffffu, 0xffffffffu, 0, NULL, 0, 0, 1);
			dfs_init(&w.bt[3], fwix, &w.sink, w.st, 0xffffffffu, 0xffffffffu, 0, NULL, 0, 1, 1);
		}
	} else {
		/* the nine objects of seededQualSearchWorkerFull (ebwt_search.cpp:2413-2539) */
		dfs_init(&w.bt[0], fwix, &w.sink, w.st, qt, mb, 0, NULL, 0, 0, 1);                       /* btf1  */
		dfs_init(&w.bt[1], fwix, &w.sink, w.st, qt, mb, 0, NULL, 1, 0, maq);                     /* bt1   */
		dfs_init(&w.bt[2], bwix, &w.sink, w.st, qt, mb, 0, NULL, 1, 0, maq);                     /* btf2  */
		dfs_init(&w.bt[3], bwix, &w.sink, w.st, qt, mb, (uint32_t)pol->mms, &w.pamRc, 1, 0, maq);/* btr2  */
		dfs_init(&w.bt[4], fwix, &w.sink, w.st, qt, mb, (uint32_t)pol->mms, &w.pamFw, 1, 0, maq);/* btf3  */
		dfs_init(&w.bt[5], fwix, &w.sink, w.st, qt, mb, 0, NULL, 1, 0, maq);                     /* btr3  */
		dfs_init(&w.bt[6], fwix, &w.sink, w.st, qt, mb, 0, NULL, 1, 1, maq);                     /* btr23 */
		dfs_init(&w.bt[7], bwix, &w.sink, w.st, qt, mb, 0, NULL, 1, 0, maq);                     /* btf4  */
		dfs_init(&w.bt[8], bwix, &w.sink, w.st, qt, mb, 0, NULL, 1, 1, maq);                     /* btf24 */
	}
	for (ri = 0; ri < nreads; ri++) { size_t l = (size_t)(offs[ri + 1] - offs[ri]); if (l > maxlen) maxlen = l; }
	memset(&r, 0, sizeof r);
	r.patFw = (uint8_t *)malloc(maxlen + 1); r.patRc = (uint8_t *)malloc(maxlen + 1);
	r.patFwRev = (uint8_t *)malloc(maxlen + 1); r.patRcRev = (uint8_t *)malloc(maxlen + 1);
	r.qual = (uint8_t *)malloc(maxlen + 1); r.qualRev = (uint8_t *)malloc(maxlen + 1);
	for (ri = 0; ri < nreads; ri++) {
		uint32_t len = (uint32_t)(offs[ri + 1] - offs[ri]);
		const uint8_t *s = seq + offs[ri], *q = qual + offs[ri];
		uint32_t ret; int maxed, unal;
		r.len = len; r.seed = seeds[ri]; r.idx = (uint32_t)ri;
		for (k = 0; k < len; k++) {
			uint8_t c = s[k], rc = s[len - 1 - k]; /* Read::constructRevComps/constructReverses (read.h:118-132) */
			r.patFw[k] = c; r.qual[k] = q[k];
			r.patRc[k] = rc < 4 ? (uint8_t)(rc ^ 3) : 4;
			r.patFwRev[k] = s[len - 1 - k];
			r.patRcRev[k] = c < 4 ? (uint8_t)(c ^ 3) : 4;
			r.qualRev[k] = q[len - 1 - k];
		}
		w.sink.hitsForThisRead = 0; w.sink.nbuf = 0; w.sink.nmm = 0;
		w.pamRc.n = 0; w.pamFw.n = 0;
		if (len > 0) {
			if (pol->mode == 0) {
				if (pol->mms == 0) search_v0(&w, &r);
				else if (pol->mms == 1) search_v1(&w, &r);
				else search_v2(&w, &r);
			} else search_n(&w, &r);
		}
		/* HitSinkPerThread::finishRead (hit.h:741-786) */
		ret = w.sink.hitsForThisRead;
		maxed = ret > w.sink.max; unal = (ret == 0);
		if (maxed) { out->maxed[ri] = 1; out->counters[2]++; }
		else if (unal) { out->counters[1]++; }
		else {
			size_t nb = w.sink.nbuf, j;
			if (nb > w.sink.n) nb = w.sink.n;
			for (j = 0; j < nb; j++) result_push(out, &w.sink.buf[j], w.sink.mmbuf + w.sink.buf[j].mm_off);
			out->nhits_per_read[ri] = (uint32_t)nb;
			out->counters[0]++; out->counters[3] += nb;
		}
	}
	free(r.patFw); free(r.patRc); free(r.patFwRev); free(r.patRcRev); free(r.qual); free(r.qualRev);
	for (k = 0; k < 9; k++) dfs_destroy(&w.bt[k]);
	free(w.sink.buf); free(w.sink.mmbuf); free(w.pamRc.v); free(w.pamFw.v);
	return 0;
}
