/*
 * bt_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the Bowtie 1.3.1 FM-index backward-search hot path
 * (non-stateful / "greedy DFS" path: -v 0/1/2 and -n 0..3 without --best), used by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER for
 * the CUDA kernels.  The product (bowtie_b200/) never includes, links or calls this.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference checkout).  Parity of this restatement is pinned against the reference's
 * own binary (oracle/_ref/bowtie-align-s, built by oracle/Makefile) on the golden
 * vectors of SURVEY.md §8c — see tests/test_oracle_vs_reference.py.
 */
#ifndef BT_ORACLE_H_
#define BT_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BTO_OFF_MASK 0xffffffffu

/* One Ebwt (forward or mirror index); fields follow Ebwt / EbwtParams (ebwt.h:116-321, 1208-1236). */
typedef struct bto_index {
	uint32_t len, bwtLen;
	int32_t  lineRate, linesPerSide, offRate, ftabChars;
	int      entireRev;
	int      fw;                 /* 1: forward index, 0: mirror (".rev") index (Ebwt::_fw) */
	uint32_t offMask;
	uint32_t sideSz, sideBwtSz, sideBwtLen, numSides, ebwtTotLen;
	uint32_t ftabLen, eftabLen, offsLen;
	uint32_t nPat, nFrag;
	uint32_t *plen, *rstarts;
	uint8_t  *ebwt;
	uint32_t zOff, zEbwtByteOff;
	int      zEbwtBpOff;
	uint32_t fchr[5];
	uint32_t *ftab, *eftab, *offs;
	char     **refnames;
	uint32_t nRefnames;
} bto_index;

/* Search policy: the option globals of ebwt_search.cpp:153-253 that reach the hot path. */
typedef struct bto_policy {
	int      mode;        /* 0: -v <mms> (end-to-end, no qualities) ; 1: -n <mms> (seeded, Maq-like) */
	int      mms;         /* -v / -n value */
	int      seedLen;     /* -l (28) */
	int      qualThresh;  /* -e (70) */
	int      maxBts;      /* --maxbts (125 without --best) */
	uint32_t khits;       /* -k (1) */
	uint32_t mhits;       /* -m (0xffffffff) */
	int      allHits;     /* -a */
	int      nofw, norc;
	int      maqRound;    /* !--nomaqround */
} bto_policy;

/* One reported alignment; fields follow Hit (hit.h:37-128) as filled by
 * EbwtSearchParams::reportHit (ebwt.h:1288-1405). */
typedef struct bto_hit {
	uint32_t read;        /* index of the read in the batch */
	uint32_t tidx, toff;  /* Hit::h */
	uint32_t oms;         /* Hit::oms = bot-top-1 */
	uint16_t cost;
	uint8_t  fw;          /* Hit::fw */
	uint8_t  stratum;
	uint32_t nmm;
	uint32_t mm_off;      /* offset of this hit's mismatches in the mm pool */
} bto_hit;

/* mismatch entry: pos = offset from the 5' end of the original read (Hit::mms bit),
 * refc = reference character code 0..3 (Hit::refcs). */
typedef struct bto_mm { uint16_t pos; uint8_t refc; uint8_t pad; } bto_mm;

/* Operation counters for the roofline accounting of SURVEY.md §8(d). */
typedef struct bto_stats {
	uint64_t lfex;        /* mapLFEx calls (2 loci)               */
	uint64_t lf;          /* single-locus mapLF / mapLF1 calls    */
	uint64_t chase;       /* of which: row-chase steps            */
	uint64_t ftab;        /* ftab jumps (2 reads each)            */
	uint64_t offs;        /* offs[] reads                         */
	uint64_t backtracks;  /* recursive frames entered beyond the first */
} bto_stats;

typedef struct bto_result {
	bto_hit  *hits;  size_t nhits,  cap_hits;
	bto_mm   *mms;   size_t nmms,   cap_mms;
	uint32_t *nhits_per_read;   /* reported hits per read (after -k truncation / -m suppression) */
	uint8_t  *maxed;            /* 1 iff the read exceeded -m */
	uint64_t counters[5];       /* aligned, unaligned, maxed, reported, reportedPaired (hit.h:169-175) */
	bto_stats stats;
} bto_result;

bto_index *bto_index_load(const char *basename, int mirror, char *err, size_t errlen);
void       bto_index_free(bto_index *ix);

/* LF primitives exposed for unit tests (ebwt.h:2334-2560). */
uint32_t bto_map_lf(const bto_index *ix, uint32_t row, int c);           /* mapLF(l, c)          */
uint32_t bto_map_lf1(const bto_index *ix, uint32_t row, int c);          /* mapLF1(row, l, c)    */
void     bto_map_lf_ex(const bto_index *ix, uint32_t row, uint32_t out[4]);
int      bto_row_l(const bto_index *ix, uint32_t row);
uint32_t bto_ftab_hi(const bto_index *ix, uint32_t i);
uint32_t bto_ftab_lo(const bto_index *ix, uint32_t i);
/* Resolve BW row -> joined-text offset (ebwt.h:2693-2746); returns #LF steps via *jumps. */
uint32_t bto_chase(const bto_index *ix, uint32_t row, uint32_t *jumps);
/* joinedToTextOff (ebwt.h:2569-2629); returns 0 and sets *tidx=OFF_MASK when straddling. */
void     bto_joined_to_text_off(const bto_index *ix, uint32_t qlen, uint32_t off,
                                uint32_t *tidx, uint32_t *textoff, uint32_t *tlen);

/* genRandSeed (pat.cpp:21-57). seq = codes 0..4, qual = phred+33 chars. */
uint32_t bto_gen_rand_seed(const uint8_t *seq, const uint8_t *qual, uint32_t len,
                           const char *name, uint32_t namelen, uint32_t global_seed);
/* RandomSource::nextU32 (random_source.h:45-54), state in *last. */
uint32_t bto_rand_next(uint32_t *last);

bto_result *bto_result_new(size_t nreads);
void        bto_result_free(bto_result *r);

/* Align a batch.  seq/qual are concatenated per-read arrays (codes 0..4 / phred+33 chars);
 * offs[i]..offs[i+1] delimits read i (offs has nreads+1 entries); seeds[i] = Read::seed.
 * Returns 0 on success, nonzero if the policy is outside the restated path. */
int bto_align(const bto_index *fw, const bto_index *mirror, const bto_policy *pol,
              size_t nreads, const uint8_t *seq, const uint8_t *qual,
              const uint64_t *offs, const uint32_t *seeds, bto_result *out);

#ifdef __cplusplus
}
#endif
#endif
