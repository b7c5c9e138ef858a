/*
 * bowtie_b200.h — C ABI of the B200-native FM-index backward-search path.
 *
 * This is the drop-in boundary for the hot path of Bowtie 1.3.1.  The reference has no plugin
 * API; its hot path sits between PatternSourcePerThread::nextReadPair() (reads in) and
 * HitSinkPerThread::reportHit()/finishRead() (hits out), inside the per-thread worker functions
 *     exactSearchWorker                    ebwt_search.cpp:1130
 *     mismatchSearchWorkerFull             ebwt_search.cpp:1606
 *     twoOrThreeMismatchSearchWorkerFull   ebwt_search.cpp:2056
 *     seededQualSearchWorkerFull           ebwt_search.cpp:2378
 * operating on two immutable Ebwt objects (ebwt.h:335-1263).  The entry points below replace
 * exactly that: load the same .ebwt files, take a batch of Read records (sequence, qualities,
 * per-read seed), and return for every read what the worker would have handed to its
 * HitSinkPerThread.  Plain pointers and sizes only; no exceptions cross the boundary; every
 * function returns 0 on success and a nonzero code otherwise (bt_last_error() has the message).
 *
 * INTEGRATION.md shows the reference-side binding.
 */
#ifndef BOWTIE_B200_H_
#define BOWTIE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BT_ABI_VERSION 5

typedef struct bt_index bt_index_t;
typedef struct bt_context bt_context_t;

/* Search policy = the option globals of ebwt_search.cpp:153-253 that reach the workers.
 * Layout is shared with the kernels (BtPolicy in bt_core.cuh). */
typedef struct bt_policy {
	int32_t  mode;        /* 0: -v <mms> end-to-end mismatches ; 1: -n <mms> seeded, quality-aware ("maqLike") */
	int32_t  mms;         /* -v: 0..3 ; -n: 0..3   (-v 3 always runs on the best-first path, ebwt_search.cpp:851-854)         */
	int32_t  seed_len;    /* -l, default 28                                  */
	uint32_t qual_thresh; /* -e, default 70                                  */
	uint32_t max_bts;     /* --maxbts, default 125 (maxBtsBetter)            */
	uint32_t khits;       /* -k, default 1                                   */
	uint32_t mhits;       /* -m, default 0xffffffff                          */
	int32_t  all_hits;    /* -a                                              */
	int32_t  nofw, norc;  /* --nofw / --norc                                 */
	int32_t  maq_round;   /* 1 unless --nomaqround                           */
	int32_t  best;        /* --best (also implied by --strata, -M, -v 3): the reference's "stateful" best-first aligners
	                         (UnpairedAlignerV2 over EbwtRangeSource drivers, aligner.h:381-599) instead of the DFS workers */
	int32_t  strata;      /* --strata: NBestFirstStratHitSinkPerThread (hit.h:1070-1129); needs best                       */
	uint32_t max_bts_best;/* --maxbts on the best-first path, default 800 (maxBts, ebwt_search.cpp:186,2644)               */
	int32_t  sample_max;  /* -M: records are kept up to the mhits ceiling (the caller samples one, hit.cpp:16-68); needs best */
	int32_t  paired;      /* paired-end (PairedBWAlignerV1, aligner.h:606-1468): reads 2p and 2p+1 of the batch are mates 1 and 2 of
	                         pair p; found/flags/hits are indexed by p; records alternate upstream/downstream mate (mate number in
	                         bits 25-26 of word 3); needs the index's X.3.ebwt/X.4.ebwt.  With best: PairedBWAlignerV2 (aligner.h:1483-2053). */
	uint32_t min_ins, max_ins;   /* -I (default 0) / -X (default 250), minus the trimmed bases as in aligner.h:975-990                 */
	int32_t  mate1fw, mate2fw;   /* --fr (default): 1, 0 ; --rf: 0, 1 ; --ff: 1, 1                                                    */
	uint32_t pair_tries;  /* --pairtries, default 100 (mixedAttemptLim)                                                              */
} bt_policy_t;

/* Per-read overflow flags (bt_hit_batch_t::flags).  Scratch-related overflows are retried inside the
 * library with a larger workspace; HITS/MM overflows mean the caller's record array was too small
 * for this read (call again for those reads with more slots / mm_cap via bt_read_batch_t::sel). */
#define BT_OVF_STACK 1u
#define BT_OVF_FRAME 2u
#define BT_OVF_PART  4u
#define BT_OVF_HITS  8u
#define BT_OVF_MM   16u

/* A batch of reads = the fields of Read (read.h:42-277) the search consumes.
 * seq: base codes 0=A 1=C 2=G 3=T 4=N (asc2dna), qual: Phred+33 characters, both concatenated;
 * offs[i]..offs[i+1] delimits read i; seeds[i] = Read::seed (genRandSeed, pat.cpp:21-57). */
typedef struct bt_read_batch {
	uint32_t        nreads;
	const uint8_t  *seq;
	const uint8_t  *qual;
	const uint64_t *offs;     /* nreads + 1 */
	const uint32_t *seeds;    /* nreads     */
	const uint32_t *sel;      /* optional: process only these read ids (length nsel), else NULL */
	uint32_t        nsel;
	uint32_t        max_len;  /* longest read in the batch; required by bt_align_batch_device, else 0 = compute */
} bt_read_batch_t;

/* Hit record: BT_HIT_HDR_WORDS header words + mm_cap mismatch words.
 *   w0 tidx (Hit::h.first)     w1 toff (Hit::h.second)     w2 oms (Hit::oms)
 *   w3 cost[15:0] | stratum[23:16] | fw[24]                w4 number of mismatches
 *   w5.. mismatch k: pos[15:0] (offset from the 5' end, Hit::mms) | refc[23:16] (0..3, Hit::refcs)          */
#define BT_HIT_HDR_WORDS 5

/* Per-read results, caller-allocated:
 *   found[i]  = hitsForThisRead_ when the worker called finishRead (hit.h:741-786); the caller applies
 *               "found > mhits => suppressed by -m" and "report min(found, khits)" exactly as finishRead does.
 *   hits      = nreads x slots x (BT_HIT_HDR_WORDS + mm_cap) words; the first min(found, n, slots) records
 *               of read i are valid, in the order the reference's sink would have buffered them. */
typedef struct bt_hit_batch {
	uint32_t *found;      /* nreads */
	uint32_t *flags;      /* nreads */
	uint32_t *hits;
	uint32_t  slots;
	uint32_t  mm_cap;
} bt_hit_batch_t;

typedef struct bt_index_info {
	uint32_t len;         /* joined reference length               */
	uint32_t n_refs;      /* number of reference sequences (nPat)  */
	int32_t  off_rate, ftab_chars;
	int32_t  has_mirror;
	uint64_t device_bytes;
} bt_index_info_t;

/* Operation counters accumulated over all bt_align_* calls on this index since the last reset;
 * units of SURVEY.md §8(d): side fetches = 2*lfex + lf. */
typedef struct bt_stats {
	uint64_t lfex, lf, chase, ftab, offs, backtracks, iters, block_loads;
} bt_stats_t;

int  bt_abi_version(void);
const char *bt_last_error(void);

/* Replaces Ebwt::Ebwt + Ebwt::loadIntoMemory (ebwt.h:402-448, 2835-3445) for X.1.ebwt/X.2.ebwt and,
 * if need_mirror, X.rev.1.ebwt/X.rev.2.ebwt; uploads to `device` and re-lays-out the BWT for the kernels. */
int  bt_index_load(const char *basename, int need_mirror, int device, bt_index_t **out);
void bt_index_free(bt_index_t *ix);
int  bt_index_info(const bt_index_t *ix, bt_index_info_t *info);
const char *bt_index_refname(const bt_index_t *ix, uint32_t i);   /* Ebwt::_refnames */
uint32_t    bt_index_reflen(const bt_index_t *ix, uint32_t i);    /* Ebwt::_plen     */

void bt_policy_init(bt_policy_t *p);                              /* resetOptions defaults (ebwt_search.cpp:153-253) */

/* Replaces one pass of the worker loop over a batch (GET_READ ... search_*.c ... FINISH_READ).
 * Host buffers in, host buffers out; the call returns when the results are in `out`.
 * `stream` is a cudaStream_t (NULL = default stream). */
int  bt_align_batch(bt_index_t *ix, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out, void *stream);

/* Same, with every pointer of `in` and `out` already resident on the index's device; enqueues on
 * `stream` and returns without synchronising (scratch overflows are retried by a second enqueued pass). */
int  bt_align_batch_device(bt_index_t *ix, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out, void *stream);

/* Contexts: everything one in-flight batch needs besides the shared immutable index (scratch, work queue,
 * device staging).  The reference runs N worker threads over one Ebwt (ebwt_search.cpp:1385-1405); the
 * equivalent here is N contexts on N CUDA streams over one bt_index_t.  One call at a time per context.
 * bt_align_batch / bt_align_batch_device use an internal default context. */
int  bt_context_create(bt_index_t *ix, bt_context_t **out);
void bt_context_free(bt_context_t *cx);
int  bt_context_align(bt_context_t *cx, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out, void *stream);
/* enqueue only (pinned host buffers, no `sel`); results are valid after bt_context_sync on the same stream */
int  bt_context_align_async(bt_context_t *cx, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out, void *stream);
int  bt_context_align_device(bt_context_t *cx, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out, void *stream);
int  bt_context_sync(bt_context_t *cx, void *stream);
/* The library runs the few very long searches of a batch ("heavy" reads) and scratch-overflow retries on an internal
 * side stream so that they overlap the next batch; bt_context_join makes `stream` wait for them (results of
 * bt_context_align_device are complete only after join or sync; bt_align_batch_device joins by itself). */
int  bt_context_join(bt_context_t *cx, void *stream);

int  bt_stats_get(bt_index_t *ix, bt_stats_t *out, int reset);    /* synchronises the device */

/* Replaces bowtie-build (ebwt_build.cpp:303-480 driver(); Ebwt::initFromVector / joinToDisk / buildToDisk, ebwt.h:3825-4388;
 * fastaRefReadSizes, ref_read.cpp:202-273): writes out_base.{1,2,3,4}.ebwt and out_base.rev.{1,2}.ebwt, byte-identical to the
 * reference's files for the same FASTA input, -o (off_rate, default 5) and -t (ftab_chars, default 10).  Everything that is
 * O(genome) runs on `device` (suffix sort, BWT, side packing with occ words, SA sample, ftab histogram, the 2-bit reference);
 * FASTA parsing and writing the files are host work. */
int  bt_index_build(const char *const *fasta_paths, uint32_t n_paths, const char *out_base, int off_rate, int ftab_chars, int device);

/* The same from an already parsed reference — what fastaRefReadSizes / fastaRefReadAppend (ref_read.cpp:10-141,202-273) hand to
 * Ebwt::initFromVector: `text` = the joined unambiguous characters (codes 0..3, one per byte, host memory, text_len < 2^32 - 1),
 * `recs` = the RefRecords (ref_read.h:57-88: `off` gap characters, then `len` unambiguous ones; `first` = 1 on the first record of
 * a sequence), `names` = one per sequence.  Used where the reference exists in memory (synthetic genomes of bench.py). */
typedef struct bt_ref_record { uint32_t off, len, first; } bt_ref_record_t;
int  bt_index_build_text(const uint8_t *text, uint64_t text_len, const bt_ref_record_t *recs, uint32_t n_recs, const char *const *names, uint32_t n_names,
                         const char *out_base, int off_rate, int ftab_chars, int device);

/* Device I/O path: read ingest and hit formatting on the device, around the same search.
 *   bt_io_parse_fastq    replaces FastqPatternSource::parse + genRandSeed (pat.cpp:858-975, 21-57; read.h:118-132) for well-formed
 *                        4-line FASTQ (Phred+33): `text` (host memory) must start at a record; the complete, regular records at its
 *                        start — at most max_reads, never the chunk's last complete one — become a device-resident read batch.
 *                        *consumed = bytes of text they cover; *irregular != 0 if a record was met that the host parser must handle
 *                        (the caller hands everything from text + *consumed on to it).
 *   bt_io_align_format   runs the policy over that batch (bt_context_align_device on the io's context) and formats the hits on the
 *                        device: the default format (VerboseHitSink::append, hit.cpp:176-240) or SAM records (SAMHitSink::append /
 *                        reportUnOrMax, sam.cpp:57-257; no header), unpaired, with finishRead's -k / -m arithmetic (hit.h:741-786), in
 *                        read order.  *out_text (host memory owned by the io, valid until its next call) holds *out_bytes bytes;
 *                        counters = { aligned, unaligned, maxed, reported } of hit.h:169-175.  Not provided here (callers format those
 *                        from hit records): paired-end, -a, -M, --suppress / --refidx / cost columns.  Returns 2 (and produces nothing)
 *                        when a read of the batch has more than 32 mismatches or more hits than -k records: the caller formats that
 *                        batch from hit records (the host parser re-reads the same text).
 * One io per in-flight chunk; calls on one io are synchronous. */
typedef struct bt_io bt_io_t;
typedef struct bt_io_format {
	int32_t  sam;             /* 0: default format, 1: SAM records               */
	int32_t  no_unal;         /* --no-unal                                        */
	int32_t  no_qname_trunc;  /* --sam-no-qname-trunc                             */
	int32_t  full_ref;        /* --fullref                                        */
	int32_t  off_base;        /* -B                                               */
	uint32_t mapq;            /* --mapq (255)                                     */
} bt_io_format_t;
int  bt_io_create(bt_context_t *cx, bt_io_t **out);
void bt_io_free(bt_io_t *io);
int  bt_io_parse_fastq(bt_io_t *io, const char *text, uint64_t nbytes, uint32_t global_seed, uint32_t max_reads, uint32_t *nreads, uint64_t *consumed, int *irregular);
int  bt_io_align_format(bt_io_t *io, const bt_policy_t *pol, const bt_io_format_t *fmt, const char **out_text, uint64_t *out_bytes, uint64_t counters[4]);

/* The path's only collective (reads shard across GPUs with no data-path exchange): sums the five summary counters of hit.h:169-175
 * { aligned, unaligned, maxed, reported, reportedPaired } over the ranks of `nccl_comm` (an ncclComm_t; NCCL is bound at call
 * time) on the calling thread's current device.  In place; returns when `counters` holds the sums. */
int  bt_counters_allreduce(void *nccl_comm, uint64_t counters[5], void *stream);

/* LF primitives on the device layout, for parity tests: computes, for each row, mapLFEx-style
 * (fchr[c] + occ(c,row)) for c = 0..3 and rowL.  rows/out are host arrays; out has 5 words per row. */
int  bt_debug_lf(bt_index_t *ix, int mirror, const uint32_t *rows, uint32_t n, uint32_t *out);

/* Page-locked host memory for the buffers of bt_context_align_async (whose copies are asynchronous only from / to pinned memory;
 * with pageable buffers the call is correct but returns only when its last copy has completed).  NULL when the allocation fails. */
void *bt_host_alloc(size_t bytes);
void  bt_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
