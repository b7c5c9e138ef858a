/*
 * bowtie_main.cpp — `bowtie`-compatible host driver over libbowtie_b200.so.
 *
 * Keeps the reference's command line (ebwt_search.cpp:443-545, 614-919), read-file formats
 * (pat.cpp: FASTQ 797-975, FASTA 531-640, -F 651-790, raw 1129-1213, -c 357-523, --12 980-1124; -1/-2, --interleaved; light parse
 * and parse() followed character for character, so malformed input behaves as it does there), the default
 * hit format (hit.cpp:73-301) and SAM (sam.cpp:20-257) for single reads and pairs, the read dumps --al/--un/--max
 * (hit.h:385-492) and the stderr summary (hit.h:270-346); the search itself (everything the reference's
 * *SearchWorker* functions do) is one bt_context_align_async() call per batch of reads or pairs.  --best, --strata,
 * -M and -v 3 select the library's best-first path, paired input its paired-end path; what is not provided is
 * rejected with a message — nothing falls back to a CPU search.
 * Threads: one parser thread fills a ring of batches (well-formed FASTQ: several conversion threads per buffer), the main
 * thread launches batch k+1 and formats batch k with -p threads; output order is input order.
 */
#include <algorithm>
#include <array>
#include <deque>
#include <map>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <getopt.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <unistd.h>
#include <sys/stat.h>
#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "../../../include/bowtie_b200.h"

#define BOWTIE_VERSION "1.3.1"

/* ---------------------------------------------------------------------------------------------- */
/* options (names follow the reference's globals)                                                  */
/* ---------------------------------------------------------------------------------------------- */
enum { FASTQ = 1, FASTA, RAW, CMDLINE, FASTA_CONT };
struct Opts {
	int format = FASTQ;
	std::string ebwtFile, outfile;
	std::vector<std::string> queries;
	int mismatches = 0, seedMms = 2, maqLike = 1, seedLen = 28, qualThresh = 70, maxBts = 125, maxBtsBest = 800;
	bool best = false, strata = false, sampleMax = false, bestFlag = false;   /* bestFlag: --best itself (it alone selects the V2 paired aligner) */
	std::vector<std::string> mates1, mates2, interleaved, tabbed;
	std::string dumpAl, dumpUn, dumpMax;                   /* --al / --un / --max: dump reads by outcome (hit.h:385-492) */
	bool bestPaired = false;   /* the pairs' --best: PairedBWAlignerV2 */
	uint32_t minInsert = 0, maxInsert = 250, pairTries = 100; bool mate1fw = true, mate2fw = false;
	bool noMaqRound = false, nofw = false, norc = false, allHits = false;
	uint32_t khits = 1, mhits = 0xffffffffu;
	uint32_t skipReads = 0, qUpto = 0xffffffffu;
	int trim3 = 0, trim5 = 0;
	bool solexaQuals = false, phred64Quals = false, integerQuals = false;
	std::vector<std::string> qualities, qualities1, qualities2;   /* -Q / --Q1 / --Q2: opened and counted, never read (pat.cpp:333-348) */
	size_t fastaContLen = 0, fastaContFreq = 0;            /* -F len,freq */
	bool pev2 = false, reorder = false;
	bool sam = false, samNoHead = false, samNoSQ = false, noUnal = false, fullRef = false, refIdx = false, printCost = false;
	bool noQnameTrunc = false, quiet = false, timing = false;
	std::string rgs;
	int defaultMapq = 255, offBase = 0;
	uint32_t seed = 0;
	std::vector<bool> suppress = std::vector<bool>(64, false);
	int nthreads = 1, device = 0;
	uint32_t batch = 1u << 20;
	std::string argstr;
};

/* Errors can surface on the parser thread while the main thread has CUDA work in flight: leaving through exit() there would run
 * the runtime's teardown under the main thread's feet, so threads other than the main one leave with _exit(). */
static const std::thread::id g_main_thread = std::this_thread::get_id();
[[noreturn]] static void leave(int rc) {
	fflush(stderr);
	if (std::this_thread::get_id() == g_main_thread) exit(rc);
	fflush(stdout);
	_exit(rc);
}
static void die(const std::string &m) { fprintf(stderr, "%s\n", m.c_str()); leave(1); }

enum {
	ARG_PHRED33 = 256, ARG_PHRED64, ARG_SOLEXA, ARG_SOLEXA13, ARG_NOMAQROUND, ARG_NOFW, ARG_NORC, ARG_MAXBTS, ARG_BEST, ARG_STRATA,
	ARG_QUIET, ARG_REFIDX, ARG_SUPPRESS, ARG_FULLREF, ARG_MAPQ, ARG_SAM_NOHEAD, ARG_SAM_NOSQ, ARG_SAM_RG, ARG_NO_UNAL, ARG_SEED,
	ARG_COST, ARG_REORDER, ARG_WRAPPER, ARG_VERSION, ARG_IGNORED0, ARG_IGNORED1, ARG_DEVICE, ARG_BATCH, ARG_SAM_NO_QNAME_TRUNC,
	ARG_LARGE_INDEX, ARG_PAIRED, ARG_INTERLEAVED, ARG_AL, ARG_UN, ARG_MAXDUMP, ARG_FF, ARG_FR, ARG_RF, ARG_PAIRTRIES,
	ARG_INTEGER_QUALS, ARG_QUALS1, ARG_QUALS2, ARG_STATEFUL, ARG_PEV2, ARG_USAGE, ARG_IGNORED_ARG
};

static const char *short_options = "fF:qrchu:v:s:at3:5:o:e:n:l:p:k:m:M:1:2:I:X:x:B:ySQ:";
static struct option long_options[] = {
	{"skip", required_argument, 0, 's'}, {"qupto", required_argument, 0, 'u'}, {"trim5", required_argument, 0, '5'},
	{"trim3", required_argument, 0, '3'}, {"phred33-quals", no_argument, 0, ARG_PHRED33}, {"phred64-quals", no_argument, 0, ARG_PHRED64},
	{"solexa-quals", no_argument, 0, ARG_SOLEXA}, {"solexa1.3-quals", no_argument, 0, ARG_SOLEXA13}, {"seedmms", required_argument, 0, 'n'},
	{"maqerr", required_argument, 0, 'e'}, {"seedlen", required_argument, 0, 'l'}, {"nomaqround", no_argument, 0, ARG_NOMAQROUND},
	{"nofw", no_argument, 0, ARG_NOFW}, {"norc", no_argument, 0, ARG_NORC}, {"maxbts", required_argument, 0, ARG_MAXBTS},
	{"tryhard", no_argument, 0, 'y'}, {"all", no_argument, 0, 'a'}, {"best", no_argument, 0, ARG_BEST}, {"strata", no_argument, 0, ARG_STRATA},
	{"time", no_argument, 0, 't'}, {"offbase", required_argument, 0, 'B'}, {"quiet", no_argument, 0, ARG_QUIET},
	{"refidx", no_argument, 0, ARG_REFIDX}, {"suppress", required_argument, 0, ARG_SUPPRESS}, {"fullref", no_argument, 0, ARG_FULLREF},
	{"sam", no_argument, 0, 'S'}, {"mapq", required_argument, 0, ARG_MAPQ}, {"sam-nohead", no_argument, 0, ARG_SAM_NOHEAD},
	{"sam-nosq", no_argument, 0, ARG_SAM_NOSQ}, {"sam-noSQ", no_argument, 0, ARG_SAM_NOSQ}, {"sam-RG", required_argument, 0, ARG_SAM_RG},
	{"no-unal", no_argument, 0, ARG_NO_UNAL}, {"sam-no-qname-trunc", no_argument, 0, ARG_SAM_NO_QNAME_TRUNC},
	{"threads", required_argument, 0, 'p'}, {"seed", required_argument, 0, ARG_SEED}, {"cost", no_argument, 0, ARG_COST},
	{"reorder", no_argument, 0, ARG_REORDER}, {"wrapper", required_argument, 0, ARG_WRAPPER}, {"version", no_argument, 0, ARG_VERSION},
	{"chunkmbs", required_argument, 0, ARG_IGNORED1}, {"mm", no_argument, 0, ARG_IGNORED0}, {"shmem", no_argument, 0, ARG_IGNORED1}, {"help", no_argument, 0, 'h'},
	{"device", required_argument, 0, ARG_DEVICE}, {"reads-per-batch", required_argument, 0, ARG_BATCH},
	{"large-index", no_argument, 0, ARG_LARGE_INDEX}, {"12", required_argument, 0, ARG_PAIRED}, {"interleaved", required_argument, 0, ARG_INTERLEAVED},
	{"ff", no_argument, 0, ARG_FF}, {"fr", no_argument, 0, ARG_FR}, {"rf", no_argument, 0, ARG_RF}, {"pairtries", required_argument, 0, ARG_PAIRTRIES},
	{"al", required_argument, 0, ARG_AL}, {"un", required_argument, 0, ARG_UN}, {"max", required_argument, 0, ARG_MAXDUMP},
	{"minins", required_argument, 0, 'I'}, {"maxins", required_argument, 0, 'X'},
	{"khits", required_argument, 0, 'k'}, {"mhits", required_argument, 0, 'm'}, {"offrate", required_argument, 0, 'o'},
	{"integer-quals", no_argument, 0, ARG_INTEGER_QUALS}, {"quals", required_argument, 0, 'Q'}, {"Q1", required_argument, 0, ARG_QUALS1},
	{"Q2", required_argument, 0, ARG_QUALS2}, {"stateful", no_argument, 0, ARG_STATEFUL}, {"pev2", no_argument, 0, ARG_PEV2},
	{"usage", no_argument, 0, ARG_USAGE},
	/* accepted for command-line compatibility, no effect here: diagnostics, memory tuning of the CPU implementation, defaults */
	{"verbose", no_argument, 0, ARG_IGNORED0}, {"startverbose", no_argument, 0, ARG_IGNORED0}, {"strandfix", no_argument, 0, ARG_IGNORED0},
	{"noreconcile", no_argument, 0, ARG_IGNORED0}, {"chunkverbose", no_argument, 0, ARG_IGNORED0}, {"filepar", no_argument, 0, ARG_IGNORED0},
	{"chunksz", required_argument, 0, ARG_IGNORED_ARG}, {"prewidth", required_argument, 0, ARG_IGNORED_ARG},
	{"thread-ceiling", required_argument, 0, ARG_IGNORED_ARG}, {"thread-piddir", required_argument, 0, ARG_IGNORED_ARG},
	{0, 0, 0, 0}
};

static long parse_int(long lo, const char *msg) {
	char *end = NULL;
	long v = strtol(optarg, &end, 10);
	if (end == optarg || *end != 0 || v < lo) die(msg);
	return v;
}
static void split(const std::string &s, char sep, std::vector<std::string> &out) {
	size_t a = 0;
	while (a <= s.size()) { size_t b = s.find(sep, a); if (b == std::string::npos) b = s.size(); if (b > a) out.push_back(s.substr(a, b - a)); a = b + 1; }
}

static void parse_options(int argc, char **argv, Opts &o) {
	for (int i = 0; i < argc; i++) { o.argstr += argv[i]; if (i < argc - 1) o.argstr += " "; }
	int c, idx = 0;
	bool vset = false;
	while ((c = getopt_long(argc, argv, short_options, long_options, &idx)) != -1) {
		switch (c) {
		case 'f': o.format = FASTA; break;
		case 'q': o.format = FASTQ; break;
		case 'r': o.format = RAW; break;
		case 'c': o.format = CMDLINE; break;
		case 'x': o.ebwtFile = optarg; break;
		case 's': o.skipReads = (uint32_t)parse_int(0, "-s arg must be positive"); break;
		case 'u': o.qUpto = (uint32_t)parse_int(1, "-u/--qupto arg must be at least 1"); break;
		case '3': o.trim3 = (int)parse_int(0, "-3/--trim3 arg must be at least 0"); break;
		case '5': o.trim5 = (int)parse_int(0, "-5/--trim5 arg must be at least 0"); break;
		case 'v': o.maqLike = 0; o.mismatches = (int)parse_int(0, "-v arg must be at least 0"); vset = true;
			if (o.mismatches > 3) die("-v arg must be at most 3");
			break;
		case 'n': o.seedMms = (int)parse_int(0, "-n/--seedmms arg must be at least 0 and at most 3"); o.maqLike = 1;
			if (o.seedMms > 3) die("-n/--seedmms arg must be at least 0 and at most 3");
			break;
		case 'e': o.qualThresh = (int)parse_int(1, "-e/--err arg must be at least 1"); break;
		case 'l': o.seedLen = (int)parse_int(5, "-l/--seedlen arg must be at least 5"); break;
		case 'k': o.khits = (uint32_t)parse_int(1, "-k arg must be at least 1"); break;
		case 'm': o.mhits = (uint32_t)parse_int(1, "-m arg must be at least 1"); break;
		case 'a': o.allHits = true; break;
		case 'p': o.nthreads = (int)parse_int(1, "-p/--threads arg must be at least 1"); break;
		case 't': o.timing = true; break;
		case 'B': o.offBase = (int)parse_int(-999999, "-B/--offbase cannot be a large negative number"); break;
		case 'S': o.sam = true; break;
		case 'y': o.maxBts = o.maxBtsBest = 0x7fffffff; break;
		case 'h': printf("Usage: bowtie-b200-align [options]* -x <ebwt> {<s> | -c <seqs>} [<hits>]\n  (option names follow bowtie 1.3.1; see DESIGN.md for the supported subset)\n"); exit(0);
		case 'M': o.sampleMax = true; o.mhits = (uint32_t)parse_int(1, "-m arg must be at least 1"); break;
		case '1': split(optarg, ',', o.mates1); break;
		case '2': split(optarg, ',', o.mates2); break;
		case 'I': o.minInsert = (uint32_t)parse_int(0, "-I arg must be positive"); break;
		case 'X': o.maxInsert = (uint32_t)parse_int(1, "-X arg must be at least 1"); break;
		case ARG_FF: o.mate1fw = true; o.mate2fw = true; break;
		case ARG_FR: o.mate1fw = true; o.mate2fw = false; break;
		case ARG_RF: o.mate1fw = false; o.mate2fw = true; break;
		case ARG_PAIRTRIES: o.pairTries = (uint32_t)parse_int(1, "--pairtries arg must be at least 1"); break;
		case ARG_INTERLEAVED: split(optarg, ',', o.interleaved); break;
		case ARG_AL: o.dumpAl = optarg; break;
		case ARG_UN: o.dumpUn = optarg; break;
		case ARG_MAXDUMP: o.dumpMax = optarg; break;
		case ARG_PAIRED: split(optarg, ',', o.tabbed); break;                       /* --12: one record per line, unpaired (3 fields) or paired (5) */
		case ARG_BEST: o.best = true; o.bestFlag = true; break;
		case ARG_STRATA: o.strata = true; break;
		case ARG_LARGE_INDEX: die("Error: large (64-bit) indexes are not supported"); break;
		case ARG_PHRED33: o.solexaQuals = false; o.phred64Quals = false; break;
		case ARG_PHRED64: case ARG_SOLEXA13: o.solexaQuals = false; o.phred64Quals = true; break;
		case ARG_SOLEXA: o.solexaQuals = true; o.phred64Quals = false; break;
		case ARG_NOMAQROUND: o.noMaqRound = true; break;
		case ARG_NOFW: o.nofw = true; break;
		case ARG_NORC: o.norc = true; break;
		case ARG_MAXBTS: o.maxBts = o.maxBtsBest = (int)parse_int(0, "--maxbts must be positive"); break;
		case ARG_QUIET: o.quiet = true; break;
		case ARG_REFIDX: o.refIdx = true; break;
		case ARG_FULLREF: o.fullRef = true; break;
		case ARG_SUPPRESS: { std::vector<std::string> f; split(optarg, ',', f); for (auto &x : f) { int ii = atoi(x.c_str()); if (ii < 1) die("--suppress arg must be at least 1"); if (ii <= 64) o.suppress[ii - 1] = true; } break; }
		case ARG_MAPQ: o.defaultMapq = (int)parse_int(0, "--mapq must be positive"); break;
		case ARG_SAM_NOHEAD: o.samNoHead = true; break;
		case ARG_SAM_NOSQ: o.samNoSQ = true; break;
		case ARG_SAM_RG: if (!o.rgs.empty()) o.rgs += '\t'; o.rgs += optarg; break;
		case ARG_NO_UNAL: o.noUnal = true; break;
		case ARG_SAM_NO_QNAME_TRUNC: o.noQnameTrunc = true; break;
		case ARG_SEED: o.seed = (uint32_t)parse_int(0, "--seed arg must be at least 0"); break;
		case ARG_COST: o.printCost = true; break;
		case ARG_VERSION: printf("%s version %s (B200 search path)\n", argv[0], BOWTIE_VERSION); exit(0);
		case ARG_DEVICE: o.device = (int)parse_int(0, "--device must be >= 0"); break;
		case ARG_BATCH: o.batch = (uint32_t)parse_int(1, "--reads-per-batch arg must be at least 1"); break;
		case ARG_REORDER: o.reorder = true; break;                                   /* output is always in read order */
		case ARG_WRAPPER: case ARG_IGNORED0: case ARG_IGNORED1: case ARG_IGNORED_ARG: break;
		case 'o': parse_int(1, "-o/--offrate arg must be at least 1"); break;          /* a sparser SA sample only slows the reference's locate down; results do not depend on it */
		case ARG_INTEGER_QUALS: o.integerQuals = true; break;
		case 'Q': split(optarg, ',', o.qualities); o.integerQuals = true; break;       /* ebwt_search.cpp:709-720 */
		case ARG_QUALS1: split(optarg, ',', o.qualities1); o.integerQuals = true; break;
		case ARG_QUALS2: split(optarg, ',', o.qualities2); o.integerQuals = true; break;
		case ARG_STATEFUL: o.best = true; break;                                      /* the best-first aligners without --best's switch to PairedBWAlignerV2 */
		case ARG_PEV2: o.pev2 = true; break;
		case ARG_USAGE: printf("Usage: bowtie-b200-align [options]* -x <ebwt> {-1 <m1> -2 <m2> | --12 <r> | --interleaved <i> | <s>} [<hit>]\n  (option names follow bowtie 1.3.1)\n"); exit(0);
		case 'F': {                                                                   /* -F <len>,<freq>: reads are substrings of the FASTA input */
			std::vector<std::string> f; split(optarg, ',', f);
			if (f.size() < 2) die("Error: -F takes <length>,<frequency>");
			o.format = FASTA_CONT; o.fastaContLen = (size_t)strtoull(f[0].c_str(), NULL, 10); o.fastaContFreq = (size_t)strtoull(f[1].c_str(), NULL, 10);
			if (o.fastaContLen == 0 || o.fastaContLen >= 1024 || o.fastaContFreq == 0) die("Error: -F needs 0 < length < 1024 and frequency > 0");
			break;
		}
		default: die("Error: unknown or unsupported option (see --help)");
		}
	}
	(void)vset;
	/* ebwt_search.cpp:851-854, 877-891 */
	if (!o.maqLike && o.mismatches == 3) o.best = true;
	if (!o.best && o.sampleMax) {
		if (!o.quiet) fprintf(stderr, "Warning: -M was specified w/o --best; automatically enabling --best\n");
		o.best = true;
	}
	if (o.strata && !o.best) die("--strata must be combined with --best");
	if (o.strata && !o.allHits && o.khits == 1 && o.mhits == 0xffffffffu) die("--strata has no effect unless combined with -m, -a, or -k N where N > 1");
	if (o.qUpto + o.skipReads > o.qUpto) o.qUpto += o.skipReads;                 /* ebwt_search.cpp:893-895 */
	if (o.ebwtFile.empty()) {
		if (optind >= argc) die("No index, query, or output file specified!");
		fprintf(stderr, "Setting the index via positional argument will be deprecated in a future release. Please use -x option instead.\n");
		o.ebwtFile = argv[optind++];
	}
	if (o.mates1.size() != o.mates2.size()) {
		fprintf(stderr, "Error: %zu mate files/sequences were specified with -1, but %zu\nmate files/sequences were specified with -2.  The same number of mate files/\nsequences must be specified with -1 and -2.\n", o.mates1.size(), o.mates2.size());
		exit(1);
	}
	if (o.nthreads == 1) o.reorder = false;                                       /* ebwt_search.cpp:835-842 */
	if (o.reorder && !o.sam) die("Bowtie will reorder its output only when outputting SAM.\nPlease specify the `-S` parameter if you intend on using this option.");
	o.bestPaired = o.bestFlag || o.pev2;                                                    /* pairs: only --best switches to PairedBWAlignerV2 (useV1 = false, ebwt_search.cpp:776); -M / -v 3 alone keep V1 */
	if (o.mates1.empty() && o.interleaved.empty() && o.tabbed.empty()) {
		if (optind >= argc) die("No query or output file specified!");
		split(argv[optind++], ',', o.queries);
	}
	if (optind < argc) o.outfile = argv[optind++];
	if (optind < argc) die(std::string("Extra parameter(s) specified: ") + argv[optind]);
	if (o.sam) std::fill(o.suppress.begin(), o.suppress.end(), false);
}

/* ---------------------------------------------------------------------------------------------- */
/* read input                                                                                      */
/* ---------------------------------------------------------------------------------------------- */
static uint8_t asc2dna[256];
static uint8_t alpha_code[256];       /* asc2dna for letters, 4 for '.', 255 for characters a read line skips */
static const unsigned char solToPhred[] = {   /* qual.cpp: Solexa (log-odds) -> Phred, index = sol + 10 */
	0, 1, 1, 1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29,
	30, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63, 64, 65,
	66, 67, 68, 69, 70, 71, 72, 73, 74, 75, 76, 77, 78, 79, 80, 81, 82, 83, 84, 85, 86, 87, 88, 89, 90, 91, 92, 93, 94, 95, 96, 97, 98, 99, 100 };

struct ReadRec { std::string name, seq /* codes 0..4 */, qual /* phred+33 */, orig /* Read::readOrigBuf: the record as it stood in the input (for --al/--un/--max) */; };

static uint32_t gen_rand_seed(const ReadRec &r, uint32_t seed);

struct Reader {
	const Opts &o;
	size_t fileIdx = 0;
	gzFile f = NULL;
	std::vector<char> buf; size_t len = 0, pos = 0; bool eof = true;   /* buf[pos, len) is unread input */
	uint64_t rdid = 0;
	bool first = true;
	size_t cmdIdx = 0;
	const std::vector<std::string> &files;
	const bool keepOrig;                /* --al/--un/--max need the records as they stood in the input */
	std::string l1, l2, l3, l4;         /* line buffers, reused from record to record */
	bool recEof = false;                /* FASTA: the last record ran to the end of the file */
	std::vector<std::string> rawPieces; size_t rawQ = 0;   /* raw: the CR-separated pieces of the current line */
	const std::vector<std::string> *qfiles;   /* -Q/--Q1/--Q2 with -f: one quality file per read file; the reference opens them and never reads them */
	/* FastaContinuousPatternSource state (pat.h:596-665): the last 1024 bases, how many to skip before the next read, where we are */
	char fcBuf[1024]; size_t fcCur = 0, fcEat = 0; bool fcBeginning = true; uint64_t fcPos = 0, fcLast = 0; std::string fcPrefix;
	void fc_reset() { fcEat = o.fastaContLen - 1; fcPrefix.clear(); fcBeginning = true; fcCur = 0; fcLast = fcPos; }
	Reader(const Opts &oo, const std::vector<std::string> &ff, const std::vector<std::string> *qf = NULL) : o(oo), files(ff),
		keepOrig(!oo.dumpAl.empty() || !oo.dumpUn.empty() || !oo.dumpMax.empty()), qfiles(qf && !qf->empty() && oo.format == FASTA ? qf : NULL) {
		if (qfiles && !files.empty() && qfiles->size() != files.size())                 /* pat.h:334-340 */
			die("Error: Different numbers of input FASTA/quality files (" + std::to_string(files.size()) + "/" + std::to_string(qfiles->size()) + ")");
		fc_reset();
	}
	/* CFilePatternSource::open (pat.cpp:283-352): files that cannot be opened are skipped with a warning; running out of files
	 * that way — rather than by reading the last one to its end — ends the run with status 1 */
	bool open_next() {
		if (f) { gzclose(f); f = NULL; }
		if (fileIdx >= files.size()) return false;
		while (fileIdx < files.size()) {
			const std::string &fn = files[fileIdx++];
			if (fn != "-") { struct stat st; if (stat(fn.c_str(), &st) != 0) perror("stat"); }   /* is_gzipped_file, pat.h:418-422 */
			f = (fn == "-") ? gzdopen(0, "rb") : gzopen(fn.c_str(), "rb");
			if (!f) { fprintf(stderr, "Warning: Could not open read file \"%s\" for reading; skipping...\n", fn.c_str()); continue; }
			if (qfiles) {
				const std::string &qn = (*qfiles)[fileIdx - 1];
				FILE *q = qn == "-" ? NULL : fopen(qn.c_str(), "rb");
				if (qn != "-" && !q) {
					fprintf(stderr, "Warning: Could not open quality file \"%s\" for reading; skipping...\n", qn.c_str());
					gzclose(f); f = NULL;
					continue;
				}
				if (q) fclose(q);
			}
			fc_reset();
			gzbuffer(f, 1 << 20);
			if (buf.empty()) buf.resize(1 << 22);
			len = pos = 0; eof = false; first = true; recEof = false;
			return true;
		}
		leave(1);
	}
	bool refill() {                     /* false at end of input */
		if (eof) return false;
		const int n = gzread(f, buf.data(), (unsigned)buf.size());
		if (n <= 0) { eof = true; len = pos = 0; return false; }
		len = (size_t)n; pos = 0;
		return true;
	}
	/* true iff what is left of the file holds exactly one or two newlines; the unread input stays in the buffer */
	bool tail_aborts() {
		size_t scanned = pos; int nl = 0;
		for (;;) {
			while (nl < 3) {
				const char *q = (const char *)memchr(buf.data() + scanned, '\n', len - scanned);
				if (!q) { scanned = len; break; }
				nl++; scanned = (size_t)(q - buf.data()) + 1;
			}
			if (nl >= 3) return false;
			if (eof) break;
			if (pos > 0) { memmove(buf.data(), buf.data() + pos, len - pos); scanned -= pos; len -= pos; pos = 0; }
			if (len == buf.size()) buf.resize(buf.size() * 2);
			const int n = gzread(f, buf.data() + len, (unsigned)(buf.size() - len));
			if (n <= 0) { eof = true; break; }
			len += (size_t)n;
		}
		return nl == 1 || nl == 2;
	}
	int getc_() {
		if (pos >= len && !refill()) return -1;
		return (unsigned char)buf[pos++];
	}
	int peek_() { int c = getc_(); if (c >= 0) pos--; return c; }
	bool line_hit_eof = false;        /* the last getline_ ended at EOF instead of a newline */
	bool getline_(std::string &s) {   /* returns false at EOF with nothing read; strips \n, keeps \r handling to callers */
		s.clear();
		line_hit_eof = false;
		if (pos >= len && !refill()) { line_hit_eof = true; return false; }
		for (;;) {                      /* whole spans at a time: memchr for the newline, append, refill when the buffer runs out */
			const char *b = buf.data() + pos;
			const char *nl = (const char *)memchr(b, '\n', len - pos);
			if (nl) { s.append(b, (size_t)(nl - b)); pos += (size_t)(nl - b) + 1; return true; }
			s.append(b, len - pos); pos = len;
			if (!refill()) { line_hit_eof = true; return true; }
		}
	}
	char int_to_phred33(int iq) const {                                   /* intToPhred33, qual.h:135-153 */
		int pq;
		if (o.solexaQuals) pq = (iq < -10 ? 0 : solToPhred[std::min(iq, 90) + 10]) + 33;
		else pq = (iq <= 93 ? iq : 93) + 33;
		if (pq < 33) die("Saw negative Phred quality " + std::to_string(pq - 33) + ".");
		return (char)pq;
	}
	static void wrong_quality_format(const std::string &name) {        /* wrongQualityFormat, pat.cpp:1215-1220 */
		die("Encountered a space parsing the quality string for read " + name + "\nIf this is a FASTQ file with integer (non-ASCII-encoded) qualities, please\nre-run Bowtie with the --integer-quals option.");
	}
	char to_phred33(int c, const std::string &name) const {
		if (c == ' ') die("Saw a space but expected an ASCII-encoded quality value.\nAre quality values formatted as integers?  If so, try --integer-quals.");
		if (o.solexaQuals) {
			int sol = c - 64; int p = sol < -10 ? 0 : solToPhred[sol + 10];
			return (char)(p + 33);
		} else if (o.phred64Quals) {
			if (c < 64) die("Saw ASCII character " + std::to_string(c) + " but expected 64-based Phred qual.\nTry not specifying --solexa1.3-quals/--phred64-quals.");
			return (char)(c - 31);
		}
		if (c < 33) die("Saw ASCII character " + std::to_string(c) + " but expected 33-based Phred qual.");
		(void)name;
		return (char)c;
	}
	void finish_seq(ReadRec &r, const std::string &raw, int &trimmed5, int &trimmed3) const {
		int nchar = 0;
		r.seq.resize(raw.size());
		size_t k = 0;
		for (char ch : raw) {                                                /* alpha_code: letters and '.' -> base code, everything else skipped */
			const uint8_t code = alpha_code[(unsigned char)ch];
			if (code != 255) { if (nchar++ >= o.trim5) r.seq[k++] = (char)code; }
		}
		r.seq.resize(k);
		trimmed5 = nchar - (int)r.seq.size();
		trimmed3 = std::min<int>(o.trim3, (int)r.seq.size());
		r.seq.resize(r.seq.size() - (size_t)trimmed3);
	}
	/* Fast path for the common case — well-formed 4-line FASTQ records, Phred+33, no trimming: the records that are complete
	 * in the current buffer are located with memchr and parsed by `nth` threads; anything unusual (blank lines, CR, a quality
	 * character below '!', lengths that disagree, a record that straddles the buffer) ends the run before that record and is
	 * left to next(), whose behaviour — including the reference's error messages — is the specification. */
	struct FqSpan { size_t l1, n1, l2, n2, l4, n4, end; };
	std::vector<FqSpan> spans_;
	double t_scan = 0, t_work = 0;      /* BT_CLI_TIMING */
	bool fast_ok() const { return o.format == FASTQ && f && !first && !keepOrig && pending.empty() && o.trim5 == 0 && o.trim3 == 0 && !o.solexaQuals && !o.phred64Quals && !o.integerQuals; }
	template <class V32, class V8, class V64>          /* (the batch's arrays: vectors over page-locked memory) */
	size_t fast_batch(std::vector<ReadRec> &recs, V32 &seeds, V8 &bseq, V8 &bqual, V64 &boffs,
	                  size_t maxRecs, size_t nth, uint32_t gseed) {
		std::vector<FqSpan> &spans = spans_;        /* (a member: worker threads must see this thread's list) */
		spans.clear();
		const auto tq0 = std::chrono::steady_clock::now();
		const char *base = buf.data();
		size_t p = pos;
		while (spans.size() < maxRecs && p < len) {
			FqSpan sp; size_t q = p;
			if (base[q] != '@') break;
			const char *nl = (const char *)memchr(base + q, '\n', len - q); if (!nl) break;
			sp.l1 = q + 1; sp.n1 = (size_t)(nl - base) - q - 1; q = (size_t)(nl - base) + 1;
			nl = (const char *)memchr(base + q, '\n', len - q); if (!nl) break;
			sp.l2 = q; sp.n2 = (size_t)(nl - base) - q; q = (size_t)(nl - base) + 1;
			if (q >= len || base[q] != '+') break;
			nl = (const char *)memchr(base + q, '\n', len - q); if (!nl) break;
			q = (size_t)(nl - base) + 1;
			nl = (const char *)memchr(base + q, '\n', len - q); if (!nl) break;
			sp.l4 = q; sp.n4 = (size_t)(nl - base) - q; q = (size_t)(nl - base) + 1;
			if (sp.n4 == 0 || sp.n2 == 0 || base[sp.l1 + sp.n1 - (sp.n1 ? 1 : 0)] == '\r' || base[sp.l2 + sp.n2 - 1] == '\r' || base[sp.l4 + sp.n4 - 1] == '\r') break;
			sp.end = q;
			spans.push_back(sp);
			p = q;
		}
		if (!spans.empty()) {                       /* the last record is next()'s unless three more newlines are in sight (tail_aborts) */
			size_t q = p; int nl = 0;
			while (nl < 3) { const char *e = (const char *)memchr(base + q, '\n', len - q); if (!e) break; nl++; q = (size_t)(e - base) + 1; }
			if (nl < 3) spans.pop_back();
		}
		const size_t n = spans.size();
		if (n == 0) return 0;
		const auto tq1 = std::chrono::steady_clock::now();
		t_scan += std::chrono::duration<double>(tq1 - tq0).count();
		const size_t r0 = recs.size();
		recs.resize(r0 + n); seeds.resize(r0 + n);
		/* every record of the run has as many bases as its sequence line is long (anything else ends the run below), so the
		 * offsets are known before the lines are touched and the threads write straight into the batch's arrays */
		const uint64_t at0 = boffs.back();
		boffs.resize(r0 + n + 1);
		for (size_t k = 0; k < n; k++) boffs[r0 + k + 1] = boffs[r0 + k] + spans[k].n2;
		bseq.resize((size_t)boffs[r0 + n]); bqual.resize((size_t)boffs[r0 + n]);
		std::vector<uint8_t> okv(n, 1);
		if (nth > n / 2048 + 1) nth = n / 2048 + 1;
		const uint64_t id0 = rdid;
		auto work = [&](size_t lo, size_t hi) {
			for (size_t k = lo; k < hi; k++) {
				const FqSpan &sp = spans[k];
				ReadRec &r = recs[r0 + k];
				if (sp.n1) r.name.assign(base + sp.l1, sp.n1); else r.name = std::to_string(id0 + k);
				r.seq.clear(); r.qual.clear(); r.orig.clear();
				uint8_t *sq = bseq.data() + boffs[r0 + k], *ql = bqual.data() + boffs[r0 + k];
				bool ok = (sp.n2 == sp.n4) && (sp.n1 == 0 || !memchr(base + sp.l1, '\r', sp.n1));
				uint32_t rseed = (gseed + 101u) * 59u * 61u * 67u * 71u * 73u * 79u * 83u;      /* genRandSeed (pat.cpp:21-57), fused into the conversion */
				for (size_t i = 0; ok && i < sp.n2; i++) {
					const uint8_t code = alpha_code[(unsigned char)base[sp.l2 + i]];
					if (code == 255) { ok = false; break; }
					sq[i] = code;
					rseed ^= ((uint32_t)code << ((i & 15) << 1));
				}
				for (size_t i = 0; ok && i < sp.n4; i++) {
					const unsigned char q = (unsigned char)base[sp.l4 + i];
					if (q < 33) { ok = false; break; }
					ql[i] = q;
					rseed ^= ((uint32_t)q << ((i & 3) << 3));
				}
				if (!ok) { okv[k] = 0; continue; }
				for (size_t i = 0; i < r.name.size(); i++) rseed ^= ((uint32_t)(uint8_t)r.name[i] << ((i & 3) << 3));
				seeds[r0 + k] = rseed;
			}
		};
		if (nth <= 1) work(0, n);
		else {
			std::vector<std::thread> th;
			for (size_t t = 0; t < nth; t++) th.emplace_back(work, n * t / nth, n * (t + 1) / nth);
			for (auto &x : th) x.join();
		}
		t_work += std::chrono::duration<double>(std::chrono::steady_clock::now() - tq1).count();
		size_t good = 0;
		while (good < n && okv[good]) good++;
		recs.resize(r0 + good); seeds.resize(r0 + good);
		boffs.resize(r0 + good + 1); bseq.resize((size_t)boffs[r0 + good]); bqual.resize((size_t)boffs[r0 + good]);
		(void)at0;
		pos = good ? spans[good - 1].end : pos;
		rdid += good;
		return good;
	}
	/* FastqPatternSource::parse (pat.cpp:858-975) works on the characters of the 4-newline chunk, not on lines: the first character
	 * is skipped whatever it is, the name ends at the first CR/LF and every CR/LF after it is skipped, the sequence is every letter
	 * or '.' up to the first '+', the rest of the '+' line and the line breaks after it are skipped, and the qualities run to the
	 * next CR/LF — the first quality character is converted unconditionally, which is what turns a misplaced blank line into
	 * "Saw ASCII character 10".  Reads past the end of the chunk (undefined in the reference) see '\n' here. */
	std::string chunk;
	void parse_fastq_chunk(const std::string &ck, ReadRec &r) const {
		const size_t n = ck.size(); size_t cur = 1;
		auto get = [&]() -> int { return cur < n ? (unsigned char)ck[cur++] : (cur++, (int)'\n'); };
		int c;
		r.name.clear(); r.seq.clear(); r.qual.clear();
		for (;;) {
			c = get();
			if (c == '\n' || c == '\r') { do { c = get(); } while ((c == '\n' || c == '\r') && cur <= n); break; }
			r.name.push_back((char)c);
		}
		int nchar = 0;
		while (c != '+' && cur < n) {
			const uint8_t code = alpha_code[c];
			if (code != 255) { if (nchar++ >= o.trim5) r.seq.push_back((char)code); }
			c = get();
		}
		const int trimmed5 = nchar - (int)r.seq.size();
		const int trimmed3 = std::min<int>(o.trim3, (int)r.seq.size());
		r.seq.resize(r.seq.size() - (size_t)trimmed3);
		do { c = get(); } while (c != '\n' && c != '\r');
		while (cur < n && (c == '\n' || c == '\r')) c = get();
		int nqual = 0;
		if (o.integerQuals) {
			/* --integer-quals (pat.cpp:905-923): space-separated integers; this branch of the reference neither trims the 3' end of
			 * the qualities nor compares their number with the sequence's — here they are cut or padded ('I') to the sequence */
			int cur_int = 0;
			while (c != '\t' && c != '\n' && c != '\r') {
				cur_int = (int)((unsigned)cur_int * 10u + (unsigned)(c - '0'));        /* (wraps, as the reference's int does in practice) */
				c = get();
				if (c == ' ' || c == '\t' || c == '\n' || c == '\r') {
					const char cadd = int_to_phred33(cur_int);
					cur_int = 0;
					if (c == ' ') c = get();
					if (++nqual > o.trim5) r.qual.push_back(cadd);
				}
			}
			r.qual.resize(r.seq.size(), 'I');
			return;
		}
		char pc = to_phred33(c, r.name);
		if (nqual++ >= trimmed5) r.qual.push_back(pc);
		while (cur < n) {
			c = get();
			if (c == '\r' || c == '\n') break;
			if (c == ' ') wrong_quality_format(r.name);
			pc = to_phred33(c, r.name);
			if (nqual++ >= trimmed5) r.qual.push_back(pc);
		}
		r.qual.resize(r.qual.size() - std::min<size_t>((size_t)trimmed3, r.qual.size()));
		if (r.qual.size() < r.seq.size()) die("Too few quality values for read: " + r.name + "\n\tare you sure this is a FASTQ-int file?");
		if (r.qual.size() > r.seq.size()) die("Reads file contained a pattern with more than 1024 quality values.\nPlease truncate reads and quality values and and re-run Bowtie");
	}
	/* TabbedPatternSource (pat.cpp:980-1124), --12: name <tab> seq <tab> quals [<tab> seq2 <tab> quals2] per line */
	bool next_tab(ReadRec &a, ReadRec &b, bool &isPair) {
		for (;;) {
			if (!f && !open_next()) return false;
			/* light parse (pat.cpp:979-1011): skip line breaks; a record is the run of characters up to the next LF or CR — an LF is
			 * kept, and so is a CR right after it */
			int c = getc_();
			while (c == '\n' || c == '\r') c = getc_();
			if (c < 0) { gzclose(f); f = NULL; continue; }
			chunk.clear();
			while (c >= 0 && c != '\n' && c != '\r') { chunk.push_back((char)c); c = getc_(); }
			if (c == '\n') { chunk.push_back('\n'); const int d = peek_(); if (d == '\r') { getc_(); chunk.push_back('\r'); } }
			rdid++;
			if (rdid - 1 < o.skipReads) { a.name.clear(); a.seq.clear(); a.qual.clear(); isPair = false; return true; }   /* -s: not parsed */
			if (keepOrig) { a.orig = chunk; b.orig.clear(); }
			/* parse() (pat.cpp:1016-1124): name TAB seq TAB quals [TAB seq2 TAB quals2]; wherever the record runs out before a field
			 * has started, it "ended prematurely" and is skipped — errors of the first end come first */
			const size_t n = chunk.size(); size_t cur = 0;
			int ch = '\t';
			bool ok = true, second = false;
			for (int e = 0; e < 2 && ch == '\t' && ok; e++) {
				ReadRec &r = e ? b : a;
				if (e == 0) {
					r.name.clear();
					ch = (unsigned char)chunk[cur++];
					while (ch != '\t' && cur < n) { r.name.push_back((char)ch); ch = (unsigned char)chunk[cur++]; }
					if (cur >= n) { ok = false; break; }
				} else { r.name = a.name; second = true; }
				int nchar = 0; r.seq.clear();
				ch = (unsigned char)chunk[cur++];
				while (ch != '\t' && cur < n) {
					if (isalpha(ch)) { if (nchar++ >= o.trim5) r.seq.push_back((char)asc2dna[ch]); }
					ch = (unsigned char)chunk[cur++];
				}
				if (cur >= n) { ok = false; break; }
				r.seq.resize(r.seq.size() - std::min<size_t>((size_t)o.trim3, r.seq.size()));
				r.qual.clear(); int nqual = 0;
				ch = (unsigned char)chunk[cur++];
				while (ch != '\t' && ch != '\n' && ch != '\r') {
					if (ch == ' ') wrong_quality_format(r.name);
					/* TabbedPatternSource is built without the quality-encoding flags (ebwt_search.cpp:2942-2943): plain Phred+33 */
					if (ch < 33) die("Saw ASCII character " + std::to_string(ch) + " but expected 33-based Phred qual.");
					if (++nqual > o.trim5) r.qual.push_back((char)ch);
					if (cur >= n) break;
					ch = (unsigned char)chunk[cur++];
				}
				if (nchar > nqual) die("Too few quality values for read: " + r.name + "\n\tare you sure this is a FASTQ-int file?");
				if (nqual > nchar) die("Reads file contained a pattern with more than 1024 quality values.\nPlease truncate reads and quality values and and re-run Bowtie");
				r.qual.resize(r.qual.size() - std::min<size_t>((size_t)o.trim3, r.qual.size()));
			}
			if (!ok) continue;
			isPair = second;
			return true;
		}
	}
	/* Returns false when all input is consumed. */
	/* FastqPatternSource::nextBatchFromFile (pat.cpp:797-856), one record at a time: the next 4-newline chunk of the input, with the
	 * light parser's end-of-file rules applied.  gid = the id the record will get (its slot in a light-parse batch is gid & 15). */
	std::deque<std::string> pending;    /* chunks gathered ahead of parsing (mate files: a light-parse batch at a time) */
	bool fq_gather(std::string &chunk, bool raw, bool mateFile, uint64_t gid) {
		for (;;) {
			if (!f && !open_next()) return false;
				if (first) { int c = peek_(); while (c == '\r' || c == '\n') { getc_(); c = peek_(); }
					if (c != '@') die("Error: reads file does not look like a FASTQ file");          /* an empty file too (pat.cpp:805-812) */
					first = false; }
				/* light parse (nextBatchFromFile): a record is whatever lies up to the fourth newline; EOF stands in for the last
				 * newline, and a record cut short earlier is dropped */
				chunk.clear();
				bool counted = false, aborted = false;
				for (int idx = 0; idx < 4; idx++) {
					if (!getline_(l1)) { if (idx == 3) { chunk += '\n'; counted = true; } else aborted = idx > 0; break; }
					chunk += l1;
					if (line_hit_eof) { if (idx == 3) { chunk += '\n'; counted = true; } else aborted = idx > 0; break; }
					chunk += '\n';
					if (idx == 3) counted = true;
				}
				/* an incomplete record in the first slot of a light-parse batch: the reference's count goes to -1 and it parses what
				 * is in that slot — the incomplete record — which ends in one of parse()'s errors (past its end: this one) */
				if (raw) {                                                             /* --interleaved: il_fill() applies the pair-counting rules */
					rawAborted = aborted;
					if (!counted) { gzclose(f); f = NULL; if (aborted || fileIdx >= files.size()) return false; continue; }
					return true;
				}
				if (aborted && (gid & 15) == 0) {
					if (mateFile) abortedSlot0 = true;                                 /* -1/-2: the two counts (-1 here) are compared first */
					else { ReadRec tmp; parse_fastq_chunk(chunk, tmp); die("Saw ASCII character 10 but expected 33-based Phred qual."); }
				}
				if (!counted) { gzclose(f); f = NULL; continue; }
				/* a file that ends inside a record — one or two newlines after this one, a stray blank line included — makes
				 * nextBatchFromFile step its read count back (pat.cpp:853-855), which discards the record BEFORE the incomplete one
				 * unless that one closed a light-parse batch of 16 */
				if ((gid & 15) != 15 && tail_aborts()) { gzclose(f); f = NULL; continue; }
				return true;
		}
	}
	/* --interleaved (pat.cpp:822-855): a light-parse batch holds up to 16 PAIRS; records are cut alternately into the two mate
	 * buffers and a pair counts once its second record is in.  A file that ends inside a record steps the pair count back by one:
	 * the last complete pair of the batch goes as well (from zero the reference goes on with -1 and parses what the first slot
	 * holds).  A last record without a mate is simply not counted.  Returns the number of pairs made pending. */
	bool rawAborted = false;
	size_t il_fill() {
		std::vector<std::string> got;
		bool aborted = false;
		rawAborted = false;
		while (got.size() < 32) {
			std::string c;
			if (!fq_gather(c, true, false, 0)) { aborted = rawAborted; if (aborted && got.empty()) got.push_back(std::move(c)); break; }
			got.push_back(std::move(c));
		}
		size_t pairs = got.size() / 2;
		if (aborted) {
			if (pairs == 0) { ReadRec tmp; parse_fastq_chunk(got[0], tmp); die("Saw ASCII character 10 but expected 33-based Phred qual."); }   /* slot 0 of the first mate's buffer */
			pairs--;
		}
		for (size_t i = 0; i < 2 * pairs; i++) pending.push_back(std::move(got[i]));
		return pairs;
	}
	/* gathers (without parsing) until `want` records are pending; returns how many are */
	size_t light_fill(size_t want) {
		while (pending.size() < want) {
			std::string c;
			if (!fq_gather(c, false, true, rdid + pending.size())) break;
			pending.push_back(std::move(c));
		}
		return pending.size();
	}
	bool abortedSlot0 = false;          /* a mate file ended inside a record that would have opened a light-parse batch */
	bool skipNext = false;              /* --interleaved: the next record belongs to a pair that -s skips */
	bool next(ReadRec &r, bool mateFile = false) {
		if (o.format == CMDLINE) {
			/* VectorPatternSource (pat.cpp:357-523): "seq[:quals]" becomes the tabbed record "<ordinal> TAB seq TAB quals" — quals
			 * default to one 'I' per character of seq — and is parsed like one: letters only, plain Phred+33, counts must agree */
			for (;;) {
				if (cmdIdx >= files.size()) return false;
				const std::string &tok = files[cmdIdx++];
				const size_t colon = tok.find(':');
				const std::string s = tok.substr(0, colon);
				const std::string q = (colon == std::string::npos || colon + 1 >= tok.size()) ? std::string(s.size(), 'I') : tok.substr(colon + 1);   /* tokenize() drops an empty second token */
				r.name = std::to_string(rdid);
				rdid++;
				if (q.empty()) continue;                                         /* "record ended prematurely": skipped, but it has used its id */
				int nchar = 0, nqual = 0;
				r.seq.clear(); r.qual.clear();
				for (char ch : s) if (isalpha((unsigned char)ch)) { if (nchar++ >= o.trim5) r.seq.push_back((char)asc2dna[(unsigned char)ch]); }
				r.seq.resize(r.seq.size() - std::min<size_t>((size_t)o.trim3, r.seq.size()));
				for (char ch : q) {
					if (ch == '\t' || ch == '\n' || ch == '\r') break;
					if (ch == ' ') wrong_quality_format(r.name);
					if ((unsigned char)ch < 33) die("Saw ASCII character " + std::to_string((int)(unsigned char)ch) + " but expected 33-based Phred qual.");
					if (++nqual > o.trim5) r.qual.push_back(ch);
				}
				if (nchar > nqual) die("Too few quality values for read: " + r.name + "\n\tare you sure this is a FASTQ-int file?");
				if (nqual > nchar) die("Reads file contained a pattern with more than 1024 quality values.\nPlease truncate reads and quality values and and re-run Bowtie");
				r.qual.resize(r.qual.size() - std::min<size_t>((size_t)o.trim3, r.qual.size()));
				return true;
			}
		}
		if (o.format == FASTA_CONT) {
			/* FastaContinuousPatternSource (pat.cpp:651-790): every freq-th window of `len` bases of each FASTA sequence is a read
			 * named <first word of the header>_<offset>; IUPAC codes and '-' count as N, anything else is skipped */
			for (;;) {
				if (!f && !open_next()) return false;
				int c = getc_();
				if (c < 0) { gzclose(f); f = NULL; continue; }
				if (c == '>') {
					fc_reset();
					c = getc_();
					bool sawSpace = false;
					while (c >= 0 && c != '\n' && c != '\r') {
						if (!sawSpace) sawSpace = isspace(c) != 0;
						if (!sawSpace) fcPrefix.push_back((char)c);
						c = getc_();
					}
					while (c == '\n' || c == '\r') c = getc_();
					if (c < 0) { gzclose(f); f = NULL; continue; }
					fcPrefix.push_back('_');
				}
				const bool acgt = strchr("ACGTacgt", c) != NULL && c != 0;
				const bool iupac = strchr("BDHKMNRSVWXYbdhkmnrsvwxy-", c) != NULL && c != 0;
				if (!acgt && !iupac) continue;
				if (iupac) c = 'N';
				fcBuf[fcCur++] = (char)c;
				if (fcCur == 1024) fcCur = 0;
				if (fcEat > 0) { fcEat--; if (!fcBeginning) fcPos++; continue; }
				r.name = fcPrefix + std::to_string(fcPos - fcLast);
				const size_t L = o.fastaContLen;
				r.seq.clear();                                                       /* (this source is built with trim3 = trim5 = 0, pat.h:600-604) */
				for (size_t i = 0; i < L; i++) {
					const char b = (L - i <= fcCur) ? fcBuf[fcCur - (L - i)] : fcBuf[fcCur + 1024 - (L - i)];
					r.seq.push_back((char)asc2dna[(unsigned char)b]);
				}
				r.qual.assign(r.seq.size(), 'I');
				fcEat = o.fastaContFreq - 1; fcPos++; fcBeginning = false;
				rdid++;
				return true;
			}
		}
		if (o.format == FASTQ) {
			/* FastqPatternSource (pat.cpp:797-975): light parse (fq_gather), then parse() */
			if (!pending.empty()) { chunk.swap(pending.front()); pending.pop_front(); }
			else if (!fq_gather(chunk, false, mateFile, rdid)) return false;
			if (rdid < o.skipReads || skipNext) { r.name.clear(); r.seq.clear(); r.qual.clear(); rdid++; return true; }   /* -s: skipped reads are never parsed (pat.cpp:108-110) */
			parse_fastq_chunk(chunk, r);
			if (keepOrig) r.orig = chunk;                                           /* Read::readOrigBuf */
			if (r.name.empty()) r.name = std::to_string(rdid);
			rdid++;
			return true;
		}
		for (;;) {
			if (!f && !open_next()) return false;
			if (o.format == FASTA) {
				/* FastaPatternSource (pat.cpp:531-640).  Light parse: a record is '>' plus everything up to the next '>' — wherever
				 * that is — or EOF.  parse(): the name ends at the first CR/LF, line breaks after it are skipped, the sequence is the
				 * letters and '.' of the next line only, and a record with nothing after its name is skipped (it keeps its id). */
				if (first) {
					int c = getc_();
					if (c < 0) { gzclose(f); f = NULL; continue; }                     /* empty file: no reads, no error */
					while (c == '\r' || c == '\n') c = getc_();
					if (c != '>') die("Error: reads file does not look like a FASTA file");
					first = false; recEof = false;
				} else if (recEof) { gzclose(f); f = NULL; continue; }
				chunk.assign(1, '>');
				for (;;) {
					if (pos >= len && !refill()) { recEof = true; break; }
					const char *b0 = buf.data() + pos;
					const char *g = (const char *)memchr(b0, '>', len - pos);
					if (g) { chunk.append(b0, (size_t)(g - b0)); pos += (size_t)(g - b0) + 1; break; }
					chunk.append(b0, len - pos); pos = len;
				}
				if (recEof && chunk.size() == 1) { gzclose(f); f = NULL; continue; }   /* "immediate EOF case" */
				const size_t n = chunk.size(); size_t cur = 1; int c = -1;
				r.name.clear(); r.seq.clear();
				while (cur < n) {
					c = (unsigned char)chunk[cur++];
					if (c == '\n' || c == '\r') {
						do { c = cur < n ? (unsigned char)chunk[cur] : 0; cur++; } while ((c == '\n' || c == '\r') && cur < n);
						break;
					}
					r.name.push_back((char)c);
				}
				if (cur >= n) { rdid++; continue; }                                    /* "FASTA ended prematurely" */
				int nchar = 0;
				/* the loop tests `cur < buflen` before it takes the character it fetched last: a sequence line that ends at EOF
				 * without a newline loses its last character (pat.cpp:607-619) */
				while (c != '\n' && cur < n) {
					const uint8_t code = alpha_code[c];
					if (code != 255) { if (nchar++ >= o.trim5) r.seq.push_back((char)code); }
					c = (unsigned char)chunk[cur++];
				}
				r.seq.resize(r.seq.size() - std::min<size_t>((size_t)o.trim3, r.seq.size()));
				r.qual.assign(r.seq.size(), 'I');
				if (keepOrig) r.orig = chunk;
			} else {
				/* RawPatternSource (pat.cpp:1129-1213): a record is a non-empty run of characters between CR/LFs — a line without a
				 * single letter is an empty read, not a skipped one — letters only ('.' is not N here), name = ordinal */
				if (rawQ >= rawPieces.size()) {
					rawPieces.clear(); rawQ = 0;
					if (!getline_(l1)) { gzclose(f); f = NULL; continue; }
					size_t p0 = 0;
					for (;;) {
						const size_t e = l1.find('\r', p0);
						const size_t stop = e == std::string::npos ? l1.size() : e;
						if (stop > p0) rawPieces.push_back(l1.substr(p0, stop - p0));
						if (e == std::string::npos) break;
						p0 = e + 1;
					}
					if (rawPieces.empty()) continue;
				}
				const std::string &pc = rawPieces[rawQ++];
				int nchar = 0;
				r.seq.clear();
				for (char ch : pc) if (isalpha((unsigned char)ch)) { if (nchar++ >= o.trim5) r.seq.push_back((char)asc2dna[(unsigned char)ch]); }
				r.seq.resize(r.seq.size() - std::min<size_t>((size_t)o.trim3, r.seq.size()));
				r.qual.assign(r.seq.size(), 'I');
				r.name.clear();
				if (keepOrig) { r.orig = pc; r.orig += '\n'; }
			}
			if (r.name.empty()) r.name = std::to_string(rdid);
			rdid++;
			return true;
		}
	}
};

/* genRandSeed (pat.cpp:21-57) */
static uint32_t gen_rand_seed(const ReadRec &r, uint32_t seed) {
	uint32_t rseed = (seed + 101u) * 59u * 61u * 67u * 71u * 73u * 79u * 83u;
	for (size_t i = 0; i < r.seq.size(); i++) rseed ^= ((uint32_t)(uint8_t)r.seq[i] << ((i & 15) << 1));
	for (size_t i = 0; i < r.qual.size(); i++) rseed ^= ((uint32_t)(uint8_t)r.qual[i] << ((i & 3) << 3));
	for (size_t i = 0; i < r.name.size(); i++) rseed ^= ((uint32_t)(uint8_t)r.name[i] << ((i & 3) << 3));
	return rseed;
}

/* ---------------------------------------------------------------------------------------------- */
/* output                                                                                          */
/* ---------------------------------------------------------------------------------------------- */
struct Out {
	FILE *fp = stdout;
	std::string buf;
	void flush() { if (!buf.empty()) { fwrite(buf.data(), 1, buf.size(), fp); buf.clear(); } }
	void maybe_flush() { if (buf.size() > (1u << 22)) flush(); }
};
static void put_uint(std::string &o, uint64_t v) { char t[24]; int n = snprintf(t, sizeof t, "%llu", (unsigned long long)v); o.append(t, (size_t)n); }
static void put_int(std::string &o, long long v) { char t[24]; int n = snprintf(t, sizeof t, "%lld", v); o.append(t, (size_t)n); }
static void put_upto_ws(std::string &o, const char *s, bool ws) {      /* printUptoWs (hit.h:1265-1276) */
	if (!ws) { o += s; return; }
	size_t n = strcspn(s, " \t");
	o.append(s, n);
}

/* What the formatters need of a read: its name and its slice of the batch's base-code / quality arrays (the arrays the search sees). */
struct RView { const std::string &name; const uint8_t *seq; const char *qual; size_t len; };
static inline void put_qual(std::string &o, const RView &r, bool fw) {
	if (fw) { o.append(r.qual, r.len); return; }
	const size_t at = o.size(); o.resize(at + r.len);
	char *d = &o[at];
	for (size_t i = 0; i < r.len; i++) d[i] = r.qual[r.len - 1 - i];
}
/* the read as it is printed: as given, or reverse-complemented for a '-' hit */
static inline void put_seq(std::string &o, const RView &r, bool fw) {
	const size_t at = o.size(); o.resize(at + r.len);
	char *d = &o[at];
	if (fw) for (size_t i = 0; i < r.len; i++) d[i] = "ACGTN"[r.seq[i]];
	else for (size_t i = 0; i < r.len; i++) { const int c = r.seq[r.len - 1 - i]; d[i] = "ACGTN"[c < 4 ? (c ^ 3) : 4]; }
}

struct HitView { uint32_t tidx, toff, oms, cost, stratum, fw, nmm; const uint32_t *mm; uint32_t mate = 0, mtoff = 0, mfw = 0, mlen = 0; };   /* mate: Hit::mate (0 = unpaired), then Hit::mh.second, mfw, mlen */

/* VerboseHitSink::append (hit.cpp:73-301), partition == 0 */
static void append_default(std::string &o, const Opts &op, const bt_index_t *ix, const RView &r, const HitView &h) {
	size_t field = 0; bool firstfield = true;
	auto sep = [&]() { if (firstfield) firstfield = false; else o += '\t'; };
	if (!op.suppress[field++]) { sep(); o += r.name; }
	if (!op.suppress[field++]) { sep(); o += (h.fw ? '+' : '-'); }
	if (!op.suppress[field++]) {
		sep();
		const char *nm = op.refIdx ? NULL : bt_index_refname(ix, h.tidx);
		if (nm) put_upto_ws(o, nm, !op.fullRef); else put_uint(o, h.tidx);
	}
	if (!op.suppress[field++]) { sep(); put_int(o, (long long)h.toff + op.offBase); }
	if (!op.suppress[field++]) {
		sep();
		put_seq(o, r, h.fw != 0);
	}
	if (!op.suppress[field++]) {
		sep();
		put_qual(o, r, h.fw != 0);
	}
	if (!op.suppress[field++]) { sep(); put_uint(o, h.oms); }
	if (!op.suppress[field++]) {
		sep();
		/* mismatches in increasing offset from the 5' end; refc printed as stored, read char from the printed sequence */
		uint32_t ord[1024]; uint32_t n = h.nmm < 1024 ? h.nmm : 1024;
		for (uint32_t i = 0; i < n; i++) ord[i] = h.mm[i];
		std::sort(ord, ord + n, [](uint32_t a, uint32_t b) { return (a & 0xffffu) < (b & 0xffffu); });
		for (uint32_t i = 0; i < n; i++) {
			uint32_t pos = ord[i] & 0xffffu, refc = (ord[i] >> 16) & 0xff;
			if (i) o += ',';
			put_uint(o, pos);
			/* qryChar = fw ? patSeq[i] : patSeq[len-i-1] where patSeq is the printed (possibly rc) sequence */
			int c = r.seq[pos]; char q = h.fw ? "ACGTN"[c] : "ACGTN"[c < 4 ? (c ^ 3) : 4];
			o += ':'; o += "ACGT"[refc & 3]; o += '>'; o += q;
		}
	}
	if (op.printCost) {
		if (!op.suppress[field++]) { sep(); put_uint(o, h.stratum); }
		if (!op.suppress[field++]) { sep(); put_uint(o, h.cost); }
	}
	o += '\n';
}

static void append_qname(std::string &o, const Opts &op, const std::string &name) {
	for (char ch : name) { if (!op.noQnameTrunc && isspace((unsigned char)ch)) break; o += ch; }
}

/* SAMHitSink::append (sam.cpp:129-257), unpaired */
static void append_sam(std::string &o, const Opts &op, const bt_index_t *ix, const RView &r, const HitView &h, int mapq, int xms) {
	const size_t len = r.len;
	append_qname(o, op, h.mate ? r.name.substr(0, r.name.size() >= 2 ? r.name.size() - 2 : 0) : r.name);
	uint32_t flags = h.fw ? 0 : 16;
	if (h.mate == 1) flags |= 1 | 64 | 2; else if (h.mate == 2) flags |= 1 | 128 | 2;      /* PAIRED | FIRST/SECOND_IN_PAIR | MAPPED_PAIRED */
	if (h.mate && !h.mfw) flags |= 32;                                                       /* MATE_STRAND */
	o += '\t'; put_uint(o, flags); o += '\t';
	const char *nm = op.refIdx ? NULL : bt_index_refname(ix, h.tidx);
	if (nm) put_upto_ws(o, nm, !op.fullRef); else put_uint(o, h.tidx);
	o += '\t'; put_uint(o, (uint64_t)h.toff + 1);
	o += '\t'; put_int(o, mapq);
	o += '\t'; put_uint(o, len); o += 'M';
	if (h.mate) {
		o += "\t=\t"; put_uint(o, (uint64_t)h.mtoff + 1); o += '\t';
		long long ins;
		if (h.toff > h.mtoff) ins = -((long long)h.toff - (long long)h.mtoff + (long long)len);
		else ins = (long long)h.mtoff - (long long)h.toff + (long long)h.mlen;
		put_int(o, (int)ins); o += '\t';
	} else o += "\t*\t0\t0\t";
	put_seq(o, r, h.fw != 0);
	o += '\t';
	put_qual(o, r, h.fw != 0);
	o += "\tXA:i:"; put_uint(o, h.stratum);
	o += "\tMD:Z:";
	/* mms[] is indexed from the 5' end; MD runs along the reference: 5'->3' for fw, reversed for rc */
	static thread_local std::vector<int8_t> refAt;
	refAt.assign(len, -1);
	for (uint32_t i = 0; i < h.nmm; i++) { uint32_t pos = h.mm[i] & 0xffffu; if (pos < len) refAt[pos] = (int8_t)((h.mm[i] >> 16) & 3); }
	int nmcnt = 0, run = 0;
	if (h.fw) { for (size_t i = 0; i < len; i++) { if (refAt[i] >= 0) { nmcnt++; put_int(o, run); o += "ACGT"[(int)refAt[i]]; run = 0; } else run++; } }
	else { for (size_t i = len; i > 0; i--) { if (refAt[i - 1] >= 0) { nmcnt++; put_int(o, run); o += "ACGT"[(int)refAt[i - 1]]; run = 0; } else run++; } }
	put_int(o, run);
	o += "\tNM:i:"; put_int(o, nmcnt);
	if (xms > 0) { o += "\tXM:i:"; put_int(o, xms); }
	o += '\n';
}

/* SAMHitSink::reportUnOrMax (sam.cpp:57-124), unpaired, un == true */
static void append_sam_unaligned(std::string &o, const Opts &op, const RView &r, int mate = 0) {
	append_qname(o, op, mate ? r.name.substr(0, r.name.size() >= 2 ? r.name.size() - 2 : 0) : r.name);
	o += mate == 0 ? "\t4" : mate == 1 ? "\t77" : "\t141";                               /* UNMAPPED [| PAIRED | FIRST/SECOND | MATE_UNMAPPED] */
	o += "\t*\t0\t0\t*\t*\t0\t0\t";
	put_seq(o, r, true);
	o += '\t'; o.append(r.qual, r.len);
	o += "\tXM:i:0\n";
}

/* SAMHitSink::appendHeaders (sam.cpp:20-50) */
static void sam_headers(std::string &o, const Opts &op, const bt_index_t *ix, uint32_t nrefs) {
	o += "@HD\tVN:1.0\tSO:unsorted\n";
	if (!op.samNoSQ) {
		for (uint32_t i = 0; i < nrefs; i++) {
			o += "@SQ\tSN:";
			const char *nm = op.refIdx ? NULL : bt_index_refname(ix, i);
			if (nm) put_upto_ws(o, nm, !op.fullRef); else put_uint(o, i);
			o += "\tLN:"; put_uint(o, bt_index_reflen(ix, i)); o += '\n';
		}
	}
	if (!op.rgs.empty()) { o += "@RG\t"; o += op.rgs; o += '\n'; }
	o += "@PG\tID:Bowtie\tVN:" BOWTIE_VERSION "\tCL:\""; o += op.argstr; o += "\"\n";
}

/* ---------------------------------------------------------------------------------------------- */
/* batches                                                                                         */
/* ---------------------------------------------------------------------------------------------- */
/* The arrays that travel to and from the device live in page-locked memory (bt_host_alloc): bt_context_align_async's copies are
 * asynchronous only then, which is what lets the search of batch k+1 overlap the formatting of batch k. */
template <class T> struct PinAlloc {
	typedef T value_type;
	PinAlloc() {}
	template <class U> PinAlloc(const PinAlloc<U> &) {}
	T *allocate(size_t n) { void *p = bt_host_alloc(n * sizeof(T)); if (!p) throw std::bad_alloc(); return (T *)p; }
	void deallocate(T *p, size_t) { bt_host_free(p); }
	template <class U> bool operator==(const PinAlloc<U> &) const { return true; }
	template <class U> bool operator!=(const PinAlloc<U> &) const { return false; }
};
template <class T> using PinVec = std::vector<T, PinAlloc<T>>;

struct Batch {
	bool paired = false;           /* every unit of the batch is a pair (mates adjacent) */
	std::vector<ReadRec> reads;
	PinVec<uint8_t> seq, qual; PinVec<uint64_t> offs; PinVec<uint32_t> seeds;
	PinVec<uint32_t> found, flags, hits;
	uint32_t slots = 1, mm_cap = 8;
	bt_context_t *cx = NULL;
	bool inflight = false;
};

int main(int argc, char **argv) {
	for (int i = 0; i < 256; i++) asc2dna[i] = 4;
	asc2dna['A'] = asc2dna['a'] = 0; asc2dna['C'] = asc2dna['c'] = 1; asc2dna['G'] = asc2dna['g'] = 2; asc2dna['T'] = asc2dna['t'] = 3;
	for (int i = 0; i < 256; i++) alpha_code[i] = isalpha(i) ? asc2dna[i] : 255;
	alpha_code['.'] = 4;
	Opts op;
	parse_options(argc, argv, op);
	auto t_start = std::chrono::steady_clock::now();

	bt_policy_t pol; bt_policy_init(&pol);
	pol.mode = op.maqLike ? 1 : 0; pol.mms = op.maqLike ? op.seedMms : op.mismatches;
	pol.seed_len = op.seedLen; pol.qual_thresh = (uint32_t)op.qualThresh; pol.max_bts = (uint32_t)op.maxBts;
	pol.khits = op.khits; pol.mhits = op.mhits; pol.all_hits = op.allHits; pol.nofw = op.nofw; pol.norc = op.norc; pol.maq_round = !op.noMaqRound;
	pol.best = op.best; pol.strata = op.strata; pol.max_bts_best = (uint32_t)op.maxBtsBest; pol.sample_max = op.sampleMax;
	const bool pairedInput = !op.mates1.empty() || !op.interleaved.empty(), interleaved = !op.interleaved.empty(), tabbed = !op.tabbed.empty();
	bt_policy_t polU = pol, polP = pol;                                           /* unpaired reads / pairs (a --12 file can hold both) */
	{
		/* aligner.h:975-990: the insert window shrinks by the bases trimmed from the outer ends of the fragment */
		const int adj = (op.mate1fw ? op.trim5 : op.trim3) + (op.mate2fw ? op.trim3 : op.trim5);
		polP.paired = 1; polP.mate1fw = op.mate1fw; polP.mate2fw = op.mate2fw; polP.pair_tries = op.pairTries; polP.best = op.bestPaired;
		polP.min_ins = (uint32_t)std::max(0, (int)op.minInsert - adj); polP.max_ins = (uint32_t)std::max(0, (int)op.maxInsert - adj);
	}
	if (tabbed) polU.best = 1;                                                    /* any paired input makes the whole run stateful (ebwt_search.cpp:3001-3002): single reads go through UnpairedAlignerV2 */
	const bool needMirror = op.maqLike || op.mismatches > 0;

	/* adjustEbwtBase (ebwt.cpp:36-85): as given, else under $BOWTIE_INDEXES */
	std::string base = op.ebwtFile;
	{
		FILE *t = fopen((base + ".1.ebwt").c_str(), "rb");
		if (!t) { const char *e = getenv("BOWTIE_INDEXES"); if (e) { std::string b2 = std::string(e) + "/" + base; FILE *t2 = fopen((b2 + ".1.ebwt").c_str(), "rb"); if (t2) { fclose(t2); base = b2; } } }
		else fclose(t);
	}
	bt_index_t *ix = NULL;
	if (bt_index_load(base.c_str(), needMirror, op.device, &ix)) die(std::string("Error: ") + bt_last_error());
	bt_index_info_t info; bt_index_info(ix, &info);
	auto t_loaded = std::chrono::steady_clock::now();

	Out out;
	if (!op.outfile.empty()) { out.fp = fopen(op.outfile.c_str(), "wb"); if (!out.fp) die("Error: could not open alignment output file " + op.outfile); }
	if (op.sam && !op.samNoHead) sam_headers(out.buf, op, ix, info.n_refs);

	Reader rd(op, tabbed ? op.tabbed : interleaved ? op.interleaved : pairedInput ? op.mates1 : op.queries, (tabbed || interleaved) ? NULL : pairedInput ? &op.qualities1 : &op.qualities), rd2(op, op.mates2, &op.qualities2);
	enum { NB = 3 };                                                              /* one batch being parsed, one on the GPU, one being formatted */
	Batch bt[NB];
	const uint32_t nlim = op.allHits ? 0xffffffffu : op.khits;
	for (auto &b : bt) {
		if (bt_context_create(ix, &b.cx)) die(std::string("Error: ") + bt_last_error());
		b.mm_cap = op.maqLike ? 10 : (uint32_t)std::max(1, op.mismatches);
	}
	uint64_t numAligned = 0, numUnaligned = 0, numMaxed = 0, numReported = 0, numReportedPaired = 0;
	const size_t fmtThreads = op.nthreads > 1 ? (size_t)op.nthreads : std::max<size_t>(1, std::min<size_t>(16, std::thread::hardware_concurrency() / 2));   /* -p: host threads that format output */
	/* HitSink::dumpAlign / dumpUnal / dumpMaxed (hit.h:385-492): files are opened when the first read lands in them; pairs go to
	 * <base>_1.<ext> / <base>_2.<ext> (openOf, hit.h:629-660); maxed reads fall back to --un when --max is not given */
	std::map<std::string, FILE *> dumps;
	auto dump_to = [&](const std::string &base, int mate, const std::string &text) {
		std::string nm = base;
		if (mate) { const size_t dot = base.find_last_of('.'); const char *sfx = mate == 1 ? "_1" : "_2"; nm = dot == std::string::npos ? base + sfx : base.substr(0, dot) + sfx + base.substr(dot); }
		FILE *&fp = dumps[nm];
		if (!fp) { fp = fopen(nm.c_str(), "wb"); if (!fp) die((mate ? "Could not open paired-end aligned/unaligned-read file for writing: " : "Could not open single-ended aligned/unaligned-read file for writing: ") + base); }
		fwrite(text.data(), 1, text.size(), fp);
	};
	auto dump_unit = [&](const std::string &base, const Batch &b, size_t i) {
		if (base.empty()) return;
		if (b.paired && tabbed) dump_to(base, 0, b.reads[2 * i].orig);           /* onePairFile_ (ebwt_search.cpp:3205,3215): a --12 line holds both mates */
		else if (b.paired) { dump_to(base, 1, b.reads[2 * i].orig); dump_to(base, 2, b.reads[2 * i + 1].orig); }
		else dump_to(base, 0, b.reads[i].orig);
	};
	ReadRec lrec, lrec2; bool haveLook = false, lookPair = false;                 /* --12: one record of lookahead (a batch is all pairs or all single reads) */
	bool input_done = false;
	ReadRec rec, rec2;
	auto fix_mate_name = [](std::string &name, int i) {                           /* Read::fixMateName (read.h:141-165) */
		const size_t n = name.size();
		if (n < 2 || name[n - 2] != '/' || name[n - 1] != "012"[i]) { name += '/'; name += "012"[i]; }
	};

	/* What the workers say about reads too short to search (the search itself leaves them unaligned): search_1mm_phase1.c:12-15,
	 * search_23mm_phase1.c:13-20, search_seeded_phase1.c:17-21, aligner.h:440-448,744-751 */
	const bool statefulU = polU.best || polU.strata || polU.sample_max || (!op.maqLike && op.mismatches == 3);
	auto short_read_check = [&](const ReadRec &a, const ReadRec *mate, size_t knownLen = (size_t)-1) {
		const size_t la = knownLen != (size_t)-1 ? knownLen : a.seq.size();
		if (mate) {
			if ((la < 4 || mate->seq.size() < 4) && !op.quiet) fprintf(stderr, "Warning: Skipping pair %s because a mate is less than 4 characters long\n", a.name.c_str());
		} else if (statefulU) {
			if (la < 4 && !op.quiet) fprintf(stderr, "Warning: Skipping read %s because it is less than 4 characters long\n", a.name.c_str());
		} else if (op.maqLike) {
			if (la < 4 && !op.quiet) fprintf(stderr, "Warning: Skipping read (%s) because it is less than 4 characters long\n", a.name.c_str());
		} else if (op.mismatches == 1) {
			if (la < 2) die("Error: Reads must be at least 2 characters long in 1-mismatch mode");
		} else if (op.mismatches == 2) {
			if (la < 3) die("Error: Read (" + a.name + ") is less than 3 characters long");
			if (la < 4) die("Error: Read (" + a.name + ") is less than 4 characters long");
		}
	};
	auto fill = [&](Batch &b) {
		b.reads.clear(); b.seq.clear(); b.qual.clear(); b.offs.assign(1, 0); b.seeds.clear();
		b.paired = pairedInput;
		while (!input_done && b.reads.size() < (size_t)op.batch * (b.paired ? 2 : 1)) {
			bool paired = pairedInput;
			if (tabbed) {
				if (!haveLook) {
					if (rd.rdid > op.qUpto) { input_done = true; break; }   /* the record with id qUpto is still read and parsed (GET_READ, ebwt_search.cpp:934-939) */
					if (!rd.next_tab(lrec, lrec2, lookPair)) { input_done = true; break; }
					haveLook = true;
				}
				if (b.reads.empty()) b.paired = lookPair; else if (lookPair != b.paired) break;   /* the other kind starts the next batch */
				rec = lrec; rec2 = lrec2; paired = lookPair; haveLook = false;
			} else {
				if (rd.rdid > op.qUpto) { input_done = true; break; }   /* the record with id qUpto is still read and parsed (GET_READ, ebwt_search.cpp:934-939) */
				if (!pairedInput && rd.fast_ok() && rd.rdid >= op.skipReads && rd.rdid < op.qUpto) {
					/* plain single-end FASTQ: whole runs of records at a time, parsed by several threads (Reader::fast_batch) */
					const size_t r0 = b.reads.size();
					const size_t want = std::min<size_t>((size_t)op.batch - r0, (size_t)(op.qUpto - rd.rdid));
					if (rd.fast_batch(b.reads, b.seeds, b.seq, b.qual, b.offs, want, fmtThreads, op.seed) > 0) {
						for (size_t k = r0; k < b.reads.size(); k++) if (b.offs[k + 1] - b.offs[k] < 4) short_read_check(b.reads[k], NULL, (size_t)(b.offs[k + 1] - b.offs[k]));
						continue;
					}
				}
				if (interleaved) {
					/* the light parser counts pairs: both records are cut out of the file before either is parsed, and a last record
					 * without a mate is dropped unparsed */
					if (rd.pending.empty() && rd.il_fill() == 0) { input_done = true; break; }
					rd.skipNext = rd.rdid < op.skipReads;                              /* the pair is skipped or parsed as one */
					rd.next(rec); rd.next(rec2);
					rd.skipNext = false;
					rd.rdid--;                                                          /* a pair is one read id */
				} else if (paired) {
					/* DualPatternComposer::nextBatch (pat.cpp:164-222) compares what the two files delivered — for FASTQ a light-parse
					 * batch of 16 records from each file at a time, before any of them is parsed */
					if (op.format == FASTQ) {
						if (rd.pending.empty() && rd2.pending.empty()) {
							const size_t na = rd.light_fill(16), nb = rd2.light_fill(16);
							const long ca = na ? (long)na : rd.abortedSlot0 ? -1 : 0, cb = nb ? (long)nb : rd2.abortedSlot0 ? -1 : 0;
							if (ca < cb) die("Error, fewer reads in file specified with -1 than in file specified with -2");
							if (cb < ca) die("Error, fewer reads in file specified with -2 than in file specified with -1");
							if (ca < 0) die("Saw ASCII character 10 but expected 33-based Phred qual.");
							if (ca == 0) { input_done = true; break; }
						}
						rd.next(rec, true); rd2.next(rec2, true);
					} else {
						const bool ga = rd.next(rec, true), gb = rd2.next(rec2, true);
						if (!ga && gb) die("Error, fewer reads in file specified with -1 than in file specified with -2");
						if (ga && !gb) die("Error, fewer reads in file specified with -2 than in file specified with -1");
						if (!ga) { input_done = true; break; }
					}
				} else if (!rd.next(rec)) { input_done = true; break; }
			}
			if (rd.rdid - 1 >= op.qUpto) { input_done = true; break; }            /* -u: ... and then dropped */
			if (rd.rdid - 1 < op.skipReads) continue;                              /* -s: skipped reads are not counted */
			if (paired) { fix_mate_name(rec.name, 1); fix_mate_name(rec2.name, 2); }   /* PatternSourcePerThread::finalizePair (pat.cpp:75-87) */
			if (rec.seq.size() < 4 || (paired && rec2.seq.size() < 4)) short_read_check(rec, paired ? &rec2 : NULL);
			for (int m = 0; m < (paired ? 2 : 1); m++) {
				ReadRec &rr = m ? rec2 : rec;
				b.seq.insert(b.seq.end(), rr.seq.begin(), rr.seq.end());
				b.qual.insert(b.qual.end(), rr.qual.begin(), rr.qual.end());
				b.offs.push_back(b.seq.size());
				b.seeds.push_back(gen_rand_seed(rr, op.seed));
				b.reads.push_back(std::move(rr));
			}
		}
	};
	auto launch = [&](Batch &b) {
		if (b.reads.empty()) return;
		const uint32_t mult = b.paired ? 2u : 1u;                                  /* HitSinkPerThreadFactory::createMult */
		const bt_policy_t &pol = b.paired ? polP : polU;
		b.slots = op.allHits ? 8 : op.khits * mult;
		if (op.sampleMax && op.mhits != 0xffffffffu) b.slots = std::max(b.slots, op.mhits * mult);   /* -M keeps every hit up to the ceiling */
		b.slots = std::min<uint32_t>(b.slots, 16);                               /* first pass: bounded records per read (-k 1000 must not size n x 1000 records); reads with more come back with BT_OVF_HITS and are re-run below with exact capacities */
		const size_t n = b.reads.size() / mult, rw = BT_HIT_HDR_WORDS + b.mm_cap;
		b.found.resize(n); b.flags.resize(n); b.hits.resize(n * b.slots * rw);    /* the library overwrites every entry it is asked for */
		bt_read_batch_t in; memset(&in, 0, sizeof in);
		in.nreads = (uint32_t)b.reads.size(); in.seq = b.seq.data(); in.qual = b.qual.data(); in.offs = b.offs.data(); in.seeds = b.seeds.data();
		bt_hit_batch_t ho = { b.found.data(), b.flags.data(), b.hits.data(), b.slots, b.mm_cap };
		if (bt_context_align_async(b.cx, &pol, &in, &ho, NULL)) die(std::string("Error: ") + bt_last_error());
		b.inflight = true;
	};
	auto finish = [&](Batch &b) {
		if (!b.inflight) return;
		if (bt_context_sync(b.cx, NULL)) die(std::string("Error: ") + bt_last_error());
		b.inflight = false;
		const bool paired = b.paired; const uint32_t mult = paired ? 2u : 1u;
		const bt_policy_t &pol = paired ? polP : polU;
		const size_t n = b.reads.size() / mult;                                    /* work units: reads or pairs */
		const uint32_t nlimU = (nlim == 0xffffffffu) ? nlim : nlim * mult, mhitsU = (op.mhits == 0xffffffffu) ? op.mhits : op.mhits * mult;
		size_t rw = BT_HIT_HDR_WORDS + b.mm_cap;
		/* reads whose records did not fit: run them again with exact capacities (ABI contract) */
		std::vector<uint32_t> need;
		for (size_t i = 0; i < n; i++) if (b.flags[i] & (BT_OVF_HITS | BT_OVF_MM)) need.push_back((uint32_t)i);
		for (size_t i = 0; i < n; i++) if (b.flags[i] & (BT_OVF_STACK | BT_OVF_FRAME | BT_OVF_PART)) die("Error: search scratch exhausted for read " + b.reads[i * mult].name);
		/* the retried reads go through compact sub-batches of bounded size (a read reported with -a can have millions of hits);
		 * their records are kept in hits2 at off2[k], rw2[k] words each */
		std::vector<uint32_t> hits2, found2(need.size(), 0); std::vector<size_t> off2(need.size(), 0), rw2(need.size(), 0);
		if (!need.empty()) {
			const uint32_t storeLim = op.sampleMax ? std::max(nlimU, mhitsU) : nlimU;
			std::vector<uint32_t> order(need.size());
			for (size_t k = 0; k < need.size(); k++) order[k] = (uint32_t)k;
			std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return b.found[need[x]] < b.found[need[y]]; });
			const size_t budgetWords = (size_t)1 << 26;
			size_t g0 = 0;
			while (g0 < order.size()) {
				size_t g1 = g0; uint32_t gSlots = 1, gLen = 1;
				while (g1 < order.size()) {
					const uint32_t i = need[order[g1]];
					uint32_t rl = (uint32_t)(b.offs[i * mult + 1] - b.offs[i * mult]); if (paired) rl = std::max<uint32_t>(rl, (uint32_t)(b.offs[i * mult + 2] - b.offs[i * mult + 1]));
					const uint32_t sl = std::max<uint32_t>(1, std::min(b.found[i], storeLim)), ln = std::max<uint32_t>(gLen, rl);
					if (g1 > g0 && (g1 - g0 + 1) * (size_t)sl * (BT_HIT_HDR_WORDS + ln) > budgetWords) break;
					gSlots = sl; gLen = ln; g1++;                                /* sorted by found: the last read sets the slot count */
				}
				const size_t gn = g1 - g0, grw = BT_HIT_HDR_WORDS + gLen;
				std::vector<uint8_t> gseq, gqual; std::vector<uint64_t> goffs(1, 0); std::vector<uint32_t> gseeds, gfound(gn, 0), gflags(gn, 0), ghits(gn * (size_t)gSlots * grw, 0);
				for (size_t k = g0; k < g1; k++) {
					for (uint32_t m = 0; m < mult; m++) {
						const uint32_t i = need[order[k]] * mult + m;
						gseq.insert(gseq.end(), b.seq.begin() + b.offs[i], b.seq.begin() + b.offs[i + 1]);
						gqual.insert(gqual.end(), b.qual.begin() + b.offs[i], b.qual.begin() + b.offs[i + 1]);
						goffs.push_back(gseq.size()); gseeds.push_back(b.seeds[i]);
					}
				}
				bt_read_batch_t in; memset(&in, 0, sizeof in);
				in.nreads = (uint32_t)(gn * mult); in.seq = gseq.data(); in.qual = gqual.data(); in.offs = goffs.data(); in.seeds = gseeds.data();
				bt_hit_batch_t ho = { gfound.data(), gflags.data(), ghits.data(), gSlots, gLen };
				if (bt_context_align(b.cx, &pol, &in, &ho, NULL)) die(std::string("Error: ") + bt_last_error());
				for (size_t k = g0; k < g1; k++) {
					const size_t j = k - g0, kk = order[k];
					if (gflags[j] & (BT_OVF_STACK | BT_OVF_FRAME | BT_OVF_PART | BT_OVF_HITS | BT_OVF_MM)) die("Error: search scratch exhausted for read " + b.reads[need[kk] * mult].name);
					const uint32_t nst = std::min(gfound[j], gSlots);
					found2[kk] = gfound[j]; off2[kk] = hits2.size(); rw2[kk] = grw;
					hits2.insert(hits2.end(), ghits.begin() + j * (size_t)gSlots * grw, ghits.begin() + (j * (size_t)gSlots + nst) * grw);
				}
				g0 = g1;
			}
		}
		/* formats the units [lo, hi) into `obuf`; cnt = { aligned, unaligned, maxed, reported, reportedPaired } */
		auto emit_range = [&](size_t lo, size_t hi, std::string &obuf, uint64_t cnt[5], bool serial) {
		size_t ni = (size_t)(std::lower_bound(need.begin(), need.end(), (uint32_t)lo) - need.begin());
		auto view = [&](size_t idx) { return RView{ b.reads[idx].name, b.seq.data() + b.offs[idx], (const char *)b.qual.data() + b.offs[idx], (size_t)(b.offs[idx + 1] - b.offs[idx]) }; };
		auto rlen = [&](size_t idx) { return (uint32_t)(b.offs[idx + 1] - b.offs[idx]); };
		for (size_t i = lo; i < hi; i++) {
			const RView r = view(i * mult);
			const uint32_t *recs = &b.hits[i * b.slots * rw]; size_t rwi = rw; uint32_t found = b.found[i];
			if (ni < need.size() && need[ni] == i) { rwi = rw2[ni]; recs = hits2.data() + off2[ni]; found = found2[ni]; ni++; }
			/* HitSinkPerThread::finishRead (hit.h:741-786) */
			const bool maxed = found > mhitsU, unal = (found == 0);
			if (maxed) {
				cnt[2]++;
				if (serial) dump_unit(op.dumpMax.empty() ? op.dumpUn : op.dumpMax, b, i);
				if (op.sampleMax) {
					/* VerboseHitSink::reportMaxed (hit.cpp:16-68) / SAMHitSink::reportMaxed (sam.cpp:263-311): one of the
					 * buffered hits of the best stratum, picked with a fresh RandomSource seeded by the read */
					const uint32_t nbuf = mhitsU;                                 /* hits buffered before the ceiling was exceeded */
					auto stratum_of = [&](uint32_t s) { return (recs[(size_t)s * rwi + 3] >> 16) & 0xff; };
					uint32_t last = b.seeds[i * mult];
					last = 1664525u * last + 1013904223u; uint32_t rr = last >> 16; last = 1664525u * last + 1013904223u; rr ^= last;   /* RandomSource::nextU32 */
					if (!paired) {
						uint32_t num = 1;
						for (uint32_t s = 1; s < nbuf; s++) { if (stratum_of(s) == stratum_of(s - 1)) num++; else break; }
						const uint32_t *w = recs + (size_t)(rr % num) * rwi;
						HitView h = { w[0], w[1], nbuf, w[3] & 0xffffu, (w[3] >> 16) & 0xff, (w[3] >> 24) & 1, w[4], w + BT_HIT_HDR_WORDS };
						if (op.sam) append_sam(obuf, op, ix, r, h, 0, (int)nbuf + 1); else append_default(obuf, op, ix, r, h);
						cnt[0]++; cnt[3]++;
					} else {
						/* pairs: among the couples whose better mate is in the best stratum (hit.cpp:28-54, sam.cpp:275-299) */
						uint32_t bestS = 999, num = 0;
						for (uint32_t s = 0; s + 1 < nbuf; s += 2) { const uint32_t st = std::min(stratum_of(s), stratum_of(s + 1)); if (st < bestS) { bestS = st; num = 1; } else if (st == bestS) num++; }
						const uint32_t pick = rr % num; num = 0;
						for (uint32_t s = 0; s + 1 < nbuf; s += 2) {
							if (std::min(stratum_of(s), stratum_of(s + 1)) != bestS) continue;
							if (num++ != pick) continue;
							for (uint32_t k = s; k < s + 2; k++) {
								const uint32_t *w = recs + (size_t)k * rwi, *mw = recs + (size_t)(k ^ 1) * rwi;
								HitView h = { w[0], w[1], nbuf / 2, w[3] & 0xffffu, (w[3] >> 16) & 0xff, (w[3] >> 24) & 1, w[4], w + BT_HIT_HDR_WORDS };
								h.mate = (w[3] >> 25) & 3; h.mtoff = mw[1]; h.mfw = (mw[3] >> 24) & 1; h.mlen = rlen(i * mult + (2 - h.mate));
								const RView rr2 = view(i * mult + (h.mate - 1));
								if (op.sam) append_sam(obuf, op, ix, rr2, h, 0, (int)(nbuf / 2) + 1); else append_default(obuf, op, ix, rr2, h);
							}
							break;
						}
						cnt[0]++; cnt[4] += 2;
					}
				}
			}
			else if (unal) {
				cnt[1]++;
				if (serial) dump_unit(op.dumpUn, b, i);
				if (op.sam && !op.noUnal) { if (paired && rlen(i * mult + 1) != 0) { append_sam_unaligned(obuf, op, r, 1); append_sam_unaligned(obuf, op, view(i * mult + 1), 2); } else append_sam_unaligned(obuf, op, r); }   /* `paired = !p.bufb().empty()`, sam.cpp:75 */
			} else {
				uint32_t nrep = std::min(found, nlimU);
				for (uint32_t s = 0; s < nrep; s++) {
					const uint32_t *w = recs + (size_t)s * rwi;
					HitView h = { w[0], w[1], w[2], w[3] & 0xffffu, (w[3] >> 16) & 0xff, (w[3] >> 24) & 1, w[4], w + BT_HIT_HDR_WORDS };
					if (op.strata) h.oms = found / mult - 1;                    /* NBestFirstStratHitSinkPerThread::finishReadImpl (hit.h:1099-1108): sz / mult - 1 */
					h.mate = (w[3] >> 25) & 3;
					size_t ridx = i * mult;
					if (h.mate) {                                               /* records come in (upstream, downstream) couples */
						const uint32_t *mw = recs + (size_t)(s ^ 1) * rwi;
						ridx = i * mult + (h.mate - 1);
						h.mtoff = mw[1]; h.mfw = (mw[3] >> 24) & 1; h.mlen = rlen(i * mult + (2 - h.mate));
					}
					const RView rv = view(ridx);
					if (op.sam) append_sam(obuf, op, ix, rv, h, op.defaultMapq, (int)(nrep / mult)); else append_default(obuf, op, ix, rv, h);
				}
				cnt[0]++; if (paired) cnt[4] += nrep; else cnt[3] += nrep;
				if (serial) dump_unit(op.dumpAl, b, i);
			}
			if (serial) { out.buf.swap(obuf); out.maybe_flush(); out.buf.swap(obuf); }
		}
		};
		/* the reference formats inside its -p worker threads; here the units of a batch are formatted by a few host threads into
		 * private buffers that are written out in unit order (read dumps keep one thread: their files are shared) */
		const bool dumping = !op.dumpAl.empty() || !op.dumpUn.empty() || !op.dumpMax.empty();
		const size_t nth = (dumping || n < 16384) ? 1 : std::min<size_t>(fmtThreads, n / 4096);
		uint64_t cnt[5] = { 0, 0, 0, 0, 0 };
		if (nth <= 1) emit_range(0, n, out.buf, cnt, true);
		else {
			std::vector<std::string> bufs(nth); std::vector<std::array<uint64_t, 5>> cs(nth);
			std::vector<std::thread> th;
			for (size_t t = 0; t < nth; t++) th.emplace_back([&, t]() { cs[t].fill(0); emit_range(n * t / nth, n * (t + 1) / nth, bufs[t], cs[t].data(), false); });
			for (auto &x : th) x.join();
			for (size_t t = 0; t < nth; t++) { out.flush(); fwrite(bufs[t].data(), 1, bufs[t].size(), out.fp); for (int k = 0; k < 5; k++) cnt[k] += cs[t][k]; }
		}
		numAligned += cnt[0]; numUnaligned += cnt[1]; numMaxed += cnt[2]; numReported += cnt[3]; numReportedPaired += cnt[4];
	};

	/* Device I/O path (bt_io_parse_fastq / bt_io_align_format): for the common case — one stream of plain
	 * single-end FASTQ, default or SAM output — the file's text goes to the GPU as it is, reads are cut out of it, searched and
	 * formatted there, and what comes back is the output text, in input order.  The host only reads and writes files.  The path
	 * covers well-formed records and the options listed below; it stops at the first record it does not cover (and always before
	 * the last record of the input), leaving the Reader exactly there: everything else — odd records, the end of the file, the
	 * other options — is the host pipeline's, whose behaviour is the specification.  BT_CLI_HOST_IO=1 turns the device path off. */
	double t_dev_io = 0; uint64_t n_dev_io = 0;
	{
		bool anySuppress = false;
		for (size_t i = 0; i < op.suppress.size(); i++) anySuppress = anySuppress || op.suppress[i];
		const bool dev_ok = !getenv("BT_CLI_HOST_IO") && !pairedInput && !tabbed && !interleaved && op.format == FASTQ && !op.allHits && !op.sampleMax &&
			op.khits >= 1 && op.khits <= 16 && !op.refIdx && !op.printCost && !anySuppress && op.dumpAl.empty() && op.dumpUn.empty() && op.dumpMax.empty() &&
			op.skipReads == 0 && op.qUpto == 0xffffffffu && op.trim5 == 0 && op.trim3 == 0 && !op.solexaQuals && !op.phred64Quals && !op.integerQuals;
		if (dev_ok && (rd.f || rd.open_next())) {
			if (rd.first) {                                                          /* as fq_gather does for the first record of a file */
				int c = rd.peek_();
				while (c == '\r' || c == '\n') { rd.getc_(); c = rd.peek_(); }
				if (c == '@') rd.first = false;
			}
			if (!rd.first && rd.fast_ok()) {
				const auto td0 = std::chrono::steady_clock::now();
				/* NIO chunks in flight: the main thread reads the file and cuts records (bt_io_parse_fastq: the cut decides where the next
				 * chunk starts), one thread per chunk searches and formats (bt_io_align_format, synchronous), outputs are written in chunk
				 * order.  A batch's search ends with its slowest reads (~0.8 s per million on an hg19-sized index), so chunks must overlap. */
				enum { NIO_MAX = 8 };
				int NIO = 4;
				if (const char *e = getenv("BT_CLI_IOS")) NIO = std::max(1, std::min<int>(NIO_MAX, atoi(e)));
				struct Job { bt_io_t *io = NULL; bt_context_t *cx = NULL; std::thread th; bool busy = false; int rc = 0; std::string err; const char *text = NULL; uint64_t bytes = 0, cnt[4] = { 0, 0, 0, 0 }, foff = 0; uint64_t rdid0 = 0; uint32_t n = 0; };
				Job jobs[NIO_MAX];
				auto ensure_io = [&](int k) {                                                /* contexts own GBs of scratch: made only for the chunks that exist */
					if (jobs[k].io) return;
					if (k < NB) jobs[k].cx = bt[k].cx;
					else if (bt_context_create(ix, &jobs[k].cx)) die(std::string("Error: ") + bt_last_error());
					if (bt_io_create(jobs[k].cx, &jobs[k].io)) die(std::string("Error: ") + bt_last_error());
				};
				size_t chunk = 192u << 20;                                                   /* ≈ 870 k 100-bp reads; measured on a B200, 2 M reads: 4 x 192 MB 2.5 s, 4 x 64 MB 3.3 s, 1 x 64 MB 4.2 s */
				if (const char *e = getenv("BT_CLI_CHUNK_MB")) chunk = (size_t)std::max(1l, atol(e)) << 20;
				if (rd.buf.size() < chunk) rd.buf.resize(chunk);
				bt_io_format_t fmt; memset(&fmt, 0, sizeof fmt);
				fmt.sam = op.sam; fmt.no_unal = op.noUnal; fmt.no_qname_trunc = op.noQnameTrunc; fmt.full_ref = op.fullRef; fmt.off_base = op.offBase; fmt.mapq = (uint32_t)op.defaultMapq;
				out.flush();
				uint64_t foff = (uint64_t)gztell(rd.f) - (uint64_t)(rd.len - rd.pos);      /* file offset (uncompressed) of the next unconsumed byte */
				bool fallback = false;                                                      /* a chunk the device formatter does not cover: the host pipeline resumes at its first record */
				/* finishes the oldest outstanding chunk: its text goes out, its counters count; returns false if it was not covered */
				auto finish = [&](Job &j) -> bool {
					j.th.join(); j.busy = false;
					if (j.rc == 2) return false;
					if (j.rc) die(std::string("Error: ") + j.err);
					if (j.bytes) fwrite(j.text, 1, (size_t)j.bytes, out.fp);
					numAligned += j.cnt[0]; numUnaligned += j.cnt[1]; numMaxed += j.cnt[2]; numReported += j.cnt[3];
					n_dev_io += j.n;
					return true;
				};
				size_t k = 0, done = 0;                                                      /* chunks started / finished */
				bool more = true;
				double t_rd = 0, t_parse = 0, t_wait = 0; const bool timing = getenv("BT_CLI_TIMING") != NULL;
				auto now = []() { return std::chrono::steady_clock::now(); };
				auto since = [&](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<double>(now() - t0).count(); };
				while (more || done < k) {
					if (more && k - done < (size_t)NIO) {
						auto t0 = now();
						if (rd.pos > 0) { memmove(rd.buf.data(), rd.buf.data() + rd.pos, rd.len - rd.pos); rd.len -= rd.pos; rd.pos = 0; }
						while (!rd.eof && rd.len < chunk) {                                  /* (the buffer may be larger than a chunk: the chunk size is what is in flight per io) */
							const int got = gzread(rd.f, rd.buf.data() + rd.len, (unsigned)std::min<size_t>(chunk - rd.len, (size_t)1 << 30));
							if (got <= 0) { rd.eof = true; break; }
							rd.len += (size_t)got;
						}
						t_rd += since(t0);
						if (rd.len == 0) { more = false; continue; }
						ensure_io((int)(k % NIO));
						Job &j = jobs[k % NIO];
						uint32_t n = 0; uint64_t used = 0; int irregular = 0;
						t0 = now();
						if (bt_io_parse_fastq(j.io, rd.buf.data(), std::min(rd.len, chunk), op.seed, 0xffffffffu, &n, &used, &irregular)) die(std::string("Error: ") + bt_last_error());
						t_parse += since(t0);
						if (n == 0) { more = false; continue; }
						j.foff = foff; j.rdid0 = rd.rdid; j.n = n; j.rc = 0; j.busy = true;
						const bt_policy_t *pp = &polU; const bt_io_format_t *pf = &fmt;
						j.th = std::thread([&j, pp, pf]() {
							j.rc = bt_io_align_format(j.io, pp, pf, &j.text, &j.bytes, j.cnt);
							if (j.rc) j.err = bt_last_error();
						});
						rd.pos = (size_t)used; rd.rdid += n; foff += used;
						k++;
						if (irregular) more = false;
						continue;
					}
					Job &j = jobs[done % NIO];
					auto t0 = now();
					const bool okj = finish(j);
					t_wait += since(t0);
					if (!okj) {
						/* not covered: drop the chunks behind it and rewind the input to this chunk's first record */
						for (size_t q = done + 1; q < k; q++) { jobs[q % NIO].th.join(); jobs[q % NIO].busy = false; }
						if (gzseek(rd.f, (z_off_t)j.foff, SEEK_SET) < 0) die("Error: could not rewind the read file for the host output path");
						rd.len = 0; rd.pos = 0; rd.eof = false; rd.rdid = j.rdid0;
						fallback = true; k = done; more = false;
						break;
					}
					done++;
				}
				(void)fallback;
				if (timing) fprintf(stderr, "device I/O chunks: %d of <= %zu MB in flight, %zu chunks; reading %.2f s, cutting records %.2f s, waiting for search+format and writing %.2f s\n", NIO, chunk >> 20, k, t_rd, t_parse, t_wait);
				for (int q = 0; q < NIO; q++) { if (jobs[q].io) bt_io_free(jobs[q].io); if (q >= NB && jobs[q].cx) bt_context_free(jobs[q].cx); }
				t_dev_io = std::chrono::duration<double>(std::chrono::steady_clock::now() - td0).count();
			}
		}
	}

	/* Three batches in a ring: a parser thread fills batch k+1 while the GPU searches batch k and this thread formats batch k-1
	 * (the reference parses and formats inside its -p worker threads; here the search needs no host thread at all). */
	std::mutex mu; std::condition_variable cv;
	bool filled[NB] = { false, false, false };
	double t_fill = 0, t_launch = 0, t_finish = 0;                               /* BT_CLI_TIMING=1 prints where the host time goes */
	std::thread parser([&]() {
		for (size_t k = 0;; k++) {
			Batch &b = bt[k % NB];
			{ std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return !filled[k % NB]; }); }
			auto t0 = std::chrono::steady_clock::now();
			fill(b);
			t_fill += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			const bool end = b.reads.empty();
			{ std::lock_guard<std::mutex> lk(mu); filled[k % NB] = true; }
			cv.notify_all();
			if (end) break;
		}
	});
	int prev = -1;
	for (size_t k = 0;; k++) {
		Batch &b = bt[k % NB];
		{ std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return filled[k % NB]; }); }
		if (b.reads.empty()) break;                                                /* end of input */
		auto t0 = std::chrono::steady_clock::now();
		launch(b);
		auto t1 = std::chrono::steady_clock::now();
		t_launch += std::chrono::duration<double>(t1 - t0).count();
		if (prev >= 0) {
			finish(bt[prev]);
			t_finish += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
			{ std::lock_guard<std::mutex> lk(mu); filled[prev] = false; }
			cv.notify_all();
		}
		prev = (int)(k % NB);
	}
	if (prev >= 0) { auto t1 = std::chrono::steady_clock::now(); finish(bt[prev]); t_finish += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count(); }
	parser.join();
	if (getenv("BT_CLI_TIMING")) fprintf(stderr, "device I/O path: %llu reads in %.2f s\n", (unsigned long long)n_dev_io, t_dev_io);
	if (getenv("BT_CLI_TIMING")) fprintf(stderr, "host pipeline: parse %.2f s (parser thread; fast path: scan %.2f s, records %.2f s), launch %.2f s, sync+format %.2f s\n", t_fill, rd.t_scan, rd.t_work, t_launch, t_finish);
	out.flush();
	if (out.fp != stdout) fclose(out.fp);
	for (auto &kv : dumps) if (kv.second) fclose(kv.second);
	auto t_end = std::chrono::steady_clock::now();

	/* HitSink::finish (hit.h:270-346); the sink's quiet_ is never set (hit.h:160), so --quiet does not silence the summary */
	{
		const uint64_t alShown = numAligned + (op.sampleMax ? 0 : numMaxed);
		uint64_t tot = alShown + numUnaligned;
		double alPct = 0, unalPct = 0, maxPct = 0;
		if (tot > 0) { alPct = 100.0 * (double)alShown / (double)tot; unalPct = 100.0 * (double)numUnaligned / (double)tot; maxPct = 100.0 * (double)numMaxed / (double)tot; }
		fprintf(stderr, "# reads processed: %llu\n", (unsigned long long)tot);
		fprintf(stderr, "# reads with at least one alignment: %llu (%.2f%%)\n", (unsigned long long)alShown, alPct);
		fprintf(stderr, "# reads that failed to align: %llu (%.2f%%)\n", (unsigned long long)numUnaligned, unalPct);
		if (numMaxed > 0) fprintf(stderr, op.sampleMax ? "# reads with alignments sampled due to -M: %llu (%.2f%%)\n" : "# reads with alignments suppressed due to -m: %llu (%.2f%%)\n", (unsigned long long)numMaxed, maxPct);
		/* HitSink::finish (hit.h:322-337) */
		if (numReported == 0 && numReportedPaired == 0) fprintf(stderr, "No alignments\n");
		else if (numReportedPaired > 0 && numReported == 0) fprintf(stderr, "Reported %llu paired-end alignments\n", (unsigned long long)(numReportedPaired >> 1));
		else if (numReported > 0 && numReportedPaired == 0) fprintf(stderr, "Reported %llu alignments\n", (unsigned long long)numReported);
		else fprintf(stderr, "Reported %llu paired-end alignments and %llu singleton alignments\n", (unsigned long long)(numReportedPaired >> 1), (unsigned long long)numReported);
	}
	if (getenv("BT_CLI_TIMING"))
		fprintf(stderr, "timing: index load %.3f s, reads to output %.3f s\n", std::chrono::duration<double>(t_loaded - t_start).count(), std::chrono::duration<double>(t_end - t_loaded).count());
	if (op.timing) {
		auto secs = [](std::chrono::steady_clock::duration d) { return (long)std::chrono::duration_cast<std::chrono::seconds>(d).count(); };
		long a = secs(t_loaded - t_start), s = secs(t_end - t_loaded), t = secs(t_end - t_start);
		fprintf(stderr, "Time loading index: %02ld:%02ld:%02ld\n", a / 3600, (a / 60) % 60, a % 60);
		fprintf(stderr, "Time searching: %02ld:%02ld:%02ld\n", s / 3600, (s / 60) % 60, s % 60);
		fprintf(stderr, "Overall time: %02ld:%02ld:%02ld\n", t / 3600, (t / 60) % 60, t % 60);
	}
	for (auto &b : bt) bt_context_free(b.cx);
	bt_index_free(ix);
	return 0;
}
