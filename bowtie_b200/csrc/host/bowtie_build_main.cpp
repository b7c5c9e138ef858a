/*
 * bowtie-b200-build — the command line of bowtie-build (ebwt_build.cpp:100-300, 488-620) over bt_index_build.
 *
 * Options that only steer the reference's CPU construction (-a/--noauto, -p/--packed, --bmax, --bmaxdivn, --dcv, --nodc,
 * --threads, --seed) are accepted and ignored: they do not change the files.  Options that would change them and are not
 * provided (--ntoa, -r/--noref, -3/--justref, --large-index, colorspace) are refused.
 */
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../../include/bowtie_b200.h"

enum { ARG_BMAX = 256, ARG_BMAXDIVN, ARG_DCV, ARG_NODC, ARG_THREADS, ARG_SEED, ARG_NTOA, ARG_USAGE, ARG_VERSION, ARG_DEVICE, ARG_LARGE, ARG_WRAPPER };
static const char *short_options = "qfcapho:t:r3";
static struct option long_options[] = {
	{"quiet", no_argument, 0, 'q'}, {"noauto", no_argument, 0, 'a'}, {"packed", no_argument, 0, 'p'}, {"bmax", required_argument, 0, ARG_BMAX},
	{"bmaxdivn", required_argument, 0, ARG_BMAXDIVN}, {"dcv", required_argument, 0, ARG_DCV}, {"nodc", no_argument, 0, ARG_NODC},
	{"threads", required_argument, 0, ARG_THREADS}, {"seed", required_argument, 0, ARG_SEED}, {"ntoa", no_argument, 0, ARG_NTOA},
	{"offrate", required_argument, 0, 'o'}, {"ftabchars", required_argument, 0, 't'}, {"noref", no_argument, 0, 'r'}, {"justref", no_argument, 0, '3'},
	{"help", no_argument, 0, 'h'}, {"usage", no_argument, 0, ARG_USAGE}, {"version", no_argument, 0, ARG_VERSION}, {"device", required_argument, 0, ARG_DEVICE},
	{"large-index", no_argument, 0, ARG_LARGE}, {"wrapper", required_argument, 0, ARG_WRAPPER},
	{0, 0, 0, 0}
};

static void usage(FILE *o) {
	fprintf(o, "Usage: bowtie-b200-build [options]* <reference_in> <ebwt_outfile_base>\n"
	           "    reference_in            comma-separated list of files with ref sequences\n"
	           "    ebwt_outfile_base       write Ebwt data to files with this dir/basename\n"
	           "Options:\n"
	           "    -f                      reference files are Fasta (default)\n"
	           "    -c                      reference sequences given on cmd line (as <seq_in>)\n"
	           "    -o/--offrate <int>      SA is sampled every 2^offRate BWT chars (default: 5)\n"
	           "    -t/--ftabchars <int>    # of chars consumed in initial lookup (default: 10)\n"
	           "    --device <int>          CUDA device that sorts the suffixes (default: 0)\n"
	           "    -q/--quiet              no progress output\n"
	           "  (bowtie-build's -a -p --bmax --bmaxdivn --dcv --nodc --threads --seed are accepted and ignored)\n");
}
static void die(const std::string &m) { fprintf(stderr, "%s\n", m.c_str()); exit(1); }

int main(int argc, char **argv) {
	int offRate = 5, ftabChars = 10, device = 0; bool cmdline = false, quiet = false;
	int c, idx = 0;
	while ((c = getopt_long(argc, argv, short_options, long_options, &idx)) != -1) {
		switch (c) {
		case 'f': break;
		case 'c': cmdline = true; break;
		case 'q': quiet = true; break;
		case 'o': offRate = atoi(optarg); if (offRate < 0) die("-o/--offRate arg must be at least 0"); break;
		case 't': ftabChars = atoi(optarg); if (ftabChars < 1) die("-t/--ftabChars arg must be at least 1"); break;
		case 'h': case ARG_USAGE: usage(stdout); return 0;
		case ARG_VERSION: printf("%s version 1.3.1 (B200 index construction)\n", argv[0]); return 0;
		case ARG_DEVICE: device = atoi(optarg); break;
		case 'a': case 'p': case ARG_BMAX: case ARG_BMAXDIVN: case ARG_DCV: case ARG_NODC: case ARG_THREADS: case ARG_SEED: case ARG_WRAPPER: break;
		case ARG_NTOA: die("Error: --ntoa is not supported");
		case 'r': case '3': die("Error: -r/--noref and -3/--justref are not supported: all six index files are always written");
		case ARG_LARGE: die("Error: large (64-bit) indexes are not supported");
		default: usage(stderr); return 1;
		}
	}
	if (optind >= argc) { fprintf(stderr, "No input sequence or sequence file specified!\n"); usage(stderr); return 1; }
	const std::string infile = argv[optind++];
	if (optind >= argc) { fprintf(stderr, "No output file specified!\n"); usage(stderr); return 1; }
	const std::string outfile = argv[optind++];
	if (optind < argc) { fprintf(stderr, "Extra parameter(s) specified: \"%s\"\n", argv[optind]); return 1; }
	std::vector<std::string> files;
	for (size_t a = 0; a <= infile.size();) { size_t b = infile.find(',', a); if (b == std::string::npos) b = infile.size(); if (b > a) files.push_back(infile.substr(a, b - a)); a = b + 1; }
	if (files.empty()) die("Error: no reference sequences");
	std::string tmp;
	if (cmdline) {                                                     /* ebwt_build.cpp:312-324: sequence i becomes the FASTA record ">i" */
		tmp = outfile + ".cmdline.fa.tmp";
		FILE *f = fopen(tmp.c_str(), "wb");
		if (!f) die("Could not open index file for writing: \"" + tmp + "\"");
		for (size_t i = 0; i < files.size(); i++) fprintf(f, ">%zu\n%s\n", i, files[i].c_str());
		fclose(f);
		files.assign(1, tmp);
	}
	std::vector<const char *> ptrs;
	for (auto &s : files) ptrs.push_back(s.c_str());
	if (!quiet) fprintf(stderr, "Building %s.{1,2,3,4,rev.1,rev.2}.ebwt (offrate %d, ftabchars %d)\n", outfile.c_str(), offRate, ftabChars);
	const int rc = bt_index_build(ptrs.data(), (uint32_t)ptrs.size(), outfile.c_str(), offRate, ftabChars, device);
	if (!tmp.empty()) remove(tmp.c_str());
	if (rc) die(bt_last_error());
	return 0;
}
