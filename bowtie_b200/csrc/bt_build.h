/*
 * bt_build.h — index construction: everything of `bowtie-build` except the suffix sort.
 *
 * the hg19-scale configurations need an index builder that does not take CPU-hours.  The reference builds
 * X.{1,2,3,4}.ebwt and X.rev.{1,2}.ebwt in ebwt_build.cpp:303-480 (driver), ref_read.cpp:10-141,202-273 (FASTA -> records),
 * ebwt.h:3825-3983 (joinToDisk), 582-611 (szsToDisk), 3602-3665 (header) and 3985-4388 (buildToDisk: BWT sides, occ words,
 * SA sample, ftab/eftab).  Here the split is:
 *   host (this file, plain C++): FASTA -> unambiguous stretches ("records") and the joined text, X.3, file headers, names,
 *        the prefix sum over the ftab histogram with its eftab exceptions, writing the files;
 *   device (bt_build_sa.cuh): everything that is O(genome): the 2-bit X.4 image, the mirror text, the suffix sort, and what
 *        buildToDisk reads off the suffix array — BWT characters packed into side pairs with their occ words, zOff, the SA
 *        sample, fchr, the ftab k-mer histogram, where the suffixes shorter than ftabChars fall.
 * Files are byte-identical to the reference's for the same input and -o/-t (the parity test); construction-only options
 * (--bmax, --dcv, --threads, --seed, --packed) have no counterpart because they do not change the output.
 *
 * Suffix order (multikey_qsort.h:25-28, blockwise_sa.h): the end of the text compares GREATER than every character, so
 * the empty suffix is the last row and a suffix is greater than any suffix it is a proper prefix of.
 */
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <zlib.h>

struct BtRefRecord { uint32_t off, len; uint8_t first; };          /* RefRecord (ref_read.h:57-88): `off` gap characters, then `len` unambiguous ones */

struct BtRefInfo {
	std::vector<BtRefRecord> recs;
	std::vector<uint32_t> plens;                                   /* length of every sequence that has unambiguous characters, gaps included */
	std::vector<std::string> names;
	std::vector<uint8_t> textStore;                                /* (FASTA path) owns the text */
	const uint8_t *text = nullptr; uint64_t textLen = 0;           /* the joined unambiguous characters, codes 0..3, in file order */
};

struct BtBuildParams { int offRate = 5, ftabChars = 10, lineRate = 6, linesPerSide = 1; bool nsToAs = false; };

/* What the device delivers for one text (length len, rows 0..len). */
struct BtIndexParts {
	std::vector<uint8_t> ebwt;                                     /* the side pairs, as in the file */
	uint32_t zOff = 0;                                             /* the row of suffix 0 */
	uint32_t fchr[5] = { 0, 0, 0, 0, 0 };
	std::vector<uint32_t> offs;                                    /* SA[row] for every row with the low offRate bits clear */
	std::vector<uint32_t> ftab;                                    /* 4^ftabChars + 1: [k + 1] = number of suffixes that start with k-mer k */
	/* suffixes shorter than ftabChars ("absorbed" into the ftab transition that follows them, ebwt.h:4146-4174): for each run
	 * of such rows, the ftabChars-mer of the next longer suffix in row order (or 4^ftabChars if none follows) and the run length */
	std::vector<std::pair<uint32_t, uint32_t>> absorb;
};

static inline int bt_dna_cat(int c) {                              /* dna4Cat (alphabet.cpp:3-25): 1 = ACGT, 2 = IUPAC ambiguity code or '-', 0 = anything else */
	switch (c) {
	case 'A': case 'C': case 'G': case 'T': case 'a': case 'c': case 'g': case 't': return 1;
	case 'B': case 'D': case 'H': case 'K': case 'M': case 'N': case 'R': case 'S': case 'V': case 'W': case 'X': case 'Y':
	case 'b': case 'd': case 'h': case 'k': case 'm': case 'n': case 'r': case 's': case 'v': case 'w': case 'x': case 'y': case '-': return 2;
	default: return 0;
	}
}
static inline uint8_t bt_dna_code(int c) { switch (c) { case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; default: return 0; } }

/* FASTA -> records.  One pass instead of the reference's two (fastaRefReadSizes, then fastaRefReadAppend per record); the
 * cases where those two disagree with each other or with common sense — empty files, sequences without a single unambiguous
 * character, '#' comment lines — are refused rather than imitated. */
static inline bool bt_build_read_fasta(const std::vector<std::string> &files, bool nsToAs, BtRefInfo &R, std::string &err) {
	if (nsToAs) { err = "Error: --ntoa is not supported"; return false; }   /* (the reference's two passes treat it differently) */
	R.recs.clear(); R.plens.clear(); R.names.clear(); R.textStore.clear(); R.text = nullptr; R.textLen = 0;
	uint64_t unambigTot = 0;
	for (const std::string &fn : files) {
		gzFile f = gzopen(fn.c_str(), "rb");
		if (!f) { err = "Error: could not open " + fn; return false; }
		gzbuffer(f, 1 << 20);
		std::vector<char> buf(1 << 22);
		size_t len = 0, pos = 0;
		int lastByte = -1;
		auto get = [&]() -> int { if (pos >= len) { int n = gzread(f, buf.data(), (unsigned)buf.size()); if (n <= 0) return -1; len = (size_t)n; pos = 0; } return lastByte = (unsigned char)buf[pos++]; };
		int c = get();
		while (c >= 0 && (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f')) c = get();
		if (c < 0) { gzclose(f); err = "Error: empty reference file " + fn; return false; }
		if (c != '>') { gzclose(f); err = "Reference file does not seem to be a FASTA file"; return false; }
		while (c == '>') {
			std::string name;
			c = get();
			while (c >= 0 && c != '\n' && c != '\r') { name.push_back((char)c); c = get(); }
			while (c == '\n' || c == '\r') c = get();
			if (c == '#') { gzclose(f); err = "Error: comment lines in the reference are not supported"; return false; }
			/* the sequence: alternating runs of gap characters and unambiguous characters up to the next '>' */
			uint32_t both = 0, unambig = 0, off = 0, run = 0; bool first = true;
			const size_t rec0 = R.recs.size();
			while (c >= 0 && c != '>') {
				int cat = bt_dna_cat(c);
				if (nsToAs && cat == 2) { c = 'A'; cat = 1; }
				if (cat == 1) { R.textStore.push_back(bt_dna_code(c)); run++; }
				else if (cat == 2) {
					if (run) { R.recs.push_back({ off, run, (uint8_t)first }); first = false; both += off + run; unambig += run; off = 0; run = 0; }
					off++;
				}
				c = get();
			}
			if (run) { R.recs.push_back({ off, run, (uint8_t)first }); first = false; both += off + run; unambig += run; off = 0; }
			else if (off == 1 && !first && c < 0 && bt_dna_cat(lastByte) == 2) {
				/* a single gap character as the very last byte of a file is consumed with the stretch before it and the reader is
				 * at EOF before a record can be made of it (ref_read.cpp:123-127,220): it is in neither the records nor plen */
			} else if (off) {
				if (first) { gzclose(f); err = "Error: reference sequence \"" + name + "\" has no unambiguous characters (not supported)"; return false; }
				R.recs.push_back({ off, 0, 0 }); both += off;                /* a trailing gap is a record of its own */
			}
			if (R.recs.size() == rec0) { gzclose(f); err = "Error: empty reference sequence \"" + name + "\" (not supported)"; return false; }
			unambigTot += unambig;
			if (unambigTot > 0xfffffffeull) { gzclose(f); err = "Error: Reference sequence has more than 2^32-1 characters!  Please try to\nbuild a large index instead using the appropiate options."; return false; }
			if (name.empty()) name = std::to_string(R.names.size());     /* ebwt.h:3904-3909 */
			R.names.push_back(name);
			R.plens.push_back(both);
		}
		gzclose(f);
	}
	if (R.textStore.empty()) { err = "Error: No unambiguous stretches of characters in the input.  Aborting..."; return false; }
	R.text = R.textStore.data(); R.textLen = R.textStore.size();
	return true;
}

static inline void bt_put_u32(std::vector<uint8_t> &o, uint32_t v) { o.push_back((uint8_t)v); o.push_back((uint8_t)(v >> 8)); o.push_back((uint8_t)(v >> 16)); o.push_back((uint8_t)(v >> 24)); }
struct BtSpan { const void *p; size_t n; };
static inline bool bt_write_file(const std::string &path, const std::vector<BtSpan> &parts, std::string &err) {
	FILE *f = fopen(path.c_str(), "wb");
	if (!f) { err = "Could not open index file for writing: \"" + path + "\"\nPlease make sure the directory exists and that permissions allow writing by\nBowtie."; return false; }
	bool ok = true;
	for (const BtSpan &s : parts) if (ok && s.n) ok = fwrite(s.p, 1, s.n, f) == s.n;
	ok = (fclose(f) == 0) && ok;
	if (!ok) err = "Error writing " + path;
	return ok;
}

/* X.3.ebwt (the records) and X.4.ebwt (the unambiguous characters, 2 bits each — packed by the device, BsaPack2):
 * ebwt_build.cpp:361-391, filebuf.h:537-590 */
static inline bool bt_build_write_ref(const std::string &base, const BtRefInfo &R, const std::vector<uint8_t> &o4, std::string &err) {
	std::vector<uint8_t> o3;
	bt_put_u32(o3, 1); bt_put_u32(o3, (uint32_t)R.recs.size());
	for (const BtRefRecord &r : R.recs) { bt_put_u32(o3, r.off); bt_put_u32(o3, r.len); o3.push_back(r.first); }
	return bt_write_file(base + ".3.ebwt", { { o3.data(), o3.size() } }, err) && bt_write_file(base + ".4.ebwt", { { o4.data(), o4.size() } }, err);
}

/* X.1.ebwt / X.2.ebwt (or X.rev.1 / X.rev.2) from what the device delivered. */
static inline bool bt_build_write_index(const std::string &path1, const std::string &path2, const BtRefInfo &R, uint32_t len,
                                        const BtBuildParams &P, BtIndexParts &S, std::string &err) {
	/* EbwtParams::init (ebwt.h:138-184) */
	const uint32_t sideSz = (1u << P.lineRate) * (uint32_t)P.linesPerSide, sideBwtSz = sideSz - 8;
	const uint32_t bwtSz = len / 4 + 1;
	const uint32_t numSidePairs = (bwtSz + 2 * sideBwtSz - 1) / (2 * sideBwtSz), numSides = numSidePairs * 2;
	const uint64_t ebwtTotLen = (uint64_t)numSides * sideSz;
	const uint32_t offsLen = (uint32_t)(((uint64_t)len + 1 + (1ull << P.offRate) - 1) >> P.offRate);
	const uint64_t ftabLen = (1ull << (2 * P.ftabChars)) + 1;
	const uint32_t eftabLen = (uint32_t)P.ftabChars * 2;
	if (S.ebwt.size() != ebwtTotLen || S.offs.size() != offsLen || S.ftab.size() != ftabLen) { err = "internal error: the device delivered index parts of the wrong shape"; return false; }

	std::vector<uint8_t> hdr;
	/* header (ebwt.h:3611-3624) */
	bt_put_u32(hdr, 1); bt_put_u32(hdr, len); bt_put_u32(hdr, (uint32_t)P.lineRate); bt_put_u32(hdr, (uint32_t)P.linesPerSide);
	bt_put_u32(hdr, (uint32_t)P.offRate); bt_put_u32(hdr, (uint32_t)P.ftabChars); bt_put_u32(hdr, (uint32_t)-1);
	/* plen[], rstarts[] (joinToDisk ebwt.h:3865-3881, szsToDisk ebwt.h:582-611; per-record reversal leaves them as they are) */
	bt_put_u32(hdr, (uint32_t)R.plens.size());
	for (uint32_t p : R.plens) bt_put_u32(hdr, p);
	uint32_t nFrag = 0;
	for (const BtRefRecord &r : R.recs) if (r.len) nFrag++;
	bt_put_u32(hdr, nFrag);
	{
		uint32_t seq = 0, off = 0, tot = 0;
		for (const BtRefRecord &r : R.recs) {
			if (r.len == 0) continue;
			if (r.first) { off = 0; seq++; }
			off += r.off;
			bt_put_u32(hdr, tot); bt_put_u32(hdr, seq - 1); bt_put_u32(hdr, off);
			tot += r.len; off += r.len;
		}
	}
	std::vector<uint8_t> mid;
	bt_put_u32(mid, S.zOff);
	for (int i = 0; i < 5; i++) bt_put_u32(mid, S.fchr[i]);            /* fchr (ebwt.h:4296-4315) */
	/* ftab / eftab (ebwt.h:4143-4174, 4317-4352): ftab[k] = first row of the suffixes that start with k-mer k; where suffixes
	 * shorter than ftabChars sit between two k-mers' blocks the entry points (x ^ 0xffffffff) at an eftab pair (lo, hi) */
	std::vector<uint32_t> &ftab = S.ftab;
	std::vector<uint32_t> eftab(eftabLen, 0);
	{
		std::vector<uint32_t> absorb_at, absorb_n;
		for (auto &a : S.absorb) { absorb_at.push_back(a.first); absorb_n.push_back(a.second); }
		uint32_t ecur = 0, hiPrev = 0;                                 /* hiPrev = ftabHi(i - 1) */
		size_t ai = 0;
		for (uint64_t i = 1; i < ftabLen; i++) {
			const uint32_t lo = ftab[(size_t)i] + hiPrev;
			while (ai < absorb_at.size() && absorb_at[ai] < i) ai++;       /* (sorted by k-mer; entries at 0 cannot occur) */
			if (ai < absorb_at.size() && absorb_at[ai] == i) {
				const uint32_t hi = lo + absorb_n[ai];
				if (ecur * 2 + 1 >= eftabLen) { err = "internal error: eftab overflow"; return false; }
				eftab[ecur * 2] = lo; eftab[ecur * 2 + 1] = hi;
				ftab[(size_t)i] = ecur ^ 0xffffffffu; ecur++;
				hiPrev = hi;
			} else { ftab[(size_t)i] = lo; hiPrev = lo; }
		}
		if (hiPrev != len + 1) { err = "internal error: ftab does not cover the suffix array"; return false; }
	}
	/* names (ebwt.h:803-811) */
	std::vector<uint8_t> names;
	for (const std::string &n : R.names) { names.insert(names.end(), n.begin(), n.end()); names.push_back('\n'); }
	names.push_back(0);
	const uint32_t one = 1;                                            /* little-endian host */
	return bt_write_file(path1, { { hdr.data(), hdr.size() }, { S.ebwt.data(), S.ebwt.size() }, { mid.data(), mid.size() },
	                              { ftab.data(), ftab.size() * 4 }, { eftab.data(), eftab.size() * 4 }, { names.data(), names.size() } }, err) &&
	       bt_write_file(path2, { { &one, 4 }, { S.offs.data(), S.offs.size() * 4 } }, err);
}

/* A parsed reference from caller-provided pieces (bt_index_build_text): checks what the FASTA reader guarantees by construction. */
static inline bool bt_build_check_ref(const BtRefInfo &R, std::string &err) {
	uint64_t tot = 0; size_t nseq = 0; bool any = false;
	for (size_t i = 0; i < R.recs.size(); i++) {
		const BtRefRecord &r = R.recs[i];
		if (i == 0 && !r.first) { err = "bt_index_build_text: the first record must start a sequence"; return false; }
		if (r.first) { if (r.len == 0) { err = "bt_index_build_text: a sequence must start with a record that has characters"; return false; } nseq++; }
		tot += r.len; any |= r.len != 0;
	}
	if (!any || tot != R.textLen) { err = "bt_index_build_text: the records do not add up to the text length"; return false; }
	if (nseq != R.plens.size() || nseq != R.names.size()) { err = "bt_index_build_text: one name per sequence is required"; return false; }
	return true;
}
