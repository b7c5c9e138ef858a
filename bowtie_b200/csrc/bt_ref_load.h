/*
 * bt_ref_load.h — host-side reader of the bit-pair reference, X.3.ebwt + X.4.ebwt
 * (BitPairReference::BitPairReference, reference.h:20-330; RefRecord, ref_read.h:57-88).
 *
 * X.3.ebwt: int32 endianness sentinel (1), uint32 record count, then per record { uint32 off, uint32 len, uint8 first }:
 * `off` ambiguous characters, then `len` unambiguous ones; `first` starts a new reference sequence.
 * X.4.ebwt: the unambiguous characters of all records, 2 bits each, 4 per byte, low bits first.
 * Only references that own at least one unambiguous character are numbered (they are the ones in the index).
 *
 * Plain C++ (host); shared by bt_lib.cu and the test-only emulation.
 */
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

struct BtHostRef {
	std::vector<uint32_t> recs;          /* 2 words per record */
	std::vector<uint32_t> refRecOffs, refOffs, approxLen;
	std::vector<uint8_t> buf;
	uint32_t nRefs = 0;
};

static inline bool bt_load_ref(const std::string &base, BtHostRef &r, std::string &err) {
	FILE *f3 = fopen((base + ".3.ebwt").c_str(), "rb");
	if (!f3) { err = "Could not open reference-string index file " + base + ".3.ebwt for reading."; return false; }
	uint32_t one = 0, sz = 0;
	if (fread(&one, 4, 1, f3) != 1 || one != 1 || fread(&sz, 4, 1, f3) != 1 || sz == 0) { fclose(f3); err = base + ".3.ebwt: bad header"; return false; }
	std::vector<uint8_t> first(sz);
	r.recs.resize(2 * (size_t)sz);
	for (uint32_t i = 0; i < sz; i++) {
		uint32_t ol[2]; int c;
		if (fread(ol, 4, 2, f3) != 2 || (c = fgetc(f3)) == EOF) { fclose(f3); err = base + ".3.ebwt: truncated"; return false; }
		r.recs[2 * (size_t)i] = ol[0]; r.recs[2 * (size_t)i + 1] = ol[1]; first[i] = c ? 1 : 0;
	}
	fclose(f3);
	uint32_t cumsz = 0, cumlen = 0, unambiglen = 0, maxlen = 0, nrefs = 0;
	for (uint32_t i = 0; i < sz; i++) {
		const uint32_t off = r.recs[2 * (size_t)i], len = r.recs[2 * (size_t)i + 1];
		if (first[i]) {
			if (unambiglen > 0 && maxlen > 1) r.approxLen.push_back(cumlen);
			if (len > 0) { r.refRecOffs.push_back(i); r.refOffs.push_back(cumsz); }
			cumlen = 0; unambiglen = 0; maxlen = 0; nrefs++;
		}
		cumsz += len;
		if (len > 0) cumlen += off + len;
		unambiglen += len;
		if (len > maxlen) maxlen = len;
	}
	r.refRecOffs.push_back(sz); r.refOffs.push_back(cumsz);
	if (unambiglen > 0 && maxlen > 1) r.approxLen.push_back(cumlen);
	r.nRefs = (uint32_t)r.refRecOffs.size() - 1;
	if (r.approxLen.size() < r.nRefs) r.approxLen.resize(r.nRefs, 0);    /* references whose longest stretch is 1 character: not in the index */
	const size_t bytes = ((size_t)cumsz + 3) / 4;
	r.buf.assign(bytes + 4, 0);
	FILE *f4 = fopen((base + ".4.ebwt").c_str(), "rb");
	if (!f4) { err = "Could not open reference-string index file " + base + ".4.ebwt for reading."; return false; }
	const size_t got = fread(r.buf.data(), 1, bytes, f4);
	fclose(f4);
	if (got != bytes) { err = base + ".4.ebwt: truncated"; return false; }
	return true;
}
