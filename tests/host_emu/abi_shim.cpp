/*
 * tests/host_emu/abi_shim.cpp — TEST-ONLY stand-in for libbowtie_b200.so on machines without a GPU.
 *
 * Exports the C ABI of include/bowtie_b200.h but runs the device state machine (bt_core.cuh) through the
 * host emulation, one lane at a time.  Its only purpose is to let `pytest -m "not gpu"` exercise the HOST
 * code of the product (the bowtie-compatible driver: option parsing, read parsing, default/SAM formatting)
 * against the reference binary.  It is built into tests/host_emu/shim/ and is only ever found through an
 * explicit LD_LIBRARY_PATH set by tests/test_cli_parity.py; the product never loads it.
 */
#define BT_HOST_EMU 1
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../bowtie_b200/csrc/bt_native.cuh"
#include "../../bowtie_b200/csrc/bt_best_prog.h"
#include "../../bowtie_b200/csrc/bt_ref_load.h"
#include "bsa_host.h"
#include "../../bowtie_b200/csrc/bt_io_run.h"
#include "../../include/bowtie_b200.h"
extern "C" {
#include "../../oracle/bt_oracle.h"
}

struct EmuIx { std::vector<uint4> blocks; bto_index *raw; BtDevIndex dev; };
struct bt_index { EmuIx *e[2]; bool mirror; std::vector<std::string> names; bt_stats_t st; std::string base; BtHostRef href; bool ref_loaded = false; };
struct bt_context { bt_index *ix; };
static std::string g_err;

static EmuIx *load_one(const char *base, int mirror) {
	char err[256];
	bto_index *ix = bto_index_load(base, mirror, err, sizeof err);
	if (!ix) { g_err = err; return NULL; }
	EmuIx *e = new EmuIx(); e->raw = ix;
	BtNativeIndex n; n.ebwt = ix->ebwt; n.len = ix->len; n.zOff = ix->zOff; n.zEbwtByteOff = ix->zEbwtByteOff; n.zEbwtBpOff = (uint32_t)ix->zEbwtBpOff;
	memcpy(n.fchr, ix->fchr, sizeof n.fchr);
	uint32_t nb = (ix->len >> 6) + 1; e->blocks.resize(2 * (size_t)nb);
	for (uint32_t k = 0; k < nb; k++) bt_relayout_block(n, k, &e->blocks[2 * (size_t)k]);
	BtDevIndex &d = e->dev;
	d.blocks = e->blocks.data(); d.offs = ix->offs; d.ftab = ix->ftab; d.eftab = ix->eftab; d.rstarts = ix->rstarts; d.plen = ix->plen;
	d.len = ix->len; d.zOff = ix->zOff; d.nFrag = ix->nFrag; d.nPat = ix->nPat; d.offMask = ix->offMask; d.offRate = ix->offRate; d.ftabChars = ix->ftabChars;
	memcpy(d.fchr, ix->fchr, sizeof d.fchr); d.fw = (uint32_t)ix->fw;
	return e;
}

extern "C" {
int bt_abi_version(void) { return BT_ABI_VERSION; }
const char *bt_last_error(void) { return g_err.c_str(); }
int bt_index_load(const char *basename, int need_mirror, int, bt_index_t **out) {
	bt_index *ix = new bt_index(); memset(&ix->st, 0, sizeof ix->st); ix->e[0] = ix->e[1] = NULL; ix->mirror = need_mirror != 0; ix->base = basename;
	ix->e[0] = load_one(basename, 0);
	if (ix->e[0] && need_mirror) ix->e[1] = load_one(basename, 1);
	if (!ix->e[0] || (need_mirror && !ix->e[1])) { delete ix; return 1; }
	bto_index *r = ix->e[0]->raw;
	for (uint32_t i = 0; i < r->nRefnames; i++) ix->names.push_back(r->refnames[i]);
	while (!ix->names.empty() && ix->names.back().empty()) ix->names.pop_back();
	*out = ix; return 0;
}
void bt_index_free(bt_index_t *ix) { if (!ix) return; for (int k = 0; k < 2; k++) if (ix->e[k]) { bto_index_free(ix->e[k]->raw); delete ix->e[k]; } delete ix; }
int bt_index_info(const bt_index_t *ix, bt_index_info_t *info) { bto_index *r = ix->e[0]->raw; info->len = r->len; info->n_refs = r->nPat; info->off_rate = r->offRate; info->ftab_chars = r->ftabChars; info->has_mirror = ix->mirror; info->device_bytes = 0; return 0; }
const char *bt_index_refname(const bt_index_t *ix, uint32_t i) { return i < ix->names.size() ? ix->names[i].c_str() : NULL; }
uint32_t bt_index_reflen(const bt_index_t *ix, uint32_t i) { bto_index *r = ix->e[0]->raw; return i < r->nPat ? r->plen[i] : 0; }
void bt_policy_init(bt_policy_t *p) { memset(p, 0, sizeof *p); p->mode = 1; p->mms = 2; p->seed_len = 28; p->qual_thresh = 70; p->max_bts = 125; p->khits = 1; p->mhits = 0xffffffffu; p->maq_round = 1; p->max_bts_best = 800; p->max_ins = 250; p->mate1fw = 1; p->pair_tries = 100; }

static int run_best(bt_index *ix, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out);
static bool stateful(const bt_policy_t *pol);
static int run(bt_index *ix, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out) {
	if (stateful(pol)) return run_best(ix, pol, in, out);
	BtKParams P; memset(&P, 0, sizeof P);
	P.ix[0] = ix->e[0]->dev; if (ix->e[1]) P.ix[1] = ix->e[1]->dev;
	memcpy(&P.pol, pol, sizeof(BtPolicy));
	bt_build_prog(pol->mode, pol->mms, pol->nofw, pol->norc, P.prog);
	P.seq = in->seq; P.qual = in->qual; P.roff = in->offs; P.seeds = in->seeds;
	P.found = out->found; P.flags = out->flags; P.hits = out->hits; P.slots = out->slots; P.mm_cap = out->mm_cap; P.rec_words = BT_HIT_HDR + out->mm_cap;
	uint32_t maxlen = 1; for (uint32_t i = 0; i < in->nreads; i++) { uint64_t l = in->offs[i + 1] - in->offs[i]; if (l > maxlen) maxlen = (uint32_t)l; }
	P.mask_rows = (maxlen + 255) >> 8;
	uint32_t R = maxlen * maxlen + 8 + (maxlen + 3) * P.mask_rows; if (R > 65000) R = 65000;
	std::vector<uint4> rows(2 * (size_t)R); std::vector<uint8_t> elims(R), stage; std::vector<BtFrame> frames(maxlen + 2); std::vector<uint64_t> parts(1 << 16);
	P.R = R; P.FCAP = maxlen + 2; P.PCAP = 1 << 16;
	BtScratch S = { rows.data(), elims.data(), frames.data(), parts.data() };
	BtLane L; BtLaneCold LK; memset(&L, 0, sizeof L); memset(&LK, 0, sizeof LK); L.K = &LK;
	const uint32_t nwork = in->sel ? in->nsel : in->nreads;
	for (uint32_t w = 0; w < nwork; w++) {
		uint32_t r = in->sel ? in->sel[w] : w;
		bt_begin_read(L, P, r);
		stage.assign(2 * (size_t)L.rlen + 2, 0);
		memcpy(stage.data(), in->seq + in->offs[r], L.rlen); memcpy(stage.data() + L.rlen, in->qual + in->offs[r], L.rlen);
		L.rseq = stage.data(); L.rqual = stage.data() + L.rlen; L.K->hasN = memchr(stage.data(), 4, L.rlen) != NULL;
		while (L.pc != PC_FINISH_READ) { if (BT_IS_FAST(L.pc)) bt_fast_iter(L, P, S); else bt_rare_iter(L, P, S); }
		bt_finish_read(L, P);
	}
	return 0;
}
static bool stateful(const bt_policy_t *pol) { return pol->best || pol->strata || pol->paired || (pol->mode == 0 && pol->mms == 3); }
/* the best-first path (bt_best.cuh), one read at a time */
static int run_best(bt_index *ix, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out) {
	BfKParams P; memset(&P, 0, sizeof P);
	P.ix[0] = ix->e[0]->dev; if (ix->e[1]) P.ix[1] = ix->e[1]->dev;
	memcpy(&P.pol, pol, sizeof(BtPolicy));
	bf_build_prog(pol->mode, pol->mms, pol->seed_len, pol->qual_thresh, pol->nofw, pol->norc, &P.prog,
	              pol->paired, pol->mate1fw, pol->mate2fw, pol->min_ins, pol->max_ins, pol->pair_tries, pol->mhits, 0, pol->best);
	if (pol->paired) {
		if (!ix->ref_loaded) { if (!bt_load_ref(ix->base, ix->href, g_err)) return 1; ix->ref_loaded = true; }
		const BtHostRef &h = ix->href;
		P.ref.recs = h.recs.data(); P.ref.refRecOffs = h.refRecOffs.data(); P.ref.refOffs = h.refOffs.data(); P.ref.approxLen = h.approxLen.data();
		P.ref.buf = h.buf.data(); P.ref.nRefs = h.nRefs;
	}
	P.seq = in->seq; P.qual = in->qual; P.roff = in->offs; P.seeds = in->seeds;
	P.found = out->found; P.flags = out->flags; P.hits = out->hits; P.slots = out->slots; P.mm_cap = out->mm_cap; P.rec_words = BT_HIT_HDR + out->mm_cap;
	static std::vector<uint32_t> arena; arena.resize((size_t)16 << 20);
	const char *env = getenv("BT_EMU_ARENA_WORDS");
	const uint32_t words = env ? (uint32_t)atol(env) : (uint32_t)arena.size();
	const char *pz = getenv("BT_EMU_ARENA_POISON"); const bool poison = pz != NULL; const size_t poison_words = pz ? (size_t)atol(pz) : 0;
	const uint32_t nwork = in->sel ? in->nsel : (pol->paired ? in->nreads / 2 : in->nreads);
	for (uint32_t w = 0; w < nwork; w++) {
		const uint32_t r = in->sel ? in->sel[w] : w;
		BfCtx X; memset(&X, 0, sizeof X);
		X.P = &P; X.rid = r;
		for (uint32_t m = 0; m < (pol->paired ? 2u : 1u); m++) {
			const uint32_t rd = pol->paired ? 2 * r + m : r;
			X.rlenM[m] = (uint32_t)(in->offs[rd + 1] - in->offs[rd]); X.seedM[m] = in->seeds[rd];
			X.seqM[m] = in->seq + in->offs[rd]; X.qualM[m] = in->qual + in->offs[rd];
		}
		X.A = arena.data(); X.acap = words; X.atop = 1;
		if (poison) memset(arena.data(), 0xA5, (size_t)(poison_words < arena.size() ? poison_words : arena.size()) * 4);   /* the GPU arenas are never zeroed either */
		if (pol->paired) { if (P.prog.pairedV2) bf_align_pair_v2(X); else bf_align_pair(X); } else bf_align_read(X);
		if ((X.flags & BT_FLAG_STACK_OVF) && words < arena.size()) {      /* what the larger-arena passes of the product do */
			memset(&X.top, 0, sizeof X.top); X.flags = 0; X.steps = 0; X.acap = (uint32_t)arena.size(); X.atop = 1;
			if (pol->paired) { if (P.prog.pairedV2) bf_align_pair_v2(X); else bf_align_pair(X); } else bf_align_read(X);
		}
		if (X.flags & (BT_FLAG_STACK_OVF | BT_FLAG_FRAME_OVF)) X.found = 0;
		if (getenv("BT_EMU_ARENA_STATS")) fprintf(stderr, "arena %u %u\n", r, X.amax);
		out->found[r] = X.found; out->flags[r] = X.flags;
	}
	return 0;
}
static int check(const bt_index *ix, const bt_policy_t *pol) {
	if (pol->mode == 0 && pol->mms > 3) { g_err = "-v must be 0..3"; return 1; }
	if ((pol->mode == 1 || pol->mms > 0) && !ix->mirror) { g_err = "mirror index needed"; return 1; }
	return 0;
}
int bt_align_batch(bt_index_t *ix, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out, void *) { if (check(ix, pol)) return 1; return run(ix, pol, in, out); }
int bt_align_batch_device(bt_index_t *, const bt_policy_t *, const bt_read_batch_t *, bt_hit_batch_t *, void *) { g_err = "emulation shim has no device entry point"; return 1; }
int bt_context_create(bt_index_t *ix, bt_context_t **out) { bt_context *c = new bt_context(); c->ix = ix; *out = c; return 0; }
void bt_context_free(bt_context_t *cx) { delete cx; }
int bt_context_align(bt_context_t *cx, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out, void *s) { return bt_align_batch(cx->ix, pol, in, out, s); }
int bt_context_align_async(bt_context_t *cx, const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out, void *s) { return bt_align_batch(cx->ix, pol, in, out, s); }
int bt_context_align_device(bt_context_t *, const bt_policy_t *, const bt_read_batch_t *, bt_hit_batch_t *, void *) { g_err = "emulation shim has no device entry point"; return 1; }
int bt_context_sync(bt_context_t *, void *) { return 0; }
int bt_context_join(bt_context_t *, void *) { return 0; }
int bt_stats_get(bt_index_t *ix, bt_stats_t *out, int) { *out = ix->st; return 0; }
int bt_debug_lf(bt_index_t *, int, const uint32_t *, uint32_t, uint32_t *) { g_err = "not in the shim"; return 1; }
void *bt_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void bt_host_free(void *p) { free(p); }
/* index construction: the product's host code (bt_build.h) and per-element device code (bt_build_sa.cuh) over the host backend */
int bt_index_build(const char *const *fasta_paths, uint32_t n_paths, const char *out_base, int off_rate, int ftab_chars, int) {
	std::vector<std::string> files;
	for (uint32_t i = 0; i < n_paths; i++) files.push_back(fasta_paths[i]);
	BtBuildParams P; P.offRate = off_rate; P.ftabChars = ftab_chars;
	std::string err;
	if (!host_build_all(files, out_base, P, err)) { g_err = err; return 1; }
	return 0;
}
int bt_index_build_text(const uint8_t *text, uint64_t text_len, const bt_ref_record_t *recs, uint32_t n_recs, const char *const *names, uint32_t n_names,
                        const char *out_base, int off_rate, int ftab_chars, int) {
	BtRefInfo R; std::string err;
	R.text = text; R.textLen = text_len;
	for (uint32_t i = 0; i < n_recs; i++) {
		R.recs.push_back({ recs[i].off, recs[i].len, (uint8_t)(recs[i].first ? 1 : 0) });
		if (recs[i].first) R.plens.push_back(0);
		if (R.plens.empty()) { g_err = "bt_index_build_text: the first record must start a sequence"; return 1; }
		R.plens.back() += recs[i].off + recs[i].len;
	}
	for (uint32_t i = 0; i < n_names; i++) R.names.push_back(names[i] ? names[i] : "");
	BtBuildParams P; P.offRate = off_rate; P.ftabChars = ftab_chars;
	BsaHost be;
	if (!bt_build_check_ref(R, err) || !bt_build_all_on(be, R, out_base, P, err)) { g_err = err; return 1; }
	return 0;
}

/* device I/O path (f1 / f2): the product's functors (bt_io.cuh) and driver (bt_io_run.h) over a host backend; "device" memory is
 * host memory here, and the search in between is this shim's bt_align_batch */
}
struct BioHost {
	bt_context_t *cx = nullptr;
	void *alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
	void release(void *p) { free(p); }
	void h2d(void *d, const void *s, size_t bytes) { memcpy(d, s, bytes); }
	void d2h(void *d, const void *s, size_t bytes) { memcpy(d, s, bytes); }
	void zero(void *d, size_t bytes) { memset(d, 0, bytes); }
	void sync() { }
	template <class F> void each(uint64_t n, F f) { for (uint64_t i = 0; i < n; i++) f(i); }
	uint64_t count_nl(const char *t, uint64_t n) { uint64_t m = 0; for (uint64_t i = 0; i < n; i++) m += t[i] == '\n'; return m; }
	void positions_nl(const char *t, uint64_t n, uint32_t *out) { uint64_t m = 0; for (uint64_t i = 0; i < n; i++) if (t[i] == '\n') out[m++] = (uint32_t)i; }
	void excl_scan(uint32_t *a, uint64_t n) { uint32_t acc = 0; for (uint64_t i = 0; i < n; i++) { const uint32_t v = a[i]; a[i] = acc; acc += v; } }
	int align(const bt_policy_t *pol, const bt_read_batch_t *in, bt_hit_batch_t *out) { return bt_align_batch(cx->ix, pol, in, out, NULL); }
	char *out_host(size_t bytes, std::vector<char> &v) { v.resize(bytes); return v.data(); }
};
struct bt_io { BioPipe<BioHost> pipe; };
extern "C" {
int bt_io_create(bt_context_t *cx, bt_io_t **out) { bt_io *io = new bt_io(); io->pipe.be.cx = cx; io->pipe.init(); *out = io; return 0; }
void bt_io_free(bt_io_t *io) { if (!io) return; io->pipe.destroy(); delete io; }
int bt_io_parse_fastq(bt_io_t *io, const char *text, uint64_t nbytes, uint32_t global_seed, uint32_t max_reads, uint32_t *nreads, uint64_t *consumed, int *irregular) {
	if (!io->pipe.parse(text, nbytes, global_seed, max_reads, nreads, consumed, irregular)) { g_err = io->pipe.err; return 1; }
	return 0;
}
int bt_io_align_format(bt_io_t *io, const bt_policy_t *pol, const bt_io_format_t *fmt, const char **out_text, uint64_t *out_bytes, uint64_t counters[4]) {
	BioPipe<BioHost> &p = io->pipe;
	if (pol->paired || pol->all_hits || pol->sample_max || pol->khits == 0 || pol->khits > 16) { g_err = "bt_io_align_format: not provided on the device output path"; return 1; }
	if (!p.have_names || p.names_full != (fmt->full_ref != 0)) {
		std::vector<std::string> names;
		const bt_index *ix = p.be.cx->ix;
		for (uint32_t i = 0; i < ix->e[0]->raw->nPat; i++) names.push_back(i < ix->names.size() ? ix->names[i] : std::to_string(i));
		p.set_names(names, fmt->full_ref != 0);
	}
	BioFmt f;
	f.sam = fmt->sam ? 1u : 0u; f.khits = pol->khits; f.mhits = pol->mhits; f.strata = pol->strata ? 1u : 0u; f.noUnal = fmt->no_unal ? 1u : 0u;
	f.noQnameTrunc = fmt->no_qname_trunc ? 1u : 0u; f.offBase = (uint32_t)fmt->off_base; f.mapq = fmt->mapq; f.slots = pol->khits; f.recWords = 0;
	if (!p.align_format(pol, f, out_text, out_bytes, counters)) { g_err = p.err; return p.not_covered ? 2 : 1; }
	return 0;
}
int bt_counters_allreduce(void *, uint64_t *, void *) { g_err = "emulation shim has no collective"; return 1; }
}
