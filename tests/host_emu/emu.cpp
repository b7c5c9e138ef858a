/*
 * tests/host_emu/emu.cpp — TEST-ONLY logic emulation of the device state machine.
 *
 * Compiles bowtie_b200/csrc/bt_core.cuh + bt_native.cuh for the host (one lane at a time) so that
 * the kernel's control logic can be checked against the oracle on a machine without a GPU
 * (pytest -m "not gpu").  It is never linked into the product library and is not a fallback:
 * libbowtie_b200.so has no host search path at all.
 */
#define BT_HOST_EMU 1
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <vector>
#ifdef BT_EMU_PROFILE
unsigned long long bt_emu_prof[16];          /* trip counts of the rare blocks' loops (tools/pc_hist.py builds with -DBT_EMU_PROFILE) */
extern "C" unsigned long long *emu_prof(void) { return bt_emu_prof; }
#endif
#include "../../bowtie_b200/csrc/bt_native.cuh"
#include "../../bowtie_b200/csrc/bt_ctxq.cuh"
extern "C" {
#include "../../oracle/bt_oracle.h"
}

struct EmuIndex { std::vector<uint4> blocks; bto_index *raw; BtDevIndex dev; };

static EmuIndex *emu_load(const char *base, int mirror) {
	char err[256];
	bto_index *ix = bto_index_load(base, mirror, err, sizeof err);
	if (!ix) return NULL;
	EmuIndex *e = new EmuIndex();
	e->raw = ix;
	BtNativeIndex n;
	n.ebwt = ix->ebwt; n.len = ix->len; n.zOff = ix->zOff; n.zEbwtByteOff = ix->zEbwtByteOff; n.zEbwtBpOff = (uint32_t)ix->zEbwtBpOff;
	memcpy(n.fchr, ix->fchr, sizeof n.fchr);
	uint32_t nblocks = (ix->len >> 6) + 1;
	e->blocks.resize(2 * (size_t)nblocks);
	for (uint32_t k = 0; k < nblocks; k++) bt_relayout_block(n, k, &e->blocks[2 * (size_t)k]);
	BtDevIndex &d = e->dev;
	d.blocks = e->blocks.data(); d.offs = ix->offs; d.ftab = ix->ftab; d.eftab = ix->eftab; d.rstarts = ix->rstarts; d.plen = ix->plen;
	d.len = ix->len; d.zOff = ix->zOff; d.nFrag = ix->nFrag; d.nPat = ix->nPat; d.offMask = ix->offMask;
	d.offRate = ix->offRate; d.ftabChars = ix->ftabChars; memcpy(d.fchr, ix->fchr, sizeof d.fchr); d.fw = (uint32_t)ix->fw;
	return e;
}

extern "C" {

void *emu_index_load(const char *base, int mirror) { return emu_load(base, mirror); }
void emu_index_free(void *p) { EmuIndex *e = (EmuIndex *)p; if (!e) return; bto_index_free(e->raw); delete e; }

/* LF via the re-laid-out blocks, for unit tests against the oracle's side arithmetic */
uint32_t emu_lf(void *p, uint32_t row, int c) { EmuIndex *e = (EmuIndex *)p; BtBlock b = bt_load_block(e->dev, row); return bt_lf(e->dev, b, row, (uint32_t)c); }
void emu_lf_ex(void *p, uint32_t row, uint32_t out[4]) { EmuIndex *e = (EmuIndex *)p; BtBlock b = bt_load_block(e->dev, row); bt_lf_ex(e->dev, b, row, out); }
int emu_row_l(void *p, uint32_t row) { EmuIndex *e = (EmuIndex *)p; BtBlock b = bt_load_block(e->dev, row); return (int)bt_row_l(b, row); }

int emu_align(void *fwp, void *bwp, const BtPolicy *pol, uint32_t nreads, const uint8_t *seq, const uint8_t *qual,
              const uint64_t *roff, const uint32_t *seeds, uint32_t *found, uint32_t *flags, uint32_t *hits,
              uint32_t slots, uint32_t mm_cap, uint32_t R, uint32_t FCAP, uint32_t PCAP, unsigned long long *stats, uint32_t *iters_per_read) {
	BtKParams P; memset(&P, 0, sizeof P);
	P.ix[0] = ((EmuIndex *)fwp)->dev;
	if (bwp) P.ix[1] = ((EmuIndex *)bwp)->dev;
	P.pol = *pol;
	P.seq = seq; P.qual = qual; P.roff = roff; P.seeds = seeds; P.nwork = nreads;
	P.found = found; P.flags = flags; P.hits = hits; P.slots = slots; P.mm_cap = mm_cap; P.rec_words = BT_HIT_HDR + mm_cap;
	{ uint32_t ml = 1; for (uint32_t r = 0; r < nreads; r++) { const uint32_t l = (uint32_t)(roff[r + 1] - roff[r]); if (l > ml) ml = l; } P.mask_rows = (ml + 255) >> 8; }
	R += (FCAP + 1) * P.mask_rows;                            /* every frame also holds its live-position mask (as enqueue_align sizes it) */
	std::vector<uint4> rows(2 * (size_t)R); std::vector<uint8_t> elims(R); std::vector<BtFrame> frames(FCAP); std::vector<uint64_t> parts(PCAP);
	P.R = R; P.FCAP = FCAP; P.PCAP = PCAP;
	BtScratch S = { rows.data(), elims.data(), frames.data(), parts.data() };
	bt_build_prog(pol->mode, pol->mms, pol->nofw, pol->norc, P.prog);
	BtLane L; BtLaneCold LK; memset(&L, 0, sizeof L); memset(&LK, 0, sizeof LK); L.K = &LK;
	std::vector<uint8_t> stage;
	/* BT_EMU_PC_HIST=1: how many transitions of each kind the batch took (where a lane's iterations go; development aid) */
	static unsigned long long hist_store[40];
	unsigned long long *pc_hist = getenv("BT_EMU_PC_HIST") ? hist_store : NULL, *lfk_hist = hist_store + 32;
	if (pc_hist) memset(hist_store, 0, sizeof hist_store);
	for (uint32_t r = 0; r < nreads; r++) {
		bt_begin_read(L, P, r);
		/* the lane's writable copy of the read (shared memory on the device) */
		stage.assign(2 * (size_t)L.rlen + 2, 0);
		memcpy(stage.data(), seq + roff[r], L.rlen);
		memcpy(stage.data() + L.rlen, qual + roff[r], L.rlen);
		L.rseq = stage.data(); L.rqual = stage.data() + L.rlen;
		L.K->hasN = memchr(stage.data(), 4, L.rlen) != NULL;
		unsigned long long guard = 0; uint32_t it0 = L.s_iter;
		while (L.pc != PC_FINISH_READ) {
			if (pc_hist) { pc_hist[L.pc < 32 ? L.pc : 31]++; if (L.pc == PC_LF) lfk_hist[L.lfk & 7]++; }
			if (BT_IS_FAST(L.pc)) bt_fast_iter(L, P, S); else bt_rare_iter(L, P, S);
			if (++guard > (1ull << 34)) return 1;
		}
		bt_finish_read(L, P);
		if (iters_per_read) iters_per_read[r] = L.s_iter - it0;
	}
	if (pc_hist) {
		static const char *names[] = { "LF", "CHASE", "PHASE", "BT_BEGIN", "FRAME_ENTER", "POS", "BTLOOP", "CHILD_RET", "POS_END", "FRAME_RET", "REPORT", "REPORT_ROW", "RESOLVE", "REPORT_RET", "BT_END" };
		unsigned long long tot = 0;
		for (int i = 0; i < 15; i++) tot += pc_hist[i];
		fprintf(stderr, "transitions per read: %.1f\n", (double)tot / nreads);
		for (int i = 0; i < 15; i++) fprintf(stderr, "  %-12s %8.2f per read  %5.1f %%\n", names[i], (double)pc_hist[i] / nreads, 100.0 * pc_hist[i] / tot);
		static const char *lk[] = { "EX", "ONE", "PAIR", "FCHR", "NONE" };
		for (int i = 0; i < 5; i++) fprintf(stderr, "  LF kind %-5s %8.2f per read\n", lk[i], (double)lfk_hist[i] / nreads);
	}
	stats[0] = L.s_lfex; stats[1] = L.s_lf; stats[2] = L.s_chase; stats[3] = L.K->s_ftab; stats[4] = L.K->s_offs; stats[5] = L.K->s_bt; stats[6] = L.s_iter; stats[7] = L.s_blk;
	return 0;
}

/* The same batch under the kernels' time slicing (bt_ctxq.cuh, checkpoint slots): a read that exceeds `budget0` transitions — or fills
 * the main pass's seedling list — is suspended into a slot (bt_slot_save_new), the lane and its scratch are poisoned, and the read is
 * resumed from the slot (bt_slot_resume) with the slot's capacities and `growth` times the budget, again and again until it ends.
 * Results must equal emu_align's; *nsusp counts the suspensions. */
int emu_align_sliced(void *fwp, void *bwp, const BtPolicy *pol, uint32_t nreads, const uint8_t *seq, const uint8_t *qual,
                     const uint64_t *roff, const uint32_t *seeds, uint32_t *found, uint32_t *flags, uint32_t *hits,
                     uint32_t slots, uint32_t mm_cap, uint32_t R, uint32_t FCAP, uint32_t PCAP, uint32_t slotFCAP, uint32_t slotPCAP,
                     uint32_t budget0, uint32_t growth, unsigned long long *stats, unsigned long long *nsusp) {
	BtKParams P; memset(&P, 0, sizeof P);
	P.ix[0] = ((EmuIndex *)fwp)->dev;
	if (bwp) P.ix[1] = ((EmuIndex *)bwp)->dev;
	P.pol = *pol;
	P.seq = seq; P.qual = qual; P.roff = roff; P.seeds = seeds; P.nwork = nreads;
	P.found = found; P.flags = flags; P.hits = hits; P.slots = slots; P.mm_cap = mm_cap; P.rec_words = BT_HIT_HDR + mm_cap;
	{ uint32_t ml = 1; for (uint32_t r = 0; r < nreads; r++) { const uint32_t l = (uint32_t)(roff[r + 1] - roff[r]); if (l > ml) ml = l; } P.mask_rows = (ml + 255) >> 8; }
	R += ((FCAP > slotFCAP ? FCAP : slotFCAP) + 1) * P.mask_rows;
	std::vector<uint4> rows(2 * (size_t)R), srows(2 * (size_t)R); std::vector<uint8_t> elims(R), selims(R);
	std::vector<BtFrame> frames(FCAP), sframes(slotFCAP); std::vector<uint64_t> parts(PCAP), sparts(slotPCAP);
	std::vector<uint32_t> sctx(BT_CTX_WORDS);
	uint32_t maxlen = 1;
	for (uint32_t r = 0; r < nreads; r++) { const uint32_t l = (uint32_t)(roff[r + 1] - roff[r]); if (l > maxlen) maxlen = l; }
	std::vector<uint8_t> sstage(2 * (size_t)maxlen + 2);
	P.slot_ctx = sctx.data(); P.slot_rows = srows.data(); P.slot_elims = selims.data(); P.slot_frames = sframes.data(); P.slot_partials = sparts.data();
	P.slot_stage = sstage.data(); P.nslot = 1; P.slot_R = R; P.slot_FCAP = slotFCAP; P.slot_PCAP = slotPCAP; P.slot_stage_len = maxlen;
	bt_build_prog(pol->mode, pol->mms, pol->nofw, pol->norc, P.prog);
	BtLane L; BtLaneCold LK; memset(&L, 0, sizeof L); memset(&LK, 0, sizeof LK); L.K = &LK;
	std::vector<uint8_t> stage;
	*nsusp = 0;
	for (uint32_t r = 0; r < nreads; r++) {
		P.R = R; P.FCAP = FCAP; P.PCAP = PCAP; P.resume = 0;
		BtScratch S = { rows.data(), elims.data(), frames.data(), parts.data() };
		uint32_t budget = budget0;
		bt_begin_read(L, P, r);
		stage.assign(2 * (size_t)L.rlen + 2, 0);
		memcpy(stage.data(), seq + roff[r], L.rlen);
		memcpy(stage.data() + L.rlen, qual + roff[r], L.rlen);
		L.rseq = stage.data(); L.rqual = stage.data() + L.rlen;
		L.K->hasN = memchr(stage.data(), 4, L.rlen) != NULL;
		unsigned long long guard = 0;
		while (L.pc != PC_FINISH_READ) {
			if (BT_IS_FAST(L.pc)) bt_fast_iter(L, P, S); else bt_rare_iter(L, P, S, budget);
			if (L.flags & BT_FLAG_PREEMPT) {
				L.flags &= ~BT_FLAG_PREEMPT;
				if (!P.resume) bt_slot_save_new(L, P, S, 0); else bt_ctx_store(L, P.slot_ctx, 1, 0);
				(*nsusp)++;
				/* nothing of the lane or of the main pass's scratch survives, except the lane's operation counters */
				uint32_t keep[8] = { L.s_lfex, L.s_lf, L.s_chase, L.K->s_ftab, L.K->s_offs, L.K->s_bt, L.s_iter, L.s_blk };
				memset(&L, 0xCD, sizeof L); memset(&LK, 0xCD, sizeof LK); L.K = &LK;
				L.s_lfex = keep[0]; L.s_lf = keep[1]; L.s_chase = keep[2]; L.K->s_ftab = keep[3]; L.K->s_offs = keep[4]; L.K->s_bt = keep[5]; L.s_iter = keep[6]; L.s_blk = keep[7];
				memset(rows.data(), 0xCD, rows.size() * sizeof(uint4)); memset(elims.data(), 0xCD, elims.size());
				memset(frames.data(), 0xCD, frames.size() * sizeof(BtFrame)); memset(parts.data(), 0xCD, parts.size() * 8);
				memset(stage.data(), 0xCD, stage.size());
				P.resume = 1; P.R = P.slot_R; P.FCAP = slotFCAP; P.PCAP = slotPCAP;
				budget = growth ? (budget > 0x7fffffffu / growth ? 0u : budget * growth) : 0u;
				bt_slot_resume(L, P, S, 0);
			}
			if (++guard > (1ull << 34)) return 1;
		}
		bt_finish_read(L, P);
	}
	stats[0] = L.s_lfex; stats[1] = L.s_lf; stats[2] = L.s_chase; stats[3] = L.K->s_ftab; stats[4] = L.K->s_offs; stats[5] = L.K->s_bt; stats[6] = L.s_iter; stats[7] = L.s_blk;
	return 0;
}

/* Warp-level replay of bt_search_kernel's scheduling (development aid, tools/warp_sim.py): `nwarps` x 32 lanes pull reads from one
 * cursor and take fast / deferred-rare transitions under the kernel's rule; what comes out is how full the warps are when they
 * execute each kind of code — the quantity ncu reports as "threads active per instruction" — for a given period / threshold /
 * budget, without a GPU.  Code paths a warp executes in one iteration: one per LF kind present among its fast lanes, one for the
 * chase step, one per distinct rare state present when the rare pass runs.  out[]: see the tool. */
int emu_warp_sim(void *fwp, void *bwp, const BtPolicy *pol, uint32_t nreads, const uint8_t *seq, const uint8_t *qual, const uint64_t *roff,
                 const uint32_t *seeds, uint32_t nwarps, uint32_t rare_period, uint32_t rare_thresh, uint32_t budget, uint32_t R, uint32_t FCAP, uint32_t PCAP,
                 double *out) {
	BtKParams P; memset(&P, 0, sizeof P);
	P.ix[0] = ((EmuIndex *)fwp)->dev;
	if (bwp) P.ix[1] = ((EmuIndex *)bwp)->dev;
	P.pol = *pol;
	P.seq = seq; P.qual = qual; P.roff = roff; P.seeds = seeds; P.nwork = nreads;
	const uint32_t slots = 1, mm_cap = 8;
	std::vector<uint32_t> found(nreads), flags(nreads), hits((size_t)nreads * slots * (BT_HIT_HDR + mm_cap));
	P.found = found.data(); P.flags = flags.data(); P.hits = hits.data(); P.slots = slots; P.mm_cap = mm_cap; P.rec_words = BT_HIT_HDR + mm_cap;
	{ uint32_t ml = 1; for (uint32_t r = 0; r < nreads; r++) { const uint32_t l = (uint32_t)(roff[r + 1] - roff[r]); if (l > ml) ml = l; } P.mask_rows = (ml + 255) >> 8; }
	R += (FCAP + 1) * P.mask_rows;
	P.R = R; P.FCAP = FCAP; P.PCAP = PCAP; P.rare_period = rare_period ? rare_period : 1; P.rare_thresh = rare_thresh; P.budget = budget;
	bt_build_prog(pol->mode, pol->mms, pol->nofw, pol->norc, P.prog);
	const uint32_t nl = nwarps * 32;
	std::vector<BtLane> lanes(nl); std::vector<BtLaneCold> colds(nl);
	std::vector<std::vector<uint4>> rows(nl); std::vector<std::vector<uint8_t>> elims(nl), stage(nl); std::vector<std::vector<BtFrame>> frames(nl); std::vector<std::vector<uint64_t>> parts(nl);
	std::vector<BtScratch> S(nl);
	for (uint32_t i = 0; i < nl; i++) {
		rows[i].resize(2 * (size_t)R); elims[i].resize(R); frames[i].resize(FCAP); parts[i].resize(PCAP);
		S[i] = BtScratch{ rows[i].data(), elims[i].data(), frames[i].data(), parts[i].data() };
		memset(&lanes[i], 0, sizeof(BtLane)); memset(&colds[i], 0, sizeof(BtLaneCold)); lanes[i].K = &colds[i]; lanes[i].pc = PC_NEXT_READ; lanes[i].K->hasN = 1;
	}
	const int rule = getenv("BT_SIM_RULE") ? atoi(getenv("BT_SIM_RULE")) : 0;
	const int unify = getenv("BT_SIM_UNIFIED") ? atoi(getenv("BT_SIM_UNIFIED")) : 0;   /* 1: one code path for the three LF kinds; 2: the chase step shares it too */
	uint64_t cursor = 0;
	/* accumulators */
	double warpIters = 0, fastPaths = 0, fastLaneSum = 0, rarePasses = 0, rarePaths = 0, rareLaneSum = 0, idleLaneIters = 0, liveLaneIters = 0, budgeted = 0;
	double fastHist[33]; memset(fastHist, 0, sizeof fastHist);
	for (uint32_t w = 0; w < nwarps; w++) {
		BtLane *L = &lanes[w * 32];
		for (uint32_t it = 0;; it++) {
			uint32_t nfast = 0, nrare = 0, kinds = 0; bool wasFast[32];
			for (int l = 0; l < 32; l++) {
				const uint32_t pc = L[l].pc;
				wasFast[l] = BT_IS_FAST(pc);
				if (wasFast[l]) { nfast++; kinds |= (pc == PC_CHASE) ? (unify >= 2 ? 1u : 16u) : (unify ? 1u : (1u << (L[l].lfk & 3))); }
				else if (pc != PC_EXIT) nrare++;
			}
			if (nfast + nrare == 0) break;
			warpIters++; liveLaneIters += nfast + nrare;
			bool run_rare = (nfast == 0) || (nrare >= P.rare_thresh) || ((it % P.rare_period) == 0);
			if (rule == 1) run_rare = run_rare || nrare > nfast;               /* more lanes waiting than working */
			if (rule == 2) run_rare = (nfast == 0) || nrare >= P.rare_thresh || nrare > nfast;
			if (run_rare && nrare) {
				uint32_t pcs = 0;
				for (int l = 0; l < 32; l++) {
					BtLane &X = L[l];
					if (X.pc == PC_FINISH_READ) { bt_finish_read(X, P); X.pc = PC_NEXT_READ; }
					if (X.pc == PC_NEXT_READ) {
						if (cursor < nreads) {
							const uint32_t rid = (uint32_t)cursor++;
							bt_begin_read(X, P, rid);
							std::vector<uint8_t> &st = stage[w * 32 + l];
							st.assign(2 * (size_t)X.rlen + 2, 0);
							memcpy(st.data(), seq + roff[rid], X.rlen); memcpy(st.data() + X.rlen, qual + roff[rid], X.rlen);
							X.rseq = st.data(); X.rqual = st.data() + X.rlen; X.K->hasN = memchr(st.data(), 4, X.rlen) != NULL;
						} else X.pc = PC_EXIT;
						pcs |= 1u << 31;
					}
					if (BT_IS_RARE_STEP(X.pc)) { pcs |= 1u << (X.pc & 31); const uint32_t f0 = X.flags; bt_rare_iter(X, P, S[w * 32 + l]); if ((X.flags & ~f0) & BT_FLAG_BUDGET) budgeted++; }
				}
				rarePasses++; rarePaths += __builtin_popcount(pcs); rareLaneSum += nrare;
			}
			if (nfast) {                                                  /* `fast` is sampled before the rare pass, as in the kernel */
				for (int l = 0; l < 32; l++) if (wasFast[l]) bt_fast_iter(L[l], P, S[w * 32 + l]);
				fastPaths += __builtin_popcount(kinds); fastLaneSum += nfast; fastHist[nfast]++;
			}
			idleLaneIters += run_rare ? 0 : nrare;
		}
	}
	out[0] = warpIters; out[1] = fastPaths; out[2] = fastLaneSum; out[3] = rarePasses; out[4] = rarePaths; out[5] = rareLaneSum; out[6] = idleLaneIters; out[7] = liveLaneIters; out[8] = budgeted;
	for (int i = 0; i <= 32; i++) out[9 + i] = fastHist[i];
	return 0;
}

}
